// MDR decoder forward (lib/models/MDR.py:124-170): vertex/joint embedding, 3 x (cross-attention block +
// unbiased-std LayerNorm + 431x431 self-attention), MDR head, dense upsample conv + template.
// fp32 parity path: dense products through gemm_f32; attention, head and embedding hand-fused here.
#include "common.cuh"

namespace gator {
namespace {

constexpr int E = 64;                 // MDR embed dim (MDR.py:74)
constexpr int V = GATOR_V_COARSE;     // 431
constexpr int VF = GATOR_V_FULL;      // 6890
constexpr int DK = 32;                // head dim (2 heads)
constexpr int MAXJ = 32;
constexpr int HEADN = 28;
constexpr int UPK = GATOR_UP_K;       // 1296

// ---------------------------------------------------------------------------------------------
// Embedding (MDR.py:127-137 with the concat of GATOR.py:19 folded in).  One CTA per sample.
//   jf[b,j,:]  = W_j[:, :5] . [pose2d, pose3d/1000]      (the feat part is added by the GEMM that follows)
//   x[b,v,:]   = VF_CONST[v,:] + W_v[:,3:6] . pose3d[b, vj[v]]/1000
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mdr_embed_kernel(const float* __restrict__ pose2d, const float* __restrict__ pose3d, const float* __restrict__ wpose,
                 const float* __restrict__ vconst, const float* __restrict__ w3, const int* __restrict__ vj,
                 float* __restrict__ jf, float* __restrict__ x, int J, int metres) {
  // jf == nullptr: vertices only; x == nullptr: joints only
  __shared__ float p5[MAXJ][5];
  __shared__ float sw3[E * 3];
  __shared__ float swp[E * 5];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < J) {
    p5[tid][0] = pose2d[((size_t)b * J + tid) * 2 + 0];
    p5[tid][1] = pose2d[((size_t)b * J + tid) * 2 + 1];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const float pv = pose3d[((size_t)b * J + tid) * 3 + t];
      p5[tid][2 + t] = metres ? pv : pv / 1000.0f;
    }
  }
  for (int i = tid; i < E * 3; i += 256) sw3[i] = w3[i];
  for (int i = tid; i < E * 5; i += 256) swp[i] = wpose[i];
  __syncthreads();
  for (int i = tid; jf && i < J * E; i += 256) {
    const int j = i / E, n = i - j * E;
    float a = 0.f;
#pragma unroll
    for (int q = 0; q < 5; ++q) a = fmaf(swp[n * 5 + q], p5[j][q], a);
    jf[((size_t)b * J + j) * E + n] = a;
  }
  if (!x) return;
  float* xb = x + (size_t)b * V * E;
  for (int i = tid; i < V * (E / 4); i += 256) {
    const int v = i / (E / 4), n = (i - v * (E / 4)) * 4;
    const int j = __ldg(vj + v);
    const float px = p5[j][2], py = p5[j][3], pz = p5[j][4];
    float4 c = __ldg(reinterpret_cast<const float4*>(vconst + (size_t)v * E + n));
    float r[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float* w = sw3 + (n + u) * 3;
      r[u] += fmaf(w[0], px, fmaf(w[1], py, w[2] * pz));
    }
    *reinterpret_cast<float4*>(xb + (size_t)v * E + n) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// Cross attention core (MDR.py:34-46): 431 vertex queries x J joint keys, 2 heads of 32.
// One CTA per sample; K/V of the sample in shared memory; one thread per (head, vertex).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mdr_cross_attn_kernel(const float* __restrict__ q, const float* __restrict__ kv, float* __restrict__ out, int J) {
  __shared__ __align__(16) float sk[MAXJ][E];
  __shared__ __align__(16) float sv[MAXJ][E];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < J * 2 * E; i += 256) {
    const int j = i / (2 * E), c = i - j * 2 * E;
    const float val = kv[((size_t)b * J + j) * 2 * E + c];
    if (c < E) sk[j][c] = val; else sv[j][c - E] = val;
  }
  __syncthreads();
  const float scale = 0.17677669529663687f;   // 32 ** -0.5
  for (int w = tid; w < 2 * V; w += 256) {
    const int h = w / V, v = w - h * V;
    const float* qr = q + ((size_t)b * V + v) * E + h * DK;
    float qq[DK];
#pragma unroll
    for (int d = 0; d < DK; d += 4) {
      float4 t = *reinterpret_cast<const float4*>(qr + d);
      qq[d] = t.x; qq[d + 1] = t.y; qq[d + 2] = t.z; qq[d + 3] = t.w;
    }
    float s[MAXJ];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      if (j < J) {
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < DK; ++d) a = fmaf(qq[d], sk[j][h * DK + d], a);
        s[j] = a * scale;
        m = fmaxf(m, s[j]);
      }
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
      if (j < J) { s[j] = expf(s[j] - m); l += s[j]; }
    const float inv = 1.0f / l;
    float o[DK];
#pragma unroll
    for (int d = 0; d < DK; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      if (j < J) {
        const float p = s[j] * inv;
#pragma unroll
        for (int d = 0; d < DK; ++d) o[d] = fmaf(p, sv[j][h * DK + d], o[d]);
      }
    }
    float* orow = out + ((size_t)b * V + v) * E + h * DK;
#pragma unroll
    for (int d = 0; d < DK; d += 4) *reinterpret_cast<float4*>(orow + d) = make_float4(o[d], o[d + 1], o[d + 2], o[d + 3]);
  }
}

// ---------------------------------------------------------------------------------------------
// Self attention core (vanilla_transformer_encoder.py:36-46): softmax(q k^T / sqrt(32)) v over the 431
// coarse vertices, 2 heads.  fp32 flash-style: one CTA per (sample, head) keeps that head's K and V
// (2 x 432 x 32 fp32 = 108 KB) in shared memory; each thread owns QPT query rows and streams the keys in
// chunks of 8 with an online softmax (exp2 domain), so the 431x431 score matrix never exists.
// ---------------------------------------------------------------------------------------------
constexpr int SA_THREADS = 224;
constexpr int SA_QPT = 2;
constexpr int SA_KC = 8;
constexpr int SA_VPAD = 432;

__global__ void __launch_bounds__(SA_THREADS)
mdr_self_attn_kernel(const float* __restrict__ qkv, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* sk = smem;                    // [432][32]
  float* sv = smem + SA_VPAD * DK;     // [432][32]
  const int b = blockIdx.x >> 1, h = blockIdx.x & 1, tid = threadIdx.x;
  const float* base = qkv + (size_t)b * V * 3 * E;
  for (int i = tid; i < SA_VPAD * (DK / 4); i += SA_THREADS) {
    const int r = i / (DK / 4), c = (i - r * (DK / 4)) * 4;
    float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
    if (r < V) {
      kk = *reinterpret_cast<const float4*>(base + (size_t)r * 3 * E + E + h * DK + c);
      vv = *reinterpret_cast<const float4*>(base + (size_t)r * 3 * E + 2 * E + h * DK + c);
    }
    *reinterpret_cast<float4*>(sk + r * DK + c) = kk;
    *reinterpret_cast<float4*>(sv + r * DK + c) = vv;
  }
  __syncthreads();
  // 1/sqrt(32) * log2(e): scores live in the exp2 domain
  const float qscale = 0.17677669529663687f * 1.4426950408889634f;
  float qq[SA_QPT][DK], o[SA_QPT][DK], m[SA_QPT], l[SA_QPT];
  int row[SA_QPT];
#pragma unroll
  for (int r = 0; r < SA_QPT; ++r) {
    row[r] = tid + r * SA_THREADS;
    const int rr = row[r] < V ? row[r] : V - 1;     // clamp: duplicate work, store is masked
    const float* qr = base + (size_t)rr * 3 * E + h * DK;
#pragma unroll
    for (int d = 0; d < DK; d += 4) {
      float4 t = *reinterpret_cast<const float4*>(qr + d);
      qq[r][d] = t.x * qscale; qq[r][d + 1] = t.y * qscale; qq[r][d + 2] = t.z * qscale; qq[r][d + 3] = t.w * qscale;
    }
#pragma unroll
    for (int d = 0; d < DK; ++d) o[r][d] = 0.f;
    m[r] = -INFINITY;
    l[r] = 0.f;
  }
  for (int k0 = 0; k0 < SA_VPAD; k0 += SA_KC) {
    float s[SA_QPT][SA_KC];
#pragma unroll
    for (int kk = 0; kk < SA_KC; ++kk) {
      const float4* kr = reinterpret_cast<const float4*>(sk + (k0 + kk) * DK);
      float a[SA_QPT];
#pragma unroll
      for (int r = 0; r < SA_QPT; ++r) a[r] = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < DK / 4; ++d4) {
        const float4 kv4 = kr[d4];
#pragma unroll
        for (int r = 0; r < SA_QPT; ++r) {
          a[r] = fmaf(qq[r][d4 * 4 + 0], kv4.x, a[r]);
          a[r] = fmaf(qq[r][d4 * 4 + 1], kv4.y, a[r]);
          a[r] = fmaf(qq[r][d4 * 4 + 2], kv4.z, a[r]);
          a[r] = fmaf(qq[r][d4 * 4 + 3], kv4.w, a[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < SA_QPT; ++r) s[r][kk] = (k0 + kk < V) ? a[r] : -INFINITY;
    }
#pragma unroll
    for (int r = 0; r < SA_QPT; ++r) {
      float cm = s[r][0];
#pragma unroll
      for (int kk = 1; kk < SA_KC; ++kk) cm = fmaxf(cm, s[r][kk]);
      const float mn = fmaxf(m[r], cm);
      const float corr = exp2f(m[r] - mn);     // m = -inf on the first chunk -> 0
      m[r] = mn;
      float ps = 0.f;
#pragma unroll
      for (int kk = 0; kk < SA_KC; ++kk) { s[r][kk] = exp2f(s[r][kk] - mn); ps += s[r][kk]; }
      l[r] = fmaf(l[r], corr, ps);
#pragma unroll
      for (int d = 0; d < DK; ++d) o[r][d] *= corr;
    }
#pragma unroll
    for (int kk = 0; kk < SA_KC; ++kk) {
      const float4* vr = reinterpret_cast<const float4*>(sv + (k0 + kk) * DK);
#pragma unroll
      for (int d4 = 0; d4 < DK / 4; ++d4) {
        const float4 vv = vr[d4];
#pragma unroll
        for (int r = 0; r < SA_QPT; ++r) {
          const float p = s[r][kk];
          o[r][d4 * 4 + 0] = fmaf(p, vv.x, o[r][d4 * 4 + 0]);
          o[r][d4 * 4 + 1] = fmaf(p, vv.y, o[r][d4 * 4 + 1]);
          o[r][d4 * 4 + 2] = fmaf(p, vv.z, o[r][d4 * 4 + 2]);
          o[r][d4 * 4 + 3] = fmaf(p, vv.w, o[r][d4 * 4 + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < SA_QPT; ++r) {
    if (row[r] >= V) continue;
    const float inv = 1.0f / l[r];
    float* orow = out + ((size_t)b * V + row[r]) * E + h * DK;
#pragma unroll
    for (int d = 0; d < DK; d += 4)
      *reinterpret_cast<float4*>(orow + d) = make_float4(o[r][d] * inv, o[r][d + 1] * inv, o[r][d + 2] * inv, o[r][d + 3] * inv);
  }
}

// ---------------------------------------------------------------------------------------------
// MDR head (MDR.py:156-166) given hd = [motion_linear(23) | bias_linear(3) | scale_linear(1) | 0](x):
// bias branch norm + GELU + Conv1d(431->20,k3,p1), softmax(A) @ B, alpha scaling, + C; then the im2col
// rows of upsample_conv's input (MDR.py:167).  One CTA per sample.
// ---------------------------------------------------------------------------------------------
constexpr int HD_NS = 4;   // samples per CTA: the 103 KB of Conv1d weights are read from L2 once per CTA and reused from
                           // registers for its samples (one sample per CTA moved 423 MB through L2 per 4096 samples: 176 us)
__global__ void __launch_bounds__(256)
mdr_head_kernel(const float* __restrict__ hd, const float* __restrict__ nscale, const float* __restrict__ nshift,
                const float* __restrict__ cw, const float* __restrict__ cb, int alpha, float* __restrict__ coarse_out,
                float* __restrict__ a3, int nb) {
  __shared__ float gB[HD_NS][V + 2][3];
  __shared__ float matB[HD_NS][20][3];
  __shared__ float sc[HD_NS][V][3];
  const int b0 = blockIdx.x * HD_NS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ns = min(HD_NS, nb - b0);
  for (int i = tid; i < ns * V; i += 256) {
    const int s = i / V, v = i - s * V;
    const float* hb = hd + ((size_t)(b0 + s) * V + v) * HEADN;
    float y[3];
    const float x0 = hb[23], x1 = hb[24], x2 = hb[25];
    if (alpha) {   // nn.LayerNorm(3)
      const float mean = (x0 + x1 + x2) / 3.0f;
      const float d0 = x0 - mean, d1 = x1 - mean, d2 = x2 - mean;
      const float rstd = rsqrtf((d0 * d0 + d1 * d1 + d2 * d2) / 3.0f + 1e-5f);
      y[0] = d0 * rstd * nscale[0] + nshift[0];
      y[1] = d1 * rstd * nscale[1] + nshift[1];
      y[2] = d2 * rstd * nscale[2] + nshift[2];
    } else {       // eval BatchNorm1d(431): channel = vertex
      const float sc_ = nscale[v], t = nshift[v];
      y[0] = fmaf(x0, sc_, t); y[1] = fmaf(x1, sc_, t); y[2] = fmaf(x2, sc_, t);
    }
#pragma unroll
    for (int t = 0; t < 3; ++t) gB[s][v][t] = gelu_erf(y[t]);
  }
  __syncthreads();
  for (int o = warp; o < 20; o += 8) {
    float p0[HD_NS], p1[HD_NS], p2[HD_NS];
#pragma unroll
    for (int s = 0; s < HD_NS; ++s) p0[s] = p1[s] = p2[s] = 0.f;
    const float* w = cw + (size_t)o * V * 3;
    for (int c = lane; c < V; c += 32) {
      const float w0 = __ldg(w + c * 3), w1 = __ldg(w + c * 3 + 1), w2 = __ldg(w + c * 3 + 2);
#pragma unroll
      for (int s = 0; s < HD_NS; ++s) {
        if (s < ns) {
          const float g0 = gB[s][c][0], g1 = gB[s][c][1], g2 = gB[s][c][2];
          p0[s] = fmaf(w1, g0, fmaf(w2, g1, p0[s]));
          p1[s] = fmaf(w0, g0, fmaf(w1, g1, fmaf(w2, g2, p1[s])));
          p2[s] = fmaf(w0, g1, fmaf(w1, g2, p2[s]));
        }
      }
    }
    const float bb = cb[o];
#pragma unroll
    for (int s = 0; s < HD_NS; ++s) {
      const float q0 = warp_sum(p0[s]), q1 = warp_sum(p1[s]), q2 = warp_sum(p2[s]);
      if (lane == 0 && s < ns) { matB[s][o][0] = q0 + bb; matB[s][o][1] = q1 + bb; matB[s][o][2] = q2 + bb; }
    }
  }
  __syncthreads();
  for (int i = tid; i < ns * V; i += 256) {
    const int s = i / V, v = i - s * V;
    const float* r = hd + ((size_t)(b0 + s) * V + v) * HEADN;
    float a[20];
    float m = -INFINITY;
#pragma unroll
    for (int o = 0; o < 20; ++o) { a[o] = r[o]; m = fmaxf(m, a[o]); }
    float l = 0.f;
#pragma unroll
    for (int o = 0; o < 20; ++o) { a[o] = expf(a[o] - m); l += a[o]; }
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int o = 0; o < 20; ++o) {
      const float p = a[o] / l;
      c0 = fmaf(p, matB[s][o][0], c0); c1 = fmaf(p, matB[s][o][1], c1); c2 = fmaf(p, matB[s][o][2], c2);
    }
    const float al = alpha ? powf(1.1f, r[26]) : 1.0f;
    sc[s][v][0] = al * c0 + r[20]; sc[s][v][1] = al * c1 + r[21]; sc[s][v][2] = al * c2 + r[22];
  }
  __syncthreads();
  for (int s = 0; s < ns; ++s) {
    const int b = b0 + s;
    if (coarse_out) {
      float* co = coarse_out + (size_t)b * V * 3;
      for (int i = tid; i < V * 3; i += 256) co[i] = (&sc[s][0][0])[i];
    }
    float* ab = a3 + (size_t)b * 3 * UPK;
    for (int i = tid; i < 3 * UPK; i += 256) {
      const int t = i / UPK, r = i - t * UPK;
      float val = 0.f;
      if (r < V * 3) {
        const int c = r / 3, k = r - c * 3, tt = t + k - 1;
        if (tt >= 0 && tt < 3) val = sc[s][c][tt];
      }
      ab[i] = val;
    }
  }
}

const char* kGlobalNames[MDR_NUM_GLOBAL] = {
    "JF_WFEAT", "JF_WPOSE", "JF_BIASROWS", "VF_CONST", "VF_W3", "VJ", "HEAD_W", "HEAD_B",
    "BNORM_SCALE", "BNORM_SHIFT", "BCONV_W", "BCONV_B", "UP_W", "UP_BIAST", "CHAIN_FINAL", "UP_W_WIDE"};
const char* kLayerNames[MDRL_NUM] = {
    "N1_W", "N1_B", "WQ", "WKV", "PROJ_W", "PROJ_B", "N2_W", "N2_B", "FC1_W", "FC1_B", "FC2_W", "FC2_B",
    "CLN_A", "CLN_B", "SQKV_W", "SQKV_B", "SO_W", "SO_B", "CHAIN"};

constexpr int kDefaultChunk = 148;   // 296 (sample, head) self-attention CTAs = one wave at 2 CTAs / SM

int resolve_chunk(int batch, int chunk) {
  if (chunk <= 0) chunk = kDefaultChunk;
  return chunk < batch ? chunk : batch;
}

struct Ws {
  float *x, *y, *q, *hid, *img, *jf, *yj, *kv, *hd, *coarse, *a3, *a3img;
  size_t a3img_bytes, bytes;
};

constexpr int kSuperChunk = 8192;   // samples per upsample_conv GEMM launch (im2col rows kept for that many)

Ws carve(float* base, int nb, int J, int nsuper) {
  Ws w;
  const size_t mv = (size_t)nb * V;
  size_t off = 0;
  auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += (n + 63) / 64 * 64; return p; };
  w.x = take(mv * E);
  w.y = take(mv * E);
  w.q = take(mv * E);
  w.hid = take(mv * 256);
  w.img = take(self_attn2_image_bytes(nb) / sizeof(float));   // fp16 [Q | K | V] operand images of the self-attention core
  const size_t msj = (size_t)nsuper * J;   // joint rows of a super-chunk: K|V of all 3 layers are computed once for it
  w.jf = take(msj * E);
  w.yj = take(msj * E);
  w.kv = take(msj * 2 * E * GATOR_MDR_LAYERS);
  w.hd = take(mv * HEADN);
  w.coarse = take((size_t)nb * V * 3);
  w.a3 = take((size_t)nsuper * 3 * UPK);
  w.a3img_bytes = wide_a_image_bytes(nsuper * 3, UPK);   // split bf16 image of a3 for the wide-N kernel
  w.a3img = take(w.a3img_bytes / sizeof(float));
  w.bytes = off * sizeof(float);
  return w;
}

}  // namespace
}  // namespace gator

extern "C" const char* gator_mdr_slot_name(int slot) {
  using namespace gator;
  if (slot < 0) return nullptr;
  if (slot < MDR_NUM_GLOBAL) return kGlobalNames[slot];
  if (slot < MDR_NUM_GLOBAL + MDRL_NUM) return kLayerNames[slot - MDR_NUM_GLOBAL];
  return nullptr;
}

extern "C" size_t gator_mdr_workspace_bytes(int32_t batch, int32_t num_joint, int32_t chunk) {
  using namespace gator;
  if (batch <= 0 || num_joint <= 0) return 0;
  return align_up(carve(nullptr, resolve_chunk(batch, chunk), num_joint, batch < kSuperChunk ? batch : kSuperChunk).bytes, 256);
}

namespace gator {
namespace {
int launch_self_attn(const float* qkv, float* out, int nb, cudaStream_t stream) {
  constexpr int sa_smem = 2 * SA_VPAD * DK * (int)sizeof(float);
  static DeviceOnce attr_once;
  GATOR_TRY(attr_once.run("mdr_self_attn", [&](int) -> cudaError_t {
    return cudaFuncSetAttribute(mdr_self_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sa_smem);
  }));
  mdr_self_attn_kernel<<<nb * 2, SA_THREADS, sa_smem, stream>>>(qkv, out);
  return check_launch("mdr_self_attn");
}
}  // namespace
}  // namespace gator

extern "C" int gator_mdr_self_attention(const float* qkv, float* out, int32_t batch, int32_t precision, void* stream) {
  using namespace gator;
  GATOR_REQUIRE(qkv && out && batch >= 0, "gator_mdr_self_attention: bad argument");
  GATOR_REQUIRE(precision == GATOR_PREC_FP32, "gator_mdr_self_attention: fp32 FFMA core only; the tensor-core core is gator_mdr_self_attention_f16");
  if (batch == 0) return GATOR_OK;
  return launch_self_attn(qkv, out, batch, (cudaStream_t)stream);
}

extern "C" size_t gator_mdr_self_attention_image_bytes(int32_t batch) { return batch > 0 ? gator::self_attn2_image_bytes(batch) : 0; }

extern "C" int gator_mdr_self_attention_f16(const float* qkv, void* image, float* out, int32_t batch, void* stream) {
  using namespace gator;
  GATOR_REQUIRE(qkv && out && image && batch >= 0, "gator_mdr_self_attention_f16: bad argument");
  if (batch == 0) return GATOR_OK;
  GATOR_TRY(launch_qkv_image(qkv, image, batch, (cudaStream_t)stream));
  return launch_self_attn2(image, out, batch, (cudaStream_t)stream);
}

extern "C" int gator_mdr_self_attention_core(const void* image, float* out, int32_t batch, void* stream) {
  using namespace gator;
  GATOR_REQUIRE(image && out && batch >= 0, "gator_mdr_self_attention_core: bad argument");
  if (batch == 0) return GATOR_OK;
  return launch_self_attn2(image, out, batch, (cudaStream_t)stream);
}

extern "C" int gator_mdr_layer_chain(const void* const* weights, int32_t layer, int32_t num_joint, int32_t precision,
                                     const float* x_in, const float* att_in, const float* kv, float* x3_out,
                                     float* qkv_out, void* image_out, int32_t batch, void* stream) {
  using namespace gator;
  GATOR_REQUIRE(weights && x_in && kv && x3_out && (qkv_out || image_out), "gator_mdr_layer_chain: null buffer");
  GATOR_REQUIRE(layer >= 0 && layer < GATOR_MDR_LAYERS && num_joint >= 2 && num_joint <= MAXJ && batch >= 0,
                "gator_mdr_layer_chain: bad argument");
  GATOR_REQUIRE(precision == GATOR_PREC_BF16 || precision == GATOR_PREC_BF16X3, "gator_mdr_layer_chain: tensor-core precisions only");
  GATOR_REQUIRE((layer == 0) == (att_in == nullptr), "gator_mdr_layer_chain: att_in must be given exactly for layers 1, 2");
  if (batch == 0) return GATOR_OK;
  const int base = MDR_NUM_GLOBAL + layer * MDRL_NUM, pbase = MDR_NUM_GLOBAL + (layer > 0 ? layer - 1 : layer) * MDRL_NUM;
  auto W = [&](int s) { return static_cast<const float*>(weights[base + s]); };
  GATOR_REQUIRE(weights[base + MDRL_CHAIN], "gator_mdr_layer_chain: CHAIN blob missing");
  const float* prm[11] = {static_cast<const float*>(weights[pbase + MDRL_SO_B]), W(MDRL_N1_W), W(MDRL_N1_B), W(MDRL_PROJ_B),
                          W(MDRL_N2_W), W(MDRL_N2_B), W(MDRL_FC1_B), W(MDRL_FC2_B), W(MDRL_CLN_A), W(MDRL_CLN_B), W(MDRL_SQKV_B)};
  return launch_mdr_chain2(x_in, att_in, kv, weights[base + MDRL_CHAIN], prm, x3_out, qkv_out, image_out, nullptr, batch, num_joint,
                           (cudaStream_t)stream);
}

extern "C" int gator_mdr_forward(const gator_mdr_args* a, void* stream_) {
  using namespace gator;
  cudaStream_t stream = (cudaStream_t)stream_;
  GATOR_REQUIRE(a, "gator_mdr_forward: null args");
  const int J = a->num_joint, B = a->batch;
  GATOR_REQUIRE(J >= 2 && J <= MAXJ, "gator_mdr_forward: num_joint=%d out of range [2,%d]", J, MAXJ);
  GATOR_REQUIRE(a->precision >= GATOR_PREC_FP32 && a->precision <= GATOR_PREC_BF16X3, "gator_mdr_forward: bad precision");
  if (B == 0) return GATOR_OK;
  GATOR_REQUIRE(B > 0 && a->weights && a->pose2d && a->pose3d && a->feat && a->mesh, "gator_mdr_forward: null buffer");
  const int nslots = MDR_NUM_GLOBAL + GATOR_MDR_LAYERS * MDRL_NUM;
  bool have_chain = true;
  for (int i = 0; i < nslots; ++i) {
    if (i == MDR_UP_W_WIDE) continue;
    if (i == MDR_CHAIN_FINAL || (i >= MDR_NUM_GLOBAL && (i - MDR_NUM_GLOBAL) % MDRL_NUM == MDRL_CHAIN)) {
      have_chain = have_chain && a->weights[i] != nullptr;
      continue;
    }
    GATOR_REQUIRE(a->weights[i], "gator_mdr_forward: weight slot %d is null", i);
  }
  const size_t need = gator_mdr_workspace_bytes(B, J, a->chunk);
  if (!a->workspace || a->workspace_bytes < need) {
    set_error("gator_mdr_forward: workspace %zu < %zu bytes", a->workspace_bytes, need);
    return GATOR_ERR_WORKSPACE;
  }
  auto G = [&](int s) { return static_cast<const float*>(a->weights[s]); };
  auto GB = [&](int s) { PackedW w; w.hi = a->weights_bf16 ? a->weights_bf16[s] : nullptr; w.lo = a->weights_bf16_lo ? a->weights_bf16_lo[s] : nullptr; return w; };
  // debug/ablation: `reserved` is a bit mask of the groups that take the bf16 kernels (0 = all):
  //   2 layer GEMMs, 4 self-attention, 8 head GEMM, 16 upsample_conv, 32 joint-feature GEMM
  const int mask = a->reserved ? a->reserved : ~0;
  auto P = [&](int bit) { return (a->precision != GATOR_PREC_FP32 && (mask & bit)) ? a->precision : (int)GATOR_PREC_FP32; };
  const int prec = P(2);
  const int cb = resolve_chunk(B, a->chunk);
  const int nsuper = B < kSuperChunk ? B : kSuperChunk;
  Ws w = carve(static_cast<float*>(a->workspace), cb, J, nsuper);

  for (int s0 = 0; s0 < B; s0 += nsuper) {
  const int ns = (B - s0 < nsuper) ? B - s0 : nsuper;
  {
    // joint tokens of the whole super-chunk: embedding, then LayerNorm1 + [Wk;Wv] of every layer (they depend
    // only on the joint features, MDR.py:130,140-153) - 8 launches instead of 4 per 148-sample chunk
    const int Msj = ns * J;
    mdr_embed_kernel<<<ns, 256, 0, stream>>>(a->pose2d + (size_t)s0 * J * 2, a->pose3d + (size_t)s0 * J * 3,
                                             G(MDR_JF_WPOSE), G(MDR_VF_CONST), G(MDR_VF_W3),
                                             static_cast<const int*>(a->weights[MDR_VJ]), w.jf, nullptr, J, a->pose3d_metres);
    GATOR_TRY(check_launch("mdr_embed_joints"));
    Epilogue e;
    e.bias_rows = G(MDR_JF_BIASROWS);
    e.bias_period = J;
    e.R = w.jf;
    e.ldr = E;
    GATOR_TRY(gemm(P(32), a->feat + (size_t)s0 * J * 128, 128, G(MDR_JF_WFEAT), 128, GB(MDR_JF_WFEAT), w.jf, E, Msj, E, 128, e, stream));
    for (int l = 0; l < GATOR_MDR_LAYERS; ++l) {
      const int base = MDR_NUM_GLOBAL + l * MDRL_NUM;
      GATOR_TRY(layernorm_rows(w.jf, w.yj, static_cast<const float*>(a->weights[base + MDRL_N1_W]),
                               static_cast<const float*>(a->weights[base + MDRL_N1_B]), Msj, E, 0, 0, stream));
      GATOR_TRY(gemm(prec, w.yj, E, static_cast<const float*>(a->weights[base + MDRL_WKV]), E, GB(base + MDRL_WKV),
                     w.kv + (size_t)l * Msj * 2 * E, 2 * E, Msj, 2 * E, E, Epilogue(), stream));
    }
  }
  for (int b0 = s0; b0 < s0 + ns; b0 += cb) {
    const int nb = (s0 + ns - b0 < cb) ? s0 + ns - b0 : cb;
    const int Mv = nb * V;
    Epilogue e;
    auto KV = [&](int l) { return w.kv + ((size_t)l * ns + (b0 - s0)) * J * 2 * E; };   // this chunk's K|V rows of layer l

    // any bit of the ablation mask disables the fused layer kernel
    const bool fused = a->precision != GATOR_PREC_FP32 && have_chain && a->reserved == 0;
    ChainEmbed emb;                  // fused path: the vertex embedding is computed inside the layer-0 chain kernel
    emb.vconst = G(MDR_VF_CONST); emb.w3 = G(MDR_VF_W3); emb.vj = static_cast<const int*>(a->weights[MDR_VJ]);
    emb.pose3d = a->pose3d + (size_t)b0 * J * 3; emb.metres = a->pose3d_metres;
    if (!fused) {
      mdr_embed_kernel<<<nb, 256, 0, stream>>>(a->pose2d + (size_t)b0 * J * 2, a->pose3d + (size_t)b0 * J * 3,
                                               G(MDR_JF_WPOSE), G(MDR_VF_CONST), G(MDR_VF_W3),
                                               static_cast<const int*>(a->weights[MDR_VJ]), nullptr, w.x, J, a->pose3d_metres);
      GATOR_TRY(check_launch("mdr_embed"));
    }
    for (int l = 0; fused && l < GATOR_MDR_LAYERS; ++l) {
      const int base = MDR_NUM_GLOBAL + l * MDRL_NUM;
      auto W = [&](int s) { return static_cast<const float*>(a->weights[base + s]); };
      auto WB = [&](int s) { return GB(base + s); };
      const int pbase = MDR_NUM_GLOBAL + (l > 0 ? l - 1 : l) * MDRL_NUM;     // previous layer's linears.3 bias
      const float* prm[11] = {static_cast<const float*>(a->weights[pbase + MDRL_SO_B]), W(MDRL_N1_W), W(MDRL_N1_B),
                              W(MDRL_PROJ_B), W(MDRL_N2_W), W(MDRL_N2_B), W(MDRL_FC1_B), W(MDRL_FC2_B),
                              W(MDRL_CLN_A), W(MDRL_CLN_B), W(MDRL_SQKV_B)};
      const float* prm2[11] = {W(MDRL_SO_B), G(MDR_HEAD_B), W(MDRL_N1_B), W(MDRL_PROJ_B), W(MDRL_N2_W), W(MDRL_N2_B),
                               W(MDRL_FC1_B), W(MDRL_FC2_B), W(MDRL_CLN_A), W(MDRL_CLN_B), W(MDRL_SQKV_B)};
      // layer chain (3-term bf16 split GEMMs, operands in tensor memory) -> fp16 operand images -> fp16 self-attention core
      GATOR_TRY(launch_mdr_chain2(l == 0 ? nullptr : w.q, l == 0 ? nullptr : w.y, KV(l), a->weights[base + MDRL_CHAIN], prm,
                                  w.q, nullptr, w.img, nullptr, nb, J, stream, l == 0 ? &emb : nullptr));
      GATOR_TRY(launch_self_attn2(w.img, w.y, nb, stream));
      if (l == GATOR_MDR_LAYERS - 1)   // last layer: x = x3 + linears.3(att) + b and the head projection, fused
        GATOR_TRY(launch_mdr_chain2(w.q, w.y, KV(l), a->weights[MDR_CHAIN_FINAL], prm2, nullptr, nullptr, nullptr, w.hd, nb, J, stream));
    }
    for (int l = 0; !fused && l < GATOR_MDR_LAYERS; ++l) {
      const int base = MDR_NUM_GLOBAL + l * MDRL_NUM;
      auto W = [&](int s) { return static_cast<const float*>(a->weights[base + s]); };
      auto WB = [&](int s) { return GB(base + s); };
      // CrossAttentionBlock
      GATOR_TRY(layernorm_rows(w.x, w.y, W(MDRL_N1_W), W(MDRL_N1_B), Mv, E, 0, 0, stream));
      GATOR_TRY(gemm(prec, w.y, E, W(MDRL_WQ), E, WB(MDRL_WQ), w.q, E, Mv, E, E, Epilogue(), stream));
      mdr_cross_attn_kernel<<<nb, 256, 0, stream>>>(w.q, KV(l), w.y, J);
      GATOR_TRY(check_launch("mdr_cross_attn"));
      e = Epilogue();
      e.bias = W(MDRL_PROJ_B);
      e.R = w.x;
      e.ldr = E;
      GATOR_TRY(gemm(prec, w.y, E, W(MDRL_PROJ_W), E, WB(MDRL_PROJ_W), w.x, E, Mv, E, E, e, stream));
      GATOR_TRY(layernorm_rows(w.x, w.y, W(MDRL_N2_W), W(MDRL_N2_B), Mv, E, 0, 0, stream));
      e = Epilogue();
      e.bias = W(MDRL_FC1_B);
      e.act = 1;
      GATOR_TRY(gemm(prec, w.y, E, W(MDRL_FC1_W), E, WB(MDRL_FC1_W), w.hid, 256, Mv, 256, E, e, stream));
      e = Epilogue();
      e.bias = W(MDRL_FC2_B);
      e.R = w.x;
      e.ldr = E;
      GATOR_TRY(gemm(prec, w.hid, 256, W(MDRL_FC2_W), 256, WB(MDRL_FC2_W), w.x, E, Mv, E, 256, e, stream));
      // unbiased-std LayerNorm, then x = x3 + selfatt(x3)
      GATOR_TRY(layernorm_rows(w.x, w.q, W(MDRL_CLN_A), W(MDRL_CLN_B), Mv, E, 1, 0, stream));
      e = Epilogue();
      e.bias = W(MDRL_SQKV_B);
      GATOR_TRY(gemm(prec, w.q, E, W(MDRL_SQKV_W), E, WB(MDRL_SQKV_W), w.hid, 3 * E, Mv, 3 * E, E, e, stream));
      if (P(4) != GATOR_PREC_FP32) {
        GATOR_TRY(launch_qkv_image(w.hid, w.img, nb, stream));
        GATOR_TRY(launch_self_attn2(w.img, w.y, nb, stream));
      } else {
        GATOR_TRY(launch_self_attn(w.hid, w.y, nb, stream));
      }
      e = Epilogue();
      e.bias = W(MDRL_SO_B);
      e.R = w.q;
      e.ldr = E;
      GATOR_TRY(gemm(prec, w.y, E, W(MDRL_SO_W), E, WB(MDRL_SO_W), w.x, E, Mv, E, E, e, stream));
    }
    // head
    if (!fused) {
      e = Epilogue();
      e.bias = G(MDR_HEAD_B);
      GATOR_TRY(gemm(P(8), w.x, E, G(MDR_HEAD_W), E, GB(MDR_HEAD_W), w.hd, HEADN, Mv, HEADN, E, e, stream));
    }
    mdr_head_kernel<<<ceil_div(nb, HD_NS), 256, 0, stream>>>(w.hd, G(MDR_BNORM_SCALE), G(MDR_BNORM_SHIFT), G(MDR_BCONV_W),
                                                             G(MDR_BCONV_B), a->alpha, a->coarse ? a->coarse + (size_t)b0 * V * 3 : nullptr,
                                                             w.a3 + (size_t)(b0 - s0) * 3 * UPK, nb);
    GATOR_TRY(check_launch("mdr_head"));
  }
  // upsample_conv + template for the whole super-chunk: (3 ns) x 1296 @ 1296 x 6890, scattered to (b, vertex, xyz)
  Epilogue e;
  e.conv3 = 1;
  e.bias_rows = G(MDR_UP_BIAST);
  if (P(16) == GATOR_PREC_BF16X3 && a->weights[MDR_UP_W_WIDE]) {
    GATOR_TRY(gemm_bf16x3_wide(w.a3, UPK, a->weights[MDR_UP_W_WIDE], w.a3img, w.a3img_bytes, a->mesh + (size_t)s0 * VF * 3, 0,
                               ns * 3, VF, UPK, e, stream));
  } else {
    GATOR_TRY(gemm(P(16), w.a3, UPK, G(MDR_UP_W), UPK, GB(MDR_UP_W), a->mesh + (size_t)s0 * VF * 3, 0, ns * 3, VF, UPK, e, stream));
  }
  }
  return GATOR_OK;
}
