// SMPL linear blend skinning forward (smplpytorch/smplpytorch/pytorch/smpl_layer.py:65-158).
//   K4 smpl_pose_kernel : Rodrigues (rodrigues_layer.py:13-52) + 24-joint kinematic chain (:103-132),
//                         one warp per sample, lane = joint, tree walked level by level with shuffles.
//   K3b blend shapes    : v_posed = T + [S|P] . [beta | R-I]   (:87-99) - dense GEMM, K = 217 (+3 pad).
//   K5 smpl_skin_kernel : per-vertex blend of the joint transforms + apply (:134-145), ELL weights,
//                         output staged through shared memory and written as aligned float4 (fp32 path;
//                         GATOR_PREC_BF16X3 uses the tensor-core kernel in smpl_skin_umma.cu).
#include "common.cuh"

namespace gator {
namespace {

constexpr int NJ = GATOR_SMPL_JOINTS;   // 24
constexpr int KB = GATOR_SMPL_K;        // 220
constexpr int NV = GATOR_V_FULL;        // 6890
constexpr int NV3 = NV * 3;             // 20670
constexpr int VP_LD = 20672;            // row stride of the v_posed workspace: a multiple of 4 floats, so the GEMM epilogue
                                        // stores 16-byte vectors (20670 would force 4-byte stores)

__global__ void smpl_flags_kernel(const float* __restrict__ betas, int nb, const float* __restrict__ trans, int nt,
                                  int* __restrict__ flags) {
  // flags[0] = any(betas != 0), flags[1] = any(trans != 0)   (norm(x) == 0  <=>  all zero)
  __shared__ int f[2];
  if (threadIdx.x < 2) f[threadIdx.x] = 0;
  __syncthreads();
  int a = 0, b = 0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) a |= (betas[i] != 0.f);
  for (int i = threadIdx.x; i < nt; i += blockDim.x) b |= (trans[i] != 0.f);
  if (a) atomicOr(&f[0], 1);
  if (b) atomicOr(&f[1], 1);
  __syncthreads();
  if (threadIdx.x < 2) flags[threadIdx.x] = f[threadIdx.x];
}

struct PoseParams {
  const float* pose;        // (B,72)
  const float* betas;       // (B,10) or null
  const float* trans;       // (B,3) or null
  const float* def_betas;   // (10)
  const float* jt;          // (24,3)
  const float* js;          // (24,3,10)
  const int* parents;       // (24)
  const int* flags;         // null or [any_betas, any_trans]
  float* aop;               // (B,220)  [beta | R-I | 0]
  float* amat;              // (B,24,12) skinning transforms (top 3 rows of G')
  float* offset;            // (B,3) added to verts
  float* jtr;               // (B,24,3)
  int batch;
  int center_idx;
  float out_scale;          // applied to jtr after the translation
  int nj;                   // joints (24 for SMPL, 16 for MANO; <= 32: one lane per joint)
  int kb;                   // row length of aop: 10 + 9 (nj - 1), zero-padded to a multiple of 4
};

__global__ void __launch_bounds__(128)
smpl_pose_kernel(PoseParams p) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= p.batch) return;
  const bool use_betas = p.betas && (!p.flags || p.flags[0]);
  const bool use_trans = p.trans && (!p.flags || p.flags[1]);
  const float* beta = use_betas ? p.betas + (size_t)b * 10 : p.def_betas;
  const int NJ = p.nj, KB = p.kb;          // (shadow the SMPL constants: the kernel serves any LBS model of <= 32 joints)
  const int i = lane < NJ ? lane : NJ - 1;

  // --- Rodrigues: axis-angle -> unit quaternion -> rotation matrix ---
  const float ax = p.pose[((size_t)b * NJ + i) * 3 + 0];
  const float ay = p.pose[((size_t)b * NJ + i) * 3 + 1];
  const float az = p.pose[((size_t)b * NJ + i) * 3 + 2];
  const float ex = ax + 1e-8f, ey = ay + 1e-8f, ez = az + 1e-8f;
  const float angle = sqrtf(ex * ex + ey * ey + ez * ez);
  const float nx = ax / angle, ny = ay / angle, nz = az / angle;
  const float half = angle * 0.5f;
  const float cs = cosf(half), sn = sinf(half);
  float qw = cs, qx = sn * nx, qy = sn * ny, qz = sn * nz;
  const float qn = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
  qw /= qn; qx /= qn; qy /= qn; qz /= qn;
  const float w2 = qw * qw, x2 = qx * qx, y2 = qy * qy, z2 = qz * qz;
  const float wx = qw * qx, wy = qw * qy, wz = qw * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz;
  float R[9];
  R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;     R[2] = 2 * wy + 2 * xz;
  R[3] = 2 * wz + 2 * xy;     R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
  R[6] = 2 * xz - 2 * wy;     R[7] = 2 * wx + 2 * yz;     R[8] = w2 - x2 - y2 + z2;

  // --- blend-shape operand row: [beta(10) | R_1..R_23 - I (207) | 0 0 0] ---
  float* arow = p.aop + (size_t)b * KB;
  if (lane < 10) arow[lane] = beta[lane];
  if (lane >= 1 && lane < NJ) {
#pragma unroll
    for (int e = 0; e < 9; ++e) arow[10 + (lane - 1) * 9 + e] = R[e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f);
  }
  {
    const int used = 10 + 9 * (NJ - 1);    // 217 for SMPL, 145 for MANO
    if (lane < KB - used) arow[used + lane] = 0.f;
  }

  // --- rest joints: J_regressor @ (T + S beta) with the regressor pre-applied to T and S ---
  float j[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    float a = 0.f;
#pragma unroll
    for (int s = 0; s < 10; ++s) a = fmaf(__ldg(p.js + (i * 3 + t) * 10 + s), beta[s], a);
    j[t] = __ldg(p.jt + i * 3 + t) + a;
  }
  int parent = (i == 0) ? 0 : p.parents[i];
  parent = parent < 0 ? 0 : (parent >= NJ ? 0 : parent);
  int depth = 0;
  for (int q = i; q != 0 && depth < NJ; q = p.parents[q]) ++depth;

  // local transform: [R | j_i - j_parent] (root: [R | j_0])
  const float pjx = __shfl_sync(0xffffffffu, j[0], parent);
  const float pjy = __shfl_sync(0xffffffffu, j[1], parent);
  const float pjz = __shfl_sync(0xffffffffu, j[2], parent);
  float t[3] = {j[0] - (i ? pjx : 0.f), j[1] - (i ? pjy : 0.f), j[2] - (i ? pjz : 0.f)};
  // G = G_parent @ local, level by level
  float G[12];
#pragma unroll
  for (int r = 0; r < 3; ++r) { G[r * 4 + 0] = R[r * 3]; G[r * 4 + 1] = R[r * 3 + 1]; G[r * 4 + 2] = R[r * 3 + 2]; G[r * 4 + 3] = t[r]; }
  for (int level = 1; level < NJ; ++level) {
    const bool mine = (depth == level) && lane < NJ;
    if (!__any_sync(0xffffffffu, mine)) break;
    float P[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) P[e] = __shfl_sync(0xffffffffu, G[e], parent);
    if (mine) {
      float N[12];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          N[r * 4 + c] = P[r * 4 + 0] * R[c] + P[r * 4 + 1] * R[3 + c] + P[r * 4 + 2] * R[6 + c];
        N[r * 4 + 3] = P[r * 4 + 0] * t[0] + P[r * 4 + 1] * t[1] + P[r * 4 + 2] * t[2] + P[r * 4 + 3];
      }
#pragma unroll
      for (int e = 0; e < 12; ++e) G[e] = N[e];
    }
  }
  // translation of verts/joints (smpl_layer.py:148-155)
  float off[3] = {0.f, 0.f, 0.f};
  if (use_trans) {
#pragma unroll
    for (int c = 0; c < 3; ++c) off[c] = p.trans[(size_t)b * 3 + c];
  } else if (p.center_idx >= 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) off[c] = -__shfl_sync(0xffffffffu, G[c * 4 + 3], p.center_idx);
  }
  if (lane < NJ) {
    float* jo = p.jtr + ((size_t)b * NJ + lane) * 3;
    jo[0] = (G[3] + off[0]) * p.out_scale; jo[1] = (G[7] + off[1]) * p.out_scale; jo[2] = (G[11] + off[2]) * p.out_scale;
    // G' = G - pack(G @ [j; 0]): last column becomes t - R j
    float* ao = p.amat + ((size_t)b * NJ + lane) * 12;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      ao[r * 4 + 0] = G[r * 4 + 0]; ao[r * 4 + 1] = G[r * 4 + 1]; ao[r * 4 + 2] = G[r * 4 + 2];
      ao[r * 4 + 3] = G[r * 4 + 3] - (G[r * 4 + 0] * j[0] + G[r * 4 + 1] * j[1] + G[r * 4 + 2] * j[2]);
    }
  }
  if (lane < 3) p.offset[(size_t)b * 3 + lane] = off[lane];
}

// ---------------------------------------------------------------------------------------------
// Skinning (smpl_layer.py:134-145).  Memory-bound: per mesh 82 680 B written and 82 680 B of v_posed read back
// (the producer GEMM runs on the same workspace chunk).  CTA = 256 consecutive vertices x SG samples, 8 warps;
// the SG x 24 joint transforms are staged once, then every warp runs on its own 32 vertices with no
// block-level synchronisation: coalesced 8-byte loads of the 96-float v_posed segment -> warp-private
// shared memory -> lane = vertex (ELL weights in registers, reused across the SG samples) -> warp-private
// shared memory -> coalesced 8-byte stores (a sample row is 82 680 B = 8 mod 16, so 8 bytes is the widest
// access that is aligned for every sample; each store instruction still covers 256 contiguous bytes).
// The next sample's segment is prefetched into registers while the current one is processed.
// ---------------------------------------------------------------------------------------------
constexpr int SK_VT = 256;
constexpr int SK_SG = 16;
constexpr int SK_PD = 4;    // prefetch depth (samples)

__global__ void __launch_bounds__(SK_VT)
smpl_skin_kernel(const float* __restrict__ vposed, const float* __restrict__ amat, const float* __restrict__ offset,
                 const int* __restrict__ sidx, const float* __restrict__ sw, int KW, float* __restrict__ verts, int batch,
                 float out_scale, int nv, int nj, int ld) {
  const int NV = nv, NJ = nj, NV3 = nv * 3, VP_LD = ld;   // (shadow the SMPL constants: any model with nj <= 24, nv even)
  __shared__ float sA[SK_SG][12 * GATOR_SMPL_JOINTS];   // component-major [e][joint]: lanes reading different joints hit different banks
  __shared__ float soff[SK_SG][4];
  __shared__ __align__(16) float stage[SK_VT / 32][96];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int v0w = blockIdx.x * SK_VT + warp * 32;          // first vertex of this warp
  const int nvw = max(0, min(32, NV - v0w));               // vertices this warp owns (last tile is ragged)
  const int v = v0w + lane;
  const bool active = lane < nvw;
  const int s_begin = blockIdx.y * SK_SG;
  const int ns = min(SK_SG, batch - s_begin);
  for (int i = tid; i < ns * NJ * 12; i += SK_VT) {
    const int s = i / (NJ * 12), r = i - s * NJ * 12, j = r / 12, e = r - j * 12;
    sA[s][e * NJ + j] = amat[(size_t)s_begin * NJ * 12 + i];
  }
  for (int i = tid; i < ns * 3; i += SK_VT) soff[i / 3][i % 3] = offset[(size_t)s_begin * 3 + i];
  int jid[4];
  float jw[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    jid[k] = (active && k < KW && KW <= 4) ? __ldg(sidx + (size_t)v * KW + k) : 0;
    jw[k] = (active && k < KW && KW <= 4) ? __ldg(sw + (size_t)v * KW + k) : 0.f;
  }
  __syncthreads();
  if (nvw == 0) return;
  const int seg2 = nvw * 3 / 2;                            // float2 per segment (nvw*3 is even: 96 or 42)
  float* st = stage[warp];
  auto seg_ptr = [&](const float* base, int s, int ld) { return base + (size_t)(s_begin + s) * ld + (size_t)v0w * 3; };
  // v_posed segments are requested SK_PD samples ahead: with one sample in flight per warp the kernel was bound by
  // memory latency (~16 KB outstanding per SM), not by bandwidth
  float2 pa[SK_PD], pb[SK_PD];
  auto request = [&](int s, int slot) {
    pa[slot] = pb[slot] = make_float2(0.f, 0.f);
    if (s < ns) {
      const float2* src = reinterpret_cast<const float2*>(seg_ptr(vposed, s, VP_LD));
      if (lane < seg2) pa[slot] = src[lane];
      if (lane + 32 < seg2) pb[slot] = src[lane + 32];
    }
  };
#pragma unroll
  for (int u = 0; u < SK_PD; ++u) request(u, u);
  for (int s0 = 0; s0 < ns; s0 += SK_PD) {
#pragma unroll
  for (int u = 0; u < SK_PD; ++u) {
    const int s = s0 + u;
    if (s >= ns) break;                                    // warp-uniform
    __syncwarp();                                          // previous iteration's readers of `st` are done
    reinterpret_cast<float2*>(st)[lane] = pa[u];
    if (lane < 16) reinterpret_cast<float2*>(st)[lane + 32] = pb[u];
    request(s + SK_PD, u);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
    if (active) {
      const float px = st[lane * 3], py = st[lane * 3 + 1], pz = st[lane * 3 + 2];
      float T[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) T[e] = 0.f;
      if (KW <= 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float* a = &sA[s][jid[k]];
#pragma unroll
          for (int e = 0; e < 12; ++e) T[e] = fmaf(jw[k], a[e * NJ], T[e]);
        }
      } else {
        for (int k = 0; k < KW; ++k) {
          const int jj = __ldg(sidx + (size_t)v * KW + k);
          const float ww = __ldg(sw + (size_t)v * KW + k);
#pragma unroll
          for (int e = 0; e < 12; ++e) T[e] = fmaf(ww, sA[s][e * NJ + jj], T[e]);
        }
      }
      o0 = (T[0] * px + T[1] * py + T[2] * pz + T[3] + soff[s][0]) * out_scale;
      o1 = (T[4] * px + T[5] * py + T[6] * pz + T[7] + soff[s][1]) * out_scale;
      o2 = (T[8] * px + T[9] * py + T[10] * pz + T[11] + soff[s][2]) * out_scale;
    }
    __syncwarp();                                          // everyone has read its inputs from `st`
    if (active) { st[lane * 3] = o0; st[lane * 3 + 1] = o1; st[lane * 3 + 2] = o2; }
    __syncwarp();
    float2* dst = reinterpret_cast<float2*>(const_cast<float*>(seg_ptr(verts, s, NV3)));
    if (lane < seg2) dst[lane] = reinterpret_cast<const float2*>(st)[lane];
    if (lane + 32 < seg2) dst[lane + 32] = reinterpret_cast<const float2*>(st)[lane + 32];
  }
  }
}

// Samples per workspace chunk.  Measured at B = 16384 (bf16x3): 592 -> 2.42 ms, 1024 -> 2.19, 2048 -> 1.91, 4096 -> 1.77,
// 8192 -> 1.64: the persistent blend-shape GEMM needs many tiles per CTA to fill its waves, and keeping v_posed
// L2-resident with small chunks does not pay.  (Numbers taken with the CUDA-core skinning kernel and 4 epilogue warps
// in the GEMM; with the tensor-core skinning kernel of smpl_skin_umma.cu and 8-16 epilogue warps the same batch takes 1.05 ms.)
constexpr int kChunk = 8192;

struct Ws {
  float *aop, *amat, *offset, *vposed;
  void* aimg;               // split bf16 image of aop for the wide-N blend-shape GEMM
  size_t aimg_bytes;
  void* timg;               // transform tile image for the tensor-core skinning
  int* flags;
  size_t bytes;
};

Ws carve(char* base, int nb) {
  Ws w;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes, 256); return p; };
  w.flags = reinterpret_cast<int*>(take(16));
  w.aop = reinterpret_cast<float*>(take((size_t)nb * KB * 4));
  w.amat = reinterpret_cast<float*>(take((size_t)nb * NJ * 12 * 4));
  w.offset = reinterpret_cast<float*>(take((size_t)nb * 3 * 4));
  w.vposed = reinterpret_cast<float*>(take((size_t)nb * VP_LD * 4));
  w.aimg_bytes = wide_a_image_bytes(nb, KB);
  w.aimg = take(w.aimg_bytes);
  w.timg = take(skin_t_image_bytes(nb));
  w.bytes = off;
  return w;
}

}  // namespace
}  // namespace gator

extern "C" size_t gator_smpl_workspace_bytes(int32_t batch) {
  using namespace gator;
  if (batch <= 0) return 0;
  return carve(nullptr, batch < kChunk ? batch : kChunk).bytes;
}

extern "C" int gator_smpl_forward(const gator_smpl_args* a, void* stream_) {
  using namespace gator;
  cudaStream_t stream = (cudaStream_t)stream_;
  GATOR_REQUIRE(a, "gator_smpl_forward: null args");
  const int B = a->batch;
  if (B == 0) return GATOR_OK;
  GATOR_REQUIRE(B > 0 && a->pose && a->verts && a->jtr, "gator_smpl_forward: null buffer");
  GATOR_REQUIRE(a->parents && a->j_template && a->j_shapedirs && a->default_betas && a->blend_w && a->v_template &&
                    a->skin_idx && a->skin_w, "gator_smpl_forward: null model buffer");
  GATOR_REQUIRE(a->weights_per_vertex >= 1 && a->weights_per_vertex <= NJ, "gator_smpl_forward: bad weights_per_vertex");
  GATOR_REQUIRE(a->center_idx >= -1 && a->center_idx < NJ, "gator_smpl_forward: bad center_idx");
  GATOR_REQUIRE(!a->has_betas || a->betas, "gator_smpl_forward: has_betas without betas");
  GATOR_REQUIRE(!a->has_trans || a->trans, "gator_smpl_forward: has_trans without trans");
  GATOR_REQUIRE(a->precision >= GATOR_PREC_FP32 && a->precision <= GATOR_PREC_BF16X3, "gator_smpl_forward: bad precision");
  const size_t need = gator_smpl_workspace_bytes(B);
  if (!a->workspace || a->workspace_bytes < need) {
    set_error("gator_smpl_forward: workspace %zu < %zu bytes", a->workspace_bytes, need);
    return GATOR_ERR_WORKSPACE;
  }
  const int cb = B < kChunk ? B : kChunk;
  const float out_scale = a->out_scale == 0.f ? 1.f : a->out_scale;
  Ws w = carve(static_cast<char*>(a->workspace), cb);
  const bool flags = a->check_zero_norm && (a->has_betas || a->has_trans);
  if (flags) {
    smpl_flags_kernel<<<1, 256, 0, stream>>>(a->has_betas ? a->betas : nullptr, a->has_betas ? B * 10 : 0,
                                             a->has_trans ? a->trans : nullptr, a->has_trans ? B * 3 : 0, w.flags);
    GATOR_TRY(check_launch("smpl_flags"));
  }
  for (int b0 = 0; b0 < B; b0 += cb) {
    const int nb = (B - b0 < cb) ? B - b0 : cb;
    PoseParams p;
    p.pose = a->pose + (size_t)b0 * 72;
    p.betas = a->has_betas ? a->betas + (size_t)b0 * 10 : nullptr;
    p.trans = a->has_trans ? a->trans + (size_t)b0 * 3 : nullptr;
    p.def_betas = a->default_betas;
    p.jt = a->j_template;
    p.js = a->j_shapedirs;
    p.parents = a->parents;
    p.flags = flags ? w.flags : nullptr;
    p.aop = w.aop;
    p.amat = w.amat;
    p.offset = w.offset;
    p.jtr = a->jtr + (size_t)b0 * NJ * 3;
    p.batch = nb;
    p.center_idx = a->center_idx;
    p.out_scale = out_scale;
    p.nj = NJ;
    p.kb = KB;
    smpl_pose_kernel<<<ceil_div(nb, 4), 128, 0, stream>>>(p);
    GATOR_TRY(check_launch("smpl_pose"));
    Epilogue e;
    e.bias = a->v_template;
    if (a->precision == GATOR_PREC_BF16X3 && a->blend_w_wide) {
      GATOR_TRY(gemm_bf16x3_wide(w.aop, KB, a->blend_w_wide, w.aimg, w.aimg_bytes, w.vposed, VP_LD, nb, NV3, KB, e, stream));
    } else {
      GATOR_TRY(gemm(a->precision, w.aop, KB, a->blend_w, KB, PackedW{a->blend_w_bf16, a->blend_w_bf16_lo}, w.vposed, VP_LD, nb, NV3, KB, e, stream));
    }
    if (a->precision == GATOR_PREC_BF16X3 && a->skin_w_img) {
      GATOR_TRY(launch_smpl_skin_umma(w.vposed, VP_LD, w.amat, w.offset, a->skin_w_img, w.timg, a->verts + (size_t)b0 * NV3, nb,
                                      out_scale, stream));
    } else {
      dim3 grid(ceil_div(NV, SK_VT), ceil_div(nb, SK_SG));
      smpl_skin_kernel<<<grid, SK_VT, 0, stream>>>(w.vposed, w.amat, w.offset, a->skin_idx, a->skin_w,
                                                   a->weights_per_vertex, a->verts + (size_t)b0 * NV3, nb, out_scale, NV, NJ, VP_LD);
      GATOR_TRY(check_launch("smpl_skin"));
    }
  }
  return GATOR_OK;
}

// ---------------------------------------------------------------------------------------------
// Any body model shaped like SMPL (MANO: 778 vertices, 16 joints, 145 blend terms - manopth/manolayer.py:170-230 is the
// same arithmetic as smpl_layer.py:87-145) on the fp32 kernels above: pose kernel -> FFMA blend-shape GEMM -> skinning.
// ---------------------------------------------------------------------------------------------
namespace gator {
namespace {
struct LbsWs { float *aop, *amat, *offset, *vposed; int* flags; size_t bytes; int ld; };
LbsWs carve_lbs(char* base, int nb, int nv, int nj, int kb) {
  LbsWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes, 256); return p; };
  w.ld = (nv * 3 + 3) & ~3;
  w.flags = reinterpret_cast<int*>(take(16));
  w.aop = reinterpret_cast<float*>(take((size_t)nb * kb * 4));
  w.amat = reinterpret_cast<float*>(take((size_t)nb * nj * 12 * 4));
  w.offset = reinterpret_cast<float*>(take((size_t)nb * 3 * 4));
  w.vposed = reinterpret_cast<float*>(take((size_t)nb * w.ld * 4));
  w.bytes = off;
  return w;
}
bool lbs_dims_ok(int nv, int nj, int kb) {
  return nv >= 2 && nv % 2 == 0 && nj >= 2 && nj <= GATOR_SMPL_JOINTS && kb == ((10 + 9 * (nj - 1) + 3) & ~3);
}
}  // namespace
}  // namespace gator

extern "C" size_t gator_lbs_workspace_bytes(int32_t batch, int32_t n_verts, int32_t n_joints) {
  using namespace gator;
  const int kb = (10 + 9 * (n_joints - 1) + 3) & ~3;
  if (batch <= 0 || !lbs_dims_ok(n_verts, n_joints, kb)) return 0;
  return carve_lbs(nullptr, batch < kChunk ? batch : kChunk, n_verts, n_joints, kb).bytes;
}

extern "C" int gator_lbs_forward(const gator_lbs_args* a, void* stream_) {
  using namespace gator;
  cudaStream_t stream = (cudaStream_t)stream_;
  GATOR_REQUIRE(a, "gator_lbs_forward: null args");
  const int B = a->batch, nv = a->n_verts, nj = a->n_joints, kb = a->k_blend;
  GATOR_REQUIRE(lbs_dims_ok(nv, nj, kb), "gator_lbs_forward: need an even vertex count, 2..24 joints and k_blend = 10 + 9 (joints - 1) rounded up to a multiple of 4");
  if (B == 0) return GATOR_OK;
  GATOR_REQUIRE(B > 0 && a->pose && a->verts && a->jtr, "gator_lbs_forward: null buffer");
  GATOR_REQUIRE(a->parents && a->j_template && a->j_shapedirs && a->default_betas && a->blend_w && a->v_template && a->skin_idx &&
                    a->skin_w, "gator_lbs_forward: null model buffer");
  GATOR_REQUIRE(a->weights_per_vertex >= 1 && a->weights_per_vertex <= nj, "gator_lbs_forward: bad weights_per_vertex");
  GATOR_REQUIRE(a->center_idx >= -1 && a->center_idx < nj, "gator_lbs_forward: bad center_idx");
  GATOR_REQUIRE(!a->has_betas || a->betas, "gator_lbs_forward: has_betas without betas");
  GATOR_REQUIRE(!a->has_trans || a->trans, "gator_lbs_forward: has_trans without trans");
  const size_t need = gator_lbs_workspace_bytes(B, nv, nj);
  if (!a->workspace || a->workspace_bytes < need) {
    set_error("gator_lbs_forward: workspace %zu < %zu bytes", a->workspace_bytes, need);
    return GATOR_ERR_WORKSPACE;
  }
  const int cb = B < kChunk ? B : kChunk;
  const float out_scale = a->out_scale == 0.f ? 1.f : a->out_scale;
  LbsWs w = carve_lbs(static_cast<char*>(a->workspace), cb, nv, nj, kb);
  const bool flags = a->check_zero_norm && (a->has_betas || a->has_trans);
  if (flags) {
    smpl_flags_kernel<<<1, 256, 0, stream>>>(a->has_betas ? a->betas : nullptr, a->has_betas ? B * 10 : 0,
                                             a->has_trans ? a->trans : nullptr, a->has_trans ? B * 3 : 0, w.flags);
    GATOR_TRY(check_launch("smpl_flags"));
  }
  for (int b0 = 0; b0 < B; b0 += cb) {
    const int nb = (B - b0 < cb) ? B - b0 : cb;
    PoseParams p;
    p.pose = a->pose + (size_t)b0 * nj * 3;
    p.betas = a->has_betas ? a->betas + (size_t)b0 * 10 : nullptr;
    p.trans = a->has_trans ? a->trans + (size_t)b0 * 3 : nullptr;
    p.def_betas = a->default_betas;
    p.jt = a->j_template;
    p.js = a->j_shapedirs;
    p.parents = a->parents;
    p.flags = flags ? w.flags : nullptr;
    p.aop = w.aop;
    p.amat = w.amat;
    p.offset = w.offset;
    p.jtr = a->jtr + (size_t)b0 * nj * 3;
    p.batch = nb;
    p.center_idx = a->center_idx;
    p.out_scale = out_scale;
    p.nj = nj;
    p.kb = kb;
    smpl_pose_kernel<<<ceil_div(nb, 4), 128, 0, stream>>>(p);
    GATOR_TRY(check_launch("smpl_pose"));
    Epilogue e;
    e.bias = a->v_template;
    GATOR_TRY(gemm(GATOR_PREC_FP32, w.aop, kb, a->blend_w, kb, PackedW{nullptr, nullptr}, w.vposed, w.ld, nb, nv * 3, kb, e, stream));
    dim3 grid(ceil_div(nv, SK_VT), ceil_div(nb, SK_SG));
    smpl_skin_kernel<<<grid, SK_VT, 0, stream>>>(w.vposed, w.amat, w.offset, a->skin_idx, a->skin_w, a->weights_per_vertex,
                                                 a->verts + (size_t)b0 * nv * 3, nb, out_scale, nv, nj, w.ld);
    GATOR_TRY(check_launch("smpl_skin"));
  }
  return GATOR_OK;
}
