// GAT lifter forward (lib/models/GAT.py:133-152): skeleton embedding, 6 GATBlocks, LN/GELU, lifter.
// Dense products go through gemm_f32 (fp32 parity path); the per-sample J x J work (attention with
// hop/path bias, modulated graph convolution, hop-masked feature mixing) is hand-fused here with the
// whole joint-token sequence of a sample in shared memory / registers.
#include "common.cuh"

namespace gator {
namespace {

constexpr int C = 128;       // embed_dim at every call site (base.py:57,59; demo/run.py:96)
constexpr int H = 8;         // heads
constexpr int DH = 16;       // head dim
constexpr int MAXJ = 32;

// ---------------------------------------------------------------------------------------------
// Embedding: GraphLinear(2,64) -> GroupNorm(4,64) -> GELU -> GraphLinear(64,128) -> + pos consts
// (GAT.py:69-72,135-144; modules.py:49-50).  One CTA (128 threads) per sample.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gat_embed_kernel(const float* __restrict__ pose2d, const float* __restrict__ w1, const float* __restrict__ b1,
                 const float* __restrict__ gnw, const float* __restrict__ gnb, const float* __restrict__ w2t,
                 const float* __restrict__ b2, const float* __restrict__ posc, float* __restrict__ x, int J) {
  __shared__ float p[MAXJ * 2];
  __shared__ __align__(16) float g[64][MAXJ];        // GELU(GroupNorm(.)) [channel][joint]; read as float4 over the joints (broadcast)
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < 2 * J) p[tid] = pose2d[(size_t)b * J * 2 + tid];
  __syncthreads();
  if (tid < 64) {
    const int c = tid;
    const float wa = w1[c * 2 + 0], wb = w1[c * 2 + 1], bb = b1[c];
    float h[MAXJ];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      if (j < J) { h[j] = fmaf(wa, p[2 * j], fmaf(wb, p[2 * j + 1], 0.f)) + bb; s += h[j]; }
    }
    // group = 16 consecutive channels = 16 consecutive lanes
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
    const float mean = s / (16.f * J);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      if (j < J) { float d = h[j] - mean; q = fmaf(d, d, q); }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o, 16);
    const float rstd = rsqrtf(q / (16.f * J) + 1e-5f);
    const float gw = gnw[c], gb = gnb[c];
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      if (j < J) g[c][j] = gelu_erf((h[j] - mean) * rstd * gw + gb);
    }
  }
  __syncthreads();
  {
    const int c = tid;   // 128 output channels
    float acc[MAXJ];
    const float bb = b2[c];
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) acc[j] = 0.f;
    const int j4n = (J + 3) >> 2;
#pragma unroll 4
    for (int k = 0; k < 64; ++k) {
      const float w = __ldg(w2t + k * C + c);
      const float4* gk = reinterpret_cast<const float4*>(g[k]);
#pragma unroll
      for (int j4 = 0; j4 < MAXJ / 4; ++j4) {
        if (j4 < j4n) {                     // (entries j >= J of a row are never written; their products are never stored)
          const float4 t = gk[j4];
          acc[4 * j4] = fmaf(w, t.x, acc[4 * j4]); acc[4 * j4 + 1] = fmaf(w, t.y, acc[4 * j4 + 1]);
          acc[4 * j4 + 2] = fmaf(w, t.z, acc[4 * j4 + 2]); acc[4 * j4 + 3] = fmaf(w, t.w, acc[4 * j4 + 3]);
        }
      }
    }
    float* xo = x + (size_t)b * J * C;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
      if (j < J) xo[j * C + c] = acc[j] + bb + __ldg(posc + j * C + c);
  }
}

// ---------------------------------------------------------------------------------------------
// Attention core (modules.py:121-138): per sample, per head: softmax_J(q k^T * 0.25 + bias) v.
// One CTA per sample, warp h = head h; qkv of the sample staged in shared memory (row stride 385
// so that the J key rows fall in distinct banks).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gat_attn_kernel(const float* __restrict__ qkv, const float* __restrict__ bias, float* __restrict__ o, int J) {
  extern __shared__ float sm[];
  __shared__ float sp[H][MAXJ];
  constexpr int LD = 3 * C + 1;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* src = qkv + (size_t)b * J * 3 * C;
  for (int i = tid; i < J * 3 * C; i += 256) {
    int r = i / (3 * C), c = i - r * 3 * C;
    sm[r * LD + c] = src[i];
  }
  __syncthreads();
  const int h = tid >> 5, lane = tid & 31;
  const float* bh = bias + (size_t)h * J * J;
  const int d = lane & 15, half = lane >> 4;
  for (int i = 0; i < J; ++i) {
    float s = -INFINITY;
    if (lane < J) {
      const float* qi = sm + i * LD + h * DH;
      const float* kj = sm + lane * LD + C + h * DH;
      float a = 0.f;
#pragma unroll
      for (int t = 0; t < DH; ++t) a = fmaf(qi[t], kj[t], a);
      s = a * 0.25f + __ldg(bh + i * J + lane);
    }
    const float m = warp_max(s);
    const float e = (lane < J) ? expf(s - m) : 0.f;
    const float denom = warp_sum(e);
    const float pj = e / denom;
    // out[d] = sum_j p_j v[j][d]; the two half-warps take alternate keys
    if (lane < J) sp[h][lane] = pj;
    __syncwarp();
    float acc = 0.f;
    for (int j = half; j < J; j += 2) acc = fmaf(sp[h][j], sm[j * LD + 2 * C + h * DH + d], acc);
    __syncwarp();
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (half == 0) o[((size_t)b * J + i) * C + h * DH + d] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// Modulated graph convolution mix (modules.py:243-255), given h01 = n @ [W0 | W1]:
//   g[b,i,c] = Adiag[i] M[i,c] h0[b,i,c] + sum_j Aoff[i,j] M[j,c] h1[b,j,c] + bias[c]
// One CTA (128 threads = channels) per sample.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gat_gcn_mix_kernel(const float* __restrict__ h01, const float* __restrict__ Mmod, const float* __restrict__ adiag,
                   const float* __restrict__ aoff, const float* __restrict__ gbias, float* __restrict__ g, int J) {
  __shared__ float sa[MAXJ * MAXJ];
  __shared__ float sd[MAXJ];
  const int b = blockIdx.x, c = threadIdx.x;
  for (int i = c; i < J * J; i += 128) sa[i] = aoff[i];
  if (c < J) sd[c] = adiag[c];
  __syncthreads();
  const float* hb = h01 + (size_t)b * J * 2 * C;
  float m1[MAXJ];
#pragma unroll
  for (int j = 0; j < MAXJ; ++j)
    if (j < J) m1[j] = __ldg(Mmod + j * C + c) * hb[j * 2 * C + C + c];
  const float bb = gbias[c];
  float* gb = g + (size_t)b * J * C;
  for (int i = 0; i < J; ++i) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
      if (j < J) acc = fmaf(sa[i * J + j], m1[j], acc);
    // torch: matmul(adj*E, M*h0) + matmul(adj*(1-E), M*h1) + bias
    const float dsum = sd[i] * (__ldg(Mmod + i * C + c) * hb[i * 2 * C + c]);
    gb[i * C + c] = (dsum + acc) + bb;
  }
}

// ---------------------------------------------------------------------------------------------
// X_Feat hop mixing (modules.py:158-177): f[:, :128] = 1[hop<=1] @ y[:, :128]; f[:,128:144] = 1[hop==2] @ y[:,128:144]
// One CTA (160 threads, 144 active channels) per sample.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160)
gat_hop_mix_kernel(const float* __restrict__ y, const float* __restrict__ mask1, const float* __restrict__ mask2,
                   float* __restrict__ f, int J) {
  constexpr int CY = 144;
  __shared__ float s1[MAXJ * MAXJ], s2[MAXJ * MAXJ];
  const int b = blockIdx.x, c = threadIdx.x;
  for (int i = c; i < J * J; i += 160) { s1[i] = mask1[i]; s2[i] = mask2[i]; }
  __syncthreads();
  if (c >= CY) return;
  const float* yb = y + (size_t)b * J * CY;
  const float* mk = c < C ? s1 : s2;
  float col[MAXJ];
#pragma unroll
  for (int j = 0; j < MAXJ; ++j)
    if (j < J) col[j] = yb[j * CY + c];
  float* fb = f + (size_t)b * J * CY;
  for (int i = 0; i < J; ++i) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
      if (j < J) acc = fmaf(mk[i * J + j], col[j], acc);
    fb[i * CY + c] = acc;
  }
}

const char* kGlobalNames[GAT_NUM_GLOBAL] = {
    "EMB_W1", "EMB_B1", "GN_W", "GN_B", "EMB_W2T", "EMB_B2", "POS_CONST", "ATTN_BIAS",
    "HOP_MASK1", "HOP_MASK2", "NORM_W", "NORM_B", "LIFT_W", "LIFT_B", "CHAIN_BLOBS", "CHAIN_PRM"};
const char* kBlockNames[GATB_NUM] = {
    "LN1_W", "LN1_B", "QKV_W", "QKV_B", "PROJ_W", "PROJ_B", "GCN_W01", "GCN_M", "GCN_ADIAG", "GCN_AOFF",
    "GCN_BIAS", "XF_W01", "XF_B01", "XF_WB", "XF_BB", "LN2_W", "LN2_B", "FC1_W", "FC1_B", "FC2_W", "FC2_B"};

// default chunk: 296 GEMM M-tiles of 128 token rows = two full waves of the 148 SMs (a 1024-sample chunk
// of J=19 gives 152 tiles: one wave plus a 4-CTA tail that costs as much as the wave)
constexpr int kTilesPerChunk = 296;
// per token row: x 128 | n 128 | big 512 | o 128 | h 256 | g 128
constexpr size_t kRowFloats = 128 + 128 + 512 + 128 + 256 + 128;

int resolve_chunk(int batch, int chunk, int J) {
  if (chunk <= 0) chunk = kTilesPerChunk * 128 / J;
  return chunk < batch ? chunk : batch;
}

}  // namespace
}  // namespace gator

extern "C" const char* gator_gat_slot_name(int slot) {
  using namespace gator;
  if (slot < 0) return nullptr;
  if (slot < GAT_NUM_GLOBAL) return kGlobalNames[slot];
  if (slot < GAT_NUM_GLOBAL + GATB_NUM) return kBlockNames[slot - GAT_NUM_GLOBAL];
  return nullptr;
}

extern "C" size_t gator_gat_workspace_bytes(int32_t batch, int32_t num_joint, int32_t chunk) {
  using namespace gator;
  if (batch <= 0 || num_joint <= 0) return 0;
  const int cb = resolve_chunk(batch, chunk, num_joint);
  return align_up((size_t)cb * num_joint * kRowFloats * sizeof(float), 256);
}

extern "C" int gator_gat_forward(const gator_gat_args* a, void* stream_) {
  using namespace gator;
  cudaStream_t stream = (cudaStream_t)stream_;
  GATOR_REQUIRE(a, "gator_gat_forward: null args");
  const int J = a->num_joint, B = a->batch;
  GATOR_REQUIRE(J >= 2 && J <= MAXJ, "gator_gat_forward: num_joint=%d out of range [2,%d]", J, MAXJ);
  GATOR_REQUIRE(a->depth >= 0 && a->depth <= 64, "gator_gat_forward: bad depth %d", a->depth);
  GATOR_REQUIRE(a->precision >= GATOR_PREC_FP32 && a->precision <= GATOR_PREC_BF16X3, "gator_gat_forward: bad precision");
  if (B == 0) return GATOR_OK;
  GATOR_REQUIRE(B > 0 && a->weights && a->pose2d && a->pose3d && a->feat, "gator_gat_forward: null buffer");
  const int nslots = GAT_NUM_GLOBAL + a->depth * GATB_NUM;
  for (int i = 0; i < nslots; ++i)
    GATOR_REQUIRE(a->weights[i] || i == GAT_CHAIN_BLOBS || i == GAT_CHAIN_PRM, "gator_gat_forward: weight slot %d is null", i);
  // tensor-core precisions run all GATBlocks in one fused kernel when its weight pieces were packed
  const bool fused = a->precision != GATOR_PREC_FP32 && a->weights[GAT_CHAIN_BLOBS] && a->weights[GAT_CHAIN_PRM] &&
                     gat_chain_supported(J) && a->reserved == 0;
  const size_t need = gator_gat_workspace_bytes(B, J, a->chunk);
  if (!a->workspace || a->workspace_bytes < need) {
    set_error("gator_gat_forward: workspace %zu < %zu bytes", a->workspace_bytes, need);
    return GATOR_ERR_WORKSPACE;
  }
  auto G = [&](int s) { return static_cast<const float*>(a->weights[s]); };
  auto GB = [&](int s) { PackedW w; w.hi = a->weights_bf16 ? a->weights_bf16[s] : nullptr; w.lo = a->weights_bf16_lo ? a->weights_bf16_lo[s] : nullptr; return w; };
  const int prec = a->precision;
  int cb = resolve_chunk(B, a->chunk, J);
  if (fused && a->chunk <= 0) {   // the fused kernel only needs x (128 floats / row): run the whole batch in one pass
    // (x plus the split-K partial sums of the lifter: at most 16 x 3J floats per sample)
    const size_t cap = a->workspace_bytes / ((size_t)J * (128 + 16 * 3) * sizeof(float));
    cb = (size_t)B < cap ? B : (int)cap;
  }
  const size_t rows_max = (size_t)cb * J;
  float* ws = static_cast<float*>(a->workspace);
  float* x = ws;
  float* n = x + rows_max * 128;
  float* big = n + rows_max * 128;
  float* o = big + rows_max * 512;
  float* h = o + rows_max * 128;
  float* g = h + rows_max * 256;

  static DeviceOnce attr_once;
  const int attn_smem = J * (3 * C + 1) * (int)sizeof(float);
  GATOR_TRY(attr_once.run("gat_attn", [&](int) -> cudaError_t {
    return cudaFuncSetAttribute(gat_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXJ * (3 * C + 1) * (int)sizeof(float));
  }));

  for (int b0 = 0; b0 < B; b0 += cb) {
    const int nb = (B - b0 < cb) ? B - b0 : cb;
    const int M = nb * J;
    gat_embed_kernel<<<nb, 128, 0, stream>>>(a->pose2d + (size_t)b0 * J * 2, G(GAT_EMB_W1), G(GAT_EMB_B1), G(GAT_GN_W),
                                             G(GAT_GN_B), G(GAT_EMB_W2T), G(GAT_EMB_B2), G(GAT_POS_CONST), x, J);
    GATOR_TRY(check_launch("gat_embed"));
    if (fused)
      GATOR_TRY(launch_gat_chain(x, M, J, a->depth, static_cast<const void* const*>(a->weights[GAT_CHAIN_BLOBS]),
                                 static_cast<const float* const*>(a->weights[GAT_CHAIN_PRM]), G(GAT_ATTN_BIAS),
                                 G(GAT_HOP_MASK1), G(GAT_HOP_MASK2), stream));
    for (int l = 0; !fused && l < a->depth; ++l) {
      const int base = GAT_NUM_GLOBAL + l * GATB_NUM;
      auto W = [&](int s) { return static_cast<const float*>(a->weights[base + s]); };
      auto WB = [&](int s) { return GB(base + s); };
      GATOR_TRY(layernorm_rows(x, n, W(GATB_LN1_W), W(GATB_LN1_B), M, C, 0, 0, stream));
      Epilogue e;
      e.bias = W(GATB_QKV_B);
      GATOR_TRY(gemm(prec, n, C, W(GATB_QKV_W), C, WB(GATB_QKV_W), big, 3 * C, M, 3 * C, C, e, stream));
      gat_attn_kernel<<<nb, 256, attn_smem, stream>>>(big, G(GAT_ATTN_BIAS), o, J);
      GATOR_TRY(check_launch("gat_attn"));
      GATOR_TRY(gemm(prec, n, C, W(GATB_GCN_W01), C, WB(GATB_GCN_W01), h, 2 * C, M, 2 * C, C, Epilogue(), stream));
      gat_gcn_mix_kernel<<<nb, 128, 0, stream>>>(h, W(GATB_GCN_M), W(GATB_GCN_ADIAG), W(GATB_GCN_AOFF),
                                                 W(GATB_GCN_BIAS), g, J);
      GATOR_TRY(check_launch("gat_gcn_mix"));
      // s = proj(o) + b + g  -> n
      e = Epilogue();
      e.bias = W(GATB_PROJ_B);
      e.R = g;
      e.ldr = C;
      GATOR_TRY(gemm(prec, o, C, W(GATB_PROJ_W), C, WB(GATB_PROJ_W), n, C, M, C, C, e, stream));
      // y = [L0(s) | L1(s)] -> h (ld 144)
      e = Epilogue();
      e.bias = W(GATB_XF_B01);
      GATOR_TRY(gemm(prec, n, C, W(GATB_XF_W01), C, WB(GATB_XF_W01), h, 144, M, 144, C, e, stream));
      gat_hop_mix_kernel<<<nb, 160, 0, stream>>>(h, G(GAT_HOP_MASK1), G(GAT_HOP_MASK2), big, J);
      GATOR_TRY(check_launch("gat_hop_mix"));
      // x = x + linearback(f)
      e = Epilogue();
      e.bias = W(GATB_XF_BB);
      e.R = x;
      e.ldr = C;
      GATOR_TRY(gemm(prec, big, 144, W(GATB_XF_WB), 144, WB(GATB_XF_WB), x, C, M, C, 144, e, stream));
      // x = x + fc2(gelu(fc1(LN2(x))))
      GATOR_TRY(layernorm_rows(x, n, W(GATB_LN2_W), W(GATB_LN2_B), M, C, 0, 0, stream));
      e = Epilogue();
      e.bias = W(GATB_FC1_B);
      e.act = 1;
      GATOR_TRY(gemm(prec, n, C, W(GATB_FC1_W), C, WB(GATB_FC1_W), big, 4 * C, M, 4 * C, C, e, stream));
      e = Epilogue();
      e.bias = W(GATB_FC2_B);
      e.R = x;
      e.ldr = C;
      GATOR_TRY(gemm(prec, big, 4 * C, W(GATB_FC2_W), 4 * C, WB(GATB_FC2_W), x, C, M, C, 4 * C, e, stream));
    }
    float* feat = a->feat + (size_t)b0 * J * C;
    GATOR_TRY(layernorm_rows(x, feat, G(GAT_NORM_W), G(GAT_NORM_B), M, C, 0, 1, stream));
    Epilogue e;
    e.bias = G(GAT_LIFT_B);
    // lifter: nb x 3J outputs over K = 128 J - a handful of 128-row tiles, so the K loop is split over blockIdx.z
    // (partials in the `n` workspace region, which is idle here)
    const PackedW lw = GB(GAT_LIFT_W);
    if (prec != GATOR_PREC_FP32 && lw.hi && (prec == GATOR_PREC_BF16 || lw.lo))
      GATOR_TRY(gemm_bf16_umma_splitk(feat, J * C, lw.hi, prec == GATOR_PREC_BF16X3 ? lw.lo : nullptr, a->pose3d + (size_t)b0 * 3 * J, 3 * J, nb,
                                      3 * J, J * C, e, n, a->workspace_bytes / sizeof(float) - rows_max * 128, stream));
    else
      GATOR_TRY(gemm(prec, feat, J * C, G(GAT_LIFT_W), J * C, GB(GAT_LIFT_W), a->pose3d + (size_t)b0 * 3 * J, 3 * J, nb, 3 * J, J * C, e, stream));
  }
  return GATOR_OK;
}
