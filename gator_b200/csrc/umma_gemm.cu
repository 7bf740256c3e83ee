// tcgen05 GEMM for the skinny products of the forward (M = tokens, K, N <= a few hundred; and the two wide
// ones: upsample_conv N = 6890, SMPL blend shapes N = 20670):
//   C[m,n] = act(sum_k A[m,k] W[n,k] + bias[n] + bias_rows[m % P, n]) + R[m,n]
// A is fp32 in HBM; it is converted to bf16 on the way into shared memory (UMMA canonical K-major
// layout, see umma.cuh).  W is pre-packed bf16 in exactly the shared-memory image, so its staging is a
// linear 16-byte copy.  One CTA = one 128-row M tile x one BN-column N tile; the K loop is double
// buffered: the next 64-wide K block is staged while the tensor core works on the current one
// (tcgen05.commit -> mbarrier releases the buffer).  The fp32 accumulator lives in TMEM (BN columns);
// the epilogue moves it TMEM -> registers (thread = row) -> shared memory -> coalesced float4 stores with
// the bias / GELU / residual applied on the coalesced side.
#include "common.cuh"
#include "umma.cuh"

namespace gator {

void umma_weight_layout(int N, int K, int* BN, int* n_tiles, int* K_pad) {
  const int n32 = (N + 31) / 32 * 32;
  const int tiles = (n32 + 255) / 256;
  int bn = (n32 / 32 + tiles - 1) / tiles * 32;
  *BN = bn;
  *n_tiles = tiles;
  *K_pad = (K + 63) / 64 * 64;
}

namespace {

using namespace umma;

constexpr int BM = 128;
constexpr int BKE = 64;                    // K elements per stage
constexpr int A_STAGE = BM * BKE * 2;      // 16 KB
constexpr int STG_LD = 36;                 // floats per staged row (32 + 4 pad)
constexpr int STG_BYTES = BM * STG_LD * 4; // 18 KB per slab

struct UmmaGemmParams {
  const float* A;
  const __nv_bfloat16* Wp;     // hi part (or the only part)
  const __nv_bfloat16* Wlo;    // lo part for the 3-term split, else null
  float* C;
  int lda, ldc, M, N, K, K_pad, BN;   // BN = columns per CTA (the packed tile width, halved in split mode if > 128)
  int stages;                         // operand buffers: 2 = double-buffered K loop; 1 for short K loops, so that
                                      // several CTAs fit on an SM and overlap each other's staging / epilogue
  Epilogue epi;
  int vec;
  int ksplit;                         // > 1: blockIdx.z takes a contiguous share of the K blocks and writes raw partial sums
  size_t part_stride;                 //      to C + z * part_stride (the epilogue is applied by splitk_reduce_kernel)
};

__global__ void __launch_bounds__(256)
umma_gemm_kernel(UmmaGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mbar[2];
  __shared__ uint32_t tmem_slot;
  const int BN = p.BN;
  const bool split = p.Wlo != nullptr;
  const int w_stage = BN * BKE * 2;
  const int a_stride = split ? 2 * A_STAGE : A_STAGE;      // [hi | lo] per stage
  const int w_stride = split ? 2 * w_stage : w_stage;
  const int stages = p.stages;
  uint8_t* sA = smem;                          // [stages][a_stride]
  uint8_t* sW = smem + stages * a_stride;      // [stages][w_stride]
  float* stg = reinterpret_cast<float*>(smem); // epilogue staging aliases the operand buffers: [2][BM][STG_LD]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  uint32_t tcols = 32;
  while ((int)tcols < BN) tcols <<= 1;

  if (warp == 0) tmem_alloc(&tmem_slot, tcols);
  if (tid == 32) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_init_fence();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = idesc_bf16(BM, BN);

  const int nk_all = p.K_pad / BKE;
  const int kb_begin = p.ksplit > 1 ? (int)((long long)nk_all * blockIdx.z / p.ksplit) : 0;
  const int kb_end = p.ksplit > 1 ? (int)((long long)nk_all * (blockIdx.z + 1) / p.ksplit) : nk_all;
  const int nk = kb_end - kb_begin;
  const int kgroups = p.K_pad / 8;             // 16-byte chunks per packed W row-group
  for (int kbi = 0; kbi < nk; ++kbi) {
    const int kb = kb_begin + kbi;
    const int buf = kbi % stages;
    if (kbi >= stages) mbar_wait(&mbar[buf], ((kbi / stages) - 1) & 1);
    // ---- stage A: 128 rows x 64 k, fp32 -> bf16, chunk c = (rg, kc, r) stored linearly ----
    {
      uint4* dst = reinterpret_cast<uint4*>(sA + buf * a_stride);
      uint4* dst_lo = reinterpret_cast<uint4*>(sA + buf * a_stride + A_STAGE);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = i * 256 + tid;
        const int r = c & 7, kc = (c >> 3) & 7, rg = c >> 6;
        const int m = m0 + rg * 8 + r, k = kb * BKE + kc * 8;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (m < p.M) {
          const float* src = p.A + (size_t)m * p.lda + k;
          if (k < p.K) a = *reinterpret_cast<const float4*>(src);
          if (k + 4 < p.K) b = *reinterpret_cast<const float4*>(src + 4);
        }
        const uint4 hi = cvt8(a, b);
        dst[c] = hi;
        if (split) dst_lo[c] = cvt8_residual(a, b, hi);
      }
    }
    // ---- stage W: BN/8 row groups x 1 KB (8 k-chunks), already in the smem image ----
    {
      uint4* dst = reinterpret_cast<uint4*>(sW + buf * w_stride);
      uint4* dst_lo = reinterpret_cast<uint4*>(sW + buf * w_stride + w_stage);
      const uint4* src = reinterpret_cast<const uint4*>(p.Wp);
      const uint4* src_lo = reinterpret_cast<const uint4*>(p.Wlo);
      const int total = BN * 8;   // 16-byte chunks
      for (int c = tid; c < total; c += 256) {
        const int ng = c >> 6, j = c & 63;
        const size_t off = ((size_t)(n0 / 8 + ng) * kgroups + kb * 8) * 8 + j;
        dst[c] = __ldg(src + off);
        if (split) dst_lo[c] = __ldg(src_lo + off);
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sA + buf * a_stride), b0 = smem_u32(sW + buf * w_stride);
#pragma unroll
      for (int ks = 0; ks < BKE / 16; ++ks) {
        const uint64_t ad = smem_desc(a0 + ks * 256, 128, BKE * 16);
        const uint64_t bd = smem_desc(b0 + ks * 256, 128, BKE * 16);
        if (split) {   // a*w ~= a_lo*w_hi + a_hi*w_lo + a_hi*w_hi  (the dropped a_lo*w_lo term is ~2^-18 relative)
          const uint64_t adl = smem_desc(a0 + A_STAGE + ks * 256, 128, BKE * 16);
          const uint64_t bdl = smem_desc(b0 + w_stage + ks * 256, 128, BKE * 16);
          mma_bf16(tmem, adl, bd, idesc, (kbi | ks) != 0);
          mma_bf16(tmem, ad, bdl, idesc, 1);
          mma_bf16(tmem, ad, bd, idesc, 1);
        } else {
          mma_bf16(tmem, ad, bd, idesc, (kbi | ks) != 0);
        }
      }
      mma_commit(&mbar[buf]);
    }
  }
  mbar_wait(&mbar[(nk - 1) % stages], ((nk - 1) / stages) & 1);
  tc_fence_after();
  __syncthreads();   // every thread is past its last operand-buffer use: staging may alias them

  // ---- epilogue ----
  const int slabs = (BN + 31) / 32;
  const Epilogue e = p.ksplit > 1 ? Epilogue() : p.epi;
  float* const Cout = p.C + (p.ksplit > 1 ? blockIdx.z * p.part_stride : 0);
  for (int s0 = 0; s0 < slabs; s0 += 2) {
    const int my = s0 + (warp >> 2);
    if (my < slabs) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + my * 32, v);
      tmem_ld_wait();
      float* row = stg + (size_t)(warp >> 2) * BM * STG_LD + ((warp & 3) * 32 + lane) * STG_LD;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(row + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int slab = s0 + h;
      if (slab >= slabs) break;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * 256;
        const int r = idx >> 3, c4 = (idx & 7) * 4;
        const int m = m0 + r, n = n0 + slab * 32 + c4;
        if (m >= p.M || n >= p.N || n >= n0 + BN) continue;
        const float4 t = *reinterpret_cast<const float4*>(stg + (size_t)h * BM * STG_LD + r * STG_LD + c4);
        float v[4] = {t.x, t.y, t.z, t.w};
        if (e.conv3) {
          const int b = m / 3, tt = m - b * 3;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < p.N) Cout[((size_t)b * p.N + (n + j)) * 3 + tt] = v[j] + __ldg(e.bias_rows + (n + j) * 3 + tt);
          continue;
        }
        const float* brow = e.bias_rows ? e.bias_rows + (size_t)(m % e.bias_period) * p.N : nullptr;
        if (p.vec && n + 3 < p.N) {
          if (e.bias) { const float4 bb = __ldg(reinterpret_cast<const float4*>(e.bias + n)); v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w; }
          if (brow) { const float4 bb = __ldg(reinterpret_cast<const float4*>(brow + n)); v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w; }
          if (e.act == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = gelu_erf(v[j]);
          }
          if (e.R) { const float4 rr = *reinterpret_cast<const float4*>(e.R + (size_t)m * e.ldr + n); v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w; }
          *reinterpret_cast<float4*>(Cout + (size_t)m * p.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n + j >= p.N) break;
            float x = v[j];
            if (e.bias) x += __ldg(e.bias + n + j);
            if (brow) x += __ldg(brow + n + j);
            if (e.act == 1) x = gelu_erf(x);
            if (e.R) x += e.R[(size_t)m * e.ldr + n + j];
            Cout[(size_t)m * p.ldc + n + j] = x;
          }
        }
      }
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tcols);
}

inline bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

// C[m,n] = epi(sum_z part[z][m][n]) in a fixed order (deterministic)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, size_t part_stride, int ksplit, float* __restrict__ C,
                                                            int ldc, int M, int N, Epilogue e) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), n = (int)(i - (long long)m * N);
  float x = 0.f;
  for (int z = 0; z < ksplit; ++z) x += part[z * part_stride + (size_t)m * ldc + n];
  if (e.bias) x += __ldg(e.bias + n);
  if (e.bias_rows) x += __ldg(e.bias_rows + (size_t)(m % e.bias_period) * N + n);
  if (e.act == 1) x = gelu_erf(x);
  if (e.R) x += e.R[(size_t)m * e.ldr + n];
  C[(size_t)m * ldc + n] = x;
}

}  // namespace

int gemm_bf16_umma(const float* A, int lda, const void* Wpacked, const void* Wpacked_lo, float* C, int ldc, int M,
                   int N, int K, const Epilogue& epi, cudaStream_t stream) {
  return gemm_bf16_umma_splitk(A, lda, Wpacked, Wpacked_lo, C, ldc, M, N, K, epi, nullptr, 0, stream);
}

// Same product for long K and few output columns (the lifter: K = 128 J, N = 3 J): the K blocks are shared out over
// blockIdx.z (8 ways) and the partial sums (ws: 8 x M x ldc floats) are added in a fixed order by a second small kernel -
// a 4096-row batch is only 32 row tiles, which left 116 of the 148 SMs idle for the length of a 38-block K loop.
int gemm_bf16_umma_splitk(const float* A, int lda, const void* Wpacked, const void* Wpacked_lo, float* C, int ldc, int M,
                          int N, int K, const Epilogue& epi, float* ws, size_t ws_floats, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return GATOR_OK;
  GATOR_REQUIRE(A && Wpacked && C, "gemm_bf16_umma: null operand");
  GATOR_REQUIRE(K > 0 && K % 4 == 0 && lda % 4 == 0, "gemm_bf16_umma: K=%d lda=%d must be multiples of 4", K, lda);
  GATOR_REQUIRE(aligned16(A) && aligned16(Wpacked), "gemm_bf16_umma: A/W must be 16-byte aligned");
  UmmaGemmParams p;
  int n_tiles;
  umma_weight_layout(N, K, &p.BN, &n_tiles, &p.K_pad);
  p.A = A; p.Wp = static_cast<const __nv_bfloat16*>(Wpacked); p.Wlo = static_cast<const __nv_bfloat16*>(Wpacked_lo); p.C = C;
  if (p.Wlo && p.BN > 128) {   // split mode doubles the operand footprint: halve the CTA's column tile
    p.BN /= 2;
    n_tiles *= 2;
  }
  p.lda = lda; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.epi = epi;
  p.vec = (!epi.conv3 && ldc % 4 == 0 && aligned16(C) && (!epi.R || (epi.ldr % 4 == 0 && aligned16(epi.R))) &&
           (!epi.bias || aligned16(epi.bias)) && (!epi.bias_rows || (aligned16(epi.bias_rows) && N % 4 == 0))) ? 1 : 0;
  p.stages = (p.K_pad / BKE <= 2) ? 1 : 2;
  const int operand = (p.Wlo ? 2 : 1) * p.stages * (A_STAGE + p.BN * BKE * 2);
  const int smem = operand > 2 * STG_BYTES ? operand : 2 * STG_BYTES;
  static DeviceOnce attr_once;
  GATOR_TRY(attr_once.run("umma_gemm", [&](int) -> cudaError_t {
    return cudaFuncSetAttribute(umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (2 * A_STAGE + 2 * 128 * BKE * 2));
  }));
  // The split depends on K only - never on M - so that a sample's result does not depend on the batch it is computed in
  // (the summation order over K is part of the result's last bits).
  const int nkb = p.K_pad / BKE;
  int ksplit = 1;
  if (ws && !epi.conv3 && nkb >= 16) {
    ksplit = nkb / 4 < 8 ? nkb / 4 : 8;
    GATOR_REQUIRE((size_t)ksplit * M * ldc <= ws_floats, "gemm_bf16_umma_splitk: workspace %zu < %zu floats", ws_floats, (size_t)ksplit * M * ldc);
  }
  p.ksplit = ksplit;
  p.part_stride = (size_t)M * ldc;
  if (ksplit > 1) p.C = ws;
  dim3 grid(ceil_div(M, BM), n_tiles, ksplit);
  umma_gemm_kernel<<<grid, 256, smem, stream>>>(p);
  GATOR_TRY(check_launch("umma_gemm"));
  if (ksplit > 1) {
    splitk_reduce_kernel<<<(unsigned)(((long long)M * N + 255) / 256), 256, 0, stream>>>(ws, p.part_stride, ksplit, C, ldc, M, N, epi);
    return check_launch("splitk_reduce");
  }
  return GATOR_OK;
}

}  // namespace gator

extern "C" int gator_umma_weight_layout(int32_t N, int32_t K, int32_t* BN, int32_t* n_tiles, int32_t* K_pad) {
  if (N <= 0 || K <= 0 || !BN || !n_tiles || !K_pad) return GATOR_ERR_BAD_ARG;
  int a, b, c;
  gator::umma_weight_layout(N, K, &a, &b, &c);
  *BN = a; *n_tiles = b; *K_pad = c;
  return GATOR_OK;
}
