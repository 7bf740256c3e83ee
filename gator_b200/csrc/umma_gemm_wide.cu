// Warp-specialised, persistent tcgen05 GEMM for the two wide products of the path, 3-term bf16 split:
//   upsample_conv  (MDR.py:167): (3B) x 1296 @ 1296 x 6890, scattered to (b, vertex, xyz)
//   SMPL blend shapes (smpl_layer.py:93-105): B x 220 @ 220 x 20670
//   C[m,n] = epi(sum_k A[m,k] W[n,k]),  A W ~= A_lo W_hi + A_hi W_lo + A_hi W_hi
//
// Both operands are pre-split bf16 (hi, lo) images stored tile-major, so that one K block of one tile is ONE
// contiguous chunk and a pipeline stage is two TMA bulk copies (cp.async.bulk + mbarrier complete_tx):
//   A image [m_tile][k_block][hi|lo][16 row groups][4 k-cores][8 rows][8 k]   16 KB per (tile, block)
//   W image [n_tile][k_block][hi|lo][32 row groups][4 k-cores][8 rows][8 k]   32 KB per (tile, block)
// (8x8 core matrices, K-major, no swizzle: LBO 128 B between K-adjacent cores, SBO 512 B between row groups).
//
// CTA = 10 warps, one CTA per SM, looping over 128 x 256 output tiles (m fastest, so concurrently running CTAs
// share W tiles in L2):
//   warp 0    producer: waits for a free stage, arms its mbarrier with the byte count, issues the two bulk copies
//   warp 1    MMA: waits for a full stage, issues 2 k-steps x 3 tcgen05.mma (128x256x16), commits the stage back
//             to the producer; after the last K block commits the accumulator to the epilogue
//   warps 2-9 epilogue (two groups of four warps taking alternate 32-column slabs): TMEM -> registers -> shared-memory
//             transpose -> coalesced stores (+bias / conv3 scatter), then hand the accumulator buffer back.  Two 256-column accumulators (all 512 TMEM columns) let the
//             epilogue of tile i overlap the main loop of tile i+1.
// 4 stages x 48 KB = 192 KB of operand ring + 2 x 16 KB staging.
#include "common.cuh"
#include "umma.cuh"

namespace gator {
namespace {

using namespace umma;

constexpr int WBM = 128, WBN = 256, WBK = 32;
constexpr int A_HALF = WBM * WBK * 2;                  // 8 KB: hi or lo of an A stage
constexpr int W_HALF = WBN * WBK * 2;                  // 16 KB
constexpr int A_STAGE = 2 * A_HALF, W_STAGE = 2 * W_HALF;
constexpr int STAGE_BYTES = A_STAGE + W_STAGE;         // 48 KB
constexpr int STG_LD = 32;                             // floats per staged row; 16-byte chunks are XOR-swizzled by the row
// Two shapes of the same kernel (operand ring depth x epilogue warp groups of 4 warps, alternating 32-column slabs):
//   <4, 2>  long K loops (upsample_conv, 41 K blocks per tile): tensor-bound, the deeper ring matters
//   <3, 4>  short K loops (SMPL blend shapes, 7 K blocks per tile): bound by draining 128 KB per tile, more drain warps
constexpr int wide_smem(int stages, int groups) { return stages * STAGE_BYTES + groups * WBM * STG_LD * 4; }
constexpr int wide_threads(int groups) { return (2 + 4 * groups) * 32; }

struct WideParams {
  const uint8_t* Aimg;
  const uint8_t* Wimg;
  float* C;
  int M, N, ldc, kblocks, m_tiles, n_tiles;
  int vec;                     // C rows are 16-byte aligned and the epilogue is bias-only: float4 stores
  Epilogue epi;
};

// fp32 A (M, K; lda) -> split bf16 tile image.  One thread = one 16-byte chunk (8 k of one row) of hi and of lo.
__global__ void __launch_bounds__(256)
wide_a_image_kernel(const float* __restrict__ A, int lda, int M, int K, int kblocks, long long chunks, uint8_t* __restrict__ img) {
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  if (c >= chunks) return;
  const int r = c & 7, kc = (c >> 3) & 3, rg = (c >> 5) & 15;
  const long long tb = c >> 9;                         // (m_tile, k_block)
  const int kb = (int)(tb % kblocks);
  const int mt = (int)(tb / kblocks);
  const int m = mt * WBM + rg * 8 + r, k = kb * WBK + kc * 8;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (m < M) {
    const float* src = A + (size_t)m * lda + k;
    if (k + 3 < K) a = *reinterpret_cast<const float4*>(src);
    else { if (k < K) a.x = src[0]; if (k + 1 < K) a.y = src[1]; if (k + 2 < K) a.z = src[2]; }
    if (k + 7 < K) b = *reinterpret_cast<const float4*>(src + 4);
    else { if (k + 4 < K) b.x = src[4]; if (k + 5 < K) b.y = src[5]; if (k + 6 < K) b.z = src[6]; }
  }
  const uint4 hi = cvt8(a, b);
  uint4* dst = reinterpret_cast<uint4*>(img + (size_t)tb * A_STAGE) + (c & 511);
  dst[0] = hi;
  dst[A_HALF / 16] = cvt8_residual(a, b, hi);
}

__device__ __forceinline__ void epi_bar(int group) { asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory"); }
// staged element (row, col): rows are 32 floats, the 16-byte chunk index is XORed with the row so that both the row-wise
// float4 writes (8 lanes = 8 rows) and the row-segment reads (8 lanes = 8 chunks of one row) are bank-conflict free
__device__ __forceinline__ int stg_idx(int r, int col) { return r * STG_LD + ((((col >> 2) ^ r) & 7) << 2) + (col & 3); }

template <int WSTAGES, int EPI_GROUPS>
__global__ void __launch_bounds__(wide_threads(EPI_GROUPS), 1)
umma_gemm_wide_kernel(WideParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[WSTAGES], empty[WSTAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  float* stg_all = reinterpret_cast<float*>(smem + WSTAGES * STAGE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int s = 0; s < WSTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 128 * EPI_GROUPS); mbar_init(&acc_empty[1], 128 * EPI_GROUPS);
    mbar_init_fence();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int total = p.m_tiles * p.n_tiles;
  const int kblocks = p.kblocks;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int nt = t / p.m_tiles, mt = t - nt * p.m_tiles;
        const uint8_t* asrc = p.Aimg + (size_t)mt * kblocks * A_STAGE;
        const uint8_t* wsrc = p.Wimg + (size_t)nt * kblocks * W_STAGE;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % WSTAGES;
          if (it >= WSTAGES) mbar_wait(&empty[s], ((it / WSTAGES) - 1) & 1);
          mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
          bulk_copy_g2s(smem + s * STAGE_BYTES, asrc + (size_t)kb * A_STAGE, A_STAGE, &full[s]);
          bulk_copy_g2s(smem + s * STAGE_BYTES + A_STAGE, wsrc + (size_t)kb * W_STAGE, W_STAGE, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(WBM, WBN);
      int it = 0, ti = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++ti) {
        const int buf = ti & 1;
        if (ti >= 2) mbar_wait(&acc_empty[buf], ((ti >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem + buf * WBN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % WSTAGES;
          mbar_wait(&full[s], (it / WSTAGES) & 1);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), w0 = a0 + A_STAGE;
#pragma unroll
          for (int ks = 0; ks < WBK / 16; ++ks) {
            const uint64_t ad = smem_desc(a0 + ks * 256, 128, 512), adl = smem_desc(a0 + A_HALF + ks * 256, 128, 512);
            const uint64_t bd = smem_desc(w0 + ks * 256, 128, 512), bdl = smem_desc(w0 + W_HALF + ks * 256, 128, 512);
            mma_bf16(d, adl, bd, idesc, (kb | ks) != 0);
            mma_bf16(d, ad, bdl, idesc, 1);
            mma_bf16(d, ad, bd, idesc, 1);
          }
          mma_commit(&empty[s]);
        }
        mma_commit(&acc_full[buf]);
      }
    }
  } else {
    const int grp = (warp - 2) >> 2;                   // epilogue group: slabs grp, grp + EPI_GROUPS, ...
    const int gw = (warp - 2) & 3;                     // warp within the group
    const int et = gw * 32 + lane;                     // 0..127 within the group
    const int q = warp & 3;                            // this warp's TMEM lane quarter
    const int row = q * 32 + lane;
    float* stg = stg_all + grp * WBM * STG_LD;
    const Epilogue& e = p.epi;
    int ti = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++ti) {
      const int nt = t / p.m_tiles, mt = t - nt * p.m_tiles;
      const int m0 = mt * WBM, n0 = nt * WBN;
      const int buf = ti & 1;
      mbar_wait(&acc_full[buf], (ti >> 1) & 1);
      tc_fence_after();
      for (int slab = grp; slab < WBN / 32; slab += EPI_GROUPS) {
        const int nb = n0 + slab * 32;
        if (nb >= p.N) break;                          // uniform over the group
        // (the bias of this thread's four columns is requested before the accumulator is read: its L2 latency used to
        //  sit between the staging barrier and the first store)
        const int c4v = (lane & 7) * 4;
        float4 bn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.vec && e.bias && nb + c4v + 3 < p.N) bn = __ldg(reinterpret_cast<const float4*>(e.bias + nb + c4v));
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * WBN + slab * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stg + stg_idx(row, j)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        epi_bar(grp);
        if (e.conv3) {
          // rows are (b, t) = (m / 3, m % 3); for one b the 32 columns x 3 taps are 96 contiguous floats of C
          const int b_first = m0 / 3;
          const int nbb = (min(m0 + WBM, p.M) - 1) / 3 - b_first + 1;
          for (int f = et; f < nbb * 96; f += 128) {
            const int bl = f / 96, rem = f - bl * 96;
            const int no = rem / 3, tt = rem - no * 3;
            const int m = (b_first + bl) * 3 + tt, r = m - m0, n = nb + no;
            if (r >= 0 && r < WBM && m < p.M && n < p.N)
              p.C[((size_t)(b_first + bl) * p.N + n) * 3 + tt] = stg[stg_idx(r, no)] + __ldg(e.bias_rows + (size_t)n * 3 + tt);
          }
        } else if (p.vec) {
          // 8 lanes x float4 = one 128-byte row segment; a warp stores 4 rows per instruction
          const int c4 = (lane & 7) * 4, n = nb + c4;
          const int rs = gw * 4 + (lane >> 3);
          if (n + 3 < p.N) {
            // all eight row segments are read before the first store: with one load / add / store per iteration the
            // compiler reused one register quad and every load waited for the previous store to drain
            float4 x[WBM / 16];
#pragma unroll
            for (int i = 0; i < WBM / 16; ++i) x[i] = *reinterpret_cast<const float4*>(stg + stg_idx(rs + i * 16, c4));
#pragma unroll
            for (int i = 0; i < WBM / 16; ++i) {
              const int m = m0 + rs + i * 16;
              if (m < p.M)
                *reinterpret_cast<float4*>(p.C + (size_t)m * p.ldc + n) = make_float4(x[i].x + bn.x, x[i].y + bn.y, x[i].z + bn.z, x[i].w + bn.w);
            }
          } else {
            for (int i = 0; i < WBM / 16; ++i) {
              const int r = rs + i * 16, m = m0 + r;
              for (int j = 0; j < 4; ++j)
                if (m < p.M && n + j < p.N) p.C[(size_t)m * p.ldc + n + j] = stg[stg_idx(r, c4 + j)] + (e.bias ? __ldg(e.bias + n + j) : 0.f);
            }
          }
        } else {
          const int n = nb + lane;
          const int ew = gw;
          if (n < p.N) {
            const float bn = e.bias ? __ldg(e.bias + n) : 0.f;
            for (int r = ew; r < WBM; r += 4) {
              const int m = m0 + r;
              if (m >= p.M) break;
              float x = stg[stg_idx(r, lane)] + bn;
              if (e.bias_rows) x += __ldg(e.bias_rows + (size_t)(m % e.bias_period) * p.N + n);
              if (e.act == 1) x = gelu_erf(x);
              if (e.R) x += e.R[(size_t)m * e.ldr + n];
              p.C[(size_t)m * p.ldc + n] = x;
            }
          }
        }
        epi_bar(grp);
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

size_t wide_a_image_bytes(int M, int K) {
  const size_t m_tiles = (M + WBM - 1) / WBM, kblocks = (K + WBK - 1) / WBK;
  return m_tiles * kblocks * A_STAGE;
}

void wide_weight_layout(int N, int K, int* n_tiles, int* kblocks) {
  *n_tiles = (N + WBN - 1) / WBN;
  *kblocks = (K + WBK - 1) / WBK;
}

int gemm_bf16x3_wide(const float* A, int lda, const void* Wimg, void* a_img, size_t a_img_bytes, float* C, int ldc,
                     int M, int N, int K, const Epilogue& epi, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return GATOR_OK;
  GATOR_REQUIRE(A && Wimg && a_img && C, "gemm_bf16x3_wide: null operand");
  GATOR_REQUIRE(K > 0 && lda % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15u) == 0, "gemm_bf16x3_wide: A must be 16-byte aligned, lda %% 4 == 0");
  GATOR_REQUIRE(((reinterpret_cast<uintptr_t>(Wimg) | reinterpret_cast<uintptr_t>(a_img)) & 15u) == 0, "gemm_bf16x3_wide: images must be 16-byte aligned");
  GATOR_REQUIRE(a_img_bytes >= wide_a_image_bytes(M, K), "gemm_bf16x3_wide: A image workspace too small");
  GATOR_REQUIRE(!epi.conv3 || epi.bias_rows, "gemm_bf16x3_wide: conv3 needs bias_rows");
  WideParams p;
  p.Aimg = static_cast<const uint8_t*>(a_img);
  p.Wimg = static_cast<const uint8_t*>(Wimg);
  p.C = C; p.M = M; p.N = N; p.ldc = ldc;
  p.m_tiles = (M + WBM - 1) / WBM;
  wide_weight_layout(N, K, &p.n_tiles, &p.kblocks);
  p.epi = epi;
  p.vec = (!epi.conv3 && !epi.bias_rows && !epi.R && epi.act == 0 && ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15u) == 0 &&
           (!epi.bias || (reinterpret_cast<uintptr_t>(epi.bias) & 15u) == 0)) ? 1 : 0;
  const long long chunks = (long long)p.m_tiles * p.kblocks * 512;
  wide_a_image_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, stream>>>(A, lda, M, K, p.kblocks, chunks, static_cast<uint8_t*>(a_img));
  GATOR_TRY(check_launch("wide_a_image"));
  static DeviceOnce attr_once;
  static int sm_count[64];
  int dev = 0;
  cudaGetDevice(&dev);
  GATOR_TRY(attr_once.run("umma_gemm_wide", [&](int d) -> cudaError_t {
    GATOR_CUDA_OK(cudaFuncSetAttribute(umma_gemm_wide_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, wide_smem(4, 2)));
    GATOR_CUDA_OK(cudaFuncSetAttribute(umma_gemm_wide_kernel<3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, wide_smem(3, 4)));
    return cudaDeviceGetAttribute(&sm_count[d & 63], cudaDevAttrMultiProcessorCount, d);
  }));
  const int sms = sm_count[dev & 63] > 0 ? sm_count[dev & 63] : 148;
  const int total = p.m_tiles * p.n_tiles;
  if (p.kblocks <= 8)
    umma_gemm_wide_kernel<3, 4><<<total < sms ? total : sms, wide_threads(4), wide_smem(3, 4), stream>>>(p);
  else
    umma_gemm_wide_kernel<4, 2><<<total < sms ? total : sms, wide_threads(2), wide_smem(4, 2), stream>>>(p);
  return check_launch("umma_gemm_wide");
}

}  // namespace gator

extern "C" int gator_umma_wide_layout(int32_t N, int32_t K, int32_t* n_tiles, int32_t* kblocks) {
  if (N <= 0 || K <= 0 || !n_tiles || !kblocks) return GATOR_ERR_BAD_ARG;
  int a, b;
  gator::wide_weight_layout(N, K, &a, &b);
  *n_tiles = a; *kblocks = b;
  return GATOR_OK;
}
extern "C" size_t gator_umma_wide_a_bytes(int32_t M, int32_t K) {
  return (M > 0 && K > 0) ? gator::wide_a_image_bytes(M, K) : 0;
}
