// Development probe for tcgen05 operand forms (one CTA, one 128 x N x K product), used by tools/umma_forms_probe.py to
// pin down on hardware the three forms the MDR self-attention core relies on:
//   bit 0 of `mode`: fp16 operands (instruction descriptor a/b format 0) instead of bf16
//   bit 1          : B operand MN-major (n contiguous) in the no-swizzle canonical layout
//                      offset(n, k) = (n/8)*SBO + (k/8)*LBO + (k%8)*16 + (n%8)*2        (bytes)
//   bit 2          : A operand read from tensor memory (tcgen05.mma [d], [a], bdesc ...), written by tcgen05.st:
//                      lane = row, 32-bit column c of a K=16 slice holds k = 2c (low half) and 2c+1 (high half)
// A (128, K) and B (N, K) arrive as fp32 and are rounded to the 16-bit format in the kernel; D (128, N) fp32.
#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace gator {
namespace {

using namespace umma;

__device__ __forceinline__ uint16_t to16(float x, bool f16) {
  if (f16) {
    __half h = __float2half_rn(x);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __nv_bfloat16 b = __float2bfloat16_rn(x);
  return *reinterpret_cast<uint16_t*>(&b);
}

__global__ void __launch_bounds__(128, 1) umma_forms_probe_kernel(const float* A, const float* B, float* D, int N, int K, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool f16 = mode & 1, b_mn = mode & 2, a_tmem = mode & 4, swap = mode & 8;   // swap: exchange LBO / SBO of the MN-major descriptor
  uint8_t* sA = smem;                         // K-major image [16 row groups][K/8][8][8]
  uint8_t* sB = smem + 128 * K * 2;           // K-major: [N/8][K/8][8][8]; MN-major: [K/8][N/8][8 k][8 n]
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 32) {
    mbar_init(&bar, 1);
    mbar_init_fence();
  }
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<uint16_t*>(sA + (r >> 3) * (K * 16) + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2) = to16(A[i], f16);
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    const size_t off = b_mn ? (size_t)(k >> 3) * (N * 16) + (n >> 3) * 128 + (k & 7) * 16 + (n & 7) * 2
                            : (size_t)(n >> 3) * (K * 16) + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<uint16_t*>(sB + off) = to16(B[i], f16);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  const uint32_t a_col = 256;                 // A operand region in tensor memory (D occupies columns [0, N))
  if (a_tmem) {                               // thread = row: pack k pairs into 32-bit columns
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t r[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int k = (c0 + c) * 2;
        r[c] = (uint32_t)to16(A[tid * K + k], f16) | ((uint32_t)to16(A[tid * K + k + 1], f16) << 16);
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + lane_addr + a_col + c0),
                   "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (tid == 0) {
    uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (!f16) idesc |= (1u << 7) | (1u << 10);
    if (b_mn) idesc |= 1u << 16;
    for (int ks = 0; ks < K / 16; ++ks) {
      // K-major: K-adjacent core matrices 128 B apart (LBO), 8-row groups K*16 B apart (SBO)
      // MN-major: n-adjacent core matrices 128 B apart (SBO), 8-k groups N*16 B apart (LBO)
      const uint64_t bd = b_mn ? smem_desc(smem_u32(sB) + ks * 2 * (N * 16), swap ? 128 : N * 16, swap ? N * 16 : 128) : smem_desc(smem_u32(sB) + ks * 256, 128, K * 16);
      if (a_tmem) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
            "}\n" ::"r"(tmem), "r"(tmem + a_col + ks * 8), "l"(bd), "r"(idesc), "r"((uint32_t)(ks > 0))
            : "memory");
      } else {
        mma_bf16(tmem, smem_desc(smem_u32(sA) + ks * 256, 128, K * 16), bd, idesc, ks > 0);
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + lane_addr + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[tid * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace
}  // namespace gator

extern "C" int gator_umma_forms_probe(const float* A, const float* B, float* D, int N, int K, int mode, void* stream) {
  using namespace gator;
  GATOR_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256, "probe: N in [16,256] step 16, K in [16,256] step 16");
  const int smem = (128 + N) * K * 2;
  cudaFuncSetAttribute(umma_forms_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_forms_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(A, B, D, N, K, mode);
  return check_launch("umma_forms_probe");
}
