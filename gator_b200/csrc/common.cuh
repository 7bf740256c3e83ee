// Shared device/host helpers for the gator_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include <mutex>

#include "../../include/gator_b200.h"

namespace gator {

// thread-local error text behind gator_last_error()
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> status

#define GATOR_REQUIRE(cond, ...)                       \
  do {                                                 \
    if (!(cond)) {                                     \
      ::gator::set_error(__VA_ARGS__);                 \
      return GATOR_ERR_BAD_ARG;                        \
    }                                                  \
  } while (0)

#define GATOR_TRY(expr)                                \
  do {                                                 \
    int _st = (expr);                                  \
    if (_st != GATOR_OK) return _st;                   \
  } while (0)

// One-time per-device setup of a kernel (cudaFuncSetAttribute is per device).  Thread-safe: concurrent first callers
// block until the setup has run, a device is only marked done after every call succeeded, a failure is reported through
// gator_last_error() and retried by the next call.  One static DeviceOnce per call site; devices 0..63.
struct DeviceOnce {
  std::mutex mu;
  unsigned long long done = 0;
  template <class F>
  int run(const char* what, F&& init) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(&done, __ATOMIC_ACQUIRE) & bit) return GATOR_OK;
    std::lock_guard<std::mutex> lock(mu);
    if (done & bit) return GATOR_OK;
    const cudaError_t e = init(dev);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("%s: kernel attribute setup failed on device %d: %s", what, dev, cudaGetErrorString(e));
      return GATOR_ERR_LAUNCH;
    }
    __atomic_or_fetch(&done, bit, __ATOMIC_RELEASE);
    return GATOR_OK;
  }
};
#define GATOR_CUDA_OK(expr)                            \
  do {                                                 \
    const cudaError_t _ce = (expr);                    \
    if (_ce != cudaSuccess) return _ce;                \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- epilogue descriptor shared by the GEMM launchers --------------------------------------
struct Epilogue {
  const float* bias = nullptr;        // (N)
  const float* bias_rows = nullptr;   // (period, N) indexed by m % period
  int bias_period = 0;
  int act = 0;                        // 1 = GELU(erf)
  const float* R = nullptr;           // residual (M, ldr); may alias C
  int ldr = 0;
  // conv3 scatter (upsample_conv): row m = (b, t), col n = o -> C[(b*N + o)*3 + t] + bias_rows[o*3+t]
  int conv3 = 0;
};

// C = epi(A (M,K;lda) * W (N,K;ldw)^T).  fp32 FFMA.
int gemm_f32(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K,
             const Epilogue& epi, cudaStream_t stream);

// Same contract on the tensor cores: A fp32 (converted to bf16 in-kernel), W pre-packed bf16 in the UMMA
// shared-memory image (see umma_weight_layout / gator_b200/packing.py), fp32 accumulate in TMEM.
// Wpacked_lo != null selects the 3-term split (a_hi w_hi + a_lo w_hi + a_hi w_lo): ~2^-17 relative error.
int gemm_bf16_umma(const float* A, int lda, const void* Wpacked, const void* Wpacked_lo, float* C, int ldc, int M,
                   int N, int K, const Epilogue& epi, cudaStream_t stream);
int gemm_bf16_umma_splitk(const float* A, int lda, const void* Wpacked, const void* Wpacked_lo, float* C, int ldc, int M,
                          int N, int K, const Epilogue& epi, float* ws, size_t ws_floats, cudaStream_t stream);
void umma_weight_layout(int N, int K, int* BN, int* n_tiles, int* K_pad);

// Wide-N variant for GATOR_PREC_BF16X3 (csrc/umma_gemm_wide.cu): persistent, warp-specialised, TMA-fed.  Wimg is the
// tile-major split image of W (gator_b200/packing.py: pack_umma_wide); a_img is caller workspace of at least
// wide_a_image_bytes(M, K) into which A is split first.
size_t wide_a_image_bytes(int M, int K);
int gemm_bf16x3_wide(const float* A, int lda, const void* Wimg, void* a_img, size_t a_img_bytes, float* C, int ldc,
                     int M, int N, int K, const Epilogue& epi, cudaStream_t stream);

// SMPL skinning with the transform blend on the tensor cores (csrc/smpl_skin_umma.cu), GATOR_PREC_BF16X3 only
size_t skin_t_image_bytes(int S);
int launch_smpl_skin_umma(const float* vposed, int ld, const float* amat, const float* offset, const void* Wimg, void* timg,
                          float* verts, int S, float scale, cudaStream_t stream);

// precision dispatch used by the stage drivers: bf16 only if a packed weight exists for the slot
struct PackedW {
  const void* hi = nullptr;
  const void* lo = nullptr;
};
inline int gemm(int precision, const float* A, int lda, const float* W, int ldw, PackedW pw, float* C, int ldc,
                int M, int N, int K, const Epilogue& epi, cudaStream_t stream) {
  if (precision == GATOR_PREC_BF16 && pw.hi) return gemm_bf16_umma(A, lda, pw.hi, nullptr, C, ldc, M, N, K, epi, stream);
  if (precision == GATOR_PREC_BF16X3 && pw.hi && pw.lo) return gemm_bf16_umma(A, lda, pw.hi, pw.lo, C, ldc, M, N, K, epi, stream);
  return gemm_f32(A, lda, W, ldw, C, ldc, M, N, K, epi, stream);
}

// round-2 self-attention core (csrc/mdr_attn2_umma.cu): fp16 [Q | K | V] operand images per (sample, head) -> out (nb*431, 64)
size_t self_attn2_image_bytes(int nb);
int launch_qkv_image(const float* qkv, void* img, int nb, cudaStream_t stream);
int launch_self_attn2(const void* img, float* out, int nb, cudaStream_t stream);

// Fused row-wise chain of one MDR layer (csrc/mdr_chain2_umma.cu): x3_prev / att_prev -> x3, q|k|v.  A operands and the
// residual stream in tensor memory, TMA-fed weight ring, always the 3-term bf16 split.  hd_out != null selects the final pass
// (x3 + linears[3](att) -> MDR head projection -> hd (rows, 28)).  q|k|v leave as fp32 rows (qkv_out) and / or as the fp16 operand images of
// launch_self_attn2 (img_out); either may be null.
struct ChainEmbed {                  // operands of the vertex embedding, for layer 0 with x_in == null
  const float* vconst = nullptr;     // (431, 64) VF_CONST
  const float* w3 = nullptr;         // (64, 3)   VF_W3
  const int* vj = nullptr;           // (431)     nearest joint of each coarse vertex
  const float* pose3d = nullptr;     // (nb, J, 3)
  int metres = 0;
};
int launch_mdr_chain2(const float* x_in, const float* att_in, const float* kv, const void* blob, const float* const* prm,
                      float* x3_out, float* qkv_out, void* img_out, float* hd_out, int nb, int J, cudaStream_t stream,
                      const ChainEmbed* embed = nullptr);

// All GATBlocks of the lifter in one kernel (csrc/gat_chain2_umma.cu; always the 3-term bf16 split).  blobs_dev / prm_dev
// are DEVICE arrays of pointers: [depth] weight-piece blobs (34 x 32 KB) and [depth * 14] fp32 parameter arrays.
bool gat_chain_supported(int J);
int launch_gat_chain(float* x, int rows, int J, int depth, const void* const* blobs_dev, const float* const* prm_dev,
                     const float* attn_bias, const float* mask1, const float* mask2, cudaStream_t stream);

// LayerNorm over the last dim (C = 64 or 128).  mode 0: nn.LayerNorm (eps 1e-5, biased var);
// mode 1: a*(x-mean)/(std_unbiased+1e-6)+b (vanilla_transformer_encoder.py:31-34).  gelu: apply after.
int layernorm_rows(const float* x, float* y, const float* w, const float* b, int rows, int C, int mode,
                   int gelu, cudaStream_t stream);

#ifdef __CUDACC__
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, i.e. at the
// fp32 rounding level of erff itself) - 15 instructions: the reciprocal and the exponential are the raw MUFU
// approximations (rcp.approx / ex2.approx, ~1 ulp; __frcp_rn and exp2f wrap the same MUFU ops in range checks, a
// Newton step and a slow-path branch, which made this 35 instructions - measured in the MDR chain, profiles/r02_*).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * (x * -0.72134752044448170368f)));   // exp(-z^2) = 2^(-x^2 log2(e) / 2)
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(hx, copysignf(erf_abs, x), hx);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif

}  // namespace gator
