// fp32 FFMA GEMM with fused epilogue: the exact-parity path for every dense product of the forward
// (GAT.py / MDR.py / smpl_layer.py nn.Linear, matmul and Conv1d call sites).
//   C[m,n] = act(sum_k A[m,k] W[n,k] + bias[n] + bias_rows[m % P, n]) + R[m,n]
// Register-tiled (8 x TN per thread), K-contiguous operands staged transposed in shared memory,
// double buffered with register prefetch.
#include "common.cuh"

namespace gator {

namespace {

constexpr int BK = 16;

template <int BM, int BN, int TN>
__global__ void __launch_bounds__((BM / 8) * (BN / TN))
gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                float* C, int ldc, int M, int N, int K, Epilogue epi, int vec_store) {
  constexpr int NT = (BM / 8) * (BN / TN);
  constexpr int TXN = BN / TN;
  constexpr int A_F4 = BM * BK / 4 / NT;   // float4 loads per thread
  constexpr int B_F4 = BN * BK / 4 / NT;
  static_assert(A_F4 >= 1 && B_F4 >= 1, "tile too small for the thread count");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_F4], rb[B_F4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      int idx = tid + i * NT, row = idx >> 2, kq = idx & 3;
      int m = m0 + row, k = k0 + kq * 4;
      ra[i] = (m < M && k < K) ? *reinterpret_cast<const float4*>(A + (size_t)m * lda + k)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
      int idx = tid + i * NT, row = idx >> 2, kq = idx & 3;
      int n = n0 + row, k = k0 + kq * 4;
      rb[i] = (n < N && k < K) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)n * ldw + k))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      int idx = tid + i * NT, row = idx >> 2, kq = (idx & 3) * 4;
      As[buf][kq + 0][row] = ra[i].x; As[buf][kq + 1][row] = ra[i].y;
      As[buf][kq + 2][row] = ra[i].z; As[buf][kq + 3][row] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
      int idx = tid + i * NT, row = idx >> 2, kq = (idx & 3) * 4;
      Bs[buf][kq + 0][row] = rb[i].x; Bs[buf][kq + 1][row] = rb[i].y;
      Bs[buf][kq + 2][row] = rb[i].z; Bs[buf][kq + 3][row] = rb[i].w;
    }
  };

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][BM / 2 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      if (TN == 8) {
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][BN / 2 + tx * 4]);
        b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
    if (m >= M) continue;
    const float* brow = epi.bias_rows && !epi.conv3 ? epi.bias_rows + (size_t)(m % epi.bias_period) * N : nullptr;
#pragma unroll
    for (int g = 0; g < TN / 4; ++g) {
      const int n = n0 + (g == 0 ? tx * 4 : BN / 2 + tx * 4);
      if (n >= N) continue;
      float v[4] = {acc[i][g * 4 + 0], acc[i][g * 4 + 1], acc[i][g * 4 + 2], acc[i][g * 4 + 3]};
      if (epi.conv3) {
        const int b = m / 3, t = m - b * 3;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) C[((size_t)b * N + (n + j)) * 3 + t] = v[j] + __ldg(epi.bias_rows + (n + j) * 3 + t);
        continue;
      }
      if (vec_store && n + 3 < N) {
        if (epi.bias) {
          float4 bb = __ldg(reinterpret_cast<const float4*>(epi.bias + n));
          v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
        }
        if (brow) {
          float4 bb = __ldg(reinterpret_cast<const float4*>(brow + n));
          v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
        }
        if (epi.act == 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = gelu_erf(v[j]);
        }
        if (epi.R) {
          float4 rr = *reinterpret_cast<const float4*>(epi.R + (size_t)m * epi.ldr + n);
          v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
        }
        *reinterpret_cast<float4*>(C + (size_t)m * ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n + j >= N) break;
          float x = v[j];
          if (epi.bias) x += __ldg(epi.bias + n + j);
          if (brow) x += __ldg(brow + n + j);
          if (epi.act == 1) x = gelu_erf(x);
          if (epi.R) x += epi.R[(size_t)m * epi.ldr + n + j];
          C[(size_t)m * ldc + n + j] = x;
        }
      }
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

int gemm_f32(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K,
             const Epilogue& epi, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return GATOR_OK;
  GATOR_REQUIRE(A && W && C, "gemm_f32: null operand");
  GATOR_REQUIRE(K > 0 && K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0, "gemm_f32: K=%d lda=%d ldw=%d must be multiples of 4", K, lda, ldw);
  GATOR_REQUIRE(aligned16(A) && aligned16(W), "gemm_f32: A/W must be 16-byte aligned");
  GATOR_REQUIRE(!(epi.bias_rows && !epi.conv3) || epi.bias_period > 0, "gemm_f32: bias_rows needs bias_period");
  int vec = (!epi.conv3 && ldc % 4 == 0 && aligned16(C) && (!epi.R || (epi.ldr % 4 == 0 && aligned16(epi.R))) &&
             (!epi.bias || aligned16(epi.bias)) && (!epi.bias_rows || (aligned16(epi.bias_rows) && N % 4 == 0)))
                ? 1 : 0;
  if (N <= 64) {
    dim3 grid(ceil_div(M, 128), ceil_div(N, 64));
    gemm_f32_kernel<128, 64, 4><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, epi, vec);
  } else {
    dim3 grid(ceil_div(M, 128), ceil_div(N, 128));
    gemm_f32_kernel<128, 128, 8><<<grid, 256, 0, stream>>>(A, lda, W, ldw, C, ldc, M, N, K, epi, vec);
  }
  return check_launch("gemm_f32");
}

}  // namespace gator

extern "C" int gator_gemm(const gator_gemm_args* a, void* stream) {
  using namespace gator;
  GATOR_REQUIRE(a, "gator_gemm: null args");
  Epilogue e;
  e.bias = a->bias;
  e.bias_rows = a->bias_rows;
  e.bias_period = a->bias_period;
  e.act = a->act;
  e.R = a->R;
  e.ldr = a->ldr;
  if (a->precision == GATOR_PREC_BF16)   // W = bf16 weights packed by gator_b200.packing.pack_umma_weight
    return gemm_bf16_umma(a->A, a->lda, a->W, nullptr, a->C, a->ldc, a->M, a->N, a->K, e, (cudaStream_t)stream);
  if (a->precision == GATOR_PREC_BF16X3 && a->W_wide)   // tile-major split image (pack_umma_wide) + caller workspace
    return gemm_bf16x3_wide(a->A, a->lda, a->W_wide, a->a_image, a->a_image_bytes, a->C, a->ldc, a->M, a->N, a->K, e,
                            (cudaStream_t)stream);
  if (a->precision == GATOR_PREC_BF16X3) {
    GATOR_REQUIRE(a->W_lo, "gator_gemm: GATOR_PREC_BF16X3 needs W_lo");
    return gemm_bf16_umma(a->A, a->lda, a->W, a->W_lo, a->C, a->ldc, a->M, a->N, a->K, e, (cudaStream_t)stream);
  }
  GATOR_REQUIRE(a->precision == GATOR_PREC_FP32, "gator_gemm: bad precision");
  return gemm_f32(a->A, a->lda, static_cast<const float*>(a->W), a->ldw, a->C, a->ldc, a->M, a->N, a->K, e, (cudaStream_t)stream);
}
