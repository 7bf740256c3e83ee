// Row LayerNorm kernels: nn.LayerNorm (GAT.py:37,41,148; MDR.py:65,68) and the unbiased-std variant of
// lib/models/vanilla_transformer_encoder.py:31-34.  One warp per row, two-pass statistics in registers.
#include "common.cuh"

namespace gator {
namespace {

template <int C>   // 64 or 128
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w,
                      const float* __restrict__ b, int rows, int mode, int gelu) {
  constexpr int PER = C / 32;   // 2 or 4 contiguous floats per lane
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * C + lane * PER;
  float v[PER];
  if (PER == 4) {
    float4 t = *reinterpret_cast<const float4*>(xr);
    v[0] = t.x; v[1] = t.y; v[PER - 2] = t.z; v[PER - 1] = t.w;
  } else {
    float2 t = *reinterpret_cast<const float2*>(xr);
    v[0] = t.x; v[1] = t.y;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { float d = v[i] - mean; q = fmaf(d, d, q); }
  q = warp_sum(q);
  float o[PER];
  if (mode == 0) {
    const float rstd = rsqrtf(q * (1.0f / C) + 1e-5f);
#pragma unroll
    for (int i = 0; i < PER; ++i) o[i] = (v[i] - mean) * rstd * __ldg(w + lane * PER + i) + __ldg(b + lane * PER + i);
  } else {
    const float denom = sqrtf(q * (1.0f / (C - 1))) + 1e-6f;
#pragma unroll
    for (int i = 0; i < PER; ++i) o[i] = __ldg(w + lane * PER + i) * (v[i] - mean) / denom + __ldg(b + lane * PER + i);
  }
  if (gelu) {
#pragma unroll
    for (int i = 0; i < PER; ++i) o[i] = gelu_erf(o[i]);
  }
  float* yr = y + (size_t)row * C + lane * PER;
  if (PER == 4) *reinterpret_cast<float4*>(yr) = make_float4(o[0], o[1], o[PER - 2], o[PER - 1]);
  else *reinterpret_cast<float2*>(yr) = make_float2(o[0], o[1]);
}

}  // namespace

int layernorm_rows(const float* x, float* y, const float* w, const float* b, int rows, int C, int mode,
                   int gelu, cudaStream_t stream) {
  if (rows <= 0) return GATOR_OK;
  GATOR_REQUIRE(C == 64 || C == 128, "layernorm_rows: C=%d unsupported", C);
  const int warps = 8;
  dim3 grid(ceil_div(rows, warps));
  if (C == 64) layernorm_rows_kernel<64><<<grid, warps * 32, 0, stream>>>(x, y, w, b, rows, mode, gelu);
  else layernorm_rows_kernel<128><<<grid, warps * 32, 0, stream>>>(x, y, w, b, rows, mode, gelu);
  return check_launch("layernorm_rows");
}

}  // namespace gator
