// Batched CSR SpMM over (B, N, F) vertex arrays: Mesh.downsample / Mesh.upsample
// (lib/models/backbones/mesh.py:93-123 via graph_layers.py:105-124) and the sparse J-regression every
// caller applies to the predicted mesh (lib/core/base.py:221, demo/run.py:142; the shipped regressors
// have ~6 non-zeros per row).  HBM-bound: each thread produces 4 consecutive output floats of the flat
// (B*rows*F) array (one aligned float4 store); the gathers hit L1/L2 (a sample's input is <= 83 KB).
#include "common.cuh"

namespace gator {
namespace {

__global__ void __launch_bounds__(256)
csr_spmm_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, const float* __restrict__ values,
                const float* __restrict__ x, float* __restrict__ y, int rows, int cols, int feat, float scale,
                long long total, int vec) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long e0 = q * 4;
  if (e0 >= total) return;
  const long long per_sample = (long long)rows * feat;
  float out[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long e = e0 + u;
    float acc = 0.f;
    if (e < total) {
      const long long b = e / per_sample;
      const int rem = (int)(e - b * per_sample);
      const int r = rem / feat, f = rem - r * feat;
      const float* xb = x + (size_t)b * cols * feat + f;
      const int k0 = __ldg(rowptr + r), k1 = __ldg(rowptr + r + 1);
      for (int k = k0; k < k1; ++k) acc = fmaf(__ldg(values + k), xb[(size_t)__ldg(colidx + k) * feat], acc);
      acc *= scale;
    }
    out[u] = acc;
  }
  if (vec && e0 + 3 < total) {
    *reinterpret_cast<float4*>(y + e0) = make_float4(out[0], out[1], out[2], out[3]);
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (e0 + u < total) y[e0 + u] = out[u];
  }
}

}  // namespace
}  // namespace gator

extern "C" int gator_csr_spmm(const gator_csr_args* a, void* stream) {
  using namespace gator;
  GATOR_REQUIRE(a, "gator_csr_spmm: null args");
  GATOR_REQUIRE(a->batch >= 0 && a->rows >= 0 && a->cols > 0 && a->feat > 0, "gator_csr_spmm: bad shape");
  const long long total = (long long)a->batch * a->rows * a->feat;
  if (total == 0) return GATOR_OK;
  GATOR_REQUIRE(a->rowptr && a->colidx && a->values && a->x && a->y, "gator_csr_spmm: null buffer");
  const int vec = (reinterpret_cast<uintptr_t>(a->y) & 15u) == 0;
  const long long quads = (total + 3) / 4;
  GATOR_REQUIRE(quads / 256 + 1 < 0x7fffffffLL, "gator_csr_spmm: too large");
  csr_spmm_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      a->rowptr, a->colidx, a->values, a->x, a->y, a->rows, a->cols, a->feat, a->scale, total, vec);
  return check_launch("csr_spmm");
}
