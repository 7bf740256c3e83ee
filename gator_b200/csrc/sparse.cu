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

// xyz fast path (feat == 3).  CTA = NS consecutive samples x all rows: the samples' inputs (cols x 3 floats each)
// are staged in shared memory with coalesced loads, so the random column gathers hit shared memory banks instead
// of touching up to 32 L1 lines per instruction; lane = output row, its <= 4 non-zeros are fetched once per
// row and reused for the NS samples; a warp's 32 rows x 3 floats are staged in warp-private shared memory and
// leave as 256-byte coalesced store instructions (8-byte accesses when the segment is 8-byte aligned).
__global__ void __launch_bounds__(256)
csr_spmm3_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, const float* __restrict__ values,
                 const float* __restrict__ x, float* __restrict__ y, int rows, int cols, float scale, int batch, int NS) {
  extern __shared__ __align__(16) float sx[];          // [NS][cols*3]
  __shared__ __align__(16) float stage[8][96];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * NS;
  const int ns = min(NS, batch - b0);
  const int xs = cols * 3;
  {
    const float* src = x + (size_t)b0 * xs;
    const int total = ns * xs;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
      for (int i = tid; i < total / 4; i += 256) reinterpret_cast<float4*>(sx)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
      for (int i = (total / 4) * 4 + tid; i < total; i += 256) sx[i] = src[i];
    } else {
      for (int i = tid; i < total; i += 256) sx[i] = src[i];
    }
  }
  __syncthreads();
  float* st = stage[warp];
  for (int r0 = warp * 32; r0 < rows; r0 += 256) {
    const int nr = min(32, rows - r0);
    const int r = r0 + lane;
    const bool active = lane < nr;
    int k0 = 0, k1 = 0;
    if (active) { k0 = __ldg(rowptr + r); k1 = __ldg(rowptr + r + 1); }
    const int nnz = k1 - k0;
    int c[4] = {0, 0, 0, 0};
    float w[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < nnz) { c[k] = __ldg(colidx + k0 + k) * 3; w[k] = __ldg(values + k0 + k); }
    const int seg = nr * 3;
    for (int s = 0; s < ns; ++s) {
      const float* xb = sx + s * xs;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      if (nnz <= 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < nnz) {
            a0 = fmaf(w[k], xb[c[k]], a0); a1 = fmaf(w[k], xb[c[k] + 1], a1); a2 = fmaf(w[k], xb[c[k] + 2], a2);
          }
        }
      } else {
        for (int k = k0; k < k1; ++k) {
          const float ww = __ldg(values + k);
          const float* xp = xb + __ldg(colidx + k) * 3;
          a0 = fmaf(ww, xp[0], a0); a1 = fmaf(ww, xp[1], a1); a2 = fmaf(ww, xp[2], a2);
        }
      }
      __syncwarp();
      if (active) { st[lane * 3] = a0 * scale; st[lane * 3 + 1] = a1 * scale; st[lane * 3 + 2] = a2 * scale; }
      __syncwarp();
      float* dst = y + ((size_t)(b0 + s) * rows + r0) * 3;
      if (((reinterpret_cast<uintptr_t>(dst) & 7u) == 0) && (seg & 1) == 0) {
        const int seg2 = seg >> 1;
        if (lane < seg2) reinterpret_cast<float2*>(dst)[lane] = reinterpret_cast<const float2*>(st)[lane];
        if (lane + 32 < seg2) reinterpret_cast<float2*>(dst)[lane + 32] = reinterpret_cast<const float2*>(st)[lane + 32];
      } else {
        for (int i = lane; i < seg; i += 32) dst[i] = st[i];
      }
    }
  }
}


// xyz fast path for operators whose input fits shared memory twice (Mesh.upsample: 431 -> 1723 -> 6890).
// "Row-stationary": a CTA owns a slice of 1024 output rows for a strided set of samples; every lane keeps the <= 4
// non-zeros of its 4 rows in registers for the whole kernel, so per sample only the input (cols x 3 floats, copied
// with cp.async into one of two shared-memory buffers while the previous sample is processed) and the outputs move:
// 12 KB of contiguous output per CTA and sample, written as 128-byte coalesced stores through warp-private staging.
constexpr int RS_ROWS = 1024;          // output rows per CTA: 8 warps x 4 groups x 32 lanes
__global__ void __launch_bounds__(256)
csr_spmm3_rows_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, const float* __restrict__ values,
                      const float* __restrict__ x, float* __restrict__ y, int rows, int cols, float scale, int batch,
                      int slices, int parts) {
  extern __shared__ __align__(16) float sx[];          // [2][cols*3]
  __shared__ __align__(16) float stage[8][4][96];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x % slices, part = blockIdx.x / slices;
  const int xs = cols * 3;
  const int v0 = slice * RS_ROWS + warp * 128;         // first row of this warp
  int c[4][4], k0[4], nnz[4];
  float w[4][4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int r = v0 + g * 32 + lane;
    k0[g] = 0; nnz[g] = 0;
    if (r < rows) { k0[g] = __ldg(rowptr + r); nnz[g] = __ldg(rowptr + r + 1) - k0[g]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      c[g][k] = 0; w[g][k] = 0.f;
      if (k < nnz[g]) { c[g][k] = __ldg(colidx + k0[g] + k) * 3; w[g][k] = __ldg(values + k0[g] + k); }
    }
  }
  // sample s -> sx[buf].  A sample starts 4-byte aligned (cols * 12 B is not a multiple of 16), so the enclosing 16-byte
  // aligned range is copied with 16-byte cp.async and read at an offset of `shift` floats; the last sample, whose
  // enclosing range could run past the end of x, is copied in 4-byte pieces.
  const int xbuf = (xs + 6) & ~3;                      // floats per buffer: room for shift <= 3, a multiple of 16 bytes
  auto copy_in = [&](int s, int buf) -> int {
    const float* src = x + (size_t)s * xs;
    const int shift = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sx + buf * xbuf);
    if (s + 1 < batch) {
      const float* a0 = src - shift;
      const int n16 = (shift + xs + 3) >> 2;
      for (int i = tid; i < n16; i += 256)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + i * 16), "l"(a0 + i * 4) : "memory");
    } else {
      for (int i = tid; i < xs; i += 256)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst + (shift + i) * 4), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    return shift;
  };
  int shift_cur = 0, shift_next = 0;
  if (part < batch) shift_cur = copy_in(part, 0);
  int it = 0;
  for (int s = part; s < batch; s += parts, ++it) {
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();                                   // sample s has landed; everyone is done with the other buffer
    if (s + parts < batch) shift_next = copy_in(s + parts, (it + 1) & 1);
    const float* xb = sx + (it & 1) * xbuf + shift_cur;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      if (nnz[g] <= 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < nnz[g]) {
            a0 = fmaf(w[g][k], xb[c[g][k]], a0); a1 = fmaf(w[g][k], xb[c[g][k] + 1], a1); a2 = fmaf(w[g][k], xb[c[g][k] + 2], a2);
          }
        }
      } else {
        for (int k = k0[g]; k < k0[g] + nnz[g]; ++k) {
          const float ww = __ldg(values + k);
          const float* xp = xb + __ldg(colidx + k) * 3;
          a0 = fmaf(ww, xp[0], a0); a1 = fmaf(ww, xp[1], a1); a2 = fmaf(ww, xp[2], a2);
        }
      }
      float* st = stage[warp][g];
      st[lane * 3] = a0 * scale; st[lane * 3 + 1] = a1 * scale; st[lane * 3 + 2] = a2 * scale;
    }
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int r0 = v0 + g * 32;
      const int seg = max(0, min(32, rows - r0)) * 3;
      const float* st = stage[warp][g];
      float* dst = y + ((size_t)s * rows + r0) * 3;
      if (lane < seg) dst[lane] = st[lane];
      if (lane + 32 < seg) dst[lane + 32] = st[lane + 32];
      if (lane + 64 < seg) dst[lane + 64] = st[lane + 64];
    }
    __syncwarp();
    shift_cur = shift_next;
  }
}


// Mesh.upsample through two operators in one launch (mesh.py:110-123 with n1=2, n2=0: 431 -> 1723 -> 6890).
// CTA = 4 consecutive samples, two CTAs per SM (one streams outputs while the other stages / runs level 1): the
// inputs (cols x 3 floats per sample) and the intermediate level (rows1 x 3) live in shared memory, so HBM sees the
// 5 KB input and the 83 KB output of a sample once.
//  * lane = one output FLOAT position (row r = f / 3, component c = f % 3), not one row: a warp's stores are 128
//    contiguous bytes per sample straight from registers (no staging pass through shared memory);
//  * the four samples of the group are interleaved in shared memory (float4 per float position), so ONE 16-byte
//    shared-memory load per non-zero serves all four samples: 3 LDS + 12 FMA + 4 STG per position;
//  * operators are ELL records of four (col, val) per row - two 16-byte loads per position and group, L2-resident.
// The first version of this kernel (scalar ELL loads, per-sample 4-byte gathers) was issue-bound: 87 instructions per
// warp-level output, 289 us at B = 4096 (profiles/r02_ncu_mesh_upsample2.txt); this one issues ~11.
constexpr int UP_NS = 4;
template <int W1, int W2>
__global__ void __launch_bounds__(512, 2)
mesh_upsample2_kernel(const int4* __restrict__ col1, const float4* __restrict__ val1, const int4* __restrict__ col2,
                      const float4* __restrict__ val2, const float* __restrict__ x, float* __restrict__ y, int cols,
                      int rows1, int rows2, float scale, int batch, int groups) {
  extern __shared__ __align__(16) float4 sm4[];
  const int xs0 = cols * 3, xs1 = rows1 * 3, xs2 = rows2 * 3;
  float4* sx0 = sm4;                  // [xs0] x 4 samples
  float4* sx1 = sm4 + xs0;            // [xs1] x 4 samples
  const int tid = threadIdx.x;
  auto gather = [](const float4* __restrict__ src, const int4 cc, const float4 ww, int c, int W) -> float4 {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    const int id[4] = {cc.x * 3 + c, cc.y * 3 + c, cc.z * 3 + c, cc.w * 3 + c};
    const float w[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < W) {
        const float4 v = src[id[k]];
        a.x = fmaf(w[k], v.x, a.x); a.y = fmaf(w[k], v.y, a.y); a.z = fmaf(w[k], v.z, a.z); a.w = fmaf(w[k], v.w, a.w);
      }
    }
    return a;
  };
  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    const int b0 = g * UP_NS;
    const int ns = min(UP_NS, batch - b0);
    __syncthreads();                  // the previous group's level 2 has finished reading sx1
    {
      // all of a thread's loads are issued before the first store (12 in flight per thread)
      const float* src = x + (size_t)b0 * xs0;
      float* dst = reinterpret_cast<float*>(sx0);
      for (int i0 = 0; i0 < xs0; i0 += 3 * 512) {
        float v[UP_NS][3];
#pragma unroll
        for (int s = 0; s < UP_NS; ++s)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int i = i0 + j * 512 + tid;
            v[s][j] = (s < ns && i < xs0) ? __ldg(src + s * xs0 + i) : 0.f;
          }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int i = i0 + j * 512 + tid;
          if (i < xs0) sx0[i] = make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
        }
      }
    }
    __syncthreads();
    for (int f = tid; f < xs1; f += 512) {
      const int r = f / 3, c = f - r * 3;
      sx1[f] = gather(sx0, __ldg(col1 + r), __ldg(val1 + r), c, W1);
    }
    __syncthreads();
    float* y0 = y + (size_t)b0 * xs2;
    if (ns == UP_NS) {
      // the operator record of the next position is requested before the current one is processed
      int4 cc = make_int4(0, 0, 0, 0);
      float4 ww = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tid < xs2) { cc = __ldg(col2 + tid / 3); ww = __ldg(val2 + tid / 3); }
      for (int f = tid; f < xs2; f += 512) {
        const int r = f / 3, c = f - r * 3;
        const int4 cc0 = cc;
        const float4 ww0 = ww;
        if (f + 512 < xs2) { cc = __ldg(col2 + (f + 512) / 3); ww = __ldg(val2 + (f + 512) / 3); }
        const float4 a = gather(sx1, cc0, ww0, c, W2);
        float* yo = y0 + f;
        yo[0] = a.x * scale; yo[xs2] = a.y * scale; yo[2 * (size_t)xs2] = a.z * scale; yo[3 * (size_t)xs2] = a.w * scale;
      }
    } else {
      for (int f = tid; f < xs2; f += 512) {
        const int r = f / 3, c = f - r * 3;
        const float4 a = gather(sx1, __ldg(col2 + r), __ldg(val2 + r), c, W2);
        float* yo = y0 + f;
        yo[0] = a.x * scale;
        if (ns > 1) yo[xs2] = a.y * scale;
        if (ns > 2) yo[2 * (size_t)xs2] = a.z * scale;
      }
    }
  }
}

template <int W1, int W2>
int launch_upsample2(const gator_upsample2_args* a, int slots, cudaStream_t stream) {
  static DeviceOnce attr_once;
  GATOR_TRY(attr_once.run("mesh_upsample2", [&](int) -> cudaError_t {
    return cudaFuncSetAttribute(mesh_upsample2_kernel<W1, W2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  }));
  const int groups = ceil_div(a->batch, UP_NS);
  const size_t smem = (size_t)UP_NS * (a->cols + a->rows1) * 12;
  mesh_upsample2_kernel<W1, W2><<<groups < slots ? groups : slots, 512, smem, stream>>>(
      reinterpret_cast<const int4*>(a->col1), reinterpret_cast<const float4*>(a->val1), reinterpret_cast<const int4*>(a->col2),
      reinterpret_cast<const float4*>(a->val2), a->x, a->y, a->cols, a->rows1, a->rows2, a->scale, a->batch, groups);
  return check_launch("mesh_upsample2");
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
  if (a->feat == 3 && a->rows >= 256 && (size_t)a->cols * 24 + 64 <= 48 * 1024 && (reinterpret_cast<uintptr_t>(a->x) & 15u) == 0) {
    // row-stationary kernel: input double-buffered in shared memory, up to 4 CTAs per SM
    const int slices = ceil_div(a->rows, RS_ROWS);
    int parts = (148 * 4) / slices;
    parts = parts < 1 ? 1 : (parts > a->batch ? a->batch : parts);
    static DeviceOnce attr_once2;
    GATOR_TRY(attr_once2.run("csr_spmm3_rows", [&](int) -> cudaError_t {
      return cudaFuncSetAttribute(csr_spmm3_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    }));
    csr_spmm3_rows_kernel<<<slices * parts, 256, (size_t)a->cols * 24 + 64, (cudaStream_t)stream>>>(
        a->rowptr, a->colidx, a->values, a->x, a->y, a->rows, a->cols, a->scale, a->batch, slices, parts);
    return check_launch("csr_spmm3_rows");
  }
  if (a->feat == 3 && a->rows >= 64 && (size_t)a->cols * 12 <= 96 * 1024) {
    const int per_sample = a->cols * 12;
    int NS = 96 * 1024 / per_sample;                 // <= 96 KB of staged inputs per CTA (2 CTAs / SM)
    NS = NS > 8 ? 8 : NS;
    static DeviceOnce attr_once;
    GATOR_TRY(attr_once.run("csr_spmm3", [&](int) -> cudaError_t {
      return cudaFuncSetAttribute(csr_spmm3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    }));
    csr_spmm3_kernel<<<ceil_div(a->batch, NS), 256, (size_t)NS * per_sample, (cudaStream_t)stream>>>(
        a->rowptr, a->colidx, a->values, a->x, a->y, a->rows, a->cols, a->scale, a->batch, NS);
    return check_launch("csr_spmm3");
  }
  const int vec = (reinterpret_cast<uintptr_t>(a->y) & 15u) == 0;
  const long long quads = (total + 3) / 4;
  GATOR_REQUIRE(quads / 256 + 1 < 0x7fffffffLL, "gator_csr_spmm: too large");
  csr_spmm_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      a->rowptr, a->colidx, a->values, a->x, a->y, a->rows, a->cols, a->feat, a->scale, total, vec);
  return check_launch("csr_spmm");
}

extern "C" int gator_mesh_upsample2(const gator_upsample2_args* a, void* stream_) {
  using namespace gator;
  cudaStream_t stream = (cudaStream_t)stream_;
  GATOR_REQUIRE(a, "gator_mesh_upsample2: null args");
  GATOR_REQUIRE(a->batch >= 0 && a->cols > 0 && a->rows1 > 0 && a->rows2 > 0, "gator_mesh_upsample2: bad shape");
  GATOR_REQUIRE(a->width1 >= 1 && a->width1 <= 4 && a->width2 >= 1 && a->width2 <= 4, "gator_mesh_upsample2: ELL width must be 1..4");
  if (a->batch == 0) return GATOR_OK;
  GATOR_REQUIRE(a->col1 && a->val1 && a->col2 && a->val2 && a->x && a->y, "gator_mesh_upsample2: null buffer");
  GATOR_REQUIRE(((reinterpret_cast<uintptr_t>(a->col1) | reinterpret_cast<uintptr_t>(a->val1) | reinterpret_cast<uintptr_t>(a->col2) |
                  reinterpret_cast<uintptr_t>(a->val2)) & 15u) == 0, "gator_mesh_upsample2: ELL arrays must be 16-byte aligned");
  GATOR_REQUIRE((size_t)(a->cols + a->rows1) * 12 * UP_NS <= 112 * 1024,
                "gator_mesh_upsample2: levels do not fit shared memory (use gator_csr_spmm twice)");
  static int sm_count[64];
  static DeviceOnce sm_once;
  int dev = 0;
  cudaGetDevice(&dev);
  GATOR_TRY(sm_once.run("mesh_upsample2 (SM count)", [&](int d) -> cudaError_t {
    return cudaDeviceGetAttribute(&sm_count[d & 63], cudaDevAttrMultiProcessorCount, d);
  }));
  const int slots = 2 * (sm_count[dev & 63] > 0 ? sm_count[dev & 63] : 148);
  // the records always hold four entries; widths <= 3 skip the fourth (its weight is 0)
  const bool w1 = a->width1 > 3, w2 = a->width2 > 3;
  if (w1) return w2 ? launch_upsample2<4, 4>(a, slots, stream) : launch_upsample2<4, 3>(a, slots, stream);
  return w2 ? launch_upsample2<3, 4>(a, slots, stream) : launch_upsample2<3, 3>(a, slots, stream);
}
