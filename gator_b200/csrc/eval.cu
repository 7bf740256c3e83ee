// Fused evaluation epilogue.
//
// Every caller of the forward runs, per batch (lib/core/base.py:219-223):
//     pred_mesh, gt_mesh = pred_mesh * 1000, gt_mesh * 1000
//     pred_pose = J_regressor[None] @ pred_mesh                       (dense 17 x 6890 matmul)
//     j_err, s_err = dataset.compute_both_err(pred_mesh, gt_mesh, pred_pose, gt_pose3d)
// and compute_both_err (data/Human36M/dataset.py:466-478, data/PW3D/dataset.py:273-286) copies both meshes
// to the host to root-align them and take mean L2 distances in numpy; the final pass repeats it per sample with
// a Procrustes alignment (dataset.py:480-504, lib/coord_utils.py:127-149).
//
// One CTA per sample does all of it with a single read of the two meshes: sparse J-regression (the shipped
// regressors have ~6 non-zeros per row), metres -> mm, root alignment, per-sample MPJPE / MPVPE.  HBM-bound:
// 2 x 82 680 B read per sample, a few hundred bytes written.  On request eval_pa_kernel adds PA-MPJPE (3x3
// one-sided Jacobi SVD in fp64, one thread per sample), and eval_mean_kernel reduces the per-sample values in a
// fixed order.
#include "common.cuh"

namespace gator {
namespace {

constexpr int kMaxJoints = 32;
constexpr int kWarpTile = 64;        // vertices per warp tile: 768 B of each mesh
constexpr int kStages = 3;           // cp.async ring depth per warp (8 warps x 3 x 1536 B = 36 KB per CTA)

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];       // same order in every thread
  return t;
}

// Similarity Procrustes of A (n x 3) onto B, mean distance after alignment (lib/coord_utils.py:127-149).
// Points are given relative to arbitrary origins; everything in fp64 by one thread.
__device__ double procrustes_error(const float (*A)[3], const float (*Bp)[3], int n) {
  double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c) { ca[c] += A[i][c]; cb[c] += Bp[i][c]; }
  for (int c = 0; c < 3; ++c) { ca[c] /= n; cb[c] /= n; }
  double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};      // H = (A - ca)^T (B - cb) / n
  double var_a = 0;
  for (int i = 0; i < n; ++i) {
    double a[3], b[3];
    for (int c = 0; c < 3; ++c) { a[c] = A[i][c] - ca[c]; b[c] = Bp[i][c] - cb[c]; var_a += a[c] * a[c]; }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) G[r][c] += a[r] * b[c];
  }
  var_a /= n;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) G[r][c] /= n;
  // one-sided Jacobi: rotate column pairs of G until orthogonal; H V = U diag(s)
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double al = 0, be = 0, ga = 0;
        for (int i = 0; i < 3; ++i) { al += G[i][p] * G[i][p]; be += G[i][q] * G[i][q]; ga += G[i][p] * G[i][q]; }
        if (fabs(ga) <= 1e-15 * sqrt(al * be) || ga == 0.0) continue;
        rotated = true;
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
        for (int i = 0; i < 3; ++i) {
          const double gp = G[i][p], gq = G[i][q];
          G[i][p] = cs * gp - sn * gq; G[i][q] = sn * gp + cs * gq;
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = cs * vp - sn * vq; V[i][q] = sn * vp + cs * vq;
        }
      }
    if (!rotated) break;
  }
  double s[3];
  int kmin = 0;
  for (int k = 0; k < 3; ++k) {
    s[k] = sqrt(G[0][k] * G[0][k] + G[1][k] * G[1][k] + G[2][k] * G[2][k]);
    if (s[k] < s[kmin]) kmin = k;
  }
  double U[3][3];
  const int k1 = (kmin + 1) % 3, k2 = (kmin + 2) % 3;
  for (int k = 0; k < 3; ++k)
    for (int i = 0; i < 3; ++i) U[i][k] = s[k] > 0 ? G[i][k] / s[k] : 0.0;
  const double smax = fmax(s[k1], s[k2]);
  if (!(s[kmin] > 1e-12 * smax)) {
    // rank-deficient H: complete U with the cross product so that (u_k1, u_k2, u_kmin) is right-handed
    U[0][kmin] = U[1][k1] * U[2][k2] - U[2][k1] * U[1][k2];
    U[1][kmin] = U[2][k1] * U[0][k2] - U[0][k1] * U[2][k2];
    U[2][kmin] = U[0][k1] * U[1][k2] - U[1][k1] * U[0][k2];
  }
  auto det3 = [](const double (*M)[3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
  };
  // R = V U^T; a reflection is repaired on the smallest singular value (coord_utils.py:135-138)
  if (det3(V) * det3(U) < 0) {
    s[kmin] = -s[kmin];
    for (int i = 0; i < 3; ++i) V[i][kmin] = -V[i][kmin];
  }
  double R[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[r][c] = V[r][0] * U[c][0] + V[r][1] * U[c][1] + V[r][2] * U[c][2];
  const double scale = (s[0] + s[1] + s[2]) / var_a;
  double tr[3];
  for (int r = 0; r < 3; ++r) tr[r] = cb[r] - scale * (R[r][0] * ca[0] + R[r][1] * ca[1] + R[r][2] * ca[2]);
  double err = 0;
  for (int i = 0; i < n; ++i) {
    double d2 = 0;
    for (int r = 0; r < 3; ++r) {
      const double a2 = scale * (R[r][0] * A[i][0] + R[r][1] * A[i][1] + R[r][2] * A[i][2]) + tr[r];
      const double d = a2 - Bp[i][r];
      d2 += d * d;
    }
    err += sqrt(d2);
  }
  return err / n;
}

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

__global__ void __launch_bounds__(256, 6)
eval_sample_kernel(gator_eval_args a) {
  __shared__ float pj[kMaxJoints][3], gj[kMaxJoints][3];
  __shared__ float red[8];
  // per-warp cp.async ring: [warp][stage][pred | gt][kWarpTile * 3] - no block-wide barrier in the streaming loop
  __shared__ __align__(16) float ring[8][kStages][2][kWarpTile * 3 + 4];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = a.verts, J = a.joints;
  const float* pm = a.pred_mesh + (size_t)b * V * 3;
  const float scale = a.scale;
  const bool surface = a.gt_mesh && a.surface_err;
  const float* gm = surface ? a.gt_mesh + (size_t)b * V * 3 : nullptr;
  const int n_tiles = (V + kWarpTile - 1) / kWarpTile;
  const bool aligned = surface && ((((uintptr_t)pm | (uintptr_t)gm) & 7u) == 0) && (V & 1) == 0;

  // Shared-memory image of a tile keeps the global 16-byte phase (a sample starts 0 or 8 bytes past a 16-byte
  // boundary: 82 680 = 8 mod 16), so whole 16-byte chunks can be copied and only the two ends are 8-byte copies.
  const int op = aligned ? (int)(((uintptr_t)pm >> 2) & 3u) : 0;          // floats past the 16-byte boundary (0 or 2)
  const int og = aligned ? (int)(((uintptr_t)gm >> 2) & 3u) : 0;
  auto copy_tile = [&](float* dst, const float* src, int o, int nf) {
    // chunk c covers tile floats [4c - o, 4c - o + 4) clipped to [0, nf)
    for (int c = lane; c * 4 < o + nf; c += 32) {
      const int lo = max(4 * c - o, 0), hi = min(4 * c - o + 4, nf);
      if (hi - lo == 4) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(dst + o + lo)), "l"(src + lo) : "memory");
      } else {
        cp_async8(dst + o + lo, src + lo);                                  // o, nf even => exactly 2 floats
      }
    }
  };
  // tile `k` of this warp (global tile warp + 8k) into ring stage k % kStages; always commits one group
  auto issue = [&](int k) {
    const int tile = warp + 8 * k;
    if (surface && tile < n_tiles) {
      const int v0 = tile * kWarpTile;
      const int nf = min(kWarpTile, V - v0) * 3;
      const float* ps = pm + (size_t)v0 * 3;
      const float* gs = gm + (size_t)v0 * 3;
      float* dp = ring[warp][k % kStages][0];
      float* dg = ring[warp][k % kStages][1];
      if (aligned) {
        copy_tile(dp, ps, op, nf);
        copy_tile(dg, gs, og, nf);
      } else {
        for (int i = lane; i < nf; i += 32) { cp_async4(dp + i, ps + i); cp_async4(dg + i, gs + i); }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  // the first tiles stream in while the joints are regressed
#pragma unroll
  for (int k = 0; k < kStages - 1; ++k) issue(k);

  // 1. joints: sparse regression of the predicted mesh (base.py:219-221) or the caller's own joints
  if (a.pred_joints_in) {
    for (int i = tid; i < J * 3; i += 256) pj[i / 3][i % 3] = a.pred_joints_in[(size_t)b * J * 3 + i];
  } else {
    for (int j = warp; j < J; j += 8) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      for (int k = __ldg(a.jreg_rowptr + j) + lane; k < __ldg(a.jreg_rowptr + j + 1); k += 32) {
        const float w = __ldg(a.jreg_values + k);
        const float* xp = pm + (size_t)__ldg(a.jreg_colidx + k) * 3;
        a0 = fmaf(w, __fmul_rn(xp[0], scale), a0);
        a1 = fmaf(w, __fmul_rn(xp[1], scale), a1);
        a2 = fmaf(w, __fmul_rn(xp[2], scale), a2);
      }
      a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
      if (lane == 0) { pj[j][0] = a0; pj[j][1] = a1; pj[j][2] = a2; }
    }
  }
  for (int i = tid; i < J * 3; i += 256) gj[i / 3][i % 3] = a.gt_joints[(size_t)b * J * 3 + i];
  __syncthreads();
  if (a.pred_joints)
    for (int i = tid; i < J * 3; i += 256) a.pred_joints[(size_t)b * J * 3 + i] = pj[i / 3][i % 3];
  const float rp0 = pj[a.root][0], rp1 = pj[a.root][1], rp2 = pj[a.root][2];
  const float rg0 = gj[a.root][0], rg1 = gj[a.root][1], rg2 = gj[a.root][2];

  // 2. surface error: both meshes streamed exactly once
  if (surface) {
    float acc = 0.f;
    for (int k = 0; warp + 8 * k < n_tiles; ++k) {
      issue(k + kStages - 1);                                  // refills the stage consumed in iteration k - 1
      asm volatile("cp.async.wait_group %0;\n" ::"n"(kStages - 1) : "memory");
      __syncwarp();
      const int nv = min(kWarpTile, V - (warp + 8 * k) * kWarpTile);
      const float* sp = ring[warp][k % kStages][0] + op;
      const float* sg = ring[warp][k % kStages][1] + og;
#pragma unroll
      for (int u = 0; u < kWarpTile / 32; ++u) {
        const int v = lane + 32 * u;
        if (v < nv) {
          // same operation order as the reference: x*1000, minus root, difference, squares, sqrt
          const float dx = __fsub_rn(__fsub_rn(__fmul_rn(sp[v * 3], scale), rp0), __fsub_rn(__fmul_rn(sg[v * 3], scale), rg0));
          const float dy = __fsub_rn(__fsub_rn(__fmul_rn(sp[v * 3 + 1], scale), rp1), __fsub_rn(__fmul_rn(sg[v * 3 + 1], scale), rg1));
          const float dz = __fsub_rn(__fsub_rn(__fmul_rn(sp[v * 3 + 2], scale), rp2), __fsub_rn(__fmul_rn(sg[v * 3 + 2], scale), rg2));
          acc += sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        }
      }
      __syncwarp();
    }
    const float tot = block_sum(acc, red);
    if (tid == 0) a.surface_err[b] = tot / (float)V;
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");

  // 3. joint error on the evaluation joints, root-aligned
  if (tid < 32) {
    float e = 0.f;
    if (lane < a.n_eval) {
      const int j = __ldg(a.eval_joints + lane);
      const float p0 = __fsub_rn(pj[j][0], rp0), p1 = __fsub_rn(pj[j][1], rp1), p2 = __fsub_rn(pj[j][2], rp2);
      const float g0 = __fsub_rn(gj[j][0], rg0), g1 = __fsub_rn(gj[j][1], rg1), g2 = __fsub_rn(gj[j][2], rg2);
      const float dx = __fsub_rn(p0, g0), dy = __fsub_rn(p1, g1), dz = __fsub_rn(p2, g2);
      e = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    }
    e = warp_sum(e);
    if (lane == 0 && a.joint_err) a.joint_err[b] = e / (float)a.n_eval;
  }
}

// PA-MPJPE: one thread per sample on the regressed joints (kept out of eval_sample_kernel so that the fp64
// Jacobi's registers do not limit the occupancy of the bandwidth-bound mesh pass)
__global__ void __launch_bounds__(64)
eval_pa_kernel(const float* pred_joints, const float* gt_joints, const int32_t* eval_joints, int n_eval, int joints,
               int root, int batch, float* pa_out) {
  const int b = blockIdx.x * 64 + threadIdx.x;
  if (b >= batch) return;
  float ea[kMaxJoints][3], eb[kMaxJoints][3];
  const float* p = pred_joints + (size_t)b * joints * 3;
  const float* g = gt_joints + (size_t)b * joints * 3;
  for (int i = 0; i < n_eval; ++i) {
    const int j = __ldg(eval_joints + i);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ea[i][c] = __fsub_rn(p[j * 3 + c], p[root * 3 + c]);
      eb[i][c] = __fsub_rn(g[j * 3 + c], g[root * 3 + c]);
    }
  }
  pa_out[b] = (float)procrustes_error(ea, eb, n_eval);
}

// mean over the batch of up to three per-sample arrays, fixed summation order, fp64 accumulation
__global__ void __launch_bounds__(256)
eval_mean_kernel(const float* j, const float* s, const float* pa, int n, float* out) {
  __shared__ double red[256];
  const float* src[3] = {j, s, pa};
  for (int w = 0; w < 3; ++w) {
    double acc = 0;
    if (src[w])
      for (int i = threadIdx.x; i < n; i += 256) acc += src[w][i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[w] = src[w] ? (float)(red[0] / n) : 0.f;
    __syncthreads();
  }
}

}  // namespace
}  // namespace gator

extern "C" int gator_eval_epilogue(const gator_eval_args* a, void* stream) {
  using namespace gator;
  GATOR_REQUIRE(a, "gator_eval_epilogue: null args");
  GATOR_REQUIRE(a->batch >= 0 && a->verts > 0 && a->joints > 0 && a->joints <= kMaxJoints,
                "gator_eval_epilogue: bad shape (joints <= %d)", kMaxJoints);
  GATOR_REQUIRE(a->n_eval > 0 && a->n_eval <= a->joints && a->root >= 0 && a->root < a->joints,
                "gator_eval_epilogue: bad eval joint set / root");
  if (a->batch == 0) return GATOR_OK;
  GATOR_REQUIRE(a->pred_mesh && a->gt_joints && a->eval_joints, "gator_eval_epilogue: null buffer");
  GATOR_REQUIRE(a->pred_joints_in || (a->jreg_rowptr && a->jreg_colidx && a->jreg_values),
                "gator_eval_epilogue: need a CSR regressor or pred_joints_in");
  GATOR_REQUIRE(!a->pa_joint_err || a->pred_joints || a->pred_joints_in,
                "gator_eval_epilogue: pa_joint_err needs pred_joints (or pred_joints_in)");
  eval_sample_kernel<<<a->batch, 256, 0, (cudaStream_t)stream>>>(*a);
  GATOR_TRY(check_launch("eval_sample"));
  if (a->pa_joint_err) {
    eval_pa_kernel<<<ceil_div(a->batch, 64), 64, 0, (cudaStream_t)stream>>>(
        a->pred_joints_in ? a->pred_joints_in : a->pred_joints, a->gt_joints, a->eval_joints, a->n_eval, a->joints,
        a->root, a->batch, a->pa_joint_err);
    GATOR_TRY(check_launch("eval_pa"));
  }
  if (a->batch_mean) {
    eval_mean_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a->joint_err, a->surface_err && a->gt_mesh ? a->surface_err : nullptr,
                                                          a->pa_joint_err, a->batch, a->batch_mean);
    GATOR_TRY(check_launch("eval_mean"));
  }
  return GATOR_OK;
}
