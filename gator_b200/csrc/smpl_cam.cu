// Camera fix-up of SMPL parameters for ground-truth mesh generation:
// the host arithmetic of Human36M.get_smpl_coord (data/Human36M/dataset.py:254-298) for a whole batch.
//
// Third-party arithmetic: transforms3d.axangles.axangle2mat / mat2axangle (unpinned in requirements.sh, absent
// from the reference tree).  Restated from its published algorithm: axangle2mat is the Rodrigues matrix
// [x*xC+c, xyC-zs, zxC+ys; ...]; mat2axangle takes the unit eigenvector of eigenvalue 1 as the axis and
// atan2(sin, cos) with cos = (trace-1)/2 as the angle - i.e. the rotation vector of the matrix, which is what is
// computed here in closed form (axis * angle does not depend on the eigenvector's sign).
//
// One thread per sample; a few hundred bytes each way - latency-bound, exists to keep the data path on the device.
#include "common.cuh"

namespace gator {
namespace {

// rotation vector (axis * angle, angle in [0, pi]) of a 3x3 rotation matrix, fp64
__device__ void rotation_log(const double M[9], double rv[3]) {
  const double vx = M[7] - M[5], vy = M[2] - M[6], vz = M[3] - M[1];          // 2 sin(a) axis
  const double n = sqrt(vx * vx + vy * vy + vz * vz);
  const double cosa = (M[0] + M[4] + M[8] - 1.0) * 0.5;
  const double ang = atan2(0.5 * n, cosa);
  if (n > 1e-6) {
    const double k = ang / n;
    rv[0] = vx * k; rv[1] = vy * k; rv[2] = vz * k;
  } else if (cosa > 0.0) {
    rv[0] = 0.5 * vx; rv[1] = 0.5 * vy; rv[2] = 0.5 * vz;                      // a -> 0: a / sin(a) -> 1
  } else {
    // a -> pi: M ~ 2 u u^T - I; take the column of (M + I) with the largest diagonal, sign from the skew part
    const double d0 = M[0] + 1.0, d1 = M[4] + 1.0, d2 = M[8] + 1.0;
    double ux, uy, uz;
    if (d0 >= d1 && d0 >= d2) { ux = d0; uy = 0.5 * (M[1] + M[3]); uz = 0.5 * (M[2] + M[6]); }
    else if (d1 >= d2) { ux = 0.5 * (M[1] + M[3]); uy = d1; uz = 0.5 * (M[5] + M[7]); }
    else { ux = 0.5 * (M[2] + M[6]); uy = 0.5 * (M[5] + M[7]); uz = d2; }
    double un = sqrt(ux * ux + uy * uy + uz * uz);
    if (ux * vx + uy * vy + uz * vz < 0.0) un = -un;
    rv[0] = ux / un * ang; rv[1] = uy / un * ang; rv[2] = uz / un * ang;
  }
}

__global__ void __launch_bounds__(128)
smpl_cam_fixup_kernel(gator_smpl_cam_args a) {
  const int b = blockIdx.x * 128 + threadIdx.x;
  if (b >= a.batch) return;
  // shape: back to the mean shape if any coefficient is implausible (dataset.py:266)
  float be[10];
  bool far = false, any = false;
#pragma unroll
  for (int k = 0; k < 10; ++k) { be[k] = a.betas[(size_t)b * 10 + k]; far |= fabsf(be[k]) > 3.f; }
#pragma unroll
  for (int k = 0; k < 10; ++k) { if (far) be[k] = 0.f; any |= be[k] != 0.f; a.betas_out[(size_t)b * 10 + k] = be[k]; }
  // pose: copy, then rotate the root orientation into the camera frame (dataset.py:268-274)
  const float* ps = a.pose + (size_t)b * 72;
  float* po = a.pose_out + (size_t)b * 72;
  for (int i = 3; i < 72; ++i) po[i] = ps[i];
  const float R[9] = {a.cam_R[(size_t)b * 9 + 0], a.cam_R[(size_t)b * 9 + 1], a.cam_R[(size_t)b * 9 + 2],
                      a.cam_R[(size_t)b * 9 + 3], a.cam_R[(size_t)b * 9 + 4], a.cam_R[(size_t)b * 9 + 5],
                      a.cam_R[(size_t)b * 9 + 6], a.cam_R[(size_t)b * 9 + 7], a.cam_R[(size_t)b * 9 + 8]};
  {
    // axangle2mat(root / |root|, |root|): float32 axis, trigonometry of the float32 angle in double, float32 entries
    const float rx = ps[0], ry = ps[1], rz = ps[2];
    const float angle = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
    float x = __fdiv_rn(rx, angle), y = __fdiv_rn(ry, angle), z = __fdiv_rn(rz, angle);
    const float n = (float)sqrt((double)__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    x = __fdiv_rn(x, n); y = __fdiv_rn(y, n); z = __fdiv_rn(z, n);
    const double cd = cos((double)angle), sd = sin((double)angle);
    const float c = (float)cd, s = (float)sd, C = (float)(1.0 - cd);
    const float xs = __fmul_rn(x, s), ys = __fmul_rn(y, s), zs = __fmul_rn(z, s);
    const float xC = __fmul_rn(x, C), yC = __fmul_rn(y, C), zC = __fmul_rn(z, C);
    const float xyC = __fmul_rn(x, yC), yzC = __fmul_rn(y, zC), zxC = __fmul_rn(z, xC);
    const float A[9] = {__fadd_rn(__fmul_rn(x, xC), c), __fsub_rn(xyC, zs), __fadd_rn(zxC, ys),
                        __fadd_rn(xyC, zs), __fadd_rn(__fmul_rn(y, yC), c), __fsub_rn(yzC, xs),
                        __fsub_rn(zxC, ys), __fadd_rn(yzC, xs), __fadd_rn(__fmul_rn(z, zC), c)};
    double M[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)                                            // np.dot(R, root_pose) in float32
        M[r * 3 + cc] = (double)fmaf(R[r * 3 + 2], A[6 + cc], fmaf(R[r * 3 + 1], A[3 + cc], __fmul_rn(R[r * 3], A[cc])));
    double rv[3];
    rotation_log(M, rv);
    po[0] = (float)rv[0]; po[1] = (float)rv[1]; po[2] = (float)rv[2];
  }
  // translation (dataset.py:289-292): R trans + t/1000 - root + R root, root = rest root joint for betas'
  float root[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = a.j_template[c];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc = fmaf(a.j_shapedirs[c * 10 + k], any ? be[k] : a.default_betas[k], acc);
    root[c] = acc;
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float* tr = a.trans + (size_t)b * 3;
    const float rt = fmaf(R[r * 3 + 2], tr[2], fmaf(R[r * 3 + 1], tr[1], __fmul_rn(R[r * 3], tr[0])));
    const float rr = fmaf(R[r * 3 + 2], root[2], fmaf(R[r * 3 + 1], root[1], __fmul_rn(R[r * 3], root[0])));
    const float t0 = __fadd_rn(rt, __fdiv_rn(a.cam_t[(size_t)b * 3 + r], 1000.f));
    a.trans_out[(size_t)b * 3 + r] = __fadd_rn(__fsub_rn(t0, root[r]), rr);
  }
}

}  // namespace
}  // namespace gator

extern "C" int gator_smpl_cam_fixup(const gator_smpl_cam_args* a, void* stream) {
  using namespace gator;
  GATOR_REQUIRE(a, "gator_smpl_cam_fixup: null args");
  GATOR_REQUIRE(a->batch >= 0, "gator_smpl_cam_fixup: negative batch");
  if (a->batch == 0) return GATOR_OK;
  GATOR_REQUIRE(a->j_template && a->j_shapedirs && a->default_betas && a->pose && a->betas && a->trans && a->cam_R &&
                    a->cam_t && a->pose_out && a->betas_out && a->trans_out,
                "gator_smpl_cam_fixup: null buffer");
  smpl_cam_fixup_kernel<<<ceil_div(a->batch, 128), 128, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("smpl_cam_fixup");
}
