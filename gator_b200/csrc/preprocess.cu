// 2D-pose pre-processing: what demo/run.py:103-133 and the datasets
// (data/Human36M/dataset.py:383-389) do per sample on the host with numpy + cv2 before every forward:
//   [add_pelvis / add_neck]  ->  get_bbox  ->  process_bbox (aspect ratio)  ->  j2d_processing (crop affine, rot 0)
//   ->  / [W, H]  ->  (x - mean) / std per axis.
// One warp per sample, lane = joint (<= 32).  The arithmetic follows the reference's dtype flow: bbox and affine
// points are rounded to float32 where numpy rounds them, the 3-point affine is solved and applied in float64 like
// cv2.getAffineTransform / np.dot, and mean / std are accumulated sequentially in float32 like np.mean(axis=0).
// Latency-bound by construction (a few hundred bytes per sample); it exists so that the batch-1 demo path is
// "pixels in -> mesh out" on the device with no host arithmetic in between.
#include "common.cuh"

namespace gator {
namespace {

__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(128)
pose2d_preprocess_kernel(gator_pose2d_args a) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= a.batch) return;
  const int Jin = a.joints_in, J = Jin + a.n_mid;
  const float* src = a.joints + (size_t)b * Jin * a.in_stride;
  // joints (float64 from here on, like the reference's npy input); synthesised joints = midpoints
  double x = 0.0, y = 0.0;
  if (lane < Jin) {
    x = src[lane * a.in_stride]; y = src[lane * a.in_stride + 1];
  } else if (lane < J) {
    const int ia = a.mid_a[lane - Jin], ib = a.mid_b[lane - Jin];
    x = ((double)src[ia * a.in_stride] + (double)src[ib * a.in_stride]) * 0.5;
    y = ((double)src[ia * a.in_stride + 1] + (double)src[ib * a.in_stride + 1]) * 0.5;
  }
  const bool on = lane < J;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  // get_bbox (coord_utils.py:21-39)
  double xmin = warp_min_d(on ? x : inf), xmax = warp_max_d(on ? x : -inf);
  double ymin = warp_min_d(on ? y : inf), ymax = warp_max_d(on ? y : -inf);
  const double xc = (xmin + xmax) / 2., wd = xmax - xmin;
  xmin = xc - 0.5 * wd; xmax = xc + 0.5 * wd;
  const double yc = (ymin + ymax) / 2., hd = ymax - ymin;
  ymin = yc - 0.5 * hd; ymax = yc + 0.5 * hd;
  const float bx = (float)xmin, by = (float)ymin, bw = (float)(xmax - xmin), bh = (float)(ymax - ymin);
  // process_bbox (coord_utils.py:42-66), float32 arithmetic
  const float x2 = __fadd_rn(bx, __fsub_rn(bw, 1.f)), y2 = __fadd_rn(by, __fsub_rn(bh, 1.f));
  const bool valid = __fmul_rn(bw, bh) > 0.f && x2 >= bx && y2 >= by;
  float w = __fsub_rn(x2, bx), h = __fsub_rn(y2, by);
  const float cx = __fadd_rn(bx, __fdiv_rn(w, 2.f)), cy = __fadd_rn(by, __fdiv_rn(h, 2.f));
  const float ar = a.aspect;
  if (w > __fmul_rn(ar, h)) h = __fdiv_rn(w, ar);
  else if (w < __fmul_rn(ar, h)) w = __fmul_rn(h, ar);
  const float pw = __fmul_rn(w, a.bbox_scale), ph = __fmul_rn(h, a.bbox_scale);
  const float px = __fsub_rn(cx, __fdiv_rn(pw, 2.f)), py = __fsub_rn(cy, __fdiv_rn(ph, 2.f));
  // get_center_scale + get_affine_transform for rot = 0 (coord_utils.py:7-18, aug_utils.py:140-172)
  const float c0 = __fadd_rn(px, __fmul_rn(pw, 0.5f)), c1 = __fadd_rn(py, __fmul_rn(ph, 0.5f));
  const float s0x = c0, s0y = c1;
  const float s1x = (float)((double)c0 + 0.0), s1y = (float)((double)c1 + (double)__fmul_rn(pw, -0.5f));
  const float d0 = __fsub_rn(s0x, s1x), d1 = __fsub_rn(s0y, s1y);           // get_3rd_point(src0, src1)
  const float s2x = __fadd_rn(s1x, -d1), s2y = __fadd_rn(s1y, d0);
  const float dw = a.out_w, dh = a.out_h;
  const float t0x = __fmul_rn(dw, 0.5f), t0y = __fmul_rn(dh, 0.5f);
  const float t1x = t0x, t1y = (float)((double)t0y + (double)__fmul_rn(dw, -0.5f));
  const float e0 = __fsub_rn(t0x, t1x), e1 = __fsub_rn(t0y, t1y);
  const float t2x = __fadd_rn(t1x, -e1), t2y = __fadd_rn(t1y, e0);
  // affine through the three pairs: M = [t1-t0, t2-t0] [s1-s0, s2-s0]^-1 in float64
  const double ux = (double)s1x - s0x, uy = (double)s1y - s0y, vx = (double)s2x - s0x, vy = (double)s2y - s0y;
  const double det = ux * vy - vx * uy;
  const double i00 = vy / det, i01 = -vx / det, i10 = -uy / det, i11 = ux / det;
  const double px1 = (double)t1x - t0x, py1 = (double)t1y - t0y, px2 = (double)t2x - t0x, py2 = (double)t2y - t0y;
  const double m00 = px1 * i00 + px2 * i10, m01 = px1 * i01 + px2 * i11;
  const double m10 = py1 * i00 + py2 * i10, m11 = py1 * i01 + py2 * i11;
  const double m02 = (double)t0x - (m00 * s0x + m01 * s0y), m12 = (double)t0y - (m10 * s0x + m11 * s0y);
  const float kx = (float)(m00 * x + m01 * y + m02), ky = (float)(m10 * x + m11 * y + m12);   // kp.astype('float32')
  // -> 0~1, then np.mean / np.std over the joints (sequential float32 accumulation)
  const float nx = __fdiv_rn(kx, dw), ny = __fdiv_rn(ky, dh);
  float sx = 0.f, sy = 0.f;
  for (int j = 0; j < J; ++j) {
    const float vx_ = __shfl_sync(0xffffffffu, nx, j), vy_ = __shfl_sync(0xffffffffu, ny, j);
    sx = j == 0 ? vx_ : __fadd_rn(sx, vx_);
    sy = j == 0 ? vy_ : __fadd_rn(sy, vy_);
  }
  const float mx = __fdiv_rn(sx, (float)J), my = __fdiv_rn(sy, (float)J);
  const float dx = __fsub_rn(nx, mx), dy = __fsub_rn(ny, my);
  const float qx = __fmul_rn(dx, dx), qy = __fmul_rn(dy, dy);
  float vxs = 0.f, vys = 0.f;
  for (int j = 0; j < J; ++j) {
    const float vx_ = __shfl_sync(0xffffffffu, qx, j), vy_ = __shfl_sync(0xffffffffu, qy, j);
    vxs = j == 0 ? vx_ : __fadd_rn(vxs, vx_);
    vys = j == 0 ? vy_ : __fadd_rn(vys, vy_);
  }
  const float sdx = __fsqrt_rn(__fdiv_rn(vxs, (float)J)), sdy = __fsqrt_rn(__fdiv_rn(vys, (float)J));
  const float nanv = __int_as_float(0x7fc00000);
  if (on) {
    float* o = a.pose2d + ((size_t)b * J + lane) * 2;
    o[0] = valid ? __fdiv_rn(dx, sdx) : nanv;
    o[1] = valid ? __fdiv_rn(dy, sdy) : nanv;
    if (a.joint_img) {
      float* q = a.joint_img + ((size_t)b * J + lane) * 2;
      q[0] = valid ? kx : nanv; q[1] = valid ? ky : nanv;
    }
  }
  if (lane == 0) {
    if (a.bbox) { float* q = a.bbox + (size_t)b * 4; q[0] = px; q[1] = py; q[2] = pw; q[3] = ph; }
    if (a.valid) a.valid[b] = valid ? 1 : 0;
  }
}

}  // namespace
}  // namespace gator

extern "C" int gator_pose2d_preprocess(const gator_pose2d_args* a, void* stream) {
  using namespace gator;
  GATOR_REQUIRE(a, "gator_pose2d_preprocess: null args");
  GATOR_REQUIRE(a->batch >= 0 && a->joints_in > 0 && a->n_mid >= 0 && a->n_mid <= 2 && a->joints_in + a->n_mid <= 32,
                "gator_pose2d_preprocess: bad joint count (<= 32 incl. synthesised joints, <= 2 synthesised)");
  GATOR_REQUIRE(a->in_stride >= 2, "gator_pose2d_preprocess: in_stride < 2");
  for (int i = 0; i < a->n_mid; ++i)
    GATOR_REQUIRE(a->mid_a[i] >= 0 && a->mid_a[i] < a->joints_in && a->mid_b[i] >= 0 && a->mid_b[i] < a->joints_in,
                  "gator_pose2d_preprocess: synthesised joint refers to a joint outside the input");
  GATOR_REQUIRE(a->out_w > 0.f && a->out_h > 0.f && a->aspect > 0.f, "gator_pose2d_preprocess: bad crop size");
  if (a->batch == 0) return GATOR_OK;
  GATOR_REQUIRE(a->joints && a->pose2d, "gator_pose2d_preprocess: null buffer");
  pose2d_preprocess_kernel<<<ceil_div(a->batch, 4), 128, 0, (cudaStream_t)stream>>>(*a);
  return check_launch("pose2d_preprocess");
}
