// MANO hand layer around the generic LBS core (manopth/manopth/manolayer.py:109-273; the core itself -
// shape/pose blend shapes, kinematic chain, skinning, :170-230 - is gator_lbs_forward in csrc/smpl.cu):
//   mano_pose_kernel : PCA coefficients -> full axis-angle pose  [root | hands_mean + coeffs @ selected_comps]  (:128-143)
//   mano_post_kernel : finger tips sampled from the vertices, optional palm root, joint re-ordering, centring /
//                      translation and the metres -> millimetres scale on vertices and joints              (:232-256)
#include "common.cuh"

namespace gator {
namespace {

constexpr int MJ = 16;     // chain joints
constexpr int MJ_OUT = 21; // + 5 finger tips
__constant__ int kManoOrder[MJ_OUT] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};   // :243

__global__ void __launch_bounds__(256)
mano_pose_kernel(const float* __restrict__ coeffs, int ld, int ncomps, const float* __restrict__ comps,
                 const float* __restrict__ hands_mean, float* __restrict__ full_pose, long long total) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / 48), c = (int)(i - (long long)b * 48);
  const float* cr = coeffs + (size_t)b * ld;
  float v;
  if (c < 3) {
    v = cr[c];
  } else if (comps) {            // th_hand_pose_coeffs.mm(th_selected_comps): k ascending, fp32
    float acc = 0.f;
    for (int k = 0; k < ncomps; ++k) acc = fmaf(cr[3 + k], __ldg(comps + k * 45 + (c - 3)), acc);
    v = __ldg(hands_mean + c - 3) + acc;
  } else {                       // use_pca = False, joint_rot_mode = 'axisang': the coefficients are the axis-angles
    v = __ldg(hands_mean + c - 3) + cr[c];
  }
  full_pose[i] = v;
}

__global__ void mano_flag_kernel(const float* __restrict__ trans, int n, int* __restrict__ flag) {
  __shared__ int f;
  if (threadIdx.x == 0) f = 0;
  __syncthreads();
  int a = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a |= (trans[i] != 0.f);
  if (a) atomicOr(&f, 1);
  __syncthreads();
  if (threadIdx.x == 0) *flag = f;
}

struct PostParams {
  const float* jtr16;      // (B,16,3) chain joints, metres, no offset
  const float* trans;      // (B,3) or null
  const int* flag;         // null, or [any(trans != 0)]
  float* verts;            // (B,nv,3) in / out
  float* jtr;              // (B,21,3) out
  int nv, center_idx, root_palm;
  int tips[5], palm[2];
  float scale;
};

__global__ void __launch_bounds__(128) mano_post_kernel(PostParams p) {
  __shared__ float j[MJ_OUT][3];
  __shared__ float off[3];
  const int b = blockIdx.x, tid = threadIdx.x;
  float* vb = p.verts + (size_t)b * p.nv * 3;
  if (tid < MJ * 3) (&j[0][0])[tid] = p.jtr16[(size_t)b * MJ * 3 + tid];
  if (tid >= 64 && tid < 64 + 15) {            // tips = th_verts[:, [745, 317, 444 | 445, 556, 673]]
    const int t = (tid - 64) / 3, c = (tid - 64) - t * 3;
    j[MJ + t][c] = vb[p.tips[t] * 3 + c];
  }
  __syncthreads();
  if (p.root_palm && tid < 3) j[0][tid] = (vb[p.palm[0] * 3 + tid] + vb[p.palm[1] * 3 + tid]) / 2.0f;   // :238-240
  __syncthreads();
  const bool use_trans = p.trans && (!p.flag || *p.flag);
  if (tid < 3) {
    float o = 0.f;
    if (use_trans) o = p.trans[(size_t)b * 3 + tid];
    else if (p.center_idx >= 0) o = -j[kManoOrder[p.center_idx]][tid];
    off[tid] = o;
  }
  __syncthreads();
  if (tid < MJ_OUT * 3) {
    const int q = tid / 3, c = tid - q * 3;
    p.jtr[((size_t)b * MJ_OUT + q) * 3 + c] = (j[kManoOrder[q]][c] + off[c]) * p.scale;
  }
  for (int i = tid; i < p.nv * 3; i += 128) vb[i] = (vb[i] + off[i % 3]) * p.scale;
}

}  // namespace
}  // namespace gator

extern "C" int gator_mano_pose(const float* coeffs, int32_t ld, int32_t ncomps, const float* comps, const float* hands_mean,
                               float* full_pose, int32_t batch, void* stream) {
  using namespace gator;
  GATOR_REQUIRE(batch >= 0 && ncomps >= 0 && ncomps <= 45 && ld >= 3 + (comps ? ncomps : 45), "gator_mano_pose: bad shape");
  if (batch == 0) return GATOR_OK;
  GATOR_REQUIRE(coeffs && hands_mean && full_pose, "gator_mano_pose: null buffer");
  const long long total = (long long)batch * 48;
  mano_pose_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(coeffs, ld, ncomps, comps, hands_mean, full_pose, total);
  return check_launch("mano_pose");
}

extern "C" int gator_mano_post(const gator_mano_post_args* a, void* stream_) {
  using namespace gator;
  cudaStream_t stream = (cudaStream_t)stream_;
  GATOR_REQUIRE(a, "gator_mano_post: null args");
  GATOR_REQUIRE(a->batch >= 0 && a->n_verts > 0 && a->center_idx >= -1 && a->center_idx < MJ_OUT, "gator_mano_post: bad argument");
  if (a->batch == 0) return GATOR_OK;
  GATOR_REQUIRE(a->jtr16 && a->verts && a->jtr, "gator_mano_post: null buffer");
  GATOR_REQUIRE(!a->has_trans || a->trans, "gator_mano_post: has_trans without trans");
  for (int t = 0; t < 5; ++t) GATOR_REQUIRE(a->tip_verts[t] >= 0 && a->tip_verts[t] < a->n_verts, "gator_mano_post: tip vertex out of range");
  for (int t = 0; t < 2; ++t) GATOR_REQUIRE(a->palm_verts[t] >= 0 && a->palm_verts[t] < a->n_verts, "gator_mano_post: palm vertex out of range");
  const bool check = a->has_trans && a->check_zero_norm;
  GATOR_REQUIRE(!check || a->flag_ws, "gator_mano_post: check_zero_norm needs 4 bytes of flag workspace");
  if (check) {
    mano_flag_kernel<<<1, 256, 0, stream>>>(a->trans, a->batch * 3, a->flag_ws);
    GATOR_TRY(check_launch("mano_flag"));
  }
  PostParams p;
  p.jtr16 = a->jtr16; p.trans = a->has_trans ? a->trans : nullptr; p.flag = check ? a->flag_ws : nullptr;
  p.verts = a->verts; p.jtr = a->jtr; p.nv = a->n_verts; p.center_idx = a->center_idx; p.root_palm = a->root_palm;
  for (int t = 0; t < 5; ++t) p.tips[t] = a->tip_verts[t];
  p.palm[0] = a->palm_verts[0]; p.palm[1] = a->palm_verts[1];
  p.scale = a->scale == 0.f ? 1.f : a->scale;
  mano_post_kernel<<<a->batch, 128, 0, stream>>>(p);
  return check_launch("mano_post");
}
