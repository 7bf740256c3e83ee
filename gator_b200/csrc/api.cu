// Error plumbing and ABI introspection for the C boundary (include/gator_b200.h).
#include "common.cuh"

namespace gator {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static long long g_launches = 0;

int check_launch(const char* what) {
  __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return GATOR_ERR_LAUNCH;
  }
  return GATOR_OK;
}

}  // namespace gator

extern "C" int gator_abi_version(void) { return GATOR_ABI_VERSION; }
extern "C" long long gator_launch_count(int reset) {
  long long v = __atomic_load_n(&gator::g_launches, __ATOMIC_RELAXED);
  if (reset) __atomic_store_n(&gator::g_launches, 0, __ATOMIC_RELAXED);
  return v;
}
extern "C" const char* gator_last_error(void) { return gator::g_err; }
extern "C" size_t gator_abi_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(gator_gat_args);
    case 1: return sizeof(gator_mdr_args);
    case 2: return sizeof(gator_smpl_args);
    case 3: return sizeof(gator_csr_args);
    case 4: return sizeof(gator_gemm_args);
    case 5: return sizeof(gator_eval_args);
    case 6: return sizeof(gator_pose2d_args);
    case 7: return sizeof(gator_smpl_cam_args);
    case 8: return sizeof(gator_upsample2_args);
    case 9: return sizeof(gator_lbs_args);
    case 10: return sizeof(gator_mano_post_args);
    default: return 0;
  }
}
