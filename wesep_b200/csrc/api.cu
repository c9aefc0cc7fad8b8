// Library-wide state and trivial entry points.
#include "common.cuh"

namespace wb {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
int g_gemm_mode = 0;
int g_gemm_backend = 1;   // tcgen05/TMA/TMEM GEMMs wherever eligible
}  // namespace wb

extern "C" int wesep_b200_version(void) { return WESEP_B200_VERSION; }
extern "C" const char* wesep_b200_last_error(void) { return wb::g_err; }
extern "C" uint64_t wesep_b200_launch_count(void) { return wb::g_launches.load(); }
extern "C" int wesep_b200_set_gemm_mode(int mode) {
  if (mode != 0 && mode != 1) return wb::fail(-2, "gemm mode must be 0 (3xTF32) or 1 (TF32)");
  wb::g_gemm_mode = mode;
  return 0;
}
extern "C" int wesep_b200_set_gemm_backend(int backend) {
  if (backend != 0 && backend != 1) return wb::fail(-2, "gemm backend must be 0 (mma.sync) or 1 (tcgen05)");
  wb::g_gemm_backend = backend;
  return 0;
}
