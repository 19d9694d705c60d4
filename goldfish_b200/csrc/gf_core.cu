// Error bookkeeping and small utilities of the C ABI.
#include "gf_common.cuh"
#include <string.h>

namespace gf {
static thread_local char g_err[512] = "";
int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return GF_ERR_CUDA;
}
static long long g_launches = 0;
void count_launch(int n) { g_launches += n; }
int check_launch(const char* where) {
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, where);
  return GF_OK;
}
}  // namespace gf

extern "C" const char* gf_last_error(void) { return gf::g_err; }
extern "C" int gf_version(void) { return 100; }
extern "C" long long gf_launch_count(void) { return gf::g_launches; }
