// Error bookkeeping and small utilities of the C ABI.
#include "gf_common.cuh"
#include <string.h>
#include <stddef.h>

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

// ABI self-description: size of a public struct and offset of its last field, so that a binding
// (goldfish_b200/_capi.py, or any other host) can check its mirror of include/goldfish_b200.h at load time.
extern "C" int gf_abi_layout(int which, int64_t* size, int64_t* last_offset) {
  if (!size || !last_offset) return gf::set_error(GF_ERR_BADARG, "gf_abi_layout: null argument");
#define GF_LAYOUT(T, last) *size = (int64_t)sizeof(T); *last_offset = (int64_t)offsetof(T, last); return GF_OK
  switch (which) {
    case 0: GF_LAYOUT(GfPatchDesc, f);
    case 1: GF_LAYOUT(GfCsr, vals);
    case 2: GF_LAYOUT(GfModel, T);
    case 3: GF_LAYOUT(GfShellOut, dt_el);
    case 4: GF_LAYOUT(GfPenalty, K_pos);
    case 5: GF_LAYOUT(GfPenaltyP, field);
    case 6: GF_LAYOUT(GfCsrT, perm);
    case 7: GF_LAYOUT(GfSchwarz, flag);
    case 8: GF_LAYOUT(GfDist, comm);
    case 9: GF_LAYOUT(GfPrecond, cinv_rows);
    case 10: GF_LAYOUT(GfPcgWork, nodes);
    case 11: GF_LAYOUT(GfGmresWork, nodes);
    default: return gf::set_error(GF_ERR_BADARG, "gf_abi_layout: unknown struct id");
  }
#undef GF_LAYOUT
}
