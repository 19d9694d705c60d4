// FP64 peak micro-benchmarks (measurement infrastructure, not on the product path): the denominators of the
// "FP64 pipe utilisation against the FP64 peak" figures.
//   mode 0: DFMA  -- 8 independent fma chains per thread
//   mode 1: DMMA  -- mma.sync.aligned.m8n8k4 f64, 4 independent accumulator tiles per warp
#include "gf_common.cuh"

namespace gf {
__global__ void __launch_bounds__(256)
k_peak_dfma(int iters, double* out) {
  double a[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) a[q] = 1.0 + 1e-3 * (threadIdx.x + q);
  const double b = 1.0 - 1e-9, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = fma(a[q], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += a[q];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256)
k_peak_dmma(int iters, double* out) {
  double c[4][2];
#pragma unroll
  for (int q = 0; q < 4; ++q) { c[q][0] = 0.0; c[q][1] = 0.0; }
  const double a = 1.0 + 1e-6 * threadIdx.x, b = 1.0 - 1e-6 * threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) s += c[q][0] + c[q][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace gf

// Launches one measurement kernel; returns the floating-point operations it performs in *flops
// (time it with CUDA events on `stream`).  out: [grid*256] doubles.
extern "C" int gf_peak_fp64(int mode, int grid, int iters, double* out, double* flops, void* stream) {
  if (!out || grid < 1 || iters < 1) return gf::set_error(GF_ERR_BADARG, "gf_peak_fp64: bad argument");
  if (mode == 0) {
    gf::k_peak_dfma<<<grid, 256, 0, (cudaStream_t)stream>>>(iters, out);
    if (flops) *flops = 2.0 * 8.0 * iters * 256.0 * grid;
  } else {
    gf::k_peak_dmma<<<grid, 256, 0, (cudaStream_t)stream>>>(iters, out);
    if (flops) *flops = 2.0 * 8 * 8 * 4 * 4.0 * iters * 8.0 * grid;   // 4 mma per warp-iteration, 8 warps per CTA
  }
  return gf::check_launch("k_peak_fp64");
}
