// CSR SpMV / SpMV^T, fused Krylov vector kernels and the PCG driver (FP64).
//
// Replaces PETSc MatMult / MatMultTranspose (A_x_b / AT_x_b,
// /root/reference/GOLDFISH/operations/disp_imop.py:68-121) and the MUMPS solves
// of /root/reference/GOLDFISH/utils/opt_utils.py:156-209 (K is symmetric by
// construction, nonmatching_opt.py:804-809, so state and adjoint solve share
// one CG).  All reductions use fixed trees over a fixed grid: results are
// bit-reproducible from run to run.
#include "gf_common.cuh"
#include <stdlib.h>

namespace gf {

constexpr int RED_THREADS = 256;
constexpr int MAX_PARTIAL = 1024;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  if (threadIdx.x == 0) sh[0] = r;
  __syncthreads();
  r = sh[0];
  return r;
}

// sum of `n` partials with stride `stride`, identical in every CTA
__device__ __forceinline__ double sum_partials(const double* p, int n, int stride, double* sh) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += p[(size_t)i * stride];
  return block_sum(s, sh);
}

// ---------------------------------------------------------------- SpMV ------
// streaming loads for the matrix (read once), cached loads for the gathered vector
template <int CS> __device__ __forceinline__ double ld_m(const double* p) { return CS ? __ldcs(p) : __ldg(p); }
template <int CS> __device__ __forceinline__ int ld_m(const int32_t* p) { return CS ? __ldcs(p) : __ldg(p); }
#define ld_stream ld_m<CS>

// simple variant: one chunk of 32 entries per loop trip
__global__ void __launch_bounds__(256)
k_spmv_simple(GfCsr A, const double* __restrict__ x, double* __restrict__ y, double alpha, double beta,
              const double* __restrict__ dotv, double* __restrict__ partial) {
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double dacc = 0.0;
  for (; row < A.nrows; row += nwarps) {
    const int64_t s = A.indptr[row], e = A.indptr[row + 1];
    double sum = 0.0;
    for (int64_t k = s + lane; k < e; k += 32) sum = fma(A.vals[k], __ldg(x + A.indices[k]), sum);
    sum = warp_sum(sum);
    if (lane == 0) {
      double v = alpha * sum;
      if (beta != 0.0) v = fma(beta, y[row], v);
      y[row] = v;
      if (dotv) dacc = fma(dotv[row], v, dacc);
    }
  }
  if (partial) {
    const double b = block_sum(dacc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = b;
  }
}

// Warp per row; the row's chunks of 32 entries are loaded four at a time before any
// FMA so that 4 x (8 + 4) bytes per lane are in flight (rows of K hold ~147 entries).
template <int CS>
__global__ void __launch_bounds__(256)
k_spmv_u4(GfCsr A, const double* __restrict__ x, double* __restrict__ y, double alpha, double beta,
          const double* __restrict__ dotv, double* __restrict__ partial) {
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double dacc = 0.0;
  for (; row < A.nrows; row += nwarps) {
    const int64_t s = A.indptr[row], e = A.indptr[row + 1];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int64_t k = s + lane;
    for (; k + 96 < e; k += 128) {
      const int c0 = ld_stream(A.indices + k), c1 = ld_stream(A.indices + k + 32),
                c2 = ld_stream(A.indices + k + 64), c3 = ld_stream(A.indices + k + 96);
      const double v0 = ld_stream(A.vals + k), v1 = ld_stream(A.vals + k + 32),
                   v2 = ld_stream(A.vals + k + 64), v3 = ld_stream(A.vals + k + 96);
      s0 = fma(v0, __ldg(x + c0), s0); s1 = fma(v1, __ldg(x + c1), s1);
      s2 = fma(v2, __ldg(x + c2), s2); s3 = fma(v3, __ldg(x + c3), s3);
    }
    {   // tail: up to four guarded chunks, still loaded before use
      const bool p0 = k < e, p1 = k + 32 < e, p2 = k + 64 < e, p3 = k + 96 < e;
      const int c0 = p0 ? ld_stream(A.indices + k) : 0, c1 = p1 ? ld_stream(A.indices + k + 32) : 0,
                c2 = p2 ? ld_stream(A.indices + k + 64) : 0, c3 = p3 ? ld_stream(A.indices + k + 96) : 0;
      const double v0 = p0 ? ld_stream(A.vals + k) : 0.0, v1 = p1 ? ld_stream(A.vals + k + 32) : 0.0,
                   v2 = p2 ? ld_stream(A.vals + k + 64) : 0.0, v3 = p3 ? ld_stream(A.vals + k + 96) : 0.0;
      s0 = fma(v0, __ldg(x + c0), s0); s1 = fma(v1, __ldg(x + c1), s1);
      s2 = fma(v2, __ldg(x + c2), s2); s3 = fma(v3, __ldg(x + c3), s3);
    }
    const double sum = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) {
      double v = alpha * sum;
      if (beta != 0.0) v = fma(beta, y[row], v);
      y[row] = v;
      if (dotv) dacc = fma(dotv[row], v, dacc);
    }
  }
  if (partial) {
    const double b = block_sum(dacc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = b;
  }
}

__global__ void __launch_bounds__(256)
k_spmv_t(GfCsr A, GfCsrT At, const double* __restrict__ x, double* __restrict__ y, double alpha,
         double beta) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (; row < At.nrows; row += nwarps) {
    const int64_t s = At.indptr[row], e = At.indptr[row + 1];
    double sum = 0.0;
    for (int64_t k = s + lane; k < e; k += 32)
      sum = fma(A.vals[At.perm[k]], __ldg(x + At.indices[k]), sum);
    sum = warp_sum(sum);
    if (lane == 0) {
      double v = alpha * sum;
      if (beta != 0.0) v = fma(beta, y[row], v);
      y[row] = v;
    }
  }
}

#undef ld_stream
static int g_spmv_variant = -1, g_spmv_grid = 1024;
static void spmv_config() {
  if (g_spmv_variant >= 0) return;
  g_spmv_variant = 0;
  if (const char* v = getenv("GF_SPMV_VARIANT")) g_spmv_variant = atoi(v);
  if (const char* v = getenv("GF_SPMV_GRID")) g_spmv_grid = atoi(v);
  if (g_spmv_grid > MAX_PARTIAL) g_spmv_grid = MAX_PARTIAL;
}
static void launch_spmv(int grid, cudaStream_t st, const GfCsr& A, const double* x, double* y, double alpha,
                        double beta, const double* dotv, double* partial) {
  spmv_config();
  if (g_spmv_variant == 1) k_spmv_u4<1><<<grid, 256, 0, st>>>(A, x, y, alpha, beta, dotv, partial);
  else if (g_spmv_variant == 2) k_spmv_u4<0><<<grid, 256, 0, st>>>(A, x, y, alpha, beta, dotv, partial);
  else k_spmv_simple<<<grid, 256, 0, st>>>(A, x, y, alpha, beta, dotv, partial);
}
static int spmv_grid(int64_t nrows) {
  spmv_config();
  int64_t g = (nrows * 32 + 255) / 256;
  if (g > g_spmv_grid) g = g_spmv_grid;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------- vector ops ---
__global__ void k_axpby(int64_t n, double a, const double* x, double b, double* y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = (b == 0.0) ? a * x[i] : fma(a, x[i], b * y[i]);
}

__global__ void __launch_bounds__(RED_THREADS)
k_dot_partial(int64_t n, const double* x, const double* y, double* partial) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s = fma(x[i], y[i], s);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(RED_THREADS)
k_finalize(const double* partial, int n, int stride, double* out) {
  __shared__ double sh[32];
  const double s = sum_partials(partial, n, stride, sh);
  if (threadIdx.x == 0) *out = s;
}

__global__ void __launch_bounds__(RED_THREADS)
k_reduce_wv(int64_t nel, const double* WV, double* out2) {
  __shared__ double sh[32];
  double w = 0.0, v = 0.0;
  for (int64_t i = threadIdx.x; i < nel; i += blockDim.x) { w += WV[2 * i]; v += WV[2 * i + 1]; }
  w = block_sum(w, sh);
  v = block_sum(v, sh);
  if (threadIdx.x == 0) { out2[0] = w; out2[1] = v; }
}

static int vec_grid(int64_t n) {
  int64_t g = (n + RED_THREADS - 1) / RED_THREADS;
  if (g > MAX_PARTIAL) g = MAX_PARTIAL;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------- PCG ----
// scal: [0] rz(even) [1] rz(odd) [2] alpha [3] beta [4] rr [5] bb [6] pAp
__global__ void __launch_bounds__(RED_THREADS)
k_pcg_init(int64_t n, const double* b, const double* dinv, double* x, double* r, double* z, double* p,
           double* partial) {
  __shared__ double sh[32];
  double rz = 0.0, bb = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double bi = b[i];
    const double zi = dinv[i] * bi;
    x[i] = 0.0; r[i] = bi; z[i] = zi; p[i] = zi;
    rz = fma(bi, zi, rz); bb = fma(bi, bi, bb);
  }
  rz = block_sum(rz, sh);
  bb = block_sum(bb, sh);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = rz; partial[2 * blockIdx.x + 1] = bb; }
}

__global__ void __launch_bounds__(RED_THREADS)
k_pcg_init_fin(const double* partial, int n, double* scal) {
  __shared__ double sh[32];
  const double rz = sum_partials(partial, n, 2, sh);
  const double bb = sum_partials(partial + 1, n, 2, sh);
  if (threadIdx.x == 0) { scal[0] = rz; scal[5] = bb; scal[4] = bb; }
}

// x += alpha p, r -= alpha Ap, z = Dinv r ; partial sums of r.z and r.r
__global__ void __launch_bounds__(RED_THREADS)
k_pcg_update(int64_t n, const double* __restrict__ p, const double* __restrict__ Ap,
             const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
             double* __restrict__ z, const double* pAp_partial, int npart, double* scal, int parity,
             double* partial2) {
  __shared__ double sh[32];
  const double pAp = sum_partials(pAp_partial, npart, 1, sh);
  const double alpha = (pAp != 0.0) ? scal[parity] / pAp : 0.0;
  double rz = 0.0, rr = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, Ap[i], r[i]);
    const double zi = dinv[i] * ri;
    r[i] = ri; z[i] = zi;
    rz = fma(ri, zi, rz); rr = fma(ri, ri, rr);
  }
  rz = block_sum(rz, sh);
  rr = block_sum(rr, sh);
  if (threadIdx.x == 0) {
    partial2[2 * blockIdx.x] = rz; partial2[2 * blockIdx.x + 1] = rr;
    if (blockIdx.x == 0) { scal[2] = alpha; scal[6] = pAp; }
  }
}

// p = z + beta p
__global__ void __launch_bounds__(RED_THREADS)
k_pcg_dir(int64_t n, const double* __restrict__ z, double* __restrict__ p, const double* partial2,
          int npart, double* scal, int parity) {
  __shared__ double sh[32];
  const double rz = sum_partials(partial2, npart, 2, sh);
  const double rr = sum_partials(partial2 + 1, npart, 2, sh);
  const double beta = (scal[parity] != 0.0) ? rz / scal[parity] : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = fma(beta, p[i], z[i]);
  if (blockIdx.x == 0 && threadIdx.x == 0) { scal[1 - parity] = rz; scal[3] = beta; scal[4] = rr; }
}

__global__ void k_jacobi(GfCsr A, double* dinv) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < A.nrows;
       row += (int64_t)gridDim.x * blockDim.x) {
    double d = 1.0;
    for (int64_t k = A.indptr[row]; k < A.indptr[row + 1]; ++k)
      if (A.indices[k] == row) { d = A.vals[k]; break; }
    dinv[row] = (d != 0.0) ? 1.0 / d : 1.0;
  }
}

__global__ void k_bc_diag(GfModel M, double diag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M.n_bc; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = M.bc_list[i];
    for (int64_t k = M.K.indptr[row]; k < M.K.indptr[row + 1]; ++k)
      M.K.vals[k] = (M.K.indices[k] == row) ? diag : 0.0;
  }
}

// Node-wise SpMV for the tangent's row layout (EXPERIMENTAL, not on the default path yet): the three field rows of a
// control point hold the same column list (tests/test_symbolic.py), so one warp takes the three rows of a node,
// reads each column index and each x entry once and multiplies three value streams: 9.33 B per non-zero instead
// of 12.  Lane partition and reduction per row are those of k_spmv_simple, so y is bitwise the same.
__global__ void __launch_bounds__(256)
k_spmv_node(GfCsr A, const int64_t* __restrict__ node_row0, const int32_t* __restrict__ node_stride, int64_t n_nodes,
            const double* __restrict__ x, double* __restrict__ y, double alpha, double beta,
            const double* __restrict__ dotv, double* __restrict__ partial) {
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double dacc = 0.0;
  for (int64_t nd = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; nd < n_nodes; nd += nwarps) {
    const int64_t r0 = node_row0[nd], st = node_stride[nd];
    const int64_t s0 = A.indptr[r0], len = A.indptr[r0 + 1] - s0;
    const int64_t s1 = A.indptr[r0 + st], s2 = A.indptr[r0 + 2 * st];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int64_t k = lane; k < len; k += 32) {
      const double xv = __ldg(x + A.indices[s0 + k]);
      a0 = fma(A.vals[s0 + k], xv, a0);
      a1 = fma(A.vals[s1 + k], xv, a1);
      a2 = fma(A.vals[s2 + k], xv, a2);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane < 3) {
      const int64_t r = r0 + lane * st;
      double v = alpha * (lane == 0 ? a0 : (lane == 1 ? a1 : a2));
      if (beta != 0.0) v = fma(beta, y[r], v);
      y[r] = v;
      if (dotv) dacc = fma(dotv[r], v, dacc);
    }
  }
  if (partial) {
    const double b = block_sum(dacc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = b;
  }
}

}  // namespace gf

using namespace gf;

extern "C" int gf_spmv(const GfCsr* A, const double* x, double* y, double alpha, double beta, void* stream) {
  if (!A || !x || !y) return set_error(GF_ERR_BADARG, "gf_spmv: null argument");
  if (A->nrows == 0) return GF_OK;
  launch_spmv(spmv_grid(A->nrows), (cudaStream_t)stream, *A, x, y, alpha, beta, nullptr, nullptr);
  return check_launch("k_spmv");
}

extern "C" int gf_spmv_node(const GfCsr* A, const int64_t* node_row0, const int32_t* node_stride, int64_t n_nodes,
                            const double* x, double* y, double alpha, double beta, void* stream) {
  if (!A || !node_row0 || !node_stride || !x || !y) return set_error(GF_ERR_BADARG, "gf_spmv_node: null argument");
  if (n_nodes == 0) return GF_OK;
  k_spmv_node<<<spmv_grid(n_nodes), 256, 0, (cudaStream_t)stream>>>(*A, node_row0, node_stride, n_nodes, x, y, alpha, beta,
                                                                    nullptr, nullptr);
  return check_launch("k_spmv_node");
}

extern "C" int gf_spmv_t(const GfCsr* A, const GfCsrT* At, const double* x, double* y, double alpha,
                         double beta, void* stream) {
  if (!A || !At || !x || !y) return set_error(GF_ERR_BADARG, "gf_spmv_t: null argument");
  if (At->nrows == 0) return GF_OK;
  k_spmv_t<<<spmv_grid(At->nrows), 256, 0, (cudaStream_t)stream>>>(*A, *At, x, y, alpha, beta);
  return check_launch("k_spmv_t");
}

extern "C" int gf_axpby(int64_t n, double a, const double* x, double b, double* y, void* stream) {
  if (n <= 0) return GF_OK;
  k_axpby<<<vec_grid(n), RED_THREADS, 0, (cudaStream_t)stream>>>(n, a, x, b, y);
  return check_launch("k_axpby");
}

extern "C" int gf_dot(int64_t n, const double* x, const double* y, double* partial, double* out_dev, void* stream) {
  const int g = vec_grid(n);
  k_dot_partial<<<g, RED_THREADS, 0, (cudaStream_t)stream>>>(n, x, y, partial);
  k_finalize<<<1, RED_THREADS, 0, (cudaStream_t)stream>>>(partial, g, 1, out_dev);
  count_launch(1);
  return check_launch("gf_dot");
}

extern "C" int gf_reduce_wv(int64_t nel, const double* WV, double* out2_dev, void* stream) {
  k_reduce_wv<<<1, RED_THREADS, 0, (cudaStream_t)stream>>>(nel, WV, out2_dev);
  return check_launch("k_reduce_wv");
}

extern "C" int gf_jacobi_setup(const GfCsr* A, double* dinv, void* stream) {
  k_jacobi<<<vec_grid(A->nrows), RED_THREADS, 0, (cudaStream_t)stream>>>(*A, dinv);
  return check_launch("k_jacobi");
}

extern "C" int gf_bc_set_diag(const GfModel* m, double diag, void* stream) {
  if (m->n_bc <= 0) return GF_OK;
  k_bc_diag<<<vec_grid(m->n_bc), RED_THREADS, 0, (cudaStream_t)stream>>>(*m, diag);
  return check_launch("k_bc_diag");
}

// rz partials (slot 0 of partial2) recomputed after a non-diagonal preconditioner
__global__ void __launch_bounds__(RED_THREADS)
k_pcg_rz(int64_t n, const double* __restrict__ r, const double* __restrict__ z, double* partial2) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s = fma(r[i], z[i], s);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partial2[2 * blockIdx.x] = s;
}
__global__ void __launch_bounds__(RED_THREADS)
k_pcg_init_p(int64_t n, const double* z, double* p) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = z[i];
}

__global__ void k_zero_list(const int32_t* list, int64_t n, double* v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[list[i]] = 0.0;
}

namespace gf {
int dist_allreduce(const GfDist* d, double* buf, int64_t n, cudaStream_t st);   // gf_dist.cu

// y = A x with x replicated: single process -> whole matrix (optionally fused dot partials dotv.y);
// sharded -> owned row ranges, the rest of y zeroed, then summed over the ranks (NCCL all-reduce on `st`).
static int apply_A(const GfCsr& A, const GfDist* dist, const GfNodeRows* nodes, const double* x, double* y,
                   const double* dotv, double* partial, int* npart, cudaStream_t st) {
  const bool sharded = dist && dist->n_ranges > 0;
  const bool nodewise = nodes && nodes->n > 0 && nodes->row0 && nodes->stride;
  const int64_t n = A.nrows;
  if (!sharded) {
    if (nodewise) {
      // the three field rows of a control point share one column list: one index / x read per three non-zeros
      const int gs = spmv_grid(nodes->n);
      k_spmv_node<<<gs, 256, 0, st>>>(A, nodes->row0, nodes->stride, nodes->n, x, y, 1.0, 0.0, dotv, partial);
      count_launch(1);
      if (npart) *npart = gs;
      return GF_OK;
    }
    const int gs = spmv_grid(n);
    launch_spmv(gs, st, A, x, y, 1.0, 0.0, dotv, partial);
    count_launch(1);
    if (npart) *npart = gs;
    return GF_OK;
  }
  cudaError_t e = cudaMemsetAsync(y, 0, (size_t)n * sizeof(double), st);
  if (e != cudaSuccess) return set_cuda_error(e, "apply_A memset");
  if (nodewise) {                              // `nodes` lists the control points of the owned patches
    k_spmv_node<<<spmv_grid(nodes->n), 256, 0, st>>>(A, nodes->row0, nodes->stride, nodes->n, x, y, 1.0, 0.0, nullptr, nullptr);
    count_launch(1);
  } else {
    for (int q = 0; q < dist->n_ranges; ++q) {
      const int64_t b0 = dist->ranges_h[2 * q], b1 = dist->ranges_h[2 * q + 1];
      if (b1 <= b0) continue;
      GfCsr sub = A; sub.indptr = A.indptr + b0; sub.nrows = b1 - b0;
      launch_spmv(spmv_grid(sub.nrows), st, sub, x, y + b0, 1.0, 0.0, nullptr, nullptr);
      count_launch(1);
    }
  }
  int rc = dist_allreduce(dist, y, n, st);
  if (rc) return rc;
  if (dotv && partial) {
    const int gv = vec_grid(n);
    k_dot_partial<<<gv, RED_THREADS, 0, st>>>(n, dotv, y, partial);
    count_launch(1);
    if (npart) *npart = gv;
  }
  return GF_OK;
}
}  // namespace gf

namespace gf {
// y[row] = A[row, :] . x for a dense row-major FP64 slab: one warp per row, 128-bit loads, x through the read-only
// path (it is a few hundred kB and stays in L2).  HBM-bound: 8 bytes per entry, read once.
__global__ void __launch_bounds__(256)
k_dense_rows(const double* __restrict__ A, int64_t rows, int64_t ncols, const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n2 = ncols >> 1;
  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += nwarps) {
    const double* a = A + row * ncols;
    double s0 = 0.0, s1 = 0.0;
    if ((((uintptr_t)a) & 15) == 0) {
      const double2* a2 = reinterpret_cast<const double2*>(a);
      const double2* x2 = reinterpret_cast<const double2*>(x);
      int64_t k = lane;
      for (; k + 96 < n2; k += 128) {
        const double2 v0 = __ldcs(a2 + k), v1 = __ldcs(a2 + k + 32), v2 = __ldcs(a2 + k + 64), v3 = __ldcs(a2 + k + 96);
        const double2 w0 = __ldg(x2 + k), w1 = __ldg(x2 + k + 32), w2 = __ldg(x2 + k + 64), w3 = __ldg(x2 + k + 96);
        s0 = fma(v0.x, w0.x, s0); s1 = fma(v0.y, w0.y, s1); s0 = fma(v1.x, w1.x, s0); s1 = fma(v1.y, w1.y, s1);
        s0 = fma(v2.x, w2.x, s0); s1 = fma(v2.y, w2.y, s1); s0 = fma(v3.x, w3.x, s0); s1 = fma(v3.y, w3.y, s1);
      }
      for (; k < n2; k += 32) { const double2 v = __ldcs(a2 + k); const double2 w = __ldg(x2 + k); s0 = fma(v.x, w.x, s0); s1 = fma(v.y, w.y, s1); }
      if ((ncols & 1) && lane == 0) s0 = fma(a[ncols - 1], x[ncols - 1], s0);
    } else {
      for (int64_t k = lane; k < ncols; k += 32) s0 = fma(a[k], __ldg(x + k), s0);
    }
    const double s = warp_sum(s0 + s1);
    if (lane == 0) y[row] = s;
  }
}
}  // namespace gf

extern "C" int gf_precond_apply(const GfPrecond* pc, const double* r, double* z, int64_t n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (pc->cinv) {
    // sharded two-level: restriction (replicated), own row slab of Kc^-1 r_c, own fine blocks, prolongation of the
    // slab, ONE all-reduce for both levels
    const int64_t nc = pc->Rt.nrows;
    int rc = gf_spmv(&pc->Rt, r, pc->rc, 1.0, 0.0, st);
    if (rc) return rc;
    if (pc->n_bc_c > 0) { k_zero_list<<<vec_grid(pc->n_bc_c), RED_THREADS, 0, st>>>(pc->bc_c, pc->n_bc_c, pc->rc); count_launch(1); }
    cudaError_t e = cudaMemsetAsync(pc->zc, 0, (size_t)nc * sizeof(double), st);
    if (e != cudaSuccess) return set_cuda_error(e, "gf_precond_apply memset");
    if (pc->cinv_rows > 0) {
      int64_t g = (pc->cinv_rows * 32 + 255) / 256; if (g > 148 * 8) g = 148 * 8;
      k_dense_rows<<<(unsigned)g, 256, 0, st>>>(pc->cinv, pc->cinv_rows, nc, pc->rc, pc->zc + pc->cinv_row0);
      count_launch(1);
    }
    rc = gf_schwarz_apply(pc->fine, r, z, n, st);
    if (rc) return rc;
    rc = gf_spmv(&pc->P, pc->zc, z, 1.0, 1.0, st);             // z += P z_c[slab]
    if (rc) return rc;
    return dist_allreduce(pc->dist, z, n, st);
  }
  if (!pc->coarse) {
    int rc0 = gf_schwarz_apply(pc->fine, r, z, n, st);
    if (rc0 == 0 && pc->dist && pc->dist->n_ranges > 0) rc0 = dist_allreduce(pc->dist, z, n, st);
    return rc0;
  }
  int rc = gf_spmv(&pc->Rt, r, pc->rc, 1.0, 0.0, st);       // restriction r_c = P^T r
  if (rc) return rc;
  if (pc->n_bc_c > 0) {
    k_zero_list<<<vec_grid(pc->n_bc_c), RED_THREADS, 0, st>>>(pc->bc_c, pc->n_bc_c, pc->rc);
    count_launch(1);
  }
  // fine block solves and the coarse solve share one cooperative launch
  rc = gf_schwarz_apply2(pc->fine, r, z, n, pc->coarse, pc->rc, pc->zc, pc->Rt.nrows, st);
  if (rc) return rc;
  if (pc->dist && pc->dist->n_ranges > 0) {                   // sum the ranks' block contributions
    rc = dist_allreduce(pc->dist, z, n, st);
    if (rc) return rc;
  }
  return gf_spmv(&pc->P, pc->zc, z, 1.0, 1.0, st);            // z += P z_c (coarse level is replicated)
}

// ---------------------------------------------------------------- r = b - A x in double-double ----
// Iterative refinement needs the residual of the CURRENT iterate, and on these systems (kappa up to 1e12) a plain FP64
// evaluation of b - A x is itself only good to ~eps |A||x| / |b| ~ 2e-9: the products are formed exactly with an FMA
// (two-product) and accumulated as an unevaluated (hi, lo) pair (two-sum), so the returned FP64 residual is correctly
// rounded to ~1e-30 |A||x|.  Memory bound like the SpMV it replaces (the extra flops are free).
namespace gf {
__device__ __forceinline__ void dd_add(double& hi, double& lo, double v) {       // (hi, lo) += v
  const double s = hi + v;
  const double bb = s - hi;
  lo += (hi - (s - bb)) + (v - bb);
  hi = s;
}
__global__ void __launch_bounds__(256)
k_residual_dd(GfCsr A, int64_t row_begin, int64_t row_end, const double* __restrict__ x, const double* __restrict__ b,
              double* __restrict__ r) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = row_begin + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); row < row_end; row += nwarps) {
    const int64_t s = A.indptr[row], e = A.indptr[row + 1];
    double hi = 0.0, lo = 0.0;
    for (int64_t k = s + lane; k < e; k += 32) {
      const double a = A.vals[k], xv = __ldg(x + A.indices[k]);
      const double p = a * xv;
      const double pe = fma(a, xv, -p);            // exact product = p + pe
      dd_add(hi, lo, p);
      lo += pe;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const double ohi = __shfl_xor_sync(0xffffffffu, hi, o), olo = __shfl_xor_sync(0xffffffffu, lo, o);
      dd_add(hi, lo, ohi);
      lo += olo;
    }
    if (lane == 0) {
      double rh = b[row], rl = 0.0;
      dd_add(rh, rl, -hi);
      r[row] = rh + (rl - lo);
    }
  }
}
}  // namespace gf

extern "C" int gf_residual_dd(const GfCsr* A, const GfDist* dist, const double* x, const double* b, double* r, void* stream) {
  if (!A || !x || !b || !r) return set_error(GF_ERR_BADARG, "gf_residual_dd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = A->nrows;
  if (dist && dist->n_ranges > 0) {
    cudaError_t e = cudaMemsetAsync(r, 0, (size_t)n * sizeof(double), st);
    if (e != cudaSuccess) return set_cuda_error(e, "gf_residual_dd memset");
    for (int q = 0; q < dist->n_ranges; ++q) {
      const int64_t b0 = dist->ranges_h[2 * q], b1 = dist->ranges_h[2 * q + 1];
      if (b1 <= b0) continue;
      k_residual_dd<<<spmv_grid(b1 - b0), 256, 0, st>>>(*A, b0, b1, x, b, r);
      count_launch(1);
    }
    return dist_allreduce(dist, r, n, st);          // disjoint rows + zeros: the sum is exact
  }
  k_residual_dd<<<spmv_grid(n), 256, 0, st>>>(*A, 0, n, x, b, r);
  return check_launch("k_residual_dd");
}

extern "C" int gf_pcg(const GfCsr* A, const double* b, double* x, const GfPcgWork* w, const GfPrecond* pre,
                      const GfDist* dist, double rtol, double atol, int max_it, int check_every, int* iters,
                      double* relres, void* stream) {
  if (!A || !b || !x || !w) return set_error(GF_ERR_BADARG, "gf_pcg: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = A->nrows;
  const int gv = vec_grid(n);
  double* part1 = w->partial;                 // pAp partials [MAX_PARTIAL]
  double* part2 = w->partial + MAX_PARTIAL;   // (rz, rr) partials [2*MAX_PARTIAL]
  if (check_every < 1) check_every = 1;
  k_pcg_init<<<gv, RED_THREADS, 0, st>>>(n, b, w->dinv, x, w->r, w->z, w->p, part2);
  if (pre) {
    int rc0 = gf_precond_apply(pre, w->r, w->z, n, st);
    if (rc0) return rc0;
    k_pcg_rz<<<gv, RED_THREADS, 0, st>>>(n, w->r, w->z, part2);
    k_pcg_init_p<<<gv, RED_THREADS, 0, st>>>(n, w->z, w->p);
    count_launch(2);
  }
  k_pcg_init_fin<<<1, RED_THREADS, 0, st>>>(part2, gv, w->scal);
  count_launch(2);
  cudaError_t e = cudaMemcpyAsync(w->scal_h, w->scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_cuda_error(e, "gf_pcg init");
  const double bb = w->scal_h[5];
  int it = 0;
  double rel = 1.0;
  if (!(bb > 0.0)) {  // zero right-hand side: x = 0
    if (iters) *iters = 0;
    if (relres) *relres = 0.0;
    return (bb == 0.0) ? GF_OK : set_error(GF_ERR_NAN, "gf_pcg: right-hand side is not finite");
  }
  const double bnorm = sqrt(bb);
  int rc = GF_ERR_NOCONV;
  while (it < max_it) {
    const int parity = it & 1;
    int npart1 = 0;
    int rca = apply_A(*A, dist, &w->nodes, w->p, w->Ap, w->p, part1, &npart1, st);
    if (rca) return rca;
    k_pcg_update<<<gv, RED_THREADS, 0, st>>>(n, w->p, w->Ap, w->dinv, x, w->r, w->z, part1, npart1, w->scal,
                                             parity, part2);
    if (pre) {
      int rc1 = gf_precond_apply(pre, w->r, w->z, n, st);
      if (rc1) return rc1;
      k_pcg_rz<<<gv, RED_THREADS, 0, st>>>(n, w->r, w->z, part2);
      count_launch(1);
    }
    k_pcg_dir<<<gv, RED_THREADS, 0, st>>>(n, w->z, w->p, part2, gv, w->scal, parity);
    count_launch(2);
    ++it;
    if (it % check_every == 0 || it == max_it) {
      e = cudaMemcpyAsync(w->scal_h, w->scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return set_cuda_error(e, "gf_pcg iteration");
      const double rr = w->scal_h[4], pAp = w->scal_h[6];
      if (!(rr == rr) || !(pAp == pAp)) { rc = set_error(GF_ERR_NAN, "gf_pcg: NaN in recurrence"); break; }
      rel = sqrt(rr) / bnorm;
      // convergence first: once r = 0 exactly (tiny systems, right-hand sides supported on identity rows)
      // the later directions are p = 0 and p.Ap = 0 without any loss of definiteness
      if (rel < rtol || sqrt(rr) < atol) { rc = GF_OK; break; }
      if (!(pAp > 0.0)) { rc = set_error(GF_ERR_BREAKDOWN, "gf_pcg: p.Ap <= 0 (matrix not SPD)"); break; }
    }
  }
  if (iters) *iters = it;
  if (relres) *relres = rel;
  if (rc == GF_ERR_NOCONV) set_error(rc, "gf_pcg: tolerance not reached within max_it");
  return rc;
}

// ------------------------------------------------------------------ GMRES ----
// Flexible (right-preconditioned) restarted GMRES(m) with the same preconditioner: the fallback of the Krylov path when the
// tangent is not positive definite (CG breaks down; the reference's LU, utils/opt_utils.py:176, still returns a
// step).  Classical Gram-Schmidt applied twice (two fused multi-dot / multi-axpy passes per iteration, fixed
// reduction trees); the (m+1) x m Hessenberg least-squares problem is updated on the host with Givens rotations.
namespace gf {
constexpr int GM_CHUNK = 8;
// partial[blk][j] = sum over the CTA's entries of V_j . w, j in [0, k)
__global__ void __launch_bounds__(RED_THREADS)
k_mdot(int64_t n, const double* __restrict__ V, int64_t ld, int k, const double* __restrict__ w, double* partial) {
  __shared__ double sh[32];
  for (int j0 = 0; j0 < k; j0 += GM_CHUNK) {
    double acc[GM_CHUNK];
#pragma unroll
    for (int q = 0; q < GM_CHUNK; ++q) acc[q] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const double wi = w[i];
#pragma unroll
      for (int q = 0; q < GM_CHUNK; ++q)
        if (j0 + q < k) acc[q] = fma(V[(int64_t)(j0 + q) * ld + i], wi, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < GM_CHUNK; ++q) {
      if (j0 + q < k) {
        const double sacc = block_sum(acc[q], sh);
        if (threadIdx.x == 0) partial[(size_t)blockIdx.x * k + j0 + q] = sacc;
      }
    }
  }
}
// h[j] (+)= sum_blk partial[blk][j]
__global__ void __launch_bounds__(RED_THREADS)
k_mdot_fin(const double* partial, int nblk, int k, double* h, int accumulate, double* hlast) {
  __shared__ double sh[32];
  for (int j = 0; j < k; ++j) {
    const double sacc = sum_partials(partial + j, nblk, k, sh);
    if (threadIdx.x == 0) { h[j] = accumulate ? h[j] + sacc : sacc; hlast[j] = sacc; }
  }
}
// w -= sum_j c[j] V_j
__global__ void __launch_bounds__(RED_THREADS)
k_maxpy(int64_t n, const double* __restrict__ V, int64_t ld, int k, const double* __restrict__ c, double sign,
        double* __restrict__ w, int overwrite) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double acc = overwrite ? 0.0 : w[i];
    for (int j = 0; j < k; ++j) acc = fma(sign * c[j], V[(int64_t)j * ld + i], acc);
    w[i] = acc;
  }
}
// out = in * (1 / sqrt(*nrm2))
__global__ void __launch_bounds__(RED_THREADS)
k_scale_inv_norm(int64_t n, const double* __restrict__ in, const double* nrm2, double* __restrict__ out) {
  const double s = (*nrm2 > 0.0) ? 1.0 / sqrt(*nrm2) : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[i] * s;
}
}  // namespace gf

extern "C" int gf_gmres(const GfCsr* A, const double* b, double* x, const GfGmresWork* w, const GfPrecond* pre,
                        const GfDist* dist, double rtol, int restart, int max_it, int* iters, double* relres,
                        void* stream) {
  if (!A || !b || !x || !w || restart < 1 || restart > 128) return set_error(GF_ERR_BADARG, "gf_gmres: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = A->nrows;
  const int gv = vec_grid(n), m = restart;
  double* V = w->V;            // [(m+1)][n]
  double* hd = w->hdev;        // [m+2] h column | [m+2] last pass | [1] norm^2
  double* hh = w->h_host;      // pinned mirror
  // host Hessenberg (column major, (m+1) x m), Givens cs/sn, rhs g
  double* H = (double*)malloc(sizeof(double) * (size_t)(m + 1) * m);
  double* cs = (double*)malloc(sizeof(double) * m), *sn = (double*)malloc(sizeof(double) * m);
  double* g = (double*)malloc(sizeof(double) * (m + 1)), *yv = (double*)malloc(sizeof(double) * (m + 1));
  int rc = GF_ERR_NOCONV, it = 0;
  double rel = 1.0, bnorm = 0.0;
  cudaError_t e = cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), st);
  bool first = true;
  auto fetch = [&](int cnt) -> int {
    cudaError_t e2 = cudaMemcpyAsync(hh, hd, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st);
    if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(st);
    return e2 == cudaSuccess ? GF_OK : set_cuda_error(e2, "gf_gmres fetch");
  };
  while (it < max_it && e == cudaSuccess) {
    // r = b - A x  (x = 0 in the first cycle)
    if (first) {
      e = cudaMemcpyAsync(w->t, b, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st);
    } else {
      int r1 = apply_A(*A, dist, &w->nodes, x, w->t, nullptr, nullptr, nullptr, st);
      if (r1) { rc = r1; break; }
      k_axpby<<<gv, RED_THREADS, 0, st>>>(n, 1.0, b, -1.0, w->t);
      count_launch(1);
    }
    k_dot_partial<<<gv, RED_THREADS, 0, st>>>(n, w->t, w->t, w->partial);
    k_finalize<<<1, RED_THREADS, 0, st>>>(w->partial, gv, 1, hd + 2 * (m + 2));
    k_scale_inv_norm<<<gv, RED_THREADS, 0, st>>>(n, w->t, hd + 2 * (m + 2), V);
    count_launch(3);
    if ((rc = fetch(2 * (m + 2) + 1)) != GF_OK) break;
    const double beta = sqrt(hh[2 * (m + 2)]);
    if (first) {
      bnorm = beta; first = false;
      if (!(bnorm > 0.0)) { rc = (bnorm == 0.0) ? GF_OK : set_error(GF_ERR_NAN, "gf_gmres: right-hand side is not finite"); rel = 0.0; break; }
    }
    rel = beta / bnorm;
    rc = GF_ERR_NOCONV;
    if (rel < rtol) { rc = GF_OK; break; }
    for (int i = 0; i <= m; ++i) g[i] = 0.0;
    g[0] = beta;
    int j = 0;
    for (; j < m && it < max_it; ++j) {
      double* vj = V + (int64_t)j * n; double* wv = V + (int64_t)(j + 1) * n;
      const double* zsrc = vj;
      if (pre) {
        // FLEXIBLE variant: z_j = M^-1 v_j is kept.  The preconditioner multiplies FP32 panels, so it is linear
        // only to ~1e-7; re-applying it to the combination V y (plain right preconditioning) would add an error
        // of that relative size times |V y|, far above the residual on these kappa ~ 1e10 systems.
        double* zj = w->Z + (int64_t)j * n;
        int r1 = gf_precond_apply(pre, vj, zj, n, st);
        if (r1) { rc = r1; goto done; }
        zsrc = zj;
      }
      { int r1 = apply_A(*A, dist, &w->nodes, zsrc, wv, nullptr, nullptr, nullptr, st); if (r1) { rc = r1; goto done; } }
      // CGS2: h = V^T w, w -= V h, twice (second pass accumulates into h)
      for (int pass = 0; pass < 2; ++pass) {
        k_mdot<<<gv, RED_THREADS, 0, st>>>(n, V, n, j + 1, wv, w->partial);
        k_mdot_fin<<<1, RED_THREADS, 0, st>>>(w->partial, gv, j + 1, hd, pass, hd + (m + 2));
        k_maxpy<<<gv, RED_THREADS, 0, st>>>(n, V, n, j + 1, hd + (m + 2), -1.0, wv, 0);
        count_launch(3);
      }
      k_dot_partial<<<gv, RED_THREADS, 0, st>>>(n, wv, wv, w->partial);
      k_finalize<<<1, RED_THREADS, 0, st>>>(w->partial, gv, 1, hd + 2 * (m + 2));
      k_scale_inv_norm<<<gv, RED_THREADS, 0, st>>>(n, wv, hd + 2 * (m + 2), wv);
      count_launch(3);
      if ((rc = fetch(2 * (m + 2) + 1)) != GF_OK) goto done;
      rc = GF_ERR_NOCONV;
      double* Hj = H + (size_t)j * (m + 1);
      for (int i = 0; i <= j; ++i) Hj[i] = hh[i];
      Hj[j + 1] = sqrt(hh[2 * (m + 2)]);
      if (!(Hj[j + 1] == Hj[j + 1])) { rc = set_error(GF_ERR_NAN, "gf_gmres: NaN in Arnoldi"); goto done; }
      for (int i = 0; i < j; ++i) {        // previous rotations
        const double t0 = cs[i] * Hj[i] + sn[i] * Hj[i + 1];
        Hj[i + 1] = -sn[i] * Hj[i] + cs[i] * Hj[i + 1]; Hj[i] = t0;
      }
      const double den = hypot(Hj[j], Hj[j + 1]);
      cs[j] = den > 0 ? Hj[j] / den : 1.0; sn[j] = den > 0 ? Hj[j + 1] / den : 0.0;
      Hj[j] = den; Hj[j + 1] = 0.0;
      g[j + 1] = -sn[j] * g[j]; g[j] = cs[j] * g[j];
      ++it;
      rel = fabs(g[j + 1]) / bnorm;
      if (rel < rtol) { ++j; break; }
    }
    // y = H^-1 g (upper triangular), x += M^-1 (V y)
    for (int i = j - 1; i >= 0; --i) {
      double sacc = g[i];
      for (int q = i + 1; q < j; ++q) sacc -= H[(size_t)q * (m + 1) + i] * yv[q];
      yv[i] = sacc / H[(size_t)i * (m + 1) + i];
    }
    for (int i = 0; i < j; ++i) hh[i] = yv[i];
    e = cudaMemcpyAsync(hd, hh, sizeof(double) * j, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) break;
    k_maxpy<<<gv, RED_THREADS, 0, st>>>(n, pre ? w->Z : V, n, j, hd, 1.0, x, 0);     // x += Z y  (V y without preconditioner)
    count_launch(1);
    e = cudaStreamSynchronize(st);         // hh is reused by the next cycle
    if (rel < rtol) { rc = GF_OK; break; }
  }
done:
  free(H); free(cs); free(sn); free(g); free(yv);
  if (e != cudaSuccess) return set_cuda_error(e, "gf_gmres");
  if (iters) *iters = it;
  if (relres) *relres = rel;
  if (rc == GF_ERR_NOCONV) set_error(rc, "gf_gmres: tolerance not reached within max_it");
  return rc;
}
