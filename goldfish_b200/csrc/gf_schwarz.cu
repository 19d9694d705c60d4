// Overlapping additive-Schwarz preconditioner with exact block solves (sm_100a, FP64).
//
// PENGoLINS' iterative option is CG + PCFIELDSPLIT (additive) with per-patch LU
// sub-solves.  With the penalty coefficient of the
// fixtures (alpha = 1e3) non-overlapping patch blocks stall (thousands of
// iterations: the stiff interface springs straddle two blocks).  Letting every
// patch block overlap the neighbouring patches by a few control-point layers
// puts each spring inside a block; CG then needs tens of iterations.
//
// One block per patch:  nodes = patch CPs + overlap layers, ordered so that the
// block matrix is banded (host: goldfish_b200/schwarz.py); 3 dofs per node
// interleaved.  The band is stored as dense nb x nb blocks, block column j
// holding blocks (j+k, j), k = 0..mb.
//   factor : right-looking block Cholesky, 3 kernels per block column,
//            all patch blocks batched in gridDim.y
//   solve  : ONE cooperative kernel, G CTAs per patch block walk the block
//            columns with a per-group global-memory barrier per step
//   z = sum_i R_i^T (L_i L_i^T)^-1 R_i r  assembled by a fixed-order gather.
#include "gf_common.cuh"

namespace gf {

constexpr int NB = 64;
constexpr int NB2 = NB * NB;

// block (row j+k, col j) of patch block i; panel heights vary per block column
// (taller next to intersections that run along the slow direction)
__device__ __forceinline__ double* sw_block(const GfSchwarz& S, int i, int j, int k) {
  return S.band + S.off_col[S.off_j[i] + j] + (size_t)k * NB2;
}
__device__ __forceinline__ int sw_mb(const GfSchwarz& S, int i, int j) { return S.mbj[S.off_j[i] + j]; }

// ---- fill the band from K ---------------------------------------------------
__global__ void __launch_bounds__(256)
k_sw_fill(GfSchwarz S, GfCsr K) {
  const int i = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int lr = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (lr >= S.n_pad[i]) return;
  const int32_t* glob = S.glob + S.off_y[i];
  const int32_t* loc = S.loc + (size_t)i * K.nrows;
  const int r = glob[lr];
  const int br = lr / NB, rr = lr % NB;
  if (r < 0) {  // padding dof: identity
    if (lane == 0) sw_block(S, i, br, 0)[rr * NB + rr] = 1.0;
    return;
  }
  for (int64_t k = K.indptr[r] + lane; k < K.indptr[r + 1]; k += 32) {
    const int lc = loc[K.indices[k]];
    if (lc < 0 || lc > lr) continue;
    const int bj = lc / NB, cc = lc % NB;
    sw_block(S, i, bj, br - bj)[rr * NB + cc] = K.vals[k];
  }
}

// ---- diagonal block: Cholesky + inverse of the factor -------------------------
__global__ void __launch_bounds__(256)
k_sw_potrf(GfSchwarz S, int j) {
  const int i = blockIdx.y;
  if (j >= S.nbr[i]) return;
  extern __shared__ double sm[];
  double (*A)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm);
  double (*Li)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm + NB * (NB + 1));
  double* blk = sw_block(S, i, j, 0);
  const int tid = threadIdx.x;
  for (int e = tid; e < NB2; e += 256) { A[e / NB][e % NB] = blk[e]; Li[e / NB][e % NB] = 0.0; }
  __syncthreads();
  for (int c = 0; c < NB; ++c) {
    if (tid == 0) {
      const double d = A[c][c];
      if (!(d > 0.0)) { atomicExch(S.flag, 1); A[c][c] = 1.0; } else A[c][c] = sqrt(d);
    }
    __syncthreads();
    const double inv = 1.0 / A[c][c];
    for (int r = c + 1 + tid; r < NB; r += 256) A[r][c] *= inv;
    __syncthreads();
    // trailing update of the lower triangle
    const int m = NB - c - 1;
    for (int e = tid; e < m * m; e += 256) {
      const int r = c + 1 + e / m, cc = c + 1 + e % m;
      if (cc <= r) A[r][cc] -= A[r][c] * A[cc][c];
    }
    __syncthreads();
  }
  // inverse of L: thread t solves L x = e_t
  if (tid < NB) {
    const int t = tid;
    Li[t][t] = 1.0 / A[t][t];
    for (int r = t + 1; r < NB; ++r) {
      double s = 0.0;
      for (int m = t; m < r; ++m) s = fma(A[r][m], Li[m][t], s);
      Li[r][t] = -s / A[r][r];
    }
  }
  __syncthreads();
  double* inv = S.invd + S.off_inv[i] + (size_t)j * NB2;
  for (int e = tid; e < NB2; e += 256) {
    const int r = e / NB, c = e % NB;
    blk[e] = (c <= r) ? A[r][c] : 0.0;
    inv[e] = (c <= r) ? Li[r][c] : 0.0;
  }
}

// ---- panel: B_k <- B_k L_jj^-T ---------------------------------------------------
__global__ void __launch_bounds__(256)
k_sw_trsm(GfSchwarz S, int j) {
  const int i = blockIdx.y, k = blockIdx.x + 1;
  if (j >= S.nbr[i] || k > sw_mb(S, i, j)) return;
  extern __shared__ double sm[];
  double (*B)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm);
  double (*Li)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm + NB * (NB + 1));
  double* blk = sw_block(S, i, j, k);
  const double* inv = S.invd + S.off_inv[i] + (size_t)j * NB2;
  const int tid = threadIdx.x;
  for (int e = tid; e < NB2; e += 256) { B[e / NB][e % NB] = blk[e]; Li[e / NB][e % NB] = inv[e]; }
  __syncthreads();
  // out[r][c] = sum_{m<=c} B[r][m] Linv[c][m]
  const int r0 = (tid / 16) * 4, c0 = (tid % 16) * 4;
  double acc[4][4] = {};
  for (int m = 0; m < NB; ++m) {
    double a[4], b[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) { a[x] = B[r0 + x][m]; b[x] = Li[c0 + x][m]; }
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
  }
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) blk[(r0 + x) * NB + c0 + y] = acc[x][y];
}

// ---- trailing update: A(j+k1, j+k2) -= B_k1 B_k2^T ------------------------------
__global__ void __launch_bounds__(256)
k_sw_update(GfSchwarz S, int j) {
  const int i = blockIdx.y;
  if (j >= S.nbr[i]) return;
  const int mb = sw_mb(S, i, j);
  // pair index -> (k1 >= k2 >= 1)
  int p = blockIdx.x, k1 = 1;
  while (p >= k1) { p -= k1; ++k1; }
  const int k2 = p + 1;
  if (k1 > mb) return;
  extern __shared__ double sm[];
  double (*B1)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm);
  double (*B2)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm + NB * (NB + 1));
  const double* b1 = sw_block(S, i, j, k1);
  const double* b2 = sw_block(S, i, j, k2);
  const int tid = threadIdx.x;
  for (int e = tid; e < NB2; e += 256) { B1[e / NB][e % NB] = b1[e]; B2[e / NB][e % NB] = b2[e]; }
  __syncthreads();
  const int r0 = (tid / 16) * 4, c0 = (tid % 16) * 4;
  double acc[4][4] = {};
#pragma unroll 8
  for (int m = 0; m < NB; ++m) {
    double a[4], b[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) { a[x] = B1[r0 + x][m]; b[x] = B2[c0 + x][m]; }
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
  }
  double* C = sw_block(S, i, j + k2, k1 - k2);
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) C[(r0 + x) * NB + c0 + y] -= acc[x][y];
}

// ---- restriction / prolongation ----------------------------------------------------
__global__ void k_sw_gather_in(GfSchwarz S, const double* __restrict__ r) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < S.n_y; t += (int64_t)gridDim.x * blockDim.x) {
    const int g = S.glob[t];
    S.y[t] = (g >= 0) ? r[g] : 0.0;
  }
}
__global__ void k_sw_gather_out(GfSchwarz S, double* __restrict__ z, int64_t n) {
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < n; d += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int64_t t = S.zptr[d]; t < S.zptr[d + 1]; ++t) s += S.y[S.zsrc[t]];
    z[d] = s;
  }
}

// ---- the banded triangular solves ------------------------------------------------------
__device__ __forceinline__ void group_barrier(unsigned* cnt, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(cnt, 1u);
    while (*((volatile unsigned*)cnt) < target) {}
    __threadfence();
  }
  __syncthreads();
}

// y_blk (64) <- M y_blk with M = Linv (forward) or Linv^T (backward); result in xs (smem)
__device__ __forceinline__ void diag_apply(const double* __restrict__ inv, const double* ys, double* xs, bool transpose) {
  // 256 threads: 4 threads per row
  const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
  double s = 0.0;
  if (!transpose) {
    for (int c = q; c <= r; c += 4) s = fma(inv[r * NB + c], ys[c], s);
  } else {
    for (int c = r + q; c < NB; c += 4) s = fma(inv[c * NB + r], ys[c], s);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (q == 0) xs[r] = s;
}

__global__ void __launch_bounds__(256)
k_sw_solve(GfSchwarz S, int G) {
  const int i = blockIdx.x / G, cta = blockIdx.x % G;
  __shared__ double ys[NB], xs[NB];
  __shared__ double red[8][NB];
  const int nbr = S.nbr[i];
  const int32_t* mbj = S.mbj + S.off_j[i];
  const int32_t* rlen = S.rlen + S.off_j[i];
  double* y = S.y + S.off_y[i];
  const double* invd = S.invd + S.off_inv[i];
  unsigned* cnt = S.barrier + i;
  unsigned step = 0;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  // forward: L x = y
  for (int j = 0; j < nbr; ++j) {
    if (tid < NB) ys[tid] = __ldcg(y + (size_t)j * NB + tid);
    __syncthreads();
    diag_apply(invd + (size_t)j * NB2, ys, xs, false);
    __syncthreads();
    if (cta == 0 && tid < NB) __stcg(y + (size_t)j * NB + tid, xs[tid]);
    const double x0 = xs[lane], x1 = xs[lane + 32];
    for (int k = 1 + cta; k <= mbj[j]; k += G) {
      // y_{j+k} -= L(j+k,j) x_j : warp w owns rows 8w..8w+7, lanes run along the row (coalesced)
      const double* L = sw_block(S, i, j, k) + (size_t)(w * 8) * NB;
      double s[8];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) s[rr] = __ldcs(L + rr * NB + lane) * x0 + __ldcs(L + rr * NB + lane + 32) * x1;
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) s[rr] = warp_sum(s[rr]);
      if (lane < 8) {
        double v = s[0];
#pragma unroll
        for (int rr = 1; rr < 8; ++rr) v = (lane == rr) ? s[rr] : v;
        double* dst = y + (size_t)(j + k) * NB + w * 8 + lane;
        __stcg(dst, __ldcg(dst) - v);
      }
    }
    ++step;
    group_barrier(cnt, step * (unsigned)G);
  }
  // backward: L^T x = y
  for (int j = nbr - 1; j >= 0; --j) {
    if (tid < NB) ys[tid] = __ldcg(y + (size_t)j * NB + tid);
    __syncthreads();
    diag_apply(invd + (size_t)j * NB2, ys, xs, true);
    __syncthreads();
    if (cta == 0 && tid < NB) __stcg(y + (size_t)j * NB + tid, xs[tid]);
    for (int k = 1 + cta; k <= rlen[j]; k += G) {
      // y_{j-k} -= L(j, j-k)^T x_j ; block (row j, col j-k) is panel j-k, offset k.
      // warp w sums its 8 rows for columns lane, lane+32; then the 8 warps are combined.
      const double* L = sw_block(S, i, j - k, k) + (size_t)(w * 8) * NB;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        const double xr = xs[w * 8 + rr];
        a0 = fma(__ldcs(L + rr * NB + lane), xr, a0);
        a1 = fma(__ldcs(L + rr * NB + lane + 32), xr, a1);
      }
      __syncthreads();
      red[w][lane] = a0; red[w][lane + 32] = a1;
      __syncthreads();
      if (tid < NB) {
        double v = 0.0;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) v += red[ww][tid];
        double* dst = y + (size_t)(j - k) * NB + tid;
        __stcg(dst, __ldcg(dst) - v);
      }
    }
    ++step;
    group_barrier(cnt, step * (unsigned)G);
  }
}

__global__ void __launch_bounds__(256)
k_dot_slot0(int64_t n, const double* x, const double* y, double* partial2) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s = fma(x[i], y[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partial2[2 * blockIdx.x] = t;
  }
}

}  // namespace gf

using namespace gf;

extern "C" int gf_schwarz_factor(const GfSchwarz* S, const GfCsr* K, void* stream) {
  if (!S || !K) return set_error(GF_ERR_BADARG, "gf_schwarz_factor: null argument");
  if (S->nb != NB) return set_error(GF_ERR_BADARG, "gf_schwarz_factor: nb must be 64");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(S->band, 0, (size_t)S->band_len * sizeof(double), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(S->flag, 0, sizeof(int), st);
  if (e != cudaSuccess) return set_cuda_error(e, "gf_schwarz_factor memset");
  dim3 gf((S->max_n_pad + 7) / 8, S->nblocks);
  k_sw_fill<<<gf, 256, 0, st>>>(*S, *K);
  const size_t smem = 2 * NB * (NB + 1) * sizeof(double);
  e = cudaFuncSetAttribute(k_sw_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sw_potrf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sw_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(k_sw_*)");
  for (int j = 0; j < S->max_nbr; ++j) {
    k_sw_potrf<<<dim3(1, S->nblocks), 256, smem, st>>>(*S, j);
    const int m = S->step_mb_h[j];      // tallest panel of this step over all patch blocks
    if (m > 0) {
      k_sw_trsm<<<dim3(m, S->nblocks), 256, smem, st>>>(*S, j);
      k_sw_update<<<dim3(m * (m + 1) / 2, S->nblocks), 256, smem, st>>>(*S, j);
    }
    count_launch(3);
  }
  int flag = 0;
  e = cudaMemcpyAsync(&flag, S->flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_cuda_error(e, "gf_schwarz_factor");
  if (flag) return set_error(GF_ERR_BREAKDOWN, "gf_schwarz_factor: a block is not positive definite");
  return check_launch("gf_schwarz_factor");
}

extern "C" int gf_schwarz_apply(const GfSchwarz* S, const double* r, double* z, int64_t n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int g = (int)((S->n_y + 255) / 256); if (g > 2048) g = 2048;
  k_sw_gather_in<<<g, 256, 0, st>>>(*S, r);
  cudaError_t e = cudaMemsetAsync(S->barrier, 0, sizeof(unsigned) * S->nblocks, st);
  if (e != cudaSuccess) return set_cuda_error(e, "gf_schwarz_apply memset");
  GfSchwarz Sv = *S;
  int G = S->ctas_per_block;
  void* args[] = {&Sv, &G};
  e = cudaLaunchCooperativeKernel((void*)k_sw_solve, dim3(S->nblocks * G), dim3(256), args, 0, st);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaLaunchCooperativeKernel(k_sw_solve)");
  int g2 = (int)((n + 255) / 256); if (g2 > 2048) g2 = 2048;
  k_sw_gather_out<<<g2, 256, 0, st>>>(*S, z, n);
  count_launch(2);
  return check_launch("gf_schwarz_apply");
}

extern "C" int gf_dot_slot0(int64_t n, const double* x, const double* y, double* partial2, int grid, void* stream) {
  k_dot_slot0<<<grid, 256, 0, (cudaStream_t)stream>>>(n, x, y, partial2);
  return check_launch("k_dot_slot0");
}
