// Overlapping additive-Schwarz preconditioner with exact block solves (sm_100a, FP64).
//
// PENGoLINS' iterative option is CG + PCFIELDSPLIT (additive) with per-patch LU
// sub-solves.  With the penalty coefficient of the
// fixtures (alpha = 1e3) non-overlapping patch blocks stall (thousands of
// iterations: the stiff interface springs straddle two blocks).  Letting every
// patch block overlap the neighbouring patches by a few control-point layers
// puts each spring inside a block; CG then needs tens of iterations.
//
// One block per sub-domain (a rectangle of a patch's control net):  nodes = own CPs + overlap layers,
// ordered so that the block matrix is banded (host: goldfish_b200/schwarz.py); 3 dofs per node
// interleaved.  The band is stored as dense nb x nb blocks, block column j holding blocks (j+k, j),
// k = 0..mb_j.
//   factor : right-looking block Cholesky, 3 kernels per block column, all blocks batched in gridDim.y,
//            then "solve form"  M(j+k,j) = L(j+k,j) L_jj^-1,  D_j = (L_jj L_jj^T)^-1  (+ FP32 copy of M)
//   solve  : k_sw_solve1          one CTA per fine block, block vector in shared memory (HBM-bound stream)
//            k_sw_coarse_cluster  the single coarse block on a 16-CTA thread-block cluster (DSMEM broadcast),
//                                 concurrent with the fine sweeps
//            k_sw_solve           fallback: G CTAs per block with a per-group global-memory barrier per step
//   z = sum_i R_i^T (L_i L_i^T)^-1 R_i r  assembled by a fixed-order gather.
#include "gf_common.cuh"
#include <stdlib.h>
#include <cuda_pipeline.h>
#include <cooperative_groups.h>

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
  const int32_t* gs = S.gs + S.off_g[i];
  const int32_t* ls = S.ls + S.off_g[i];
  const int ng = (int)(S.off_g[i + 1] - S.off_g[i]);
  const int r = glob[lr];
  const int br = lr / NB, rr = lr % NB;
  if (r < 0) {  // padding dof: identity
    if (lane == 0) sw_block(S, i, br, 0)[rr * NB + rr] = 1.0;
    return;
  }
  for (int64_t k = K.indptr[r] + lane; k < K.indptr[r + 1]; k += 32) {
    // global column -> local index of this block (binary search in the sorted dof list)
    const int c = K.indices[k];
    int lo = 0, hi = ng;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (gs[mid] < c) lo = mid + 1; else hi = mid; }
    if (lo >= ng || gs[lo] != c) continue;
    const int lc = ls[lo];
    if (lc > lr) continue;
    const int bj = lc / NB, cc = lc % NB;
    sw_block(S, i, bj, br - bj)[rr * NB + cc] = K.vals[k];
  }
}

// ---- in-shared-memory Cholesky of a 64 x 64 block (lower), all 256 threads ------
__device__ __forceinline__ void chol64(double (*A)[NB + 1], int32_t* flag) {
  // right-looking, 256 threads: thread (r = tid/4, q = tid%4) owns row r, columns 16q..16q+15
  const int tid = threadIdx.x, r = tid >> 2, q = tid & 3;
  for (int c = 0; c < NB; ++c) {
    if (tid == 0) {
      const double d = A[c][c];
      if (!(d > 0.0)) { atomicExch(flag, 1); A[c][c] = 1.0; } else A[c][c] = sqrt(d);
    }
    __syncthreads();
    if (q == 0 && r > c) A[r][c] *= 1.0 / A[c][c];
    __syncthreads();
    if (r > c) {
      const double lrc = A[r][c];
      const int lo = max(c + 1, 16 * q), hi = min(r, 16 * q + 15);
      for (int cc = lo; cc <= hi; ++cc) A[r][cc] = fma(-lrc, A[cc][c], A[r][cc]);
    }
    __syncthreads();
  }
}

// ---- diagonal block of step j: Cholesky and the inverse of the factor, one CTA per block ----
__global__ void __launch_bounds__(256)
k_sw_potrf(GfSchwarz S, int j) {
  const int i = blockIdx.x;
  if (j >= S.nbr[i]) return;
  extern __shared__ double sm[];
  double (*A)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm);
  double (*Li)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm + NB * (NB + 1));
  const double* dblk = sw_block(S, i, j, 0);
  const int tid = threadIdx.x;
  for (int e = tid; e < NB2; e += 256) A[e / NB][e % NB] = dblk[e];
  __syncthreads();
  chol64(A, S.flag);
  // inverse of L: column t by 4 threads (t = tid/4), rows sequential, dot products split 4 ways
  {
    const int t = tid >> 2, q = tid & 3;
    if (q == 0) Li[t][t] = 1.0 / A[t][t];
    __syncwarp();
    for (int r = 1; r < NB; ++r) {          // uniform trip count: the shuffles need the whole warp
      double sdot = 0.0;
      if (r > t) for (int m = t + q; m < r; m += 4) sdot = fma(A[r][m], Li[m][t], sdot);
      sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
      sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
      if (q == 0 && r > t) Li[r][t] = -sdot / A[r][r];
      __syncwarp();
    }
  }
  __syncthreads();
  double* inv = S.invd + S.off_inv[i] + (size_t)j * NB2;
  for (int e = tid; e < NB2; e += 256) { const int r = e / NB, c = e % NB; inv[e] = (c <= r) ? Li[r][c] : 0.0; }
}

// ---- panel of step j: B_k <- B_k L_jj^-T as a 64^3 product with the inverted factor ----
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
  const int r0 = (tid / 16) * 4, c0 = (tid % 16) * 4;   // out[r][c] = sum_{m<=c} B[r][m] Linv[c][m]
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
    // release-increment / acquire-spin at gpu scope (bar.sync makes the CTA's stores
    // happen-before the release; cumulativity publishes them to the other CTAs)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cnt) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

// ---- after the factorisation: turn the factor into "solve form" -----------------
//   M(j+k, j) = L(j+k, j) L_jj^-1        (k >= 1)
//   D_j       = L_jj^-T L_jj^-1 = (A_jj - ...)^-1
// so that the triangular sweeps need no diagonal solve inside their dependency
// chain:  forward  y_{j+k} -= M(j+k,j) y_j ;  w_j = D_j y_j (parallel) ;
//         backward x_j = w_j - s_j,  s_{j-k} += M(j,j-k)^T x_j.
__global__ void __launch_bounds__(256)
k_sw_convert_panel(GfSchwarz S) {
  const int i = blockIdx.z, j = blockIdx.y, k = blockIdx.x + 1;
  if (j >= S.nbr[i] || k > sw_mb(S, i, j)) return;
  extern __shared__ double sm[];
  double (*B)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm);
  double (*Li)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm + NB * (NB + 1));
  double* blk = sw_block(S, i, j, k);
  const double* inv = S.invd + S.off_inv[i] + (size_t)j * NB2;
  const int tid = threadIdx.x;
  for (int e = tid; e < NB2; e += 256) { B[e / NB][e % NB] = blk[e]; Li[e % NB][e / NB] = inv[e]; }  // Li transposed
  __syncthreads();
  // out[r][c] = sum_m B[r][m] Linv[m][c] = sum_m B[r][m] LiT[c][m]
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

__global__ void __launch_bounds__(256)
k_sw_convert_diag(GfSchwarz S) {
  const int i = blockIdx.y, j = blockIdx.x;
  if (j >= S.nbr[i]) return;
  extern __shared__ double sm[];
  double (*Li)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm);
  double* inv = S.invd + S.off_inv[i] + (size_t)j * NB2;
  const int tid = threadIdx.x;
  for (int e = tid; e < NB2; e += 256) Li[e / NB][e % NB] = inv[e];
  __syncthreads();
  // D[r][c] = sum_m Linv[m][r] Linv[m][c],  m >= max(r, c)
  const int r0 = (tid / 16) * 4, c0 = (tid % 16) * 4;
  double acc[4][4] = {};
  for (int m = 0; m < NB; ++m) {
    double a[4], b[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) { a[x] = Li[m][r0 + x]; b[x] = Li[m][c0 + x]; }
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
  }
  __syncthreads();
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) inv[(r0 + x) * NB + c0 + y] = acc[x][y];
}

// The sweeps are bound by streaming the factor from HBM (it is read twice per
// preconditioner application), so the solve-form panels are also kept in FP32:
// a fixed, symmetric (M forward, M^T backward) and positive definite operator,
// which is all CG needs from a preconditioner.  Arithmetic stays FP64.
__global__ void k_sw_to_f32(const double* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (float)src[i];
}
__device__ __forceinline__ const float* sw_block32(const GfSchwarz& S, int i, int j, int k) {
  return S.band32 + S.off_col[S.off_j[i] + j] + (size_t)k * NB2;
}

#ifndef GF_PF
#define GF_PF 4
#endif
constexpr int PF = GF_PF;   // panel blocks of the NEXT step prefetched into shared memory per CTA
constexpr int LDS = NB + 4;  // padded row stride (floats) of a prefetched block: 16-B aligned, conflict-light
constexpr int PBLK = NB * LDS;

// cp.async the CTA's first PF panel blocks of step `j` (forward: column panel j, backward:
// row j of the factor) into shared memory; they land while the CTA waits at the group barrier.
__device__ __forceinline__ void sw_prefetch(const GfSchwarz& S, int i, int j, int k0, int G, int kmax,
                                            bool backward, float* pre) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int t = 0; t < PF; ++t) {
    const int k = k0 + t * G;
    if (k > kmax) break;
    const float* src = backward ? sw_block32(S, i, j - k, k) : sw_block32(S, i, j, k);
    float* dst = pre + t * PBLK;
#pragma unroll
    for (int e = tid; e < NB2 / 4; e += 256)                // 16-byte chunks: row e/16, chunk e%16
      __pipeline_memcpy_async(dst + (e >> 4) * LDS + (e & 15) * 4, src + e * 4, 16);
  }
  __pipeline_commit();
}

// The sweeps stream the factor once forward and once backward; measured on B200 they were
// bound by FP32->FP64 conversions (16/clk/SM) and FP64 shuffle reductions, not by HBM.  The
// block GEMVs therefore run in FP32 (FFMA, x converted once per step, 4 threads per row /
// column, 2-step reductions); results are accumulated into the FP64 vectors.
__device__ __forceinline__ void sw_group_solve(const GfSchwarz& S, const int i, const int cta, const int G, float* pre) {
  // pre: [PF][64][LDS] floats of dynamic shared memory
  __shared__ double xs[NB];
  __shared__ __align__(16) float xf[NB];
  __shared__ float red[4][NB];
  const int nbr = S.nbr[i];
  const int32_t* mbj = S.mbj + S.off_j[i];
  const int32_t* rlen = S.rlen + S.off_j[i];
  double* y = S.y + S.off_y[i];
  double* sv = S.s + S.off_y[i];
  const double* invd = S.invd + S.off_inv[i];
  unsigned* cnt = S.barrier + i;
  unsigned step = 0;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int dbg = S.debug_flags;               // timing experiments only: 1 = skip GEMVs, 2 = skip barriers
  const int k0 = (dbg & 1) ? 1000000 : 1 + cta;
  const int fr = tid >> 2, fq = tid & 3;           // forward: row fr, quarter fq (16 columns)
  const int bc = tid & 63, bq = tid >> 6;          // backward: column bc, quarter bq (16 rows)

  // ---------------- forward: y_{j+k} -= M(j+k, j) y_j ----------------
  if (nbr > 0) sw_prefetch(S, i, 0, k0, G, mbj[0], false, pre);
  for (int j = 0; j < nbr; ++j) {
    if (tid < NB) xf[tid] = (float)__ldcg(y + (size_t)j * NB + tid);
    __pipeline_wait_prior(0);
    __syncthreads();
    float xr[16];
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(xf + fq * 16 + c);
      xr[c] = v.x; xr[c + 1] = v.y; xr[c + 2] = v.z; xr[c + 3] = v.w;
    }
    int t = 0;
    for (int k = k0; k <= mbj[j]; k += G, ++t) {
      float a = 0.f;
      if (t < PF) {
        const float* p = pre + t * PBLK + fr * LDS + fq * 16;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          const float4 m = *reinterpret_cast<const float4*>(p + c);
          a = fmaf(m.x, xr[c], a); a = fmaf(m.y, xr[c + 1], a); a = fmaf(m.z, xr[c + 2], a); a = fmaf(m.w, xr[c + 3], a);
        }
      } else {
        const float4* p = reinterpret_cast<const float4*>(sw_block32(S, i, j, k) + fr * NB + fq * 16);
        const float4 m0 = __ldcs(p), m1 = __ldcs(p + 1), m2 = __ldcs(p + 2), m3 = __ldcs(p + 3);
        a = m0.x * xr[0] + m0.y * xr[1] + m0.z * xr[2] + m0.w * xr[3] + m1.x * xr[4] + m1.y * xr[5] + m1.z * xr[6] + m1.w * xr[7]
          + m2.x * xr[8] + m2.y * xr[9] + m2.z * xr[10] + m2.w * xr[11] + m3.x * xr[12] + m3.y * xr[13] + m3.z * xr[14] + m3.w * xr[15];
      }
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      // fire-and-forget reduction: every entry receives at most one contribution per step and
      // steps are separated by the group barrier, so the summation order is fixed.
      if (fq == 0) atomicAdd(y + (size_t)(j + k) * NB + fr, -(double)a);
    }
    __syncthreads();                                   // shared panel blocks consumed
    if (j + 1 < nbr) sw_prefetch(S, i, j + 1, k0, G, mbj[j + 1], false, pre);
    ++step;
    if (!(dbg & 2)) group_barrier(cnt, step * (unsigned)G);
  }
  // ---------------- diagonal: w_j = D_j y_j (FP64) ; s = 0 ----------------
  if (nbr > 0) sw_prefetch(S, i, nbr - 1, k0, G, rlen[nbr - 1], true, pre);
  for (int j = cta; j < nbr; j += G) {
    if (tid < NB) xs[tid] = __ldcg(y + (size_t)j * NB + tid);
    __syncthreads();
    const double* D = invd + (size_t)j * NB2 + (size_t)(w * 8) * NB;
    double sacc[8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) sacc[rr] = warp_sum(D[rr * NB + lane] * xs[lane] + D[rr * NB + lane + 32] * xs[lane + 32]);
    __syncthreads();
    if (lane < 8) {
      double v = sacc[0];
#pragma unroll
      for (int rr = 1; rr < 8; ++rr) v = (lane == rr) ? sacc[rr] : v;
      __stcg(y + (size_t)j * NB + w * 8 + lane, v);
      __stcg(sv + (size_t)j * NB + w * 8 + lane, 0.0);
    }
  }
  ++step;
  group_barrier(cnt, step * (unsigned)G);
  // ---------------- backward: x_j = w_j - s_j ; s_{j-k} += M(j, j-k)^T x_j ----------------
  for (int j = nbr - 1; j >= 0; --j) {
    if (tid < NB) {
      const double xv = __ldcg(y + (size_t)j * NB + tid) - __ldcg(sv + (size_t)j * NB + tid);
      xs[tid] = xv; xf[tid] = (float)xv;
    }
    __pipeline_wait_prior(0);
    __syncthreads();
    float xr[16];
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(xf + bq * 16 + c);
      xr[c] = v.x; xr[c + 1] = v.y; xr[c + 2] = v.z; xr[c + 3] = v.w;
    }
    int t = 0;
    for (int k = k0; k <= rlen[j]; k += G, ++t) {
      // block (row j, col j-k) lives in panel j-k at offset k; apply its transpose:
      // thread (column bc, quarter bq) sums 16 rows, the four quarters are combined in smem
      float a = 0.f;
      if (t < PF) {
        const float* p = pre + t * PBLK + (bq * 16) * LDS + bc;
#pragma unroll
        for (int r = 0; r < 16; ++r) a = fmaf(p[r * LDS], xr[r], a);
      } else {
        const float* p = sw_block32(S, i, j - k, k) + (bq * 16) * NB + bc;
        float m[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) m[r] = __ldcs(p + r * NB);
#pragma unroll
        for (int r = 0; r < 16; ++r) a = fmaf(m[r], xr[r], a);
      }
      __syncthreads();
      red[bq][bc] = a;
      __syncthreads();
      if (tid < NB) atomicAdd(sv + (size_t)(j - k) * NB + tid, (double)((red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid])));
    }
    __syncthreads();
    if (j > 0) sw_prefetch(S, i, j - 1, k0, G, rlen[j - 1], true, pre);
    ++step;
    if (!(dbg & 2)) group_barrier(cnt, step * (unsigned)G);
    // everyone has read w_j and s_j: the final x_j can now replace w_j
    if (cta == 0 && tid < NB) __stcg(y + (size_t)j * NB + tid, xs[tid]);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_sw_solve(GfSchwarz Sf, int G_f, int first_block, int nf, GfSchwarz Sc, int G_c) {
  extern __shared__ __align__(16) unsigned char sw_smem[];
  float* pre = reinterpret_cast<float*>(sw_smem);
  if ((int)blockIdx.x >= nf * G_f) sw_group_solve(Sc, 0, (int)blockIdx.x - nf * G_f, G_c, pre);
  else sw_group_solve(Sf, first_block + (int)blockIdx.x / G_f, (int)blockIdx.x % G_f, G_f, pre);
}

// ---- one CTA per block: the whole block-local vector lives in shared memory -------------
// With >= ~1 block per SM there is nothing to gain from splitting a block over CTAs: a single
// CTA needs no inter-CTA barrier (two bar.sync per step instead of an L2 round trip), no atomics
// and no s-vector; the factor is streamed exactly once per sweep, the next panels are pulled
// into L2 while the current step computes.  Backward runs in GATHER form over the same column
// panels the forward sweep uses:  x_j = w_j - sum_k M(j+k, j)^T x_{j+k}.
constexpr int XW = 32;        // ring of FP32 copies of the last XW solution blocks (needs max_mb < XW)
constexpr int SW1_PD = 2;     // panels prefetched ahead into L2

__device__ __forceinline__ void l2_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void sw1_prefetch(const float* panel, int mb, const double* D) {
  const char* p = reinterpret_cast<const char*>(panel + NB2);          // tile k = 0 is not used by the sweeps
  const int lines = mb * (NB2 * 4 / 128);
  for (int l = threadIdx.x; l < lines; l += 256) l2_prefetch(p + (size_t)l * 128);
  if (D) l2_prefetch(reinterpret_cast<const char*>(D) + (size_t)threadIdx.x * 128);   // 32 KB = 256 lines
}

__device__ __forceinline__ void sw_single_solve(const GfSchwarz& S, const int i, const double* __restrict__ r, double* ys) {
  __shared__ __align__(16) float xf1[NB];
  __shared__ __align__(16) float xw[XW * NB];
  __shared__ double red1[4][NB];
  const int tid = threadIdx.x;
  const int nbr = S.nbr[i], n_pad = S.n_pad[i];
  const int32_t* mbj = S.mbj + S.off_j[i];
  const int64_t* offc = S.off_col + S.off_j[i];
  const double* invd = S.invd + S.off_inv[i];
  const int32_t* glob = S.glob + S.off_y[i];
  const int fr = tid >> 2, fq = tid & 3;           // forward: row fr, quarter fq (16 columns)
  const int bc = tid & 63, bq = tid >> 6;          // transposed products: column bc, quarter bq (16 rows)
  for (int p = 0; p < SW1_PD && p < nbr; ++p) sw1_prefetch(S.band32 + offc[p], mbj[p], invd + (size_t)p * NB2);
  for (int l = tid; l < n_pad; l += 256) { const int g = glob[l]; ys[l] = (g >= 0) ? r[g] : 0.0; }
  __syncthreads();
  // ---------------- forward: y_{j+k} -= M(j+k, j) y_j ; then w_j = D_j y_j replaces y_j ----------------
  for (int j = 0; j < nbr; ++j) {
    const int mb = mbj[j];
    if (tid < NB) xf1[tid] = (float)ys[j * NB + tid];
    if (j + SW1_PD < nbr) sw1_prefetch(S.band32 + offc[j + SW1_PD], mbj[j + SW1_PD], invd + (size_t)(j + SW1_PD) * NB2);
    __syncthreads();
    float xr[16];
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(xf1 + fq * 16 + c);
      xr[c] = v.x; xr[c + 1] = v.y; xr[c + 2] = v.z; xr[c + 3] = v.w;
    }
    const float* panel = S.band32 + offc[j];
    for (int k = 1; k <= mb; k += 4) {
      float4 m[4][4];
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (k + t <= mb) {
          const float4* p = reinterpret_cast<const float4*>(panel + (size_t)(k + t) * NB2 + fr * NB + fq * 16);
          m[t][0] = __ldcs(p); m[t][1] = __ldcs(p + 1); m[t][2] = __ldcs(p + 2); m[t][3] = __ldcs(p + 3);
        }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (k + t <= mb) {
          float a = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            a = fmaf(m[t][c].x, xr[4 * c], a); a = fmaf(m[t][c].y, xr[4 * c + 1], a);
            a = fmaf(m[t][c].z, xr[4 * c + 2], a); a = fmaf(m[t][c].w, xr[4 * c + 3], a);
          }
          a += __shfl_xor_sync(0xffffffffu, a, 1);
          a += __shfl_xor_sync(0xffffffffu, a, 2);
          if (fq == 0) ys[(j + k + t) * NB + fr] -= (double)a;      // one owner per entry and step
        }
    }
    {   // D_j is symmetric: read it row-wise as its own transpose (coalesced over the column index)
      const double* D = invd + (size_t)j * NB2 + (size_t)(bq * 16) * NB + bc;
      const double* yj = ys + j * NB + bq * 16;
      double d[16];
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) d[rr] = __ldcs(D + rr * NB);
      double acc = 0.0;
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) acc = fma(d[rr], yj[rr], acc);
      red1[bq][bc] = acc;
    }
    __syncthreads();
    if (tid < NB) ys[j * NB + tid] = (red1[0][tid] + red1[1][tid]) + (red1[2][tid] + red1[3][tid]);
  }
  // ---------------- backward: x_j = w_j - sum_k M(j+k, j)^T x_{j+k} ----------------
  for (int j = nbr - 1; j >= 0; --j) {
    const int mb = mbj[j];
    if (j - SW1_PD >= 0) sw1_prefetch(S.band32 + offc[j - SW1_PD], mbj[j - SW1_PD], nullptr);
    const float* panel = S.band32 + offc[j];
    double acc = 0.0;
    for (int k = 1; k <= mb; k += 4) {
      float m[4][16];
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (k + t <= mb) {
          const float* p = panel + (size_t)(k + t) * NB2 + (bq * 16) * NB + bc;
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) m[t][rr] = __ldcs(p + rr * NB);
        }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (k + t <= mb) {
          const float4* xv = reinterpret_cast<const float4*>(xw + ((j + k + t) & (XW - 1)) * NB + bq * 16);
          float a = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 v = xv[c];
            a = fmaf(m[t][4 * c], v.x, a); a = fmaf(m[t][4 * c + 1], v.y, a);
            a = fmaf(m[t][4 * c + 2], v.z, a); a = fmaf(m[t][4 * c + 3], v.w, a);
          }
          acc += (double)a;
        }
    }
    red1[bq][bc] = acc;
    __syncthreads();
    if (tid < NB) {
      const double xv = ys[j * NB + tid] - ((red1[0][tid] + red1[1][tid]) + (red1[2][tid] + red1[3][tid]));
      ys[j * NB + tid] = xv;
      xw[(j & (XW - 1)) * NB + tid] = (float)xv;
    }
    __syncthreads();
  }
  double* y = S.y + S.off_y[i];
  for (int l = tid; l < n_pad; l += 256) y[l] = ys[l];
}

// CTAs [0, G_c) sweep the coarse block as one barrier group; every other CTA owns one fine block.
__global__ void __launch_bounds__(256, 2)
k_sw_solve1(GfSchwarz Sf, const double* __restrict__ r_f, int first_block, GfSchwarz Sc, int G_c) {
  extern __shared__ __align__(16) unsigned char sw_smem[];
  if ((int)blockIdx.x < G_c) sw_group_solve(Sc, 0, (int)blockIdx.x, G_c, reinterpret_cast<float*>(sw_smem));
  else sw_single_solve(Sf, first_block + (int)blockIdx.x - G_c, r_f, reinterpret_cast<double*>(sw_smem));
}

// ---- the coarse block on ONE thread-block cluster ----------------------------------------
// The coarse block is a single long dependency chain (2 x nbr steps); with CTAs that only share
// global memory every step costs two L2 round trips plus a polled barrier (~3 us measured).  On a
// cluster the step is: the owner of block row j broadcasts its 64 FP32 values into every CTA's
// shared memory (DSMEM), barrier.cluster, every CTA updates the rows IT owns out of its own
// shared memory.  Block row R lives in CTA R % CL ("owner computes"): no atomics, no global
// traffic on the chain, tiles arrive through a 3-step cp.async ring.
//   forward : y_R -= M(R, j) y_j          for the tiles (R, j) with R % CL == rank
//   backward: x_C -= M(j, C)^T x_j        for the tiles (j, C) with C % CL == rank
namespace cg = cooperative_groups;

__device__ __forceinline__ int cc_k0(int c, int j, int CL, bool backward) {
  int d = (backward ? (j - c) : (c - j)) % CL;
  if (d < 0) d += CL;
  return d == 0 ? CL : d;
}
__device__ __forceinline__ void cc_prefetch(const float* __restrict__ band32, const int64_t* __restrict__ offc,
                                            const int32_t* __restrict__ lim, int nbr, int j, int c, int CL,
                                            bool backward, float* stage) {
  if (j >= 0 && j < nbr) {
    const int kmax = lim[j];
    int t = 0;
    for (int k = cc_k0(c, j, CL, backward); k <= kmax; k += CL, ++t) {
      const float* src = band32 + (backward ? offc[j - k] : offc[j]) + (size_t)k * NB2;
      float* dst = stage + t * PBLK;
      for (int e = threadIdx.x; e < NB2 / 4; e += 256)
        __pipeline_memcpy_async(dst + (e >> 4) * LDS + (e & 15) * 4, src + e * 4, 16);
    }
  }
  __pipeline_commit();                       // one group per step, empty or not
}

__global__ void __launch_bounds__(256, 1)
k_sw_coarse_cluster(GfSchwarz S, int TPS) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), c = (int)cluster.block_rank();
  extern __shared__ __align__(16) unsigned char sw_smem[];
  float* ring = reinterpret_cast<float*>(sw_smem);                                       // [3][TPS][PBLK]
  double* yl = reinterpret_cast<double*>(sw_smem + (size_t)3 * TPS * PBLK * sizeof(float));  // [rows_loc][NB]
  __shared__ __align__(16) float xb[2][NB];
  __shared__ float redc[4][NB];
  __shared__ double redd[4][NB];
  const int nbr = S.nbr[0];
  const double* invd = S.invd + S.off_inv[0];
  double* y = S.y + S.off_y[0];
  const int tid = threadIdx.x;
  const int fr = tid >> 2, fq = tid & 3;
  const int bc = tid & 63, bq = tid >> 6;
  const int rows_loc = (nbr + CL - 1) / CL;
  const size_t stage_sz = (size_t)TPS * PBLK;
  // the envelope tables are read on the dependency chain: keep them in shared memory
  // (cluster.sync invalidates L1, every global read after it would be an L2 round trip)
  int64_t* offc = reinterpret_cast<int64_t*>(yl + (size_t)rows_loc * NB);                 // [nbr]
  int32_t* mbj = reinterpret_cast<int32_t*>(offc + nbr);                                  // [nbr]
  int32_t* rlen = mbj + nbr;                                                              // [nbr]
  for (int t = tid; t < nbr; t += 256) {
    offc[t] = S.off_col[S.off_j[0] + t]; mbj[t] = S.mbj[S.off_j[0] + t]; rlen[t] = S.rlen[S.off_j[0] + t];
  }
  __syncthreads();

  cc_prefetch(S.band32, offc, mbj, nbr, 0, c, CL, false, ring);
  cc_prefetch(S.band32, offc, mbj, nbr, 1, c, CL, false, ring + stage_sz);
  for (int idx = tid; idx < rows_loc * NB; idx += 256) {
    const int R = (idx / NB) * CL + c;
    yl[idx] = (R < nbr) ? y[(size_t)R * NB + (idx % NB)] : 0.0;
  }
  __syncthreads();
  if (c == 0 && tid < NB) {
    const float v = (float)yl[tid];
    for (int q = 0; q < CL; ++q) cluster.map_shared_rank(&xb[0][0], q)[tid] = v;
  }
  cluster.sync();
  // ---------------- forward ----------------
  for (int j = 0; j < nbr; ++j) {
    cc_prefetch(S.band32, offc, mbj, nbr, j + 2, c, CL, false, ring + ((j + 2) % 3) * stage_sz);
    __pipeline_wait_prior(2);
    __syncthreads();
    float xr[16];
#pragma unroll
    for (int cc = 0; cc < 16; cc += 4) {
      const float4 v = *reinterpret_cast<const float4*>(&xb[j & 1][fq * 16 + cc]);
      xr[cc] = v.x; xr[cc + 1] = v.y; xr[cc + 2] = v.z; xr[cc + 3] = v.w;
    }
    const float* stage = ring + (j % 3) * stage_sz;
    const int kmax = mbj[j];
    int t = 0;
    for (int k = cc_k0(c, j, CL, false); k <= kmax; k += CL, ++t) {
      const float* p = stage + t * PBLK + fr * LDS + fq * 16;
      float a = 0.f;
#pragma unroll
      for (int cc = 0; cc < 16; cc += 4) {
        const float4 m = *reinterpret_cast<const float4*>(p + cc);
        a = fmaf(m.x, xr[cc], a); a = fmaf(m.y, xr[cc + 1], a); a = fmaf(m.z, xr[cc + 2], a); a = fmaf(m.w, xr[cc + 3], a);
      }
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      if (fq == 0) yl[((j + k) / CL) * NB + fr] -= (double)a;
    }
    __syncthreads();
    if (j + 1 < nbr && c == (j + 1) % CL && tid < NB) {
      const float v = (float)yl[((j + 1) / CL) * NB + tid];
      for (int q = 0; q < CL; ++q) cluster.map_shared_rank(&xb[0][0], q)[((j + 1) & 1) * NB + tid] = v;
    }
    cluster.sync();
  }
  __pipeline_wait_prior(0);
  // ---------------- diagonal: w_R = D_R y_R for the rows this CTA owns (FP64) ----------------
  for (int R = c; R < nbr; R += CL) {
    const int lr = R / CL;
    const double* D = invd + (size_t)R * NB2 + (size_t)(bq * 16) * NB + bc;     // D symmetric: rows as columns
    const double* yr = yl + lr * NB + bq * 16;
    double d[16];
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) d[rr] = __ldcs(D + rr * NB);
    double acc = 0.0;
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) acc = fma(d[rr], yr[rr], acc);
    redd[bq][bc] = acc;
    __syncthreads();
    if (tid < NB) yl[lr * NB + tid] = (redd[0][tid] + redd[1][tid]) + (redd[2][tid] + redd[3][tid]);
    __syncthreads();
  }
  // ---------------- backward ----------------
  const int last = nbr - 1;
  cc_prefetch(S.band32, offc, rlen, nbr, last, c, CL, true, ring + (last % 3) * stage_sz);
  cc_prefetch(S.band32, offc, rlen, nbr, last - 1, c, CL, true, ring + ((last + 2) % 3) * stage_sz);
  if (c == last % CL && tid < NB) {
    const float v = (float)yl[(last / CL) * NB + tid];
    for (int q = 0; q < CL; ++q) cluster.map_shared_rank(&xb[0][0], q)[(last & 1) * NB + tid] = v;
  }
  cluster.sync();
  for (int j = last; j >= 0; --j) {
    cc_prefetch(S.band32, offc, rlen, nbr, j - 2, c, CL, true, ring + ((j + 1) % 3) * stage_sz);   // (j-2) % 3 == (j+1) % 3
    __pipeline_wait_prior(2);
    __syncthreads();
    float xr[16];
#pragma unroll
    for (int cc = 0; cc < 16; cc += 4) {
      const float4 v = *reinterpret_cast<const float4*>(&xb[j & 1][bq * 16 + cc]);
      xr[cc] = v.x; xr[cc + 1] = v.y; xr[cc + 2] = v.z; xr[cc + 3] = v.w;
    }
    const float* stage = ring + (j % 3) * stage_sz;
    const int kmax = rlen[j];
    int t = 0;
    for (int k = cc_k0(c, j, CL, true); k <= kmax; k += CL, ++t) {
      const float* p = stage + t * PBLK + (bq * 16) * LDS + bc;
      float a = 0.f;
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) a = fmaf(p[rr * LDS], xr[rr], a);
      redc[bq][bc] = a;
      __syncthreads();
      if (tid < NB) yl[((j - k) / CL) * NB + tid] -= (double)((redc[0][tid] + redc[1][tid]) + (redc[2][tid] + redc[3][tid]));
      __syncthreads();
    }
    if (j >= 1 && c == (j - 1) % CL && tid < NB) {
      const float v = (float)yl[((j - 1) / CL) * NB + tid];
      for (int q = 0; q < CL; ++q) cluster.map_shared_rank(&xb[0][0], q)[((j - 1) & 1) * NB + tid] = v;
    }
    cluster.sync();
  }
  __pipeline_wait_prior(0);
  for (int idx = tid; idx < rows_loc * NB; idx += 256) {
    const int R = (idx / NB) * CL + c;
    if (R < nbr) y[(size_t)R * NB + (idx % NB)] = yl[idx];
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
    const int m = S->step_mb_h[j];      // tallest panel of this step over all patch blocks
    k_sw_potrf<<<S->nblocks, 256, smem, st>>>(*S, j);
    if (m > 0) {
      k_sw_trsm<<<dim3(m, S->nblocks), 256, smem, st>>>(*S, j);
      k_sw_update<<<dim3(m * (m + 1) / 2, S->nblocks), 256, smem, st>>>(*S, j);
    }
    count_launch(3);
  }
  e = cudaFuncSetAttribute(k_sw_convert_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sw_convert_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(k_sw_convert)");
  if (S->max_mb > 0) k_sw_convert_panel<<<dim3(S->max_mb, S->max_nbr, S->nblocks), 256, smem, st>>>(*S);
  k_sw_convert_diag<<<dim3(S->max_nbr, S->nblocks), 256, smem, st>>>(*S);
  k_sw_to_f32<<<2048, 256, 0, st>>>(S->band, S->band32, S->band_len);
  count_launch(3);
  int flag = 0;
  e = cudaMemcpyAsync(&flag, S->flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_cuda_error(e, "gf_schwarz_factor");
  if (flag) return set_error(GF_ERR_BREAKDOWN, "gf_schwarz_factor: a block is not positive definite");
  return check_launch("gf_schwarz_factor");
}

static int sw_caps(int* sms_out, int* occ_out) {
  static int sms = 0, occ = 0;
  if (!sms) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(k_sw_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PF * PBLK * sizeof(float)));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sw_solve, 256, PF * PBLK * sizeof(float));
    int lim = 4;
    if (const char* ev = getenv("GF_SW_OCC")) lim = atoi(ev);
    if (occ > lim) occ = lim;
    if (occ < 1) occ = 1;
  }
  *sms_out = sms; *occ_out = occ;
  return sms * occ;
}

// Single-CTA-per-block path: capacity and dynamic shared memory of k_sw_solve1 for blocks of
// at most max_n_pad dofs riding with (or without) a coarse group.  Returns 0 if it does not fit.
static int sw1_caps(int max_n_pad, bool with_coarse, size_t* smem_out) {
  size_t smem = (size_t)max_n_pad * sizeof(double);
  if (with_coarse && smem < PF * PBLK * sizeof(float)) smem = PF * PBLK * sizeof(float);
  static size_t smem_set = 0, smem_cached = 0;
  static int cap_cached = 0;
  *smem_out = smem;
  if (smem == smem_cached) return cap_cached;
  int dev = 0, sms = 0, lim = 0, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaFuncAttributes fa;
  int cap = 0;
  if (cudaFuncGetAttributes(&fa, k_sw_solve1) == cudaSuccess && smem + fa.sharedSizeBytes <= (size_t)lim) {
    bool ok = true;
    if (smem > smem_set) {
      ok = cudaFuncSetAttribute(k_sw_solve1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
      if (ok) smem_set = smem;
    }
    if (ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sw_solve1, 256, smem) == cudaSuccess && occ >= 1) cap = sms * occ;
  }
  cudaGetLastError();                              // a failed probe must not poison the next launch check
  smem_cached = smem; cap_cached = cap;
  return cap;
}

// Coarse block on one thread-block cluster (16 CTAs, else 8).  Returns 1 when launched, 0 when the
// block does not fit (caller falls back to the barrier-group kernel), <0 on a launch error.
static int launch_coarse_cluster(const GfSchwarz* Sc, cudaStream_t st) {
  static int CL = -1;                            // cluster size found usable on this device (0: none)
  static size_t smem_set = 0;
  const int tries[2] = {16, 8};
  for (int a = 0; a < 2; ++a) {
    const int cl = (CL > 0) ? CL : tries[a];
    if (CL == 0) return 0;
    const int TPS = (Sc->max_mb + cl - 1) / cl > 0 ? (Sc->max_mb + cl - 1) / cl : 1;
    const int rows_loc = (Sc->max_nbr + cl - 1) / cl;
    const size_t smem = (size_t)3 * TPS * PBLK * sizeof(float) + (size_t)rows_loc * NB * sizeof(double) + (size_t)Sc->max_nbr * 16;
    bool ok = smem <= 200 * 1024;
    if (ok && cl > 8 && CL < 0) ok = cudaFuncSetAttribute(k_sw_coarse_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    if (ok && smem > smem_set) {
      ok = cudaFuncSetAttribute(k_sw_coarse_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
      if (ok) smem_set = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cl); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (ok && CL < 0) {
      int ncl = 0;
      ok = cudaOccupancyMaxActiveClusters(&ncl, k_sw_coarse_cluster, &cfg) == cudaSuccess && ncl >= 1;
    }
    if (ok) {
      cudaError_t e = cudaLaunchKernelEx(&cfg, k_sw_coarse_cluster, *Sc, TPS);
      if (e != cudaSuccess) { set_cuda_error(e, "launch k_sw_coarse_cluster"); return -1; }
      CL = cl;
      count_launch(1);
      return 1;
    }
    cudaGetLastError();
    if (CL > 0) return 0;                        // this coarse block is too large for the known cluster size
  }
  CL = 0;
  return 0;
}

// z_f = sum_i R_i^T A_i^-1 R_i r_f  (fine blocks of Sf)  and, if Sc != NULL, z_c = Kc^-1 r_c
// (single block of Sc) with both sets of triangular sweeps running concurrently.
extern "C" int gf_schwarz_apply2(const GfSchwarz* Sf, const double* r_f, double* z_f, int64_t n_f,
                                 const GfSchwarz* Sc, const double* r_c, double* z_c, int64_t n_c, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  int sms, occ;
  const int cap = sw_caps(&sms, &occ);           // all CTAs of one cooperative launch must be co-resident
  GfSchwarz Sfv = *Sf, Scv = Sc ? *Sc : *Sf;
  static int dbg = -1, single = -1;
  if (dbg < 0) { dbg = 0; if (const char* ev = getenv("GF_SW_DEBUG")) dbg = atoi(ev); }
  if (single < 0) { single = 2; if (const char* ev = getenv("GF_SW_SINGLE")) single = atoi(ev); }   // 0 off, 1 on, 2 auto
  Sfv.debug_flags = Sf->debug_flags | dbg; Scv.debug_flags = dbg;
  int mode = single;                              // GfSchwarz.debug_flags: 4 forces one CTA per block, 8 forces CTA groups
  if (Sfv.debug_flags & 4) mode = 1;
  if (Sfv.debug_flags & 8) mode = 0;
  int G_c = 0;
  bool coarse_side = false;                       // coarse sweeps on their own stream, concurrent with the fine ones
  static cudaStream_t side = nullptr;
  static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t sf = st;                           // stream of the fine sweeps
  if (Sc) {
    int gc = (int)((Sc->n_y + 255) / 256); if (gc > 2048) gc = 2048;
    k_sw_gather_in<<<gc, 256, 0, st>>>(*Sc, r_c);
    count_launch(1);
    if (mode && !(Sfv.debug_flags & 16)) {
      if (!side) {
        if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess)
          return set_error(GF_ERR_CUDA, "gf_schwarz_apply: cannot create the second sweep stream");
      }
      // The cluster needs 16 SMs of ONE GPC with room for it, which the fine CTAs would not leave:
      // it goes first, on the caller's stream; the fine sweeps follow on a second stream (behind an
      // event recorded before the cluster launch) and fill the rest of the machine around it.
      e = cudaEventRecord(ev_fork, st);
      if (e != cudaSuccess) return set_cuda_error(e, "gf_schwarz_apply fork");
      const int lc = launch_coarse_cluster(Sc, st);
      if (lc < 0) return GF_ERR_CUDA;
      if (lc == 1) {
        int g3 = (int)((n_c + 255) / 256); if (g3 > 2048) g3 = 2048;
        k_sw_gather_out<<<g3, 256, 0, st>>>(*Sc, z_c, n_c);
        count_launch(1);
        e = cudaStreamWaitEvent(side, ev_fork, 0);
        if (e != cudaSuccess) return set_cuda_error(e, "gf_schwarz_apply fork");
        coarse_side = true;
        sf = side;
      }
    }
    if (!coarse_side) {
      e = cudaMemsetAsync(Sc->barrier, 0, sizeof(unsigned) * Sc->nblocks, st);
      if (e != cudaSuccess) return set_cuda_error(e, "gf_schwarz_apply memset");
      G_c = cap / 6; if (G_c > Sc->max_mb) G_c = Sc->max_mb; if (G_c < 1) G_c = 1;
    }
  }
  // one CTA per block whenever a block's vector fits in shared memory (otherwise: G CTAs per block)
  size_t smem1 = 0;
  int cap1 = 0;
  if (mode && Sf->max_mb < XW) cap1 = sw1_caps(Sf->max_n_pad, Sc != nullptr && !coarse_side, &smem1);
  if (cap1 > G_c) {
    for (int b0 = (Sfv.debug_flags & 32) ? Sf->nblocks : 0; b0 < Sf->nblocks;) {   // 32: timing experiments, coarse only
      const int gcl = (b0 == 0) ? G_c : 0;       // the coarse group rides along with the first batch
      int nb_l = Sf->nblocks - b0;
      if (nb_l > cap1 - gcl) nb_l = cap1 - gcl;
      int first = b0;
      void* args[] = {&Sfv, &r_f, &first, &Scv, (void*)&gcl};
      if (gcl) e = cudaLaunchCooperativeKernel((void*)k_sw_solve1, dim3(nb_l + gcl), dim3(256), args, smem1, sf);
      else { k_sw_solve1<<<nb_l, 256, smem1, sf>>>(Sfv, r_f, first, Scv, 0); e = cudaGetLastError(); }
      if (e != cudaSuccess) return set_cuda_error(e, "launch k_sw_solve1");
      count_launch(1);
      b0 += nb_l;
    }
  } else {
    int g = (int)((Sf->n_y + 255) / 256); if (g > 2048) g = 2048;
    k_sw_gather_in<<<g, 256, 0, sf>>>(*Sf, r_f);
    e = cudaMemsetAsync(Sf->barrier, 0, sizeof(unsigned) * Sf->nblocks, sf);
    if (e != cudaSuccess) return set_cuda_error(e, "gf_schwarz_apply memset");
    count_launch(1);
    int G = (cap - G_c) / Sf->nblocks;
    if (G > Sf->max_mb) G = Sf->max_mb;            // one panel block per CTA and step is enough
    if (G < 1) G = 1;
    const int per_launch = (cap - G_c) / G;        // fine blocks per cooperative launch
    for (int b0 = 0; b0 < Sf->nblocks; b0 += per_launch) {
      int nb_l = Sf->nblocks - b0; if (nb_l > per_launch) nb_l = per_launch;
      int first = b0;
      int gcl = (b0 == 0) ? G_c : 0;               // the coarse block rides along with the first batch
      void* args[] = {&Sfv, &G, &first, &nb_l, &Scv, &gcl};
      e = cudaLaunchCooperativeKernel((void*)k_sw_solve, dim3(nb_l * G + gcl), dim3(256), args, PF * PBLK * sizeof(float), sf);
      if (e != cudaSuccess) return set_cuda_error(e, "cudaLaunchCooperativeKernel(k_sw_solve)");
      count_launch(1);
    }
  }
  int g2 = (int)((n_f + 255) / 256); if (g2 > 2048) g2 = 2048;
  k_sw_gather_out<<<g2, 256, 0, sf>>>(*Sf, z_f, n_f);
  if (Sc && !coarse_side) {
    int g3 = (int)((n_c + 255) / 256); if (g3 > 2048) g3 = 2048;
    k_sw_gather_out<<<g3, 256, 0, st>>>(*Sc, z_c, n_c);
    count_launch(1);
  }
  if (coarse_side) {
    e = cudaEventRecord(ev_join, side);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev_join, 0);
    if (e != cudaSuccess) return set_cuda_error(e, "gf_schwarz_apply join");
  }
  count_launch(1);
  return check_launch("gf_schwarz_apply");
}

// Fine sweeps alone (one CTA per block), result left in S->y: the launch bench.py times for the roofline.
extern "C" int gf_schwarz_sweeps(const GfSchwarz* S, const double* r, void* stream) {
  if (!S || !r) return set_error(GF_ERR_BADARG, "gf_schwarz_sweeps: null argument");
  size_t smem1 = 0;
  const int cap1 = (S->max_mb < XW) ? sw1_caps(S->max_n_pad, false, &smem1) : 0;
  if (cap1 < 1) return set_error(GF_ERR_BADARG, "gf_schwarz_sweeps: blocks do not fit the one-CTA-per-block kernel");
  for (int b0 = 0; b0 < S->nblocks; b0 += cap1) {
    const int nb_l = (S->nblocks - b0 < cap1) ? S->nblocks - b0 : cap1;
    k_sw_solve1<<<nb_l, 256, smem1, (cudaStream_t)stream>>>(*S, r, b0, *S, 0);
    count_launch(1);
  }
  return check_launch("k_sw_solve1");
}

extern "C" int gf_schwarz_apply(const GfSchwarz* S, const double* r, double* z, int64_t n, void* stream) {
  return gf_schwarz_apply2(S, r, z, n, nullptr, nullptr, nullptr, 0, stream);
}

extern "C" int gf_dot_slot0(int64_t n, const double* x, const double* y, double* partial2, int grid, void* stream) {
  k_dot_slot0<<<grid, 256, 0, (cudaStream_t)stream>>>(n, x, y, partial2);
  return check_launch("k_dot_slot0");
}
