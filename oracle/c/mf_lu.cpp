// ORACLE / CPU BASELINE (test + measurement infrastructure, never the product path).
//
// Multifrontal sparse LU on a nested-dissection tree: the CPU restatement of the reference's linear solves,
// `solve_nonmatching_mat(A, x, b, solver='direct')` = PETSc PCLU with MUMPS
// (/root/reference/GOLDFISH/utils/opt_utils.py:176 state, :199-204 adjoint: explicit transpose + fresh LU).
// MUMPS is not in this image; this is its published algorithm (Duff & Reid multifrontal method: per tree node a
// dense frontal matrix, partial LU of the fully summed block with pivoting restricted to that block, Schur
// complement extend-added into the parent), with the dense kernels taken from the BLAS/LAPACK the image has
// (scipy's OpenBLAS, passed in as function pointers) and OpenMP over independent subtrees -- i.e. a competent
// multi-core CPU code, so that bench.py's CPU arm uses the host cores the way the reference's MPI run would.
//
// The ordering / symbolic phase is oracle/nested_dissection.py.  All indices here are PERMUTED dof positions.
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include <omp.h>
#include <sys/mman.h>

namespace {
typedef void (*dgemm_t)(char*, char*, int*, int*, int*, double*, double*, int*, double*, int*, double*, double*, int*);
typedef void (*dtrsm_t)(char*, char*, char*, char*, int*, int*, double*, double*, int*, double*, int*);
typedef void (*dgetrf_t)(int*, int*, double*, int*, int*, int*);
typedef void (*dlaswp_t)(int*, double*, int*, int*, int*, int*, int*);
typedef void (*dgemv_t)(char*, int*, int*, double*, double*, int*, double*, int*, double*, double*, int*);
typedef void (*dtrsv_t)(char*, char*, char*, int*, double*, int*, double*, int*);
typedef void (*setthr_t)(int);

struct Blas { dgemm_t gemm; dtrsm_t trsm; dgetrf_t getrf; dlaswp_t laswp; dgemv_t gemv; dtrsv_t trsv; setthr_t set_threads; };

struct Front {
  int k = 0, u = 0;            // fully summed dofs, border dofs
  int64_t first = 0;           // first own permuted position
  const int64_t* border = nullptr;   // [u] permuted positions, sorted
  int parent = -1, level = 0;
  std::vector<int> child;
  std::vector<int> rel;        // [u] index of each border dof in the parent's [own | border] numbering
  double* P = nullptr;         // (k+u) x k panel: [LU11; L21], column major, ld = k+u
  double* U12 = nullptr;       // k x u, ld = k
  double* S = nullptr;         // u x u Schur complement (freed once the parent has taken it)
  int* ipiv = nullptr;
  double* z = nullptr;         // [u] forward-solve update vector
};

struct Solver {
  int nf = 0; int64_t N = 0;
  std::vector<Front> F;
  std::vector<int64_t> border_store;
  std::vector<std::vector<int>> levels;
  Blas blas{};
  int big = 1200;              // fronts of at least this order use the threaded BLAS, one at a time
  int nthreads = 1;
  double flops = 0.0; int64_t lu_doubles = 0;
  int info = 0;
  double t_small = 0, t_large = 0, t_asm_large = 0, t_getrf_large = 0, t_trsm_large = 0, t_gemm_large = 0;
};

// large blocks: 2 MB aligned + MADV_HUGEPAGE, so first touch costs one page fault per 2 MB instead of per 4 KB
inline double* big_alloc(size_t n) {
  const size_t bytes = n * sizeof(double);
  if (bytes < ((size_t)8 << 20)) return (double*)malloc(bytes);
  const size_t al = (size_t)2 << 20;
  void* p = aligned_alloc(al, (bytes + al - 1) / al * al);
  if (p) madvise(p, (bytes + al - 1) / al * al, MADV_HUGEPAGE);
  return (double*)p;
}

inline int local_index(const Front& f, int64_t p) {     // position p -> index in [own | border]
  if (p >= f.first && p < f.first + f.k) return (int)(p - f.first);
  const int64_t* b = std::lower_bound(f.border, f.border + f.u, p);
  return f.k + (int)(b - f.border);
}

void free_numeric(Solver& s) {
  for (auto& f : s.F) { free(f.P); free(f.U12); free(f.S); free(f.ipiv); free(f.z); f.P = f.U12 = f.S = f.z = nullptr; f.ipiv = nullptr; }
}

// assemble + partial factorisation of one front
void do_front(Solver& s, int t, const int64_t* indptr, const int32_t* indices, const double* vals,
              const int64_t* indptrT, const int32_t* indicesT, const double* valsT,
              const int64_t* perm, const int64_t* pos) {
  Front& f = s.F[t];
  const int k = f.k, u = f.u, m = k + u;
  const bool lg = m >= s.big || s.nthreads == 1;
  double t0 = omp_get_wtime();
  const bool par = m >= s.big;           // large fronts are processed one at a time: use the threads inside
  const size_t nP = (size_t)m * std::max(k, 1), nU = (size_t)std::max(k, 1) * std::max(u, 1), nS = (size_t)std::max(u, 1) * std::max(u, 1);
  f.P = big_alloc(nP);
  f.U12 = big_alloc(nU);
  f.S = big_alloc(nS);
  f.ipiv = (int*)calloc(std::max(k, 1), sizeof(int));
  {
    const size_t CH = 1 << 16;           // first touch in parallel: page faults are the cost of fresh memory
    const int64_t cP = (int64_t)((nP + CH - 1) / CH), cU = (int64_t)((nU + CH - 1) / CH), cS = (int64_t)((nS + CH - 1) / CH);
#pragma omp parallel for schedule(static) if (par)
    for (int64_t c = 0; c < cP + cU + cS; ++c) {
      double* base; size_t n, c0;
      if (c < cP) { base = f.P; n = nP; c0 = (size_t)c; }
      else if (c < cP + cU) { base = f.U12; n = nU; c0 = (size_t)(c - cP); }
      else { base = f.S; n = nS; c0 = (size_t)(c - cP - cU); }
      const size_t b0 = c0 * CH, b1 = std::min(n, b0 + CH);
      memset(base + b0, 0, (b1 - b0) * sizeof(double));
    }
  }
  // original entries: own rows (columns >= first) and own columns (rows beyond the own range)
#pragma omp parallel for schedule(static) if (par)
  for (int i = 0; i < k; ++i) {
    const int64_t r = perm[f.first + i];
    for (int64_t q = indptr[r]; q < indptr[r + 1]; ++q) {
      const int64_t pc = pos[indices[q]];
      if (pc < f.first) continue;
      const int j = local_index(f, pc);
      if (j < k) f.P[(size_t)j * m + i] += vals[q]; else f.U12[(size_t)(j - k) * k + i] += vals[q];
    }
    for (int64_t q = indptrT[r]; q < indptrT[r + 1]; ++q) {      // column r of A = row r of A^T
      const int64_t pr = pos[indicesT[q]];
      if (pr < f.first + k) continue;                            // own block already taken from the rows
      const int j = local_index(f, pr);
      f.P[(size_t)i * m + j] += valsT[q];
    }
  }
  // extend-add the children's Schur complements (one destination column per source column: no write conflicts)
  for (int c : f.child) {
    Front& g = s.F[c];
    const int uc = g.u;
#pragma omp parallel for schedule(static) if (par)
    for (int b = 0; b < uc; ++b) {
      const int jb = g.rel[b];
      const double* col = g.S + (size_t)b * uc;
      if (jb < k) {
        double* dst = f.P + (size_t)jb * m;
        for (int a = 0; a < uc; ++a) dst[g.rel[a]] += col[a];
      } else {
        double* dstU = f.U12 + (size_t)(jb - k) * k;
        double* dstS = f.S + (size_t)(jb - k) * u;
        for (int a = 0; a < uc; ++a) {
          const int ia = g.rel[a];
          if (ia < k) dstU[ia] += col[a]; else dstS[ia - k] += col[a];
        }
      }
    }
    free(g.S); g.S = nullptr;
  }
  if (k == 0) return;
  if (lg) { const double t1 = omp_get_wtime(); s.t_asm_large += t1 - t0; t0 = t1; }
  // F11 = P L11 U11 (pivoting inside the fully summed block only)
  int K = k, M = m, U = u, info = 0, one = 1;
  double done = 1.0, dmone = -1.0;
  char L = 'L', Uc = 'U', Nn = 'N', Rr = 'R', Un = 'U';
  if (!par) {
    s.blas.getrf(&K, &K, f.P, &M, f.ipiv, &info);
    if (u > 0) {
      s.blas.laswp(&U, f.U12, &K, &one, &K, f.ipiv, &one);                     // U12 <- P^T U12
      s.blas.trsm(&L, &L, &Nn, &Un, &K, &U, &done, f.P, &M, f.U12, &K);        // U12 <- L11^-1 U12
      s.blas.trsm(&Rr, &Uc, &Nn, &Nn, &U, &K, &done, f.P, &M, f.P + k, &M);    // L21 <- F21 U11^-1
      s.blas.gemm(&Nn, &Nn, &U, &U, &K, &dmone, f.P + k, &M, f.U12, &K, &done, f.S, &U);   // S -= L21 U12
    }
  } else {
    // Large front, processed alone: the BLAS stays single-threaded and OpenMP splits every level-3 operation
    // into independent column / row chunks (one runtime owns the cores; no spinning worker pools fighting).
    const int NBK = 128, CHK = 256;
    for (int j0 = 0; j0 < k && info == 0; j0 += NBK) {       // right-looking blocked LU of the k x k block
      int jb = std::min(NBK, k - j0), rows = k - j0, inf2 = 0;
      double* Ajj = f.P + (size_t)j0 * m + j0;
      s.blas.getrf(&rows, &jb, Ajj, &M, f.ipiv + j0, &inf2);
      if (inf2 != 0) info = j0 + inf2;
      for (int i = 0; i < jb; ++i) f.ipiv[j0 + i] += j0;       // global row numbers (1-based)
      int k1 = j0 + 1, k2 = j0 + jb;
      const int ncl = (j0 + CHK - 1) / CHK, nrest = k - j0 - jb, ncr = (nrest + CHK - 1) / CHK;
#pragma omp parallel for schedule(dynamic, 1)
      for (int c = 0; c < ncl + ncr; ++c) {
        if (c < ncl) {                                         // row swaps in the columns to the left
          int c0 = c * CHK, nc = std::min(CHK, j0 - c0);
          s.blas.laswp(&nc, f.P + (size_t)c0 * m, &M, &k1, &k2, f.ipiv, &one);
        } else {                                               // swaps + block row of U + trailing update
          int c0 = j0 + jb + (c - ncl) * CHK, nc = std::min(CHK, k - c0), below = k - j0 - jb;
          double* B = f.P + (size_t)c0 * m;
          s.blas.laswp(&nc, B, &M, &k1, &k2, f.ipiv, &one);
          s.blas.trsm(&L, &L, &Nn, &Un, &jb, &nc, &done, Ajj, &M, B + j0, &M);
          if (below > 0) s.blas.gemm(&Nn, &Nn, &below, &nc, &jb, &dmone, Ajj + jb, &M, B + j0, &M, &done, B + j0 + jb, &M);
        }
      }
    }
    if (lg) { const double t1 = omp_get_wtime(); s.t_getrf_large += t1 - t0; t0 = t1; }
    if (u > 0) {
      const int ncu = (u + CHK - 1) / CHK;
#pragma omp parallel for schedule(dynamic, 1)
      for (int c = 0; c < 2 * ncu; ++c) {
        if (c < ncu) {                                         // U12 columns: P^T, then L11^-1
          int c0 = c * CHK, nc = std::min(CHK, u - c0);
          double* B = f.U12 + (size_t)c0 * k;
          s.blas.laswp(&nc, B, &K, &one, &K, f.ipiv, &one);
          s.blas.trsm(&L, &L, &Nn, &Un, &K, &nc, &done, f.P, &M, B, &K);
        } else {                                               // L21 rows: F21 U11^-1
          int r0 = (c - ncu) * CHK, nr = std::min(CHK, u - r0);
          s.blas.trsm(&Rr, &Uc, &Nn, &Nn, &nr, &K, &done, f.P, &M, f.P + k + r0, &M);
        }
      }
      if (lg) { const double t1 = omp_get_wtime(); s.t_trsm_large += t1 - t0; t0 = t1; }
#pragma omp parallel for schedule(dynamic, 1)
      for (int c = 0; c < ncu; ++c) {                          // S -= L21 U12, by column chunks
        int c0 = c * CHK, nc = std::min(CHK, u - c0);
        s.blas.gemm(&Nn, &Nn, &U, &nc, &K, &dmone, f.P + k, &M, f.U12 + (size_t)c0 * k, &K, &done, f.S + (size_t)c0 * u, &U);
      }
      if (lg) { const double t1 = omp_get_wtime(); s.t_gemm_large += t1 - t0; t0 = t1; }
    }
  }
  if (info != 0) {
#pragma omp critical
    s.info = info;
  }
}
}  // namespace

extern "C" {

void* mf_create(int nf, const int32_t* k, const int32_t* u, const int64_t* first, const int64_t* border_ptr,
                const int64_t* border_idx, const int32_t* parent, int64_t N) {
  Solver* s = new Solver();
  s->nf = nf; s->N = N;
  s->F.resize(nf);
  s->border_store.assign(border_idx, border_idx + border_ptr[nf]);
  for (int t = 0; t < nf; ++t) {
    Front& f = s->F[t];
    f.k = k[t]; f.u = u[t]; f.first = first[t]; f.parent = parent[t];
    f.border = s->border_store.data() + border_ptr[t];
  }
  for (int t = 0; t < nf; ++t) if (parent[t] >= 0) s->F[parent[t]].child.push_back(t);
  int maxlev = 0;
  for (int t = 0; t < nf; ++t) {                 // post-order: children come first
    int lv = 0;
    for (int c : s->F[t].child) lv = std::max(lv, s->F[c].level + 1);
    s->F[t].level = lv; maxlev = std::max(maxlev, lv);
  }
  s->levels.resize(maxlev + 1);
  for (int t = 0; t < nf; ++t) s->levels[s->F[t].level].push_back(t);
  for (int t = 0; t < nf; ++t) {
    Front& f = s->F[t];
    if (f.parent < 0) continue;
    const Front& p = s->F[f.parent];
    f.rel.resize(f.u);
    for (int b = 0; b < f.u; ++b) f.rel[b] = local_index(p, f.border[b]);
    const double kk = f.k, uu = f.u;
    (void)kk; (void)uu;
  }
  for (int t = 0; t < nf; ++t) {
    const double kk = s->F[t].k, uu = s->F[t].u;
    s->flops += 2.0 / 3.0 * kk * kk * kk + 2.0 * kk * kk * uu + 2.0 * kk * uu * uu;
    s->lu_doubles += (int64_t)((kk + uu) * kk + kk * uu);
  }
  s->nthreads = omp_get_max_threads();
  return s;
}

void mf_set_blas(void* h, void* gemm, void* trsm, void* getrf, void* laswp, void* gemv, void* trsv, void* set_threads, int big) {
  Solver* s = (Solver*)h;
  s->blas.gemm = (dgemm_t)gemm; s->blas.trsm = (dtrsm_t)trsm; s->blas.getrf = (dgetrf_t)getrf; s->blas.laswp = (dlaswp_t)laswp;
  s->blas.gemv = (dgemv_t)gemv; s->blas.trsv = (dtrsv_t)trsv; s->blas.set_threads = (setthr_t)set_threads;
  if (big > 0) s->big = big;
}

double mf_flops(void* h) { return ((Solver*)h)->flops; }
int64_t mf_lu_doubles(void* h) { return ((Solver*)h)->lu_doubles; }

// Numeric factorisation of A (CSR, original numbering) with A^T given as a second CSR (pass the same arrays for a
// symmetric matrix).  perm[p] = original dof at permuted position p, pos = its inverse.
int mf_factor(void* h, const int64_t* indptr, const int32_t* indices, const double* vals,
              const int64_t* indptrT, const int32_t* indicesT, const double* valsT,
              const int64_t* perm, const int64_t* pos) {
  Solver& s = *(Solver*)h;
  free_numeric(s);
  s.info = 0;
  s.t_small = s.t_large = s.t_asm_large = s.t_getrf_large = s.t_trsm_large = s.t_gemm_large = 0;
  if (s.blas.set_threads) s.blas.set_threads(1);     // single-threaded BLAS throughout; OpenMP owns the cores
  for (auto& lv : s.levels) {
    std::vector<int> small, large;
    for (int t : lv) (s.F[t].k + s.F[t].u >= s.big ? large : small).push_back(t);
    double tl0 = omp_get_wtime();
    if (!small.empty()) {
      const int ns = (int)small.size();
#pragma omp parallel for schedule(dynamic, 1)
      for (int i = 0; i < ns; ++i) do_front(s, small[i], indptr, indices, vals, indptrT, indicesT, valsT, perm, pos);
    }
    { const double t1 = omp_get_wtime(); s.t_small += t1 - tl0; tl0 = t1; }
    if (!large.empty()) {
      for (int t : large) do_front(s, t, indptr, indices, vals, indptrT, indicesT, valsT, perm, pos);
    }
    { const double t1 = omp_get_wtime(); s.t_large += t1 - tl0;
      if (getenv("GFO_MF_VERBOSE") && atoi(getenv("GFO_MF_VERBOSE")) > 1) {
        double fl = 0; int mx = 0;
        for (int t : lv) { const double kk = s.F[t].k, uu = s.F[t].u; fl += 2.0 / 3.0 * kk * kk * kk + 2.0 * kk * kk * uu + 2.0 * kk * uu * uu; mx = std::max(mx, s.F[t].k + s.F[t].u); }
        fprintf(stderr, "  level: %zu small %zu large, max order %d, %.2e flops, cumulative small %.2fs large %.2fs\n", small.size(), large.size(), mx, fl, s.t_small, s.t_large);
      } }
  }
  if (s.blas.set_threads) s.blas.set_threads(s.nthreads);
  if (getenv("GFO_MF_VERBOSE"))
    fprintf(stderr, "mf_factor: small fronts %.2fs  large fronts %.2fs (assemble %.2f getrf %.2f trsm %.2f gemm %.2f)\n",
            s.t_small, s.t_large, s.t_asm_large, s.t_getrf_large, s.t_trsm_large, s.t_gemm_large);
  return s.info;
}

// x = A^-1 b (original numbering in and out)
int mf_solve(void* h, const double* b, double* x, const int64_t* perm) {
  Solver& s = *(Solver*)h;
  std::vector<double> y(s.N);
  for (int64_t p = 0; p < s.N; ++p) y[p] = b[perm[p]];
  char L = 'L', Uc = 'U', Nn = 'N', Un = 'U';
  int one = 1;
  double done = 1.0, dmone = -1.0;
  if (s.blas.set_threads) s.blas.set_threads(1);
  // forward: L y = P b, front by front up the tree
  for (auto& lv : s.levels) {
    const int n = (int)lv.size();
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n; ++i) {
      Front& f = s.F[lv[i]];
      int k = f.k, u = f.u, m = k + u;
      f.z = (double*)calloc(std::max(u, 1), sizeof(double));
      double* r = y.data() + f.first;
      for (int c : f.child) {
        Front& g = s.F[c];
        for (int a = 0; a < g.u; ++a) { const int ia = g.rel[a]; if (ia < k) r[ia] += g.z[a]; else f.z[ia - k] += g.z[a]; }
        free(g.z); g.z = nullptr;
      }
      if (k == 0) continue;
      s.blas.laswp(&one, r, &k, &one, &k, f.ipiv, &one);
      s.blas.trsv(&L, &Nn, &Un, &k, f.P, &m, r, &one);
      if (u > 0) s.blas.gemv(&Nn, &u, &k, &dmone, f.P + k, &m, r, &one, &done, f.z, &one);
    }
  }
  for (auto& f : s.F) { free(f.z); f.z = nullptr; }
  // backward: U x = y, down the tree
  for (int li = (int)s.levels.size() - 1; li >= 0; --li) {
    auto& lv = s.levels[li];
    const int n = (int)lv.size();
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n; ++i) {
      Front& f = s.F[lv[i]];
      int k = f.k, u = f.u, m = k + u;
      if (k == 0) continue;
      double* r = y.data() + f.first;
      if (u > 0) {
        std::vector<double> xb(u);
        for (int a = 0; a < u; ++a) xb[a] = y[f.border[a]];
        s.blas.gemv(&Nn, &k, &u, &dmone, f.U12, &k, xb.data(), &one, &done, r, &one);
      }
      s.blas.trsv(&Uc, &Nn, &Nn, &k, f.P, &m, r, &one);
    }
  }
  if (s.blas.set_threads) s.blas.set_threads(s.nthreads);
  for (int64_t p = 0; p < s.N; ++p) x[perm[p]] = y[p];
  return 0;
}

void mf_set_threads(void* h, int n) { if (n > 0) { omp_set_num_threads(n); ((Solver*)h)->nthreads = n; } }
void mf_free_numeric(void* h) { free_numeric(*(Solver*)h); }
void mf_destroy(void* h) { Solver* s = (Solver*)h; free_numeric(*s); delete s; }
}
