// Shell quadrature + scatter kernels (sm_100a, FP64).
//
// One warp per Bezier element.  Elements are processed colour by colour
// ((p+1)^2 = 16 colours for bicubics): two elements of one colour share no
// control point, so every CSR slot / vector entry is written by exactly one
// warp per launch with plain read-modify-write -> the scatter needs no atomics
// and the summation order (colour order) is fixed: bit-reproducible.
//
// Per quadrature point the warp
//   A. builds the 16 rational basis functions (6 derivative kinds) from the
//      1-D span tables, the covariant vectors g_X, g_u and the thickness,
//   B. evaluates the hand-derived first variation with a dual number whose
//      direction differs per lane (lanes 0..14: d/dg_u, lane 15: d/dt,
//      lanes 16..30: d/dg_X at fixed u_hom) -> each lane owns one column of
//      the point Hessian (replaces UFL `derivative`, nonmatching_opt.py:440,449),
//   C. contracts it with the basis:  K_e += w Phi^T H Phi  in two stages
//      (G = H Phi in shared memory, then a register-tiled Phi^T G),
// and finally scatters the element tensors straight into the IGA-space CSR
// pattern (the M^T K M extraction of nonmatching_opt.py:688 is fused away:
// we never leave the spline basis).
#include "gf_common.cuh"
#include <stdlib.h>

namespace gf {

struct WarpSmem {
  double Xc[16][4];
  double uc[16][3];
  double Phi[6][16];
  double Nraw[16];
  double g[32];
  double Ed[32];
  double Jd[32];
  double Hc[15][33];
  double G[15][48];
  double tw[16];
  double the[16];
  int ninfo[16][8];  // cpl, I, J, Ilo, WI, Jlo, S, nlow
  int tdof[16];
};

constexpr int MODE_K = 0, MODE_P = 1, MODE_T = 2;

template <int MODE>
__global__ void __launch_bounds__(128)
k_shell(GfModel M, GfShellOut O, int what, int color_begin, int color_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSmem& S = reinterpret_cast<WarpSmem*>(smem_raw)[warp];
  const int slot = blockIdx.x * (blockDim.x >> 5) + warp;
  if (slot >= color_count) return;
  const int el = M.color_elem[color_begin + slot];
  const GfPatchDesc P = M.patches[M.elem_patch[el]];
  const int eu = M.elem_eu[el], ev = M.elem_ev[el];
  const int su = P.span_u_off + eu, sv = P.span_v_off + ev;
  const int I0 = M.first_cp_u[su], J0 = M.first_cp_v[sv];
  const int ncp = P.n_u * P.n_v;
  const double area = M.span_h_u[su] * M.span_h_v[sv];
  const int nq = M.nq;

  // ---- element setup -------------------------------------------------------
  if (lane < 16) {
    const int lu = lane & 3, lv = lane >> 2;
    const int I = I0 + lu, J = J0 + lv;
    const int cpl = I + J * P.n_u;
    const double4 c = reinterpret_cast<const double4*>(M.cp)[P.cp_off + cpl];
    S.Xc[lane][0] = c.x; S.Xc[lane][1] = c.y; S.Xc[lane][2] = c.z; S.Xc[lane][3] = c.w;
    const double* up = M.u + P.dof_off + cpl;
    S.uc[lane][0] = up[0]; S.uc[lane][1] = up[ncp]; S.uc[lane][2] = up[2 * (size_t)ncp];
    const int Ilo = M.cp_lo_u[P.cpd_u_off + I], Ihi = M.cp_hi_u[P.cpd_u_off + I];
    const int Jlo = M.cp_lo_v[P.cpd_v_off + J], Jhi = M.cp_hi_v[P.cpd_v_off + J];
    S.ninfo[lane][0] = cpl; S.ninfo[lane][1] = I; S.ninfo[lane][2] = J;
    S.ninfo[lane][3] = Ilo; S.ninfo[lane][4] = Ihi - Ilo + 1; S.ninfo[lane][5] = Jlo;
    S.ninfo[lane][6] = (Ihi - Ilo + 1) * (Jhi - Jlo + 1);
    S.ninfo[lane][7] = M.row_nlow[P.cp_off + cpl];
  }
  // thickness dofs of the element
  int nt = 1;
  if (P.th_kind == GF_TH_LINEAR) nt = 4; else if (P.th_kind == GF_TH_IGA) nt = 16;
  if (lane < nt) {
    int td;
    if (P.th_kind == GF_TH_CONST) td = 0;
    else if (P.th_kind == GF_TH_LINEAR) td = (eu + (lane & 1)) + (ev + (lane >> 1)) * (P.neu + 1);
    else td = (I0 + (lane & 3)) + (J0 + (lane >> 2)) * P.n_u;
    S.tdof[lane] = td;
    S.the[lane] = M.theta[P.th_off + td];
  }
  __syncwarp();

  // ---- accumulators ----------------------------------------------------------
  constexpr int NACC = (MODE == MODE_T) ? 24 : 72;
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
  double racc0 = 0.0, racc1 = 0.0, vacc0 = 0.0, vacc1 = 0.0, wsum = 0.0, vsum = 0.0;
  const int ag = lane >> 3, bg = lane & 7;  // register tile of stage 2
  const int half = lane >> 4, an = lane & 15;
  const int mh = (nt + 1) >> 1, m0 = half * mh;

  const double* tu = M.tab_u + (size_t)su * nq * 12;
  const double* tv = M.tab_v + (size_t)sv * nq * 12;

  for (int q = 0; q < nq; ++q) {
    // ---- A. basis -------------------------------------------------------------
    {
      const int lu = an & 3, lv = an >> 2;
      const double* a = tu + q * 12;
      const double* b = tv + q * 12;
      const double u0 = a[lu], u1 = a[4 + lu], u2 = a[8 + lu];
      const double v0 = b[lv], v1 = b[4 + lv], v2 = b[8 + lv];
      double N = u0 * v0, Nu = u1 * v0, Nv = u0 * v1, Nuu = u2 * v0, Nvv = u0 * v2, Nuv = u1 * v1;
      const double nraw = N;
      if (P.rational) {
        const double w = S.Xc[an][3];
        double W = N * w, Wu = Nu * w, Wv = Nv * w, Wuu = Nuu * w, Wvv = Nvv * w, Wuv = Nuv * w;
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
          W += __shfl_xor_sync(0xffffffffu, W, o);
          Wu += __shfl_xor_sync(0xffffffffu, Wu, o);
          Wv += __shfl_xor_sync(0xffffffffu, Wv, o);
          Wuu += __shfl_xor_sync(0xffffffffu, Wuu, o);
          Wvv += __shfl_xor_sync(0xffffffffu, Wvv, o);
          Wuv += __shfl_xor_sync(0xffffffffu, Wuv, o);
        }
        const double iW = 1.0 / W;
        const double f = N * iW;
        const double fu = (Nu - f * Wu) * iW;
        const double fv = (Nv - f * Wv) * iW;
        const double fuu = (Nuu - 2.0 * fu * Wu - f * Wuu) * iW;
        const double fvv = (Nvv - 2.0 * fv * Wv - f * Wvv) * iW;
        const double fuv = (Nuv - fu * Wv - fv * Wu - f * Wuv) * iW;
        N = f; Nu = fu; Nv = fv; Nuu = fuu; Nvv = fvv; Nuv = fuv;
      }
      if (lane < 16) {
        S.Phi[0][an] = N; S.Phi[1][an] = Nu; S.Phi[2][an] = Nv;
        S.Phi[3][an] = Nuu; S.Phi[4][an] = Nvv; S.Phi[5][an] = Nuv;
        S.Nraw[an] = nraw;
        if (P.th_kind == GF_TH_IGA) S.tw[an] = nraw;
        else if (P.th_kind == GF_TH_LINEAR) { if (an < 4) S.tw[an] = M.tw_lin[q * 4 + an]; }
        else if (an == 0) S.tw[0] = 1.0;
      }
    }
    __syncwarp();
    // covariant vectors and thickness
    {
      double val = 0.0;
      if (lane < 15) {
        const int k = lane / 3 + 1, c = lane % 3;
#pragma unroll
        for (int a = 0; a < 16; ++a) val = fma(S.Phi[k][a], S.Xc[a][c], val);
      } else if (lane == 15) {
        for (int m = 0; m < nt; ++m) val = fma(S.tw[m], S.the[m], val);
      } else if (lane < 31) {
        const int k = (lane - 16) / 3 + 1, c = (lane - 16) % 3;
#pragma unroll
        for (int a = 0; a < 16; ++a) val = fma(S.Phi[k][a], S.uc[a][c], val);
      }
      // layout g: [gX 0..14 | t 15 | gu 16..30]
      S.g[lane] = val;
    }
    __syncwarp();
    // ---- B. point first variation with a per-lane dual direction ----------------
    Dual gX[15], gu[15], grad[15], e, J;
#pragma unroll
    for (int k = 0; k < 15; ++k) {
      gX[k] = Dual(S.g[k], (lane == 16 + k) ? 1.0 : 0.0);
      gu[k] = Dual(S.g[16 + k], (lane == k) ? 1.0 : 0.0);
    }
    const double tq = S.g[15];
    const Dual td(tq, (lane == 15) ? 1.0 : 0.0);
    kl_shell_point<Dual>(gX, gu, td, P.E, P.nu, e, J, grad);
    const double wq = M.qw[q] * area;
#pragma unroll
    for (int m = 0; m < 15; ++m) S.Hc[m][lane] = grad[m].d;
    S.Ed[lane] = e.d;
    S.Jd[lane] = J.d;
    __syncwarp();

    if (MODE == MODE_K) {
      // ---- C1. G[m][cb] = w sum_l Huu[m][(l,j)] Phi[1+l][b],  cb = 3 b + j ----
      for (int o = lane; o < 720; o += 32) {
        const int m = o / 48, cb = o - m * 48;
        const int b = cb / 3, j = cb - 3 * b;
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < 5; ++l) s = fma(S.Hc[m][l * 3 + j], S.Phi[1 + l][b], s);
        S.G[m][cb] = wq * s;
      }
      __syncwarp();
      // ---- C2. acc[i][aa][cc] += Phi[1+k][4ag+aa] G[3k+i][6bg+cc] ----
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        double ph[4];
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) ph[aa] = S.Phi[1 + k][4 * ag + aa];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double gg[6];
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) gg[cc] = S.G[3 * k + i][6 * bg + cc];
#pragma unroll
          for (int aa = 0; aa < 4; ++aa)
#pragma unroll
            for (int cc = 0; cc < 6; ++cc)
              acc[(i * 4 + aa) * 6 + cc] = fma(ph[aa], gg[cc], acc[(i * 4 + aa) * 6 + cc]);
        }
      }
      // residual: lanes 0..15 -> (a, i=0,1); lanes 16..31 -> (a, i=2)
      {
        const int i0 = half ? 2 : 0;
        double r0 = 0.0, r1 = 0.0;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          r0 = fma(S.Phi[1 + k][an], grad[3 * k + i0].v, r0);
          r1 = fma(S.Phi[1 + k][an], grad[3 * k + 1].v, r1);
        }
        const double jf = J.v * S.Phi[0][an];
        racc0 += wq * (r0 - jf * P.f[i0]);
        racc1 += wq * (r1 - jf * P.f[1]);
      }
      wsum += wq * e.v;
      vsum += wq * J.v * tq;
    } else if (MODE == MODE_P) {
      // ---- G_f[m][b] = w sum_l HuX[m][(l,f)] Phi[1+l][b] -> S.G[m][16 f + b] ----
      for (int o = lane; o < 720; o += 32) {
        const int m = o / 48, fb = o - m * 48;
        const int f = fb >> 4, b = fb & 15;
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < 5; ++l) s = fma(S.Hc[m][16 + l * 3 + f], S.Phi[1 + l][b], s);
        S.G[m][fb] = wq * s;
      }
      __syncwarp();
      // acc[f][i][aa][bb] += Phi[1+k][4ag+aa] G[3k+i][16 f + 2bg + bb]
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        double ph[4];
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) ph[aa] = S.Phi[1 + k][4 * ag + aa];
#pragma unroll
        for (int f = 0; f < 3; ++f)
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const double g0 = S.G[3 * k + i][16 * f + 2 * bg], g1 = S.G[3 * k + i][16 * f + 2 * bg + 1];
#pragma unroll
            for (int aa = 0; aa < 4; ++aa) {
              acc[((f * 3 + i) * 4 + aa) * 2 + 0] = fma(ph[aa], g0, acc[((f * 3 + i) * 4 + aa) * 2 + 0]);
              acc[((f * 3 + i) * 4 + aa) * 2 + 1] = fma(ph[aa], g1, acc[((f * 3 + i) * 4 + aa) * 2 + 1]);
            }
          }
      }
      // dead-load part: - w f_i Phi[0][a] (sum_l dJ/dgX[(l,f)] Phi[1+l][b]);  dJ only for l = 0,1
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        double dj[2];
#pragma unroll
        for (int bb = 0; bb < 2; ++bb)
          dj[bb] = wq * (S.Jd[16 + f] * S.Phi[1][2 * bg + bb] + S.Jd[19 + f] * S.Phi[2][2 * bg + bb]);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int aa = 0; aa < 4; ++aa) {
            const double c = -P.f[i] * S.Phi[0][4 * ag + aa];
            acc[((f * 3 + i) * 4 + aa) * 2 + 0] = fma(c, dj[0], acc[((f * 3 + i) * 4 + aa) * 2 + 0]);
            acc[((f * 3 + i) * 4 + aa) * 2 + 1] = fma(c, dj[1], acc[((f * 3 + i) * 4 + aa) * 2 + 1]);
          }
      }
      // dW/dCP_f[b], dV/dCP_f[b]: lanes 0..15 -> f = 0,1 ; lanes 16..31 -> f = 2
      {
        const int f0 = half ? 2 : 0;
        double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int l = 0; l < 5; ++l) {
          s0 = fma(S.Ed[16 + 3 * l + f0], S.Phi[1 + l][an], s0);
          s1 = fma(S.Ed[16 + 3 * l + 1], S.Phi[1 + l][an], s1);
        }
#pragma unroll
        for (int l = 0; l < 2; ++l) {
          t0 = fma(S.Jd[16 + 3 * l + f0], S.Phi[1 + l][an], t0);
          t1 = fma(S.Jd[16 + 3 * l + 1], S.Phi[1 + l][an], t1);
        }
        racc0 += wq * s0; racc1 += wq * s1;
        vacc0 += wq * tq * t0; vacc1 += wq * tq * t1;
      }
    } else {  // MODE_T
      // r[(a,i)] = w sum_k Phi[1+k][a] d grad[(k,i)]/dt ; acc[i][mm] += r * tw[m0+mm]
      double r[3], du[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double s = 0.0, s2 = 0.0;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          s = fma(S.Phi[1 + k][an], S.Hc[3 * k + i][15], s);
          s2 = fma(S.Phi[1 + k][an], grad[3 * k + i].v, s2);
        }
        r[i] = wq * s; du[i] = wq * s2;
      }
#pragma unroll
      for (int mm = 0; mm < 8; ++mm) {
        const int m = m0 + mm;
        const double t = (mm < mh && m < nt) ? S.tw[m] : 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) acc[i * 8 + mm] = fma(r[i], t, acc[i * 8 + mm]);
      }
      // dW/du: lanes 0..15 -> i = 0,1 ; lanes 16..31 -> i = 2
      racc0 += half ? du[2] : du[0];
      racc1 += du[1];
      // dW/dt[m], dV/dt[m] on lanes m < nt
      if (lane < nt) {
        vacc0 += wq * S.Ed[15] * S.tw[lane];
        vacc1 += wq * J.v * S.tw[lane];
      }
    }
    __syncwarp();
  }

  // ---- scatter ---------------------------------------------------------------
  const size_t dof0 = (size_t)P.dof_off;
  if (MODE == MODE_K) {
    if (what & GF_OUT_K) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) {
          const int a = 4 * ag + aa;
          const int* na = S.ninfo[a];
          const size_t row = dof0 + (size_t)i * ncp + na[0];
          const bool rbc = M.bc[row];
          const int64_t base = M.K.indptr[row] + na[7];
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) {
            const int cb = 6 * bg + cc, b = cb / 3, j = cb - 3 * b;
            const int* nb = S.ninfo[b];
            const size_t col = dof0 + (size_t)j * ncp + nb[0];
            if (rbc || M.bc[col]) continue;
            const int64_t pos = base + (int64_t)j * na[6] + (nb[2] - na[5]) * na[4] + (nb[1] - na[3]);
            M.K.vals[pos] += acc[(i * 4 + aa) * 6 + cc];
          }
        }
    }
    if (what & GF_OUT_R) {
      const int cpl = S.ninfo[an][0];
      if (half == 0) {
        O.R[dof0 + cpl] += racc0;
        O.R[dof0 + ncp + cpl] += racc1;
      } else {
        O.R[dof0 + 2 * (size_t)ncp + cpl] += racc0;
      }
    }
    if (what & GF_OUT_W) {
      // all lanes hold identical wsum / vsum
      if (lane == 0) { O.WV[2 * (size_t)el] = wsum; O.WV[2 * (size_t)el + 1] = vsum; }
    }
  } else if (MODE == MODE_P) {
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      if (P.pcol_off[f] < 0 || M.P[f].vals == nullptr) continue;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) {
          const int a = 4 * ag + aa;
          const int* na = S.ninfo[a];
          const size_t row = dof0 + (size_t)i * ncp + na[0];
          if (M.bc[row]) continue;
          const int64_t base = M.P[f].indptr[row];
#pragma unroll
          for (int bb = 0; bb < 2; ++bb) {
            const int* nb = S.ninfo[2 * bg + bb];
            const int64_t pos = base + (nb[2] - na[5]) * na[4] + (nb[1] - na[3]);
            M.P[f].vals[pos] += acc[((f * 3 + i) * 4 + aa) * 2 + bb];
          }
        }
    }
    {
      const int cpl = S.ninfo[an][0];
      const int f0 = half ? 2 : 0;
      if (P.pcol_off[f0] >= 0 && O.dWdP[f0]) {
        O.dWdP[f0][P.pcol_off[f0] + cpl] += racc0;
        if (O.dVdP[f0]) O.dVdP[f0][P.pcol_off[f0] + cpl] += vacc0;
      }
      if (half == 0 && P.pcol_off[1] >= 0 && O.dWdP[1]) {
        O.dWdP[1][P.pcol_off[1] + cpl] += racc1;
        if (O.dVdP[1]) O.dVdP[1][P.pcol_off[1] + cpl] += vacc1;
      }
    }
  } else {
    if (M.T.vals != nullptr) {
      const int* na = S.ninfo[an];
      // stencil of the thickness columns of row node `an`
      int lo_u = 0, wu = 1, lo_v = 0;
      if (P.th_kind == GF_TH_LINEAR) {
        lo_u = M.el_lo_u[P.cpd_u_off + na[1]];
        wu = M.el_hi_u[P.cpd_u_off + na[1]] - lo_u + 2;
        lo_v = M.el_lo_v[P.cpd_v_off + na[2]];
      } else if (P.th_kind == GF_TH_IGA) {
        lo_u = na[3]; wu = na[4]; lo_v = na[5];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const size_t row = dof0 + (size_t)i * ncp + na[0];
        const int64_t base = M.T.indptr[row];
#pragma unroll
        for (int mm = 0; mm < 8; ++mm) {
          const int m = m0 + mm;
          if (!(mm < mh && m < nt)) continue;
          int64_t pos = base;
          if (P.th_kind == GF_TH_LINEAR) {
            const int vI = eu + (m & 1), vJ = ev + (m >> 1);
            pos += (vJ - lo_v) * wu + (vI - lo_u);
          } else if (P.th_kind == GF_TH_IGA) {
            const int* nb = S.ninfo[m];
            pos += (nb[2] - lo_v) * wu + (nb[1] - lo_u);
          }
          M.T.vals[pos] += acc[i * 8 + mm];
        }
      }
    }
    if (O.dWdu) {
      const int cpl = S.ninfo[an][0];
      if (half == 0) { O.dWdu[dof0 + cpl] += racc0; O.dWdu[dof0 + ncp + cpl] += racc1; }
      else O.dWdu[dof0 + 2 * (size_t)ncp + cpl] += racc0;
    }
    if (P.th_kind == GF_TH_CONST) {
      // one thickness dof per patch: every element of the patch hits it, so the
      // per-element values are reduced afterwards in a fixed order (k_reduce_dt)
      if (lane == 0 && O.dt_el) { O.dt_el[2 * (size_t)el] = vacc0; O.dt_el[2 * (size_t)el + 1] = vacc1; }
    } else if (lane < nt) {
      if (O.dWdt) O.dWdt[P.th_off + S.tdof[lane]] += vacc0;
      if (O.dVdt) O.dVdt[P.th_off + S.tdof[lane]] += vacc1;
    }
  }
}

// dW/dt, dV/dt of constant-thickness patches: fixed-order sum over the patch's elements
__global__ void __launch_bounds__(256) k_reduce_dt(GfModel M, GfShellOut O) {
  __shared__ double sh[2][8];
  const GfPatchDesc P = M.patches[blockIdx.x];
  if (P.th_kind != GF_TH_CONST) return;
  const int nel = P.neu * P.nev;
  double a = 0.0, b = 0.0;
  for (int e = threadIdx.x; e < nel; e += blockDim.x) {
    a += O.dt_el[2 * (size_t)(P.el_off + e)];
    b += O.dt_el[2 * (size_t)(P.el_off + e) + 1];
  }
  a = warp_sum(a); b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0;
    for (int w = 0; w < 8; ++w) { sa += sh[0][w]; sb += sh[1][w]; }
    if (O.dWdt) O.dWdt[P.th_off] += sa;
    if (O.dVdt) O.dVdt[P.th_off] += sb;
  }
}

template <int MODE>
static int launch_mode(const GfModel* m, int what, const GfShellOut* out, cudaStream_t st) {
  const size_t smem = 4 * sizeof(WarpSmem);
  cudaError_t e = cudaFuncSetAttribute(k_shell<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(k_shell)");
  for (int c = 0; c < m->num_colors; ++c) {
    const int b = m->color_ptr_h[c], n = m->color_ptr_h[c + 1] - b;
    if (n <= 0) continue;
    k_shell<MODE><<<(n + 3) / 4, 128, smem, st>>>(*m, *out, what, b, n);
    count_launch(1);
  }
  count_launch(-1);
  return check_launch("k_shell");
}


// ---------------------------------------------------------------------------------
// Tangent pass, version 2: two quadrature points per warp pass.
// Only the 15 g_u directions are needed for K, so lanes 0..15 serve point 2*it and
// lanes 16..31 point 2*it+1 (lane & 15 = direction, lane 15 / 31 carry no seed).
// The reference-configuration part of the point routine runs in plain doubles.
// ---------------------------------------------------------------------------------
struct WarpSmemK2 {
  double Xc[16][4];
  double uc[16][3];
  double Phi[2][6][16];
  double g[2][32];       // [gX 0..14 | t 15 | gu 16..30]
  double Gv[2][16];      // grad values, [15] = J
  double Ev[2];
  double Hc[2][15][17];  // Hc[h][m][d] = d grad_m / d gu_d
  double G[15][48];
  double tw[2][16];
  double the[16];
  int ninfo[16][8];
};

template <int OCC>      // resident CTAs per SM asked of the register allocator: 2 -> 255 registers, 3 -> 168 (experiment knob GF_SHELL_OCC)
__global__ void __launch_bounds__(128, OCC)
k_shell_k2(GfModel M, GfShellOut O, int what, int color_begin, int color_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSmemK2& S = reinterpret_cast<WarpSmemK2*>(smem_raw)[warp];
  const int slot = blockIdx.x * (blockDim.x >> 5) + warp;
  if (slot >= color_count) return;
  const int el = M.color_elem[color_begin + slot];
  const GfPatchDesc P = M.patches[M.elem_patch[el]];
  const int eu = M.elem_eu[el], ev = M.elem_ev[el];
  const int su = P.span_u_off + eu, sv = P.span_v_off + ev;
  const int I0 = M.first_cp_u[su], J0 = M.first_cp_v[sv];
  const int ncp = P.n_u * P.n_v;
  const double area = M.span_h_u[su] * M.span_h_v[sv];
  const int nq = M.nq;
  const int half = lane >> 4, an = lane & 15;

  if (lane < 16) {
    const int lu = lane & 3, lv = lane >> 2;
    const int I = I0 + lu, J = J0 + lv;
    const int cpl = I + J * P.n_u;
    const double4 c = reinterpret_cast<const double4*>(M.cp)[P.cp_off + cpl];
    S.Xc[lane][0] = c.x; S.Xc[lane][1] = c.y; S.Xc[lane][2] = c.z; S.Xc[lane][3] = c.w;
    const double* up = M.u + P.dof_off + cpl;
    S.uc[lane][0] = up[0]; S.uc[lane][1] = up[ncp]; S.uc[lane][2] = up[2 * (size_t)ncp];
    const int Ilo = M.cp_lo_u[P.cpd_u_off + I], Ihi = M.cp_hi_u[P.cpd_u_off + I];
    const int Jlo = M.cp_lo_v[P.cpd_v_off + J], Jhi = M.cp_hi_v[P.cpd_v_off + J];
    S.ninfo[lane][0] = cpl; S.ninfo[lane][1] = I; S.ninfo[lane][2] = J;
    S.ninfo[lane][3] = Ilo; S.ninfo[lane][4] = Ihi - Ilo + 1; S.ninfo[lane][5] = Jlo;
    S.ninfo[lane][6] = (Ihi - Ilo + 1) * (Jhi - Jlo + 1);
    S.ninfo[lane][7] = M.row_nlow[P.cp_off + cpl];
  }
  int nt = 1;
  if (P.th_kind == GF_TH_LINEAR) nt = 4; else if (P.th_kind == GF_TH_IGA) nt = 16;
  if (lane < nt) {
    int td;
    if (P.th_kind == GF_TH_CONST) td = 0;
    else if (P.th_kind == GF_TH_LINEAR) td = (eu + (lane & 1)) + (ev + (lane >> 1)) * (P.neu + 1);
    else td = (I0 + (lane & 3)) + (J0 + (lane >> 2)) * P.n_u;
    S.the[lane] = M.theta[P.th_off + td];
  }
  __syncwarp();

  double acc[72];
#pragma unroll
  for (int i = 0; i < 72; ++i) acc[i] = 0.0;
  double racc0 = 0.0, racc1 = 0.0, wsum = 0.0, vsum = 0.0;
  const int ag = lane >> 3, bg = lane & 7;
  const double* tu = M.tab_u + (size_t)su * nq * 12;
  const double* tv = M.tab_v + (size_t)sv * nq * 12;

  for (int q0 = 0; q0 < nq; q0 += 2) {
    // this half's quadrature point (the second half idles with zero weight on an odd tail)
    const bool valid = (q0 + half) < nq;
    const int q = valid ? q0 + half : q0;
    // ---- A. basis of both points (lanes 0..15: point q0, lanes 16..31: point q0+1) ----
    {
      const int lu = an & 3, lv = an >> 2;
      const double* a = tu + q * 12;
      const double* b = tv + q * 12;
      const double u0 = a[lu], u1 = a[4 + lu], u2 = a[8 + lu];
      const double v0 = b[lv], v1 = b[4 + lv], v2 = b[8 + lv];
      double N = u0 * v0, Nu = u1 * v0, Nv = u0 * v1, Nuu = u2 * v0, Nvv = u0 * v2, Nuv = u1 * v1;
      const double nraw = N;
      if (P.rational) {
        const double w = S.Xc[an][3];
        double W = N * w, Wu = Nu * w, Wv = Nv * w, Wuu = Nuu * w, Wvv = Nvv * w, Wuv = Nuv * w;
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
          W += __shfl_xor_sync(0xffffffffu, W, o);
          Wu += __shfl_xor_sync(0xffffffffu, Wu, o);
          Wv += __shfl_xor_sync(0xffffffffu, Wv, o);
          Wuu += __shfl_xor_sync(0xffffffffu, Wuu, o);
          Wvv += __shfl_xor_sync(0xffffffffu, Wvv, o);
          Wuv += __shfl_xor_sync(0xffffffffu, Wuv, o);
        }
        const double iW = 1.0 / W;
        const double f = N * iW;
        const double fu = (Nu - f * Wu) * iW;
        const double fv = (Nv - f * Wv) * iW;
        const double fuu = (Nuu - 2.0 * fu * Wu - f * Wuu) * iW;
        const double fvv = (Nvv - 2.0 * fv * Wv - f * Wvv) * iW;
        const double fuv = (Nuv - fu * Wv - fv * Wu - f * Wuv) * iW;
        N = f; Nu = fu; Nv = fv; Nuu = fuu; Nvv = fvv; Nuv = fuv;
      }
      S.Phi[half][0][an] = N; S.Phi[half][1][an] = Nu; S.Phi[half][2][an] = Nv;
      S.Phi[half][3][an] = Nuu; S.Phi[half][4][an] = Nvv; S.Phi[half][5][an] = Nuv;
      if (P.th_kind == GF_TH_IGA) S.tw[half][an] = nraw;
      else if (P.th_kind == GF_TH_LINEAR) { if (an < 4) S.tw[half][an] = M.tw_lin[q * 4 + an]; }
      else if (an == 0) S.tw[half][0] = 1.0;
    }
    __syncwarp();
    // covariant vectors and thickness of both points: two rounds of 31 dot products
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double val = 0.0;
      if (lane < 15) {
        const int k = lane / 3 + 1, c = lane % 3;
#pragma unroll
        for (int a = 0; a < 16; ++a) val = fma(S.Phi[h][k][a], S.Xc[a][c], val);
      } else if (lane == 15) {
        for (int m = 0; m < nt; ++m) val = fma(S.tw[h][m], S.the[m], val);
      } else if (lane < 31) {
        const int k = (lane - 16) / 3 + 1, c = (lane - 16) % 3;
#pragma unroll
        for (int a = 0; a < 16; ++a) val = fma(S.Phi[h][k][a], S.uc[a][c], val);
      }
      S.g[h][lane] = val;
    }
    __syncwarp();
    // ---- B. first variation; lane (half, d): direction d of point `half` ----
    {
      double gXd[15];
      Dual gu[15], grad[15], e;
#pragma unroll
      for (int k = 0; k < 15; ++k) {
        gXd[k] = S.g[half][k];
        gu[k] = Dual(S.g[half][16 + k], (an == k) ? 1.0 : 0.0);
      }
      KlRef<double> R;
      kl_reference<double>(gXd, P.E, P.nu, R);
      kl_shell_point_fixed_ref<Dual>(gXd, R, gu, Dual(S.g[half][15]), e, grad);
      if (an < 15) {
#pragma unroll
        for (int m = 0; m < 15; ++m) S.Hc[half][m][an] = grad[m].d;
      } else {
#pragma unroll
        for (int m = 0; m < 15; ++m) S.Gv[half][m] = grad[m].v;
        S.Gv[half][15] = R.J;
        S.Ev[half] = e.v;
      }
    }
    __syncwarp();
    // ---- C. contraction, one point after the other, all 32 lanes ----
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      if (q0 + h >= nq) break;
      const double wq = M.qw[q0 + h] * area;
      {
        // G[m][3b+j] = w sum_l H[m][(l,j)] Phi[1+l][b]: lane = b + 16 s owns column b;
        // s = 0: j = 0 (all m) and j = 2 (m < 8); s = 1: j = 1 (all m) and j = 2 (m >= 8).
        // Fully unrolled: no index arithmetic, the five Phi values stay in registers and the
        // H loads are warp-wide broadcasts.
        double pl[5];
#pragma unroll
        for (int l = 0; l < 5; ++l) pl[l] = wq * S.Phi[h][1 + l][an];
        const int j0 = half;                       // 0 or 1
#pragma unroll
        for (int m = 0; m < 15; ++m) {
          double s0 = 0.0;
#pragma unroll
          for (int l = 0; l < 5; ++l) s0 = fma(S.Hc[h][m][l * 3 + j0], pl[l], s0);
          S.G[m][3 * an + j0] = s0;
        }
#pragma unroll
        for (int mm = 0; mm < 8; ++mm) {
          const int m = mm + 8 * half;
          if (m < 15) {
            double s2 = 0.0;
#pragma unroll
            for (int l = 0; l < 5; ++l) s2 = fma(S.Hc[h][m][l * 3 + 2], pl[l], s2);
            S.G[m][3 * an + 2] = s2;
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        double ph[4];
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) ph[aa] = S.Phi[h][1 + k][4 * ag + aa];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double gg[6];
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) gg[cc] = S.G[3 * k + i][6 * bg + cc];
#pragma unroll
          for (int aa = 0; aa < 4; ++aa)
#pragma unroll
            for (int cc = 0; cc < 6; ++cc)
              acc[(i * 4 + aa) * 6 + cc] = fma(ph[aa], gg[cc], acc[(i * 4 + aa) * 6 + cc]);
        }
      }
      {
        const int i0 = half ? 2 : 0;
        double r0 = 0.0, r1 = 0.0;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          r0 = fma(S.Phi[h][1 + k][an], S.Gv[h][3 * k + i0], r0);
          r1 = fma(S.Phi[h][1 + k][an], S.Gv[h][3 * k + 1], r1);
        }
        const double Jv = S.Gv[h][15];
        const double jf = Jv * S.Phi[h][0][an];
        racc0 += wq * (r0 - jf * P.f[i0]);
        racc1 += wq * (r1 - jf * P.f[1]);
        wsum += wq * S.Ev[h];
        vsum += wq * Jv * S.g[h][15];
      }
      __syncwarp();
    }
  }

  const size_t dof0 = (size_t)P.dof_off;
  if (what & GF_OUT_K) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int aa = 0; aa < 4; ++aa) {
        const int a = 4 * ag + aa;
        const int* na = S.ninfo[a];
        const size_t row = dof0 + (size_t)i * ncp + na[0];
        const bool rbc = M.bc[row];
        const int64_t base = M.K.indptr[row] + na[7];
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {
          const int cb = 6 * bg + cc, b = cb / 3, j = cb - 3 * b;
          const int* nb = S.ninfo[b];
          const size_t col = dof0 + (size_t)j * ncp + nb[0];
          if (rbc || M.bc[col]) continue;
          const int64_t pos = base + (int64_t)j * na[6] + (nb[2] - na[5]) * na[4] + (nb[1] - na[3]);
          M.K.vals[pos] += acc[(i * 4 + aa) * 6 + cc];
        }
      }
  }
  if (what & GF_OUT_R) {
    const int cpl = S.ninfo[an][0];
    if (half == 0) { O.R[dof0 + cpl] += racc0; O.R[dof0 + ncp + cpl] += racc1; }
    else O.R[dof0 + 2 * (size_t)ncp + cpl] += racc0;
  }
  if (what & GF_OUT_W) {
    if (lane == 0) { O.WV[2 * (size_t)el] = wsum; O.WV[2 * (size_t)el + 1] = vsum; }
  }
}

static int launch_k2(const GfModel* m, int what, const GfShellOut* out, cudaStream_t st) {
  const size_t smem = 4 * sizeof(WarpSmemK2);
  static const int occ = getenv("GF_SHELL_OCC") ? atoi(getenv("GF_SHELL_OCC")) : 2;
  cudaError_t e = cudaFuncSetAttribute(k_shell_k2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_shell_k2<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(k_shell_k2)");
  for (int c = 0; c < m->num_colors; ++c) {
    const int b = m->color_ptr_h[c], n = m->color_ptr_h[c + 1] - b;
    if (n <= 0) continue;
    if (occ == 3) k_shell_k2<3><<<(n + 3) / 4, 128, smem, st>>>(*m, *out, what, b, n);
    else k_shell_k2<2><<<(n + 3) / 4, 128, smem, st>>>(*m, *out, what, b, n);
    count_launch(1);
  }
  count_launch(-1);
  return check_launch("k_shell_k2");
}

// ---------------------------------------------------------------------------------
// Shape / thickness passes, version 2: two quadrature points per warp pass, same lane
// layout as k_shell_k2.  PT = 1: dR/dCP_f, dW/dCP_f, dV/dCP_f (lane & 15 = g_X direction);
// PT = 2: dR/dt, dW/dt, dV/dt, dW/du (the single t direction, reference part in plain doubles).
// ---------------------------------------------------------------------------------
struct WarpSmemP2 {
  double Xc[16][4];
  double uc[16][3];
  double Phi[2][6][16];
  double g[2][32];
  double Gv[2][16];      // grad values; [15] = J
  double Ht[2][16];      // d grad / d t ; [15] = d e / d t
  double Ed[2][16];
  double Jd[2][16];
  double Hc[2][15][17];  // Hc[h][m][dx] = d grad_m / d gX_dx
  double G[15][48];
  double tw[2][16];
  double the[16];
  int ninfo[16][8];
  int tdof[16];
};

template <int PT>
__global__ void __launch_bounds__(128)
k_shell_p2(GfModel M, GfShellOut O, int color_begin, int color_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSmemP2& S = reinterpret_cast<WarpSmemP2*>(smem_raw)[warp];
  const int slot = blockIdx.x * (blockDim.x >> 5) + warp;
  if (slot >= color_count) return;
  const int el = M.color_elem[color_begin + slot];
  const GfPatchDesc P = M.patches[M.elem_patch[el]];
  const int eu = M.elem_eu[el], ev = M.elem_ev[el];
  const int su = P.span_u_off + eu, sv = P.span_v_off + ev;
  const int I0 = M.first_cp_u[su], J0 = M.first_cp_v[sv];
  const int ncp = P.n_u * P.n_v;
  const double area = M.span_h_u[su] * M.span_h_v[sv];
  const int nq = M.nq;
  const int half = lane >> 4, an = lane & 15;

  if (lane < 16) {
    const int lu = lane & 3, lv = lane >> 2;
    const int I = I0 + lu, J = J0 + lv;
    const int cpl = I + J * P.n_u;
    const double4 c = reinterpret_cast<const double4*>(M.cp)[P.cp_off + cpl];
    S.Xc[lane][0] = c.x; S.Xc[lane][1] = c.y; S.Xc[lane][2] = c.z; S.Xc[lane][3] = c.w;
    const double* up = M.u + P.dof_off + cpl;
    S.uc[lane][0] = up[0]; S.uc[lane][1] = up[ncp]; S.uc[lane][2] = up[2 * (size_t)ncp];
    const int Ilo = M.cp_lo_u[P.cpd_u_off + I], Ihi = M.cp_hi_u[P.cpd_u_off + I];
    const int Jlo = M.cp_lo_v[P.cpd_v_off + J], Jhi = M.cp_hi_v[P.cpd_v_off + J];
    S.ninfo[lane][0] = cpl; S.ninfo[lane][1] = I; S.ninfo[lane][2] = J;
    S.ninfo[lane][3] = Ilo; S.ninfo[lane][4] = Ihi - Ilo + 1; S.ninfo[lane][5] = Jlo;
    S.ninfo[lane][6] = (Ihi - Ilo + 1) * (Jhi - Jlo + 1);
    S.ninfo[lane][7] = M.row_nlow[P.cp_off + cpl];
  }
  int nt = 1;
  if (P.th_kind == GF_TH_LINEAR) nt = 4; else if (P.th_kind == GF_TH_IGA) nt = 16;
  if (lane < nt) {
    int td;
    if (P.th_kind == GF_TH_CONST) td = 0;
    else if (P.th_kind == GF_TH_LINEAR) td = (eu + (lane & 1)) + (ev + (lane >> 1)) * (P.neu + 1);
    else td = (I0 + (lane & 3)) + (J0 + (lane >> 2)) * P.n_u;
    S.tdof[lane] = td;
    S.the[lane] = M.theta[P.th_off + td];
  }
  __syncwarp();

  constexpr int NACC = (PT == 2) ? 24 : 72;
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
  double racc0 = 0.0, racc1 = 0.0, vacc0 = 0.0, vacc1 = 0.0;
  const int ag = lane >> 3, bg = lane & 7;
  const int mh = (nt + 1) >> 1, m0 = half * mh;
  const double* tu = M.tab_u + (size_t)su * nq * 12;
  const double* tv = M.tab_v + (size_t)sv * nq * 12;

  for (int q0 = 0; q0 < nq; q0 += 2) {
    const bool valid = (q0 + half) < nq;
    const int q = valid ? q0 + half : q0;
    {
      const int lu = an & 3, lv = an >> 2;
      const double* a = tu + q * 12;
      const double* b = tv + q * 12;
      const double u0 = a[lu], u1 = a[4 + lu], u2 = a[8 + lu];
      const double v0 = b[lv], v1 = b[4 + lv], v2 = b[8 + lv];
      double N = u0 * v0, Nu = u1 * v0, Nv = u0 * v1, Nuu = u2 * v0, Nvv = u0 * v2, Nuv = u1 * v1;
      const double nraw = N;
      if (P.rational) {
        const double w = S.Xc[an][3];
        double W = N * w, Wu = Nu * w, Wv = Nv * w, Wuu = Nuu * w, Wvv = Nvv * w, Wuv = Nuv * w;
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
          W += __shfl_xor_sync(0xffffffffu, W, o);
          Wu += __shfl_xor_sync(0xffffffffu, Wu, o);
          Wv += __shfl_xor_sync(0xffffffffu, Wv, o);
          Wuu += __shfl_xor_sync(0xffffffffu, Wuu, o);
          Wvv += __shfl_xor_sync(0xffffffffu, Wvv, o);
          Wuv += __shfl_xor_sync(0xffffffffu, Wuv, o);
        }
        const double iW = 1.0 / W;
        const double f = N * iW;
        const double fu = (Nu - f * Wu) * iW;
        const double fv = (Nv - f * Wv) * iW;
        const double fuu = (Nuu - 2.0 * fu * Wu - f * Wuu) * iW;
        const double fvv = (Nvv - 2.0 * fv * Wv - f * Wvv) * iW;
        const double fuv = (Nuv - fu * Wv - fv * Wu - f * Wuv) * iW;
        N = f; Nu = fu; Nv = fv; Nuu = fuu; Nvv = fvv; Nuv = fuv;
      }
      S.Phi[half][0][an] = N; S.Phi[half][1][an] = Nu; S.Phi[half][2][an] = Nv;
      S.Phi[half][3][an] = Nuu; S.Phi[half][4][an] = Nvv; S.Phi[half][5][an] = Nuv;
      if (P.th_kind == GF_TH_IGA) S.tw[half][an] = nraw;
      else if (P.th_kind == GF_TH_LINEAR) { if (an < 4) S.tw[half][an] = M.tw_lin[q * 4 + an]; }
      else if (an == 0) S.tw[half][0] = 1.0;
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double val = 0.0;
      if (lane < 15) {
        const int k = lane / 3 + 1, c = lane % 3;
#pragma unroll
        for (int a = 0; a < 16; ++a) val = fma(S.Phi[h][k][a], S.Xc[a][c], val);
      } else if (lane == 15) {
        for (int m = 0; m < nt; ++m) val = fma(S.tw[h][m], S.the[m], val);
      } else if (lane < 31) {
        const int k = (lane - 16) / 3 + 1, c = (lane - 16) % 3;
#pragma unroll
        for (int a = 0; a < 16; ++a) val = fma(S.Phi[h][k][a], S.uc[a][c], val);
      }
      S.g[h][lane] = val;
    }
    __syncwarp();
    if (PT == 1) {
      // lane (half, d): derivative along g_X[d] at fixed u_hom (x = X + u moves with X)
      Dual gX[15], gu[15], grad[15], e, J;
#pragma unroll
      for (int k = 0; k < 15; ++k) {
        gX[k] = Dual(S.g[half][k], (an == k) ? 1.0 : 0.0);
        gu[k] = Dual(S.g[half][16 + k]);
      }
      kl_shell_point<Dual>(gX, gu, Dual(S.g[half][15]), P.E, P.nu, e, J, grad);
      if (an < 15) {
#pragma unroll
        for (int m = 0; m < 15; ++m) S.Hc[half][m][an] = grad[m].d;
        S.Ed[half][an] = e.d; S.Jd[half][an] = J.d;
      }
    } else {
      double gXd[15];
      Dual gu[15], grad[15], e;
#pragma unroll
      for (int k = 0; k < 15; ++k) { gXd[k] = S.g[half][k]; gu[k] = Dual(S.g[half][16 + k]); }
      KlRef<double> R;
      kl_reference<double>(gXd, P.E, P.nu, R);
      kl_shell_point_fixed_ref<Dual>(gXd, R, gu, Dual(S.g[half][15], 1.0), e, grad);
      if (an == 0) {
#pragma unroll
        for (int m = 0; m < 15; ++m) { S.Gv[half][m] = grad[m].v; S.Ht[half][m] = grad[m].d; }
        S.Gv[half][15] = R.J; S.Ht[half][15] = e.d;
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      if (q0 + h >= nq) break;
      const double wq = M.qw[q0 + h] * area;
      const double tq = S.g[h][15];
      if (PT == 1) {
        {
          // G_f[m][16 f + b]: lane = b + 16 s; s = 0: f = 0 (all m), f = 2 (m < 8); s = 1: f = 1, f = 2 (m >= 8)
          double pl[5];
#pragma unroll
          for (int l = 0; l < 5; ++l) pl[l] = wq * S.Phi[h][1 + l][an];
          const int f0 = half;
#pragma unroll
          for (int m = 0; m < 15; ++m) {
            double s0 = 0.0;
#pragma unroll
            for (int l = 0; l < 5; ++l) s0 = fma(S.Hc[h][m][l * 3 + f0], pl[l], s0);
            S.G[m][16 * f0 + an] = s0;
          }
#pragma unroll
          for (int mm = 0; mm < 8; ++mm) {
            const int m = mm + 8 * half;
            if (m < 15) {
              double s2 = 0.0;
#pragma unroll
              for (int l = 0; l < 5; ++l) s2 = fma(S.Hc[h][m][l * 3 + 2], pl[l], s2);
              S.G[m][32 + an] = s2;
            }
          }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          double ph[4];
#pragma unroll
          for (int aa = 0; aa < 4; ++aa) ph[aa] = S.Phi[h][1 + k][4 * ag + aa];
#pragma unroll
          for (int f = 0; f < 3; ++f)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const double g0 = S.G[3 * k + i][16 * f + 2 * bg], g1 = S.G[3 * k + i][16 * f + 2 * bg + 1];
#pragma unroll
              for (int aa = 0; aa < 4; ++aa) {
                acc[((f * 3 + i) * 4 + aa) * 2 + 0] = fma(ph[aa], g0, acc[((f * 3 + i) * 4 + aa) * 2 + 0]);
                acc[((f * 3 + i) * 4 + aa) * 2 + 1] = fma(ph[aa], g1, acc[((f * 3 + i) * 4 + aa) * 2 + 1]);
              }
            }
        }
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          double dj[2];
#pragma unroll
          for (int bb = 0; bb < 2; ++bb)
            dj[bb] = wq * (S.Jd[h][f] * S.Phi[h][1][2 * bg + bb] + S.Jd[h][3 + f] * S.Phi[h][2][2 * bg + bb]);
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int aa = 0; aa < 4; ++aa) {
              const double c = -P.f[i] * S.Phi[h][0][4 * ag + aa];
              acc[((f * 3 + i) * 4 + aa) * 2 + 0] = fma(c, dj[0], acc[((f * 3 + i) * 4 + aa) * 2 + 0]);
              acc[((f * 3 + i) * 4 + aa) * 2 + 1] = fma(c, dj[1], acc[((f * 3 + i) * 4 + aa) * 2 + 1]);
            }
        }
        {
          const int f0 = half ? 2 : 0;
          double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll
          for (int l = 0; l < 5; ++l) {
            s0 = fma(S.Ed[h][3 * l + f0], S.Phi[h][1 + l][an], s0);
            s1 = fma(S.Ed[h][3 * l + 1], S.Phi[h][1 + l][an], s1);
          }
#pragma unroll
          for (int l = 0; l < 2; ++l) {
            t0 = fma(S.Jd[h][3 * l + f0], S.Phi[h][1 + l][an], t0);
            t1 = fma(S.Jd[h][3 * l + 1], S.Phi[h][1 + l][an], t1);
          }
          racc0 += wq * s0; racc1 += wq * s1;
          vacc0 += wq * tq * t0; vacc1 += wq * tq * t1;
        }
      } else {
        double r[3], du[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double s = 0.0, s2 = 0.0;
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            s = fma(S.Phi[h][1 + k][an], S.Ht[h][3 * k + i], s);
            s2 = fma(S.Phi[h][1 + k][an], S.Gv[h][3 * k + i], s2);
          }
          r[i] = wq * s; du[i] = wq * s2;
        }
#pragma unroll
        for (int mm = 0; mm < 8; ++mm) {
          const int m = m0 + mm;
          const double t = (mm < mh && m < nt) ? S.tw[h][m] : 0.0;
#pragma unroll
          for (int i = 0; i < 3; ++i) acc[i * 8 + mm] = fma(r[i], t, acc[i * 8 + mm]);
        }
        racc0 += half ? du[2] : du[0];
        racc1 += du[1];
        if (lane < nt) {
          vacc0 += wq * S.Ht[h][15] * S.tw[h][lane];
          vacc1 += wq * S.Gv[h][15] * S.tw[h][lane];
        }
      }
      __syncwarp();
    }
  }

  const size_t dof0 = (size_t)P.dof_off;
  if (PT == 1) {
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      if (P.pcol_off[f] < 0 || M.P[f].vals == nullptr) continue;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) {
          const int a = 4 * ag + aa;
          const int* na = S.ninfo[a];
          const size_t row = dof0 + (size_t)i * ncp + na[0];
          if (M.bc[row]) continue;
          const int64_t base = M.P[f].indptr[row];
#pragma unroll
          for (int bb = 0; bb < 2; ++bb) {
            const int* nb = S.ninfo[2 * bg + bb];
            const int64_t pos = base + (nb[2] - na[5]) * na[4] + (nb[1] - na[3]);
            M.P[f].vals[pos] += acc[((f * 3 + i) * 4 + aa) * 2 + bb];
          }
        }
    }
    const int cpl = S.ninfo[an][0];
    const int f0 = half ? 2 : 0;
    if (P.pcol_off[f0] >= 0 && O.dWdP[f0]) {
      O.dWdP[f0][P.pcol_off[f0] + cpl] += racc0;
      if (O.dVdP[f0]) O.dVdP[f0][P.pcol_off[f0] + cpl] += vacc0;
    }
    if (half == 0 && P.pcol_off[1] >= 0 && O.dWdP[1]) {
      O.dWdP[1][P.pcol_off[1] + cpl] += racc1;
      if (O.dVdP[1]) O.dVdP[1][P.pcol_off[1] + cpl] += vacc1;
    }
  } else {
    if (M.T.vals != nullptr) {
      const int* na = S.ninfo[an];
      int lo_u = 0, wu = 1, lo_v = 0;
      if (P.th_kind == GF_TH_LINEAR) {
        lo_u = M.el_lo_u[P.cpd_u_off + na[1]];
        wu = M.el_hi_u[P.cpd_u_off + na[1]] - lo_u + 2;
        lo_v = M.el_lo_v[P.cpd_v_off + na[2]];
      } else if (P.th_kind == GF_TH_IGA) {
        lo_u = na[3]; wu = na[4]; lo_v = na[5];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const size_t row = dof0 + (size_t)i * ncp + na[0];
        const int64_t base = M.T.indptr[row];
#pragma unroll
        for (int mm = 0; mm < 8; ++mm) {
          const int m = m0 + mm;
          if (!(mm < mh && m < nt)) continue;
          int64_t pos = base;
          if (P.th_kind == GF_TH_LINEAR) {
            const int vI = eu + (m & 1), vJ = ev + (m >> 1);
            pos += (vJ - lo_v) * wu + (vI - lo_u);
          } else if (P.th_kind == GF_TH_IGA) {
            const int* nb = S.ninfo[m];
            pos += (nb[2] - lo_v) * wu + (nb[1] - lo_u);
          }
          M.T.vals[pos] += acc[i * 8 + mm];
        }
      }
    }
    if (O.dWdu) {
      const int cpl = S.ninfo[an][0];
      if (half == 0) { O.dWdu[dof0 + cpl] += racc0; O.dWdu[dof0 + ncp + cpl] += racc1; }
      else O.dWdu[dof0 + 2 * (size_t)ncp + cpl] += racc0;
    }
    if (P.th_kind == GF_TH_CONST) {
      if (lane == 0 && O.dt_el) { O.dt_el[2 * (size_t)el] = vacc0; O.dt_el[2 * (size_t)el + 1] = vacc1; }
    } else if (lane < nt) {
      if (O.dWdt) O.dWdt[P.th_off + S.tdof[lane]] += vacc0;
      if (O.dVdt) O.dVdt[P.th_off + S.tdof[lane]] += vacc1;
    }
  }
}

template <int PT>
static int launch_p2(const GfModel* m, const GfShellOut* out, cudaStream_t st) {
  const size_t smem = 4 * sizeof(WarpSmemP2);
  cudaError_t e = cudaFuncSetAttribute(k_shell_p2<PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(k_shell_p2)");
  for (int c = 0; c < m->num_colors; ++c) {
    const int b = m->color_ptr_h[c], n = m->color_ptr_h[c + 1] - b;
    if (n <= 0) continue;
    k_shell_p2<PT><<<(n + 3) / 4, 128, smem, st>>>(*m, *out, b, n);
    count_launch(1);
  }
  count_launch(-1);
  return check_launch("k_shell_p2");
}
}  // namespace gf

extern "C" int gf_shell_assemble(const GfModel* m, int what, const GfShellOut* out, void* stream) {
  if (!m || !out) return gf::set_error(GF_ERR_BADARG, "gf_shell_assemble: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = GF_OK;
  static const bool v1 = getenv("GF_SHELL_V1") != nullptr;   // keep the one-point-per-pass kernels reachable
  if (what & (GF_OUT_R | GF_OUT_K | GF_OUT_W)) {
    if ((what & GF_OUT_R) && !out->R) return gf::set_error(GF_ERR_BADARG, "GF_OUT_R without out->R");
    if ((what & GF_OUT_W) && !out->WV) return gf::set_error(GF_ERR_BADARG, "GF_OUT_W without out->WV");
    rc = v1 ? gf::launch_mode<gf::MODE_K>(m, what, out, st) : gf::launch_k2(m, what, out, st);
    if (rc) return rc;
  }
  if (what & GF_OUT_P) { rc = v1 ? gf::launch_mode<gf::MODE_P>(m, what, out, st) : gf::launch_p2<1>(m, out, st); if (rc) return rc; }
  if (what & GF_OUT_T) {
    rc = v1 ? gf::launch_mode<gf::MODE_T>(m, what, out, st) : gf::launch_p2<2>(m, out, st);
    if (rc) return rc;
    if (out->dt_el && (out->dWdt || out->dVdt)) {
      gf::k_reduce_dt<<<m->num_patches, 256, 0, st>>>(*m, *out);
      rc = gf::check_launch("k_reduce_dt");
    }
  }
  return rc;
}
