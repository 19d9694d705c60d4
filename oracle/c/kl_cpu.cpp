// ORACLE / CPU BASELINE (test + measurement infrastructure, never the product path).
//
// Compiled C++/OpenMP restatement of the shell quadrature + scatter of the
// reference's CPU path: what FFC-generated tabulate_tensor + DOLFIN assemble +
// PETSc MatPtAP do per form (/root/reference/GOLDFISH/nonmatching_opt.py:733-739,
// 779-781,852,936,688), written against the same plain-data model struct
// (include/goldfish_b200.h, HOST pointers here) so that bench.py can time the
// CPU side on the GPU box's host cores with all threads (`cpu_baseline.kind = "port"`).
// Derivatives by forward-mode dual numbers over the hand-derived first variation
// (same point mathematics header as the CUDA kernels, compiled for the host).
// Parity unpinned; checked against oracle/model.py in tests/.
//
//   g++ -O3 -march=native -fopenmp -shared -fPIC -I../../include oracle/c/kl_cpu.cpp
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>
#include <omp.h>
#include "../../include/goldfish_b200.h"
#include "../../goldfish_b200/csrc/kl_point.cuh"

using gf::Dual;

namespace {

struct Elem {
  double Xc[16][4], uc[16][3], Phi[6][16], tw[16], the[16];
  int cpl[16], I[16], J[16], Ilo[16], WI[16], Jlo[16], S[16], nlow[16], tdof[16];
};

void basis(const GfModel& M, const GfPatchDesc& P, int su, int sv, int q, Elem& E, int nt) {
  const double* a = M.tab_u + ((size_t)su * M.nq + q) * 12;
  const double* b = M.tab_v + ((size_t)sv * M.nq + q) * 12;
  double N[6][16];
  for (int n = 0; n < 16; ++n) {
    const int lu = n & 3, lv = n >> 2;
    N[0][n] = a[lu] * b[lv]; N[1][n] = a[4 + lu] * b[lv]; N[2][n] = a[lu] * b[4 + lv];
    N[3][n] = a[8 + lu] * b[lv]; N[4][n] = a[lu] * b[8 + lv]; N[5][n] = a[4 + lu] * b[4 + lv];
  }
  if (P.th_kind == GF_TH_IGA) for (int n = 0; n < 16; ++n) E.tw[n] = N[0][n];
  else if (P.th_kind == GF_TH_LINEAR) for (int n = 0; n < 4; ++n) E.tw[n] = M.tw_lin[q * 4 + n];
  else E.tw[0] = 1.0;
  (void)nt;
  if (P.rational) {
    double W[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 6; ++k) for (int n = 0; n < 16; ++n) W[k] += N[k][n] * E.Xc[n][3];
    const double iW = 1.0 / W[0];
    for (int n = 0; n < 16; ++n) {
      const double f = N[0][n] * iW;
      const double fu = (N[1][n] - f * W[1]) * iW, fv = (N[2][n] - f * W[2]) * iW;
      E.Phi[0][n] = f; E.Phi[1][n] = fu; E.Phi[2][n] = fv;
      E.Phi[3][n] = (N[3][n] - 2.0 * fu * W[1] - f * W[3]) * iW;
      E.Phi[4][n] = (N[4][n] - 2.0 * fv * W[2] - f * W[4]) * iW;
      E.Phi[5][n] = (N[5][n] - fu * W[2] - fv * W[1] - f * W[5]) * iW;
    }
  } else {
    memcpy(E.Phi, N, sizeof(N));
  }
}

void element(const GfModel& M, int what, const GfShellOut& O, int el) {
  const GfPatchDesc& P = M.patches[M.elem_patch[el]];
  const int eu = M.elem_eu[el], ev = M.elem_ev[el];
  const int su = P.span_u_off + eu, sv = P.span_v_off + ev;
  const int I0 = M.first_cp_u[su], J0 = M.first_cp_v[sv];
  const int ncp = P.n_u * P.n_v;
  const double area = M.span_h_u[su] * M.span_h_v[sv];
  Elem E;
  for (int n = 0; n < 16; ++n) {
    const int I = I0 + (n & 3), J = J0 + (n >> 2), cpl = I + J * P.n_u;
    for (int c = 0; c < 4; ++c) E.Xc[n][c] = M.cp[(size_t)(P.cp_off + cpl) * 4 + c];
    for (int c = 0; c < 3; ++c) E.uc[n][c] = M.u[P.dof_off + (size_t)c * ncp + cpl];
    const int Ilo = M.cp_lo_u[P.cpd_u_off + I], Ihi = M.cp_hi_u[P.cpd_u_off + I];
    const int Jlo = M.cp_lo_v[P.cpd_v_off + J], Jhi = M.cp_hi_v[P.cpd_v_off + J];
    E.cpl[n] = cpl; E.I[n] = I; E.J[n] = J; E.Ilo[n] = Ilo; E.WI[n] = Ihi - Ilo + 1; E.Jlo[n] = Jlo;
    E.S[n] = (Ihi - Ilo + 1) * (Jhi - Jlo + 1); E.nlow[n] = M.row_nlow[P.cp_off + cpl];
  }
  const int nt = P.th_kind == GF_TH_LINEAR ? 4 : (P.th_kind == GF_TH_IGA ? 16 : 1);
  for (int m = 0; m < nt; ++m) {
    int td = 0;
    if (P.th_kind == GF_TH_LINEAR) td = (eu + (m & 1)) + (ev + (m >> 1)) * (P.neu + 1);
    else if (P.th_kind == GF_TH_IGA) td = E.cpl[m];
    E.tdof[m] = td; E.the[m] = M.theta[P.th_off + td];
  }
  static thread_local std::vector<double> Kbuf, Pbuf, Tbuf;
  Kbuf.assign(48 * 48, 0.0); Pbuf.assign(3 * 48 * 16, 0.0); Tbuf.assign(48 * 16, 0.0);
  double Re[48] = {0}, dWdu[48] = {0}, dWdP[3][16] = {{0}}, dVdP[3][16] = {{0}}, dWdt[16] = {0}, dVdt[16] = {0};
  double W = 0.0, V = 0.0;
  const bool doK = what & GF_OUT_K, doP = what & GF_OUT_P, doT = what & GF_OUT_T;
  for (int q = 0; q < M.nq; ++q) {
    basis(M, P, su, sv, q, E, nt);
    const double wq = M.qw[q] * area;
    double gX[15], gu[15], tq = 0.0;
    for (int l = 0; l < 15; ++l) {
      double sx = 0.0, sg = 0.0;
      for (int n = 0; n < 16; ++n) { sx += E.Phi[1 + l / 3][n] * E.Xc[n][l % 3]; sg += E.Phi[1 + l / 3][n] * E.uc[n][l % 3]; }
      gX[l] = sx; gu[l] = sg;
    }
    for (int m = 0; m < nt; ++m) tq += E.tw[m] * E.the[m];
    // columns of the point Hessian: direction d = 0..14 g_u, 15 t, 16..30 g_X
    double H[31][15], Ed[31], Jd[31], gv[15], ev_ = 0.0, Jv = 0.0;
    const int d0 = doK ? 0 : 15, d1 = doP ? 31 : 16;
    for (int d = (doK || doT || doP) ? d0 : 31; d < d1; ++d) {
      if (d < 15 && !doK) continue;
      if (d == 15 && !doT) continue;
      Dual X[15], U[15], g[15], e, Jq;
      for (int k = 0; k < 15; ++k) { X[k] = Dual(gX[k], d == 16 + k ? 1.0 : 0.0); U[k] = Dual(gu[k], d == k ? 1.0 : 0.0); }
      gf::kl_shell_point<Dual>(X, U, Dual(tq, d == 15 ? 1.0 : 0.0), P.E, P.nu, e, Jq, g);
      for (int m = 0; m < 15; ++m) H[d][m] = g[m].d;
      Ed[d] = e.d; Jd[d] = Jq.d;
    }
    {
      double e, Jq;
      gf::kl_shell_point<double>(gX, gu, tq, P.E, P.nu, e, Jq, gv);
      ev_ = e; Jv = Jq;
    }
    W += wq * ev_; V += wq * Jv * tq;
    for (int n = 0; n < 16; ++n)
      for (int i = 0; i < 3; ++i) {
        double r = 0.0;
        for (int k = 0; k < 5; ++k) r += E.Phi[1 + k][n] * gv[3 * k + i];
        dWdu[n * 3 + i] += wq * r;
        Re[n * 3 + i] += wq * (r - Jv * P.f[i] * E.Phi[0][n]);
      }
    if (doK) {
      double G[15][48];
      for (int m = 0; m < 15; ++m)
        for (int b = 0; b < 16; ++b)
          for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int l = 0; l < 5; ++l) s += H[l * 3 + j][m] * E.Phi[1 + l][b];
            G[m][b * 3 + j] = wq * s;
          }
      for (int a = 0; a < 16; ++a)
        for (int i = 0; i < 3; ++i)
          for (int k = 0; k < 5; ++k) {
            const double ph = E.Phi[1 + k][a];
            double* Kr = &Kbuf[(a * 3 + i) * 48];
            const double* Gr = G[3 * k + i];
            for (int c = 0; c < 48; ++c) Kr[c] += ph * Gr[c];
          }
    }
    if (doP) {
      for (int f = 0; f < 3; ++f) {
        double G[15][16];
        for (int m = 0; m < 15; ++m)
          for (int b = 0; b < 16; ++b) {
            double s = 0.0;
            for (int l = 0; l < 5; ++l) s += H[16 + l * 3 + f][m] * E.Phi[1 + l][b];
            G[m][b] = wq * s;
          }
        for (int a = 0; a < 16; ++a)
          for (int i = 0; i < 3; ++i) {
            double* Pr = &Pbuf[((f * 16 + a) * 3 + i) * 16];
            for (int k = 0; k < 5; ++k) { const double ph = E.Phi[1 + k][a]; for (int b = 0; b < 16; ++b) Pr[b] += ph * G[3 * k + i][b]; }
            for (int b = 0; b < 16; ++b)
              Pr[b] -= wq * P.f[i] * E.Phi[0][a] * (Jd[16 + f] * E.Phi[1][b] + Jd[19 + f] * E.Phi[2][b]);
          }
        for (int b = 0; b < 16; ++b) {
          double s = 0.0, t = 0.0;
          for (int l = 0; l < 5; ++l) s += Ed[16 + 3 * l + f] * E.Phi[1 + l][b];
          for (int l = 0; l < 2; ++l) t += Jd[16 + 3 * l + f] * E.Phi[1 + l][b];
          dWdP[f][b] += wq * s; dVdP[f][b] += wq * tq * t;
        }
      }
    }
    if (doT) {
      for (int a = 0; a < 16; ++a)
        for (int i = 0; i < 3; ++i) {
          double r = 0.0;
          for (int k = 0; k < 5; ++k) r += E.Phi[1 + k][a] * H[15][3 * k + i];
          for (int m = 0; m < nt; ++m) Tbuf[(a * 3 + i) * 16 + m] += wq * r * E.tw[m];
        }
      for (int m = 0; m < nt; ++m) { dWdt[m] += wq * Ed[15] * E.tw[m]; dVdt[m] += wq * Jv * E.tw[m]; }
    }
  }
  // ---- scatter (same CSR arithmetic as the CUDA kernels) ----
  const size_t dof0 = P.dof_off;
  for (int a = 0; a < 16; ++a)
    for (int i = 0; i < 3; ++i) {
      const size_t row = dof0 + (size_t)i * ncp + E.cpl[a];
      if (what & GF_OUT_R) O.R[row] += Re[a * 3 + i];
      if (doT && O.dWdu) O.dWdu[row] += dWdu[a * 3 + i];
      if (doK) {
        const int64_t base = M.K.indptr[row] + E.nlow[a];
        for (int b = 0; b < 16; ++b)
          for (int j = 0; j < 3; ++j) {
            const size_t col = dof0 + (size_t)j * ncp + E.cpl[b];
            if (M.bc[row] || M.bc[col]) continue;
            M.K.vals[base + (int64_t)j * E.S[a] + (E.J[b] - E.Jlo[a]) * E.WI[a] + (E.I[b] - E.Ilo[a])] += Kbuf[(a * 3 + i) * 48 + b * 3 + j];
          }
      }
      if (doP)
        for (int f = 0; f < 3; ++f) {
          if (P.pcol_off[f] < 0 || !M.P[f].vals || M.bc[row]) continue;
          const int64_t base = M.P[f].indptr[row];
          for (int b = 0; b < 16; ++b)
            M.P[f].vals[base + (E.J[b] - E.Jlo[a]) * E.WI[a] + (E.I[b] - E.Ilo[a])] += Pbuf[((f * 16 + a) * 3 + i) * 16 + b];
        }
      if (doT && M.T.vals) {
        const int64_t base = M.T.indptr[row];
        for (int m = 0; m < nt; ++m) {
          int64_t pos = base;
          if (P.th_kind == GF_TH_LINEAR) {
            const int lo_u = M.el_lo_u[P.cpd_u_off + E.I[a]], wu = M.el_hi_u[P.cpd_u_off + E.I[a]] - lo_u + 2;
            const int lo_v = M.el_lo_v[P.cpd_v_off + E.J[a]];
            pos += ((ev + (m >> 1)) - lo_v) * wu + ((eu + (m & 1)) - lo_u);
          } else if (P.th_kind == GF_TH_IGA) {
            pos += (E.J[m] - E.Jlo[a]) * E.WI[a] + (E.I[m] - E.Ilo[a]);
          }
          M.T.vals[pos] += Tbuf[(a * 3 + i) * 16 + m];
        }
      }
    }
  if (doP)
    for (int f = 0; f < 3; ++f)
      if (P.pcol_off[f] >= 0 && O.dWdP[f])
        for (int b = 0; b < 16; ++b) {
          O.dWdP[f][P.pcol_off[f] + E.cpl[b]] += dWdP[f][b];
          if (O.dVdP[f]) O.dVdP[f][P.pcol_off[f] + E.cpl[b]] += dVdP[f][b];
        }
  if (doT)
    for (int m = 0; m < nt; ++m) {
      // constant thickness: all elements of a patch hit one dof -> per-element slot, summed by the caller
      if (P.th_kind == GF_TH_CONST) { if (O.dt_el) { O.dt_el[2 * (size_t)el] = dWdt[0]; O.dt_el[2 * (size_t)el + 1] = dVdt[0]; } }
      else { if (O.dWdt) O.dWdt[P.th_off + E.tdof[m]] += dWdt[m]; if (O.dVdt) O.dVdt[P.th_off + E.tdof[m]] += dVdt[m]; }
    }
  if ((what & GF_OUT_W) && O.WV) { O.WV[2 * (size_t)el] = W; O.WV[2 * (size_t)el + 1] = V; }
}

}  // namespace

// All pointers inside *m and *out are HOST pointers.  Elements of one colour share no
// control point, so a colour is an OpenMP parallel loop without atomics.
extern "C" int gfo_shell_assemble(const GfModel* m, int what, const GfShellOut* out) {
  for (int c = 0; c < m->num_colors; ++c) {
    const int b = m->color_ptr_h[c], n = m->color_ptr_h[c + 1] - b;
#pragma omp parallel for schedule(dynamic, 16)
    for (int s = 0; s < n; ++s) element(*m, what, *out, m->color_elem[b + s]);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Penalty coupling on the host (PENGoLINS transfer_penalty_residual[_deriv], GOLDFISH transfer_dRmdcpm_sub:
// /root/reference/GOLDFISH/nonmatching_opt.py:745-752, 789-801, 864-867; utils/opt_utils.py:212-260): point
// Hessians by forward-mode duals over penalty_point, then the same host-built destination lists as the CUDA
// path (GfPenalty / GfPenaltyP with HOST pointers), OpenMP over evaluations / destinations.
extern "C" int gfo_penalty_points(const GfModel* m, const GfPenalty* p, int with_X) {
  const GfModel& M = *m; const GfPenalty& Q = *p;
#pragma omp parallel for schedule(static)
  for (int64_t ev = 0; ev < Q.n_eval; ++ev) {
    double g[36];
    for (int lane = 0; lane < 18; ++lane) {
      const int side = lane / 9, kc = lane % 9, k = kc / 3, c = kc % 3;
      const int32_t* conn = (side ? Q.connB : Q.connA) + ev * 16;
      const double* bas = (side ? Q.basB : Q.basA) + ev * 48 + k * 16;
      const int32_t* dd = (side ? Q.dofB : Q.dofA) + ev * 3;
      const double* u = M.u + dd[0] + (size_t)c * dd[1] - dd[2];
      double s = 0.0;
      for (int a = 0; a < 16; ++a) s += bas[a] * u[conn[a]];
      g[lane] = s;
      const int blk = lane / 3;
      const int32_t* cx; const double* bx;
      if (blk == 0) { cx = Q.connC0 + ev * 16; bx = Q.basC0 + ev * 16; }
      else if (blk == 1) { cx = Q.connC1 + ev * 16; bx = Q.basC1 + ev * 16; }
      else if (blk < 4) { cx = Q.connA + ev * 16; bx = Q.basA + ev * 48 + (blk - 1) * 16; }
      else { cx = Q.connB + ev * 16; bx = Q.basB + ev * 48 + (blk - 3) * 16; }
      double x = 0.0;
      for (int a = 0; a < 16; ++a) x += bx[a] * M.cp[(size_t)cx[a] * 4 + (lane % 3)];
      g[18 + lane] = x;
    }
    const double tp[2] = {Q.tpar[ev * 2], Q.tpar[ev * 2 + 1]};
    const double ad = Q.alpha[ev * 2], ar = Q.alpha[ev * 2 + 1];
    for (int pass = 0; pass < (with_X ? 2 : 1); ++pass)
      for (int d = 0; d < 18; ++d) {
        Dual uv[18], Xv[18], grad[18], e;
        for (int k = 0; k < 18; ++k) {
          uv[k] = Dual(g[k], (pass == 0 && d == k) ? 1.0 : 0.0);
          Xv[k] = Dual(g[18 + k], (pass == 1 && d == k) ? 1.0 : 0.0);
        }
        gf::penalty_point<Dual>(uv, Xv, tp, ad, ar, e, grad);
        double* H = (pass == 0 ? Q.Huu : Q.HuX) + ev * 324;
        for (int mm = 0; mm < 18; ++mm) H[mm * 18 + d] = grad[mm].d;
        if (pass == 0 && d == 0) for (int mm = 0; mm < 18; ++mm) Q.g[ev * 18 + mm] = grad[mm].v;
      }
  }
  return 0;
}

extern "C" int gfo_penalty_gather_R(const GfPenalty* p, double* R) {
  const GfPenalty& Q = *p;
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < Q.nR; ++n) {
    double s[3] = {0.0, 0.0, 0.0};
    for (int64_t it = Q.R_ptr[n]; it < Q.R_ptr[n + 1]; ++it) {
      const int32_t item = Q.R_item[it];
      const int64_t ev = item >> 5;
      const int ln = item & 31, side = ln >> 4, a = ln & 15;
      const double* bas = (side ? Q.basB : Q.basA) + ev * 48;
      const double* g = Q.g + ev * 18 + side * 9;
      for (int k = 0; k < 3; ++k) for (int i = 0; i < 3; ++i) s[i] += bas[k * 16 + a] * g[3 * k + i];
    }
    for (int i = 0; i < 3; ++i) R[Q.R_row[n * 3 + i]] += s[i];
  }
  return 0;
}

extern "C" int gfo_penalty_gather_K(const GfPenalty* p, double* Kvals) {
  const GfPenalty& Q = *p;
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < Q.nK; ++n) {
    double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t it = Q.K_ptr[n]; it < Q.K_ptr[n + 1]; ++it) {
      const int32_t item = Q.K_item[it];
      const int64_t ev = item >> 10;
      const int la = (item >> 5) & 31, lb = item & 31;
      const int sa = la >> 4, a = la & 15, sb = lb >> 4, b = lb & 15;
      const double* bR = (sa ? Q.basB : Q.basA) + ev * 48;
      const double* bC = (sb ? Q.basB : Q.basA) + ev * 48;
      const double* H = Q.Huu + ev * 324 + (sa * 9) * 18 + sb * 9;
      for (int k = 0; k < 3; ++k)
        for (int l = 0; l < 3; ++l) {
          const double w = bR[k * 16 + a] * bC[l * 16 + b];
          for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s[i * 3 + j] += w * H[(3 * k + i) * 18 + 3 * l + j];
        }
    }
    for (int i = 0; i < 9; ++i) { const int64_t pos = Q.K_pos[n * 9 + i]; if (pos >= 0) Kvals[pos] += s[i]; }
  }
  return 0;
}

extern "C" int gfo_penalty_gather_P(const GfPenalty* p, const GfPenaltyP* pp) {
  const GfPenalty& Q = *p; const GfPenaltyP& PP = *pp;
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < PP.n_dest; ++n) {
    double s[3] = {0.0, 0.0, 0.0};
    for (int64_t it = PP.ptr[n]; it < PP.ptr[n + 1]; ++it) {
      const int64_t ev = PP.item_eval[it];
      const int code = PP.item_code[it];
      const int la = code & 31, xb = (code >> 5) & 7, lb = (code >> 8) & 15;
      const int sa = la >> 4, a = la & 15;
      const double* bR = (sa ? Q.basB : Q.basA) + ev * 48;
      double cC;
      if (xb == 0) cC = Q.basC0[ev * 16 + lb];
      else if (xb == 1) cC = Q.basC1[ev * 16 + lb];
      else if (xb < 4) cC = Q.basA[ev * 48 + (xb - 1) * 16 + lb];
      else cC = Q.basB[ev * 48 + (xb - 3) * 16 + lb];
      const double* H = Q.HuX + ev * 324 + (sa * 9) * 18 + xb * 3 + PP.field;
      for (int k = 0; k < 3; ++k) { const double w = bR[k * 16 + a] * cC; for (int i = 0; i < 3; ++i) s[i] += w * H[(3 * k + i) * 18]; }
    }
    for (int i = 0; i < 3; ++i) { const int64_t pos = PP.pos[n * 3 + i]; if (pos >= 0) PP.vals[pos] += s[i]; }
  }
  return 0;
}

// K[bc rows / cols] were masked during scatter; write the unit diagonal (zeroRowsColumns, nonmatching_opt.py:693-700)
extern "C" int gfo_bc_set_diag(const GfModel* m, double diag) {
  for (int64_t i = 0; i < m->n_bc; ++i) {
    const int64_t row = m->bc_list[i];
    for (int64_t k = m->K.indptr[row]; k < m->K.indptr[row + 1]; ++k) m->K.vals[k] = (m->K.indices[k] == row) ? diag : 0.0;
  }
  return 0;
}

// torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm sets its thread count explicitly
extern "C" void gfo_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
