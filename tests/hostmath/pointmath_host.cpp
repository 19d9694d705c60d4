// TEST-ONLY host build of the device point mathematics (goldfish_b200/csrc/kl_point.cuh)
// so that the hand-derived first variations and the dual-number second
// derivatives can be checked against the oracle without a GPU.  Not part of
// the product library; compiled by tests/test_pointmath_host.py with g++.
#include "../../goldfish_b200/csrc/kl_point.cuh"
using gf::Dual;
extern "C" {
// dir: [dgX(15) | dgu(15) | dt]; outputs value and directional derivative.
void gf_test_shell_point(const double* gX, const double* gu, double t, double E, double nu,
                         const double* dir, double* e, double* J, double* grad,
                         double* de, double* dJ, double* dgrad) {
  Dual X[15], U[15], g[15], ee, JJ;
  for (int k = 0; k < 15; ++k) { X[k] = Dual(gX[k], dir[k]); U[k] = Dual(gu[k], dir[15 + k]); }
  gf::kl_shell_point<Dual>(X, U, Dual(t, dir[30]), E, nu, ee, JJ, g);
  *e = ee.v; *de = ee.d; *J = JJ.v; *dJ = JJ.d;
  for (int k = 0; k < 15; ++k) { grad[k] = g[k].v; dgrad[k] = g[k].d; }
}
// dir: [duv(18) | dXv(18)]
void gf_test_penalty_point(const double* uv, const double* Xv, const double* tp, double ad,
                           double ar, const double* dir, double* e, double* grad,
                           double* de, double* dgrad) {
  Dual U[18], X[18], g[18], ee;
  for (int k = 0; k < 18; ++k) { U[k] = Dual(uv[k], dir[k]); X[k] = Dual(Xv[k], dir[18 + k]); }
  gf::penalty_point<Dual>(U, X, tp, ad, ar, ee, g);
  *e = ee.v; *de = ee.d;
  for (int k = 0; k < 18; ++k) { grad[k] = g[k].v; dgrad[k] = g[k].d; }
}
}
