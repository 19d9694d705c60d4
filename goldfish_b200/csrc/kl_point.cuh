// Point-level Kirchhoff-Love / penalty mathematics of the hot path.
//
// Replaces the FFC/UFLACS-generated `tabulate_tensor` bodies the reference JIT
// compiles from UFL (every `Form(...)`/`assemble(...)` site,
// /root/reference/GOLDFISH/nonmatching_opt.py:442,451,734,780,852,936).  The
// reference differentiates ONE energy symbolically; here the FIRST variation is
// hand-derived (Kiendl et al. 2009) on a generic scalar type S and every second
// derivative (tangent, dR/dCP, dR/dt) is the derivative of that first variation
// along one direction carried by a dual number -- one direction per GPU lane.
//
// Energy (ShNAPr `surfaceEnergyDensitySVK`), written in
// curvilinear (invariant) form, identical to the local-Cartesian form for the
// isotropic D:   e = J [ t/2 eps^T D eps + t^3/24 kap^T D kap ],  J = |X_1 x X_2|
//   eps = (e11, e22, 2 e12),  e_ab = (a_ab - A_ab)/2
//   kap = (k11, k22, 2 k12),  k_ab = B_ab - b_ab,  b_ab = x_,ab . a3
//   D   = C [[A11^2, nu A11 A22 + (1-nu) A12^2, A11 A12],
//            [ .   , A22^2                     , A22 A12],
//            [ .   , .  , ((1-nu) A11 A22 + (1+nu) A12^2)/2 ]]   (contravariant A^ab)
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define GF_HD __host__ __device__ __forceinline__
#else
#define GF_HD inline
#endif

namespace gf {

struct Dual {
  double v, d;
  GF_HD Dual() : v(0.0), d(0.0) {}
  GF_HD Dual(double v_) : v(v_), d(0.0) {}
  GF_HD Dual(double v_, double d_) : v(v_), d(d_) {}
};
GF_HD Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
GF_HD Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
GF_HD Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
GF_HD Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, fma(a.v, b.d, a.d * b.v)); }
GF_HD Dual operator*(double a, Dual b) { return Dual(a * b.v, a * b.d); }
GF_HD Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
GF_HD Dual operator+(Dual a, double b) { return Dual(a.v + b, a.d); }
GF_HD Dual operator+(double a, Dual b) { return Dual(a + b.v, b.d); }
GF_HD Dual operator-(Dual a, double b) { return Dual(a.v - b, a.d); }
GF_HD Dual operator-(double a, Dual b) { return Dual(a - b.v, -b.d); }
GF_HD Dual operator/(Dual a, Dual b) {
  double r = 1.0 / b.v;
  double q = a.v * r;
  return Dual(q, (a.d - q * b.d) * r);
}
GF_HD Dual operator/(double a, Dual b) {
  double r = 1.0 / b.v;
  double q = a * r;
  return Dual(q, -q * b.d * r);
}
GF_HD Dual gf_sqrt(Dual a) {
  double s = sqrt(a.v);
  return Dual(s, 0.5 * a.d / s);
}
GF_HD double gf_sqrt(double a) { return sqrt(a); }
GF_HD double gf_val(double a) { return a; }
GF_HD double gf_val(Dual a) { return a.v; }
GF_HD double gf_dir(double) { return 0.0; }
GF_HD double gf_dir(Dual a) { return a.d; }

template <class S> struct V3 { S x, y, z; };
template <class S> GF_HD V3<S> mk(S x, S y, S z) { V3<S> r; r.x = x; r.y = y; r.z = z; return r; }
template <class S> GF_HD S dot(const V3<S>& a, const V3<S>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class S> GF_HD V3<S> cross(const V3<S>& a, const V3<S>& b) {
  return mk<S>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <class S> GF_HD V3<S> operator+(const V3<S>& a, const V3<S>& b) { return mk<S>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class S> GF_HD V3<S> operator-(const V3<S>& a, const V3<S>& b) { return mk<S>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class S> GF_HD V3<S> scal(S s, const V3<S>& a) { return mk<S>(s * a.x, s * a.y, s * a.z); }
template <class S> GF_HD V3<S> ld3(const S* p) { return mk<S>(p[0], p[1], p[2]); }
template <class S> GF_HD void st3(S* p, const V3<S>& a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

// ---------------------------------------------------------------------------
// Shell point.  gX, gu: [.,1 (3) | .,2 (3) | .,11 (3) | .,22 (3) | .,12 (3)].
// Outputs: e = J*psi (energy per unit parametric area), J, and
// grad[15] = d e / d gu  (first variation w.r.t. the displacement derivatives,
// which equals d e / d gx at fixed reference configuration).
// ---------------------------------------------------------------------------
template <class S> GF_HD V3<S> lift(const V3<S>& a) { return a; }
GF_HD V3<Dual> lift_d(const V3<double>& a) { return mk<Dual>(Dual(a.x), Dual(a.y), Dual(a.z)); }

// Reference-configuration quantities (depend on g_X only).
template <class S> struct KlRef { S J, A11, A22, A12, B11, B22, B12, D11, D12, D13, D22, D23, D33; };

template <class S>
GF_HD void kl_reference(const S* gX, double E, double nu, KlRef<S>& R) {
  const V3<S> X1 = ld3(gX), X2 = ld3(gX + 3), X11 = ld3(gX + 6), X22 = ld3(gX + 9), X12 = ld3(gX + 12);
  const V3<S> Nr = cross(X1, X2);
  const S det = dot(Nr, Nr);
  R.J = gf_sqrt(det);
  const S iJ = 1.0 / R.J;
  const V3<S> A3 = scal(iJ, Nr);
  R.A11 = dot(X1, X1); R.A22 = dot(X2, X2); R.A12 = dot(X1, X2);
  R.B11 = dot(X11, A3); R.B22 = dot(X22, A3); R.B12 = dot(X12, A3);
  const S idet = 1.0 / det;
  const S c11 = R.A22 * idet, c22 = R.A11 * idet, c12 = -(R.A12 * idet);
  const double C = E / (1.0 - nu * nu);
  R.D11 = C * (c11 * c11); R.D22 = C * (c22 * c22);
  R.D12 = C * (nu * (c11 * c22) + (1.0 - nu) * (c12 * c12));
  R.D13 = C * (c11 * c12); R.D23 = C * (c22 * c12);
  R.D33 = (0.5 * C) * ((1.0 - nu) * (c11 * c22) + (1.0 + nu) * (c12 * c12));
}

// Deformed configuration with PLAIN reference data (directions in g_u and t only):
// the tangent and dR/dt passes use this -- the reference part costs no dual arithmetic.
template <class S>
GF_HD void kl_shell_point_fixed_ref(const double* gXd, const KlRef<double>& R, const S* gu, S t,
                                    S& e, S* grad) {
  S gx[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) gx[k] = gXd[k] + gu[k];
  const V3<S> x1 = ld3(gx), x2 = ld3(gx + 3), x11 = ld3(gx + 6), x22 = ld3(gx + 9), x12 = ld3(gx + 12);
  const V3<S> n = cross(x1, x2);
  const S j = gf_sqrt(dot(n, n));
  const S ij = 1.0 / j;
  const V3<S> a3 = scal(ij, n);
  const S a11 = dot(x1, x1), a22 = dot(x2, x2), a12 = dot(x1, x2);
  const S b11 = dot(x11, a3), b22 = dot(x22, a3), b12 = dot(x12, a3);
  const S e0 = 0.5 * (a11 - R.A11), e1 = 0.5 * (a22 - R.A22), e2 = a12 - R.A12;
  const S k0 = R.B11 - b11, k1 = R.B22 - b22, k2 = 2.0 * (R.B12 - b12);
  const S tb = (t * t * t) * (1.0 / 12.0);
  const S n0 = t * (R.D11 * e0 + R.D12 * e1 + R.D13 * e2);
  const S n1 = t * (R.D12 * e0 + R.D22 * e1 + R.D23 * e2);
  const S n2 = t * (R.D13 * e0 + R.D23 * e1 + R.D33 * e2);
  const S m0 = tb * (R.D11 * k0 + R.D12 * k1 + R.D13 * k2);
  const S m1 = tb * (R.D12 * k0 + R.D22 * k1 + R.D23 * k2);
  const S m2 = tb * (R.D13 * k0 + R.D23 * k1 + R.D33 * k2);
  e = (0.5 * R.J) * (e0 * n0 + e1 * n1 + e2 * n2 + k0 * m0 + k1 * m1 + k2 * m2);
  const V3<S> hm = scal(m0, x11) + scal(m1, x22) + scal(2.0 * m2, x12);
  const V3<S> hp = hm - scal(dot(hm, a3), a3);
  const S Jj = R.J * ij;
  const V3<S> g1 = scal(S(R.J), scal(n0, x1) + scal(n2, x2)) - scal(Jj, cross(x2, hp));
  const V3<S> g2 = scal(S(R.J), scal(n1, x2) + scal(n2, x1)) - scal(Jj, cross(hp, x1));
  st3(grad, g1);
  st3(grad + 3, g2);
  st3(grad + 6, scal(-(R.J * m0), a3));
  st3(grad + 9, scal(-(R.J * m1), a3));
  st3(grad + 12, scal(-(2.0 * (R.J * m2)), a3));
}

template <class S>
GF_HD void kl_shell_point(const S* gX, const S* gu, S t, double E, double nu,
                          S& e, S& J, S* grad) {
  const V3<S> X1 = ld3(gX), X2 = ld3(gX + 3), X11 = ld3(gX + 6), X22 = ld3(gX + 9), X12 = ld3(gX + 12);
  const V3<S> x1 = X1 + ld3(gu), x2 = X2 + ld3(gu + 3), x11 = X11 + ld3(gu + 6),
              x22 = X22 + ld3(gu + 9), x12 = X12 + ld3(gu + 12);
  // reference configuration
  const V3<S> Nr = cross(X1, X2);
  const S det = dot(Nr, Nr);
  J = gf_sqrt(det);
  const S iJ = 1.0 / J;
  const V3<S> A3 = scal(iJ, Nr);
  const S A11 = dot(X1, X1), A22 = dot(X2, X2), A12 = dot(X1, X2);
  const S B11 = dot(X11, A3), B22 = dot(X22, A3), B12 = dot(X12, A3);
  const S idet = 1.0 / det;
  const S c11 = A22 * idet, c22 = A11 * idet, c12 = -(A12 * idet);
  const double C = E / (1.0 - nu * nu);
  const S D11 = C * (c11 * c11), D22 = C * (c22 * c22);
  const S D12 = C * (nu * (c11 * c22) + (1.0 - nu) * (c12 * c12));
  const S D13 = C * (c11 * c12), D23 = C * (c22 * c12);
  const S D33 = (0.5 * C) * ((1.0 - nu) * (c11 * c22) + (1.0 + nu) * (c12 * c12));
  // deformed configuration
  const V3<S> n = cross(x1, x2);
  const S j = gf_sqrt(dot(n, n));
  const S ij = 1.0 / j;
  const V3<S> a3 = scal(ij, n);
  const S a11 = dot(x1, x1), a22 = dot(x2, x2), a12 = dot(x1, x2);
  const S b11 = dot(x11, a3), b22 = dot(x22, a3), b12 = dot(x12, a3);
  const S e0 = 0.5 * (a11 - A11), e1 = 0.5 * (a22 - A22), e2 = a12 - A12;
  const S k0 = B11 - b11, k1 = B22 - b22, k2 = 2.0 * (B12 - b12);
  const S tb = (t * t * t) * (1.0 / 12.0);
  const S n0 = t * (D11 * e0 + D12 * e1 + D13 * e2);
  const S n1 = t * (D12 * e0 + D22 * e1 + D23 * e2);
  const S n2 = t * (D13 * e0 + D23 * e1 + D33 * e2);
  const S m0 = tb * (D11 * k0 + D12 * k1 + D13 * k2);
  const S m1 = tb * (D12 * k0 + D22 * k1 + D23 * k2);
  const S m2 = tb * (D13 * k0 + D23 * k1 + D33 * k2);
  e = (0.5 * J) * (e0 * n0 + e1 * n1 + e2 * n2 + k0 * m0 + k1 * m1 + k2 * m2);
  // first variation
  const V3<S> hm = scal(m0, x11) + scal(m1, x22) + scal(2.0 * m2, x12);
  const V3<S> hp = hm - scal(dot(hm, a3), a3);
  const S Jj = J * ij;
  const V3<S> g1 = scal(J, scal(n0, x1) + scal(n2, x2)) - scal(Jj, cross(x2, hp));
  const V3<S> g2 = scal(J, scal(n1, x2) + scal(n2, x1)) - scal(Jj, cross(hp, x1));
  st3(grad, g1);
  st3(grad + 3, g2);
  st3(grad + 6, scal(-(J * m0), a3));
  st3(grad + 9, scal(-(J * m1), a3));
  st3(grad + 12, scal(-(2.0 * (J * m2)), a3));
}

// Reference area measure J = |X_1 x X_2| only (volume functional).
template <class S>
GF_HD S kl_area(const S* gX) {
  const V3<S> Nr = cross(ld3(gX), ld3(gX + 3));
  return gf_sqrt(dot(Nr, Nr));
}

// ---------------------------------------------------------------------------
// Penalty point (PENGoLINS `penalty_energy`), one
// (mortar cell, end vertex) evaluation.
//   uv[18]  = [uA(3) | uA,1 | uA,2 | uB(3) | uB,1 | uB,2]
//   Xv[18]  = [XA(c)(3) | XA(c+1)(3) | XA,1 | XA,2 | XB,1 | XB,2]   (c: cell vertices)
//   tp[2]   = parametric tangent of the cell on side A
// Outputs e and grad[18] = d e / d uv.
// ---------------------------------------------------------------------------
template <class S>
GF_HD void penalty_point(const S* uv, const S* Xv, const double* tp, double alpha_d,
                         double alpha_r, S& e, S* grad) {
  const V3<S> uA = ld3(uv), uA1 = ld3(uv + 3), uA2 = ld3(uv + 6);
  const V3<S> uB = ld3(uv + 9), uB1 = ld3(uv + 12), uB2 = ld3(uv + 15);
  const V3<S> chord = ld3(Xv + 3) - ld3(Xv);
  const S w = 0.5 * gf_sqrt(dot(chord, chord));
  const V3<S> XA1 = ld3(Xv + 6), XA2 = ld3(Xv + 9), XB1 = ld3(Xv + 12), XB2 = ld3(Xv + 15);
  const double t1 = tp[0], t2 = tp[1];
  // reference frames
  V3<S> N = cross(XA1, XA2);
  const V3<S> A3A = scal(1.0 / gf_sqrt(dot(N, N)), N);
  V3<S> T = scal(S(t1), XA1) + scal(S(t2), XA2);
  const V3<S> AtA = scal(1.0 / gf_sqrt(dot(T, T)), T);
  const V3<S> AnA = cross(AtA, A3A);
  N = cross(XB1, XB2);
  const V3<S> A3B = scal(1.0 / gf_sqrt(dot(N, N)), N);
  const S R1 = dot(A3A, A3B), R2 = dot(AnA, A3B);
  // deformed frames
  const V3<S> xA1 = XA1 + uA1, xA2 = XA2 + uA2, xB1 = XB1 + uB1, xB2 = XB2 + uB2;
  const V3<S> nA = cross(xA1, xA2);
  const S inA = 1.0 / gf_sqrt(dot(nA, nA));
  const V3<S> a3A = scal(inA, nA);
  const V3<S> tA = scal(S(t1), xA1) + scal(S(t2), xA2);
  const S itA = 1.0 / gf_sqrt(dot(tA, tA));
  const V3<S> atA = scal(itA, tA);
  const V3<S> anA = cross(atA, a3A);
  const V3<S> nB = cross(xB1, xB2);
  const S inB = 1.0 / gf_sqrt(dot(nB, nB));
  const V3<S> a3B = scal(inB, nB);
  const V3<S> du = uA - uB;
  const S r1 = dot(a3A, a3B) - R1;
  const S r2 = dot(anA, a3B) - R2;
  e = w * ((0.5 * alpha_d) * dot(du, du) + (0.5 * alpha_r) * (r1 * r1 + r2 * r2));
  // ---- first variation (reverse sweep by hand) ----
  const S wd = w * alpha_d, wr = w * alpha_r;
  st3(grad, scal(wd, du));
  st3(grad + 9, scal(-wd, du));
  // adjoints of the unit vectors
  //   e_r = wr/2 (r1^2 + r2^2): bar(a3A) = wr (r1 a3B + r2 (a3B x atA)),
  //   bar(atA) = wr r2 (a3A x a3B),  bar(a3B) = wr (r1 a3A + r2 anA)
  const V3<S> ba3A = scal(wr, scal(r1, a3B) + scal(r2, cross(a3B, atA)));
  const V3<S> batA = scal(wr * r2, cross(a3A, a3B));
  const V3<S> ba3B = scal(wr, scal(r1, a3A) + scal(r2, anA));
  // through normalisation v/|v|: bar(v) = (bar(u) - (bar(u).u) u)/|v|
  const V3<S> bnA = scal(inA, ba3A - scal(dot(ba3A, a3A), a3A));
  const V3<S> btA = scal(itA, batA - scal(dot(batA, atA), atA));
  const V3<S> bnB = scal(inB, ba3B - scal(dot(ba3B, a3B), a3B));
  // n = x1 x x2: bar(x1) = x2 x bar(n), bar(x2) = bar(n) x x1 ; t = t1 x1 + t2 x2
  st3(grad + 3, cross(xA2, bnA) + scal(S(t1), btA));
  st3(grad + 6, cross(bnA, xA1) + scal(S(t2), btA));
  st3(grad + 12, cross(xB2, bnB));
  st3(grad + 15, cross(bnB, xB1));
}

}  // namespace gf
