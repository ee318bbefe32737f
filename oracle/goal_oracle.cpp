// goal_oracle.cpp -- CPU oracle: a plain restatement of bgranzow/goal's assembly hot path.
//
// TEST INFRASTRUCTURE ONLY (see goal_oracle.h).  Serial C++17, no dependencies.
// Every block cites the reference file:line it restates (paths under
// /root/reference/src/).  The third-party arithmetic the reference leans on
// (Sacado SLFad<double,16>, MiniTensor 3x3 helpers, apf linear-tet shape
// functions, Tpetra sorted-row CRS) is restated from its published semantics
// because none of those sources exist in this environment;
// tests/test_oracle_goldens.py pins the result against the reference's golden
// functional values.
//
// Structure: one templated `chain<T>()` evaluates the evaluator vector that
// Mechanics::build_resid<T> (goal_mechanics.cpp:97-146) / build_error (:169-218)
// assembles, for one element, in the reference's order, with T = double (ST) or
// Fad (FADT).  The drivers below loop elements like assemble()
// (goal_assembly.cpp:65-88) and scatter like Displacement/Pressure::scatter_*.

#include "goal_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

namespace {

// ---------------------------------------------------------------------------
// Forward-mode AD scalar with Sacado::Fad::SLFad<double,16> semantics
// (goal_scalar_types.hpp:9, CMakeLists.txt:13): a value plus up to 16
// derivative components; a constant has a zero-length derivative array, and
// comparisons look at the value only.
// ---------------------------------------------------------------------------
constexpr int ND = 16;

struct Fad {
  double v;
  int n;  // 0 (constant) or ND
  double d[ND];
  Fad() : v(0.0), n(0) {}
  Fad(double x) : v(x), n(0) {}
  void diff(int i, int sz) {  // Sacado: resize, zero, seed
    n = sz;
    for (int k = 0; k < ND; ++k) d[k] = 0.0;
    d[i] = 1.0;
  }
  double dx(int i) const { return n ? d[i] : 0.0; }
};

inline double value(double x) { return x; }
inline double value(Fad const& x) { return x.v; }

inline Fad operator-(Fad const& a) {
  Fad r; r.v = -a.v; r.n = a.n;
  if (a.n) for (int k = 0; k < ND; ++k) r.d[k] = -a.d[k];
  return r;
}
inline Fad operator+(Fad const& a, Fad const& b) {
  Fad r; r.v = a.v + b.v; r.n = a.n | b.n;
  if (a.n && b.n) for (int k = 0; k < ND; ++k) r.d[k] = a.d[k] + b.d[k];
  else if (a.n) for (int k = 0; k < ND; ++k) r.d[k] = a.d[k];
  else if (b.n) for (int k = 0; k < ND; ++k) r.d[k] = b.d[k];
  return r;
}
inline Fad operator-(Fad const& a, Fad const& b) {
  Fad r; r.v = a.v - b.v; r.n = a.n | b.n;
  if (a.n && b.n) for (int k = 0; k < ND; ++k) r.d[k] = a.d[k] - b.d[k];
  else if (a.n) for (int k = 0; k < ND; ++k) r.d[k] = a.d[k];
  else if (b.n) for (int k = 0; k < ND; ++k) r.d[k] = -b.d[k];
  return r;
}
inline Fad operator*(Fad const& a, Fad const& b) {
  Fad r; r.v = a.v * b.v; r.n = a.n | b.n;
  if (a.n && b.n) for (int k = 0; k < ND; ++k) r.d[k] = a.v * b.d[k] + a.d[k] * b.v;
  else if (a.n) for (int k = 0; k < ND; ++k) r.d[k] = a.d[k] * b.v;
  else if (b.n) for (int k = 0; k < ND; ++k) r.d[k] = a.v * b.d[k];
  return r;
}
inline Fad operator/(Fad const& a, Fad const& b) {
  Fad r; r.v = a.v / b.v; r.n = a.n | b.n;
  if (a.n && b.n) {
    double const b2 = b.v * b.v;
    for (int k = 0; k < ND; ++k) r.d[k] = (a.d[k] * b.v - a.v * b.d[k]) / b2;
  } else if (a.n) {
    for (int k = 0; k < ND; ++k) r.d[k] = a.d[k] / b.v;
  } else if (b.n) {
    double const c = -a.v / (b.v * b.v);
    for (int k = 0; k < ND; ++k) r.d[k] = c * b.d[k];
  }
  return r;
}
inline Fad& operator+=(Fad& a, Fad const& b) { a = a + b; return a; }
inline Fad& operator-=(Fad& a, Fad const& b) { a = a - b; return a; }
inline Fad& operator*=(Fad& a, Fad const& b) { a = a * b; return a; }
inline Fad& operator/=(Fad& a, Fad const& b) { a = a / b; return a; }
inline bool operator>(Fad const& a, double b) { return a.v > b; }
inline bool operator<(Fad const& a, double b) { return a.v < b; }

inline Fad fn1(Fad const& a, double val, double dval) {  // chain rule helper
  Fad r; r.v = val; r.n = a.n;
  if (a.n) for (int k = 0; k < ND; ++k) r.d[k] = dval * a.d[k];
  return r;
}
inline Fad sqrt(Fad const& a) { double s = std::sqrt(a.v); return fn1(a, s, 0.5 / s); }
inline Fad cbrt(Fad const& a) { double c = std::cbrt(a.v); return fn1(a, c, c / (3.0 * a.v)); }
inline Fad pow(Fad const& a, double e) {
  double p = std::pow(a.v, e);
  return fn1(a, p, a.v == 0.0 ? 0.0 : e * p / a.v);
}
inline Fad abs(Fad const& a) { return a.v >= 0.0 ? a : -a; }
using std::abs;
using std::cbrt;
using std::pow;
using std::sqrt;

// ---------------------------------------------------------------------------
// 3x3 tensor helpers with MiniTensor semantics (used by goal_kinematics.cpp:24,
// goal_neohookean.cpp:36,67,69,83, goal_J2.cpp:84,90,93,130,154,
// goal_stabilization.cpp:76): closed-form det / inverse, dev(A) = A - tr(A)/3 I,
// Frobenius norm, Pade scaling-and-squaring exp.
// ---------------------------------------------------------------------------
template <class T>
struct Ten {
  T a[3][3];
  T& operator()(int i, int j) { return a[i][j]; }
  T const& operator()(int i, int j) const { return a[i][j]; }
};

template <class T> Ten<T> zeros() { Ten<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = T(0.0); return r; }
template <class T> Ten<T> eye() { Ten<T> r = zeros<T>(); for (int i = 0; i < 3; ++i) r(i, i) = T(1.0); return r; }
template <class T> Ten<T> transpose(Ten<T> const& A) { Ten<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = A(j, i); return r; }
template <class T> Ten<T> operator*(Ten<T> const& A, Ten<T> const& B) {
  Ten<T> r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    T s = A(i, 0) * B(0, j);
    for (int k = 1; k < 3; ++k) s += A(i, k) * B(k, j);
    r(i, j) = s;
  }
  return r;
}
template <class T, class S> Ten<T> scale(S const& s, Ten<T> const& A) { Ten<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = s * A(i, j); return r; }
template <class T> Ten<T> operator+(Ten<T> const& A, Ten<T> const& B) { Ten<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = A(i, j) + B(i, j); return r; }
template <class T> Ten<T> operator-(Ten<T> const& A, Ten<T> const& B) { Ten<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = A(i, j) - B(i, j); return r; }
template <class T> T trace(Ten<T> const& A) { return A(0, 0) + A(1, 1) + A(2, 2); }
template <class T> Ten<T> dev(Ten<T> const& A) {
  Ten<T> r = A;
  T th = trace(A) / 3.0;
  for (int i = 0; i < 3; ++i) r(i, i) -= th;
  return r;
}
template <class T> T norm(Ten<T> const& A) {
  T s = T(0.0);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s += A(i, j) * A(i, j);
  return sqrt(s);
}
template <class T> T det(Ten<T> const& A) {
  return -A(0, 2) * A(1, 1) * A(2, 0) + A(0, 1) * A(1, 2) * A(2, 0) + A(0, 2) * A(1, 0) * A(2, 1)
         - A(0, 0) * A(1, 2) * A(2, 1) - A(0, 1) * A(1, 0) * A(2, 2) + A(0, 0) * A(1, 1) * A(2, 2);
}
template <class T> Ten<T> inverse(Ten<T> const& A) {
  T d = det(A);
  Ten<T> r;
  r(0, 0) = (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) / d;
  r(0, 1) = (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2)) / d;
  r(0, 2) = (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1)) / d;
  r(1, 0) = (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2)) / d;
  r(1, 1) = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) / d;
  r(1, 2) = (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2)) / d;
  r(2, 0) = (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0)) / d;
  r(2, 1) = (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1)) / d;
  r(2, 2) = (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0)) / d;
  return r;
}
template <class T> double norm_1_value(Ten<T> const& A) {
  double m = 0.0;
  for (int j = 0; j < 3; ++j) {
    double s = 0.0;
    for (int i = 0; i < 3; ++i) s += std::fabs(value(A(i, j)));
    m = std::max(m, s);
  }
  return m;
}
// Matrix exponential: [m/m] Pade approximant with scaling and squaring
// (Higham 2005 thresholds), which is what minitensor::exp does.
template <class T> Ten<T> exp_pade(Ten<T> const& A) {
  static const double theta[5] = {1.495585217958292e-2, 2.539398330063230e-1, 9.504178996162932e-1,
                                  2.097847961257068e0, 5.371920351148152e0};
  static const int orders[5] = {3, 5, 7, 9, 13};
  static const double b3[] = {120., 60., 12., 1.};
  static const double b5[] = {30240., 15120., 3360., 420., 30., 1.};
  static const double b7[] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
  static const double b9[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240.,
                              2162160., 110880., 3960., 90., 1.};
  static const double b13[] = {64764752532480000., 32382376266240000., 7771770303897600.,
                               1187353796428800., 129060195264000., 10559470521600., 670442572800.,
                               33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.};
  static const double* const coef[5] = {b3, b5, b7, b9, b13};
  double const nrm = norm_1_value(A);
  int which = 4, squarings = 0;
  for (int k = 0; k < 5; ++k) if (nrm <= theta[k]) { which = k; break; }
  Ten<T> As = A;
  if (nrm > theta[4]) {
    squarings = std::max(0, (int)std::ceil(std::log2(nrm / theta[4])));
    As = scale(std::ldexp(1.0, -squarings), A);
  }
  int const m = orders[which];
  double const* b = coef[which];
  Ten<T> const I = eye<T>();
  Ten<T> const A2 = As * As;
  Ten<T> U = scale(b[1], I), V = scale(b[0], I), P = I;  // P = A2^k
  for (int k = 1; 2 * k <= m; ++k) {
    P = P * A2;
    U = U + scale(b[2 * k + 1], P);
    V = V + scale(b[2 * k], P);
  }
  U = As * U;
  Ten<T> R = inverse(V - U) * (V + U);
  for (int s = 0; s < squarings; ++s) R = R * R;
  return R;
}

// ---------------------------------------------------------------------------
// apf linear tetrahedron (restated from knowledge):
//   N = (1-xi0-xi1-xi2, xi0, xi1, xi2); Jacobian rows x1-x0, x2-x0, x3-x0;
//   grad N_n = J^{-1} dN_n/dxi; getDV = det J; order-1 rule: xi=(1/4,1/4,1/4), w=1/6
// (goal_assembly.cpp:78-80).
// ---------------------------------------------------------------------------
struct TetGeom {
  double BF[4];
  double GBF[4][3];
  double dv;  // det J
  double w;   // 1/6
  double h;   // goal_stabilization.cpp:59-66
};

bool tet_geometry(double const x[4][3], TetGeom& g) {
  Ten<double> J;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J(i, j) = x[i + 1][j] - x[0][j];
  g.dv = det(J);
  g.w = 1.0 / 6.0;
  Ten<double> Ji = inverse(J);
  static const double dN[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int n = 0; n < 4; ++n)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += Ji(j, k) * dN[n][k];
      g.GBF[n][j] = s;
    }
  double const xi = 0.25;
  g.BF[0] = 1.0 - xi - xi - xi;
  g.BF[1] = g.BF[2] = g.BF[3] = xi;
  // get_size: sqrt(sum over the 6 edges of length^2 / 6)
  static const int ev[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
  double h2 = 0.0;
  for (int e = 0; e < 6; ++e) {
    double l2 = 0.0;
    for (int j = 0; j < 3; ++j) { double d = x[ev[e][1]][j] - x[ev[e][0]][j]; l2 += d * d; }
    double const l = std::sqrt(l2);  // apf::measure(edge) is a length; reference squares it again
    h2 += l * l;
  }
  g.h = std::sqrt(h2 / 6);
  return g.dv > 0.0;
}

}  // namespace

// ---------------------------------------------------------------------------
struct go_ctx {
  int nn = 0, ne = 0, nsets = 1, model = 0;
  std::vector<int32_t> conn, eset;
  std::vector<double> coords, mats;
  std::vector<double> u, p;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> colind;
  // states, AoS (goal_mechanics.cpp:87-95)
  std::vector<double> sigma, eqps, eqps_old, Fp, Fp_old;
  int64_t plastic = 0;
  std::string err;
};

namespace {

struct Material {
  double E, nu, K, Y, c0, kappa, mu;
};
Material material(go_ctx const& c, int es) {
  Material m;
  m.E = c.mats[es * 5 + 0]; m.nu = c.mats[es * 5 + 1]; m.K = c.mats[es * 5 + 2];
  m.Y = c.mats[es * 5 + 3]; m.c0 = c.mats[es * 5 + 4];
  m.kappa = m.E / (3.0 * (1.0 - 2.0 * m.nu));  // goal_neohookean.cpp:50, goal_J2.cpp:62
  m.mu = m.E / (2.0 * (1.0 + m.nu));           // goal_neohookean.cpp:51, goal_J2.cpp:63
  return m;
}

// Test functions.  Plain weights: goal_vector_weight.cpp:13-28, goal_scalar_weight.cpp:16-31.
// Adjoint weights: goal_displacement_adjoint.cpp:37-53, goal_pressure_adjoint.cpp:38-49.
struct Weights {
  double uw_grad[4][3][3];  // (n,i,j)
  double pw_val[4];
  double pw_grad[4][3];   // "pw"  (PResidual)
  double pwc_grad[4][3];  // "pwc" (Stabilization in build_error), == pw_grad in build_resid
};

void plain_weights(TetGeom const& g, Weights& w) {
  for (int n = 0; n < 4; ++n) {
    w.pw_val[n] = g.BF[n];
    for (int j = 0; j < 3; ++j) {
      w.pw_grad[n][j] = w.pwc_grad[n][j] = g.GBF[n][j];
      for (int i = 0; i < 3; ++i) w.uw_grad[n][i][j] = g.GBF[n][j];
    }
  }
}

void adjoint_weights(TetGeom const& g, double const zu[4][3], double const zp[4], double const zpc[4],
                     Weights& w) {
  double z[3], gz[3][3];  // gz[i][j] = d z_i / d x_j (the reference transposes apf's getVectorGrad)
  for (int i = 0; i < 3; ++i) {
    z[i] = 0.0;
    for (int n = 0; n < 4; ++n) z[i] += zu[n][i] * g.BF[n];
    for (int j = 0; j < 3; ++j) {
      gz[i][j] = 0.0;
      for (int n = 0; n < 4; ++n) gz[i][j] += zu[n][i] * g.GBF[n][j];
    }
  }
  double s = 0.0, sc = 0.0, gs[3] = {0, 0, 0}, gsc[3] = {0, 0, 0};
  for (int n = 0; n < 4; ++n) {
    s += zp[n] * g.BF[n];
    sc += zpc[n] * g.BF[n];
    for (int j = 0; j < 3; ++j) { gs[j] += zp[n] * g.GBF[n][j]; gsc[j] += zpc[n] * g.GBF[n][j]; }
  }
  for (int n = 0; n < 4; ++n) {
    w.pw_val[n] = s * g.BF[n];
    for (int j = 0; j < 3; ++j) {
      w.pw_grad[n][j] = gs[j] * g.BF[n] + s * g.GBF[n][j];
      w.pwc_grad[n][j] = gsc[j] * g.BF[n] + sc * g.GBF[n][j];
      for (int i = 0; i < 3; ++i) w.uw_grad[n][i][j] = gz[i][j] * g.BF[n] + z[i] * g.GBF[n][j];
    }
  }
}

template <class T> void store_tensor(double* dst, Ten<T> const& A) {  // set_tensor, goal_states.cpp:47-57
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) dst[3 * i + j] = value(A(i, j));
}

// One element through the evaluator chain, reference order:
//   u, uw, p, pw, kinematics, model, mixed, mresidual, presidual, stabilization.
// ru[n*3+i], rp[n] are the element residuals the u / p evaluators own.
template <class T>
bool chain(go_ctx& c, int e, TetGeom const& g, Weights const& w, bool save, T ru[12], T rp[4],
           T* uval /*[3], optional*/, Ten<T>* cauchy = nullptr /* model->get_cauchy() after Mixed, optional */) {
  Material const m = material(c, c.eset.empty() ? 0 : c.eset[e]);
  int32_t const* nd = &c.conn[4 * (size_t)e];

  // --- Displacement<T>::gather / Pressure<T>::gather: nodal values, FAD seeds
  //     dx[n*4+d] (goal_displacement.cpp:139-156, goal_pressure.cpp:136-151)
  T un[4][3], pn[4];
  for (int n = 0; n < 4; ++n) {
    for (int d = 0; d < 3; ++d) {
      un[n][d] = T(c.u[3 * (size_t)nd[n] + d]);
      if constexpr (std::is_same<T, Fad>::value) un[n][d].diff(n * 4 + d, ND);
      ru[n * 3 + d] = T(0.0);
    }
    pn[n] = T(c.p[nd[n]]);
    if constexpr (std::is_same<T, Fad>::value) pn[n].diff(n * 4 + 3, ND);
    rp[n] = T(0.0);
  }
  // --- Displacement<T>::at_point (goal_displacement.cpp:158-172)
  T uv[3];
  Ten<T> gradu;
  for (int i = 0; i < 3; ++i) {
    uv[i] = un[0][i] * g.BF[0];
    for (int n = 1; n < 4; ++n) uv[i] += un[n][i] * g.BF[n];
    for (int j = 0; j < 3; ++j) {
      gradu(i, j) = un[0][i] * g.GBF[0][j];
      for (int n = 1; n < 4; ++n) gradu(i, j) += un[n][i] * g.GBF[n][j];
    }
  }
  if (uval) for (int i = 0; i < 3; ++i) uval[i] = uv[i];
  // --- Pressure<T>::at_point (goal_pressure.cpp:153-165)
  T pv = pn[0] * g.BF[0];
  for (int n = 1; n < 4; ++n) pv += pn[n] * g.BF[n];
  T gradp[3];
  for (int i = 0; i < 3; ++i) {
    gradp[i] = pn[0] * g.GBF[0][i];
    for (int n = 1; n < 4; ++n) gradp[i] += pn[n] * g.GBF[n][i];
  }
  // --- Kinematics<T>::at_point (goal_kinematics.cpp:18-25)
  Ten<T> F = gradu;
  for (int i = 0; i < 3; ++i) F(i, i) += 1.0;
  T J = det(F);
  if (!(value(J) > 0.0)) { c.err = "inverted deformation (det F <= 0) in element " + std::to_string(e); return false; }

  Ten<T> const I = eye<T>();
  Ten<T> sigma;
  double* sig_dst = &c.sigma[9 * (size_t)e];
  if (c.model == GO_MODEL_NEOHOOKEAN) {
    // --- Neohookean<T>::at_point (goal_neohookean.cpp:60-72)
    T Jm13 = 1.0 / cbrt(J);
    T Jm23 = Jm13 * Jm13;
    T Jm53 = Jm23 * Jm23 * Jm13;
    Ten<T> b = F * transpose(F);
    T pr = 0.5 * m.kappa * (J - 1.0 / J);
    sigma = scale(m.mu * Jm53, dev(b)) + scale(pr, I);
    if (save) store_tensor(sig_dst, sigma);
  } else {
    // --- J2<T>::at_point (goal_J2.cpp:72-143)
    double const sq23 = std::sqrt(2.0 / 3.0);
    T Jm23 = pow(J, -2.0 / 3.0);
    Ten<T> Fp;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Fp(i, j) = T(c.Fp_old[9 * (size_t)e + 3 * i + j]);
    Ten<T> Fpinv = inverse(Fp);
    Ten<T> Cpinv = Fpinv * transpose(Fpinv);
    Ten<T> be = scale(Jm23, F * Cpinv * transpose(F));
    Ten<T> s = scale(m.mu, dev(be));
    T mubar = trace(be) * m.mu / 3;
    T smag = norm(s);
    T eqps = T(c.eqps_old[e]);
    T f = smag - sq23 * (m.Y + m.K * eqps);
    if (f > 1.0e-12) {  // plastic increment, radial return (goal_J2.cpp:98-132)
      int iter = 0;
      bool converged = false;
      T H(0.0), dH(0.0), alpha(0.0), res(0.0);
      T X = 0.0;
      T R = f;
      T dRdX = -2.0 * mubar * (1.0 + H / (3.0 * mubar));
      while ((!converged) && (iter < 30)) {
        iter++;
        X = X - R / dRdX;
        alpha = eqps + sq23 * X;
        H = m.K * alpha;
        dH = m.K;
        R = smag - (2.0 * mubar * X + sq23 * (m.Y + H));
        dRdX = -2.0 * mubar * (1.0 + dH / (3.0 * mubar));
        res = abs(R);
        if ((res < 1.0e-11) || (res / m.Y < 1.0e-11) || (res / f < 1.0e-11)) converged = true;
        if (iter == 30) { c.err = "J2: return mapping failed in element " + std::to_string(e); return false; }
      }
      T dgam = X;
      Ten<T> N = scale(1.0 / smag, s);
      s = s - scale(2.0 * mubar * dgam, N);
      if (save) c.eqps[e] = value(alpha);
      Ten<T> Fpn = exp_pade(scale(dgam, N)) * Fp;
      if (save) store_tensor(&c.Fp[9 * (size_t)e], Fpn);
      c.plastic++;
    } else {
      if (save) c.eqps[e] = value(eqps);  // elastic: Fp is NOT written (goal_J2.cpp:135-136)
    }
    T pr = 0.5 * m.kappa * (J - 1.0 / J);
    sigma = scale(1.0 / J, s) + scale(pr, I);  // sigma = s/J + p*I
    if (save) store_tensor(sig_dst, sigma);
  }
  // --- Mixed<T>::at_point (goal_mixed.cpp:34-46)
  {
    T pbar = T(0.0);
    for (int i = 0; i < 3; ++i) pbar += sigma(i, i);
    pbar /= 3;
    for (int i = 0; i < 3; ++i) sigma(i, i) += pv - pbar;
    if (save) store_tensor(sig_dst, sigma);
    if (cauchy) *cauchy = sigma;
  }
  // --- MResidual<T>::at_point (goal_mresidual.cpp:26-32) with
  //     Model::get_first_pk (goal_neohookean.cpp:80-86, goal_J2.cpp:151-157)
  {
    Ten<T> Finv = inverse(F);
    Ten<T> P = scale(J, sigma * transpose(Finv));
    for (int n = 0; n < 4; ++n)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) ru[n * 3 + i] += P(i, j) * w.uw_grad[n][i][j] * g.w * g.dv;
  }
  // --- PResidual<T>::do_large_strain (goal_presidual.cpp:54-59)
  {
    T dUdJ = 0.5 * (J - 1.0 / J);
    for (int n = 0; n < 4; ++n) rp[n] += ((pv / m.kappa) - dUdJ) * w.pw_val[n] * g.w * g.dv;
  }
  // --- Stabilization<T>::at_point (goal_stabilization.cpp:69-81); weight is "pw"
  //     in build_resid and "pwc" in build_error (goal_mechanics.cpp:142, 214)
  {
    double const tau = 0.5 * m.c0 * g.h * g.h / m.mu;
    Ten<T> Cinv = inverse(transpose(F) * F);
    for (int n = 0; n < 4; ++n)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) rp[n] += tau * J * Cinv(i, j) * gradp[i] * w.pwc_grad[n][j] * g.w * g.dv;
  }
  return true;
}

bool elem_geometry(go_ctx& c, int e, TetGeom& g) {
  double x[4][3];
  for (int n = 0; n < 4; ++n)
    for (int j = 0; j < 3; ++j) x[n][j] = c.coords[3 * (size_t)c.conn[4 * (size_t)e + n] + j];
  if (!tet_geometry(x, g)) { c.err = "inverted element (dv <= 0): " + std::to_string(e); return false; }
  return true;
}

// Position of (row, col) in the sorted CRS row, what Tpetra's sumIntoLocalValues
// resolves internally (goal_displacement.cpp:191).
inline int64_t crs_pos(go_ctx const& c, int row, int col) {
  int32_t const* b = &c.colind[c.rowptr[row]];
  int32_t const* en = &c.colind[c.rowptr[row + 1]];
  int32_t const* it = std::lower_bound(b, en, col);
  return (it != en && *it == col) ? (it - &c.colind[0]) : -1;
}

// Disc::compute_graphs restated (goal_disc.cpp:307-332): every (row dof, col dof)
// pair of every element; Tpetra's fillComplete leaves each row sorted and unique.
void build_graph(go_ctx& c) {
  std::vector<std::vector<int32_t>> nbr(c.nn);
  for (int e = 0; e < c.ne; ++e)
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) nbr[c.conn[4 * (size_t)e + a]].push_back(c.conn[4 * (size_t)e + b]);
  c.rowptr.assign(4 * (size_t)c.nn + 1, 0);
  for (int n = 0; n < c.nn; ++n) {
    std::sort(nbr[n].begin(), nbr[n].end());
    nbr[n].erase(std::unique(nbr[n].begin(), nbr[n].end()), nbr[n].end());
    for (int eq = 0; eq < 4; ++eq) c.rowptr[4 * (size_t)n + eq + 1] = 4 * (int64_t)nbr[n].size();
  }
  for (size_t r = 0; r < 4 * (size_t)c.nn; ++r) c.rowptr[r + 1] += c.rowptr[r];
  c.colind.resize(c.rowptr.back());
  for (int n = 0; n < c.nn; ++n)
    for (int eq = 0; eq < 4; ++eq) {
      int32_t* dst = &c.colind[c.rowptr[4 * (size_t)n + eq]];
      for (int32_t m : nbr[n]) for (int k = 0; k < 4; ++k) *dst++ = 4 * m + k;
    }
}

}  // namespace

extern "C" {

go_ctx* go_create(int n_nodes, int n_elems, const int32_t* conn, const double* coords,
                  const int32_t* elem_set, int n_sets, int model, const double* materials) {
  go_ctx* c = new go_ctx;
  c->nn = n_nodes; c->ne = n_elems; c->nsets = n_sets; c->model = model;
  c->conn.assign(conn, conn + 4 * (size_t)n_elems);
  c->coords.assign(coords, coords + 3 * (size_t)n_nodes);
  if (elem_set) c->eset.assign(elem_set, elem_set + n_elems);
  c->mats.assign(materials, materials + 5 * (size_t)n_sets);
  c->u.assign(3 * (size_t)n_nodes, 0.0);  // Mechanics::make_displacement zeroes (goal_mechanics.cpp:71-77)
  c->p.assign(n_nodes, 0.0);
  // Mechanics::make_states (goal_mechanics.cpp:87-95) + States::add identity init (goal_states.cpp:87-128)
  c->sigma.assign(9 * (size_t)n_elems, 0.0);
  if (model == GO_MODEL_J2) {
    c->eqps.assign(n_elems, 0.0);
    c->eqps_old.assign(n_elems, 0.0);
    c->Fp.assign(9 * (size_t)n_elems, 0.0);
    for (int e = 0; e < n_elems; ++e) for (int i = 0; i < 3; ++i) c->Fp[9 * (size_t)e + 4 * i] = 1.0;
    c->Fp_old = c->Fp;
  }
  build_graph(*c);
  return c;
}
void go_destroy(go_ctx* c) { delete c; }

int64_t go_graph_nnz(go_ctx* c) { return c->rowptr.back(); }
const int64_t* go_graph_rowptr(go_ctx* c) { return c->rowptr.data(); }
const int32_t* go_graph_colind(go_ctx* c) { return c->colind.data(); }

void go_set_solution(go_ctx* c, const double* u, const double* p) {
  c->u.assign(u, u + 3 * (size_t)c->nn);
  c->p.assign(p, p + c->nn);
}

double* go_state(go_ctx* c, const char* name) {
  std::string n(name);
  if (n == "sigma") return c->sigma.data();
  if (n == "eqps") return c->eqps.empty() ? nullptr : c->eqps.data();
  if (n == "eqps_old") return c->eqps_old.empty() ? nullptr : c->eqps_old.data();
  if (n == "Fp") return c->Fp.empty() ? nullptr : c->Fp.data();
  if (n == "Fp_old") return c->Fp_old.empty() ? nullptr : c->Fp_old.data();
  return nullptr;
}
void go_update_states(go_ctx* c) {
  c->eqps_old = c->eqps;
  c->Fp_old = c->Fp;
}

int go_assemble_residual(go_ctx* c, int save_state, double* R) {
  c->plastic = 0; c->err.clear();
  for (int e = 0; e < c->ne; ++e) {
    TetGeom g; Weights w;
    if (!elem_geometry(*c, e, g)) return 1;
    plain_weights(g, w);
    double ru[12], rp[4];
    if (!chain<double>(*c, e, g, w, save_state != 0, ru, rp, nullptr)) return 2;
    // Displacement<ST>/Pressure<ST>::scatter_primal (goal_displacement.cpp:77-86, goal_pressure.cpp:76-83)
    for (int n = 0; n < 4; ++n) {
      int const nd = c->conn[4 * (size_t)e + n];
      for (int d = 0; d < 3; ++d) R[4 * (size_t)nd + d] += ru[n * 3 + d];
      R[4 * (size_t)nd + 3] += rp[n];
    }
  }
  return 0;
}

int go_assemble_jacobian(go_ctx* c, int mode, int save_state, double* R, double* values) {
  c->plastic = 0; c->err.clear();
  for (int e = 0; e < c->ne; ++e) {
    TetGeom g; Weights w;
    if (!elem_geometry(*c, e, g)) return 1;
    plain_weights(g, w);
    Fad ru[12], rp[4];
    if (!chain<Fad>(*c, e, g, w, save_state != 0, ru, rp, nullptr)) return 2;
    int cols[16];  // Disc::get_lids (goal_disc.cpp:214-222)
    for (int n = 0; n < 4; ++n) for (int eq = 0; eq < 4; ++eq) cols[4 * n + eq] = 4 * c->conn[4 * (size_t)e + n] + eq;
    for (int n = 0; n < 4; ++n)
      for (int eq = 0; eq < 4; ++eq) {
        Fad const& v = (eq < 3) ? ru[n * 3 + eq] : rp[n];
        int const row = cols[4 * n + eq];
        R[row] += v.v;
        if (!values || mode == GO_MODE_NONE) continue;
        for (int j = 0; j < 16; ++j) {
          // scatter_primal: A(row, cols[j]) += dx[j]; scatter_adjoint: A(cols[j], row) += dx[j]
          // (goal_displacement.cpp:177-214, goal_pressure.cpp:170-203)
          int64_t const pos = (mode == GO_MODE_PRIMAL) ? crs_pos(*c, row, cols[j]) : crs_pos(*c, cols[j], row);
          if (pos < 0) { c->err = "CRS entry missing"; return 3; }
          values[pos] += v.dx(j);
        }
      }
  }
  return 0;
}

double go_functional_avg_disp(go_ctx* c, double* dMdu) {
  // AvgDisp<T>::at_point (goal_avg_disp.cpp:17-21) inside the save=false chain of
  // Functional (goal_functional.cpp:21-46); QoI<FADT>::scatter (goal_qoi.cpp:63-76).
  double J = 0.0;
  for (int e = 0; e < c->ne; ++e) {
    TetGeom g;
    if (!elem_geometry(*c, e, g)) return NAN;
    int32_t const* nd = &c->conn[4 * (size_t)e];
    double ev = 0.0;
    for (int i = 0; i < 3; ++i) {
      double ui = c->u[3 * (size_t)nd[0] + i] * g.BF[0];
      for (int n = 1; n < 4; ++n) ui += c->u[3 * (size_t)nd[n] + i] * g.BF[n];
      ev += ui * g.w * g.dv;
    }
    ev /= 3;
    J += ev;
    if (dMdu)
      for (int n = 0; n < 4; ++n)
        for (int i = 0; i < 3; ++i) dMdu[4 * (size_t)nd[n] + i] += g.BF[n] * g.w * g.dv / 3;
  }
  return J;
}

}  // extern "C"

namespace {
// compute_von_mises (goal_von_mises.cpp:6-18)
template <class T> T von_mises(Ten<T> const& sigma) {
  T s1 = (sigma(0, 0) - sigma(1, 1)) * (sigma(0, 0) - sigma(1, 1));
  T s2 = (sigma(1, 1) - sigma(2, 2)) * (sigma(1, 1) - sigma(2, 2));
  T s3 = (sigma(2, 2) - sigma(0, 0)) * (sigma(2, 2) - sigma(0, 0));
  T s4 = sigma(0, 1) * sigma(0, 1);
  T s5 = sigma(1, 2) * sigma(1, 2);
  T s6 = sigma(2, 0) * sigma(2, 0);
  T s7 = 0.5 * (s1 + s2 + s3 + 6.0 * (s4 + s5 + s6));
  return sqrt(s7);
}
inline Fad exp(Fad const& a) { double e = std::exp(a.v); return fn1(a, e, e); }
using std::exp;

// One element of the functional chain: build_resid<T>(save = false) followed by the QoI evaluator
// (goal_functional.cpp:41-45, goal_mechanics.cpp:149-167).  ks = {max, scale, rho} for "max vm".
template <class T>
bool qoi_element(go_ctx& c, int e, int type, int es_idx, double const ks[3], T& elem_value) {
  TetGeom g; Weights w;
  if (!elem_geometry(c, e, g)) return false;
  plain_weights(g, w);
  T ru[12], rp[4], uv[3];
  Ten<T> sigma;
  if (!chain<T>(c, e, g, w, false, ru, rp, uv, &sigma)) return false;
  // QoI<T>::gather: elem_value = 0 (FADT: diff(0, num_dofs) with dx(0) = 0, goal_qoi.cpp:54-60)
  elem_value = T(0.0);
  if constexpr (std::is_same<T, Fad>::value) { elem_value.diff(0, ND); elem_value.d[0] = 0.0; }
  int const es = c.eset.empty() ? 0 : c.eset[e];
  switch (type) {
    case GO_QOI_AVG_DISP:  // goal_avg_disp.cpp:17-21
      for (int i = 0; i < 3; ++i) elem_value += uv[i] * g.w * g.dv;
      elem_value /= 3;
      break;
    case GO_QOI_AVG_DISP_SUBDOMAIN:  // goal_avg_disp_subdomain.cpp:37-53
      if (es == es_idx) {
        for (int i = 0; i < 3; ++i) elem_value += uv[i] * g.w * g.dv;
        elem_value /= 3;
      }
      break;
    case GO_QOI_AVG_VM:  // goal_avg_vm.cpp:43-61
      if (es == es_idx) {
        T vm = von_mises(sigma);
        elem_value += vm * g.w * g.dv;
      }
      break;
    case GO_QOI_KS_VM: {  // goal_ks_vm.cpp:89-99
      T vm = von_mises(sigma);
      elem_value += (1.0 / (ks[2] * ks[1])) * exp(ks[2] * (vm - ks[0])) * g.w * g.dv;
      break;
    }
    default: return false;
  }
  return true;
}
}  // namespace

extern "C" double go_functional(go_ctx* c, int type, int elem_set, double rho, int point_node, int point_idx,
                                double* dMdu) {
  c->err.clear();
  c->plastic = 0;
  if (type == GO_QOI_POINT_WISE) {  // PointWise<T>::post_process (goal_point_wise.cpp:37-56)
    if (dMdu) dMdu[4 * (size_t)point_node + point_idx] = 1.0;
    return c->u[3 * (size_t)point_node + point_idx];
  }
  double ks[3] = {0.0, 0.0, rho};
  if (type == GO_QOI_KS_VM) {
    // KSVM<T>::pre_process: get_max_vm / get_scale over the saved "sigma" state (goal_ks_vm.cpp:36-87)
    for (int e = 0; e < c->ne; ++e) {
      Ten<double> sg;
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sg(i, j) = c->sigma[9 * (size_t)e + 3 * i + j];
      ks[0] = std::max(ks[0], von_mises(sg));
    }
    for (int e = 0; e < c->ne; ++e) {
      TetGeom g;
      if (!elem_geometry(*c, e, g)) return NAN;
      Ten<double> sg;
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sg(i, j) = c->sigma[9 * (size_t)e + 3 * i + j];
      ks[1] += std::exp(rho * (von_mises(sg) - ks[0])) * g.w * g.dv;
    }
  }
  double J = 0.0;
  for (int e = 0; e < c->ne; ++e) {
    if (dMdu) {  // QoI<FADT>::scatter (goal_qoi.cpp:63-76)
      Fad ev;
      if (!qoi_element<Fad>(*c, e, type, elem_set, ks, ev)) return NAN;
      for (int n = 0; n < 4; ++n)
        for (int eq = 0; eq < 4; ++eq) dMdu[4 * (size_t)c->conn[4 * (size_t)e + n] + eq] += ev.dx(4 * n + eq);
      J += ev.v;
    } else {  // QoI<ST>::scatter (goal_qoi.cpp:31-34)
      double ev;
      if (!qoi_element<double>(*c, e, type, elem_set, ks, ev)) return NAN;
      J += ev;
    }
  }
  if (type == GO_QOI_KS_VM) J = ks[0] + (1.0 / rho) * std::log(ks[1]);  // KSVM<T>::post_process (goal_ks_vm.cpp:102-105)
  return J;
}

extern "C" {

int go_assemble_error(go_ctx* c, const double* zu_diff, const double* zp_diff, const double* zp_coarse,
                      double* R) {
  c->plastic = 0; c->err.clear();
  for (int e = 0; e < c->ne; ++e) {
    TetGeom g; Weights w;
    if (!elem_geometry(*c, e, g)) return 1;
    double zu[4][3], zp[4], zpc[4];
    for (int n = 0; n < 4; ++n) {
      int const nd = c->conn[4 * (size_t)e + n];
      for (int i = 0; i < 3; ++i) zu[n][i] = zu_diff[3 * (size_t)nd + i];
      zp[n] = zp_diff[nd];
      zpc[n] = zp_coarse[nd];
    }
    adjoint_weights(g, zu, zp, zpc, w);
    double ru[12], rp[4];
    if (!chain<double>(*c, e, g, w, false, ru, rp, nullptr)) return 2;
    for (int n = 0; n < 4; ++n) {
      int const nd = c->conn[4 * (size_t)e + n];
      for (int d = 0; d < 3; ++d) R[4 * (size_t)nd + d] += ru[n * 3 + d];
      R[4 * (size_t)nd + 3] += rp[n];
    }
  }
  return 0;
}

// BForce<T>::at_point (goal_bforce.cpp:58-68), wired behind MResidual in build_resid / build_error
// (goal_mechanics.cpp:132-136, 204-208):  u->resid(n, i) -= b[i] * w->val(n, i) * ipw * dv.
// The reference evaluates b from a named expression at the integration point ("elastic squared", a constant,
// :51-56); here the caller hands the value per element (b is data of the problem, not of the path).
// w = VectorWeight: val(n,i) = N_n (goal_vector_weight.cpp:13-28); in the error chain DisplacementAdjoint:
// val(n,i) = z_i(xi) N_n (goal_displacement_adjoint.cpp:48-49) with z = u_z_diff.
int go_apply_bforce(go_ctx* c, const double* b, const double* zu_diff, double* R) {
  c->err.clear();
  for (int e = 0; e < c->ne; ++e) {
    TetGeom g;
    if (!elem_geometry(*c, e, g)) return 1;
    double z[3] = {1.0, 1.0, 1.0};
    if (zu_diff) {
      for (int i = 0; i < 3; ++i) {
        z[i] = 0.0;
        for (int n = 0; n < 4; ++n) z[i] += zu_diff[3 * (size_t)c->conn[4 * (size_t)e + n] + i] * g.BF[n];
      }
    }
    for (int n = 0; n < 4; ++n) {
      int const nd = c->conn[4 * (size_t)e + n];
      for (int i = 0; i < 3; ++i) R[4 * (size_t)nd + i] -= b[3 * (size_t)e + i] * (z[i] * g.BF[n]) * g.w * g.dv;
    }
  }
  return 0;
}

double go_element_error(go_ctx* c, const double* u_err, const double* p_err, double* eta_elem,
                        const int32_t* parent, int n_parent, double* eta_parent) {
  // compute_error (goal_error.cpp:7-35): |sum_d u_err_d(xi_c) + p_err(xi_c)|, N_n(xi_c) = 1/4
  for (int e = 0; e < c->ne; ++e) {
    int32_t const* nd = &c->conn[4 * (size_t)e];
    double const bf[4] = {1.0 - 0.25 - 0.25 - 0.25, 0.25, 0.25, 0.25};
    double pe = 0.0, ue[3] = {0, 0, 0};
    for (int n = 0; n < 4; ++n) {
      pe += p_err[nd[n]] * bf[n];
      for (int d = 0; d < 3; ++d) ue[d] += u_err[3 * (size_t)nd[n] + d] * bf[n];
    }
    double total = 0.0;
    for (int d = 0; d < 3; ++d) total += ue[d];
    total += pe;
    eta_elem[e] = std::fabs(total);
  }
  // Nested::set_error (goal_nested.cpp:395-412): parent error = sum of children
  if (parent && eta_parent) {
    for (int k = 0; k < n_parent; ++k) eta_parent[k] = 0.0;
    for (int e = 0; e < c->ne; ++e) eta_parent[parent[e]] += eta_elem[e];
  }
  // sum_contribs (goal_error.cpp:37-56), single part
  double sum = 0.0;
  for (int v = 0; v < c->nn; ++v) {
    double tmp = 0.0;
    for (int d = 0; d < 3; ++d) tmp += u_err[3 * (size_t)v + d];
    tmp += p_err[v];
    sum += std::fabs(tmp);
  }
  return sum;
}

// get_iso_target_size (goal_size_field.cpp:39-150), single part; returns G = sum_contributions
double go_size_field(go_ctx* c, const double* eta, int target, int p_order, double* vtx_size) {
  double const d = 3.0, p = p_order, alpha = 0.25, beta = 2.0;
  double G = 0.0;
  for (int e = 0; e < c->ne; ++e) G += std::pow(std::abs(eta[e]), ((2.0 * d) / (2.0 * p + d)));  // :39-52
  double const size_factor = std::pow((G / (double)target), (1.0 / d));                             // :54-59
  std::vector<double> hn(c->ne);
  static const int ev[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
  for (int e = 0; e < c->ne; ++e) {
    double h = 0.0;  // get_current_size (:61-69)
    for (int k = 0; k < 6; ++k) {
      double l2 = 0.0;
      for (int j = 0; j < 3; ++j) {
        double t = c->coords[3 * (size_t)c->conn[4 * (size_t)e + ev[k][1]] + j] - c->coords[3 * (size_t)c->conn[4 * (size_t)e + ev[k][0]] + j];
        l2 += t * t;
      }
      double const l = std::sqrt(l2);
      h += l * l;
    }
    h = std::sqrt(h / 6);
    double const r = std::pow(std::abs(eta[e]), ((-2.0) / (2.0 * p + d)));  // get_new_size (:71-81)
    double h_new = size_factor * r * h;
    if (h_new < alpha * h) h_new = alpha * h;
    if (h_new > beta * h) h_new = beta * h;
    hn[e] = h_new;
  }
  std::vector<double> sum(c->nn, 0.0);
  std::vector<int> cnt(c->nn, 0);
  for (int e = 0; e < c->ne; ++e)  // avg_to_vtx (:95-106)
    for (int n = 0; n < 4; ++n) { sum[c->conn[4 * (size_t)e + n]] += hn[e]; cnt[c->conn[4 * (size_t)e + n]]++; }
  for (int v = 0; v < c->nn; ++v) vtx_size[v] = cnt[v] ? sum[v] / cnt[v] : 0.0;
  return G;
}

int64_t go_last_plastic_count(go_ctx* c) { return c->plastic; }
const char* go_last_error(go_ctx* c) { return c->err.c_str(); }

double go_time_jacobian_elements(go_ctx* c, int64_t e0, int64_t e1, int save_state) {
  double chk = 0.0;
  for (int64_t e = e0; e < e1; ++e) {
    TetGeom g; Weights w;
    if (!elem_geometry(*c, (int)e, g)) return NAN;
    plain_weights(g, w);
    Fad ru[12], rp[4];
    if (!chain<Fad>(*c, (int)e, g, w, save_state != 0, ru, rp, nullptr)) return NAN;
    for (int k = 0; k < 12; ++k) chk += ru[k].v + ru[k].dx(k);
    for (int k = 0; k < 4; ++k) chk += rp[k].v + rp[k].dx(4 * k + 3);
  }
  return chk;
}

}  // extern "C"
