// element_math.cuh -- per-element arithmetic of the mixed u/p tet assembly path.
//
// One linear tet, one integration point (goal_assembly.cpp:78-80).  The functions
// here are what each CUDA thread runs; they are also compiled for the host by
// tests/hostcheck (a test-only build used to verify this arithmetic against the
// oracle in a container without a GPU -- it is not a product path).
//
// What replaces Sacado FAD (goal_scalar_types.hpp:9): the reference seeds 16
// directions dx[n*4+d] (goal_displacement.cpp:139-156, goal_pressure.cpp:136-151)
// and pushes 17-wide numbers through every evaluator.  Here forward mode is
// carried out per seed direction in closed form.  A displacement seed (m,k)
// perturbs F by the rank-one tensor e_k (x) G_m, so with the spatial shape
// gradients w_n = F^{-T} G_n every directional derivative collapses to a few
// scalars per (m,k) and a few 3-vectors per node m:
//     dJ        = J w_m[k]
//     d w_n     = -w_m w_n[k]
//     d B       = e_k (x) r_m + r_m (x) e_k,   B = F Cp^{-1} F^T,  r_m = F Cp^{-1} G_m
// (neo-Hookean is Cp = I).  The residual is written in spatial form,
// P G_n = tau w_n with tau = J sigma (Kirchhoff), so
//     K[(n,i),(m,k)] = vol { (d tau w_n)_i - (tau w_m)_i w_n[k] }.
// Everything stays in registers; nothing 16-wide is ever formed.
//
// Evaluator order and formulas follow (file:line under /root/reference/src):
//   kinematics      goal_kinematics.cpp:18-25
//   neohookean      goal_neohookean.cpp:44-52, 60-72, 80-86
//   J2              goal_J2.cpp:54-64, 72-143, 151-157
//   mixed           goal_mixed.cpp:34-46
//   mresidual       goal_mresidual.cpp:26-32
//   presidual       goal_presidual.cpp:54-59
//   stabilization   goal_stabilization.cpp:59-81
//   adjoint weights goal_displacement_adjoint.cpp:37-53, goal_pressure_adjoint.cpp:38-49
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define GX_HD __host__ __device__ __forceinline__
#else
#define GX_HD inline
#endif

namespace gx {

enum { MODEL_NEOHOOKEAN = 0, MODEL_J2 = 1 };

// error codes reported per launch (first failing element wins)
enum { ERR_NONE = 0, ERR_INVERTED_ELEMENT = 1, ERR_INVERTED_DEFORMATION = 2, ERR_J2_RETURN_MAP = 3 };

struct Material {  // per elem set; kappa/mu as in goal_neohookean.cpp:50-51
  double kappa, mu, K, Y, c0;
  double rkappa;  // 1/kappa
  double tauc;    // 0.5*c0/(6*mu): tau = tauc * (sum of the 6 squared edge lengths)  goal_stabilization.cpp:59-72
};
GX_HD Material make_material(double E, double nu, double K, double Y, double c0) {
  Material m;
  m.kappa = E / (3.0 * (1.0 - 2.0 * nu));
  m.mu = E / (2.0 * (1.0 + nu));
  m.K = K; m.Y = Y; m.c0 = c0;
  m.rkappa = 1.0 / m.kappa;
  m.tauc = 0.5 * c0 / (6.0 * m.mu);
  return m;
}
// 1/cbrt(x) and 1/sqrt(x): single device intrinsics, plain libm on the host build
template <class S> GX_HD S gx_rcbrt(S x) {
#if defined(__CUDA_ARCH__)
  return rcbrt(x);
#else
  return S(1.0) / cbrt(x);
#endif
}
template <class S> GX_HD S gx_rsqrt(S x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return S(1.0) / sqrt(x);
#endif
}
// Cp^{-1} = Fp^{-1} Fp^{-T} (symmetric, 00 11 22 01 02 12) of a plastic deformation gradient
// (goal_J2.cpp:84-87).  Depends on the old state only, so it is cached per element, not per pass.
template <class S> GX_HD void cp_inverse(S const Fp[9], S Cp[6]) {
  S Fpi[9];
  S const d = Fp[0] * (Fp[4] * Fp[8] - Fp[5] * Fp[7]) - Fp[1] * (Fp[3] * Fp[8] - Fp[5] * Fp[6]) + Fp[2] * (Fp[3] * Fp[7] - Fp[4] * Fp[6]);
  S const rd = S(1.0) / d;
  Fpi[0] = (Fp[4] * Fp[8] - Fp[5] * Fp[7]) * rd; Fpi[1] = (Fp[2] * Fp[7] - Fp[1] * Fp[8]) * rd; Fpi[2] = (Fp[1] * Fp[5] - Fp[2] * Fp[4]) * rd;
  Fpi[3] = (Fp[5] * Fp[6] - Fp[3] * Fp[8]) * rd; Fpi[4] = (Fp[0] * Fp[8] - Fp[2] * Fp[6]) * rd; Fpi[5] = (Fp[2] * Fp[3] - Fp[0] * Fp[5]) * rd;
  Fpi[6] = (Fp[3] * Fp[7] - Fp[4] * Fp[6]) * rd; Fpi[7] = (Fp[1] * Fp[6] - Fp[0] * Fp[7]) * rd; Fpi[8] = (Fp[0] * Fp[4] - Fp[1] * Fp[3]) * rd;
  Cp[0] = Fpi[0] * Fpi[0] + Fpi[1] * Fpi[1] + Fpi[2] * Fpi[2];
  Cp[1] = Fpi[3] * Fpi[3] + Fpi[4] * Fpi[4] + Fpi[5] * Fpi[5];
  Cp[2] = Fpi[6] * Fpi[6] + Fpi[7] * Fpi[7] + Fpi[8] * Fpi[8];
  Cp[3] = Fpi[0] * Fpi[3] + Fpi[1] * Fpi[4] + Fpi[2] * Fpi[5];
  Cp[4] = Fpi[0] * Fpi[6] + Fpi[1] * Fpi[7] + Fpi[2] * Fpi[8];
  Cp[5] = Fpi[3] * Fpi[6] + Fpi[4] * Fpi[7] + Fpi[5] * Fpi[8];
}

// symmetric 3x3 storage: 00 11 22 01 02 12
template <class S> GX_HD void sym_mv(S const t[6], S const v[3], S o[3]) {
  o[0] = t[0] * v[0] + t[3] * v[1] + t[4] * v[2];
  o[1] = t[3] * v[0] + t[1] * v[1] + t[5] * v[2];
  o[2] = t[4] * v[0] + t[5] * v[1] + t[2] * v[2];
}
template <class S> GX_HD S dot3(S const a[3], S const b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class S> GX_HD void cross3(S const a[3], S const b[3], S o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
template <class S> GX_HD S det3(S const A[9]) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}
// inverse given the determinant's reciprocal
template <class S> GX_HD void inv3(S const A[9], S rdet, S o[9]) {
  o[0] = (A[4] * A[8] - A[5] * A[7]) * rdet;
  o[1] = (A[2] * A[7] - A[1] * A[8]) * rdet;
  o[2] = (A[1] * A[5] - A[2] * A[4]) * rdet;
  o[3] = (A[5] * A[6] - A[3] * A[8]) * rdet;
  o[4] = (A[0] * A[8] - A[2] * A[6]) * rdet;
  o[5] = (A[2] * A[3] - A[0] * A[5]) * rdet;
  o[6] = (A[3] * A[7] - A[4] * A[6]) * rdet;
  o[7] = (A[1] * A[6] - A[0] * A[7]) * rdet;
  o[8] = (A[0] * A[4] - A[1] * A[3]) * rdet;
}
template <class S> GX_HD void mm3(S const A[9], S const B[9], S o[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

// exp of a 3x3 matrix by Pade approximation with Higham's (2005) order selection, which is what
// minitensor::exp does: [3/3] for ||A||_1 <= 1.4956e-2, [5/5] for <= 2.5394e-1 (every realistic plastic
// increment dgam*N lands in one of these two), otherwise [13/13] with scaling and squaring.
template <class S> GX_HD void expm3(S const A[9], S o[9]) {
  S n1 = 0;
  for (int j = 0; j < 3; ++j) {
    S c = fabs(A[j]) + fabs(A[3 + j]) + fabs(A[6 + j]);
    n1 = c > n1 ? c : n1;
  }
  S A2[9], U[9], V[9], T[9];
  if (n1 <= S(2.539398330063230e-1)) {
    mm3(A, A, A2);
    if (n1 <= S(1.495585217958292e-2)) {
      for (int i = 0; i < 9; ++i) { T[i] = A2[i]; V[i] = S(12.) * A2[i]; }
      T[0] += S(60.); T[4] += S(60.); T[8] += S(60.);
      V[0] += S(120.); V[4] += S(120.); V[8] += S(120.);
    } else {
      S A4[9];
      mm3(A2, A2, A4);
      for (int i = 0; i < 9; ++i) { T[i] = A4[i] + S(420.) * A2[i]; V[i] = S(30.) * A4[i] + S(3360.) * A2[i]; }
      T[0] += S(15120.); T[4] += S(15120.); T[8] += S(15120.);
      V[0] += S(30240.); V[4] += S(30240.); V[8] += S(30240.);
    }
    mm3(A, T, U);
  } else {
    int sq = 0;
    S As[9];
    S const th13 = S(5.371920351148152);
    S sc = S(1.0);
    while (n1 * sc > th13 && sq < 60) { sc *= S(0.5); ++sq; }
    for (int i = 0; i < 9; ++i) As[i] = A[i] * sc;
    S A4[9], A6[9];
    mm3(As, As, A2); mm3(A2, A2, A4); mm3(A2, A4, A6);
    S const b[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                     129060195264000., 10559470521600., 670442572800., 33522128640., 1323241920.,
                     40840800., 960960., 16380., 182., 1.};
    S W1[9], W2[9];
    for (int i = 0; i < 9; ++i) W1[i] = b[13] * A6[i] + b[11] * A4[i] + b[9] * A2[i];
    mm3(A6, W1, W2);
    for (int i = 0; i < 9; ++i) W2[i] += b[7] * A6[i] + b[5] * A4[i] + b[3] * A2[i];
    W2[0] += b[1]; W2[4] += b[1]; W2[8] += b[1];
    mm3(As, W2, U);
    for (int i = 0; i < 9; ++i) W1[i] = b[12] * A6[i] + b[10] * A4[i] + b[8] * A2[i];
    mm3(A6, W1, V);
    for (int i = 0; i < 9; ++i) V[i] += b[6] * A6[i] + b[4] * A4[i] + b[2] * A2[i];
    V[0] += b[0]; V[4] += b[0]; V[8] += b[0];
    for (int i = 0; i < 9; ++i) { T[i] = V[i] - U[i]; A2[i] = V[i] + U[i]; }
    S Ti[9];
    inv3(T, S(1.0) / det3(T), Ti);
    mm3(Ti, A2, o);
    for (int s = 0; s < sq; ++s) { mm3(o, o, T); for (int i = 0; i < 9; ++i) o[i] = T[i]; }
    return;
  }
  for (int i = 0; i < 9; ++i) { T[i] = V[i] - U[i]; A2[i] = V[i] + U[i]; }
  S Ti[9];
  inv3(T, S(1.0) / det3(T), Ti);
  mm3(Ti, A2, o);
}

// ---------------------------------------------------------------------------
// Value-level state of one element shared by the residual and every Jacobian
// column.  All vectors are indexed by local node n = 0..3.
// ---------------------------------------------------------------------------
template <class S>
struct Core {
  S vol;       // w*dv = det(J_geom)/6                       goal_assembly.cpp:78-80
  S taus;      // 0.5*c0*h^2/mu                              goal_stabilization.cpp:72
  S J;         // det F                                      goal_kinematics.cpp:24
  S pv;        // p at the centroid                          goal_pressure.cpp:153-158
  S w[4][3];   // spatial shape gradients F^{-T} G_n
  S r[4][3];   // F Cp^{-1} G_n (neo-Hookean: F G_n)
  S q[3];      // F^{-T} grad p
  S tau[6];    // Kirchhoff stress of the mixed Cauchy stress, J*sigma
  S s[6];      // trial deviatoric Kirchhoff stress (goal_J2.cpp:89)
  S beta;      // radial-return scaling of s (1 when elastic)
  S c1;        // mu*J^{-2/3}
  S mubar;     // mu*tr(be)/3 (J2)
  int plastic;
  // ---- tangent data, pre-scaled by vol (see "closed-form tangent" below)
  // vol * tau = vb s + Jpv I and vol * (g_r I + g_N N) = gNs s + vgr I are carried as scalars next to s
  S vb;        // vol * beta
  S gNs;       // vol * g_N / |s|
  S vgr;       // vol * g_r
  S gwv;       // vol * g_w
  S A1v;       // vol * beta * c1
  S Jpv;       // vol * J * p
  S upc;       // vol * J / 4                  d R_u / d p_m       = upc * w_n
  S va;        // vol * (-1/8)(1 + 1/J^2) J    d R_p / d u (volumetric part)
  S tjv;       // vol * taus * J
  S ppc;       // vol / (16 kappa)
  S rb;        // vol * (p/kappa - (J - 1/J)/2) / 4     R_p without stabilization
  S rc1, tb3;  // 1/c1 and tr(B)/3:  r_n = B w_n = rc1 (s w_n) + tb3 w_n   (B = s/c1 + tr(B)/3 I), see node_r()
  S dN[6];     // plastic branch: flow increment dgam*N (symmetric), input of plastic_update
  // geometry kept for the adjoint-weighted residual
  S G[4][3];
  S Finv[9];
};

// Geometry + kinematics + stress update.  x,u: [4][3]; p: [4].
// Cp (= Cp^{-1} of the old plastic state, see cp_inverse) and eqps_old are read only for MODEL_J2.
// The plastic flow increment dgam*N is returned in c.dN; the caller turns it into the new Fp with
// plastic_update() after the Jacobian work (keeps the matrix exponential's temporaries out of the
// register-critical section).  When `save` is set the state the
// reference's evaluators would leave behind is written to sigma_out (mixed Cauchy,
// goal_mixed.cpp:44-45) and eqps_out; Fp is the caller's job -- plastic branch only
// (goal_J2.cpp:128-136; on the elastic branch Fp is deliberately left untouched).
// Returns an ERR_* code.
template <int MODEL, class S>
GX_HD int element_core(S const x[4][3], S const u[4][3], S const p[4], Material const& mat, S const Cp[6],
                       S eqps_old, bool save, S sigma_out[9], S& eqps_out, Core<S>& c) {
  // ---- linear tet geometry: G_n = grad N_n, vol = det/6, h^2 = mean squared edge length
  S e1[3], e2[3], e3[3];
  for (int j = 0; j < 3; ++j) { e1[j] = x[1][j] - x[0][j]; e2[j] = x[2][j] - x[0][j]; e3[j] = x[3][j] - x[0][j]; }
  S c23[3], c31[3], c12[3];
  cross3(e2, e3, c23); cross3(e3, e1, c31); cross3(e1, e2, c12);
  S const dv = dot3(e1, c23);
  if (!(dv > S(0.0))) return ERR_INVERTED_ELEMENT;
  // grad u = (1/dv) sum_n (u_n - u_0) (x) c_n with c_n the cofactor columns
  S Fd[9];  // dv * grad u
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      Fd[3 * i + j] = (u[1][i] - u[0][i]) * c23[j] + (u[2][i] - u[0][i]) * c31[j] + (u[3][i] - u[0][i]) * c12[j];
  S const rdv = S(1.0) / dv;
  for (int j = 0; j < 3; ++j) {
    c.G[1][j] = c23[j] * rdv; c.G[2][j] = c31[j] * rdv; c.G[3][j] = c12[j] * rdv;
    c.G[0][j] = -(c.G[1][j] + c.G[2][j] + c.G[3][j]);
  }
  c.vol = dv * S(1.0 / 6.0);
  S h2 = dot3(e1, e1) + dot3(e2, e2) + dot3(e3, e3);
  for (int j = 0; j < 3; ++j) {
    S a = x[2][j] - x[1][j], b = x[3][j] - x[1][j], d = x[3][j] - x[2][j];
    h2 += a * a + b * b + d * d;
  }
  c.taus = mat.tauc * h2;

  // ---- F = I + grad u, J, F^{-1}; p and grad p at the centroid
  S F[9];
  for (int i = 0; i < 9; ++i) F[i] = Fd[i] * rdv;
  F[0] += S(1.0); F[4] += S(1.0); F[8] += S(1.0);
  S const J = det3(F);
  if (!(J > S(0.0))) return ERR_INVERTED_DEFORMATION;
  c.J = J;
  S const rJ = S(1.0) / J;
  inv3(F, rJ, c.Finv);
  c.pv = S(0.25) * p[0] + S(0.25) * p[1] + S(0.25) * p[2] + S(0.25) * p[3];
  S gp[3];
  for (int j = 0; j < 3; ++j) gp[j] = (p[1] - p[0]) * c.G[1][j] + (p[2] - p[0]) * c.G[2][j] + (p[3] - p[0]) * c.G[3][j];
  for (int k = 0; k < 3; ++k) c.q[k] = c.Finv[k] * gp[0] + c.Finv[3 + k] * gp[1] + c.Finv[6 + k] * gp[2];
  for (int n = 1; n < 4; ++n)
    for (int k = 0; k < 3; ++k) c.w[n][k] = c.Finv[k] * c.G[n][0] + c.Finv[3 + k] * c.G[n][1] + c.Finv[6 + k] * c.G[n][2];
  for (int k = 0; k < 3; ++k) c.w[0][k] = -(c.w[1][k] + c.w[2][k] + c.w[3][k]);

  S const pr = S(0.5) * mat.kappa * (J - rJ);  // U'(J) kappa, replaced by p in Mixed
  S Jm23;
  S snew[6];
  c.plastic = 0;
  c.beta = S(1.0);
  c.mubar = S(0.0);
  S g_r = S(0.0), g_N = S(0.0), g_w = S(-2.0 / 3.0);  // elastic: gamma_m = -(2/3) w_m
  S gNs = S(0.0);
  if (MODEL == MODEL_NEOHOOKEAN) {
    S const Jm13 = gx_rcbrt(J);
    Jm23 = Jm13 * Jm13;
    // b = F F^T
    S b[6];
    b[0] = F[0] * F[0] + F[1] * F[1] + F[2] * F[2];
    b[1] = F[3] * F[3] + F[4] * F[4] + F[5] * F[5];
    b[2] = F[6] * F[6] + F[7] * F[7] + F[8] * F[8];
    b[3] = F[0] * F[3] + F[1] * F[4] + F[2] * F[5];
    b[4] = F[0] * F[6] + F[1] * F[7] + F[2] * F[8];
    b[5] = F[3] * F[6] + F[4] * F[7] + F[5] * F[8];
    S const tr3 = (b[0] + b[1] + b[2]) * S(1.0 / 3.0);
    S const c1 = mat.mu * Jm23;
    c.s[0] = c1 * (b[0] - tr3); c.s[1] = c1 * (b[1] - tr3); c.s[2] = c1 * (b[2] - tr3);
    c.s[3] = c1 * b[3]; c.s[4] = c1 * b[4]; c.s[5] = c1 * b[5];
    c.c1 = c1;
    c.rc1 = S(1.0) / c1; c.tb3 = tr3;
    for (int n = 1; n < 4; ++n)
      for (int i = 0; i < 3; ++i) c.r[n][i] = F[3 * i] * c.G[n][0] + F[3 * i + 1] * c.G[n][1] + F[3 * i + 2] * c.G[n][2];
    for (int i = 0; i < 6; ++i) snew[i] = c.s[i];
  } else {
    {
      S const Jm13 = gx_rcbrt(J);  // J^{-2/3} (goal_J2.cpp:80 uses pow; same value to round-off)
      Jm23 = Jm13 * Jm13;
    }
    // M = F Cp^{-1}; B = M F^T (symmetric)
    S M[9];
    for (int i = 0; i < 3; ++i) {
      S const* f = &F[3 * i];
      M[3 * i + 0] = f[0] * Cp[0] + f[1] * Cp[3] + f[2] * Cp[4];
      M[3 * i + 1] = f[0] * Cp[3] + f[1] * Cp[1] + f[2] * Cp[5];
      M[3 * i + 2] = f[0] * Cp[4] + f[1] * Cp[5] + f[2] * Cp[2];
    }
    S B[6];
    B[0] = M[0] * F[0] + M[1] * F[1] + M[2] * F[2];
    B[1] = M[3] * F[3] + M[4] * F[4] + M[5] * F[5];
    B[2] = M[6] * F[6] + M[7] * F[7] + M[8] * F[8];
    B[3] = M[0] * F[3] + M[1] * F[4] + M[2] * F[5];
    B[4] = M[0] * F[6] + M[1] * F[7] + M[2] * F[8];
    B[5] = M[3] * F[6] + M[4] * F[7] + M[5] * F[8];
    S const c1 = mat.mu * Jm23;
    c.c1 = c1;
    S const trB = B[0] + B[1] + B[2];
    S const tr3 = trB * S(1.0 / 3.0);
    c.s[0] = c1 * (B[0] - tr3); c.s[1] = c1 * (B[1] - tr3); c.s[2] = c1 * (B[2] - tr3);
    c.s[3] = c1 * B[3]; c.s[4] = c1 * B[4]; c.s[5] = c1 * B[5];
    c.mubar = c1 * tr3;  // mu * trace(be) / 3
    c.rc1 = S(1.0) / c1; c.tb3 = tr3;
    for (int n = 1; n < 4; ++n)
      for (int i = 0; i < 3; ++i) c.r[n][i] = M[3 * i] * c.G[n][0] + M[3 * i + 1] * c.G[n][1] + M[3 * i + 2] * c.G[n][2];
    S const s2 = c.s[0] * c.s[0] + c.s[1] * c.s[1] + c.s[2] * c.s[2] +
                 S(2.0) * (c.s[3] * c.s[3] + c.s[4] * c.s[4] + c.s[5] * c.s[5]);
    S const rs = s2 > S(0.0) ? gx_rsqrt(s2) : S(0.0);
    S const smag = s2 * rs;
    S const sq23 = S(0.81649658092772603273);  // sqrt(2/3)
    S const f = smag - sq23 * (mat.Y + mat.K * eqps_old);
    eqps_out = eqps_old;
    for (int i = 0; i < 6; ++i) snew[i] = c.s[i];
    if (f > S(1.0e-12)) {
      // Radial return.  With linear hardening the reference's Newton loop on X (goal_J2.cpp:108-121) is exact after
      // two iterations: X1 = f / (2 mubar), R1 = -(2/3) K X1, X2 = f / (2 mubar + 2K/3), R2 = 0.  It stops after the
      // FIRST iteration when |R1| already passes its test (|R| < 1e-11, |R|/Y < 1e-11 or |R|/f < 1e-11: elements that
      // barely yield, f below about 1e-9, or K = 0) and then differentiates X1, not X2 -- reproduced here through the
      // denominator, which is all the closed-form tangent below depends on.
      c.plastic = 1;
      S const mubar = c.mubar;
      S const tm = S(2.0) * mubar, r1n = S(2.0 / 3.0) * mat.K * f;  // |R1| = r1n / tm
      bool const first = (r1n < S(1.0e-11) * tm) || (r1n < S(1.0e-11) * mat.Y * tm) || (S(2.0 / 3.0) * mat.K < S(1.0e-11) * tm);
      S const rD = S(1.0) / (first ? tm : tm + S(2.0 / 3.0) * mat.K);
      S const dgam = f * rD;
      // goal_J2.cpp:119-120: fail("J2: return mapping failed") after 30 iterations -- with a linear residual that
      // only happens when the iteration runs on non-finite numbers (f = +Inf, NaN state)
      if (!(dgam - dgam == S(0.0))) return ERR_J2_RETURN_MAP;
      S N[6];
      for (int i = 0; i < 6; ++i) N[i] = c.s[i] * rs;
      c.beta = S(1.0) - S(2.0) * mubar * dgam * rs;
      for (int i = 0; i < 6; ++i) snew[i] = c.s[i] - S(2.0) * mubar * dgam * N[i];
      eqps_out = eqps_old + sq23 * dgam;
      // gamma_mk = d beta(m,k) - (2/3) beta w_m[k] is linear in (r_m, N r_m, w_m):
      //   gamma_m = (g_r I + g_N N) r_m + g_w w_m
      g_r = S(-4.0 / 3.0) * rs * dgam * c1 * (S(1.0) - S(2.0) * mubar * rD);
      g_N = S(4.0) * c1 * mubar * rs * (dgam * rs - rD);
      g_w = S(2.0 / 3.0) * c.beta * (S(2.0) * mubar * rD - S(1.0));
      gNs = g_N * rs;
      for (int i = 0; i < 6; ++i) c.dN[i] = dgam * N[i];
    }
  }
  c.r[0][0] = -(c.r[1][0] + c.r[2][0] + c.r[3][0]);
  c.r[0][1] = -(c.r[1][1] + c.r[2][1] + c.r[3][1]);
  c.r[0][2] = -(c.r[1][2] + c.r[2][2] + c.r[3][2]);
  // sigma = s/J + pr*I (model), then Mixed: sigma_ii += p - tr(sigma)/3
  S sig[6];
  for (int i = 0; i < 3; ++i) sig[i] = snew[i] * rJ + pr;
  for (int i = 3; i < 6; ++i) sig[i] = snew[i] * rJ;
  S const pbar = (sig[0] + sig[1] + sig[2]) * S(1.0 / 3.0);
  for (int i = 0; i < 3; ++i) sig[i] += c.pv - pbar;
  if (save) {
    sigma_out[0] = sig[0]; sigma_out[4] = sig[1]; sigma_out[8] = sig[2];
    sigma_out[1] = sigma_out[3] = sig[3];
    sigma_out[2] = sigma_out[6] = sig[4];
    sigma_out[5] = sigma_out[7] = sig[5];
  }
  for (int i = 0; i < 6; ++i) c.tau[i] = J * sig[i];
  // ---- pre-scaled tangent data
  S const vol = c.vol;
  c.vb = vol * c.beta;
  c.gNs = vol * gNs;
  c.vgr = vol * g_r;
  c.gwv = vol * g_w;
  c.A1v = vol * c.beta * c.c1;
  S const vJ = vol * J;
  c.Jpv = vJ * c.pv;
  c.upc = S(0.25) * vJ;
  c.va = S(-0.125) * (vJ + vol * rJ);  // vol * (-1/8)(1 + 1/J^2) J
  c.tjv = c.taus * vJ;
  c.ppc = S(1.0 / 16.0) * vol * mat.rkappa;
  c.rb = S(0.25) * vol * (c.pv * mat.rkappa - S(0.5) * (J - rJ));
  return ERR_NONE;
}

// Fp = exp(dgam N) Fp_old  (goal_J2.cpp:128-131).  Only called for elements on the plastic branch; on the
// elastic branch the reference leaves Fp untouched (goal_J2.cpp:135-136).
template <class S> GX_HD void plastic_update(S const dN[6], S const Fp_old[9], S Fp_new[9]) {
  S const A[9] = {dN[0], dN[3], dN[4], dN[3], dN[1], dN[5], dN[4], dN[5], dN[2]};
  S E[9];
  expm3(A, E);
  mm3(E, Fp_old, Fp_new);
}

// ---------------------------------------------------------------------------
// Residual.  R_u[n] = vol tau w_n,  R_p[n] = rb + vol taus J (q . w_n)
// ---------------------------------------------------------------------------
// One node's rows: out = (R_u[n][0..2], R_p[n]) for the node whose spatial gradient is wn.
// (vol tau) w = vb (s w) + Jpv w
template <class S> GX_HD void tau_mv(Core<S> const& c, S const w[3], S out[3]) {
  S sw[3];
  sym_mv(c.s, w, sw);
  for (int k = 0; k < 3; ++k) out[k] = c.vb * sw[k] + c.Jpv * w[k];
}
template <class S> GX_HD void element_residual_row(Core<S> const& c, S const wn[3], S out[4]) {
  tau_mv(c, wn, out);
  out[3] = c.rb + c.tjv * dot3(c.q, wn);
}
// Element residual: ru[n*3+i] (momentum), rp[n] (pressure + stabilization).
template <class S> GX_HD void element_residual(Core<S> const& c, S ru[12], S rp[4]) {
  for (int n = 0; n < 4; ++n) {
    S o[4];
    element_residual_row(c, c.w[n], o);
    ru[3 * n] = o[0]; ru[3 * n + 1] = o[1]; ru[3 * n + 2] = o[2]; rp[n] = o[3];
  }
}

// ---------------------------------------------------------------------------
// Closed-form tangent.  For the seed (m,k) (displacement k of node m):
//   K[(n,i),(m,k)] = w_n[k] A_m[i] + w_n[i] B_m[k] + g_m[k] (s w_n)[i] + delta_ik (rA_m . w_n)
//     A_m  = vol (beta c1 r_m - tau w_m)          material + geometric stiffness
//     B_m  = vol (-(2/3) beta c1 r_m + J p w_m)   d(J^{-2/3}) and d(J p) terms
//     g_m  = vol (Gamma r_m + g_w w_m)            radial-return scaling d beta (zero Gamma when elastic)
//     rA_m = vol beta c1 r_m
//   K[(n,p),(m,k)] = w_m[k] cq_n - (w_m . w_n) tq[k] - tqw_m w_n[k],
//     cq_n = va + tjv (q . w_n),  tq = tjv q,  tqw_m = tjv (q . w_m)
//   K[(n,i),(m,p)] = upc w_n[i]        K[(n,p),(m,p)] = ppc + tjv (w_m . w_n)
// ---------------------------------------------------------------------------
template <class S>
struct ColNode {
  S w[3];   // w_m
  S A[3], B[3], g[3], rA[3];
  S tqw;
};

// wm = w_m, rm = r_m (passed explicitly so that callers with a run-time node index can select them
// without indexing the register-resident Core dynamically)
template <class S> GX_HD void column_node(Core<S> const& c, S const wm[3], S const rm[3], ColNode<S>& cn) {
  S tw[3], gr[3], sr[3];
  tau_mv(c, wm, tw);
  sym_mv(c.s, rm, sr);
  for (int k = 0; k < 3; ++k) gr[k] = c.gNs * sr[k] + c.vgr * rm[k];
  S const m23 = S(-2.0 / 3.0);
  for (int k = 0; k < 3; ++k) {
    cn.w[k] = wm[k];
    cn.rA[k] = c.A1v * rm[k];
    cn.A[k] = cn.rA[k] - tw[k];
    cn.B[k] = m23 * cn.rA[k] + c.Jpv * wm[k];
    cn.g[k] = gr[k] + c.gwv * wm[k];
  }
  cn.tqw = c.tjv * dot3(c.q, wm);
}

// r_m = F Cp^{-1} G_m = B w_m with B = F Cp^{-1} F^T = s/c1 + tr(B)/3 I: rebuilt from the spatial gradient, so the
// tangent records of the two-kernel Jacobian pass carry w_n only.  sw = s w (returned: the callers need it too).
template <class S> GX_HD void node_r(Core<S> const& c, S const w[3], S sw[3], S r[3]) {
  sym_mv(c.s, w, sw);
  for (int k = 0; k < 3; ++k) r[k] = c.rc1 * sw[k] + c.tb3 * w[k];
}
// column_node for callers that hold w_m only
template <class S> GX_HD void column_node_w(Core<S> const& c, S const wm[3], ColNode<S>& cn) {
  S sw[3], rm[3], sr[3];
  node_r(c, wm, sw, rm);
  sym_mv(c.s, rm, sr);
  S const m23 = S(-2.0 / 3.0);
  for (int k = 0; k < 3; ++k) {
    cn.w[k] = wm[k];
    cn.rA[k] = c.A1v * rm[k];
    cn.A[k] = -(c.vb * sw[k]) + (-(c.Jpv * wm[k]) + cn.rA[k]);  // rA - vol tau w_m as two fused multiply-adds
    cn.B[k] = m23 * cn.rA[k] + c.Jpv * wm[k];
    cn.g[k] = (c.gNs * sr[k] + c.vgr * rm[k]) + c.gwv * wm[k];
  }
  cn.tqw = c.tjv * dot3(c.q, wm);
}

// Row-node quantities shared by the four column nodes.
template <class S>
struct RowNode {
  S w[3];   // w_n
  S sw[3];  // s w_n
  S cq;     // va + tjv (q . w_n)
  S up[3];  // upc w_n
};
template <class S> GX_HD void row_node(Core<S> const& c, S const wn[3], RowNode<S>& rn) {
  for (int k = 0; k < 3; ++k) { rn.w[k] = wn[k]; rn.up[k] = c.upc * wn[k]; }
  sym_mv(c.s, wn, rn.sw);
  rn.cq = c.va + c.tjv * dot3(c.q, wn);
}

// 4x4 block K[(n,i),(m,k)], i,k = 0..3 (eq 3 = pressure), row-major in `blk`.
template <class S>
GX_HD void jacobian_block(Core<S> const& c, RowNode<S> const& rn, ColNode<S> const& cn, S blk[16]) {
  S const d = dot3(cn.rA, rn.w);
  S const W = dot3(cn.w, rn.w);
  for (int i = 0; i < 3; ++i) {
    for (int k = 0; k < 3; ++k) {
      S v = rn.w[k] * cn.A[i] + rn.w[i] * cn.B[k] + cn.g[k] * rn.sw[i];
      if (i == k) v += d;
      blk[4 * i + k] = v;
    }
    blk[4 * i + 3] = rn.up[i];
  }
  for (int k = 0; k < 3; ++k) blk[12 + k] = cn.w[k] * rn.cq - W * (c.tjv * c.q[k]) - cn.tqw * rn.w[k];
  blk[15] = c.ppc + c.tjv * W;
}

// acc += K[(n,.),(m,.)] (TRANSPOSE: acc += the transposed block), every term as one fused multiply-add into the
// accumulator: 3 DFMA per displacement entry instead of DMUL + 2 DFMA for the block and a DADD to accumulate it.
template <bool TRANSPOSE, class S>
GX_HD void jacobian_block_add(Core<S> const& c, RowNode<S> const& rn, ColNode<S> const& cn, S acc[16]) {
  S const d = dot3(cn.rA, rn.w);
  S const W = dot3(cn.w, rn.w);
  for (int i = 0; i < 3; ++i) {
    for (int k = 0; k < 3; ++k) {
      S& a = acc[TRANSPOSE ? 4 * k + i : 4 * i + k];
      a = rn.w[k] * cn.A[i] + a;
      a = rn.w[i] * cn.B[k] + a;
      a = cn.g[k] * rn.sw[i] + a;
      if (i == k) a += d;
    }
    S& u = acc[TRANSPOSE ? 12 + i : 4 * i + 3];
    u = c.upc * rn.w[i] + u;
  }
  S const tW = c.tjv * W;
  for (int k = 0; k < 3; ++k) {
    S& a = acc[TRANSPOSE ? 4 * k + 3 : 12 + k];
    a = cn.w[k] * rn.cq + a;
    a = -(tW * c.q[k]) + a;
    a = -(cn.tqw * rn.w[k]) + a;
  }
  acc[15] += c.ppc + tW;
}

// Residual of the error chain: the same integrand tested with the adjoint-weighted
// partition of unity  w_n^i = z_i N_n,  d_j w_n^i = d_j z_i N_n + z_i d_j N_n
// (goal_displacement_adjoint.cpp:48-52); PResidual uses z_p-diff, Stabilization uses
// z_p-coarse (goal_mechanics.cpp:214).  zu: [4][3] nodal u_z_diff; zp, zpc: [4].
template <class S>
GX_HD void element_error_residual(Core<S> const& c, S const zu[4][3], S const zp[4],
                                  S const zpc[4], S ru[12], S rp[4]) {
  S z[3], gz[9];
  for (int i = 0; i < 3; ++i) {
    z[i] = S(0.25) * zu[0][i] + S(0.25) * zu[1][i] + S(0.25) * zu[2][i] + S(0.25) * zu[3][i];
    for (int j = 0; j < 3; ++j)
      gz[3 * i + j] = (zu[1][i] - zu[0][i]) * c.G[1][j] + (zu[2][i] - zu[0][i]) * c.G[2][j] + (zu[3][i] - zu[0][i]) * c.G[3][j];
  }
  S const zs = S(0.25) * zp[0] + S(0.25) * zp[1] + S(0.25) * zp[2] + S(0.25) * zp[3];
  S const zc = S(0.25) * zpc[0] + S(0.25) * zpc[1] + S(0.25) * zpc[2] + S(0.25) * zpc[3];
  S gzc[3], hz[3];
  for (int j = 0; j < 3; ++j) gzc[j] = (zpc[1] - zpc[0]) * c.G[1][j] + (zpc[2] - zpc[0]) * c.G[2][j] + (zpc[3] - zpc[0]) * c.G[3][j];
  for (int k = 0; k < 3; ++k) hz[k] = c.Finv[k] * gzc[0] + c.Finv[3 + k] * gzc[1] + c.Finv[6 + k] * gzc[2];  // F^{-T} grad z_pc
  // P = tau F^{-T}:  P_ij = sum_l tau_il Finv_jl ;  pg_i = sum_j P_ij gz_ij
  S const T[9] = {c.tau[0], c.tau[3], c.tau[4], c.tau[3], c.tau[1], c.tau[5], c.tau[4], c.tau[5], c.tau[2]};
  S pg[3];
  for (int i = 0; i < 3; ++i) {
    S acc = S(0.0);
    for (int j = 0; j < 3; ++j) {
      S const Pij = T[3 * i] * c.Finv[3 * j] + T[3 * i + 1] * c.Finv[3 * j + 1] + T[3 * i + 2] * c.Finv[3 * j + 2];
      acc += Pij * gz[3 * i + j];
    }
    pg[i] = acc;
  }
  S const base = c.rb * zs;
  S const qh = dot3(c.q, hz) * S(0.25);
  for (int n = 0; n < 4; ++n) {
    S tw[3];
    tau_mv(c, c.w[n], tw);
    for (int i = 0; i < 3; ++i) ru[3 * n + i] = S(0.25) * c.vol * pg[i] + z[i] * tw[i];
    rp[n] = base + c.tjv * (qh + zc * dot3(c.q, c.w[n]));
  }
}

// ---------------------------------------------------------------------------
// Functionals of the stress (AvgVM goal_avg_vm.cpp:43-61, KSVM goal_ks_vm.cpp:89-99): von Mises stress of the
// mixed Cauchy stress (compute_von_mises, goal_von_mises.cpp:6-18) and, in place of the FADT evaluation, its
// closed-form derivative.  dev(sigma) = beta s / J, so vm = sqrt(3/2) beta |s| / J, independent of p, and for the
// seed (m,k), with d(beta s) = gamma_mk s + beta c1 dev(e_k (x) r_m + r_m (x) e_k)  (gamma as in element_core),
//   d vm = sqrt(3/2)/J { |s| gamma_mk + 2 beta c1 (N r_m)[k] - beta |s| w_m[k] },   N = s/|s|.
// Returns vm; dvm[m][k] = vol * d vm / d u_(m,k)  (vol = w dv, the weight both functionals integrate with).
// ---------------------------------------------------------------------------
template <class S> GX_HD S element_von_mises(Core<S> const& c, S dvm[4][3]) {
  S const s2 = c.s[0] * c.s[0] + c.s[1] * c.s[1] + c.s[2] * c.s[2] + S(2.0) * (c.s[3] * c.s[3] + c.s[4] * c.s[4] + c.s[5] * c.s[5]);
  S const rs = s2 > S(0.0) ? gx_rsqrt(s2) : S(0.0);
  S const smag = s2 * rs;
  S const k = S(1.2247448713915890491) / c.J;  // sqrt(3/2) / J
  for (int m = 0; m < 4; ++m) {
    S sr[3];
    sym_mv(c.s, c.r[m], sr);
    for (int d = 0; d < 3; ++d) {
      S const g = c.gNs * sr[d] + c.vgr * c.r[m][d] + c.gwv * c.w[m][d];  // vol * gamma_m[d]
      dvm[m][d] = k * (smag * g + S(2.0) * c.A1v * rs * sr[d] - c.vb * smag * c.w[m][d]);
    }
  }
  return k * (c.vb / c.vol) * smag;
}
// von Mises stress of a stored stress tensor (row-major 3x3), the reference's formula
template <class S> GX_HD S von_mises9(S const t[9]) {
  S const s1 = (t[0] - t[4]) * (t[0] - t[4]), s2 = (t[4] - t[8]) * (t[4] - t[8]), s3 = (t[8] - t[0]) * (t[8] - t[0]);
  S const s4 = t[1] * t[1], s5 = t[5] * t[5], s6 = t[6] * t[6];
  return sqrt(S(0.5) * (s1 + s2 + s3 + S(6.0) * (s4 + s5 + s6)));
}

}  // namespace gx
