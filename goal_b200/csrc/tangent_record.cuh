// tangent_record.cuh -- the per-element tangent record of the Jacobian pass and the block contributions built from it.
//
// Stage A (elem_record_kernel) evaluates an element once and leaves a 38-double record; stage B (patch_pair_kernel)
// stages the records of a patch in shared memory and every work item adds, per contributing element, either
//   * a PAIR   : the two 4x4 blocks K[(n,.),(m,.)] and K[(m,.),(n,.)] of an edge (n != m), or
//   * a DIAGONAL: the block K[(n,.),(n,.)] and the node's four residual entries,
// into register accumulators.  The functions are __host__ __device__ so that tests/hostcheck replays the schedule on
// the CPU with exactly this arithmetic (test-only; there is no CPU product path).
//
// The closed-form tangent of element_math.cuh, reduced further.  With  r_m = F Cp^{-1} G_m = (s/c1 + tr(B)/3 I) w_m
// (node_r) the column vectors of a block become linear images of the spatial gradient w_m alone:
//     A_m = vol (beta c1 r_m - tau w_m)           = a0 w_m                       a0 = A1v tb3 - Jpv
//     B_m = vol (-(2/3) beta c1 r_m + J p w_m)    = b1 (s w_m) + b0 w_m          b1 = -(2/3) vb,  b0 = Jpv - (2/3) A1v tb3
//     g_m = vol (Gamma r_m + g_w w_m)             = G w_m                        G  = gNs rc1 s^2 + (gNs tb3 + vgr rc1) s + (vgr tb3 + gwv) I
//     rA_m . w_n                                  = vb (w_m . s w_n) + a1 (w_m . w_n)     a1 = A1v tb3
// so  K[(n,i),(m,k)] = a0 w_n[k] w_m[i] + w_n[i] B_m[k] + (s w_n)[i] g_m[k] + delta_ik d_nm  with d_nm = d_mn, and a record
// carries per node only w_n (and tqw_n = tjv q.w_n), per element the symmetric s and G and 14 scalars.
// Reference: Displacement/Pressure::scatter_primal / scatter_adjoint (goal_displacement.cpp:177-214,
// goal_pressure.cpp:166-203) receive exactly these blocks from the FAD chain.
//
// Record layout (38 doubles = 19 chunks of 16 B -- an odd number, so consecutive staged records start in different
// shared-memory bank groups):
//   [4n .. 4n+2] w_n   [4n+3] tqw_n            n = 0..3       chunks 0-7
//   [16..20] s00 s11 s01 s02 s12  (s is deviatoric: s22 = -s00 - s11 is not stored)
//   [21..26] G  (00 11 22 01 02 12)
//   [27..36] vb a1 Jpv upc va ppc tjv tq0 tq1 tq2             (a0, b1, b0 follow from vb, a1, Jpv)
//   [37] rb                                                   (residual rows of the diagonal items)
// chunks 8-18 are the same for every block of the element ("common"): 11 loads, plus 2 per node.
#pragma once

#include "element_math.cuh"

namespace gx {

constexpr int TREC = 38;  // doubles per record; 304 B = 19 x 16 B
constexpr int TREC_S = 16, TREC_G = 21, TREC_SC = 27;  // offsets of s (5 stored), G (6), the 11 scalars

// Core -> record
template <class S> GX_HD void pack_trec(Core<S> const& c, S rec[TREC]) {
  for (int n = 0; n < 4; ++n) {
    rec[4 * n] = c.w[n][0]; rec[4 * n + 1] = c.w[n][1]; rec[4 * n + 2] = c.w[n][2];
    rec[4 * n + 3] = c.tjv * dot3(c.q, c.w[n]);
  }
  S const* s = c.s;
  rec[16] = s[0]; rec[17] = s[1]; rec[18] = s[3]; rec[19] = s[4]; rec[20] = s[5];
  S const k2 = c.gNs * c.rc1, k1 = c.gNs * c.tb3 + c.vgr * c.rc1, k0 = c.vgr * c.tb3 + c.gwv;
  rec[21] = k2 * (s[0] * s[0] + s[3] * s[3] + s[4] * s[4]) + k1 * s[0] + k0;
  rec[22] = k2 * (s[3] * s[3] + s[1] * s[1] + s[5] * s[5]) + k1 * s[1] + k0;
  rec[23] = k2 * (s[4] * s[4] + s[5] * s[5] + s[2] * s[2]) + k1 * s[2] + k0;
  rec[24] = k2 * (s[0] * s[3] + s[3] * s[1] + s[4] * s[5]) + k1 * s[3];
  rec[25] = k2 * (s[0] * s[4] + s[3] * s[5] + s[4] * s[2]) + k1 * s[4];
  rec[26] = k2 * (s[3] * s[4] + s[1] * s[5] + s[5] * s[2]) + k1 * s[5];
  rec[27] = c.vb;
  rec[28] = c.A1v * c.tb3;  // a1
  rec[29] = c.Jpv;
  rec[30] = c.upc; rec[31] = c.va; rec[32] = c.ppc; rec[33] = c.tjv;
  rec[34] = c.tjv * c.q[0]; rec[35] = c.tjv * c.q[1]; rec[36] = c.tjv * c.q[2];
  rec[37] = c.rb;
}
// the stored five entries of s -> symmetric storage (00 11 22 01 02 12)
GX_HD void trec_s6(double const* s5, double s[6]) {
  s[0] = s5[0]; s[1] = s5[1]; s[2] = -(s5[0] + s5[1]); s[3] = s5[2]; s[4] = s5[3]; s[5] = s5[4];
}

// One node in its two roles.  Row role: w, sw = s w, cq = va + tqw.  Column role: w, aw = a0 w, B, g, tqw.
// The functions below take the pieces of a record by pointer -- nq = rec + 4 n (node n), s[6] (trec_s6 of rec + 16),
// G = rec + 21, sc = rec + 27 (the 11 scalars) -- so that the device code can hand them registers it filled with
// 128-bit shared loads.
struct TNode {
  double w[3], tqw, sw[3], g[3];
};
GX_HD void trec_node(double const* nq, double const* s, double const* G, TNode& t) {
  t.w[0] = nq[0]; t.w[1] = nq[1]; t.w[2] = nq[2]; t.tqw = nq[3];
  sym_mv(s, t.w, t.sw);
  sym_mv(G, t.w, t.g);
}
enum { SC_VB = 0, SC_A1, SC_JPV, SC_UPC, SC_VA, SC_PPC, SC_TJV, SC_TQ0, SC_TQ1, SC_TQ2, SC_RB, SC_N };

// acc += K[(row node),(col node)]  (TRANSPOSE: acc += its transpose).  d, W, Wtq are symmetric in the two nodes and
// shared by the two blocks of a pair.
template <bool TRANSPOSE>
GX_HD void trec_block_add(double const* sc, TNode const& r, TNode const& c, double d, double W, double const Wtq[3], double acc[16]) {
  double const m23 = -2.0 / 3.0;
  double const a0 = sc[SC_A1] - sc[SC_JPV], b1 = m23 * sc[SC_VB], b0 = m23 * sc[SC_A1] + sc[SC_JPV];
  double aw[3], B[3];
  for (int k = 0; k < 3; ++k) { aw[k] = a0 * c.w[k]; B[k] = b1 * c.sw[k] + b0 * c.w[k]; }
  for (int i = 0; i < 3; ++i) {
    for (int k = 0; k < 3; ++k) {
      double& a = acc[TRANSPOSE ? 4 * k + i : 4 * i + k];
      a = r.w[k] * aw[i] + a;
      a = r.w[i] * B[k] + a;
      a = r.sw[i] * c.g[k] + a;
      if (i == k) a += d;
    }
    double& u = acc[TRANSPOSE ? 12 + i : 4 * i + 3];
    u = sc[SC_UPC] * r.w[i] + u;  // upc w_n[i]
  }
  double const cq = sc[SC_VA] + r.tqw;  // va + tjv (q . w_n)
  for (int k = 0; k < 3; ++k) {
    double& a = acc[TRANSPOSE ? 4 * k + 3 : 12 + k];
    a = c.w[k] * cq + a;
    a = a - Wtq[k];
    a = -(c.tqw * r.w[k]) + a;
  }
  acc[15] += sc[SC_PPC] + sc[SC_TJV] * W;  // ppc + tjv (w_m . w_n)
}

// symmetric scalars of the node pair (n, m)
GX_HD void trec_pair_scalars(double const* sc, TNode const& a, TNode const& b, double& d, double& W, double Wtq[3]) {
  W = dot3(a.w, b.w);
  double const S = dot3(b.w, a.sw);
  d = sc[SC_VB] * S + sc[SC_A1] * W;  // vb (w_m . s w_n) + a1 (w_m . w_n)
  for (int k = 0; k < 3; ++k) Wtq[k] = W * sc[SC_TQ0 + k];
}

// PAIR contribution of one element record, local nodes n != m (nq, mq = their node quadruples):
//   primal   : acc1 += K[(n,.),(m,.)]      acc2 += K[(m,.),(n,.)]          (blocks (a_n,a_m) and (a_m,a_n) of A)
//   transpose: acc1 += K[(m,.),(n,.)]^T    acc2 += K[(n,.),(m,.)]^T        (the same two blocks of A^T)
template <bool TRANSPOSE>
GX_HD void trec_pair_add(double const* nq, double const* mq, double const* s, double const* G, double const* sc, double acc1[16],
                         double acc2[16]) {
  TNode tn, tm;
  trec_node(nq, s, G, tn);
  trec_node(mq, s, G, tm);
  double d, W, Wtq[3];
  trec_pair_scalars(sc, tn, tm, d, W, Wtq);
  if (!TRANSPOSE) {
    trec_block_add<false>(sc, tn, tm, d, W, Wtq, acc1);
    trec_block_add<false>(sc, tm, tn, d, W, Wtq, acc2);
  } else {
    trec_block_add<true>(sc, tm, tn, d, W, Wtq, acc1);
    trec_block_add<true>(sc, tn, tm, d, W, Wtq, acc2);
  }
}

// DIAGONAL contribution, local node n: acc += K[(n,.),(n,.)] (or its transpose) and the node's residual entries
//   r4 += (vol tau w_n, rb + tjv q.w_n)        (MResidual / PResidual / Stabilization, element_residual_row)
template <bool TRANSPOSE>
GX_HD void trec_diag_add(double const* nq, double const* s, double const* G, double const* sc, double acc[16], double r4[4]) {
  TNode tn;
  trec_node(nq, s, G, tn);
  double d, W, Wtq[3];
  trec_pair_scalars(sc, tn, tn, d, W, Wtq);
  trec_block_add<TRANSPOSE>(sc, tn, tn, d, W, Wtq, acc);
  for (int k = 0; k < 3; ++k) r4[k] += sc[SC_VB] * tn.sw[k] + sc[SC_JPV] * tn.w[k];
  r4[3] += sc[SC_RB] + tn.tqw;
}
// host-side convenience (tests/hostcheck): the same two calls on a whole record
template <bool TRANSPOSE>
GX_HD void trec_pair_add_rec(double const* rec, int n, int m, double acc1[16], double acc2[16]) {
  double s[6];
  trec_s6(rec + TREC_S, s);
  trec_pair_add<TRANSPOSE>(rec + 4 * n, rec + 4 * m, s, rec + TREC_G, rec + TREC_SC, acc1, acc2);
}
template <bool TRANSPOSE>
GX_HD void trec_diag_add_rec(double const* rec, int n, double acc[16], double r4[4]) {
  double s[6];
  trec_s6(rec + TREC_S, s);
  trec_diag_add<TRANSPOSE>(rec + 4 * n, s, rec + TREC_G, rec + TREC_SC, acc, r4);
}

}  // namespace gx
