// gx_kernels.cuh -- the assembly kernels.
//
// Replaces the element loop of goal::assemble (src/goal_assembly.cpp:65-88) together
// with the gather/scatter halves of Displacement/Pressure (src/goal_displacement.cpp,
// src/goal_pressure.cpp) and States get/set (src/goal_states.cpp:21-57).
//
// Data layout in HBM
//   nodes     NodeRec[Nn]  64 B per node: x, u, p and the node's block-row descriptor; an element
//                          gathers 4 records = 4 aligned 64 B segments (4 LDG.128 each)
//   conn      int4[Ne]     one 128-bit load per element
//   bpos      uint4[Ne]    element -> nonzero scatter map, 16 x uint8 block positions
//   adj       int2[4 Ne]   node -> (element, local node) incidences + the 4 block positions that
//                          incidence writes, grouped by node (adj_off[Nn+1]); the row-owner work list
//   state_in  double[Ne][8]   Cp^{-1}[6] of Fp_old (cached), eqps_old, pad   64 B record (4 LDG.128)
//   fp_old    double[Ne][9]   read only by the one incidence that saves a plastic element's Fp
//   state_out double[Ne][20]  sigma[9], eqps, Fp[9], pad                     160 B record
//   R         double[4 Nn] ghost layout;   values  double[nnz]  CRS order of gx_graph
//
// Two schedules, both free of atomics on the data path and bit-reproducible:
//  (1) row-owner (Jacobian pass; default in its two-kernel form, see (1b) below): one warp per node a, one lane per incident element
//      (a, e).  Each lane evaluates its element and the four 4x4 blocks of node a's rows, the warp
//      sums them per target block through shared memory in a fixed order and writes node a's four
//      CRS rows exactly once, fully coalesced.  No zeroing pass, no read-modify-write traffic.
//  (2) gather form of the residual / error-localisation passes (default): element residual vectors, then one
//      thread per node sums its incidences.
//  (3) coloured elements (fallback for every pass, gx_set_option("kernel", 1)): one thread per
//      element, launches cover one colour (no two elements of a colour share a node), plain
//      read-modify-write into R / values.
//
// The element bodies are __host__ __device__ so that tests/hostcheck can run schedule (2) exactly as
// written, launch order included, on the CPU (test-only; there is no CPU product path).
#pragma once

#include <stdint.h>

#include "element_math.cuh"
#include "gx_internal.h"

namespace gx {

struct KParams {
  NodeRec const* nodes;
  ZRec const* z;
  int4 const* conn;
  uint4 const* bpos;
  uint8_t const* eset;  // may be null (single elem set)
  int32_t const* elems; // colour schedule: slot -> element
  uint32_t const* adj_off;
  int2 const* adj;
  uint32_t const* fold_ord;
  int32_t const* nblk_g;  // partitioned contexts: local (ghost) block count per node; blocks beyond it are phantom
  int fold_ld;            // row stride of the sorted fold's staging array: odd, > max incidences per node (<= 33)
  int32_t const* node_order;  // stage B visiting order (Z-curve)
  double const* state_in;
  double const* fp_old;
  double* state_out;
  double* R;
  double* values;
  int* err;                     // {code, element}
  unsigned long long* plastic;  // counter
  int e0, e1;                   // slot range of this launch (coloured schedule)
  int nn;
  int max_nblk;
  Material mat[GX_MAX_ELEM_SETS];
};

enum { PASS_RESIDUAL = 0, PASS_JACOBIAN = 1, PASS_JACOBIAN_T = 2, PASS_ERROR = 3 };

template <class T> GX_HD T ldg(T const* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one request where two 128-bit ones were needed, which halves
// the L1TEX tag-stage work of the scattered record gathers and block stores.  p must be 32 B aligned.
GX_HD void ldg256(double const* p, double& a, double& b, double& c, double& d) {
#if defined(__CUDA_ARCH__)
  asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
#else
  a = p[0]; b = p[1]; c = p[2]; d = p[3];
#endif
}
GX_HD void stg256(double* p, double a, double b, double c, double d) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#else
  p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#endif
}

GX_HD void load_node(NodeRec const* nodes, int id, double x[3], double u[3], double& p, int& blk0, int& nblk) {
  double const* q = reinterpret_cast<double const*>(nodes + id);
  double t3;
  ldg256(q, x[0], x[1], x[2], u[0]);
  ldg256(q + 4, u[1], u[2], p, t3);
#if defined(__CUDA_ARCH__)
  blk0 = __double2loint(t3);
  nblk = __double2hiint(t3);
#else
  int32_t t[2];
  __builtin_memcpy(t, &t3, 8);
  blk0 = t[0]; nblk = t[1];
#endif
}

GX_HD void report_error(int* err, int code, int e) {
#if defined(__CUDA_ARCH__)
  if (atomicCAS(err, 0, code) == 0) err[1] = e;
#else
  if (err[0] == 0) { err[0] = code; err[1] = e; }
#endif
}

// add a row of 4 doubles (32 B aligned) into the CRS values
GX_HD void add4(double* dst, double a, double b, double c, double d) {
  double2* q = reinterpret_cast<double2*>(dst);
  double2 v0 = q[0], v1 = q[1];
  v0.x += a; v0.y += b; v1.x += c; v1.y += d;
  q[0] = v0; q[1] = v1;
}

// gather + stress update of one element.  want_state: also return the state the evaluators would save
// (mixed Cauchy stress sig[9], eqps) -- writing it is the caller's job.
template <int MODEL>
GX_HD int load_element(KParams const& P, int e, bool want_state, int nd[4], int blk0[4], int nblk[4], Material const*& mat,
                       Core<double>& c, double sig[9], double& eqps_new) {
  int4 const cn = ldg(P.conn + e);
  nd[0] = cn.x; nd[1] = cn.y; nd[2] = cn.z; nd[3] = cn.w;
  double x[4][3], u[4][3], p[4];
#pragma unroll
  for (int n = 0; n < 4; ++n) load_node(P.nodes, nd[n], x[n], u[n], p[n], blk0[n], nblk[n]);
  mat = &P.mat[P.eset ? P.eset[e] : 0];
  double Cp[6], eqps_old = 0.0;
  if (MODEL == MODEL_J2) {
    double const* q = P.state_in + (int64_t)STATE_IN * e;
    double pad;
    ldg256(q, Cp[0], Cp[1], Cp[2], Cp[3]);
    ldg256(q + 4, Cp[4], Cp[5], eqps_old, pad);
  }
  eqps_new = 0.0;
  return element_core<MODEL>(x, u, p, *mat, Cp, eqps_old, want_state, sig, eqps_new, c);
}

// sigma and eqps of one element -> pairs 0-4 of its state record: two 256-bit stores and one 128-bit store
template <int MODEL>
GX_HD void store_sigma_eqps(KParams const& P, int e, double const sig[9], double eqps_new) {
  double* so = P.state_out + (int64_t)STATE_OUT * e;  // 160 B records: 32 B aligned
  stg256(so, sig[0], sig[1], sig[2], sig[3]);
  stg256(so + 4, sig[4], sig[5], sig[6], sig[7]);
  if (MODEL == MODEL_J2) *reinterpret_cast<double2*>(so + 8) = make_double2(sig[8], eqps_new);
  else so[8] = sig[8];
}

// gather + stress update of one element; writes the history state when SAVE && write_state
template <int MODEL, bool SAVE>
GX_HD int load_and_update(KParams const& P, int e, bool write_state, int nd[4], int blk0[4], int nblk[4],
                          Material const*& mat, Core<double>& c) {
  double sig[9], eqps_new = 0.0;
  int const rc = load_element<MODEL>(P, e, SAVE && write_state, nd, blk0, nblk, mat, c, sig, eqps_new);
  if (rc != ERR_NONE) return rc;
  if (SAVE && write_state) store_sigma_eqps<MODEL>(P, e, sig, eqps_new);
  return ERR_NONE;
}

// Fp of a plastic element, done after the Jacobian work; elastic elements leave Fp untouched (goal_J2.cpp:135-136)
GX_HD void save_plastic_Fp(KParams const& P, int e, double const dN[6]) {
  double Fpo[9], Fpn[9];
  double const* src = P.fp_old + 9 * (int64_t)e;
#pragma unroll
  for (int k = 0; k < 9; ++k) Fpo[k] = ldg(src + k);
  plastic_update(dN, Fpo, Fpn);
  double* so = P.state_out + (int64_t)STATE_OUT * e + SO_FP;
#pragma unroll
  for (int k = 0; k < 9; ++k) so[k] = Fpn[k];
}

// ---------------------------------------------------------------------------
// Schedule (2): one element of one colour.  Returns 1 when the element took the plastic branch.
// ---------------------------------------------------------------------------
template <int MODEL, int PASS, bool SAVE>
GX_HD int assemble_element(KParams const& P, int slot) {
  int const e = ldg(P.elems + slot);
  int nd[4], blk0[4], nblk[4];
  Material const* matp;
  Core<double> c;
  int const rc = load_and_update<MODEL, SAVE>(P, e, true, nd, blk0, nblk, matp, c);
  if (rc != ERR_NONE) { report_error(P.err, rc, e); return 0; }
  (void)matp;

  // ---- residual (Displacement/Pressure::scatter_primal, R[row] += resid)
  double ru[12], rp[4];
  if (PASS == PASS_ERROR) {
    double zu[4][3], zp[4], zpc[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      double2 const* q = reinterpret_cast<double2 const*>(P.z + nd[n]);
      double2 const a = ldg(q), b = ldg(q + 1), d = ldg(q + 2);
      zu[n][0] = a.x; zu[n][1] = a.y; zu[n][2] = b.x; zp[n] = b.y; zpc[n] = d.x;
    }
    element_error_residual(c, zu, zp, zpc, ru, rp);
  } else {
    element_residual(c, ru, rp);
  }
#pragma unroll
  for (int n = 0; n < 4; ++n) add4(P.R + 4 * (int64_t)nd[n], ru[3 * n], ru[3 * n + 1], ru[3 * n + 2], rp[n]);

  // ---- Jacobian, one column node m at a time; 4x4 node blocks go straight into the CRS
  if (PASS == PASS_JACOBIAN || PASS == PASS_JACOBIAN_T) {
    uint4 const bq = ldg(P.bpos + e);
    uint32_t const bw[4] = {bq.x, bq.y, bq.z, bq.w};  // bw[n] byte m = position of block (n,m)
    RowNode<double> rn[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) row_node(c, c.w[n], rn[n]);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      ColNode<double> cnm;
      column_node(c, c.w[m], c.r[m], cnm);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        double blk[16];
        jacobian_block(c, rn[n], cnm, blk);
        if (PASS == PASS_JACOBIAN) {
          // A(row (n,i), col (m,k)) += blk[i][k]      (scatter_primal, goal_displacement.cpp:177-194)
          int64_t const rowlen = 4 * (int64_t)nblk[n];
          double* base = P.values + 16 * (int64_t)blk0[n] + 4 * (int64_t)((bw[n] >> (8 * m)) & 0xffu);
#pragma unroll
          for (int i = 0; i < 4; ++i) add4(base + i * rowlen, blk[4 * i], blk[4 * i + 1], blk[4 * i + 2], blk[4 * i + 3]);
        } else {
          // A(row (m,k), col (n,i)) += blk[i][k]      (scatter_adjoint, goal_displacement.cpp:196-214)
          int64_t const rowlen = 4 * (int64_t)nblk[m];
          double* base = P.values + 16 * (int64_t)blk0[m] + 4 * (int64_t)((bw[m] >> (8 * n)) & 0xffu);
#pragma unroll
          for (int k = 0; k < 4; ++k) add4(base + k * rowlen, blk[k], blk[4 + k], blk[8 + k], blk[12 + k]);
        }
      }
    }
  }
  if (SAVE && MODEL == MODEL_J2 && c.plastic) save_plastic_Fp(P, e, c.dN);
  return c.plastic;
}

#if defined(__CUDACC__)
template <int MODEL, int PASS, bool SAVE>
__global__ void __launch_bounds__(128) assemble_kernel(const __grid_constant__ KParams P) {
  int const slot = P.e0 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
  int plastic = 0;
  if (slot < P.e1) plastic = assemble_element<MODEL, PASS, SAVE>(P, slot);
  if (MODEL == MODEL_J2) {
    unsigned const b = __ballot_sync(0xffffffffu, plastic != 0);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(P.plastic, (unsigned long long)__popc(b));
  }
}

// ---------------------------------------------------------------------------
// Schedule (1): row-owner Jacobian kernel.  Warp = node a; lane = incidence (a, e) with a = local
// node n of e.  Lane work: element core, then for each column node m the 4x4 block
//   PRIMAL : K[(n,.),(m,.)]           -> block (a, a_m) of node a's rows
//   ADJOINT: K[(m,.),(n,.)]^T         -> block (a, a_m) of the transposed operator
// staged in shared memory (stg[16][33]).  The blocks are then folded into the node's row accumulator
// by a fixed schedule: half-warp h = 0/1 walks the staged blocks of the lower / upper half of the active
// lanes in ascending lane order (= ascending element id), lane t of the half adding entry t of each block into its own
// accumulator copy acc[h][16 j + t]; the two copies are summed at the end.  Every CRS entry is therefore
// produced by one thread in a fixed order: deterministic, atomics-free, written exactly once.
// Nodes with more than 32 incident elements take several rounds.
// Shared memory per warp: stg 16*33*8 B + wr 24*32*8 B + acc 2*128*max_nblk B.
// ---------------------------------------------------------------------------
GX_HD int std_min_int(int a, int b) { return a < b ? a : b; }
constexpr int STG_LD = 33;
constexpr int WR_LD = 32;  // per-lane spatial vectors w_n, r_n (n = 0..3): wr[24][32]
GX_HD size_t row_owner_smem_per_warp(int max_nblk) {
  return (size_t)(16 * STG_LD + 24 * WR_LD + 2 * 16 * max_nblk) * sizeof(double);
}

template <int MODEL, bool TRANSPOSE, bool SAVE, int MINB>
__global__ void __launch_bounds__(128, MINB) row_owner_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int const wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int const half = lane >> 4, t16 = lane & 15;
  double* stg = reinterpret_cast<double*>(smem_raw + row_owner_smem_per_warp(P.max_nblk) * wib);
  double* wr = stg + 16 * STG_LD + lane;  // this lane's column of wr[24][32]: w_n[k] at 3n+k, r_n[k] at 12+3n+k
  double* acc = stg + 16 * STG_LD + 24 * WR_LD;  // [2][16*max_nblk], block-major: acc[h][16 j + 4 i + k]
  int const accld = 16 * P.max_nblk;

  int const a = blockIdx.x * (blockDim.x >> 5) + wib;
  if (a >= P.nn) return;  // whole warp exits together
  uint32_t const o0 = __ldg(P.adj_off + a), o1 = __ldg(P.adj_off + a + 1);
  int blk0a, nblka;
  {
    double2 const d3 = __ldg(reinterpret_cast<double2 const*>(P.nodes + a) + 3);
    blk0a = __double2loint(d3.y);
    nblka = __double2hiint(d3.y);
  }
  int const nent = 16 * nblka;
  for (int g = lane; g < nent; g += 32) { acc[g] = 0.0; acc[accld + g] = 0.0; }
  double racc[4] = {0.0, 0.0, 0.0, 0.0};
  int nplastic = 0;

  for (uint32_t r0 = o0; r0 < o1; r0 += 32) {
    int const nact = min(32, (int)(o1 - r0));  // active lanes of this round: 0 .. nact-1
    bool const active = lane < nact;
    int n = 0, e = 0;
    uint32_t jpack = 0;
    Core<double> c;
    bool ok = false;
    if (active) {
      int2 const ad = __ldg(P.adj + r0 + lane);
      e = ad.x >> 2;
      n = ad.x & 3; jpack = (uint32_t)ad.y;
      int nd[4], b0[4], nb[4];
      Material const* matp;
      int const rc = load_and_update<MODEL, SAVE>(P, e, n == 0, nd, b0, nb, matp, c);
      if (rc != ERR_NONE) report_error(P.err, rc, e);
      ok = rc == ERR_NONE;
      if (ok && n == 0) nplastic += c.plastic;
      // park the 24 spatial vectors in shared memory: frees 48 registers for the phases and lets the
      // run-time node indices (n, m) address them directly
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          wr[(3 * q + k) * WR_LD] = c.w[q][k];
          wr[(12 + 3 * q + k) * WR_LD] = c.r[q][k];
        }
    }
    double wn[3], rn3[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      wn[k] = wr[(3 * n + k) * WR_LD];
      rn3[k] = wr[(12 + 3 * n + k) * WR_LD];
    }
    RowNode<double> rown;  // PRIMAL: this lane's row node
    ColNode<double> coln;  // ADJOINT: this lane's column node
    if (ok) {
      double r4[4];
      element_residual_row(c, wn, r4);
      racc[0] += r4[0]; racc[1] += r4[1]; racc[2] += r4[2]; racc[3] += r4[3];
      if (!TRANSPOSE) row_node(c, wn, rown);
      else column_node(c, wn, rn3, coln);
    }
#pragma unroll 1
    for (int m = 0; m < 4; ++m) {
      uint32_t jm = 0;
      if (ok) {
        // column node m of this phase
        double wm[3], rm[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          wm[k] = wr[(3 * m + k) * WR_LD];
          rm[k] = wr[(12 + 3 * m + k) * WR_LD];
        }
        double blk[16];
        if (!TRANSPOSE) {
          ColNode<double> cnm;
          column_node(c, wm, rm, cnm);
          jacobian_block(c, rown, cnm, blk);
#pragma unroll
          for (int t = 0; t < 16; ++t) stg[t * STG_LD + lane] = blk[t];
        } else {
          RowNode<double> rnm;
          row_node(c, wm, rnm);
          jacobian_block(c, rnm, coln, blk);
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) stg[(4 * i + k) * STG_LD + lane] = blk[4 * k + i];
        }
        jm = (jpack >> (8 * m)) & 0xffu;
      } else if (active) {
#pragma unroll
        for (int t = 0; t < 16; ++t) stg[t * STG_LD + lane] = 0.0;  // failed element: contributes nothing
      }
      __syncwarp();
      // fold: the staged blocks of lanes [0, split) go to half-warp 0, [split, nact) to half-warp 1,
      // each in ascending lane order; lane t of a half adds entry t of every block it walks
      {
        int const split = (nact + 1) >> 1;
        int const lbase = half ? split : 0;
        int const lend = half ? nact : split;
        double* my = acc + half * accld + t16;
#pragma unroll 4
        for (int it = 0; it < split; ++it) {
          int const l = lbase + it;
          uint32_t const j = __shfl_sync(0xffffffffu, jm, l & 31);
          if (l < lend) my[16 * j] += stg[t16 * STG_LD + l];
        }
      }
      __syncwarp();
    }
    if (SAVE && MODEL == MODEL_J2 && ok && n == 0 && c.plastic) save_plastic_Fp(P, e, c.dN);
  }
  // ---- R rows of node a: fixed butterfly over the lanes
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double v = racc[i];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    racc[i] = v;
  }
  if (lane == 0) {
    double2* q = reinterpret_cast<double2*>(P.R + 4 * (int64_t)a);
    q[0] = make_double2(racc[0], racc[1]);
    q[1] = make_double2(racc[2], racc[3]);
  }
  // ---- node a's four CRS rows, written once: row i = [4 nblk] contiguous doubles, gathered from the
  //      block-major accumulators acc[h][16 j + 4 i + k]
  double* out = P.values + 16 * (int64_t)blk0a;
  int const rl = 4 * nblka;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    for (int cidx = lane; cidx < rl; cidx += 32) {
      int const g = 16 * (cidx >> 2) + 4 * i + (cidx & 3);
      out[i * rl + cidx] = acc[g] + acc[accld + g];
    }
  if (MODEL == MODEL_J2) {
    unsigned const tot = __reduce_add_sync(0xffffffffu, (unsigned)nplastic);
    if (lane == 0 && tot) atomicAdd(P.plastic, (unsigned long long)tot);
  }
}

// ---------------------------------------------------------------------------
// Schedule (1b), the default Jacobian pass: the row-owner schedule split in two kernels so that the
// element core is evaluated once per element instead of once per incidence.
//   stage A  elem_record_kernel : one thread per element.  Gather, stress update, state save, and the
//            34-double "tangent record" of the element (everything the 4x4 blocks are built from):
//              w_n[3] x 4 nodes | s[6] | q[3] gwv A1v Jpv upc va tjv ppc vb gNs vgr rb rc1 tb3
//            (r_n = F Cp^{-1} G_n is not stored: r_n = rc1 (s w_n) + tb3 w_n, element_math.cuh node_r)
//   stage B  row_fold_kernel    : one warp per node, one lane per incidence.  Each lane reads its
//            element's record (272 B, contiguous), builds the four blocks of the node's rows, and the
//            warp folds and writes the node's CRS rows once, exactly like row_owner_kernel.
// Costs 272 B written + read per element of extra HBM traffic and removes 3 of the 4 evaluations of the
// element core (about 60 % of all instructions of the fused kernel).
// ---------------------------------------------------------------------------
constexpr int ELEM_REC = 34;  // doubles; 272 B = 17 x 16 B: an odd number of 16 B chunks, see patch_gather_kernel

// chunks 6..16 of a record (16 B each) -> the tangent fields of Core
template <bool GLOBAL>
__device__ __forceinline__ void unpack_tangent(double2 const* q, Core<double>& c) {
  auto ld = [&](int k) { return GLOBAL ? __ldg(q + k) : q[k]; };
#pragma unroll
  for (int k = 0; k < 3; ++k) { double2 const v = ld(6 + k); c.s[2 * k] = v.x; c.s[2 * k + 1] = v.y; }
  double2 v = ld(9); c.q[0] = v.x; c.q[1] = v.y;
  v = ld(10); c.q[2] = v.x; c.gwv = v.y;
  v = ld(11); c.A1v = v.x; c.Jpv = v.y;
  v = ld(12); c.upc = v.x; c.va = v.y;
  v = ld(13); c.tjv = v.x; c.ppc = v.y;
  v = ld(14); c.vb = v.x; c.gNs = v.y;
  v = ld(15); c.vgr = v.x; c.rb = v.y;
  v = ld(16); c.rc1 = v.x; c.tb3 = v.y;
}
// chunks 0..5 of a record -> w_n of the four nodes
template <bool GLOBAL>
__device__ __forceinline__ void unpack_w(double2 const* q, double wv[4][3]) {
  auto ld = [&](int k) { return GLOBAL ? __ldg(q + k) : q[k]; };
  double2 const v0 = ld(0), v1 = ld(1), v2 = ld(2), v3 = ld(3), v4 = ld(4), v5 = ld(5);
  wv[0][0] = v0.x; wv[0][1] = v0.y; wv[0][2] = v1.x;
  wv[1][0] = v1.y; wv[1][1] = v2.x; wv[1][2] = v2.y;
  wv[2][0] = v3.x; wv[2][1] = v3.y; wv[2][2] = v4.x;
  wv[3][0] = v4.y; wv[3][1] = v5.x; wv[3][2] = v5.y;
}

// Warp-cooperative Fp update of 32 consecutive elements e0 .. e0+nrec-1 (lane = element), plastic branch only:
// Fp = exp(dgam N) Fp_old (goal_J2.cpp:128-131); on the elastic branch the reference leaves Fp untouched (:135-136).
// Per-thread accesses of the 72 B Fp_old / Fp records touch 32 different 128 B lines per instruction (9 + 9
// instructions); staged through shared memory the warp reads its 2.3 KB of Fp_old with 5 coalesced 128-bit loads and
// writes the Fp halves of its state records (pairs 5-9, 80 B each) with 5 coalesced 128-bit stores.
// buf: >= WSAVE_DOUBLES doubles of shared memory that belong to the warp and are free (callers __syncwarp first).
constexpr int WSAVE_ST = 288, WSAVE_LD = 11;  // Fp_old block at [0, 288), Fp rows of 11 doubles (odd: conflict-free) behind it
constexpr int WSAVE_DOUBLES = WSAVE_ST + 32 * WSAVE_LD;
__device__ __forceinline__ void warp_save_Fp(KParams const& P, int e0, int nrec, int lane, int plastic, double const dN[6], double* buf) {
  unsigned const pmask = __ballot_sync(0xffffffffu, plastic != 0);
  if (!pmask) return;
  double const* src = P.fp_old + 9 * (int64_t)e0;  // 72 B * e0, e0 a multiple of 32: 16 B aligned
  int const tot = 9 * nrec;
  for (int g = 2 * lane; g < tot; g += 64) {
    if (g + 1 < tot) { double2 const v = __ldg(reinterpret_cast<double2 const*>(src + g)); buf[g] = v.x; buf[g + 1] = v.y; }
    else buf[g] = __ldg(src + g);
  }
  __syncwarp();
  if (plastic) {
    double Fpo[9], Fpn[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Fpo[k] = buf[9 * lane + k];
    plastic_update(dN, Fpo, Fpn);
    double* st = buf + WSAVE_ST + lane * WSAVE_LD;
#pragma unroll
    for (int k = 0; k < 9; ++k) st[k] = Fpn[k];
    st[9] = 0.0;
  }
  __syncwarp();
  double* dst = P.state_out + (int64_t)STATE_OUT * e0 + SO_FP;
  for (int g = lane; g < nrec * 5; g += 32) {
    int const r = g / 5, j = g - r * 5;
    if (!((pmask >> r) & 1u)) continue;
    double const* q = buf + WSAVE_ST + r * WSAVE_LD + 2 * j;
    *reinterpret_cast<double2*>(dst + STATE_OUT * r + 2 * j) = make_double2(q[0], q[1]);
  }
}

template <int MODEL, bool SAVE>
__global__ void __launch_bounds__(64, 8) elem_record_kernel(const __grid_constant__ KParams P, double* __restrict__ rec, int ne) {
  // records leave through shared memory so that a warp writes its 32 records (8.5 KB, contiguous) with
  // fully coalesced 128-bit stores instead of 17 stride-272 B stores per thread; the Fp update reuses the buffer
  __shared__ double srec[2][32 * ELEM_REC + 32];  // 64-thread blocks; row stride 35 doubles (odd): conflict-free column writes
  static_assert(32 * ELEM_REC + 32 >= WSAVE_DOUBLES, "state staging must fit the record buffer");
  int const e = blockIdx.x * blockDim.x + threadIdx.x;
  int const wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int const e0 = blockIdx.x * blockDim.x + wib * 32;  // first element of this warp
  int const nrec = min(32, ne - e0);
  double* mine = &srec[wib][lane * (ELEM_REC + 1)];
  int plastic = 0;
  double dN[6];
  if (SAVE && MODEL == MODEL_J2 && nrec > 0 && lane < 19) {  // the warp's Fp_old block is read last: have it in L2 by then
    char const* f = reinterpret_cast<char const*>(P.fp_old + 9 * (int64_t)e0) + 128 * lane;
    if (f < reinterpret_cast<char const*>(P.fp_old + 9 * (int64_t)(e0 + nrec))) asm volatile("prefetch.global.L2 [%0];" ::"l"(f));
  }
  if (e < ne) {
    int nd[4], b0[4], nb[4];
    Material const* matp;
    Core<double> c;
    int const rc = load_and_update<MODEL, SAVE>(P, e, true, nd, b0, nb, matp, c);
    if (rc != ERR_NONE) {
      report_error(P.err, rc, e);
#pragma unroll
      for (int k = 0; k < ELEM_REC; ++k) mine[k] = 0.0;
    } else {
      plastic = c.plastic;
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 3; ++k) mine[3 * q + k] = c.w[q][k];
#pragma unroll
      for (int k = 0; k < 6; ++k) mine[12 + k] = c.s[k];
      mine[18] = c.q[0]; mine[19] = c.q[1]; mine[20] = c.q[2]; mine[21] = c.gwv; mine[22] = c.A1v; mine[23] = c.Jpv;
      mine[24] = c.upc; mine[25] = c.va; mine[26] = c.tjv; mine[27] = c.ppc; mine[28] = c.vb; mine[29] = c.gNs;
      mine[30] = c.vgr; mine[31] = c.rb; mine[32] = c.rc1; mine[33] = c.tb3;
      if (SAVE && MODEL == MODEL_J2 && plastic) {
#pragma unroll
        for (int k = 0; k < 6; ++k) dN[k] = c.dN[k];
      }
    }
  }
  __syncwarp();
  if (nrec > 0) {
    double* dst = rec + (int64_t)ELEM_REC * e0;
    int const total = nrec * ELEM_REC;  // doubles, contiguous in global memory
    for (int g = 2 * lane; g < total; g += 64) {
      int const r = g / ELEM_REC, k = g - r * ELEM_REC;  // ELEM_REC is even: the pair stays inside one record
      double const* src = &srec[wib][r * (ELEM_REC + 1) + k];
      *reinterpret_cast<double2*>(dst + g) = make_double2(src[0], src[1]);
    }
    if (SAVE && MODEL == MODEL_J2) {
      __syncwarp();  // the record buffer is free again
      warp_save_Fp(P, e0, nrec, lane, plastic, dN, srec[wib]);
    }
  }
  if (MODEL == MODEL_J2) {
    unsigned const b = __ballot_sync(0xffffffffu, plastic != 0);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(P.plastic, (unsigned long long)__popc(b));
  }
}


template <bool TRANSPOSE, int MINB>
__global__ void __launch_bounds__(128, MINB) row_fold_kernel(const __grid_constant__ KParams P, double const* __restrict__ rec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int const wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int const half = lane >> 4, t16 = lane & 15;
  double* stg = reinterpret_cast<double*>(smem_raw + row_owner_smem_per_warp(P.max_nblk) * wib);
  double* wr = stg + 16 * STG_LD + lane;
  double* acc = stg + 16 * STG_LD + 24 * WR_LD;  // [2][16*max_nblk], one copy per half-warp
  int const accld = 16 * P.max_nblk;

  // Persistent warps: warp gw handles nodes gw, gw + W, gw + 2W, ...  The node -> incidence -> record chain is
  // three dependent global loads; it is software-pipelined across nodes: while node a is processed the
  // incidences of node a + 2W are being loaded and the records of node a + W are being pulled into L2.
  int const W = gridDim.x * (blockDim.x >> 5);
  int slot = blockIdx.x * (blockDim.x >> 5) + wib;  // position in the visiting order
  int a = 0, ap = 0;
  uint32_t o0 = 0, o1 = 0, p0 = 0, p1 = 0;
  int2 ad = make_int2(0, 0), adp = make_int2(0, 0);
  if (slot < P.nn) {
    a = __ldg(P.node_order + slot);
    o0 = __ldg(P.adj_off + a); o1 = __ldg(P.adj_off + a + 1);
    if (o0 + lane < o1) ad = __ldg(P.adj + o0 + lane);
  }
  if (slot + W < P.nn) {
    ap = __ldg(P.node_order + slot + W);
    p0 = __ldg(P.adj_off + ap); p1 = __ldg(P.adj_off + ap + 1);
    if (p0 + lane < p1) adp = __ldg(P.adj + p0 + lane);
  }
  for (; slot < P.nn; slot += W) {
    // stage 1 of the pipeline: records of the next node -> L2 (3 lines cover the 272 B record)
    if (p0 + lane < p1) {
      char const* r = reinterpret_cast<char const*>(rec + (int64_t)ELEM_REC * (adp.x >> 2));
#pragma unroll
      for (int k = 0; k < 3; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(r + 128 * k));
    }
    // stage 0: incidences of the node after next
    uint32_t q0 = 0, q1 = 0;
    int aq = 0;
    int2 adq = make_int2(0, 0);
    if (slot + 2 * W < P.nn) {
      aq = __ldg(P.node_order + slot + 2 * W);
      q0 = __ldg(P.adj_off + aq); q1 = __ldg(P.adj_off + aq + 1);
      if (q0 + lane < q1) adq = __ldg(P.adj + q0 + lane);
    }
    if ((int)(o1 - o0) >= P.e0) {  // nodes below the threshold are left to row_fold_sorted_kernel
      int blk0a, nblka;
      {
        double2 const d3 = __ldg(reinterpret_cast<double2 const*>(P.nodes + a) + 3);
        blk0a = __double2loint(d3.y);
        nblka = __double2hiint(d3.y);
      }
      int const nent = 16 * nblka;
      for (int g = lane; g < nent; g += 32) { acc[g] = 0.0; acc[accld + g] = 0.0; }
      double racc[4] = {0.0, 0.0, 0.0, 0.0};

      for (uint32_t r0 = o0; r0 < o1; r0 += 32) {
        int const nact = min(32, (int)(o1 - r0));
        bool const active = lane < nact;
        int n = 0;
        uint32_t jpack = 0;
        Core<double> c;  // only the tangent fields are filled
        if (active) {
          int2 const adr = r0 == o0 ? ad : __ldg(P.adj + r0 + lane);
          int const e = adr.x >> 2;
          n = adr.x & 3; jpack = (uint32_t)adr.y;
          double2 const* q = reinterpret_cast<double2 const*>(rec + (int64_t)ELEM_REC * e);
          unpack_tangent<true>(q, c);
          double wv[4][3];
          unpack_w<true>(q, wv);
#pragma unroll
          for (int n4 = 0; n4 < 4; ++n4) {  // w, r -> shared memory (this lane's column)
            double sw[3], r3[3];
            node_r(c, wv[n4], sw, r3);
#pragma unroll
            for (int k = 0; k < 3; ++k) { wr[(6 * n4 + k) * WR_LD] = wv[n4][k]; wr[(6 * n4 + 3 + k) * WR_LD] = r3[k]; }
          }
        }
        double wn[3], rn3[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          wn[k] = wr[(6 * n + k) * WR_LD];
          rn3[k] = wr[(6 * n + 3 + k) * WR_LD];
        }
        RowNode<double> rown;
        ColNode<double> coln;
        if (active) {
          double r4[4];
          element_residual_row(c, wn, r4);
          racc[0] += r4[0]; racc[1] += r4[1]; racc[2] += r4[2]; racc[3] += r4[3];
          if (!TRANSPOSE) row_node(c, wn, rown);
          else column_node(c, wn, rn3, coln);
        }
#pragma unroll 1
        for (int m = 0; m < 4; ++m) {
          uint32_t jm = 0;
          if (active) {
            double wm[3], rm[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              wm[k] = wr[(6 * m + k) * WR_LD];
              rm[k] = wr[(6 * m + 3 + k) * WR_LD];
            }
            double blk[16];
            if (!TRANSPOSE) {
              ColNode<double> cnm;
              column_node(c, wm, rm, cnm);
              jacobian_block(c, rown, cnm, blk);
#pragma unroll
              for (int t = 0; t < 16; ++t) stg[t * STG_LD + lane] = blk[t];
            } else {
              RowNode<double> rnm;
              row_node(c, wm, rnm);
              jacobian_block(c, rnm, coln, blk);
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) stg[(4 * i + k) * STG_LD + lane] = blk[4 * k + i];
            }
            jm = (jpack >> (8 * m)) & 0xffu;
          }
          __syncwarp();
          {
            int const split = (nact + 1) >> 1;
            int const lbase = half ? split : 0;
            int const lend = half ? nact : split;
            double* my = acc + half * accld + t16;
#pragma unroll 4
            for (int it = 0; it < split; ++it) {
              int const l = lbase + it;
              uint32_t const j = __shfl_sync(0xffffffffu, jm, l & 31);
              if (l < lend) my[16 * j] += stg[t16 * STG_LD + l];
            }
          }
          __syncwarp();
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double v = racc[i];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        racc[i] = v;
      }
      if (lane == 0) {
        double2* q = reinterpret_cast<double2*>(P.R + 4 * (int64_t)a);
        q[0] = make_double2(racc[0], racc[1]);
        q[1] = make_double2(racc[2], racc[3]);
      }
      double* out = P.values + 16 * (int64_t)blk0a;
      int const rl = 4 * nblka;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        for (int cidx = lane; cidx < rl; cidx += 32) {
          int const g = 16 * (cidx >> 2) + 4 * i + (cidx & 3);
          out[i * rl + cidx] = acc[g] + acc[accld + g];
        }
      __syncwarp();
    }
    a = ap; o0 = p0; o1 = p1; ad = adp;
    ap = aq; p0 = q0; p1 = q1; adp = adq;
  }
}

// Stage B, sorted fold (nodes with at most 32 incidences -- every node of a Kuhn mesh).  All four blocks of
// every incidence are staged (stg4[64][33] per warp); then the warp walks the node's precomputed schedule
// (KParams::fold_ord, gx_setup.cpp): the staged blocks grouped by target block, two per word.  Half-warp 0
// takes the first block of a word, half-warp 1 the second; lane t accumulates entry t in a register.  At the
// end of a group the two halves are joined by one shuffle and half-warp 0 stores the finished 4x4 block
// straight into the CRS rows.  Control flow is warp-uniform; no accumulator array, no zeroing, no second pass.
GX_HD int fold_row_stride(int max_deg) {  // staging row stride: one pad column past the widest node, odd, <= 33
  int const ld = (std_min_int(max_deg, 32) + 1) | 1;
  return ld;
}
GX_HD size_t row_fold_smem_per_warp(int max_nblk, int fold_ld, bool need_generic) {
  size_t const sorted = (size_t)64 * fold_ld * sizeof(double);
  size_t const generic = need_generic ? row_owner_smem_per_warp(max_nblk) : 0;
  return sorted > generic ? sorted : generic;
}

template <bool TRANSPOSE, int MINB>
__global__ void __launch_bounds__(128, MINB) row_fold_sorted_kernel(const __grid_constant__ KParams P, double const* __restrict__ rec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int const wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int const half = lane >> 4, t16 = lane & 15;
  int const ld = P.fold_ld;
  double* stg = reinterpret_cast<double*>(smem_raw + (size_t)64 * ld * sizeof(double) * wib);
  stg[lane * ld + ld - 1] = 0.0;         // pad column: the schedule's "no block" slot reads as zero
  stg[(32 + lane) * ld + ld - 1] = 0.0;

  // persistent warps with the same three-deep software pipeline as row_fold_kernel
  int const W = gridDim.x * (blockDim.x >> 5);
  int slot = blockIdx.x * (blockDim.x >> 5) + wib;  // position in the visiting order
  int a = 0, ap = 0;
  uint32_t o0 = 0, o1 = 0, p0 = 0, p1 = 0;
  int2 ad = make_int2(0, 0), adp = make_int2(0, 0);
  if (slot < P.nn) {
    a = __ldg(P.node_order + slot);
    o0 = __ldg(P.adj_off + a); o1 = __ldg(P.adj_off + a + 1);
    if (o0 + lane < o1) ad = __ldg(P.adj + o0 + lane);
  }
  if (slot + W < P.nn) {
    ap = __ldg(P.node_order + slot + W);
    p0 = __ldg(P.adj_off + ap); p1 = __ldg(P.adj_off + ap + 1);
    if (p0 + lane < p1) adp = __ldg(P.adj + p0 + lane);
  }
  for (; slot < P.nn; slot += W) {
    if (p0 + lane < p1) {
      char const* r = reinterpret_cast<char const*>(rec + (int64_t)ELEM_REC * (adp.x >> 2));
#pragma unroll
      for (int k = 0; k < 3; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(r + 128 * k));
    }
    if (lane < 4 && p1 - p0 <= 32) {  // next node's schedule (at most 132 words) -> L2
      asm volatile("prefetch.global.L2 [%0];" ::"l"(P.fold_ord + 4 * (int64_t)p0 + 8 * (int64_t)ap + 32 * lane));
    }
    uint32_t q0 = 0, q1 = 0;
    int aq = 0;
    int2 adq = make_int2(0, 0);
    if (slot + 2 * W < P.nn) {
      aq = __ldg(P.node_order + slot + 2 * W);
      q0 = __ldg(P.adj_off + aq); q1 = __ldg(P.adj_off + aq + 1);
      if (q0 + lane < q1) adq = __ldg(P.adj + q0 + lane);
    }
    int const deg = (int)(o1 - o0);
    if (deg <= 32) {  // larger nodes are handled by row_fold_kernel
      int blk0a, nblka;
      {
        double2 const d3 = __ldg(reinterpret_cast<double2 const*>(P.nodes + a) + 3);
        blk0a = __double2loint(d3.y);
        nblka = __double2hiint(d3.y);
      }
      uint4 const* ord = reinterpret_cast<uint4 const*>(P.fold_ord + 4 * (int64_t)o0 + 8 * (int64_t)a);
      int nw = 0;
      uint32_t const nop = (uint32_t)(ld - 1) | ((uint32_t)(ld - 1) << 11);
      uint4 wq = make_uint4(nop, nop, nop, nop);
      if (deg > 0) { nw = (int)__ldg(reinterpret_cast<uint32_t const*>(ord)); wq = __ldg(ord + 1); }  // issued early
      double r4[4] = {0.0, 0.0, 0.0, 0.0};
      if (lane < deg) {
        int const e = ad.x >> 2, n = ad.x & 3;
        double2 const* q = reinterpret_cast<double2 const*>(rec + (int64_t)ELEM_REC * e);
        Core<double> c;  // only the tangent fields are filled
        double wv[4][3], rv[4][3];
        unpack_tangent<true>(q, c);
        unpack_w<true>(q, wv);
#pragma unroll
        for (int n4 = 0; n4 < 4; ++n4) { double sw[3]; node_r(c, wv[n4], sw, rv[n4]); }
        double wn[3], rn3[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          wn[k] = n == 0 ? wv[0][k] : n == 1 ? wv[1][k] : n == 2 ? wv[2][k] : wv[3][k];
          rn3[k] = n == 0 ? rv[0][k] : n == 1 ? rv[1][k] : n == 2 ? rv[2][k] : rv[3][k];
        }
        element_residual_row(c, wn, r4);
        RowNode<double> rown;
        ColNode<double> coln;
        if (!TRANSPOSE) row_node(c, wn, rown);
        else column_node(c, wn, rn3, coln);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          double blk[16];
          double* dst = stg + (m * 16) * ld + lane;
          if (!TRANSPOSE) {
            ColNode<double> cnm;
            column_node(c, wv[m], rv[m], cnm);
            jacobian_block(c, rown, cnm, blk);
#pragma unroll
            for (int t = 0; t < 16; ++t) dst[t * ld] = blk[t];
          } else {
            RowNode<double> rnm;
            row_node(c, wv[m], rnm);
            jacobian_block(c, rnm, coln, blk);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int k = 0; k < 4; ++k) dst[(4 * i + k) * ld] = blk[4 * k + i];
          }
        }
      }
      __syncwarp();
      // ---- R rows of node a: fixed butterfly over the lanes
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double v = r4[i];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        r4[i] = v;
      }
      if (lane == 0) {
        double2* q = reinterpret_cast<double2*>(P.R + 4 * (int64_t)a);
        q[0] = make_double2(r4[0], r4[1]);
        q[1] = make_double2(r4[2], r4[3]);
      }
      // ---- sorted fold, warp-uniform control flow
      double const* src = stg + t16 * ld;
      double* out = P.values + 16 * (int64_t)blk0a + (int64_t)(t16 >> 2) * (4 * nblka) + (t16 & 3);
      int const sh = half ? 11 : 0;
      double acc = 0.0;
      for (int i = 0; i < nw; i += 4) {
        uint4 const w = wq;
        if (i + 4 < nw) wq = __ldg(ord + 2 + (i >> 2));  // next group, in flight while this one is folded
        // the four staged values first (independent shared loads), then the dependent adds
        double const v0 = src[(w.x >> sh) & 0x7ffu], v1 = src[(w.y >> sh) & 0x7ffu];
        double const v2 = src[(w.z >> sh) & 0x7ffu], v3 = src[(w.w >> sh) & 0x7ffu];
        uint32_t const ws[4] = {w.x, w.y, w.z, w.w};
        double const vs[4] = {v0, v1, v2, v3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc += vs[k];
          if (ws[k] & 0x40000000u) {  // same word in every lane: uniform branch
            double const tot = acc + __shfl_xor_sync(0xffffffffu, acc, 16);
            if (half == 0) out[4 * ((ws[k] >> 22) & 0xffu)] = tot;
            acc = 0.0;
          }
        }
      }
      // phantom blocks (columns that live only on other parts) receive remote contributions later: start at zero
      if (P.nblk_g)
        for (int j = __ldg(P.nblk_g + a) + half; j < nblka; j += 2) out[4 * j] = 0.0;
      __syncwarp();
    }
    a = ap; o0 = p0; o1 = p1; ad = adp;
    ap = aq; p0 = q0; p1 = q1; adp = adq;
  }
}

// ---------------------------------------------------------------------------
// Schedule (1c), patch gather (option kernel = 3): stage A as above, then one thread block per patch of the
// precomputed patch schedule (build_patch_schedule, gx_setup.cpp).  The block stages the records of the patch's
// elements in shared memory with bulk asynchronous copies (cp.async.bulk, one per record: every record crosses
// the L1 data pipe once per patch instead of once per incident node and lane), then every thread runs one work item: up to 8 contributions to
// one 4x4 block, rebuilt from the staged records and accumulated in registers.  Blocks with more contributions
// are finished by their primary item from the secondaries' partial sums (fixed order).  Every block of the patch's
// rows, and the rows' residual entries, are written exactly once.
// ---------------------------------------------------------------------------
constexpr int PATCH_REC_LD = ELEM_REC;  // staged records keep their global stride: 272 B = 17 x 16 B (odd), so the bank
                                        // group of a record's chunk k is (slot + k) mod 8
GX_HD size_t patch_smem_bytes() { return ((size_t)PATCH_RECS * PATCH_REC_LD + (size_t)PATCH_PARTS * 20) * sizeof(double); }

// One work item: up to PATCH_ITEM_LEN contributions to one 4x4 block (and, for diagonal items, to the node's residual
// entries), rebuilt from the records staged at srec and accumulated in registers.
template <bool TRANSPOSE>
__device__ __forceinline__ void patch_item(double const* srec, uint4 const it, bool const diag, double acc[16], double r4[4]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.0;
  r4[0] = r4[1] = r4[2] = r4[3] = 0.0;
  uint64_t elo = (uint64_t)it.x | ((uint64_t)it.y << 32), ehi = (uint64_t)it.z | ((uint64_t)it.w << 32);
#pragma unroll 1
  for (int k = 0; k < PATCH_ITEM_LEN; ++k) {
    uint32_t const ent = (uint32_t)elo & 0xffffu;
    elo = (elo >> 16) | (ehi << 48); ehi >>= 16;
    if (!(ent & 0x8000u)) {  // an empty round of this item (bank-conflict avoidance), or the end of its list
      if ((elo | ehi) == 0) break;
      continue;
    }
    int const slot = (int)(ent & 0xffu), n = (int)((ent >> 10) & 3u), m = (int)((ent >> 8) & 3u);
    double const* rp = srec + slot * PATCH_REC_LD;
    Core<double> c;  // only the tangent fields are filled
    unpack_tangent<false>(reinterpret_cast<double2 const*>(rp), c);
    // Every lane reads the same 17 chunks of its record, so a quarter-warp whose records sit in 8 different bank
    // groups (the host schedule sees to that) reads without conflicts; the two nodes the block needs are then
    // selected in registers.  (Loading only w_n / w_m would put the chunk offset, and with it the bank group,
    // at the mercy of the local node numbers.)
    double wv[4][3];
    unpack_w<false>(reinterpret_cast<double2 const*>(rp), wv);
    // row node = the node of this block row in the primal operator; roles swap for the transpose
    int const nr = TRANSPOSE ? m : n, nc = TRANSPOSE ? n : m;
    double wr[3], wc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double const a01 = (nr & 1) ? wv[1][k] : wv[0][k], a23 = (nr & 1) ? wv[3][k] : wv[2][k];
      wr[k] = (nr & 2) ? a23 : a01;
      double const b01 = (nc & 1) ? wv[1][k] : wv[0][k], b23 = (nc & 1) ? wv[3][k] : wv[2][k];
      wc[k] = (nc & 2) ? b23 : b01;
    }
    RowNode<double> rown;
    ColNode<double> coln;
    row_node(c, wr, rown);
    column_node_w(c, wc, coln);
    jacobian_block_add<TRANSPOSE>(c, rown, coln, acc);
    if (diag) {  // n == m: the residual entries of the node
      double t4[4];
      element_residual_row(c, wr, t4);
      r4[0] += t4[0]; r4[1] += t4[1]; r4[2] += t4[2]; r4[3] += t4[3];
    }
  }
}

template <bool TRANSPOSE>
__global__ void __launch_bounds__(PATCH_THREADS, PATCH_MINB) patch_gather_kernel(const __grid_constant__ KParams P, double const* __restrict__ rec,
                                                                        uint32_t const* __restrict__ sched) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t mbar;
  double* srec = reinterpret_cast<double*>(smem_raw);
  int const tid = threadIdx.x;
  uint32_t const* w = sched + (size_t)blockIdx.x * PATCH_WORDS;
  int const n_recs = (int)__ldg(w);
  int const n_runs = (int)__ldg(w + 2);
  uint2 const my_run = __ldg(reinterpret_cast<uint2 const*>(w + 4 + PATCH_RECS + 8 * PATCH_THREADS) + min(tid, PATCH_RECS - 1));
  uint4 const it = __ldg(reinterpret_cast<uint4 const*>(w + 4 + PATCH_RECS) + tid);
  uint4 const ot = __ldg(reinterpret_cast<uint4 const*>(w + 4 + PATCH_RECS + 4 * PATCH_THREADS) + tid);
  // Record staging: thread r issues one bulk asynchronous copy (global -> shared) for run r of the schedule -- a run is
  // a number of consecutive elements' records (272 B each, contiguous in global memory) that go to consecutive slots;
  // the copies report their bytes to an mbarrier that the whole block then waits on.
  static_assert(PATCH_RECS <= PATCH_THREADS, "one thread per run");
  uint32_t const mb = (uint32_t)__cvta_generic_to_shared(&mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(n_recs * (ELEM_REC * 8)) : "memory");
  if (tid < n_runs) {
    uint32_t const sl = my_run.y & 0xffu, len = my_run.y >> 8;
    uint32_t const dst = (uint32_t)__cvta_generic_to_shared(srec) + sl * (uint32_t)(PATCH_REC_LD * 8);
    double const* src = rec + (int64_t)ELEM_REC * (int64_t)my_run.x;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(len * (uint32_t)(ELEM_REC * 8)), "r"(mb)
                 : "memory");
  }
  {
    uint32_t done;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb) : "memory");
    } while (!done);
  }
  int const kind = (int)(ot.z >> 30);
  bool const diag = (ot.w & 0x80000000u) != 0;
  double acc[16], r4[4];
  patch_item<TRANSPOSE>(srec, it, diag, acc, r4);
  // Finish.  Items without secondaries write their block and leave; only the few items that exchange partial sums
  // (diagonal blocks, edges of high valence: the longest items, i.e. the first warp) meet at the barrier.
  int const part = (int)((ot.z >> 16) & 0xffu);
  int const nsec = (int)((ot.z >> 24) & 0x3fu);
  double* spart = srec + (size_t)PATCH_RECS * PATCH_REC_LD;  // [PATCH_PARTS][20]
  auto write_out = [&]() {
    int64_t const voff = (int64_t)(((uint64_t)ot.y << 32) | (uint64_t)ot.x);
    int const rl = (int)(ot.z & 0xffffu);
    double* out = P.values + voff;
#pragma unroll
    for (int i = 0; i < 4; ++i) stg256(out + (int64_t)i * rl, acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);  // one 32 B sector each
    if (diag) stg256(P.R + 4 * (int64_t)(ot.w & 0x7fffffffu), r4[0], r4[1], r4[2], r4[3]);
  };
  if (kind == 0) return;
  if (kind == 1 && nsec == 0) { write_out(); return; }
  if (kind == 2) {
    double2* d = reinterpret_cast<double2*>(spart + 20 * part);
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = make_double2(acc[2 * k], acc[2 * k + 1]);
    d[8] = make_double2(r4[0], r4[1]);
    d[9] = make_double2(r4[2], r4[3]);
  }
  __syncthreads();  // the threads that are still here
  if (kind == 1) {
    for (int s2 = 0; s2 < nsec; ++s2) {
      double2 const* d = reinterpret_cast<double2 const*>(spart + 20 * (part + s2));
#pragma unroll
      for (int k = 0; k < 8; ++k) { double2 const v = d[k]; acc[2 * k] += v.x; acc[2 * k + 1] += v.y; }
      if (diag) {
        double2 v = d[8]; r4[0] += v.x; r4[1] += v.y;
        v = d[9]; r4[2] += v.x; r4[3] += v.y;
      }
    }
    write_out();
  }
}


// ---------------------------------------------------------------------------
// Residual and error-localisation passes, gather form (default): no colouring, no zeroing, no atomics.
//   elem_residual_kernel : one thread per element -> its 16 residual entries rvec[e][n][4] (128 B, one line),
//                          state save as in stage A of the Jacobian pass.  ERROR selects the adjoint-weighted
//                          residual of the error chain (goal_mechanics.cpp:169-218).
//   node_gather_kernel   : 8 lanes per node sum rvec[e][n] over the node's incidences (fixed order) and write
//                          R[4a .. 4a+3] once.
// ---------------------------------------------------------------------------
template <int MODEL, bool SAVE, bool ERROR>
__global__ void __launch_bounds__(128, 4) elem_residual_kernel(const __grid_constant__ KParams P, double* __restrict__ rvec, int ne) {
  // the element residual lines (128 B each) and the state leave through shared memory: coalesced 128-bit stores
  // instead of 32 different 128 B lines per store instruction
  __shared__ double sbuf[4][WSAVE_DOUBLES];
  static_assert(WSAVE_DOUBLES >= 32 * 17, "residual staging must fit");
  int const e = blockIdx.x * blockDim.x + threadIdx.x;
  int const wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int const e0 = blockIdx.x * blockDim.x + wib * 32;
  int const nrec = min(32, ne - e0);
  double* buf = sbuf[wib];
  int plastic = 0;
  double dN[6];
  if (SAVE && MODEL == MODEL_J2 && nrec > 0 && lane < 19) {
    char const* f = reinterpret_cast<char const*>(P.fp_old + 9 * (int64_t)e0) + 128 * lane;
    if (f < reinterpret_cast<char const*>(P.fp_old + 9 * (int64_t)(e0 + nrec))) asm volatile("prefetch.global.L2 [%0];" ::"l"(f));
  }
  if (e < ne) {
    int nd[4], b0[4], nb[4];
    Material const* matp;
    Core<double> c;
    int const rc = load_and_update<MODEL, SAVE>(P, e, true, nd, b0, nb, matp, c);
    double ru[12], rp[4];
    if (rc != ERR_NONE) {
      report_error(P.err, rc, e);
#pragma unroll
      for (int k = 0; k < 12; ++k) ru[k] = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) rp[k] = 0.0;
    } else {
      plastic = c.plastic;
      if (ERROR) {
        double zu[4][3], zp[4], zpc[4];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          double2 const* q = reinterpret_cast<double2 const*>(P.z + nd[n]);
          double2 const a = ldg(q), b = ldg(q + 1), d = ldg(q + 2);
          zu[n][0] = a.x; zu[n][1] = a.y; zu[n][2] = b.x; zp[n] = b.y; zpc[n] = d.x;
        }
        element_error_residual(c, zu, zp, zpc, ru, rp);
      } else {
        element_residual(c, ru, rp);
      }
      if (SAVE && MODEL == MODEL_J2 && plastic) {
#pragma unroll
        for (int k = 0; k < 6; ++k) dN[k] = c.dN[k];
      }
    }
    double* o = buf + 17 * lane;  // row stride 17 (odd): conflict-free
#pragma unroll
    for (int n = 0; n < 4; ++n) { o[4 * n] = ru[3 * n]; o[4 * n + 1] = ru[3 * n + 1]; o[4 * n + 2] = ru[3 * n + 2]; o[4 * n + 3] = rp[n]; }
  }
  __syncwarp();
  if (nrec > 0) {
    double* dst = rvec + 16 * (int64_t)e0;
    for (int g = lane; g < nrec * 8; g += 32) {
      double const* q = buf + 17 * (g >> 3) + 2 * (g & 7);
      *reinterpret_cast<double2*>(dst + 2 * g) = make_double2(q[0], q[1]);
    }
    if (SAVE && MODEL == MODEL_J2) {
      __syncwarp();
      warp_save_Fp(P, e0, nrec, lane, plastic, dN, buf);
    }
  }
  if (MODEL == MODEL_J2) {
    unsigned const b = __ballot_sync(0xffffffffu, plastic != 0);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(P.plastic, (unsigned long long)__popc(b));
  }
}

__global__ void __launch_bounds__(256) node_gather_kernel(const __grid_constant__ KParams P, double const* __restrict__ rvec) {
  // 8 lanes per node: lane j sums the incidences j, j + 8, j + 16, ... (ascending element order), then the 8 partial
  // sums are added in a fixed xor tree -- the same order on every run, so the result is bit-reproducible.  Compared
  // with one thread walking all (about 24) incidences this puts 8 times as many independent 32 B loads in flight.
  int const t = blockIdx.x * blockDim.x + threadIdx.x;
  int const a = t >> 3, sub = t & 7;
  double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
  if (a < P.nn) {
    uint32_t const o0 = __ldg(P.adj_off + a), o1 = __ldg(P.adj_off + a + 1);
    for (uint32_t k = o0 + sub; k < o1; k += 8) {
      int const en = __ldg(P.adj + k).x;  // e*4 + n: rvec is [e][n][4]
      double v0, v1, v2, v3;
      ldg256(rvec + 4 * (int64_t)en, v0, v1, v2, v3);
      r0 += v0; r1 += v1; r2 += v2; r3 += v3;
    }
  }
#pragma unroll
  for (int w = 4; w > 0; w >>= 1) {
    r0 += __shfl_xor_sync(0xffffffffu, r0, w);
    r1 += __shfl_xor_sync(0xffffffffu, r1, w);
    r2 += __shfl_xor_sync(0xffffffffu, r2, w);
    r3 += __shfl_xor_sync(0xffffffffu, r3, w);
  }
  if (a < P.nn && sub == 0) stg256(P.R + 4 * (int64_t)a, r0, r1, r2, r3);
}
// ---------------------------------------------------------------------------
// Functionals (Mechanics::build_functional, goal_mechanics.cpp:149-167; QoI<T> goal_qoi.cpp:21-82): one thread per
// element evaluates the QoI evaluator behind the save=false chain and writes the element value ev[e] and, when
// rvec != nullptr, d elem_value / d dof as one 128 B line rvec[e][n][4] -- node_gather_kernel then sums them per
// node into dMdu exactly like the residual (QoI<FADT>::scatter, goal_qoi.cpp:63-76), no atomics.
//   type: GX_QOI_AVG_DISP / _SUBDOMAIN (goal_avg_disp.cpp:17-21, goal_avg_disp_subdomain.cpp:37-53),
//         GX_QOI_AVG_VM (goal_avg_vm.cpp:43-61), GX_QOI_KS_VM (goal_ks_vm.cpp:89-99; ks = max, scale, rho)
// ---------------------------------------------------------------------------
struct QoiParams { int type, es_idx; double ks_max, ks_scale, rho; };
enum { QOI_AVG_DISP = 0, QOI_AVG_DISP_SUBDOMAIN = 1, QOI_AVG_VM = 2, QOI_KS_VM = 3 };

template <int MODEL>
__global__ void __launch_bounds__(128) elem_qoi_kernel(const __grid_constant__ KParams P, QoiParams const Q, double* __restrict__ rvec,
                                                       double* __restrict__ ev, int ne) {
  int const e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  double d[4][3], val = 0.0;
#pragma unroll
  for (int n = 0; n < 4; ++n) d[n][0] = d[n][1] = d[n][2] = 0.0;
  bool const in_set = Q.type == QOI_AVG_DISP || Q.type == QOI_KS_VM || (P.eset ? P.eset[e] : 0) == Q.es_idx;
  if (in_set) {
    int nd[4], b0[4], nb[4];
    Material const* matp;
    Core<double> c;
    // the whole chain runs in the reference too (Functional builds build_resid<T> first, goal_functional.cpp:41-45),
    // so an inverted element / deformation is reported here as well
    int const rc = load_and_update<MODEL, false>(P, e, false, nd, b0, nb, matp, c);
    if (rc != ERR_NONE) {
      report_error(P.err, rc, e);
    } else if (Q.type == QOI_AVG_DISP || Q.type == QOI_AVG_DISP_SUBDOMAIN) {
      double us = 0.0;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        double2 const* q = reinterpret_cast<double2 const*>(P.nodes + nd[n]);
        double2 const d1 = ldg(q + 1), d2 = ldg(q + 2);
        us += 0.25 * (d1.y + d2.x + d2.y);
      }
      val = us * c.vol * (1.0 / 3.0);
      double const g = 0.25 * c.vol * (1.0 / 3.0);
#pragma unroll
      for (int n = 0; n < 4; ++n) d[n][0] = d[n][1] = d[n][2] = g;
    } else {
      double const vm = element_von_mises(c, d);
      double f = 1.0;  // d elem_value / d vm, per unit volume
      if (Q.type == QOI_KS_VM) {
        double const ex = exp(Q.rho * (vm - Q.ks_max));
        val = (1.0 / (Q.rho * Q.ks_scale)) * ex * c.vol;
        f = ex / Q.ks_scale;
      } else {
        val = vm * c.vol;
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) { d[n][0] *= f; d[n][1] *= f; d[n][2] *= f; }
    }
  }
  ev[e] = val;
  if (rvec) {
    double2* o = reinterpret_cast<double2*>(rvec + 16 * (int64_t)e);
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      o[2 * n] = make_double2(d[n][0], d[n][1]);
      o[2 * n + 1] = make_double2(d[n][2], 0.0);
    }
  }
}

// KSVM<T>::pre_process (goal_ks_vm.cpp:36-87) over the saved "sigma" state: pass 0 = max vm per block,
// pass 1 = sum of exp(rho (vm - max)) w dv per block (fixed-shape tree: deterministic)
__global__ void __launch_bounds__(256) ks_vm_partial_kernel(double* partial, double const* state_out, NodeRec const* nodes, int4 const* conn,
                                                            int ne, int pass, double ks_max, double rho) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
    double t[9];
    for (int k = 0; k < 9; ++k) t[k] = state_out[(int64_t)STATE_OUT * e + k];
    double const vm = von_mises9(t);
    if (pass == 0) {
      s = fmax(s, vm);
    } else {
      int4 const cn = conn[e];
      int const nd[4] = {cn.x, cn.y, cn.z, cn.w};
      double x[4][3];
      for (int n = 0; n < 4; ++n) { x[n][0] = nodes[nd[n]].x[0]; x[n][1] = nodes[nd[n]].x[1]; x[n][2] = nodes[nd[n]].x[2]; }
      double e1[3], e2[3], e3[3], c23[3];
      for (int j = 0; j < 3; ++j) { e1[j] = x[1][j] - x[0][j]; e2[j] = x[2][j] - x[0][j]; e3[j] = x[3][j] - x[0][j]; }
      cross3(e2, e3, c23);
      s += exp(rho * (vm - ks_max)) * (dot3(e1, c23) * (1.0 / 6.0));
    }
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] = pass == 0 ? fmax(sh[threadIdx.x], sh[threadIdx.x + w]) : sh[threadIdx.x] + sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
// sum of x[0..n) in a fixed-shape tree (block partials; the caller adds them in index order)
__global__ void __launch_bounds__(256) sum_partial_kernel(double* partial, double const* x, int64_t n) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += x[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
#endif


}  // namespace gx
