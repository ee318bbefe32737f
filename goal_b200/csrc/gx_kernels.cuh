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
//   adj       int2[4 Ne]   node -> (element, local node) incidences, grouped by node (adj_off[Nn+1]): work list of the
//                          element-line gather (node_gather_kernel) and of the host-side schedule builders
//   state_in  double[Ne][10]  Fp_old[9], eqps_old                             80 B record (5 LDG.128); Cp^{-1} is
//                             formed in registers (cp_inverse): the old state is read once per element and pass
//   state_out double[Ne][20]  sigma[9], eqps, Fp[9], pad                     160 B record
//   R         double[4 Nn] ghost layout;   values  double[nnz]  CRS order of gx_graph
//
// Schedules, all free of atomics on the data path and bit-reproducible:
//  (1) Jacobian pass (default): stage A elem_record_kernel (one thread per element: stress update, state save, the
//      element's 38-double tangent record, tangent_record.cuh) + stage B patch_pair_kernel (one thread block per patch
//      of nodes: the patch's records staged in shared memory by bulk copies, one work item per thread -- an edge's two
//      mirror blocks or a diagonal block with the node's residual entries -- accumulated in registers and written once).
//      No zeroing pass, no read-modify-write traffic.
//  (2) residual / error-localisation passes (default): block-reduced -- elem_residual_block_kernel keeps the residual
//      lines of 128 consecutive elements in shared memory, sums them per node there and writes R or one 32 B partial
//      sum per node shared with other blocks; node_partial_sum_kernel adds those.  The element-line form
//      (elem_residual_kernel + node_gather_kernel: a 128 B line per element through global memory, 8 lanes per node
//      sum its incidences) remains as gx_set_option("residual_kernel", 1) and carries the functionals' dMdu.
//  (3) coloured elements (fallback for every pass, gx_set_option("kernel", 1), and for meshes with a node that does
//      not fit a patch): one thread per element, launches cover one colour (no two elements of a colour share a
//      node), plain read-modify-write into R / values.
//
// The element bodies are __host__ __device__ so that tests/hostcheck can run schedule (3) exactly as written, launch
// order included, and replay schedule (1) from its schedule words on the CPU (test-only; there is no CPU product path).
#pragma once

#include <stdint.h>

#include "element_math.cuh"
#include "gx_internal.h"
#include "tangent_record.cuh"

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
  double const* state_in;
  double* state_out;
  double* R;
  double* values;
  int* err;                     // {code, element}
  unsigned long long* plastic;  // counter
  int e0, e1;                   // slot range of this launch (coloured schedule)
  int pf_dist;                  // stage B: patches ahead whose records are pulled into L2 (0 = no prefetch)
  int pf_elems;                 // element kernels: elements ahead whose connectivity / old state are pulled into L2 (0 = off)
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
  if (MODEL == MODEL_J2) {  // Fp_old, eqps_old: one 80 B record (16 B aligned); Cp^{-1} = Fp^{-1} Fp^{-T} (goal_J2.cpp:84-87)
    double2 const* q = reinterpret_cast<double2 const*>(P.state_in + (int64_t)STATE_IN * e);
    double2 const a = ldg(q), b = ldg(q + 1), c2 = ldg(q + 2), d = ldg(q + 3), f = ldg(q + 4);
    double const Fp[9] = {a.x, a.y, b.x, b.y, c2.x, c2.y, d.x, d.y, f.x};
    eqps_old = f.y;
    cp_inverse(Fp, Cp);
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
  double const* src = P.state_in + (int64_t)STATE_IN * e;
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
// Schedule (1), the default Jacobian pass.
//   stage A  elem_record_kernel : one thread per element.  Gather, stress update, state save, and the element's
//            tangent record (tangent_record.cuh: w_n and tqw_n per node, s, G, 11 scalars -- 38 doubles).
//   stage B  patch_pair_kernel  : one thread block per patch of the schedule built by build_patch_schedule
//            (gx_setup.cpp); see below.
// ---------------------------------------------------------------------------
constexpr int ELEM_REC = TREC;  // doubles; 304 B = 19 x 16 B: an odd number of 16 B chunks, see patch_pair_kernel

// Warp-cooperative Fp update of 32 consecutive elements e0 .. e0+nrec-1 (lane = element), plastic branch only:
// Fp = exp(dgam N) Fp_old (goal_J2.cpp:128-131); on the elastic branch the reference leaves Fp untouched (:135-136).
// Per-thread stores of the 72 B Fp records touch 32 different 128 B lines per instruction (9 instructions); staged
// through shared memory the warp re-reads its 2.5 KB of old-state records (just gathered by its threads: L1 / L2
// hits) with 5 coalesced 128-bit loads and writes the Fp halves of its state records (pairs 5-9, 80 B each) with 5
// coalesced 128-bit stores.  The block load (wsave_load) is split from the rest so that callers can issue it before
// they wait for the staging buffer to become free.
// buf: >= WSAVE_DOUBLES doubles of shared memory, 16 B aligned, that belong to the warp and are free.
constexpr int WSAVE_ST = 32 * STATE_IN, WSAVE_LD = 11;  // old-state block at [0, 320), Fp rows of 11 doubles (odd: conflict-free) behind it
constexpr int WSAVE_DOUBLES = WSAVE_ST + 32 * WSAVE_LD;
__device__ __forceinline__ void wsave_load(KParams const& P, int e0, int nrec, int lane, double2 v[5]) {
  double2 const* src = reinterpret_cast<double2 const*>(P.state_in + (int64_t)STATE_IN * e0);  // 80 B records: 16 B aligned
  int const tot = (STATE_IN / 2) * nrec;
#pragma unroll
  for (int i = 0; i < 5; ++i) { int const g = lane + 32 * i; v[i] = g < tot ? __ldg(src + g) : make_double2(0.0, 0.0); }
}
__device__ __forceinline__ void wsave_finish(KParams const& P, int e0, int nrec, int lane, unsigned pmask, int plastic, double const dN[6],
                                             double2 const v[5], double* buf) {
#pragma unroll
  for (int i = 0; i < 5; ++i) reinterpret_cast<double2*>(buf)[lane + 32 * i] = v[i];
  __syncwarp();
  if (plastic) {
    double2 const* q = reinterpret_cast<double2 const*>(buf + STATE_IN * lane);  // stride 5 x 16 B (odd): conflict-free
    double2 const a = q[0], b = q[1], c = q[2], d = q[3], f = q[4];
    double const Fpo[9] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y, f.x};
    double Fpn[9];
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
// L2 prefetch of the streamed inputs (connectivity, old state) of the 32 elements starting at ep: two bulk prefetches
// issued by one warp.  The element kernels are latency-bound at their register-limited occupancy (conn -> nodes is a
// dependent chain of two DRAM round trips); with the first link served from L2 the chain is one DRAM trip shorter.
template <int MODEL>
__device__ __forceinline__ void prefetch_elements(KParams const& P, int ep, int ne, int lane) {
  if (ep >= ne) return;
  uint32_t const n = (uint32_t)min(32, ne - ep);
  if (lane == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.conn + ep), "r"(n * 16u) : "memory");
  if (MODEL == MODEL_J2 && lane == 1)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.state_in + (int64_t)STATE_IN * ep), "r"(n * (uint32_t)(STATE_IN * 8)) : "memory");
}

#ifndef GX_ELEM_MINB
#define GX_ELEM_MINB 8
#endif
template <int MODEL, bool SAVE>
__global__ void __launch_bounds__(64, GX_ELEM_MINB) elem_record_kernel(const __grid_constant__ KParams P, double* __restrict__ rec, int ne) {
  // A warp's 32 records are contiguous in global memory (9.5 KB): the threads put them into shared memory (record
  // stride 19 x 16 B, odd: conflict-free 128-bit stores) and one bulk asynchronous copy writes the block -- no
  // per-thread stride-304 B stores, no copy loop through the LSU.  The Fp update reuses the buffer afterwards.
  __shared__ __align__(128) double srec[2][32 * ELEM_REC];  // 64-thread blocks
  static_assert(32 * ELEM_REC >= WSAVE_DOUBLES, "state staging must fit the record buffer");
  int const e = blockIdx.x * blockDim.x + threadIdx.x;
  int const wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int const e0 = blockIdx.x * blockDim.x + wib * 32;  // first element of this warp
  int const nrec = min(32, ne - e0);
  if (P.pf_elems > 0) prefetch_elements<MODEL>(P, e0 + P.pf_elems, ne, lane);
  double2* mine = reinterpret_cast<double2*>(&srec[wib][lane * ELEM_REC]);
  int plastic = 0;
  double dN[6];
  if (e < ne) {
    int nd[4], b0[4], nb[4];
    Material const* matp;
    Core<double> c;
    int const rc = load_and_update<MODEL, SAVE>(P, e, true, nd, b0, nb, matp, c);
    if (rc != ERR_NONE) {
      report_error(P.err, rc, e);
#pragma unroll
      for (int k = 0; k < ELEM_REC / 2; ++k) mine[k] = make_double2(0.0, 0.0);
    } else {
      plastic = c.plastic;
      {
        double r[ELEM_REC];
        pack_trec(c, r);
#pragma unroll
        for (int k = 0; k < ELEM_REC / 2; ++k) mine[k] = make_double2(r[2 * k], r[2 * k + 1]);
      }
      if (SAVE && MODEL == MODEL_J2 && plastic) {
#pragma unroll
        for (int k = 0; k < 6; ++k) dN[k] = c.dN[k];
      }
    }
  }
  if (nrec > 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes above -> visible to the bulk copy
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(rec + (int64_t)ELEM_REC * e0),
                   "r"((uint32_t)__cvta_generic_to_shared(srec[wib])), "r"((uint32_t)nrec * (uint32_t)(ELEM_REC * 8))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    unsigned const pmask = SAVE && MODEL == MODEL_J2 ? __ballot_sync(0xffffffffu, plastic != 0) : 0u;
    double2 v[5];
    if (pmask) wsave_load(P, e0, nrec, lane, v);  // in flight while the bulk copy drains the buffer
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the buffer has been read: free (and safe to exit)
    __syncwarp();
    if (pmask) wsave_finish(P, e0, nrec, lane, pmask, plastic, dN, v, srec[wib]);
  }
  if (MODEL == MODEL_J2) {
    unsigned const b = __ballot_sync(0xffffffffu, plastic != 0);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(P.plastic, (unsigned long long)__popc(b));
  }
}


// ---------------------------------------------------------------------------
// Stage B: one thread block per patch of the precomputed patch schedule (build_patch_schedule, gx_setup.cpp).
// The block stages the tangent records of the patch's elements in shared memory with bulk asynchronous copies
// (cp.async.bulk, one per run of consecutive elements, completion on an mbarrier: every record crosses the L1 data
// pipe once per patch instead of once per incident node and lane), then every thread runs one work item:
//   PAIR : the elements around one mesh edge (a,b).  Per element the two node quadruples, s, G and the scalars are
//          read once (4 + 11 = 15 128-bit shared loads) and give BOTH mirror blocks (a,b) and (b,a), which share the node
//          vectors s w, G w and the symmetric scalars -- 32 accumulators in registers;
//   DIAG : the elements around one node: block (a,a) and the node's four residual entries;
//   ZERO : a phantom block of a partitioned context, written as zeros.
// Lists longer than PATCH_ITEM_LEN are finished by their primary item from the secondaries' partial sums (fixed
// order).  Every block of the operator, and every residual entry, is written exactly once.
// Bank conflicts: record stride 19 x 16 B (odd), so the bank group of chunk k of the record in slot s is
// (3 s + k) mod 8; the host schedule gives the 8 lanes of a quarter-warp records in 8 different groups where it can.
// ---------------------------------------------------------------------------
constexpr int PATCH_REC_LD = ELEM_REC;
GX_HD size_t patch_smem_bytes() { return ((size_t)PATCH_RECS * PATCH_REC_LD + (size_t)PATCH_PARTS * PATCH_PART_LD) * sizeof(double); }

#if defined(__CUDACC__)
// the pieces of a staged record as registers (128-bit shared loads)
struct TRegs { double s[6], G[6], sc[SC_N]; };
// chunks 8..18 of a staged record: (s00 s11)(s01 s02)(s12 G0)(G1 G2)(G3 G4)(G5 vb)(a1 Jpv)(upc va)(ppc tjv)(tq0 tq1)(tq2 rb)
__device__ __forceinline__ void trec_load_common(double2 const* q, TRegs& t) {
  double2 const c8 = q[8], c9 = q[9], c10 = q[10], c11 = q[11], c12 = q[12], c13 = q[13], c14 = q[14], c15 = q[15], c16 = q[16],
                c17 = q[17], c18 = q[18];
  t.s[0] = c8.x; t.s[1] = c8.y; t.s[2] = -(c8.x + c8.y); t.s[3] = c9.x; t.s[4] = c9.y; t.s[5] = c10.x;
  t.G[0] = c10.y; t.G[1] = c11.x; t.G[2] = c11.y; t.G[3] = c12.x; t.G[4] = c12.y; t.G[5] = c13.x;
  t.sc[SC_VB] = c13.y; t.sc[SC_A1] = c14.x; t.sc[SC_JPV] = c14.y; t.sc[SC_UPC] = c15.x; t.sc[SC_VA] = c15.y;
  t.sc[SC_PPC] = c16.x; t.sc[SC_TJV] = c16.y; t.sc[SC_TQ0] = c17.x; t.sc[SC_TQ1] = c17.y; t.sc[SC_TQ2] = c18.x; t.sc[SC_RB] = c18.y;
}
__device__ __forceinline__ void trec_load_node(double2 const* q, int n, double nq[4]) {
  double2 const a = q[2 * n], b = q[2 * n + 1];
  nq[0] = a.x; nq[1] = a.y; nq[2] = b.x; nq[3] = b.y;
}

template <bool TRANSPOSE>
__global__ void __launch_bounds__(PATCH_THREADS, PATCH_MINB) patch_pair_kernel(const __grid_constant__ KParams P, double const* __restrict__ rec,
                                                                               uint32_t const* __restrict__ sched) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t mbar;
  double* srec = reinterpret_cast<double*>(smem_raw);
  int const tid = threadIdx.x;
  uint32_t const* w = sched + (size_t)blockIdx.x * PATCH_WORDS;
  // every schedule word this thread needs, requested up front (one round trip to L2 / HBM)
  uint4 const hdr = __ldg(reinterpret_cast<uint4 const*>(w));  // n_recs, n_items, n_runs, block-wide partial sums
  uint4 const it = __ldg(reinterpret_cast<uint4 const*>(w + 4) + tid);
  uint4 const ot = __ldg(reinterpret_cast<uint4 const*>(w + 4 + 4 * PATCH_THREADS) + tid);
  uint2 const* runs = reinterpret_cast<uint2 const*>(w + 4 + 8 * PATCH_THREADS);
  static_assert(PATCH_RECS <= 2 * PATCH_THREADS && PATCH_THREADS % 32 == 0, "two runs per thread at most");
  // Run j goes to lane j / W of warp j % W (W warps): issuing a bulk copy is serialised over the lanes of a warp
  // (uniform-datapath operands), so the patch's copies are spread over all warps instead of filling warp 0 first.
  constexpr int NW = PATCH_THREADS / 32;
  int const jrun = (tid & 31) * NW + (tid >> 5);
  uint2 const run0 = jrun < PATCH_RECS ? __ldg(runs + jrun) : make_uint2(0u, 0u);
  uint2 const run1 = jrun + PATCH_THREADS < PATCH_RECS ? __ldg(runs + jrun + PATCH_THREADS) : make_uint2(0u, 0u);
  // L2 prefetch, two stages deep: the schedule words of patch b + 2D, and -- from its words, fetched by block b - D --
  // the records of patch b + D.  Blocks are dispatched in index order, so both arrive one to two generations early
  // and the two dependent round trips of a block (words, then records) hit L2 instead of HBM.
  uint2 pf0 = make_uint2(0u, 0u), pf1 = make_uint2(0u, 0u);
  if (P.pf_dist > 0) {
    unsigned const p2 = blockIdx.x + 2u * (unsigned)P.pf_dist, p1 = blockIdx.x + (unsigned)P.pf_dist;
    if (p2 < gridDim.x && 32 * tid < PATCH_WORDS)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(sched + (size_t)p2 * PATCH_WORDS + 32 * tid));
    if (p1 < gridDim.x) {
      uint2 const* r1 = reinterpret_cast<uint2 const*>(sched + (size_t)p1 * PATCH_WORDS + 4 + 8 * PATCH_THREADS);
      if (jrun < PATCH_RECS) pf0 = __ldg(r1 + jrun);
      if (jrun + PATCH_THREADS < PATCH_RECS) pf1 = __ldg(r1 + jrun + PATCH_THREADS);
    }
  }
  int const n_recs = (int)hdr.x;
  // Record staging: one bulk asynchronous copy (global -> shared) per run of the schedule -- a run is a number of
  // consecutive elements' records (304 B each, contiguous in global memory) that go to consecutive slots; the copies
  // report their bytes to an mbarrier that the whole block then waits on.
  uint32_t const mb = (uint32_t)__cvta_generic_to_shared(&mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(n_recs * (ELEM_REC * 8)) : "memory");
  auto stage_run = [&](uint2 const run) {
    uint32_t const sl = run.y & 0xffu, len = run.y >> 8;
    if (len == 0) return;  // past the end of the run list (the words are zero there)
    uint32_t const dst = (uint32_t)__cvta_generic_to_shared(srec) + sl * (uint32_t)(PATCH_REC_LD * 8);
    double const* src = rec + (int64_t)ELEM_REC * (int64_t)run.x;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(len * (uint32_t)(ELEM_REC * 8)), "r"(mb)
                 : "memory");
  };
  stage_run(run0);
  stage_run(run1);
  auto prefetch_run = [&](uint2 const run) {
    uint32_t const len = run.y >> 8;
    if (len == 0) return;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(rec + (int64_t)ELEM_REC * (int64_t)run.x), "r"(len * (uint32_t)(ELEM_REC * 8)) : "memory");
  };
  prefetch_run(pf0);
  prefetch_run(pf1);
  {
    uint32_t done;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb) : "memory");
    } while (!done);
  }
  int const kind = (int)(ot.w & 3u), type = (int)((ot.w >> 2) & 3u);
  int const part = (int)((ot.w >> 4) & 0xffu), nsec = (int)((ot.w >> 12) & 0x3fu);
  if (kind == 0) return;
  double acc1[16], acc2[16], r4[4];
#pragma unroll
  for (int k = 0; k < 16; ++k) { acc1[k] = 0.0; acc2[k] = 0.0; }
  r4[0] = r4[1] = r4[2] = r4[3] = 0.0;
  uint64_t elo = (uint64_t)it.x | ((uint64_t)it.y << 32), ehi = (uint64_t)it.z | ((uint64_t)it.w << 32);
  if (type == 2) {
#pragma unroll 1
    for (int k = 0; k < PATCH_ITEM_LEN; ++k) {
      uint32_t const ent = (uint32_t)elo & 0xffffu;
      elo = (elo >> 16) | (ehi << 48); ehi >>= 16;
      if (!(ent & 0x8000u)) {  // an empty round of this item (bank-conflict avoidance), or the end of its list
        if ((elo | ehi) == 0) break;
        continue;
      }
      double2 const* q = reinterpret_cast<double2 const*>(srec + (ent & 0xffu) * PATCH_REC_LD);
      double nq[4], mq[4];
      trec_load_node(q, (int)((ent >> 10) & 3u), nq);
      trec_load_node(q, (int)((ent >> 8) & 3u), mq);
      TRegs t;
      trec_load_common(q, t);
      trec_pair_add<TRANSPOSE>(nq, mq, t.s, t.G, t.sc, acc1, acc2);
    }
  } else if (type == 1) {
#pragma unroll 1
    for (int k = 0; k < PATCH_ITEM_LEN; ++k) {
      uint32_t const ent = (uint32_t)elo & 0xffffu;
      elo = (elo >> 16) | (ehi << 48); ehi >>= 16;
      if (!(ent & 0x8000u)) {
        if ((elo | ehi) == 0) break;
        continue;
      }
      double2 const* q = reinterpret_cast<double2 const*>(srec + (ent & 0xffu) * PATCH_REC_LD);
      double nq[4];
      trec_load_node(q, (int)((ent >> 10) & 3u), nq);
      TRegs t;
      trec_load_common(q, t);
      trec_diag_add<TRANSPOSE>(nq, t.s, t.G, t.sc, acc1, r4);
    }
  }
  // Finish.  Items without secondaries write their block(s) and leave; only the few items that exchange partial sums
  // (diagonal blocks, edges of high valence) meet at the barrier.
  double* spart = srec + (size_t)PATCH_RECS * PATCH_REC_LD;  // [PATCH_PARTS][PATCH_PART_LD]
  auto write_out = [&]() {
    int const j1 = (int)(ot.z & 0xffu), nb1 = (int)((ot.z >> 8) & 0xffu);
    double* out = P.values + 16 * (int64_t)ot.x + 4 * j1;
#pragma unroll
    for (int i = 0; i < 4; ++i) stg256(out + (int64_t)i * (4 * nb1), acc1[4 * i], acc1[4 * i + 1], acc1[4 * i + 2], acc1[4 * i + 3]);  // one 32 B sector each
    if (type == 2) {
      int const j2 = (int)((ot.z >> 16) & 0xffu), nb2 = (int)(ot.z >> 24);
      double* out2 = P.values + 16 * (int64_t)ot.y + 4 * j2;
#pragma unroll
      for (int i = 0; i < 4; ++i) stg256(out2 + (int64_t)i * (4 * nb2), acc2[4 * i], acc2[4 * i + 1], acc2[4 * i + 2], acc2[4 * i + 3]);
    } else if (type == 1) {
      stg256(P.R + 4 * (int64_t)ot.y, r4[0], r4[1], r4[2], r4[3]);
    }
  };
  if (kind == 1 && nsec == 0) { write_out(); return; }
  if (kind == 2) {
    double2* d = reinterpret_cast<double2*>(spart + PATCH_PART_LD * part);
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = make_double2(acc1[2 * k], acc1[2 * k + 1]);
    if (type == 2) {
#pragma unroll
      for (int k = 0; k < 8; ++k) d[8 + k] = make_double2(acc2[2 * k], acc2[2 * k + 1]);
    }
    d[16] = make_double2(r4[0], r4[1]);
    d[17] = make_double2(r4[2], r4[3]);
  }
  // The schedule keeps a primary and its secondaries in one warp wherever it can (header word 3 = 0): the hand-over
  // then needs only a warp barrier and the warps of a block never wait for each other.
  if (hdr.w) __syncthreads();  // the threads that are still here
  else __syncwarp();
  if (kind == 1) {
    for (int s2 = 0; s2 < nsec; ++s2) {
      double2 const* d = reinterpret_cast<double2 const*>(spart + PATCH_PART_LD * (part + s2));
#pragma unroll
      for (int k = 0; k < 8; ++k) { double2 const v = d[k]; acc1[2 * k] += v.x; acc1[2 * k + 1] += v.y; }
      if (type == 2) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { double2 const v = d[8 + k]; acc2[2 * k] += v.x; acc2[2 * k + 1] += v.y; }
      }
      if (type == 1) {
        double2 v = d[16]; r4[0] += v.x; r4[1] += v.y;
        v = d[17]; r4[2] += v.x; r4[3] += v.y;
      }
    }
    write_out();
  }
}
#endif

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
  __shared__ __align__(16) double sbuf[4][WSAVE_DOUBLES];
  static_assert(WSAVE_DOUBLES >= 32 * 17 && WSAVE_DOUBLES % 2 == 0, "residual staging must fit");
  int const e = blockIdx.x * blockDim.x + threadIdx.x;
  int const wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int const e0 = blockIdx.x * blockDim.x + wib * 32;
  int const nrec = min(32, ne - e0);
  if (P.pf_elems > 0) prefetch_elements<MODEL>(P, e0 + P.pf_elems, ne, lane);
  double* buf = sbuf[wib];
  int plastic = 0;
  double dN[6];
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
    unsigned const pmask = SAVE && MODEL == MODEL_J2 ? __ballot_sync(0xffffffffu, plastic != 0) : 0u;
    double2 v[5];
    if (pmask) wsave_load(P, e0, nrec, lane, v);  // in flight during the copy loop
    double* dst = rvec + 16 * (int64_t)e0;
    for (int g = lane; g < nrec * 8; g += 32) {
      double const* q = buf + 17 * (g >> 3) + 2 * (g & 7);
      *reinterpret_cast<double2*>(dst + 2 * g) = make_double2(q[0], q[1]);
    }
    if (pmask) {
      __syncwarp();
      wsave_finish(P, e0, nrec, lane, pmask, plastic, dN, v, buf);
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
// Residual and error-localisation passes, block-reduced form (default; schedule: build_residual_schedule, gx_setup.cpp).
//   elem_residual_block_kernel : RES_BLOCK consecutive elements per thread block, one per thread.  The 16 residual
//                          entries of every element stay in shared memory; after a block barrier the threads sum them
//                          per node of the block in the schedule's fixed order and write R (node complete in this
//                          block) or one 32 B partial sum (node shared with other blocks).  State save as in stage A.
//   node_partial_sum_kernel    : adds the partial sums of every shared node in ascending block order and writes R.
// Against the element-line form above, the 128 B per element written and read back (plus the 32 B of adjacency) shrink
// to the partial sums and the schedule words: about 2 x 22 + 14 B per element on an x-fastest numbered Kuhn cube.
// ---------------------------------------------------------------------------
template <int MODEL, bool SAVE, bool ERROR>
__global__ void __launch_bounds__(RES_BLOCK, RES_MINB) elem_residual_block_kernel(const __grid_constant__ KParams P, uint32_t const* __restrict__ sched,
                                                                          uint32_t const* __restrict__ boff, double* __restrict__ partial, int ne) {
  __shared__ __align__(16) double srv[RES_BLOCK * 16];                   // [local element][local node][4]
  __shared__ __align__(16) double sbuf[RES_BLOCK / 32][WSAVE_DOUBLES];   // Fp staging, one buffer per warp
  __shared__ __align__(16) uint32_t ssched[RES_MAX_WORDS];
  __shared__ __align__(8) unsigned long long mbar;
  int const tid = threadIdx.x, wib = tid >> 5, lane = tid & 31;
  int const e = blockIdx.x * RES_BLOCK + tid;
  int const e0 = blockIdx.x * RES_BLOCK + wib * 32;
  int const nrec = min(32, ne - e0);
  // the block's schedule words arrive by one bulk copy while the elements are evaluated
  uint32_t const mb = (uint32_t)__cvta_generic_to_shared(&mbar);
  if (tid == 0) {
    uint32_t const w0 = __ldg(boff + blockIdx.x), w1 = __ldg(boff + blockIdx.x + 1);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((w1 - w0) * 4u) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(ssched)),
                 "l"(sched + w0), "r"((w1 - w0) * 4u), "r"(mb)
                 : "memory");
  }
  if (P.pf_elems > 0) {
    prefetch_elements<MODEL>(P, e0 + P.pf_elems, ne, lane);
    int const bp = blockIdx.x + P.pf_elems / RES_BLOCK;
    if (tid == 2 && bp < (int)gridDim.x) {
      uint32_t const w0 = __ldg(boff + bp), w1 = __ldg(boff + bp + 1);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(sched + w0), "r"((w1 - w0) * 4u) : "memory");
    }
  }
  int plastic = 0;
  double dN[6];
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
    // 128 B rows: chunk k of row r sits at k ^ (r & 7), so that the 8 lanes of a quarter-warp store to 8 different bank
    // groups; the schedule's entries are the swizzled chunk numbers
    double2* o = reinterpret_cast<double2*>(srv + 16 * tid);
    int const x = tid & 7;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      o[(2 * n) ^ x] = make_double2(ru[3 * n], ru[3 * n + 1]);
      o[(2 * n + 1) ^ x] = make_double2(ru[3 * n + 2], rp[n]);
    }
  }
  if (SAVE && MODEL == MODEL_J2 && nrec > 0) {  // warp-local, before the block barrier: its loads overlap the other warps' work
    unsigned const pmask = __ballot_sync(0xffffffffu, plastic != 0);
    if (pmask) {
      double2 v[5];
      wsave_load(P, e0, nrec, lane, v);
      wsave_finish(P, e0, nrec, lane, pmask, plastic, dN, v, sbuf[wib]);
    }
  }
  if (MODEL == MODEL_J2) {
    unsigned const b = __ballot_sync(0xffffffffu, plastic != 0);
    if (lane == 0 && b) atomicAdd(P.plastic, (unsigned long long)__popc(b));
  }
  // Block barrier, then all warps share the reduction.  (Measured alternative: the last warp to arrive reduces alone
  // and the others leave -- 1.37 instead of 1.09 ms, the lone warp's six serial rounds hold the block's resources.)
  __syncthreads();
  {
    uint32_t done;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb) : "memory");
    } while (!done);
  }
  int const S = (int)ssched[0];
  uint16_t const* ent = reinterpret_cast<uint16_t const*>(ssched + RES_HDR + 2 * S);
  double2 const* sv = reinterpret_cast<double2 const*>(srv);
  for (int item = tid; item < 2 * S; item += RES_BLOCK) {  // 2 threads per node: residual entries {0, 1} and {2, 3}
    int const s = item >> 1, h = item & 1;
    uint32_t const w0 = ssched[RES_HDR + 2 * s], w1 = ssched[RES_HDR + 2 * s + 1];
    uint32_t const first = w1 & 0xffffu, cnt = w1 >> 16;
    // the entry is the (swizzled) chunk of the node's first half.  Two independent chains (even / odd entries), added
    // at the end: a fixed order, half the dependent shared-memory latency
    double2 acc = make_double2(0.0, 0.0), acc1 = make_double2(0.0, 0.0);
    uint32_t k = first;
    uint32_t const end = first + cnt;
    for (; k + 1 < end; k += 2) {
      double2 const v0 = sv[(uint32_t)ent[k] ^ (uint32_t)h], v1 = sv[(uint32_t)ent[k + 1] ^ (uint32_t)h];
      acc.x += v0.x; acc.y += v0.y; acc1.x += v1.x; acc1.y += v1.y;
    }
    if (k < end) { double2 const v0 = sv[(uint32_t)ent[k] ^ (uint32_t)h]; acc.x += v0.x; acc.y += v0.y; }
    acc.x += acc1.x; acc.y += acc1.y;
    double* dst = (w0 & 0x80000000u) ? P.R : partial;
    *reinterpret_cast<double2*>(dst + 4 * (int64_t)(w0 & 0x7fffffffu) + 2 * h) = acc;
  }
}

__global__ void __launch_bounds__(256) node_partial_sum_kernel(int32_t const* __restrict__ pnode, uint32_t const* __restrict__ poff,
                                                               double const* __restrict__ partial, double* __restrict__ R, int np) {
  int const t = blockIdx.x * blockDim.x + threadIdx.x;
  int const i = t >> 1, h = t & 1;  // 2 threads per node
  if (i >= np) return;
  uint32_t const p0 = __ldg(poff + i), p1 = __ldg(poff + i + 1);
  double2 acc = make_double2(0.0, 0.0);
  for (uint32_t p = p0; p < p1; ++p) {
    double2 const v = __ldg(reinterpret_cast<double2 const*>(partial + 4 * (int64_t)p) + h);
    acc.x += v.x; acc.y += v.y;
  }
  *reinterpret_cast<double2*>(R + 4 * (int64_t)__ldg(pnode + i) + 2 * h) = acc;
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
