// gx_kernels.cuh -- the assembly kernels: one element per thread.
//
// Replaces the element loop of goal::assemble (src/goal_assembly.cpp:65-88) together
// with the gather/scatter halves of Displacement/Pressure (src/goal_displacement.cpp,
// src/goal_pressure.cpp) and States get/set (src/goal_states.cpp:21-57).
//
// Data layout in HBM
//   nodes   NodeRec[Nn]  64 B per node: x, u, p and the node's block-row descriptor;
//                        one element gathers 4 records = 4 full 64 B segments (4 LDG.128 each)
//   conn    int4[Ne]     one coalesced 128-bit load per thread
//   bpos    uint4[Ne]    16 x uint8 block positions, one coalesced 128-bit load per thread
//   state   SoA, component-major with stride `sstride`: Fp_old[9][Ne], eqps_old[Ne], ...
//                        -> every state load/store of a warp is one 256 B contiguous run
//   R       double[4 Nn] ghost layout;   values  double[nnz]  CRS order of gx_graph
// Elements are stored colour-sorted (gx_setup.cpp); a launch covers one colour, so all
// read-modify-writes below are conflict free without atomics.
//
// The element bodies are __host__ __device__ so that tests/hostcheck can run exactly
// this code, launch order included, on the CPU (test-only; there is no CPU product path).
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
  double const* Fp_old;
  double const* eqps_old;
  double* Fp;
  double* eqps;
  double* sigma;
  int64_t sstride;
  double* R;
  double* values;
  int* err;                     // {code, device element}
  unsigned long long* plastic;  // counter
  int e0, e1;                   // device element range of this launch
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

GX_HD void load_node(NodeRec const* nodes, int id, double x[3], double u[3], double& p, int& blk0, int& nblk) {
  double2 const* q = reinterpret_cast<double2 const*>(nodes + id);
  double2 const d0 = ldg(q), d1 = ldg(q + 1), d2 = ldg(q + 2), d3 = ldg(q + 3);
  x[0] = d0.x; x[1] = d0.y; x[2] = d1.x;
  u[0] = d1.y; u[1] = d2.x; u[2] = d2.y;
  p = d3.x;
#if defined(__CUDA_ARCH__)
  blk0 = __double2loint(d3.y);
  nblk = __double2hiint(d3.y);
#else
  int32_t t[2];
  __builtin_memcpy(t, &d3.y, 8);
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

// One element of one colour.  Returns 1 when the element took the plastic branch.
template <int MODEL, int PASS, bool SAVE>
GX_HD int assemble_element(KParams const& P, int e) {
  int4 const cn = ldg(P.conn + e);
  int const nd[4] = {cn.x, cn.y, cn.z, cn.w};
  double x[4][3], u[4][3], p[4];
  int blk0[4], nblk[4];
#pragma unroll
  for (int n = 0; n < 4; ++n) load_node(P.nodes, nd[n], x[n], u[n], p[n], blk0[n], nblk[n]);
  Material const& mat = P.mat[P.eset ? P.eset[e] : 0];

  double Fp_old[9], eqps_old = 0.0;
  if (MODEL == MODEL_J2) {
#pragma unroll
    for (int k = 0; k < 9; ++k) Fp_old[k] = ldg(P.Fp_old + k * P.sstride + e);
    eqps_old = ldg(P.eqps_old + e);
  }
  Core<double> c;
  double sig[9], eqps_new = 0.0, Fp_new[9];
  bool write_Fp = false;
  int const rc = element_core<MODEL>(x, u, p, mat, Fp_old, eqps_old, SAVE, sig, eqps_new, Fp_new, write_Fp, c);
  if (rc != ERR_NONE) { report_error(P.err, rc, e); return 0; }
  if (SAVE) {
#pragma unroll
    for (int k = 0; k < 9; ++k) P.sigma[k * P.sstride + e] = sig[k];
    if (MODEL == MODEL_J2) {
      P.eqps[e] = eqps_new;
      if (write_Fp) {
#pragma unroll
        for (int k = 0; k < 9; ++k) P.Fp[k * P.sstride + e] = Fp_new[k];
      }
    }
  }

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
    element_error_residual(c, mat, zu, zp, zpc, ru, rp);
  } else {
    element_residual(c, mat, ru, rp);
  }
#pragma unroll
  for (int n = 0; n < 4; ++n) add4(P.R + 4 * (int64_t)nd[n], ru[3 * n], ru[3 * n + 1], ru[3 * n + 2], rp[n]);

  // ---- Jacobian, one column node m at a time; 4x4 node blocks go straight into the CRS
  if (PASS == PASS_JACOBIAN || PASS == PASS_JACOBIAN_T) {
    uint4 const bq = ldg(P.bpos + e);
    uint32_t const bw[4] = {bq.x, bq.y, bq.z, bq.w};  // bw[n] byte m = position of block (n,m)
    double sw[4][3];
#pragma unroll
    for (int n = 0; n < 4; ++n) sym_mv(c.s, c.w[n], sw[n]);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      ColNode<double> cnm;
      column_node(c, m, cnm);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        double blk[16];
        jacobian_block(c, mat, n, cnm, sw[n], blk);
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
  return c.plastic;
}

#if defined(__CUDACC__)
template <int MODEL, int PASS, bool SAVE>
__global__ void __launch_bounds__(128) assemble_kernel(const __grid_constant__ KParams P) {
  int const e = P.e0 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
  int plastic = 0;
  if (e < P.e1) plastic = assemble_element<MODEL, PASS, SAVE>(P, e);
  if (MODEL == MODEL_J2) {
    unsigned const b = __ballot_sync(0xffffffffu, plastic != 0);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(P.plastic, (unsigned long long)__popc(b));
  }
}
#endif

}  // namespace gx
