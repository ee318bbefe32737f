// gx_internal.h -- context layout shared by the host setup code and the CUDA side.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <string>
#include <vector>

#include "../../include/goal_b200.h"
#include "element_math.cuh"

namespace gx {

// One node as the kernels gather it: 64 bytes, one aligned line-half.
//   d0 = (x, y)  d1 = (z, ux)  d2 = (uy, uz)  d3 = (p, {blk0, nblk})
// blk0 = index of the node's first 4x4 block in the block-CRS (value offset of
// dof row 4a+i is 16*blk0 + i*4*nblk), nblk = number of node blocks in its row.
struct alignas(64) NodeRec {
  double x[3];
  double u[3];
  double p;
  int32_t blk0;
  int32_t nblk;
};
static_assert(sizeof(NodeRec) == 64, "NodeRec must be 64 bytes");

// adjoint weights gathered by the error-localisation kernel: (zu0,zu1) (zu2,zp) (zpc,-) -> 48 B padded to 64
struct alignas(64) ZRec {
  double zu[3];
  double zp;
  double zpc;
  double pad[3];
};
static_assert(sizeof(ZRec) == 64, "ZRec must be 64 bytes");

struct Peer {
  int rank = -1;
  std::vector<int32_t> nodes;       // shared nodes, agreed order
  std::vector<int32_t> send_nodes;  // subset owned by the peer   (we send our partial rows)
  std::vector<int32_t> recv_nodes;  // subset owned by this rank  (we receive and add)
  // structure exchange: per send node [nblk, global column node ids...]
  std::vector<int64_t> struct_out, struct_in;
  bool struct_have = false;
  // value exchange plan
  std::vector<int64_t> send_off, recv_off;  // [n+1] offsets (doubles) of each node's packed block rows
  std::vector<int64_t> recv_moff;           // [n_recv+1] offsets into recv_map
  std::vector<int32_t> recv_cnt;            // sender's block count per recv node
  std::vector<int32_t> recv_map;            // sender block -> block position in the owner's extended row
  int64_t send_vals = 0, recv_vals = 0;     // doubles of CRS payload
  // device side
  int32_t *d_send_nodes = nullptr, *d_recv_nodes = nullptr, *d_recv_map = nullptr, *d_recv_cnt = nullptr;
  int64_t *d_send_off = nullptr, *d_recv_off = nullptr, *d_recv_moff = nullptr;
  double *d_send = nullptr, *d_recv = nullptr, *d_sendR = nullptr, *d_recvR = nullptr;
};

struct NcclApi;  // resolved with dlopen (gx_nccl.cpp)

}  // namespace gx

struct gx_ctx {
  // ---- description
  int nn = 0, ne = 0, nsets = 1, model = 0, device = 0;
  uint32_t flags = 0;
  int rank = 0, nranks = 1;
  gx::Material mats[GX_MAX_ELEM_SETS];
  std::vector<int32_t> conn;    // user order
  std::vector<double> coords;
  std::vector<int32_t> eset;
  std::vector<int64_t> node_gid;
  std::vector<int32_t> node_owner;
  // ---- graph (host)
  std::vector<int64_t> nrow;    // [nn+1] block-row offsets
  std::vector<int32_t> ncol;    // [nblocks] neighbour node of each block, sorted per row
  int64_t nnz = 0;
  std::vector<int64_t> rowptr;  // lazily materialised dof-level CRS
  std::vector<int32_t> colind;
  std::vector<uint8_t> bpos;    // [ne*16] user order
  // node -> (element, local node) incidences, elements ascending: the row-owner kernel's work list
  std::vector<uint32_t> adj_off;  // [nn+1]
  std::vector<int2> adj;          // [4*ne]  x = e*4+n, y = block positions of (a, a_m), m = 0..3, one byte each
  int max_nblk = 0, max_deg = 0;
  bool has_isolated_nodes = false;  // some node belongs to no element (its R entries must be zeroed by the pass)
  // order in which stage B visits the nodes: Morton (Z-curve) order of the node coordinates, so that the four
  // incidences of an element are processed close in time and its tangent record is fetched from HBM once
  std::vector<int32_t> node_order;
  std::vector<uint8_t> diag_pos;  // position of block (a,a) in node a's block row
  bool block_lists_built = false;
  // ---- patch schedule of the Jacobian pass (stage B, patch_pair_kernel), see build_patch_schedule()
  std::vector<uint32_t> patch_sched;                  // flat (flatten_patch_schedule), or empty while ...
  std::vector<std::vector<uint32_t>> patch_chunks;    // ... the builder's chunks are still waiting for the upload
  int n_patches = 0;
  int n_patches_iface = 0;  // partitioned contexts: the leading patches that write every block of the interface rows
  int patch_state = 0;  // 0 = not built, 1 = built, -1 = mesh does not fit (a node exceeds a patch)
  // ---- block-reduced schedule of the residual / error-localisation passes, see build_residual_schedule()
  std::vector<std::vector<uint32_t>> res_chunks;  // the blocks' words, one chunk per builder thread (concatenated = the schedule)
  std::vector<uint32_t> res_boff;                 // [blocks + 1] first word of every block
  std::vector<int32_t> res_pnode;                 // nodes finished by node_partial_sum_kernel (shared between blocks, or without elements)
  std::vector<uint32_t> res_poff;                 // [res_pnode.size() + 1] their runs in the partial-sum buffer
  int64_t res_npartial = 0;
  int res_state = 0;  // 0 = not built, 1 = built, -1 = does not fit
  // ---- schedule
  int ncolors = 0;
  std::vector<int32_t> color_off;  // [ncolors+1] in device element order
  std::vector<int32_t> perm;       // device slot -> user element
  // ---- device
  cudaStream_t stream = nullptr;
  cudaStream_t comm_stream = nullptr;  // interface exchange overlapped with the interior patches (option "overlap")
  cudaEvent_t ev_iface = nullptr, ev_b2 = nullptr, ev_comm = nullptr;
  cudaEvent_t ev_stage = nullptr;  // between the two kernels of a two-stage pass (stage_ms)
  bool staged = false;
  double stage_ms[2] = {0, 0};
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  gx::NodeRec* d_nodes = nullptr;
  gx::ZRec* d_z = nullptr;
  int4* d_conn = nullptr;       // user element order
  uint4* d_bpos = nullptr;      // user element order
  uint8_t* d_eset = nullptr;    // user element order
  int32_t* d_perm = nullptr;    // colour schedule: slot -> user element
  uint32_t* d_adj_off = nullptr;
  int2* d_adj = nullptr;
  uint8_t* d_diag_pos = nullptr;   // position of block (a,a) in node a's block row
  uint32_t* d_patch_sched = nullptr;
  uint32_t* d_res_sched = nullptr;
  uint32_t* d_res_boff = nullptr;
  int32_t* d_res_pnode = nullptr;
  uint32_t* d_res_poff = nullptr;
  double* d_res_partial = nullptr;  // [res_npartial][4]
  // history state, one record per element (user order):
  //   in  : Fp_old[9], eqps_old                              (80 B)  read once per element and pass
  //   out : sigma[9], eqps, Fp[9], pad                       (160 B)
  double* d_state_in = nullptr;
  double* d_state_out = nullptr;
  double* d_elemrec = nullptr;  // [ne][ELEM_REC] tangent records of the two-kernel Jacobian pass (lazy)
  double* d_R = nullptr;
  double* d_values = nullptr;
  double* d_stage = nullptr;  // staging for host<->device field copies, >= max(4*nn, 10*ne) doubles
  int64_t stage_len = 0;
  int* d_err = nullptr;                 // {code, element} + the plastic counter behind it (one 16-byte block)
  unsigned long long* d_plastic = nullptr;  // = d_err + 2
  int* h_status = nullptr;              // pinned host copy of that block
  double* d_red = nullptr;              // reduction scratch
  double* d_dMdu = nullptr;             // [4 nn] ghost dMdu of the last gx_functional (lazy)
  bool have_dMdu = false;
  int32_t* d_child_off = nullptr;       // parent -> children CRS for set_error
  int32_t* d_child = nullptr;
  int n_parent_cached = -1;
  std::vector<int32_t> parent_cached;
  // ---- partition
  std::vector<gx::Peer> peers;
  gx::NcclApi* nccl = nullptr;
  void* comm = nullptr;
  bool struct_done = true;
  // extended block rows: ghost row + phantom columns appended (owned interface nodes only)
  std::vector<int64_t> nrow_x;    // [nn+1]
  std::vector<int64_t> xcol_gid;  // [nrow_x[nn]] global node id of every block
  std::vector<int32_t> xcol_owner;  // [nrow_x[nn]] owning rank of that node
  // owned matrix in Tpetra's local layout (gx_owned_tpetra_graph): column-map node gids, dof-level rows, and per stored
  // block of an owned row its position in the Tpetra-ordered row
  std::vector<int64_t> tp_colmap, tp_rowptr;
  std::vector<int32_t> tp_colind;
  std::vector<uint8_t> tp_perm;
  int64_t nnz_x = 0;
  std::vector<int32_t> owned_nodes;
  std::vector<int64_t> owned_rowptr, owned_colgid;
  int32_t *d_blk0_x = nullptr, *d_nblk_x = nullptr, *d_nblk_g = nullptr;
  int64_t* d_blk0_g = nullptr;
  double* d_ghost_vals = nullptr;
  // ---- bookkeeping
  int64_t last_plastic = 0;
  double timing[4] = {0, 0, 0, 0};
  int launches = 0;
  bool have_result = false;
  bool have_values = false;
  int64_t opt_block = 128;
  // element kernels: L2 prefetch distance in elements, {passes that save the state, passes that do not}.  Measured on
  // 12.6M tets (profiles/r02u_stage_a_prefetch.txt): saving passes are best at 9.5k-57k (the prefetched lines must survive
  // 574 B/element of streaming traffic in L2), the others at 150k-200k.
  int64_t opt_prefetch_elems[2] = {18944, 151552};
  int64_t opt_prefetch = 256;  // stage B L2 prefetch distance in patches (measured best on 12.6M tets: 150-300; one generation of resident blocks is 592)
  int64_t overlap_now = 0;  // `what` of the exchange fused into the pass being enqueued (0: none)
  bool overlapped = false;  // the last pass reduced the interfaces itself
  int64_t opt_overlap = 0;  // bit mask like gx_reduce_interfaces' `what` (1 = R, 2 = dRdu): the Jacobian pass reduces the
                            // interfaces itself, on a second stream, while the interior patches are still being assembled
  int64_t opt_residual = 0;  // residual / localisation passes: 0 = block-reduced, 1 = element lines + node gather
  int64_t opt_kernel = 0;  // 0 = owner-computes schedules (patch pairs / gather form), 1 = coloured elements
  int num_sms = 148;
  std::string err;
};

namespace gx {
// host setup (gx_setup.cpp)
struct SetupTimer {  // GX_SETUP_TIMING=1 prints the wall time of every setup section to stderr
  bool on = getenv("GX_SETUP_TIMING") != nullptr;
  double t0 = now();
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
  void lap(char const* what) { if (!on) return; double const t = now(); fprintf(stderr, "[gx setup] %-28s %.3f s\n", what, t - t0); t0 = t; }
};

int build_graph_and_schedule(gx_ctx* c);
int build_colouring(gx_ctx* c);  // lazily: only the coloured fallback needs it
void materialise_crs(gx_ctx* c);
// host images of the device arrays (user element order; the colour schedule indexes them through perm)
struct HostPack {
  std::vector<NodeRec> nodes;
  std::vector<int4> conn4;   // user element order
  std::vector<uint4> bpos;   // user element order
  std::vector<uint8_t> eset;
};
constexpr int STATE_IN = 10;   // doubles per element: Fp_old[9], eqps_old
constexpr int STATE_OUT = 20;   // sigma[9] | eqps | Fp[9] | pad: 16 B pairs 0-4 are always written, pairs 5-9 only on the plastic branch
constexpr int SO_SIGMA = 0, SO_EQPS = 9, SO_FP = 10;
void pack_host(gx_ctx const* c, HostPack& h);
void build_block_lists(gx_ctx* c);
bool build_patch_schedule(gx_ctx* c);
bool build_residual_schedule(gx_ctx* c);
void flatten_patch_schedule(gx_ctx* c);

// patch schedule geometry (shared by gx_setup.cpp and the kernel)
#ifndef GX_PATCH_THREADS
#define GX_PATCH_THREADS 96
#define GX_PATCH_RECS 160
#define GX_PATCH_MINB 4
#define GX_PATCH_PARTS 24
#endif
constexpr int PATCH_THREADS = GX_PATCH_THREADS;  // work items per patch, one per thread
constexpr int PATCH_RECS = GX_PATCH_RECS;        // element records staged per patch (304 B each); slots are 8 bit
constexpr int PATCH_MINB = GX_PATCH_MINB;        // thread blocks per SM the kernel is compiled for
constexpr int PATCH_ITEM_LEN = 8;   // contributions per work item
constexpr int PATCH_PARTS = GX_PATCH_PARTS;  // secondary items (partial sums handed to a primary) per patch
constexpr int PATCH_DIAG_LANES = 32;  // lanes [0, 32) run the DIAG items, the other warps the PAIR items
constexpr int PATCH_PART_LD = 36;   // doubles per partial sum: two 4x4 blocks + the residual entries
constexpr int PATCH_WORDS = 4 + 4 * PATCH_THREADS + 4 * PATCH_THREADS + 2 * PATCH_RECS;  // uint32 words per patch
static_assert(PATCH_RECS <= 256, "record slots are 8 bit");
// Block-reduced schedule of the residual / error-localisation passes (build_residual_schedule, gx_setup.cpp): blocks of
// RES_BLOCK consecutive elements.  Words of one block: {slots S, words of this block, 0, 0}, then per slot (= distinct
// node of the block, in order of first appearance) {target | complete << 31, first | count << 16}, then the block's
// 4 * elements incidence entries grouped by slot, ascending elements inside a slot (uint16 = 8 * local element +
// ((2 * local node) ^ (local element & 7)): the 16 B chunk, in the kernel's swizzled shared-memory rows, that holds
// the first two of the node's four residual entries; the other two sit in chunk ^ 1); padded to a multiple of 4 words.  target = the node when the block holds all its elements (the block writes R), else
// the position of this block's partial sum in the partial buffer (node_partial_sum_kernel adds those in block order).
#ifndef GX_RES_BLOCK
#define GX_RES_BLOCK 128
#endif
constexpr int RES_BLOCK = GX_RES_BLOCK;  // 64, 128 or 256
#ifndef GX_RES_MINB
#define GX_RES_MINB (512 / GX_RES_BLOCK)
#endif
constexpr int RES_MINB = GX_RES_MINB;  // thread blocks per SM the kernel is compiled for (4 x 128 threads x 128 registers)
constexpr int RES_HDR = 4;
constexpr int RES_MAX_WORDS = RES_HDR + 2 * 4 * RES_BLOCK + 2 * RES_BLOCK;  // every incidence a slot of its own
// gx_comm.cu
void comm_destroy(gx_ctx*);
int comm_setup_lists(gx_ctx*, const gx_desc*);
int ghost_values_dev(gx_ctx* ctx, double** out);
int comm_enqueue_reduce(gx_ctx* ctx, int what, cudaStream_t stream);  // pack, send/recv, unpack-add: enqueued, not synchronised
}  // namespace gx
