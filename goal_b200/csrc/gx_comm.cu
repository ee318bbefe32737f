// gx_comm.cu -- mesh parts: owned rows, interface exchange, NCCL transport.
//
// What this replaces in the reference:
//   C4  Disc::compute_graphs   owned_graph = Export(ghost_graph, INSERT)        src/goal_disc.cpp:327-329
//   C1  SolInfo::gather_R      owned R    += Export(ghost R,    ADD)            src/goal_sol_info.cpp:33-35
//   C2  SolInfo::gather_dRdu   owned dRdu += Export(ghost dRdu, ADD)            src/goal_sol_info.cpp:41-43
//   C6  PCU_Add_Doubles        scalar all-reduce of the error bound             src/goal_error.cpp:54
//
// Model.  Every part assembles its own elements into rows of all the nodes it holds (ghost layout).
// A node on a part boundary has partial rows on every part that holds it; its owner must end up with
// the sum.  The owner's row can need columns that the owner's own elements never touch (a neighbour of
// the node that lives only on another part), so at setup each non-owner tells the owner the global
// column ids of the rows it will send ("structure exchange", the INSERT export).  The owner appends the
// missing columns ("phantom" blocks) to the END of that node's block row.  Appending keeps every local
// block position -- and with it the element scatter map -- unchanged; only the row offsets move.  After
// that a value exchange is: non-owner packs its contiguous block rows, owner adds them through a
// precomputed block map, peers in ascending rank order (deterministic, == Tpetra Export/ADD).
//
// Transport is separable: gx_struct_* / gx_pack_interface / gx_unpack_add_interface move opaque
// buffers that any host transport can carry (MPI in the reference, torch.distributed in the tests);
// gx_comm_init + gx_reduce_interfaces do the same over NCCL send/recv on the context's stream.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <unordered_map>

#include "gx_internal.h"

using namespace gx;

#define GX_CUDA(call)                                                    \
  do {                                                                   \
    cudaError_t e_ = (call);                                             \
    if (e_ != cudaSuccess) {                                             \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);     \
      return GX_ERR_CUDA;                                                \
    }                                                                    \
  } while (0)

// ---------------------------------------------------------------------------
// NCCL through dlopen: the library must load (and the CPU-side tests must run) without NCCL, and
// inside a torch process it must bind to the libnccl.so.2 torch already loaded.
// ---------------------------------------------------------------------------
namespace gx {
struct Uid { char b[128]; };  // ncclUniqueId
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Uid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
}  // namespace gx
static constexpr int kNcclFloat64 = 8, kNcclSum = 0, kNcclInt64 = 4;

static NcclApi* load_nccl(std::string& err) {
  static NcclApi api;
  if (api.h) return &api;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { err = std::string("cannot load libnccl: ") + dlerror(); return nullptr; }
  auto sym = [&](const char* n) { return dlsym(h, n); };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.Send = (decltype(api.Send))sym("ncclSend");
  api.Recv = (decltype(api.Recv))sym("ncclRecv");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.GroupStart || !api.GroupEnd || !api.AllReduce) {
    err = "libnccl lacks required symbols";
    return nullptr;
  }
  api.h = h;
  return &api;
}

#define GX_NCCL(call)                                                                                   \
  do {                                                                                                  \
    int r_ = (call);                                                                                    \
    if (r_ != 0) {                                                                                      \
      ctx->err = std::string(#call) + ": " + (ctx->nccl->GetErrorString ? ctx->nccl->GetErrorString(r_) : "nccl error"); \
      return GX_ERR_NCCL;                                                                               \
    }                                                                                                   \
  } while (0)

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
__global__ void set_blocks_kernel(NodeRec* nodes, int32_t const* blk0, int32_t const* nblk, int nn) {
  int const n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nn) return;
  nodes[n].blk0 = blk0[n];
  nodes[n].nblk = nblk[n];
}

// send side: node rows are contiguous (non-owned nodes carry no phantom blocks): plain segmented copy
__global__ void pack_rows_kernel(double* buf, double const* values, int32_t const* nodes, int64_t const* off,
                                 int32_t const* blk0, int32_t const* nblk_g, int nsend) {
  int const s = blockIdx.x;
  if (s >= nsend) return;
  int const a = nodes[s];
  double const* src = values + 16 * (int64_t)blk0[a];
  double* dst = buf + off[s];
  int const n = 16 * nblk_g[a];
  for (int t = threadIdx.x; t < n; t += blockDim.x) dst[t] = src[t];
}
__global__ void pack_R_kernel(double* buf, double const* R, int32_t const* nodes, int nsend) {
  int const t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4 * nsend) return;
  buf[t] = R[4 * (int64_t)nodes[t >> 2] + (t & 3)];
}
// receive side: sender's block j of node s goes to block map[moff[s] + j] of the owner's (extended) row
__global__ void unpack_rows_kernel(double* values, double const* buf, int32_t const* nodes, int64_t const* off,
                                   int32_t const* cnt, int32_t const* map, int64_t const* moff, int32_t const* blk0,
                                   int32_t const* nblk_x, int nrecv) {
  int const s = blockIdx.x;
  if (s >= nrecv) return;
  int const a = nodes[s];
  int const nb_s = cnt[s];          // sender's blocks in this row
  int const rl_s = 4 * nb_s;        // sender's dof-row length
  int64_t const rl = 4 * (int64_t)nblk_x[a];
  double* dst = values + 16 * (int64_t)blk0[a];
  double const* src = buf + off[s];
  int32_t const* m = map + moff[s];
  for (int t = threadIdx.x; t < 16 * nb_s; t += blockDim.x) {
    int const i = t / rl_s, c = t - i * rl_s;  // dof row, column within the sender's row
    dst[i * rl + 4 * (int64_t)m[c >> 2] + (c & 3)] += src[t];
  }
}
__global__ void unpack_R_kernel(double* R, double const* buf, int32_t const* nodes, int nrecv) {
  int const t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4 * nrecv) return;
  R[4 * (int64_t)nodes[t >> 2] + (t & 3)] += buf[t];
}
// apf::synchronize of "u","p" (src/goal_disc.cpp:420-421): the owner's nodal values overwrite the copies
__global__ void pack_sol_kernel(double* buf, NodeRec const* nd, int32_t const* nodes, int n) {
  int const t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  NodeRec const& r = nd[nodes[t]];
  buf[4 * (int64_t)t] = r.u[0]; buf[4 * (int64_t)t + 1] = r.u[1]; buf[4 * (int64_t)t + 2] = r.u[2]; buf[4 * (int64_t)t + 3] = r.p;
}
__global__ void unpack_sol_kernel(NodeRec* nd, double const* buf, int32_t const* nodes, int n) {
  int const t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  NodeRec& r = nd[nodes[t]];
  r.u[0] = buf[4 * (int64_t)t]; r.u[1] = buf[4 * (int64_t)t + 1]; r.u[2] = buf[4 * (int64_t)t + 2]; r.p = buf[4 * (int64_t)t + 3];
}
// extended rows -> reference ghost layout (drop the phantom blocks)
__global__ void compact_ghost_kernel(double* ghost, double const* ext, int32_t const* blk0_x, int32_t const* nblk_x,
                                     int64_t const* blk0_g, int nn) {
  int const a = blockIdx.x;
  if (a >= nn) return;
  int const nb_g = (int)(blk0_g[a + 1] - blk0_g[a]);
  int64_t const rl_x = 4 * (int64_t)nblk_x[a];
  int const rl_g = 4 * nb_g;
  double const* src = ext + 16 * (int64_t)blk0_x[a];
  double* dst = ghost + 16 * blk0_g[a];
  for (int t = threadIdx.x; t < 16 * nb_g; t += blockDim.x) {
    int const i = t / rl_g, c = t - i * rl_g;
    dst[t] = src[i * rl_x + c];
  }
}

// ---------------------------------------------------------------------------
namespace gx {

void comm_destroy(gx_ctx* ctx) {
  if (ctx->device >= 0) cudaSetDevice(ctx->device);
  for (auto& p : ctx->peers) {
    void* ptrs[] = {p.d_send_nodes, p.d_recv_nodes, p.d_send_off, p.d_recv_off, p.d_recv_map, p.d_recv_moff,
                    p.d_recv_cnt, p.d_send, p.d_recv, p.d_sendR, p.d_recvR};
    for (void* q : ptrs) if (q) cudaFree(q);
  }
  void* ptrs[] = {ctx->d_blk0_x, ctx->d_nblk_x, ctx->d_nblk_g, ctx->d_blk0_g, ctx->d_ghost_vals};
  for (void* q : ptrs) if (q) cudaFree(q);
  if (ctx->comm && ctx->nccl && ctx->nccl->CommDestroy) ctx->nccl->CommDestroy(ctx->comm);
  ctx->comm = nullptr;
}

int comm_setup_lists(gx_ctx* c, const gx_desc* d) {
  int const nn = c->nn;
  c->nrow_x = c->nrow;
  c->nnz_x = c->nnz;
  c->owned_nodes.clear();
  if (d->n_ranks <= 1) {
    c->struct_done = true;
    return GX_OK;
  }
  if (!d->node_gid || !d->node_owner || d->n_peers < 0 || (d->n_peers > 0 && (!d->peer_rank || !d->peer_offset || !d->peer_nodes))) {
    c->err = "partitioned context needs node_gid, node_owner and the peer lists";
    return GX_ERR_ARG;
  }
  c->node_gid.assign(d->node_gid, d->node_gid + nn);
  c->node_owner.assign(d->node_owner, d->node_owner + nn);
  std::vector<char> listed(nn, 0);
  c->peers.resize(d->n_peers);
  for (int p = 0; p < d->n_peers; ++p) {
    Peer& P = c->peers[p];
    P.rank = d->peer_rank[p];
    if (P.rank < 0 || P.rank >= d->n_ranks || P.rank == c->rank) { c->err = "bad peer rank"; return GX_ERR_ARG; }
    if (p > 0 && d->peer_rank[p] <= d->peer_rank[p - 1]) { c->err = "peer ranks must be ascending"; return GX_ERR_ARG; }
    for (int k = d->peer_offset[p]; k < d->peer_offset[p + 1]; ++k) {
      int const a = d->peer_nodes[k];
      if (a < 0 || a >= nn) { c->err = "peer node out of range"; return GX_ERR_ARG; }
      P.nodes.push_back(a);
      if (c->node_owner[a] == P.rank) { P.send_nodes.push_back(a); listed[a] = 1; }
      else if (c->node_owner[a] == c->rank) P.recv_nodes.push_back(a);
    }
  }
  for (int a = 0; a < nn; ++a) {
    if (c->node_owner[a] < 0 || c->node_owner[a] >= d->n_ranks) { c->err = "node_owner out of range"; return GX_ERR_ARG; }
    if (c->node_owner[a] != c->rank && !listed[a]) { c->err = "a node owned by another rank is missing from that peer's list"; return GX_ERR_ARG; }
  }
  c->struct_done = false;
  return GX_OK;
}

}  // namespace gx

static int check_peer(gx_ctx* ctx, int peer_index) {
  if (!ctx) return GX_ERR_ARG;
  if (peer_index < 0 || peer_index >= (int)ctx->peers.size()) { ctx->err = "peer index out of range"; return GX_ERR_ARG; }
  return GX_OK;
}

// upload the exchange plan and the extended block layout
static int upload_plan(gx_ctx* ctx) {
  if (ctx->device < 0) return GX_OK;
  GX_CUDA(cudaSetDevice(ctx->device));
  int const nn = ctx->nn;
  std::vector<int32_t> b0(nn), nbx(nn), nbg(nn);
  for (int a = 0; a < nn; ++a) {
    b0[a] = (int32_t)ctx->nrow_x[a];
    nbx[a] = (int32_t)(ctx->nrow_x[a + 1] - ctx->nrow_x[a]);
    nbg[a] = (int32_t)(ctx->nrow[a + 1] - ctx->nrow[a]);
  }
  auto up = [&](auto*& dptr, auto const& v) -> cudaError_t {
    using T = typename std::remove_reference<decltype(v)>::type::value_type;
    if (dptr) { cudaFree(dptr); dptr = nullptr; }
    if (v.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc(&dptr, sizeof(T) * v.size());
    if (e != cudaSuccess) return e;
    return cudaMemcpy(dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
  };
  GX_CUDA(up(ctx->d_blk0_x, b0));
  GX_CUDA(up(ctx->d_nblk_x, nbx));
  GX_CUDA(up(ctx->d_nblk_g, nbg));
  GX_CUDA(up(ctx->d_blk0_g, ctx->nrow));
  set_blocks_kernel<<<(nn + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_nodes, ctx->d_blk0_x, ctx->d_nblk_x, nn);
  GX_CUDA(cudaGetLastError());
  GX_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->nnz_x != ctx->nnz) {
    if (ctx->d_values) cudaFree(ctx->d_values);
    ctx->d_values = nullptr;
    GX_CUDA(cudaMalloc(&ctx->d_values, sizeof(double) * (size_t)ctx->nnz_x));
    GX_CUDA(cudaMemset(ctx->d_values, 0, sizeof(double) * (size_t)ctx->nnz_x));
  }
  for (auto& P : ctx->peers) {
    GX_CUDA(up(P.d_send_nodes, P.send_nodes));
    GX_CUDA(up(P.d_recv_nodes, P.recv_nodes));
    GX_CUDA(up(P.d_send_off, P.send_off));
    GX_CUDA(up(P.d_recv_off, P.recv_off));
    GX_CUDA(up(P.d_recv_map, P.recv_map));
    GX_CUDA(up(P.d_recv_moff, P.recv_moff));
    GX_CUDA(up(P.d_recv_cnt, P.recv_cnt));
    if (P.send_vals) GX_CUDA(cudaMalloc(&P.d_send, sizeof(double) * (size_t)P.send_vals));
    if (P.recv_vals) GX_CUDA(cudaMalloc(&P.d_recv, sizeof(double) * (size_t)P.recv_vals));
    if (!P.send_nodes.empty()) GX_CUDA(cudaMalloc(&P.d_sendR, sizeof(double) * 4 * P.send_nodes.size()));
    if (!P.recv_nodes.empty()) GX_CUDA(cudaMalloc(&P.d_recvR, sizeof(double) * 4 * P.recv_nodes.size()));
  }
  return GX_OK;
}

extern "C" {

int gx_struct_pack(gx_ctx* ctx, int peer_index, const void** blob, int64_t* bytes) {
  int rc = check_peer(ctx, peer_index);
  if (rc) return rc;
  if (!blob || !bytes) return GX_ERR_ARG;
  Peer& P = ctx->peers[peer_index];
  P.struct_out.clear();
  for (int a : P.send_nodes) {
    int64_t const nb = ctx->nrow[a + 1] - ctx->nrow[a];
    P.struct_out.push_back(nb);
    for (int64_t k = ctx->nrow[a]; k < ctx->nrow[a + 1]; ++k) {  // (global column node, its owning rank)
      P.struct_out.push_back(ctx->node_gid[ctx->ncol[k]]);
      P.struct_out.push_back((int64_t)ctx->node_owner[ctx->ncol[k]]);
    }
  }
  *blob = P.struct_out.data();
  *bytes = (int64_t)(P.struct_out.size() * sizeof(int64_t));
  return GX_OK;
}

int gx_struct_unpack(gx_ctx* ctx, int peer_index, const void* blob, int64_t bytes) {
  int rc = check_peer(ctx, peer_index);
  if (rc) return rc;
  if ((bytes && !blob) || bytes % 8) { ctx->err = "bad structure blob"; return GX_ERR_ARG; }
  Peer& P = ctx->peers[peer_index];
  int64_t const n = bytes / 8;
  int64_t const* w = static_cast<int64_t const*>(blob);
  P.struct_in.assign(w, w + n);
  // validate framing against the receive list
  int64_t pos = 0;
  for (size_t s = 0; s < P.recv_nodes.size(); ++s) {
    if (pos >= n || w[pos] < 0 || pos + 1 + 2 * w[pos] > n) { ctx->err = "structure blob does not match the shared-node list"; return GX_ERR_ARG; }
    pos += 1 + 2 * w[pos];
  }
  if (pos != n) { ctx->err = "structure blob has trailing data"; return GX_ERR_ARG; }
  P.struct_have = true;
  return GX_OK;
}

int gx_struct_finalize(gx_ctx* ctx) {
  if (!ctx) return GX_ERR_ARG;
  if (ctx->struct_done) return GX_OK;
  int const nn = ctx->nn;
  for (auto& P : ctx->peers)
    if (!P.recv_nodes.empty() && !P.struct_have) { ctx->err = "gx_struct_finalize: structure of peer " + std::to_string(P.rank) + " not received"; return GX_ERR_ARG; }
  std::unordered_map<int64_t, int32_t> g2l;
  g2l.reserve(nn * 2);
  for (int a = 0; a < nn; ++a) g2l[ctx->node_gid[a]] = a;
  // phantom columns per owned node: global ids the senders have and this part's row lacks
  std::vector<std::vector<int64_t>> phantom(nn);
  std::unordered_map<int64_t, int32_t> phantom_owner;  // owning rank of every column that lives only on other parts
  auto local_pos = [&](int a, int64_t gid) -> int {
    auto it = g2l.find(gid);
    if (it == g2l.end()) return -1;
    int32_t const* b = ctx->ncol.data() + ctx->nrow[a];
    int32_t const* e = ctx->ncol.data() + ctx->nrow[a + 1];
    int32_t const* f = std::lower_bound(b, e, it->second);
    return (f != e && *f == it->second) ? (int)(f - b) : -1;
  };
  for (auto& P : ctx->peers) {
    int64_t pos = 0;
    for (int a : P.recv_nodes) {
      int64_t const nb = P.struct_in[pos++];
      for (int64_t k = 0; k < nb; ++k) {
        int64_t const gid = P.struct_in[pos++];
        int64_t const own = P.struct_in[pos++];
        if (own < 0 || own >= ctx->nranks) { ctx->err = "structure blob: column owner out of range"; return GX_ERR_ARG; }
        if (local_pos(a, gid) < 0) { phantom[a].push_back(gid); phantom_owner[gid] = (int32_t)own; }
      }
    }
  }
  ctx->nrow_x.assign(nn + 1, 0);
  ctx->block_lists_built = false;
  ctx->patch_state = 0;
  for (int a = 0; a < nn; ++a) {
    auto& ph = phantom[a];
    std::sort(ph.begin(), ph.end());
    ph.erase(std::unique(ph.begin(), ph.end()), ph.end());
    int64_t const nb = ctx->nrow[a + 1] - ctx->nrow[a] + (int64_t)ph.size();
    if (nb > 255) { ctx->err = "an owned interface node has more than 255 block columns"; return GX_ERR_UNSUPPORTED; }
    ctx->nrow_x[a + 1] = ctx->nrow_x[a] + nb;
    ctx->max_nblk = std::max<int>(ctx->max_nblk, (int)nb);
  }
  if (ctx->nrow_x[nn] > 0x7fffffffLL) { ctx->err = "more than 2^31 node blocks"; return GX_ERR_UNSUPPORTED; }
  ctx->nnz_x = 16 * ctx->nrow_x[nn];
  ctx->xcol_gid.resize(ctx->nrow_x[nn]);
  ctx->xcol_owner.resize(ctx->nrow_x[nn]);
  for (int a = 0; a < nn; ++a) {
    int64_t o = ctx->nrow_x[a];
    for (int64_t k = ctx->nrow[a]; k < ctx->nrow[a + 1]; ++k) { ctx->xcol_gid[o] = ctx->node_gid[ctx->ncol[k]]; ctx->xcol_owner[o++] = ctx->node_owner[ctx->ncol[k]]; }
    for (int64_t g : phantom[a]) { ctx->xcol_gid[o] = g; ctx->xcol_owner[o++] = phantom_owner[g]; }
  }
  // exchange plan
  for (auto& P : ctx->peers) {
    P.send_off.assign(1, 0);
    for (int a : P.send_nodes) P.send_off.push_back(P.send_off.back() + 16 * (ctx->nrow[a + 1] - ctx->nrow[a]));
    P.send_vals = P.send_off.back();
    P.recv_off.assign(1, 0);
    P.recv_moff.assign(1, 0);
    P.recv_cnt.clear();
    P.recv_map.clear();
    int64_t pos = 0;
    for (int a : P.recv_nodes) {
      int64_t const nb = P.struct_in[pos++];
      int64_t const nloc = ctx->nrow[a + 1] - ctx->nrow[a];
      for (int64_t k = 0; k < nb; ++k) {
        int64_t const gid = P.struct_in[pos++];
        ++pos;  // the column's owner (used above)
        int lp = local_pos(a, gid);
        if (lp < 0) lp = (int)(nloc + (std::lower_bound(phantom[a].begin(), phantom[a].end(), gid) - phantom[a].begin()));
        P.recv_map.push_back(lp);
      }
      P.recv_cnt.push_back((int32_t)nb);
      P.recv_off.push_back(P.recv_off.back() + 16 * nb);
      P.recv_moff.push_back(P.recv_moff.back() + nb);
    }
    P.recv_vals = P.recv_off.back();
  }
  // owned view: nodes this rank owns, ascending local id (apf::numberOwnedNodes keeps mesh order)
  ctx->owned_nodes.clear();
  for (int a = 0; a < nn; ++a)
    if (ctx->node_owner[a] == ctx->rank) ctx->owned_nodes.push_back(a);
  ctx->owned_rowptr.clear();
  ctx->owned_colgid.clear();
  ctx->tp_rowptr.clear();
  ctx->struct_done = true;
  return upload_plan(ctx);
}

int gx_owned_graph(gx_ctx* ctx, int32_t* n_owned_nodes, const int32_t** owned_nodes, int64_t* nnz_owned,
                   const int64_t** rowptr, const int64_t** col_gid) {
  if (!ctx) return GX_ERR_ARG;
  if (!ctx->struct_done) { ctx->err = "gx_owned_graph: structure exchange not finished"; return GX_ERR_ARG; }
  if (ctx->owned_nodes.empty() && ctx->nranks <= 1) {
    ctx->owned_nodes.resize(ctx->nn);
    for (int a = 0; a < ctx->nn; ++a) ctx->owned_nodes[a] = a;
  }
  if (ctx->owned_rowptr.empty()) {
    size_t const no = ctx->owned_nodes.size();
    ctx->owned_rowptr.assign(4 * no + 1, 0);
    for (size_t s = 0; s < no; ++s) {
      int const a = ctx->owned_nodes[s];
      int64_t const len = 4 * (ctx->nrow_x[a + 1] - ctx->nrow_x[a]);
      for (int i = 0; i < 4; ++i) ctx->owned_rowptr[4 * s + i + 1] = ctx->owned_rowptr[4 * s + i] + len;
    }
    ctx->owned_colgid.resize(ctx->owned_rowptr.back());
    for (size_t s = 0; s < no; ++s) {
      int const a = ctx->owned_nodes[s];
      for (int i = 0; i < 4; ++i) {
        int64_t* dst = ctx->owned_colgid.data() + ctx->owned_rowptr[4 * s + i];
        for (int64_t k = ctx->nrow_x[a]; k < ctx->nrow_x[a + 1]; ++k) {
          int64_t const g = ctx->nranks > 1 ? ctx->xcol_gid[k] : (int64_t)ctx->ncol[k];
          for (int c = 0; c < 4; ++c) *dst++ = 4 * g + c;
        }
      }
    }
  }
  if (n_owned_nodes) *n_owned_nodes = (int32_t)ctx->owned_nodes.size();
  if (owned_nodes) *owned_nodes = ctx->owned_nodes.data();
  if (nnz_owned) *nnz_owned = ctx->owned_rowptr.back();
  if (rowptr) *rowptr = ctx->owned_rowptr.data();
  if (col_gid) *col_gid = ctx->owned_colgid.data();
  return GX_OK;
}

// ---------------------------------------------------------------------------
// The owned matrix in Tpetra's LOCAL layout, i.e. what sol_info->owned->dRdu holds after
// owned_graph->fillComplete() (src/goal_disc.cpp:327-332) and gather_dRdu (src/goal_sol_info.cpp:41-43), so that
// the values can be written straight into the matrix goal::solve consumes (src/goal_linear_solve.cpp:72-90).
// Restated rule (Tpetra::CrsGraph::fillComplete -> Tpetra::Details::makeColMap; Trilinos is not in /root/reference,
// version unpinned -- SURVEY.md 8c):
//   rows     owned dofs in owned_map order = owned nodes in apf's owned numbering (mesh order, i.e. ascending
//            overlap-local id restricted to owned nodes), dof = node * 4 + eq          (goal_disc.cpp:270-290)
//   columns  column map = (1) the dofs of the domain map (= owned_map) in owned_map order, then (2) the remote dofs
//            grouped by owning rank in ascending rank order and, inside a rank, in ascending global id;
//            local column index = position in that list
//   a row    its local column indices in ascending order
// ---------------------------------------------------------------------------
static void build_tpetra_view(gx_ctx* ctx) {
  if (!ctx->tp_rowptr.empty()) return;
  int const nn = ctx->nn;
  if (ctx->owned_nodes.empty() && ctx->nranks <= 1) {
    ctx->owned_nodes.resize(nn);
    for (int a = 0; a < nn; ++a) ctx->owned_nodes[a] = a;
  }
  size_t const no = ctx->owned_nodes.size();
  bool const parts = ctx->nranks > 1;
  auto gid_of = [&](int64_t k) { return parts ? ctx->xcol_gid[k] : (int64_t)ctx->ncol[k]; };
  auto own_of = [&](int64_t k) { return parts ? ctx->xcol_owner[k] : 0; };
  std::unordered_map<int64_t, int32_t> lid;  // global node -> column-map node index
  lid.reserve(2 * no);
  ctx->tp_colmap.clear();
  for (size_t s = 0; s < no; ++s) {
    int64_t const g = parts ? ctx->node_gid[ctx->owned_nodes[s]] : (int64_t)ctx->owned_nodes[s];
    lid[g] = (int32_t)s;
    ctx->tp_colmap.push_back(g);
  }
  std::vector<std::pair<int32_t, int64_t>> remote;  // (owner, gid)
  for (size_t s = 0; s < no; ++s) {
    int const a = ctx->owned_nodes[s];
    for (int64_t k = ctx->nrow_x[a]; k < ctx->nrow_x[a + 1]; ++k)
      if (own_of(k) != ctx->rank) remote.emplace_back(own_of(k), gid_of(k));
  }
  std::sort(remote.begin(), remote.end());
  remote.erase(std::unique(remote.begin(), remote.end()), remote.end());
  for (auto const& r : remote) { lid[r.second] = (int32_t)ctx->tp_colmap.size(); ctx->tp_colmap.push_back(r.second); }
  ctx->tp_rowptr.assign(4 * no + 1, 0);
  for (size_t s = 0; s < no; ++s) {
    int const a = ctx->owned_nodes[s];
    int64_t const len = 4 * (ctx->nrow_x[a + 1] - ctx->nrow_x[a]);
    for (int i = 0; i < 4; ++i) ctx->tp_rowptr[4 * s + i + 1] = ctx->tp_rowptr[4 * s + i] + len;
  }
  ctx->tp_colind.resize(ctx->tp_rowptr.back());
  ctx->tp_perm.resize(ctx->nrow_x[nn], 0);  // per stored block of an owned row: its position in the Tpetra-ordered row
  std::vector<std::pair<int32_t, int32_t>> ord;
  for (size_t s = 0; s < no; ++s) {
    int const a = ctx->owned_nodes[s];
    int64_t const b0 = ctx->nrow_x[a], nb = ctx->nrow_x[a + 1] - b0;
    ord.clear();
    for (int64_t k = 0; k < nb; ++k) ord.emplace_back(lid[gid_of(b0 + k)], (int32_t)k);
    std::sort(ord.begin(), ord.end());
    for (int64_t j = 0; j < nb; ++j) {
      ctx->tp_perm[b0 + ord[j].second] = (uint8_t)j;
      for (int i = 0; i < 4; ++i)
        for (int c = 0; c < 4; ++c) ctx->tp_colind[ctx->tp_rowptr[4 * s + i] + 4 * j + c] = 4 * ord[j].first + c;
    }
  }
}

// values of the owned rows, blocks permuted into the Tpetra order of each row: one thread block per owned node
__global__ void tpetra_rows_kernel(double* out, double const* values, int32_t const* owned, int64_t const* out_off, int32_t const* blk0,
                                   int32_t const* nblk, uint8_t const* perm, int n_owned) {
  int const s = blockIdx.x;
  if (s >= n_owned) return;
  int const a = owned[s];
  int const nb = nblk[a];
  double const* src = values + 16 * (int64_t)blk0[a];
  double* dst = out + out_off[s];
  for (int t = threadIdx.x; t < 16 * nb; t += blockDim.x) {
    int const i = t / (4 * nb), r = t - i * 4 * nb, k = r >> 2, c = r & 3;
    dst[i * 4 * nb + 4 * perm[blk0[a] + k] + c] = src[t];
  }
}

extern "C" int gx_owned_tpetra_graph(gx_ctx* ctx, int32_t* n_owned_nodes, int32_t* n_col_nodes, const int64_t** colmap_node_gid,
                                     int64_t* nnz_owned, const int64_t** rowptr, const int32_t** colind) {
  if (!ctx) return GX_ERR_ARG;
  if (!ctx->struct_done) { ctx->err = "gx_owned_tpetra_graph: structure exchange not finished"; return GX_ERR_ARG; }
  build_tpetra_view(ctx);
  if (n_owned_nodes) *n_owned_nodes = (int32_t)ctx->owned_nodes.size();
  if (n_col_nodes) *n_col_nodes = (int32_t)ctx->tp_colmap.size();
  if (colmap_node_gid) *colmap_node_gid = ctx->tp_colmap.data();
  if (nnz_owned) *nnz_owned = ctx->tp_rowptr.back();
  if (rowptr) *rowptr = ctx->tp_rowptr.data();
  if (colind) *colind = ctx->tp_colind.data();
  return GX_OK;
}

extern "C" int gx_fetch_owned_tpetra(gx_ctx* ctx, double* R_owned, double* values_owned) {
  if (!ctx) return GX_ERR_ARG;
  if (ctx->device < 0) { ctx->err = "gx_fetch_owned_tpetra: host-only context (device = -1) cannot compute"; return GX_ERR_CUDA; }
  if (!ctx->have_result || (values_owned && !ctx->have_values)) { ctx->err = "gx_fetch_owned_tpetra: no matching result on the device"; return GX_ERR_ARG; }
  if (!ctx->struct_done) { ctx->err = "gx_fetch_owned_tpetra: structure exchange not finished"; return GX_ERR_ARG; }
  build_tpetra_view(ctx);
  GX_CUDA(cudaSetDevice(ctx->device));
  int const no = (int)ctx->owned_nodes.size();
  if (R_owned)
    for (int s = 0; s < no;) {  // owned nodes come in runs of consecutive local ids
      int e = s;
      while (e + 1 < no && ctx->owned_nodes[e + 1] == ctx->owned_nodes[e] + 1) ++e;
      GX_CUDA(cudaMemcpyAsync(R_owned + 4 * (size_t)s, ctx->d_R + 4 * (size_t)ctx->owned_nodes[s], sizeof(double) * 4 * (size_t)(e + 1 - s),
                              cudaMemcpyDeviceToHost, ctx->stream));
      s = e + 1;
    }
  if (values_owned && no > 0) {
    int const nn = ctx->nn;
    std::vector<int32_t> b0(nn), nbx(nn);
    for (int a = 0; a < nn; ++a) { b0[a] = (int32_t)ctx->nrow_x[a]; nbx[a] = (int32_t)(ctx->nrow_x[a + 1] - ctx->nrow_x[a]); }
    std::vector<int64_t> off(no);
    for (int s = 0; s < no; ++s) off[s] = ctx->tp_rowptr[4 * (size_t)s];
    double* d_out = nullptr; int32_t *d_owned = nullptr, *d_b0 = nullptr, *d_nb = nullptr; int64_t* d_off = nullptr; uint8_t* d_perm = nullptr;
    auto body = [&]() -> int {
      GX_CUDA(cudaMalloc(&d_out, sizeof(double) * (size_t)ctx->tp_rowptr.back()));
      GX_CUDA(cudaMalloc(&d_owned, sizeof(int32_t) * (size_t)no));
      GX_CUDA(cudaMalloc(&d_b0, sizeof(int32_t) * (size_t)nn));
      GX_CUDA(cudaMalloc(&d_nb, sizeof(int32_t) * (size_t)nn));
      GX_CUDA(cudaMalloc(&d_off, sizeof(int64_t) * (size_t)no));
      GX_CUDA(cudaMalloc(&d_perm, ctx->tp_perm.size()));
      GX_CUDA(cudaMemcpyAsync(d_owned, ctx->owned_nodes.data(), sizeof(int32_t) * (size_t)no, cudaMemcpyHostToDevice, ctx->stream));
      GX_CUDA(cudaMemcpyAsync(d_b0, b0.data(), sizeof(int32_t) * (size_t)nn, cudaMemcpyHostToDevice, ctx->stream));
      GX_CUDA(cudaMemcpyAsync(d_nb, nbx.data(), sizeof(int32_t) * (size_t)nn, cudaMemcpyHostToDevice, ctx->stream));
      GX_CUDA(cudaMemcpyAsync(d_off, off.data(), sizeof(int64_t) * (size_t)no, cudaMemcpyHostToDevice, ctx->stream));
      GX_CUDA(cudaMemcpyAsync(d_perm, ctx->tp_perm.data(), ctx->tp_perm.size(), cudaMemcpyHostToDevice, ctx->stream));
      tpetra_rows_kernel<<<no, 128, 0, ctx->stream>>>(d_out, ctx->d_values, d_owned, d_off, d_b0, d_nb, d_perm, no);
      GX_CUDA(cudaGetLastError());
      GX_CUDA(cudaMemcpyAsync(values_owned, d_out, sizeof(double) * (size_t)ctx->tp_rowptr.back(), cudaMemcpyDeviceToHost, ctx->stream));
      GX_CUDA(cudaStreamSynchronize(ctx->stream));
      return GX_OK;
    };
    int const rc = body();
    cudaStreamSynchronize(ctx->stream);
    for (void* q : {(void*)d_out, (void*)d_owned, (void*)d_b0, (void*)d_nb, (void*)d_off, (void*)d_perm}) if (q) cudaFree(q);
    if (rc) return rc;
  }
  GX_CUDA(cudaStreamSynchronize(ctx->stream));
  return GX_OK;
}

// Exchange plan of one peer, for hosts that emulate or verify the exchange:
// what[0]=n_send what[1]=n_recv; pointers stay valid until gx_destroy.
int gx_exchange_plan(gx_ctx* ctx, int peer_index, int32_t* peer_rank, int32_t counts[2], const int32_t** send_nodes,
                     const int32_t** recv_nodes, const int32_t** recv_cnt, const int32_t** recv_map) {
  int rc = check_peer(ctx, peer_index);
  if (rc) return rc;
  if (!ctx->struct_done) { ctx->err = "gx_exchange_plan: structure exchange not finished"; return GX_ERR_ARG; }
  Peer& P = ctx->peers[peer_index];
  if (peer_rank) *peer_rank = P.rank;
  if (counts) { counts[0] = (int32_t)P.send_nodes.size(); counts[1] = (int32_t)P.recv_nodes.size(); }
  if (send_nodes) *send_nodes = P.send_nodes.data();
  if (recv_nodes) *recv_nodes = P.recv_nodes.data();
  if (recv_cnt) *recv_cnt = P.recv_cnt.data();
  if (recv_map) *recv_map = P.recv_map.data();
  return GX_OK;
}

int gx_num_peers(gx_ctx* ctx, int32_t* n) {
  if (!ctx || !n) return GX_ERR_ARG;
  *n = (int32_t)ctx->peers.size();
  return GX_OK;
}

static int need_device(gx_ctx* ctx, const char* fn) {
  if (ctx->device < 0) { ctx->err = std::string(fn) + ": host-only context (device = -1) cannot compute"; return GX_ERR_CUDA; }
  return GX_OK;
}

int gx_interface_bytes(gx_ctx* ctx, int peer_index, int what, int64_t* send_bytes, int64_t* recv_bytes) {
  int rc = check_peer(ctx, peer_index);
  if (rc) return rc;
  if (!ctx->struct_done) { ctx->err = "structure exchange not finished"; return GX_ERR_ARG; }
  Peer& P = ctx->peers[peer_index];
  int64_t s = 0, r = 0;
  if (what & 5) { s += 32 * (int64_t)P.send_nodes.size(); r += 32 * (int64_t)P.recv_nodes.size(); }
  if (what & 2) { s += 8 * P.send_vals; r += 8 * P.recv_vals; }
  if (send_bytes) *send_bytes = s;
  if (recv_bytes) *recv_bytes = r;
  return GX_OK;
}

static int pack_peer(gx_ctx* ctx, Peer& P, int what, cudaStream_t st) {
  int const ns = (int)P.send_nodes.size();
  if (!ns) return GX_OK;
  // what = 4: dMdu of the last gx_functional travels like R (gather_dMdu, goal_sol_info.cpp:37-39)
  if (what & 5) pack_R_kernel<<<(4 * ns + 255) / 256, 256, 0, st>>>(P.d_sendR, (what & 4) ? ctx->d_dMdu : ctx->d_R, P.d_send_nodes, ns);
  if (what & 2) pack_rows_kernel<<<ns, 128, 0, st>>>(P.d_send, ctx->d_values, P.d_send_nodes, P.d_send_off, ctx->d_blk0_x, ctx->d_nblk_g, ns);
  GX_CUDA(cudaGetLastError());
  return GX_OK;
}
static int unpack_peer(gx_ctx* ctx, Peer& P, int what, double const* bufR, double const* bufV, cudaStream_t st) {
  int const nr = (int)P.recv_nodes.size();
  if (!nr) return GX_OK;
  if (what & 5) unpack_R_kernel<<<(4 * nr + 255) / 256, 256, 0, st>>>((what & 4) ? ctx->d_dMdu : ctx->d_R, bufR, P.d_recv_nodes, nr);
  if (what & 2)
    unpack_rows_kernel<<<nr, 128, 0, st>>>(ctx->d_values, bufV, P.d_recv_nodes, P.d_recv_off, P.d_recv_cnt, P.d_recv_map,
                                           P.d_recv_moff, ctx->d_blk0_x, ctx->d_nblk_x, nr);
  GX_CUDA(cudaGetLastError());
  return GX_OK;
}
// SolInfo::gather_* over NCCL, enqueued on `st`: pack, grouped send/recv of the packed interface rows, then the
// owner adds peer by peer in ascending rank order.  No synchronisation here.
}  // extern "C"
namespace gx {
int comm_enqueue_reduce(gx_ctx* ctx, int what, cudaStream_t st) {
  int rc;
  for (auto& P : ctx->peers) if ((rc = pack_peer(ctx, P, what, st))) return rc;
  GX_NCCL(ctx->nccl->GroupStart());
  for (auto& P : ctx->peers) {
    size_t const ns = P.send_nodes.size(), nr = P.recv_nodes.size();
    if ((what & 5) && ns) GX_NCCL(ctx->nccl->Send(P.d_sendR, 4 * ns, kNcclFloat64, P.rank, ctx->comm, st));
    if ((what & 5) && nr) GX_NCCL(ctx->nccl->Recv(P.d_recvR, 4 * nr, kNcclFloat64, P.rank, ctx->comm, st));
    if ((what & 2) && P.send_vals) GX_NCCL(ctx->nccl->Send(P.d_send, (size_t)P.send_vals, kNcclFloat64, P.rank, ctx->comm, st));
    if ((what & 2) && P.recv_vals) GX_NCCL(ctx->nccl->Recv(P.d_recv, (size_t)P.recv_vals, kNcclFloat64, P.rank, ctx->comm, st));
  }
  GX_NCCL(ctx->nccl->GroupEnd());
  for (auto& P : ctx->peers) if ((rc = unpack_peer(ctx, P, what, P.d_recvR, P.d_recv, st))) return rc;  // ascending rank
  return GX_OK;
}
}  // namespace gx
extern "C" {

// what must be 1 (R, 4 doubles per send node) or 2 (CRS rows); *send_dev is valid until the next pack
int gx_pack_interface(gx_ctx* ctx, int peer_index, int what, void** send_dev) {
  int rc = check_peer(ctx, peer_index);
  if (rc) return rc;
  if ((rc = need_device(ctx, "gx_pack_interface"))) return rc;
  if (what != 1 && what != 2 && what != 4) { ctx->err = "what must be 1 (R), 2 (dRdu) or 4 (dMdu)"; return GX_ERR_ARG; }
  if (what == 4 && !ctx->have_dMdu) { ctx->err = "no functional derivative on the device"; return GX_ERR_ARG; }
  Peer& P = ctx->peers[peer_index];
  GX_CUDA(cudaSetDevice(ctx->device));
  if ((rc = pack_peer(ctx, P, what, ctx->stream))) return rc;
  GX_CUDA(cudaStreamSynchronize(ctx->stream));
  if (send_dev) *send_dev = what != 2 ? (void*)P.d_sendR : (void*)P.d_send;
  return GX_OK;
}

int gx_unpack_add_interface(gx_ctx* ctx, int peer_index, int what, const void* recv_dev) {
  int rc = check_peer(ctx, peer_index);
  if (rc) return rc;
  if ((rc = need_device(ctx, "gx_unpack_add_interface"))) return rc;
  if (what != 1 && what != 2 && what != 4) { ctx->err = "what must be 1 (R), 2 (dRdu) or 4 (dMdu)"; return GX_ERR_ARG; }
  if (what == 4 && !ctx->have_dMdu) { ctx->err = "no functional derivative on the device"; return GX_ERR_ARG; }
  Peer& P = ctx->peers[peer_index];
  GX_CUDA(cudaSetDevice(ctx->device));
  if ((rc = unpack_peer(ctx, P, what, (double const*)recv_dev, (double const*)recv_dev, ctx->stream))) return rc;
  GX_CUDA(cudaStreamSynchronize(ctx->stream));
  return GX_OK;
}

// Solution synchronisation, owner -> copies: the reverse direction of the R exchange, so the owner packs its
// recv_nodes (into the R receive buffer) and the copy holder overwrites its send_nodes.
int gx_pack_solution(gx_ctx* ctx, int peer_index, void** send_dev, int64_t* send_bytes) {
  int rc = check_peer(ctx, peer_index);
  if (rc) return rc;
  if ((rc = need_device(ctx, "gx_pack_solution"))) return rc;
  Peer& P = ctx->peers[peer_index];
  int const n = (int)P.recv_nodes.size();
  GX_CUDA(cudaSetDevice(ctx->device));
  if (n) pack_sol_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(P.d_recvR, ctx->d_nodes, P.d_recv_nodes, n);
  GX_CUDA(cudaGetLastError());
  GX_CUDA(cudaStreamSynchronize(ctx->stream));
  if (send_dev) *send_dev = P.d_recvR;
  if (send_bytes) *send_bytes = 32 * (int64_t)n;
  return GX_OK;
}
int gx_unpack_solution(gx_ctx* ctx, int peer_index, const void* recv_dev) {
  int rc = check_peer(ctx, peer_index);
  if (rc) return rc;
  if ((rc = need_device(ctx, "gx_unpack_solution"))) return rc;
  Peer& P = ctx->peers[peer_index];
  int const n = (int)P.send_nodes.size();
  if (!n) return GX_OK;
  if (!recv_dev) { ctx->err = "gx_unpack_solution: null buffer"; return GX_ERR_ARG; }
  GX_CUDA(cudaSetDevice(ctx->device));
  unpack_sol_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_nodes, (double const*)recv_dev, P.d_send_nodes, n);
  GX_CUDA(cudaGetLastError());
  GX_CUDA(cudaStreamSynchronize(ctx->stream));
  return GX_OK;
}

int gx_nccl_unique_id(void* out, size_t* id_bytes) {
  std::string err;
  NcclApi* api = load_nccl(err);
  if (!api || !out) return GX_ERR_NCCL;
  Uid id;
  if (api->GetUniqueId(&id) != 0) return GX_ERR_NCCL;
  memcpy(out, &id, sizeof id);
  if (id_bytes) *id_bytes = sizeof id;
  return GX_OK;
}

// structure exchange over NCCL: lengths first, then the blobs (int64 words).  Temporary device buffers are owned by
// the caller's `bufs` so that every exit path frees them.
static int struct_exchange_nccl(gx_ctx* ctx, std::vector<void*>& bufs) {
  int rc;
  size_t const np = ctx->peers.size();
  std::vector<int64_t> slen(np), rlen(np, 0);
  for (size_t p = 0; p < np; ++p) {
    const void* b; int64_t bytes;
    if ((rc = gx_struct_pack(ctx, (int)p, &b, &bytes))) return rc;
    slen[p] = bytes / 8;
  }
  auto dev_alloc = [&](int64_t** q, size_t bytes) -> cudaError_t {
    cudaError_t const e = cudaMalloc(q, bytes);
    if (e == cudaSuccess) bufs.push_back(*q);
    return e;
  };
  int64_t *d_s = nullptr, *d_r = nullptr;
  GX_CUDA(dev_alloc(&d_s, 8 * std::max<size_t>(np, 1)));
  GX_CUDA(dev_alloc(&d_r, 8 * std::max<size_t>(np, 1)));
  GX_CUDA(cudaMemcpyAsync(d_s, slen.data(), 8 * np, cudaMemcpyHostToDevice, ctx->stream));
  GX_NCCL(ctx->nccl->GroupStart());
  for (size_t p = 0; p < np; ++p) {
    GX_NCCL(ctx->nccl->Send(d_s + p, 1, kNcclInt64, ctx->peers[p].rank, ctx->comm, ctx->stream));
    GX_NCCL(ctx->nccl->Recv(d_r + p, 1, kNcclInt64, ctx->peers[p].rank, ctx->comm, ctx->stream));
  }
  GX_NCCL(ctx->nccl->GroupEnd());
  GX_CUDA(cudaMemcpyAsync(rlen.data(), d_r, 8 * np, cudaMemcpyDeviceToHost, ctx->stream));
  GX_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<int64_t*> ds(np, nullptr), dr(np, nullptr);
  for (size_t p = 0; p < np; ++p) {
    if (slen[p]) { GX_CUDA(dev_alloc(&ds[p], 8 * slen[p])); GX_CUDA(cudaMemcpyAsync(ds[p], ctx->peers[p].struct_out.data(), 8 * slen[p], cudaMemcpyHostToDevice, ctx->stream)); }
    if (rlen[p]) GX_CUDA(dev_alloc(&dr[p], 8 * rlen[p]));
  }
  GX_NCCL(ctx->nccl->GroupStart());
  for (size_t p = 0; p < np; ++p) {
    if (slen[p]) GX_NCCL(ctx->nccl->Send(ds[p], slen[p], kNcclInt64, ctx->peers[p].rank, ctx->comm, ctx->stream));
    if (rlen[p]) GX_NCCL(ctx->nccl->Recv(dr[p], rlen[p], kNcclInt64, ctx->peers[p].rank, ctx->comm, ctx->stream));
  }
  GX_NCCL(ctx->nccl->GroupEnd());
  std::vector<std::vector<int64_t>> in(np);
  for (size_t p = 0; p < np; ++p) {
    in[p].resize(rlen[p]);
    if (rlen[p]) GX_CUDA(cudaMemcpyAsync(in[p].data(), dr[p], 8 * rlen[p], cudaMemcpyDeviceToHost, ctx->stream));
  }
  GX_CUDA(cudaStreamSynchronize(ctx->stream));
  for (size_t p = 0; p < np; ++p)
    if ((rc = gx_struct_unpack(ctx, (int)p, in[p].data(), 8 * rlen[p]))) return rc;
  return gx_struct_finalize(ctx);
}

int gx_comm_init(gx_ctx* ctx, const void* nccl_unique_id, size_t id_bytes) {
  if (!ctx || !nccl_unique_id || id_bytes != sizeof(Uid)) { if (ctx) ctx->err = "gx_comm_init: bad unique id"; return GX_ERR_ARG; }
  int rc = need_device(ctx, "gx_comm_init");
  if (rc) return rc;
  ctx->nccl = load_nccl(ctx->err);
  if (!ctx->nccl) return GX_ERR_NCCL;
  GX_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->comm) {  // a retry after a failed structure exchange keeps the live communicator
    Uid id;
    memcpy(&id, nccl_unique_id, sizeof id);
    GX_NCCL(ctx->nccl->CommInitRank(&ctx->comm, ctx->nranks, id, ctx->rank));
  }
  if (ctx->struct_done) return GX_OK;
  std::vector<void*> bufs;
  rc = struct_exchange_nccl(ctx, bufs);
  if (rc) cudaStreamSynchronize(ctx->stream);  // nothing may still be using the buffers
  for (void* q : bufs) cudaFree(q);
  return rc;
}

// SolInfo::gather_R / gather_dRdu over NCCL: grouped send/recv of the packed interface rows, then the
// owner adds peer by peer in ascending rank order.
int gx_reduce_interfaces(gx_ctx* ctx, int what) {
  if (!ctx) return GX_ERR_ARG;
  if (ctx->nranks <= 1 || ctx->peers.empty()) { ctx->timing[2] = 0.0; return GX_OK; }
  int rc = need_device(ctx, "gx_reduce_interfaces");
  if (rc) return rc;
  if (!ctx->comm || !ctx->struct_done) { ctx->err = "gx_reduce_interfaces: call gx_comm_init first"; return GX_ERR_ARG; }
  if (!(what & 7)) return GX_OK;
  if ((what & 4) && (what != 4 || !ctx->have_dMdu)) { ctx->err = "gx_reduce_interfaces: what = 4 (dMdu) goes alone, after gx_functional with a derivative"; return GX_ERR_ARG; }
  if ((what & 1) && !ctx->have_result) { ctx->err = "gx_reduce_interfaces: no residual on the device"; return GX_ERR_ARG; }
  if ((what & 2) && !ctx->have_values) { ctx->err = "gx_reduce_interfaces: what & 2 needs the CRS values of a Jacobian pass (the last pass left none)"; return GX_ERR_ARG; }
  GX_CUDA(cudaSetDevice(ctx->device));
  GX_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
  if ((rc = comm_enqueue_reduce(ctx, what, ctx->stream))) return rc;
  GX_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
  GX_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  GX_CUDA(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
  ctx->timing[2] = ms;
  return GX_OK;
}

// apf::synchronize(u), apf::synchronize(p) after Disc::add_soln (src/goal_disc.cpp:420-421) over NCCL
int gx_sync_solution(gx_ctx* ctx) {
  if (!ctx) return GX_ERR_ARG;
  if (ctx->nranks <= 1 || ctx->peers.empty()) return GX_OK;
  int rc = need_device(ctx, "gx_sync_solution");
  if (rc) return rc;
  if (!ctx->comm || !ctx->struct_done) { ctx->err = "gx_sync_solution: call gx_comm_init first"; return GX_ERR_ARG; }
  GX_CUDA(cudaSetDevice(ctx->device));
  for (auto& P : ctx->peers) {
    int const n = (int)P.recv_nodes.size();
    if (n) pack_sol_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(P.d_recvR, ctx->d_nodes, P.d_recv_nodes, n);
  }
  GX_CUDA(cudaGetLastError());
  GX_NCCL(ctx->nccl->GroupStart());
  for (auto& P : ctx->peers) {
    size_t const ns = P.send_nodes.size(), nr = P.recv_nodes.size();
    if (nr) GX_NCCL(ctx->nccl->Send(P.d_recvR, 4 * nr, kNcclFloat64, P.rank, ctx->comm, ctx->stream));
    if (ns) GX_NCCL(ctx->nccl->Recv(P.d_sendR, 4 * ns, kNcclFloat64, P.rank, ctx->comm, ctx->stream));
  }
  GX_NCCL(ctx->nccl->GroupEnd());
  for (auto& P : ctx->peers) {
    int const n = (int)P.send_nodes.size();
    if (n) unpack_sol_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_nodes, P.d_sendR, P.d_send_nodes, n);
  }
  GX_CUDA(cudaGetLastError());
  GX_CUDA(cudaStreamSynchronize(ctx->stream));
  return GX_OK;
}

int gx_allreduce_sum(gx_ctx* ctx, double* x, int n) {
  if (!ctx || !x || n < 0) return GX_ERR_ARG;
  if (ctx->nranks <= 1) return GX_OK;
  int rc = need_device(ctx, "gx_allreduce_sum");
  if (rc) return rc;
  if (!ctx->comm) { ctx->err = "gx_allreduce_sum: call gx_comm_init first"; return GX_ERR_ARG; }
  if (n > 1024) { ctx->err = "gx_allreduce_sum: n <= 1024"; return GX_ERR_ARG; }
  GX_CUDA(cudaSetDevice(ctx->device));
  GX_CUDA(cudaMemcpyAsync(ctx->d_red, x, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  GX_NCCL(ctx->nccl->AllReduce(ctx->d_red, ctx->d_red, (size_t)n, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream));
  GX_CUDA(cudaMemcpyAsync(x, ctx->d_red, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  GX_CUDA(cudaStreamSynchronize(ctx->stream));
  return GX_OK;
}

// owned rows (ascending owned node) of R and of the CRS values, extended-row layout of gx_owned_graph
int gx_fetch_owned(gx_ctx* ctx, double* R_owned, double* values_owned) {
  if (!ctx) return GX_ERR_ARG;
  int rc = need_device(ctx, "gx_fetch_owned");
  if (rc) return rc;
  if (!ctx->have_result) { ctx->err = "gx_fetch_owned: no result yet"; return GX_ERR_ARG; }
  int32_t no; const int32_t* on; int64_t nnzo; const int64_t* rp;
  if ((rc = gx_owned_graph(ctx, &no, &on, &nnzo, &rp, nullptr))) return rc;
  GX_CUDA(cudaSetDevice(ctx->device));
  // owned nodes come in runs of consecutive local ids: copy run by run
  for (int s = 0; s < no;) {
    int e = s;
    while (e + 1 < no && on[e + 1] == on[e] + 1) ++e;
    int const a0 = on[s], a1 = on[e] + 1;
    if (R_owned) GX_CUDA(cudaMemcpyAsync(R_owned + 4 * (size_t)s, ctx->d_R + 4 * (size_t)a0, sizeof(double) * 4 * (size_t)(a1 - a0), cudaMemcpyDeviceToHost, ctx->stream));
    if (values_owned && ctx->have_values)
      GX_CUDA(cudaMemcpyAsync(values_owned + rp[4 * (size_t)s], ctx->d_values + 16 * ctx->nrow_x[a0],
                              sizeof(double) * 16 * (size_t)(ctx->nrow_x[a1] - ctx->nrow_x[a0]), cudaMemcpyDeviceToHost, ctx->stream));
    s = e + 1;
  }
  GX_CUDA(cudaStreamSynchronize(ctx->stream));
  return GX_OK;
}

}  // extern "C"

// ghost-layout view of the (possibly extended) device values; used by gx_api.cu's fetch
namespace gx {
int ghost_values_dev(gx_ctx* ctx, double** out) {
  if (ctx->nnz_x == ctx->nnz) { *out = ctx->d_values; return GX_OK; }
  if (!ctx->d_ghost_vals) GX_CUDA(cudaMalloc(&ctx->d_ghost_vals, sizeof(double) * (size_t)ctx->nnz));
  compact_ghost_kernel<<<ctx->nn, 128, 0, ctx->stream>>>(ctx->d_ghost_vals, ctx->d_values, ctx->d_blk0_x, ctx->d_nblk_x, ctx->d_blk0_g, ctx->nn);
  GX_CUDA(cudaGetLastError());
  *out = ctx->d_ghost_vals;
  return GX_OK;
}
}  // namespace gx
