// gx_setup.cpp -- host-side, once-per-mesh setup: the CRS operator skeleton, the
// element -> nonzero scatter map and the conflict-free element schedule.
//
// Replaces Disc::compute_graphs (src/goal_disc.cpp:307-332), which inserts every
// (row dof, col dof) pair of every element one column at a time and lets Tpetra
// sort and merge at fillComplete.  Because dofs are node-blocked with 4 equations
// per node (src/goal_disc.cpp:195-201) the dof graph is the node adjacency graph
// with every entry expanded to a 4x4 block, so the graph is built at node level:
//   nrow/ncol : block-CRS of "nodes sharing an element", each row sorted
//   dof row 4a+i  = the same block row, columns 4b+k, b in ncol order, k = 0..3
// which is exactly Tpetra's sorted-unique local layout for the ghost graph (the
// ghost column map equals the ghost row map).
//
// The schedule is a greedy element colouring: two elements of one colour share no
// node, so their scatters into R and into the CRS values touch disjoint rows and
// the assembly needs no atomics and is bit-reproducible run to run.
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <numeric>
#include <omp.h>
#include <parallel/algorithm>

#include "gx_internal.h"

namespace gx {

int build_graph_and_schedule(gx_ctx* c) {
  SetupTimer tm;
  int const nn = c->nn, ne = c->ne;
  int32_t const* conn = c->conn.data();
  {
    int64_t const none = INT64_MAX;
    int64_t bad_range = none, bad_rep = none;  // first offending element of either kind (deterministic: min over threads)
#pragma omp parallel for schedule(static) reduction(min : bad_range, bad_rep)
    for (int e = 0; e < ne; ++e) {
      int32_t const* en = conn + 4 * (int64_t)e;
      bool const out = en[0] < 0 || en[0] >= nn || en[1] < 0 || en[1] >= nn || en[2] < 0 || en[2] >= nn || en[3] < 0 || en[3] >= nn;
      bool const rep = en[0] == en[1] || en[0] == en[2] || en[0] == en[3] || en[1] == en[2] || en[1] == en[3] || en[2] == en[3];
      if (out) bad_range = std::min<int64_t>(bad_range, e);
      if (rep) bad_rep = std::min<int64_t>(bad_rep, e);
    }
    if (bad_range != none) { c->err = "conn entry out of range"; return GX_ERR_ARG; }
    if (bad_rep != none) { c->err = "element " + std::to_string(bad_rep) + " repeats a node"; return GX_ERR_ARG; }
  }
  if ((int64_t)ne >= (1ll << 29)) { c->err = "more than 2^29 elements per part"; return GX_ERR_UNSUPPORTED; }

  tm.lap("validate");
  // ---- node -> elements (counting sort, elements ascending per node).  Parallel over contiguous element chunks with one
  //      histogram per chunk: chunk t's entries of a node go behind those of the chunks before it, so the result is
  //      the serial one whatever the number of chunks (which is capped to keep the histograms below 256 MB).
  std::vector<int64_t> n2e_off(nn + 1, 0);
  std::vector<int32_t> n2e(4 * (size_t)ne);
  {
    int const Tc = (int)std::max<int64_t>(1, std::min<int64_t>(omp_get_max_threads(), ((int64_t)64 << 20) / std::max(nn, 1)));
    std::unique_ptr<uint32_t[]> hist(new uint32_t[(size_t)Tc * nn]);
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < Tc; ++t) {
      uint32_t* h = hist.get() + (size_t)t * nn;
      memset(h, 0, sizeof(uint32_t) * (size_t)nn);
      int64_t const i0 = 4 * ((int64_t)ne * t / Tc), i1 = 4 * ((int64_t)ne * (t + 1) / Tc);
      for (int64_t i = i0; i < i1; ++i) h[conn[i]]++;
    }
#pragma omp parallel for schedule(static)
    for (int n = 0; n < nn; ++n) {
      int64_t tot = 0;
      for (int t = 0; t < Tc; ++t) tot += hist[(size_t)t * nn + n];
      n2e_off[n + 1] = tot;
    }
    for (int n = 0; n < nn; ++n) n2e_off[n + 1] += n2e_off[n];
#pragma omp parallel for schedule(static)
    for (int n = 0; n < nn; ++n) {  // counts -> start positions (fits 32 bits: 4 ne < 2^31)
      uint32_t off = (uint32_t)n2e_off[n];
      for (int t = 0; t < Tc; ++t) { uint32_t const k = hist[(size_t)t * nn + n]; hist[(size_t)t * nn + n] = off; off += k; }
    }
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < Tc; ++t) {
      uint32_t* h = hist.get() + (size_t)t * nn;
      int64_t const i0 = 4 * ((int64_t)ne * t / Tc), i1 = 4 * ((int64_t)ne * (t + 1) / Tc);
      for (int64_t i = i0; i < i1; ++i) n2e[h[conn[i]]++] = (int32_t)(i >> 2);
    }
  }

  tm.lap("node->elements");
  // ---- row-owner work list: incidences (e, n) of every node, elements ascending
  c->adj_off.resize(nn + 1);
  c->max_deg = 0;
  for (int n = 0; n <= nn; ++n) c->adj_off[n] = (uint32_t)n2e_off[n];
  c->has_isolated_nodes = false;
  for (int n = 0; n < nn; ++n) {
    c->max_deg = std::max<int>(c->max_deg, (int)(n2e_off[n + 1] - n2e_off[n]));
    if (n2e_off[n + 1] == n2e_off[n]) c->has_isolated_nodes = true;
  }

  // ---- node adjacency, sorted unique per row (two passes: count, fill).  Duplicates are recognised with a per-thread
  //      stamp array (stamp[b] == a: b already seen for row a), so only the unique neighbours are ever sorted.
  c->nrow.assign(nn + 1, 0);
  int bad_row = 0;
#pragma omp parallel
  {
    std::vector<int32_t> stamp(nn, -1);
#pragma omp for schedule(dynamic, 4096)
    for (int n = 0; n < nn; ++n) {
      int64_t cnt = 0;
      for (int64_t k = n2e_off[n]; k < n2e_off[n + 1]; ++k) {
        int32_t const* en = conn + 4 * (int64_t)n2e[k];
        for (int q = 0; q < 4; ++q) if (stamp[en[q]] != n) { stamp[en[q]] = n; ++cnt; }
      }
      if (cnt > 255) bad_row = 1;
      c->nrow[n + 1] = cnt;
    }
  }
  if (bad_row) { c->err = "a node has more than 255 neighbours (scatter map is 8-bit)"; return GX_ERR_UNSUPPORTED; }
  for (int n = 0; n < nn; ++n) c->nrow[n + 1] += c->nrow[n];
  if (c->nrow[nn] > 0x7fffffffLL) { c->err = "more than 2^31 node blocks"; return GX_ERR_UNSUPPORTED; }
  c->ncol.resize(c->nrow[nn]);
  c->nnz = 16 * c->nrow[nn];
#pragma omp parallel
  {
    std::vector<int32_t> stamp(nn, -1);
#pragma omp for schedule(dynamic, 4096)
    for (int n = 0; n < nn; ++n) {
      int32_t* dst = c->ncol.data() + c->nrow[n];
      int cnt = 0;
      for (int64_t k = n2e_off[n]; k < n2e_off[n + 1]; ++k) {
        int32_t const* en = conn + 4 * (int64_t)n2e[k];
        for (int q = 0; q < 4; ++q) if (stamp[en[q]] != n) { stamp[en[q]] = n; dst[cnt++] = en[q]; }
      }
      std::sort(dst, dst + cnt);
    }
  }
  tm.lap("node adjacency");
  // ---- scatter map: position of block (a_n, a_m) in a_n's block row, and the node -> (element, local node) incidences.
  //      One pass over the nodes: pos[b] = position of b in the current row (every node of an element incident to a is
  //      in a's row), so a block position is an array read instead of a binary search.  The four bytes of (e, n) are
  //      written by the thread that owns node a_n: distinct bytes, no two threads write the same one.
  c->bpos.resize(16 * (size_t)ne);
  c->adj.resize(n2e.size());
  c->max_nblk = 0;
  for (int n = 0; n < nn; ++n) c->max_nblk = std::max<int>(c->max_nblk, (int)(c->nrow[n + 1] - c->nrow[n]));
#pragma omp parallel
  {
    std::vector<uint8_t> pos(nn, 0);
#pragma omp for schedule(dynamic, 4096)
    for (int a = 0; a < nn; ++a) {
      int32_t const* b = c->ncol.data() + c->nrow[a];
      int const nb = (int)(c->nrow[a + 1] - c->nrow[a]);
      for (int j = 0; j < nb; ++j) pos[b[j]] = (uint8_t)j;
      for (int64_t k = n2e_off[a]; k < n2e_off[a + 1]; ++k) {
        int const e = n2e[k];
        int32_t const* en = conn + 4 * (int64_t)e;
        int n = 0;
        while (en[n] != a) ++n;
        uint8_t* d = &c->bpos[16 * (size_t)e + 4 * n];
        for (int m = 0; m < 4; ++m) d[m] = pos[en[m]];
        c->adj[k].x = e * 4 + n;
        c->adj[k].y = (int)((uint32_t)d[0] | ((uint32_t)d[1] << 8) | ((uint32_t)d[2] << 16) | ((uint32_t)d[3] << 24));
      }
    }
  }

  tm.lap("scatter map + incidences");
  // ---- diagonal block position per node (Dirichlet rows put their 1 there)
  c->diag_pos.assign(nn, 0);
#pragma omp parallel for schedule(static)
  for (int a = 0; a < nn; ++a) {
    int32_t const* b = c->ncol.data() + c->nrow[a];
    int32_t const* e = c->ncol.data() + c->nrow[a + 1];
    c->diag_pos[a] = (uint8_t)(std::lower_bound(b, e, a) - b);
  }

  // ---- stage B node order: Z-curve over the bounding box (10 bits per axis), ties by node id
  {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int a = 0; a < nn; ++a)
      for (int d = 0; d < 3; ++d) {
        lo[d] = std::min(lo[d], c->coords[3 * (size_t)a + d]);
        hi[d] = std::max(hi[d], c->coords[3 * (size_t)a + d]);
      }
    auto spread = [](uint64_t v) {  // 10 bits -> every third bit
      v &= 0x3ff;
      v = (v | (v << 16)) & 0x30000ff;
      v = (v | (v << 8)) & 0x300f00f;
      v = (v | (v << 4)) & 0x30c30c3;
      v = (v | (v << 2)) & 0x9249249;
      return v;
    };
    std::vector<uint64_t> key(nn);
#pragma omp parallel for schedule(static)
    for (int a = 0; a < nn; ++a) {
      uint64_t code = 0;
      for (int d = 0; d < 3; ++d) {
        double const ext = hi[d] - lo[d];
        uint64_t const q = ext > 0 ? (uint64_t)std::min(1023.0, (c->coords[3 * (size_t)a + d] - lo[d]) / ext * 1024.0) : 0;
        code |= spread(q) << d;
      }
      key[a] = (code << 32) | (uint32_t)a;
    }
    __gnu_parallel::sort(key.begin(), key.end());  // keys are unique (node id in the low word): the order is the serial one
    c->node_order.resize(nn);
    for (int a = 0; a < nn; ++a) c->node_order[a] = (int32_t)(key[a] & 0xffffffffu);
  }

  tm.lap("diag + Z-curve order");
  return GX_OK;
}

// The coloured element schedule (fallback / cross-check path only): built on first use.
int build_colouring(gx_ctx* c) {
  if (c->ncolors > 0) return GX_OK;
  SetupTimer tm;
  int const nn = c->nn, ne = c->ne;
  int32_t const* conn = c->conn.data();
  // ---- greedy colouring over node conflicts (elements sharing a node get different colours)
  constexpr int W = 4;  // 256 colours at most
  std::vector<uint64_t> used((size_t)nn * W, 0);
  std::vector<uint8_t> color(ne);
  int ncolors = 0;
  for (int e = 0; e < ne; ++e) {
    int32_t const* en = conn + 4 * (int64_t)e;
    int col = -1;
    for (int w = 0; w < W && col < 0; ++w) {
      uint64_t m = used[(size_t)en[0] * W + w] | used[(size_t)en[1] * W + w] | used[(size_t)en[2] * W + w] |
                   used[(size_t)en[3] * W + w];
      if (~m) col = 64 * w + __builtin_ctzll(~m);
    }
    if (col < 0) { c->err = "element colouring needs more than 256 colours"; return GX_ERR_UNSUPPORTED; }
    color[e] = (uint8_t)col;
    for (int a = 0; a < 4; ++a) used[(size_t)en[a] * W + col / 64] |= 1ull << (col % 64);
    ncolors = std::max(ncolors, col + 1);
  }
  c->ncolors = ncolors;
  c->color_off.assign(ncolors + 1, 0);
  for (int e = 0; e < ne; ++e) c->color_off[color[e] + 1]++;
  for (int k = 0; k < ncolors; ++k) c->color_off[k + 1] += c->color_off[k];
  c->perm.resize(ne);
  {
    std::vector<int32_t> cur(c->color_off.begin(), c->color_off.end() - 1);
    for (int e = 0; e < ne; ++e) c->perm[cur[color[e]]++] = e;  // stable: natural order inside a colour
  }
  tm.lap("colouring");
  return GX_OK;
}

// (The per-block contribution lists the patch schedule needs -- for block (a,b) the (element, n, m) with node n = a, node
// m = b -- are built node by node inside build_patch_schedule from the node's incidences; a global list would be 16
// entries per element.)
void build_block_lists(gx_ctx* c) { c->block_lists_built = true; }

// Patch schedule of the Jacobian pass (stage B, patch_pair_kernel).  A patch is a run of nodes of the Z-curve visiting
// order whose incident elements (at most PATCH_RECS) are staged once in shared memory by one thread block.  The
// patch's work is cut into items, one per thread:
//   PAIR  an edge (a,b) of the mesh, owned by the patch of whichever end comes first in the visiting order (all
//         elements around the edge are incident to that end, hence staged): the two mirror blocks (a,b) and (b,a),
//         built from the same staged data of every element around the edge;
//   DIAG  the diagonal block (a,a) and the four residual entries of node a;
//   ZERO  a phantom block of a partitioned context (a column that lives on another part): written as zeros.
// An item holds at most PATCH_ITEM_LEN contributions; longer lists (the diagonal: one per incident element, edges
// of high valence) are cut into a primary item and secondaries whose partial sums the primary adds in a fixed
// order.  Layout per patch (uint32 words, PATCH_WORDS):
//   [0..3]   n_recs, lanes in use (idle lanes in between have kind 0), n_runs, 1 if some unit (primary + secondaries) straddles a warp (block barrier needed)
//   then     items[PATCH_THREADS][4]   8 rounds x 16 bit: slot | m << 8 | n << 10 | 0x8000; 0 = sits the round out
//   then     outs[PATCH_THREADS][4]    w0 = first block of row a (extended block rows), w1 = first block of row b (PAIR)
//                                      or the node id a (DIAG); w2 = j1 | nblk1 << 8 | j2 << 16 | nblk2 << 24 (position
//                                      of the block in its row, blocks per row); w3 = kind | type << 2 | part << 4 | nsec << 12
//            kind: 0 idle thread, 1 primary, 2 secondary;  type: 0 ZERO, 1 DIAG, 2 PAIR
//   then     runs[PATCH_RECS][2]       bulk copies: first element id, first slot | number of records << 8
bool build_patch_schedule(gx_ctx* c) {
  using namespace gx;
  SetupTimer tm;
  if (!c->block_lists_built) build_block_lists(c);
  tm.lap("block lists");
  int const nn = c->nn;
  std::vector<int64_t> const& nx = c->nrow_x;
  int const CH = 4096;  // nodes per independent chunk of the visiting order
  int const nch = (nn + CH - 1) / CH + 1;  // upper bound (the two node classes are chunked separately)
  std::vector<std::vector<uint32_t>> out(nch);
  bool const stats = getenv("GX_SCHED_STATS") != nullptr;
  bool const nomatch = getenv("GX_SCHED_NOMATCH") != nullptr;
  std::vector<int64_t> st_wave(nch, 0), st_rounds(nch, 0), st_runs(nch, 0), st_recs(nch, 0), st_items(nch, 0), st_contrib(nch, 0);
  bool ok = true;
  // Long contribution lists are cut into items of at most `split` contributions, as evenly as possible.  A thread
  // block lives as long as its longest item, so the cut follows the length of the ordinary items (an edge of a Kuhn
  // mesh has 4 or 6 elements) rather than the capacity of an item.
  int split = 6, split_diag = 6, max_run = 32;  // records per bulk copy at most
  if (char const* e = getenv("GX_SCHED_MAXRUN")) max_run = std::max(1, std::min(64, atoi(e)));
  if (char const* e = getenv("GX_SCHED_SPLIT")) split = std::max(1, std::min(PATCH_ITEM_LEN, atoi(e)));
  if (char const* e = getenv("GX_SCHED_SPLIT_DIAG")) split_diag = std::max(1, std::min(PATCH_ITEM_LEN, atoi(e)));
  auto n_parts = [&](int cnt, bool diag) { int const sp = diag ? split_diag : split; return std::max(1, (cnt + sp - 1) / sp); };
  // Visiting order: the Z-curve order of the nodes; on a partitioned context the nodes on part boundaries and their
  // neighbours come first (each class in Z-curve order).  Every block of an interface node's rows -- its diagonal, and
  // its pairs, which belong to the earlier end of an edge whose two ends are both in that class -- is then written by
  // the first n_patches_iface patches, so the interface exchange can start while the interior patches still run.
  std::vector<int32_t> order(c->node_order);
  int n_first = 0;
  if (!c->peers.empty()) {
    std::vector<char> mark(nn, 0);
    for (auto const& P : c->peers) for (int a : P.nodes) mark[a] = 1;
    std::vector<char> near(mark);
    for (int a = 0; a < nn; ++a)
      if (mark[a]) for (int64_t k = c->nrow[a]; k < c->nrow[a + 1]; ++k) near[c->ncol[k]] = 1;
    std::stable_partition(order.begin(), order.end(), [&](int32_t a) { return near[a] != 0; });
    for (int a = 0; a < nn; ++a) n_first += near[a] != 0;
  }
  std::vector<int32_t> rank(nn);
  for (int s2 = 0; s2 < nn; ++s2) rank[order[s2]] = s2;
  // chunks of the visiting order, built independently; no chunk straddles the two classes
  std::vector<std::pair<int, int>> chunk_rng;
  for (int b0 = 0; b0 < n_first; b0 += CH) chunk_rng.emplace_back(b0, std::min(n_first, b0 + CH));
  int const n_chunks_first = (int)chunk_rng.size();
  for (int b0 = n_first; b0 < nn; b0 += CH) chunk_rng.emplace_back(b0, std::min(nn, b0 + CH));
#pragma omp parallel for schedule(dynamic, 1)
  for (int ch = 0; ch < (int)chunk_rng.size(); ++ch) {
    struct Item { uint16_t ent[PATCH_ITEM_LEN]; int n; uint32_t w0, w1, w2; int kind, type, nsec, part; };
    std::vector<Item> items;
    std::vector<int32_t> recs;
    int32_t hkey[1024]; int16_t hval[1024];
    auto hclear = [&]() { for (int i = 0; i < 1024; ++i) hkey[i] = -1; };
    auto hfind = [&](int32_t e) -> int {
      uint32_t h = ((uint32_t)e * 2654435761u) >> 22;
      while (hkey[h] != -1) { if (hkey[h] == e) return hval[h]; h = (h + 1) & 1023u; }
      return -1;
    };
    auto hput = [&](int32_t e, int v) {
      uint32_t h = ((uint32_t)e * 2654435761u) >> 22;
      while (hkey[h] != -1) h = (h + 1) & 1023u;
      hkey[h] = e; hval[h] = (int16_t)v;
    };
    int nparts = 0, cur_d = 0, cur_p = 0;  // secondaries, DIAG lanes and PAIR lanes of the patch being filled
    // scratch of flush(), allocated once per chunk
    struct UnitS { int prim, nsec, len; };
    std::vector<int> ord, zeros, res, slot_of, byel, feed_off, feed_q, qcap;
    std::vector<uint32_t> run_e0, run_sl;
    std::vector<std::array<int, 8>> qload;
    std::vector<UnitS> groups_s, singles_s;
    auto flush = [&]() {
      if (items.empty()) return;
      // Lane order: typed warps.  The first PATCH_DIAG_LANES lanes (whole warps) hold the DIAG items, the rest the
      // PAIR items, so that no warp runs both code paths (a warp that did would take twice as long and hold the
      // block's shared memory while the others idle); ZERO items take whatever lanes stay free.  A "unit" is a
      // primary item with its secondaries; a unit that sits inside one warp hands its partial sums over with a warp
      // barrier only.  Inside a region: units with secondaries first, by size (equal power-of-two sizes are aligned by
      // construction; a unit that would straddle a warp boundary is preceded by the shortest single items of the
      // region as filler), then the single items, longest first: the lanes of a warp then run the same number of
      // contributions.  ord[t] = item of lane t, -1 = idle lane.
      ord.clear();
      bool block_sync = false;
      {
        using Unit = UnitS;  // the secondaries of item i are items i+1 .. i+nsec (pushed right behind it)
        std::vector<Unit>& groups = groups_s; std::vector<Unit>& singles = singles_s;
        auto layout = [&](int type, int first_lane) {
          groups.clear(); singles.clear();
          for (size_t i = 0; i < items.size(); ++i) {
            if (items[i].kind != 1 || items[i].type != type) continue;
            Unit const u{(int)i, items[i].nsec, items[i].n};
            (u.nsec == 0 ? singles : groups).push_back(u);
          }
          std::stable_sort(groups.begin(), groups.end(), [](Unit const& x, Unit const& y) {
            if (x.nsec != y.nsec) return x.nsec > y.nsec;
            return x.len > y.len;
          });
          std::stable_sort(singles.begin(), singles.end(), [](Unit const& x, Unit const& y) { return x.len > y.len; });
          ord.resize(first_lane, -1);
          size_t pool_end = singles.size();
          auto fill = [&](int lanes) {  // the shortest singles, from the back; idle lanes when they run out
            while (lanes > 0 && pool_end > 0) { ord.push_back(singles[--pool_end].prim); --lanes; }
            while (lanes-- > 0) ord.push_back(-1);
          };
          for (Unit const& u : groups) {
            int const sz = 1 + u.nsec;
            int const room = 32 - (int)(ord.size() % 32);
            if (sz > 32) block_sync = true;
            else if (sz > room) fill(room);
            for (int q = 0; q < sz; ++q) ord.push_back(u.prim + q);
          }
          for (size_t i = 0; i < pool_end; ++i) ord.push_back(singles[i].prim);
        };
        layout(1, 0);
        int const diag_end = (int)ord.size();
        layout(2, std::max(diag_end, PATCH_DIAG_LANES));
        zeros.clear();
        for (size_t i = 0; i < items.size(); ++i) if (items[i].type == 0) zeros.push_back((int)i);
        // zeros into the idle lanes, then behind
        size_t zi = 0;
        for (size_t t = 0; t < ord.size() && zi < zeros.size(); ++t) if (ord[t] < 0) ord[t] = zeros[zi++];
        while (zi < zeros.size()) ord.push_back(zeros[zi++]);
        if ((int)ord.size() > PATCH_THREADS) {
          // Alignment padding does not fit: drop it (lanes packed in region order) and let the block barrier
          // hand the partial sums over.  Rare: needs units that straddle warps in an almost full patch.
          std::vector<int> packed;
          for (int v : ord) if (v >= 0) packed.push_back(v);
          ord.swap(packed);
          block_sync = true;
        }
      }
      int const n_real_items = (int)items.size();
      items.push_back(Item{});  // the idle lane: kind 0, no contributions
      for (int& v : ord) if (v < 0) v = n_real_items;
      // Shared-memory bank conflicts: a 128-bit load is served per quarter-warp (8 lanes), and the bank group of a
      // staged record is a permutation of its slot modulo 8 (record stride 21 x 16 B, odd).  The lanes of a quarter
      // read one record each per round; they collide only when two of them read different records of one bank group
      // in the same round.  So (1) the records get their bank groups such that, for every quarter-warp, the
      // contributions of its eight items are spread evenly over the eight groups (no group more than the warp's
      // rounds, if possible), and (2) every quarter-warp orders its contributions by an edge colouring (below).
      int const nrec = (int)recs.size();
      res.assign(nrec, -1);
      slot_of.assign(nrec, -1);
      run_e0.clear(); run_sl.clear();  // runs of consecutive elements in consecutive slots: one bulk copy each
      int const nquart = ((int)ord.size() + 7) / 8;
      // quarter-warps a record feeds, with multiplicity (CRS: feed_off / feed_q)
      feed_off.assign(nrec + 1, 0);
      for (size_t t = 0; t < ord.size(); ++t) {
        Item const& it = items[ord[t]];
        for (int q = 0; q < it.n; ++q) feed_off[(it.ent[q] & 0xff) + 1]++;
      }
      for (int l = 0; l < nrec; ++l) feed_off[l + 1] += feed_off[l];
      feed_q.resize(feed_off[nrec]);
      {
        int cur[PATCH_RECS];
        for (int l = 0; l < nrec; ++l) cur[l] = feed_off[l];
        for (size_t t = 0; t < ord.size(); ++t) {
          Item const& it = items[ord[t]];
          for (int q = 0; q < it.n; ++q) feed_q[cur[it.ent[q] & 0xff]++] = (int)(t / 8);
        }
      }
      qload.assign(nquart, std::array<int, 8>{});  // contributions of quarter Q that read bank group r
      qcap.assign(nquart, 0);                       // rounds of the quarter's warp
      for (int Q = 0; Q < nquart; ++Q) {
        int R = 0;
        for (size_t t = ((size_t)Q / 4) * 32; t < std::min(ord.size(), ((size_t)Q / 4) * 32 + 32); ++t) R = std::max(R, items[ord[t]].n);
        qcap[Q] = R;
      }
      // cost of giving record l the bank group r (computed below for all eight r at once): readers already there, and
      // heavily (+16 each), readers beyond the rounds of their warp
      auto place = [&](int l, int r) {
        res[l] = r;
        for (int k = feed_off[l]; k < feed_off[l + 1]; ++k) qload[feed_q[k]][r]++;
      };
      {
        // Records of consecutive elements are consecutive in global memory: placed in consecutive slots they arrive with
        // ONE bulk copy.  Longest runs first, each at the free position with the fewest conflicts (ties: lowest slot).
        byel.resize(nrec);
        for (int l = 0; l < nrec; ++l) byel[l] = l;
        std::sort(byel.begin(), byel.end(), [&](int x, int y) { return recs[x] < recs[y]; });
        struct Run { int first, len; };
        std::vector<Run> runs;
        for (int i = 0; i < nrec;) {
          int j = i + 1;
          while (j < nrec && recs[byel[j]] == recs[byel[j - 1]] + 1 && j - i < max_run) ++j;
          runs.push_back({i, j - i});
          i = j;
        }
        std::stable_sort(runs.begin(), runs.end(), [](Run const& a, Run const& b) { return a.len > b.len; });
        bool taken[PATCH_RECS] = {};
        int free_len[PATCH_RECS + 1];  // free slots in a row starting at s
        for (int s0 = 0; s0 <= PATCH_RECS; ++s0) free_len[s0] = PATCH_RECS - s0;
        for (size_t ri = 0; ri < runs.size(); ++ri) {
          Run const run = runs[ri];
          // the cost of a position depends on its residue modulo 8 only: rate the eight residues, then take the
          // first free position of the best residue that has one
          int rcost[8] = {};
          for (int j = 0; j < run.len; ++j) {
            int const l = byel[run.first + j];
            int cl[8] = {};  // conflicts(l, r) for the eight bank groups at once
            for (int k = feed_off[l]; k < feed_off[l + 1]; ++k) {
              int const Q = feed_q[k], cap = qcap[Q];
              for (int r = 0; r < 8; ++r) cl[r] += qload[Q][r] + (qload[Q][r] >= cap ? 16 : 0);
            }
            for (int r0 = 0; r0 < 8; ++r0) rcost[r0] += cl[(r0 + j) & 7];
          }
          // = the feasible position with the smallest (cost of its residue, slot): residues in order of cost, and among
          // residues of equal cost the one whose first free position comes first
          int best = -1;
          {
            int byc[8] = {0, 1, 2, 3, 4, 5, 6, 7};
            std::stable_sort(byc, byc + 8, [&](int x, int y) { return rcost[x] < rcost[y]; });
            for (int g = 0; g < 8 && best < 0;) {
              int h = g;
              while (h < 8 && rcost[byc[h]] == rcost[byc[g]]) ++h;
              for (int q = g; q < h; ++q)
                for (int s0 = byc[q]; s0 + run.len <= PATCH_RECS && (best < 0 || s0 < best); s0 += 8)
                  if (free_len[s0] >= run.len) { best = s0; break; }
              g = h;
            }
          }
          if (best < 0) {  // fragmented: cut the run in two and place the halves
            runs.push_back({run.first, run.len / 2});
            runs.push_back({run.first + run.len / 2, run.len - run.len / 2});
            continue;
          }
          for (int j = 0; j < run.len; ++j) {
            int const l = byel[run.first + j];
            taken[best + j] = true; slot_of[l] = best + j; place(l, (best + j) & 7);
          }
          for (int s0 = best; s0 < best + run.len; ++s0) free_len[s0] = 0;
          for (int s0 = best - 1; s0 >= 0 && !taken[s0]; --s0) free_len[s0] = best - s0;  // the free stretch in front now ends at best
          run_e0.push_back((uint32_t)recs[byel[run.first]]);
          run_sl.push_back((uint32_t)best | ((uint32_t)run.len << 8));
        }
      }
      // Round assignment per quarter-warp (8 lanes): a proper edge colouring of the bipartite multigraph
      // items x bank groups (one edge per contribution, colour = round).  Koenig: with R rounds available (the
      // warp's longest item) a conflict-free assignment exists iff no bank group carries more than R of the eight
      // items' contributions; the colouring below finds it (alternating-path recolouring).  Where a group is
      // over-subscribed the surplus edges go to the round with the fewest readers of that group.  An item with
      // fewer contributions than rounds sits the other rounds out (empty entries).
      for (size_t g0 = 0; g0 < ord.size() && !nomatch; g0 += 8) {
        int const gn = (int)std::min<size_t>(8, ord.size() - g0);
        int R = 0;
        for (size_t t = (g0 / 32) * 32; t < std::min(ord.size(), (g0 / 32) * 32 + 32); ++t) R = std::max(R, items[ord[t]].n);
        if (R == 0) continue;
        struct Edge { int u, v, q, col; };
        Edge edges[8 * PATCH_ITEM_LEN];
        int n_edges = 0;
        int degv[8] = {};
        for (int i = 0; i < gn; ++i) {
          Item const& it = items[ord[g0 + i]];
          for (int q = 0; q < it.n; ++q) { int const v = res[it.ent[q] & 0xff]; edges[n_edges++] = {i, v, q, -1}; degv[v]++; }
        }
        int C = R;
        for (int v = 0; v < 8; ++v) C = std::max(C, degv[v]);  // <= 8 * PATCH_ITEM_LEN
        int atu[8 * 8 * PATCH_ITEM_LEN], atv[8 * 8 * PATCH_ITEM_LEN];  // edge with colour c at item u / at bank group v
        for (int k = 0; k < 8 * C; ++k) { atu[k] = -1; atv[k] = -1; }
        uint64_t const all = C >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << C) - 1);
        uint64_t freeu[8], freev[8];  // bit c set = colour c unused at the item / at the bank group (mirrors atu / atv)
        for (int k = 0; k < 8; ++k) { freeu[k] = all; freev[k] = all; }
        auto put = [&](int ei, int col) {
          Edge& pe = edges[ei];
          pe.col = col;
          atu[pe.u * C + col] = ei; atv[pe.v * C + col] = ei;
          freeu[pe.u] &= ~((uint64_t)1 << col); freev[pe.v] &= ~((uint64_t)1 << col);
        };
        auto take = [&](int ei) {
          Edge const& pe = edges[ei];
          atu[pe.u * C + pe.col] = -1; atv[pe.v * C + pe.col] = -1;
          freeu[pe.u] |= (uint64_t)1 << pe.col; freev[pe.v] |= (uint64_t)1 << pe.col;
        };
        for (int ei = 0; ei < n_edges; ++ei) {
          Edge& e = edges[ei];
          if (!freeu[e.u] || !freev[e.v]) { e.col = -2; continue; }  // cannot happen (C >= every degree); placed below if it does
          int const ca = __builtin_ctzll(freeu[e.u]), cb = __builtin_ctzll(freev[e.v]);  // lowest free colour at either end
          if (atv[e.v * C + ca] >= 0) {
            // colour ca is taken at v: swap ca <-> cb along the alternating path that starts at v with colour ca
            int path[8 * PATCH_ITEM_LEN], n_path = 0;
            int cur = atv[e.v * C + ca];
            bool at_v = true;  // the path edge was reached through its v end
            int want = ca;
            while (cur >= 0) {
              path[n_path++] = cur;
              Edge const& pe = edges[cur];
              want = want == ca ? cb : ca;
              cur = at_v ? atu[pe.u * C + want] : atv[pe.v * C + want];
              at_v = !at_v;
            }
            int newcol[8 * PATCH_ITEM_LEN];
            for (int k = 0; k < n_path; ++k) { newcol[k] = edges[path[k]].col == ca ? cb : ca; take(path[k]); }
            for (int k = 0; k < n_path; ++k) put(path[k], newcol[k]);
          }
          put(ei, ca);
        }
        uint16_t sched_ent[8][PATCH_ITEM_LEN] = {};
        bool used_round[8][PATCH_ITEM_LEN] = {};
        int readers[PATCH_ITEM_LEN][8] = {};  // readers of bank group v in round k (distinct records not tracked: a bound)
        for (int ei = 0; ei < n_edges; ++ei) {
          Edge const& e = edges[ei];
          if (e.col >= 0 && e.col < R) {
            sched_ent[e.u][e.col] = items[ord[g0 + e.u]].ent[e.q]; used_round[e.u][e.col] = true; readers[e.col][e.v]++;
          }
        }
        for (int ei = 0; ei < n_edges; ++ei) {
          Edge const& e = edges[ei];
          if (e.col < 0 || e.col >= R) {  // surplus: the free round of this item where the group has the fewest readers
            int best = -1;
            for (int k = 0; k < R; ++k)
              if (!used_round[e.u][k] && (best < 0 || readers[k][e.v] < readers[best][e.v])) best = k;
            sched_ent[e.u][best] = items[ord[g0 + e.u]].ent[e.q]; used_round[e.u][best] = true; readers[best][e.v]++;
          }
        }
        for (int i = 0; i < gn; ++i) {
          Item& it = items[ord[g0 + i]];
          if (it.n == 0) continue;
          for (int k = 0; k < PATCH_ITEM_LEN; ++k) it.ent[k] = sched_ent[i][k];
        }
      }
      for (auto& it : items)  // provisional record numbers -> final slots
        for (int k = 0; k < PATCH_ITEM_LEN; ++k)
          if (it.ent[k] & 0x8000) it.ent[k] = (uint16_t)((it.ent[k] & 0xff00) | slot_of[it.ent[k] & 0xff]);
      if (stats) {  // wavefronts per 128-bit load and round: the fullest bank group (distinct records)
        st_runs[ch] += (int64_t)run_e0.size();
        st_recs[ch] += (int64_t)nrec;
        st_items[ch] += (int64_t)n_real_items;
        for (auto const& it : items) st_contrib[ch] += it.n;
        for (size_t g0 = 0; g0 < ord.size(); g0 += 8) {
          int const gn = (int)std::min<size_t>(8, ord.size() - g0);
          for (int k = 0; k < PATCH_ITEM_LEN; ++k) {
            int cnt[8] = {};
            bool any = false;
            for (int i = 0; i < gn; ++i) {
              Item const& it = items[ord[g0 + i]];
              if (!(it.ent[k] & 0x8000)) continue;
              any = true;
              bool dup = false;
              for (int j = 0; j < i; ++j) {
                Item const& jt = items[ord[g0 + j]];
                if ((jt.ent[k] & 0x8000) && (jt.ent[k] & 0xff) == (it.ent[k] & 0xff)) dup = true;
              }
              if (!dup) cnt[it.ent[k] & 7]++;
            }
            if (!any) continue;
            int mx = 0;
            for (int r = 0; r < 8; ++r) mx = std::max(mx, cnt[r]);
            st_wave[ch] += mx; st_rounds[ch] += 1;
          }
        }
      }
      size_t const base = out[ch].size();
      out[ch].resize(base + PATCH_WORDS, 0u);
      uint32_t* w = out[ch].data() + base;
      w[0] = (uint32_t)nrec; w[1] = (uint32_t)ord.size(); w[2] = (uint32_t)run_e0.size(); w[3] = block_sync ? 1u : 0u;
      uint32_t* wi = w + 4;
      uint32_t* wo = wi + 4 * PATCH_THREADS;
      uint32_t* wr = wo + 4 * PATCH_THREADS;  // runs[PATCH_RECS][2]: first element, first slot | length << 8
      for (size_t i = 0; i < run_e0.size(); ++i) { wr[2 * i] = run_e0[i]; wr[2 * i + 1] = run_sl[i]; }
      for (size_t t = 0; t < ord.size(); ++t) {
        Item const& it = items[ord[t]];
        for (int k = 0; k < 4; ++k) wi[4 * t + k] = (uint32_t)it.ent[2 * k] | ((uint32_t)it.ent[2 * k + 1] << 16);
        wo[4 * t] = it.w0; wo[4 * t + 1] = it.w1; wo[4 * t + 2] = it.w2;
        wo[4 * t + 3] = (uint32_t)it.kind | ((uint32_t)it.type << 2) | ((uint32_t)it.part << 4) | ((uint32_t)it.nsec << 12);
      }
      items.clear(); recs.clear(); hclear(); nparts = 0; cur_d = 0; cur_p = 0;
    };
    hclear();
    bool bad = false;
    int const s1 = chunk_rng[ch].second;
    // the items of node a: its diagonal block, the edges it owns, its phantom blocks
    struct Blk { int64_t t; int type; int cnt; int parts; };
    std::vector<Blk> blks;
    std::vector<int> lc_off, lc_cur;
    std::vector<int32_t> lc;
    for (int s = chunk_rng[ch].first; s < s1; ++s) {
      int const a = order[s];
      if (nx[a + 1] == nx[a]) continue;  // a node without elements: no blocks, no work (its R entries are zeroed by the pass)
      int add = 0;  // new records this node would add
      for (uint32_t k = c->adj_off[a]; k < c->adj_off[a + 1]; ++k) if (hfind(c->adj[k].x >> 2) < 0) ++add;
      blks.clear();
      int nit = 0, nsecs = 0, nd = 0, np = 0;  // items, secondaries, DIAG lanes, PAIR lanes this node needs
      int64_t const nloc = c->nrow[a + 1] - c->nrow[a];  // local blocks; the extended row may continue with phantom ones
      int const nb_x = (int)(nx[a + 1] - nx[a]);
      // contributions of this node's blocks, grouped by block position, elements ascending (the incidences are):
      // lc_off[j] .. lc_off[j+1] index lc[], an entry is e*16 + n*4 + m
      lc_off.assign(nb_x + 1, 0);
      for (uint32_t k = c->adj_off[a]; k < c->adj_off[a + 1]; ++k) {
        uint32_t const jp = (uint32_t)c->adj[k].y;
        for (int m = 0; m < 4; ++m) lc_off[((jp >> (8 * m)) & 0xffu) + 1]++;
      }
      for (int j = 0; j < nb_x; ++j) lc_off[j + 1] += lc_off[j];
      lc.resize(lc_off[nb_x]);
      lc_cur.assign(lc_off.begin(), lc_off.end() - 1);
      for (uint32_t k = c->adj_off[a]; k < c->adj_off[a + 1]; ++k) {
        int const en = c->adj[k].x;  // e*4 + n
        uint32_t const jp = (uint32_t)c->adj[k].y;
        for (int m = 0; m < 4; ++m) lc[lc_cur[(jp >> (8 * m)) & 0xffu]++] = (int32_t)(4 * en + m);
      }
      int const jdiag = c->diag_pos[a];
      for (int j = 0; j < nb_x; ++j) {
        int64_t const t = nx[a] + j;
        int const cnt = lc_off[j + 1] - lc_off[j];
        int type;
        if (j == jdiag) type = 1;
        else if (cnt == 0) type = 0;
        else {
          int const b = c->ncol[c->nrow[a] + j];
          (void)nloc;
          if (rank[b] < s) continue;  // the edge belongs to the patch of b
          type = 2;
        }
        int const parts = n_parts(cnt, type == 1);
        blks.push_back({t, type, cnt, parts});
        nit += parts; nsecs += parts - 1;
        if (type == 1) nd += parts;
        if (type == 2) np += parts;
      }
      // a node must fit an empty patch; the lane regions are filled with a little slack for alignment padding
      if (nd > PATCH_DIAG_LANES || np > PATCH_THREADS - PATCH_DIAG_LANES || nit > PATCH_THREADS ||
          (int)(c->adj_off[a + 1] - c->adj_off[a]) > PATCH_RECS || nsecs > PATCH_PARTS) { bad = true; break; }
      if (cur_d + nd > PATCH_DIAG_LANES || cur_p + np > PATCH_THREADS - PATCH_DIAG_LANES || (int)items.size() + nit > PATCH_THREADS ||
          (int)recs.size() + add > PATCH_RECS || nparts + nsecs > PATCH_PARTS) flush();
      cur_d += nd; cur_p += np;
      for (uint32_t k = c->adj_off[a]; k < c->adj_off[a + 1]; ++k) {
        int32_t const e = c->adj[k].x >> 2;
        if (hfind(e) < 0) { hput(e, (int)recs.size()); recs.push_back(e); }
      }
      uint32_t const nblk_a = (uint32_t)(nx[a + 1] - nx[a]);
      for (Blk const& bk : blks) {
        uint32_t const j1 = (uint32_t)(bk.t - nx[a]);
        int const c0 = lc_off[j1];
        uint32_t w1 = 0, j2 = 0, nblk_b = 0;
        if (bk.type == 1) w1 = (uint32_t)a;
        if (bk.type == 2) {
          int const b = c->ncol[c->nrow[a] + j1];
          int32_t const* rb = c->ncol.data() + c->nrow[b];
          j2 = (uint32_t)(std::lower_bound(rb, (int32_t const*)(c->ncol.data() + c->nrow[b + 1]), (int32_t)a) - rb);
          nblk_b = (uint32_t)(nx[b + 1] - nx[b]);
          w1 = (uint32_t)nx[b];
        }
        int first = 0;  // contributions in ascending element order, cut into `parts` runs of almost equal length
        for (int pi = 0; pi < bk.parts; ++pi) {
          Item it{};
          it.n = bk.cnt / bk.parts + (pi < bk.cnt % bk.parts ? 1 : 0);
          for (int q = 0; q < it.n; ++q) {
            int32_t const ent = lc[c0 + first + q];
            it.ent[q] = (uint16_t)(hfind(ent >> 4) | ((ent & 15) << 8) | 0x8000);  // n*4+m -> bits 8..11
          }
          it.w0 = (uint32_t)nx[a]; it.w1 = w1;
          it.w2 = j1 | (nblk_a << 8) | (j2 << 16) | (nblk_b << 24);
          it.type = bk.type;
          if (pi == 0) { it.kind = 1; it.nsec = bk.parts - 1; it.part = nparts; }
          else { it.kind = 2; it.nsec = 0; it.part = nparts + pi - 1; }
          first += it.n;
          items.push_back(it);
        }
        nparts += bk.parts - 1;
      }
    }
    flush();
    if (bad) {
#pragma omp atomic write
      ok = false;
    }
  }
  tm.lap("patches");
  if (!ok) { c->patch_state = -1; return false; }
  size_t total = 0;
  std::vector<size_t> out_sizes;
  for (auto& v : out) { total += v.size(); out_sizes.push_back(v.size()); }
  // the chunks stay as they are: the device upload copies them one by one (upload_patch_schedule), a flat host copy
  // is made only on request (flatten_patch_schedule: gx_patch_schedule, the CPU replay of tests/hostcheck)
  c->patch_sched.clear();
  c->patch_chunks.swap(out);
  c->n_patches = (int)(total / PATCH_WORDS);
  c->n_patches_iface = 0;
  for (int ch = 0; ch < n_chunks_first; ++ch) c->n_patches_iface += (int)(out_sizes[ch] / PATCH_WORDS);
  c->patch_state = 1;
  tm.lap("concatenate");
  if (stats) {
    int64_t wv = 0, rd = 0, runs = 0, nrecs = 0, nit = 0, nco = 0;
    for (int i = 0; i < nch; ++i) { wv += st_wave[i]; rd += st_rounds[i]; runs += st_runs[i]; nrecs += st_recs[i]; nit += st_items[i]; nco += st_contrib[i]; }
    double const np = std::max(1, c->n_patches);
    fprintf(stderr, "[gx] patch schedule: %d patches, %.2f nodes/patch, %.1f records/patch in %.1f runs (staging factor %.2f), %.1f items/patch, %.2f contributions/item, %.3f wavefronts per quarter-warp round\n",
            c->n_patches, (double)nn / np, (double)nrecs / np, (double)runs / np, (double)nrecs / std::max(1, c->ne), (double)nit / np,
            nit ? (double)nco / (double)nit : 0.0, rd ? (double)wv / (double)rd : 0.0);
  }
  return true;
}

void flatten_patch_schedule(gx_ctx* c) {
  if (!c->patch_sched.empty() || c->patch_chunks.empty()) return;
  std::vector<size_t> off(c->patch_chunks.size() + 1, 0);
  for (size_t i = 0; i < c->patch_chunks.size(); ++i) off[i + 1] = off[i] + c->patch_chunks[i].size();
  c->patch_sched.resize(off.back());
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < (int)c->patch_chunks.size(); ++i)
    if (!c->patch_chunks[i].empty()) memcpy(c->patch_sched.data() + off[i], c->patch_chunks[i].data(), sizeof(uint32_t) * c->patch_chunks[i].size());
  std::vector<std::vector<uint32_t>>().swap(c->patch_chunks);
}

// ---------------------------------------------------------------------------
// Block-reduced schedule of the residual / error-localisation passes (layout: gx_internal.h, RES_*).
// A thread block evaluates RES_BLOCK consecutive elements, leaves their 16 residual entries in shared memory and sums
// them per node there, in a fixed order; only one 32 B partial sum per (block, node shared with another block) goes
// through global memory instead of the 128 B per element of an element-by-element record (SolInfo ghost R,
// src/goal_sol_info.cpp:51-64 is what both produce).  Nodes whose elements all sit in one block are written by it.
// ---------------------------------------------------------------------------
bool build_residual_schedule(gx_ctx* c) {
  if (c->res_state != 0) return c->res_state == 1;
  SetupTimer tm;
  int const ne = c->ne, nn = c->nn;
  if (4 * (int64_t)ne >= ((int64_t)1 << 31)) { c->res_state = -1; return false; }  // partial positions are 31 bit
  int const nb = (ne + RES_BLOCK - 1) / RES_BLOCK;
  int const T = std::max(1, omp_get_max_threads());
  c->res_chunks.assign(T, std::vector<uint32_t>());
  c->res_boff.assign((size_t)nb + 1, 0);
  std::vector<uint32_t> npart(nn, 0);  // blocks that hold some, not all, elements of the node
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < T; ++t) {  // chunk t = a contiguous range of blocks, whatever thread runs it
    int const b0 = (int)((int64_t)nb * t / T), b1 = (int)((int64_t)nb * (t + 1) / T);
    std::vector<uint32_t>& out = c->res_chunks[t];
    std::vector<int32_t> stamp(nn, -1);
    std::vector<uint16_t> slot_of(nn, 0);
    uint32_t snode[4 * RES_BLOCK], scnt[4 * RES_BLOCK], sfirst[4 * RES_BLOCK];
    out.reserve((size_t)(b1 - b0) * 512);
    for (int b = b0; b < b1; ++b) {
      int const e0 = b * RES_BLOCK, cnt = std::min(RES_BLOCK, ne - e0);
      int S = 0;
      for (int i = 0; i < 4 * cnt; ++i) {
        int32_t const a = c->conn[4 * (size_t)e0 + i];
        if (stamp[a] != b) { stamp[a] = b; slot_of[a] = (uint16_t)S; snode[S] = (uint32_t)a; scnt[S] = 0; ++S; }
        ++scnt[slot_of[a]];
      }
      uint32_t run = 0;
      for (int s = 0; s < S; ++s) { sfirst[s] = run; run += scnt[s]; }
      size_t const base = out.size();
      uint32_t const nw = (uint32_t)((RES_HDR + 2 * S + (4 * cnt + 1) / 2 + 3) & ~3);
      out.resize(base + nw, 0u);
      uint32_t* w = out.data() + base;
      w[0] = (uint32_t)S; w[1] = nw;
      uint16_t* ent = reinterpret_cast<uint16_t*>(w + RES_HDR + 2 * S);
      for (int s = 0; s < S; ++s) {
        int32_t const a = (int32_t)snode[s];
        bool const complete = c->adj_off[a + 1] - c->adj_off[a] == scnt[s];
        w[RES_HDR + 2 * s] = snode[s] | (complete ? 0x80000000u : 0u);
        w[RES_HDR + 2 * s + 1] = sfirst[s] | (scnt[s] << 16);
        if (!complete) {
#pragma omp atomic
          ++npart[a];
        }
        scnt[s] = 0;  // reused as the fill cursor
      }
      for (int i = 0; i < 4 * cnt; ++i) {  // ascending (element, local node): ascending elements inside every slot
        int const s = slot_of[c->conn[4 * (size_t)e0 + i]];
        int const row = i >> 2, n = i & 3;
        ent[sfirst[s] + scnt[s]++] = (uint16_t)(8 * row + ((2 * n) ^ (row & 7)));  // 16 B chunk of the kernel's swizzled rows
      }
      c->res_boff[b + 1] = nw;  // block sizes; prefix sum below
    }
  }
  uint64_t total_words = 0;
  for (int b = 0; b < nb; ++b) { total_words += c->res_boff[b + 1]; c->res_boff[b + 1] = (uint32_t)total_words; }
  if (total_words >= ((uint64_t)1 << 32)) {  // block offsets are 32 bit
    std::vector<std::vector<uint32_t>>().swap(c->res_chunks);
    c->res_state = -1;
    return false;
  }
  tm.lap("residual blocks");
  // nodes finished by the second kernel: shared between blocks, or without any element (their R entries are zero)
  c->res_pnode.clear();
  for (int a = 0; a < nn; ++a)
    if (npart[a] > 0 || c->adj_off[a + 1] == c->adj_off[a]) c->res_pnode.push_back(a);
  c->res_poff.assign(c->res_pnode.size() + 1, 0);
  for (size_t i = 0; i < c->res_pnode.size(); ++i) {
    int32_t const a = c->res_pnode[i];
    c->res_poff[i + 1] = c->res_poff[i] + npart[a];
    npart[a] = c->res_poff[i];  // from here on: the node's cursor into the partial buffer
  }
  c->res_npartial = c->res_poff.back();
  // positions in block order: the second kernel adds a node's partial sums in ascending block order
  for (int t = 0; t < T; ++t) {
    std::vector<uint32_t>& ch = c->res_chunks[t];
    for (size_t o = 0; o < ch.size(); o += ch[o + 1]) {
      int const S = (int)ch[o];
      for (int s = 0; s < S; ++s) {
        uint32_t& w0 = ch[o + RES_HDR + 2 * s];
        if (!(w0 & 0x80000000u)) w0 = npart[w0]++;
      }
    }
  }
  tm.lap("partial positions");
  c->res_state = 1;
  return true;
}

void pack_host(gx_ctx const* c, HostPack& h) {
  int const nn = c->nn, ne = c->ne;
  h.nodes.resize(nn);
  for (int n = 0; n < nn; ++n) {
    NodeRec& r = h.nodes[n];
    for (int j = 0; j < 3; ++j) { r.x[j] = c->coords[3 * (size_t)n + j]; r.u[j] = 0.0; }
    r.p = 0.0;
    r.blk0 = (int32_t)c->nrow[n];
    r.nblk = (int32_t)(c->nrow[n + 1] - c->nrow[n]);
  }
  h.conn4.resize(ne);
  h.bpos.resize(ne);
  h.eset.clear();
  if (!c->eset.empty()) h.eset.resize(ne);
  for (int e = 0; e < ne; ++e) {
    int32_t const* c4 = &c->conn[4 * (size_t)e];
    h.conn4[e].x = c4[0]; h.conn4[e].y = c4[1]; h.conn4[e].z = c4[2]; h.conn4[e].w = c4[3];
    memcpy(&h.bpos[e], &c->bpos[16 * (size_t)e], 16);
    if (!h.eset.empty()) h.eset[e] = (uint8_t)c->eset[e];
  }
}

void materialise_crs(gx_ctx* c) {
  if (!c->rowptr.empty()) return;
  int const nn = c->nn;
  c->rowptr.resize(4 * (size_t)nn + 1);
  c->colind.resize(c->nnz);
  c->rowptr[0] = 0;
  for (int n = 0; n < nn; ++n) {
    int64_t const len = 4 * (c->nrow[n + 1] - c->nrow[n]);
    for (int i = 0; i < 4; ++i) c->rowptr[4 * (size_t)n + i + 1] = 16 * c->nrow[n] + (i + 1) * len;
  }
#pragma omp parallel for schedule(static)
  for (int n = 0; n < nn; ++n) {
    int64_t const nb = c->nrow[n + 1] - c->nrow[n];
    for (int i = 0; i < 4; ++i) {
      int32_t* dst = c->colind.data() + c->rowptr[4 * (size_t)n + i];
      for (int64_t j = 0; j < nb; ++j) {
        int32_t const b = c->ncol[c->nrow[n] + j];
        dst[4 * j] = 4 * b; dst[4 * j + 1] = 4 * b + 1; dst[4 * j + 2] = 4 * b + 2; dst[4 * j + 3] = 4 * b + 3;
      }
    }
  }
}

}  // namespace gx
