// hostcheck.cpp -- TEST-ONLY host build of goal_b200/csrc/element_math.cuh.
// Lets the CPU test-suite (no GPU in the build container) verify the arithmetic
// every CUDA thread runs against the oracle.  Not linked into, nor reachable
// from, the product library.
#include <cmath>

#include "../../goal_b200/csrc/element_math.cuh"
#include "../../goal_b200/csrc/tangent_record.cuh"

extern "C" int hc_element(int model, const double* x, const double* u, const double* p, const double* mat5,
                          const double* Fp_old, double eqps_old, int save, double* K /*16x16 row-major, dof=n*4+eq*/,
                          double* R /*16*/, double* sigma /*9*/, double* eqps, double* Fp /*9*/, int* wrote_Fp,
                          int* plastic) {
  gx::Material m = gx::make_material(mat5[0], mat5[1], mat5[2], mat5[3], mat5[4]);
  double X[4][3], U[4][3];
  for (int n = 0; n < 4; ++n) for (int j = 0; j < 3; ++j) { X[n][j] = x[3 * n + j]; U[n][j] = u[3 * n + j]; }
  gx::Core<double> c;
  bool wf = false;
  double Cp[6];
  gx::cp_inverse(Fp_old, Cp);
  int rc = model == 0 ? gx::element_core<gx::MODEL_NEOHOOKEAN>(X, U, p, m, Cp, eqps_old, (save & 1) != 0, sigma, *eqps, c)
                      : gx::element_core<gx::MODEL_J2>(X, U, p, m, Cp, eqps_old, (save & 1) != 0, sigma, *eqps, c);
  if (rc) return rc;
  if ((save & 1) && model == 1 && c.plastic) { gx::plastic_update(c.dN, Fp_old, Fp); wf = true; }
  *wrote_Fp = wf; *plastic = c.plastic;
  double ru[12], rp[4];
  gx::element_residual(c, ru, rp);
  for (int n = 0; n < 4; ++n) { for (int i = 0; i < 3; ++i) R[4 * n + i] = ru[3 * n + i]; R[4 * n + 3] = rp[n]; }
  if (save & 16) {  // the 42-double tangent record and the pair / diagonal contributions stage B builds from it
    double rec[gx::TREC];
    gx::pack_trec(c, rec);
    bool const tr = (save & 8) != 0;
    auto put = [&](int n, int mm, double const* b) {  // b = block (n,mm) of A, or of A^T when tr
      for (int i = 0; i < 4; ++i) for (int k = 0; k < 4; ++k) {
        if (!tr) K[(4 * n + i) * 16 + 4 * mm + k] = b[4 * i + k];
        else K[(4 * mm + k) * 16 + 4 * n + i] = b[4 * i + k];  // undo the transpose: K(row dof of A, col dof of A)
      }
    };
    for (int n = 0; n < 4; ++n) {
      double a[16] = {}, r4[4] = {};
      if (tr) gx::trec_diag_add_rec<true>(rec, n, a, r4); else gx::trec_diag_add_rec<false>(rec, n, a, r4);
      put(n, n, a);
      for (int i = 0; i < 4; ++i) R[4 * n + i] = r4[i];
      for (int mm = n + 1; mm < 4; ++mm) {
        double a1[16] = {}, a2[16] = {};
        if (tr) gx::trec_pair_add_rec<true>(rec, n, mm, a1, a2); else gx::trec_pair_add_rec<false>(rec, n, mm, a1, a2);
        put(n, mm, a1); put(mm, n, a2);
      }
    }
    return 0;
  }
  for (int mm = 0; mm < 4; ++mm) {
    gx::ColNode<double> cn;
    if (save & 2) gx::column_node_w(c, c.w[mm], cn);  // the form the tangent records use: r_m rebuilt from w_m
    else gx::column_node(c, c.w[mm], c.r[mm], cn);
    for (int n = 0; n < 4; ++n) {
      double blk[16];
      gx::RowNode<double> rn;
      gx::row_node(c, c.w[n], rn);
      if (save & 4) {  // the accumulating form the patch gather uses (fused multiply-adds into the accumulator)
        for (int q = 0; q < 16; ++q) blk[q] = 0.0;
        if (save & 8) {  // transposed accumulation, undone here
          double t[16] = {};
          gx::jacobian_block_add<true>(c, rn, cn, t);
          for (int i = 0; i < 4; ++i) for (int k = 0; k < 4; ++k) blk[4 * i + k] = t[4 * k + i];
        } else {
          gx::jacobian_block_add<false>(c, rn, cn, blk);
        }
      } else {
        gx::jacobian_block(c, rn, cn, blk);
      }
      for (int i = 0; i < 4; ++i) for (int k = 0; k < 4; ++k) K[(4 * n + i) * 16 + 4 * mm + k] = blk[4 * i + k];
    }
  }
  return 0;
}

extern "C" int hc_error_residual(int model, const double* x, const double* u, const double* p, const double* mat5,
                                 const double* Fp_old, double eqps_old, const double* zu, const double* zp,
                                 const double* zpc, double* R) {
  gx::Material m = gx::make_material(mat5[0], mat5[1], mat5[2], mat5[3], mat5[4]);
  double X[4][3], U[4][3], Z[4][3];
  for (int n = 0; n < 4; ++n) for (int j = 0; j < 3; ++j) { X[n][j] = x[3 * n + j]; U[n][j] = u[3 * n + j]; Z[n][j] = zu[3 * n + j]; }
  gx::Core<double> c;
  double sg[9], eq;
  double Cp[6];
  gx::cp_inverse(Fp_old, Cp);
  int rc = model == 0 ? gx::element_core<gx::MODEL_NEOHOOKEAN>(X, U, p, m, Cp, eqps_old, false, sg, eq, c)
                      : gx::element_core<gx::MODEL_J2>(X, U, p, m, Cp, eqps_old, false, sg, eq, c);
  if (rc) return rc;
  double ru[12], rp[4];
  gx::element_error_residual(c, Z, zp, zpc, ru, rp);
  for (int n = 0; n < 4; ++n) { for (int i = 0; i < 3; ++i) R[4 * n + i] = ru[3 * n + i]; R[4 * n + 3] = rp[n]; }
  return 0;
}

// von Mises stress and vol * d vm / d u of one element (element_von_mises): out[0] = vm, out[1] = vol, dvm[12]
extern "C" int hc_von_mises(int model, const double* x, const double* u, const double* p, const double* mat5,
                            const double* Fp_old, double eqps_old, double* out, double* dvm) {
  gx::Material m = gx::make_material(mat5[0], mat5[1], mat5[2], mat5[3], mat5[4]);
  double X[4][3], U[4][3];
  for (int n = 0; n < 4; ++n) for (int j = 0; j < 3; ++j) { X[n][j] = x[3 * n + j]; U[n][j] = u[3 * n + j]; }
  gx::Core<double> c;
  double sg[9], eq, Cp[6], d[4][3];
  gx::cp_inverse(Fp_old, Cp);
  int rc = model == 0 ? gx::element_core<gx::MODEL_NEOHOOKEAN>(X, U, p, m, Cp, eqps_old, false, sg, eq, c)
                      : gx::element_core<gx::MODEL_J2>(X, U, p, m, Cp, eqps_old, false, sg, eq, c);
  if (rc) return rc;
  out[0] = gx::element_von_mises(c, d);
  out[1] = c.vol;
  for (int n = 0; n < 4; ++n) for (int k = 0; k < 3; ++k) dvm[3 * n + k] = d[n][k];
  return 0;
}

extern "C" void hc_expm3(const double* A, double* o) { gx::expm3(A, o); }

// ---------------------------------------------------------------------------
// Whole-mesh emulation: the product's own setup code (gx_setup.cpp) and the
// product's own element body (gx_kernels.cuh), run colour by colour on the host.
// Checks indexing, the scatter map, the colour schedule and the SoA state layout
// before any GPU time is spent.
// ---------------------------------------------------------------------------
#include "../../goal_b200/csrc/gx_kernels.cuh"

template <int MODEL, int PASS, bool SAVE>
static int64_t run_colours(gx_ctx& c, gx::KParams& P) {
  int64_t plastic = 0;
  for (int k = 0; k < c.ncolors; ++k)
    for (int e = c.color_off[k]; e < c.color_off[k + 1]; ++e) plastic += gx::assemble_element<MODEL, PASS, SAVE>(P, e);
  return plastic;
}

// states in/out are AoS user order: sigma[ne*9], eqps[ne], eqps_old[ne], Fp[ne*9], Fp_old[ne*9]
extern "C" int hc_assemble(int model, int pass, int save, int nn, int ne, const int32_t* conn, const double* coords,
                           const double* mat5, const double* u, const double* p, const double* z5 /*[nn*5] or NULL*/,
                           double* sigma, double* eqps, const double* eqps_old, double* Fp, const double* Fp_old,
                           double* R, double* values, int64_t* nnz_out, int64_t* rowptr_out, int32_t* colind_out,
                           int32_t* ncolors, int64_t* plastic) {
  gx_ctx c;
  c.nn = nn; c.ne = ne; c.model = model; c.nsets = 1;
  c.conn.assign(conn, conn + 4 * (size_t)ne);
  c.coords.assign(coords, coords + 3 * (size_t)nn);
  c.mats[0] = gx::make_material(mat5[0], mat5[1], mat5[2], mat5[3], mat5[4]);
  int rc = gx::build_graph_and_schedule(&c);
  if (rc) return rc;
  if ((rc = gx::build_colouring(&c))) return rc;
  *nnz_out = c.nnz; *ncolors = c.ncolors;
  if (rowptr_out) {
    gx::materialise_crs(&c);
    std::copy(c.rowptr.begin(), c.rowptr.end(), rowptr_out);
    std::copy(c.colind.begin(), c.colind.end(), colind_out);
  }
  if (!R) return 0;
  gx::HostPack hp;
  gx::pack_host(&c, hp);
  for (int n = 0; n < nn; ++n) {
    for (int j = 0; j < 3; ++j) hp.nodes[n].u[j] = u[3 * (size_t)n + j];
    hp.nodes[n].p = p[n];
  }
  std::vector<gx::ZRec> z(nn);
  if (z5) for (int n = 0; n < nn; ++n) { for (int j = 0; j < 3; ++j) z[n].zu[j] = z5[5 * (size_t)n + j]; z[n].zp = z5[5 * (size_t)n + 3]; z[n].zpc = z5[5 * (size_t)n + 4]; }
  std::vector<double> sin((size_t)gx::STATE_IN * ne, 0.0), sout((size_t)gx::STATE_OUT * ne, 0.0);
  for (int e = 0; e < ne; ++e) {
    for (int k = 0; k < 9; ++k) sout[(size_t)gx::STATE_OUT * e + k] = sigma[9 * (size_t)e + k];
    if (model == 1) {
      sout[(size_t)gx::STATE_OUT * e + gx::SO_EQPS] = eqps[e]; sin[(size_t)gx::STATE_IN * e + 9] = eqps_old[e];
      for (int k = 0; k < 9; ++k) { sout[(size_t)gx::STATE_OUT * e + gx::SO_FP + k] = Fp[9 * (size_t)e + k]; sin[(size_t)gx::STATE_IN * e + k] = Fp_old[9 * (size_t)e + k]; }
    }
  }
  int err[2] = {0, 0};
  unsigned long long pl = 0;
  gx::KParams P;
  P.nodes = hp.nodes.data(); P.z = z.data(); P.conn = hp.conn4.data(); P.bpos = hp.bpos.data(); P.eset = nullptr;
  P.elems = c.perm.data(); P.adj_off = c.adj_off.data(); P.adj = c.adj.data();
  P.state_in = sin.data(); P.state_out = sout.data();
  P.R = R; P.values = values; P.err = err; P.plastic = &pl; P.e0 = 0; P.e1 = ne; P.nn = nn; P.max_nblk = c.max_nblk; P.pf_dist = 0; P.pf_elems = 0;
  for (int s = 0; s < GX_MAX_ELEM_SETS; ++s) P.mat[s] = c.mats[0];
  int64_t npl = 0;
  using namespace gx;
#define HC_RUN(M) \
  switch (pass) { \
    case PASS_RESIDUAL: npl = save ? run_colours<M, PASS_RESIDUAL, true>(c, P) : run_colours<M, PASS_RESIDUAL, false>(c, P); break; \
    case PASS_JACOBIAN: npl = save ? run_colours<M, PASS_JACOBIAN, true>(c, P) : run_colours<M, PASS_JACOBIAN, false>(c, P); break; \
    case PASS_JACOBIAN_T: npl = save ? run_colours<M, PASS_JACOBIAN_T, true>(c, P) : run_colours<M, PASS_JACOBIAN_T, false>(c, P); break; \
    default: npl = run_colours<M, PASS_ERROR, false>(c, P); }
  if (model == 1) { HC_RUN(MODEL_J2) } else { HC_RUN(MODEL_NEOHOOKEAN) }
  *plastic = npl;
  for (int e = 0; e < ne; ++e) {
    for (int k = 0; k < 9; ++k) sigma[9 * (size_t)e + k] = sout[(size_t)gx::STATE_OUT * e + k];
    if (model == 1) {
      eqps[e] = sout[(size_t)gx::STATE_OUT * e + gx::SO_EQPS];
      for (int k = 0; k < 9; ++k) Fp[9 * (size_t)e + k] = sout[(size_t)gx::STATE_OUT * e + gx::SO_FP + k];
    }
  }
  return err[0] ? 100 + err[0] : 0;
}

// ---------------------------------------------------------------------------
// Operation counts of the kernel's own formulation ("counted, not guessed"): the element
// math instantiated with a counting scalar.  add/sub, mul, div and special (sqrt, cbrt, fabs, compare)
// are tallied separately; a fused multiply-add executes as one instruction but is counted here as
// 1 mul + 1 add = 2 flops, the usual convention.
// ---------------------------------------------------------------------------
namespace {
struct Cnt {
  static long add, mul, dv, sp;
  double v;
  Cnt() : v(0) {}
  Cnt(double x) : v(x) {}
  explicit operator double() const { return v; }
};
long Cnt::add = 0, Cnt::mul = 0, Cnt::dv = 0, Cnt::sp = 0;
inline Cnt operator+(Cnt a, Cnt b) { ++Cnt::add; return Cnt(a.v + b.v); }
inline Cnt operator-(Cnt a, Cnt b) { ++Cnt::add; return Cnt(a.v - b.v); }
inline Cnt operator-(Cnt a) { return Cnt(-a.v); }
inline Cnt operator*(Cnt a, Cnt b) { ++Cnt::mul; return Cnt(a.v * b.v); }
inline Cnt operator/(Cnt a, Cnt b) { ++Cnt::dv; return Cnt(a.v / b.v); }
inline Cnt& operator+=(Cnt& a, Cnt b) { a = a + b; return a; }
inline Cnt& operator-=(Cnt& a, Cnt b) { a = a - b; return a; }
inline Cnt& operator*=(Cnt& a, Cnt b) { a = a * b; return a; }
inline bool operator>(Cnt a, Cnt b) { ++Cnt::sp; return a.v > b.v; }
inline bool operator<(Cnt a, Cnt b) { ++Cnt::sp; return a.v < b.v; }
inline bool operator==(Cnt a, Cnt b) { ++Cnt::sp; return a.v == b.v; }
inline bool operator<=(Cnt a, Cnt b) { ++Cnt::sp; return a.v <= b.v; }
inline Cnt sqrt(Cnt a) { ++Cnt::sp; return Cnt(std::sqrt(a.v)); }
inline Cnt cbrt(Cnt a) { ++Cnt::sp; return Cnt(std::cbrt(a.v)); }
inline Cnt fabs(Cnt a) { ++Cnt::sp; return Cnt(std::fabs(a.v)); }
// Material fields are doubles: mixed double*Cnt goes through Cnt(double)
}  // namespace

// out[0..3] = add, mul, div, special for: what = 0 element core (+ state save), 1 one incidence of the
// row-owner kernel (core + residual row + 4 blocks), 2 one whole element matrix (core + residual + 16 blocks)
extern "C" int hc_count_ops(int model, int what, int save, const double* x, const double* u, const double* p,
                            const double* mat5, const double* Fp_old, double eqps_old, long* out, int* plastic) {
  gx::Material m = gx::make_material(mat5[0], mat5[1], mat5[2], mat5[3], mat5[4]);
  Cnt X[4][3], U[4][3], Pn[4], Cp[6], sig[9], eq, FpO[9], FpN[9];
  double Cpd[6];
  gx::cp_inverse(Fp_old, Cpd);
  for (int n = 0; n < 4; ++n) { for (int j = 0; j < 3; ++j) { X[n][j] = x[3 * n + j]; U[n][j] = u[3 * n + j]; } Pn[n] = p[n]; }
  for (int k = 0; k < 6; ++k) Cp[k] = Cpd[k];
  for (int k = 0; k < 9; ++k) FpO[k] = Fp_old[k];
  Cnt::add = Cnt::mul = Cnt::dv = Cnt::sp = 0;
  gx::Core<Cnt> c;
  int rc = model == 0 ? gx::element_core<gx::MODEL_NEOHOOKEAN>(X, U, Pn, m, Cp, Cnt(eqps_old), save != 0, sig, eq, c)
                      : gx::element_core<gx::MODEL_J2>(X, U, Pn, m, Cp, Cnt(eqps_old), save != 0, sig, eq, c);
  if (rc) return rc;
  *plastic = c.plastic;
  if (save && model == 1 && c.plastic) gx::plastic_update(c.dN, FpO, FpN);
  if (what == 1) {
    Cnt r4[4], blk[16];
    gx::element_residual_row(c, c.w[1], r4);
    gx::RowNode<Cnt> rn;
    gx::row_node(c, c.w[1], rn);
    for (int mm = 0; mm < 4; ++mm) { gx::ColNode<Cnt> cn; gx::column_node(c, c.w[mm], c.r[mm], cn); gx::jacobian_block(c, rn, cn, blk); }
  } else if (what == 2) {
    Cnt ru[12], rp[4], blk[16];
    gx::element_residual(c, ru, rp);
    gx::RowNode<Cnt> rn[4];
    for (int n = 0; n < 4; ++n) gx::row_node(c, c.w[n], rn[n]);
    for (int mm = 0; mm < 4; ++mm) {
      gx::ColNode<Cnt> cn;
      gx::column_node(c, c.w[mm], c.r[mm], cn);
      for (int n = 0; n < 4; ++n) gx::jacobian_block(c, rn[n], cn, blk);
    }
  }
  out[0] = Cnt::add; out[1] = Cnt::mul; out[2] = Cnt::dv; out[3] = Cnt::sp;
  return 0;
}

// ---------------------------------------------------------------------------
// CPU replay of the Jacobian pass's stage B (the default GPU schedule): builds the patch schedule with the product's
// own host code (gx_setup.cpp), then interprets its words the way patch_pair_kernel does -- staged record slots from
// the bulk-copy runs, PAIR / DIAG / ZERO work items of up to 8 contributions, primaries adding their secondaries'
// partial sums in order, one writer per 4x4 block and per node's residual entries -- with the device's own record
// packing and contribution arithmetic (tangent_record.cuh) compiled for the host.
//   values [nnz] and R [4 nn] must come in NaN-filled; every entry the schedule owns is written exactly once.
// ---------------------------------------------------------------------------
template <bool TRANSPOSE>
static int replay_patch_pairs(gx_ctx& c, std::vector<double> const& rec, double* R, double* values) {
  using namespace gx;
  uint32_t const* sched = c.patch_sched.data();
  for (int pch = 0; pch < c.n_patches; ++pch) {
    uint32_t const* w = sched + (size_t)pch * PATCH_WORDS;
    int const n_recs = (int)w[0], n_runs = (int)w[2];
    std::vector<int64_t> slot_elem(PATCH_RECS, -1);  // what the bulk copies stage: runs of consecutive elements
    int staged = 0;
    uint32_t const* wi = w + 4;
    uint32_t const* wo = wi + 4 * PATCH_THREADS;
    uint32_t const* wr = wo + 4 * PATCH_THREADS;
    for (int r = 0; r < n_runs; ++r)
      for (uint32_t j = 0; j < (wr[2 * r + 1] >> 8); ++j) {
        uint32_t const sl = (wr[2 * r + 1] & 0xffu) + j;
        if (sl >= (uint32_t)PATCH_RECS || slot_elem[sl] >= 0) return 10;
        slot_elem[sl] = (int64_t)wr[2 * r] + j; ++staged;
      }
    if (staged != n_recs) return 10;
    static thread_local double acc1[PATCH_THREADS][16], acc2[PATCH_THREADS][16], r4[PATCH_THREADS][4], parts[PATCH_PARTS][PATCH_PART_LD];
    for (int t = 0; t < PATCH_THREADS; ++t) {
      uint32_t const* ot = wo + 4 * t;
      int const kind = (int)(ot[3] & 3u), type = (int)((ot[3] >> 2) & 3u);
      for (int k = 0; k < 16; ++k) acc1[t][k] = acc2[t][k] = 0.0;
      for (int k = 0; k < 4; ++k) r4[t][k] = 0.0;
      if (kind == 0) continue;
      for (int k = 0; k < PATCH_ITEM_LEN; ++k) {
        uint32_t const ent = (wi[4 * t + k / 2] >> (16 * (k & 1))) & 0xffffu;
        if (!(ent & 0x8000u)) continue;
        if (type == 0) return 14;  // a ZERO item has no contributions
        int const slot = (int)(ent & 0xffu), n = (int)((ent >> 10) & 3u), m = (int)((ent >> 8) & 3u);
        if (slot_elem[slot] < 0) return 11;
        double const* rp = rec.data() + (size_t)TREC * (size_t)slot_elem[slot];
        if (type == 2) { if (n == m) return 15; trec_pair_add_rec<TRANSPOSE>(rp, n, m, acc1[t], acc2[t]); }
        else { if (n != m) return 15; trec_diag_add_rec<TRANSPOSE>(rp, n, acc1[t], r4[t]); }
      }
      if (kind == 2) {
        int const part = (int)((ot[3] >> 4) & 0xffu);
        if (part >= PATCH_PARTS) return 12;
        for (int k = 0; k < 16; ++k) { parts[part][k] = acc1[t][k]; parts[part][16 + k] = acc2[t][k]; }
        for (int k = 0; k < 4; ++k) parts[part][32 + k] = r4[t][k];
      }
    }
    for (int t = 0; t < PATCH_THREADS; ++t) {
      uint32_t const* ot = wo + 4 * t;
      if ((ot[3] & 3u) != 1u) continue;
      int const type = (int)((ot[3] >> 2) & 3u);
      int const part = (int)((ot[3] >> 4) & 0xffu), nsec = (int)((ot[3] >> 12) & 0x3fu);
      for (int s2 = 0; s2 < nsec; ++s2) {
        for (int k = 0; k < 16; ++k) { acc1[t][k] += parts[part + s2][k]; acc2[t][k] += parts[part + s2][16 + k]; }
        for (int k = 0; k < 4; ++k) r4[t][k] += parts[part + s2][32 + k];
      }
      auto put = [&](int64_t blk0, int j, int nb, double const* a) -> bool {
        for (int i = 0; i < 4; ++i)
          for (int k = 0; k < 4; ++k) {
            double& v = values[16 * blk0 + 4 * j + (int64_t)i * (4 * nb) + k];
            if (v == v) return false;  // a second writer (the array comes in NaN-filled)
            v = a[4 * i + k];
          }
        return true;
      };
      if (!put((int64_t)ot[0], (int)(ot[2] & 0xffu), (int)((ot[2] >> 8) & 0xffu), acc1[t])) return 13;
      if (type == 2 && !put((int64_t)ot[1], (int)((ot[2] >> 16) & 0xffu), (int)(ot[2] >> 24), acc2[t])) return 13;
      if (type == 1) for (int k = 0; k < 4; ++k) {
        double& v = R[4 * (size_t)ot[1] + k];
        if (v == v) return 13;
        v = r4[t][k];
      }
    }
  }
  return 0;
}

extern "C" int hc_patch_gather(int model, int transpose, int nn, int ne, const int32_t* conn, const double* coords,
                               const double* mat5, const double* u, const double* p, const double* eqps_old,
                               const double* Fp_old, double* R, double* values, int32_t* n_patches) {
  gx_ctx c;
  c.nn = nn; c.ne = ne; c.model = model; c.nsets = 1;
  c.conn.assign(conn, conn + 4 * (size_t)ne);
  c.coords.assign(coords, coords + 3 * (size_t)nn);
  gx::Material const mat = gx::make_material(mat5[0], mat5[1], mat5[2], mat5[3], mat5[4]);
  int rc = gx::build_graph_and_schedule(&c);
  if (rc) return rc;
  c.nrow_x = c.nrow;  // single part: no phantom blocks (comm_setup_lists does this in the library)
  if (!gx::build_patch_schedule(&c)) return 20;
  gx::flatten_patch_schedule(&c);
  *n_patches = c.n_patches;
  std::vector<double> rec((size_t)gx::TREC * ne);
  for (int e = 0; e < ne; ++e) {  // stage A: the element core and its tangent record, once per element
    double X[4][3], U[4][3], P4[4], Cp[6], sg[9], eq;
    for (int n = 0; n < 4; ++n) {
      int const a = conn[4 * (size_t)e + n];
      for (int j = 0; j < 3; ++j) { X[n][j] = coords[3 * (size_t)a + j]; U[n][j] = u[3 * (size_t)a + j]; }
      P4[n] = p[a];
    }
    if (model == 1) gx::cp_inverse(Fp_old + 9 * (size_t)e, Cp);
    gx::Core<double> core;
    rc = model == 0 ? gx::element_core<gx::MODEL_NEOHOOKEAN>(X, U, P4, mat, Cp, 0.0, false, sg, eq, core)
                    : gx::element_core<gx::MODEL_J2>(X, U, P4, mat, Cp, eqps_old[e], false, sg, eq, core);
    if (rc) return rc;
    gx::pack_trec(core, rec.data() + (size_t)gx::TREC * e);
  }
  int64_t const nnz = c.nnz;
  for (int64_t i = 0; i < nnz; ++i) values[i] = std::nan("");
  for (int i = 0; i < 4 * nn; ++i) R[i] = std::nan("");
  rc = transpose ? replay_patch_pairs<true>(c, rec, R, values) : replay_patch_pairs<false>(c, rec, R, values);
  if (rc) return rc;
  for (int64_t i = 0; i < nnz; ++i) if (values[i] != values[i]) return 16;  // an entry nobody wrote
  for (int i = 0; i < 4 * nn; ++i) if (R[i] != R[i]) return 17;
  return 0;
}

// ---------------------------------------------------------------------------
// CPU replay of the block-reduced residual schedule (build_residual_schedule; elem_residual_block_kernel +
// node_partial_sum_kernel read it exactly like this): rvec [ne][4][4] are the element residual lines (any numbers),
// R [4 nn] comes back as their per-node sums.  Every entry of R and of the partial buffer must be written exactly once,
// every incidence used exactly once.  stats = {blocks, schedule words, partial sums, nodes finished by the second kernel}.
// ---------------------------------------------------------------------------
static int replay_residual_schedule(gx_ctx& c, const double* rvec, double* R, int64_t* stats) {
  using namespace gx;
  int const nn = c.nn, ne = c.ne;
  int32_t const* conn = c.conn.data();
  if (!build_residual_schedule(&c)) return 20;
  int const nb = (ne + RES_BLOCK - 1) / RES_BLOCK;
  std::vector<uint32_t> sched;
  for (auto const& v : c.res_chunks) sched.insert(sched.end(), v.begin(), v.end());
  if (c.res_boff.size() != (size_t)nb + 1 || c.res_boff[nb] != sched.size()) return 21;
  if (stats) { stats[0] = nb; stats[1] = (int64_t)sched.size(); stats[2] = c.res_npartial; stats[3] = (int64_t)c.res_pnode.size(); }
  std::vector<double> partial(4 * (size_t)std::max<int64_t>(c.res_npartial, 1), std::nan(""));
  for (int i = 0; i < 4 * nn; ++i) R[i] = std::nan("");
  std::vector<uint8_t> used(4 * (size_t)ne, 0);
  for (int b = 0; b < nb; ++b) {
    uint32_t const* w = sched.data() + c.res_boff[b];
    int const S = (int)w[0], cnt = std::min(RES_BLOCK, ne - b * RES_BLOCK);
    if (w[1] != c.res_boff[b + 1] - c.res_boff[b] || (w[1] & 3u) || w[1] > (uint32_t)RES_MAX_WORDS || S > 4 * cnt) return 22;
    uint16_t const* ent = reinterpret_cast<uint16_t const*>(w + RES_HDR + 2 * S);
    uint32_t covered = 0;
    for (int s = 0; s < S; ++s) {
      uint32_t const w0 = w[RES_HDR + 2 * s], w1 = w[RES_HDR + 2 * s + 1], first = w1 & 0xffffu, n = w1 >> 16;
      if (first != covered || n == 0) return 23;  // slots tile the entry list
      covered += n;
      bool const complete = (w0 & 0x80000000u) != 0;
      uint32_t const tgt = w0 & 0x7fffffffu;
      int32_t node = -1;
      double acc[4] = {0, 0, 0, 0};
      uint32_t prev = 0;
      for (uint32_t k = first; k < first + n; ++k) {
        uint32_t const row = ent[k] >> 3, n2 = (ent[k] & 7u) ^ (row & 7u);  // swizzled chunk -> local element, 2 * local node
        if (row >= (uint32_t)cnt || (n2 & 1u) || (k > first && row <= prev)) return 24;  // ascending elements inside a slot
        prev = row;
        size_t const g = 4 * (size_t)b * RES_BLOCK + 4 * row + (n2 >> 1);  // 4 * element + local node
        if (used[g]++) return 25;
        if (node < 0) node = conn[g]; else if (node != conn[g]) return 26;  // one node per slot
        for (int q = 0; q < 4; ++q) acc[q] += rvec[4 * g + q];
      }
      if (complete) {
        if ((int32_t)tgt != node || (uint32_t)(c.adj_off[node + 1] - c.adj_off[node]) != n) return 27;
        for (int q = 0; q < 4; ++q) { if (R[4 * (size_t)node + q] == R[4 * (size_t)node + q]) return 28; R[4 * (size_t)node + q] = acc[q]; }
      } else {
        if ((int64_t)tgt >= c.res_npartial) return 29;
        auto it = std::lower_bound(c.res_pnode.begin(), c.res_pnode.end(), node);
        if (it == c.res_pnode.end() || *it != node) return 30;
        size_t const pi = it - c.res_pnode.begin();
        if (tgt < c.res_poff[pi] || tgt >= c.res_poff[pi + 1]) return 31;  // inside the node's run
        for (int q = 0; q < 4; ++q) { if (partial[4 * (size_t)tgt + q] == partial[4 * (size_t)tgt + q]) return 32; partial[4 * (size_t)tgt + q] = acc[q]; }
      }
    }
    if (covered != 4u * (uint32_t)cnt) return 33;
  }
  for (size_t i = 0; i < c.res_pnode.size(); ++i) {
    int32_t const a = c.res_pnode[i];
    for (int q = 0; q < 4; ++q) {
      double acc = 0.0;
      for (uint32_t p = c.res_poff[i]; p < c.res_poff[i + 1]; ++p) {
        if (partial[4 * (size_t)p + q] != partial[4 * (size_t)p + q]) return 34;  // a partial sum nobody wrote
        acc += partial[4 * (size_t)p + q];
      }
      if (R[4 * (size_t)a + q] == R[4 * (size_t)a + q]) return 35;
      R[4 * (size_t)a + q] = acc;
    }
  }
  for (size_t g = 0; g < 4 * (size_t)ne; ++g) if (used[g] != 1) return 36;
  for (int i = 0; i < 4 * nn; ++i) if (R[i] != R[i]) return 37;
  return 0;
}

extern "C" int hc_residual_schedule(int nn, int ne, const int32_t* conn, const double* coords, const double* rvec, double* R,
                                    int64_t* stats) {
  gx_ctx c;
  c.nn = nn; c.ne = ne; c.model = 0; c.nsets = 1;
  c.conn.assign(conn, conn + 4 * (size_t)ne);
  c.coords.assign(coords, coords + 3 * (size_t)nn);
  int rc = gx::build_graph_and_schedule(&c);
  if (rc) return rc;
  return replay_residual_schedule(c, rvec, R, stats);
}

// The whole block-reduced residual / error-localisation pass on the CPU: the element residual lines from the device's
// own element code (element_core + element_residual, or element_error_residual with the adjoint weights z5 =
// [nn][zu(3), zp, zpc]), summed through the schedule exactly as the two kernels do.
extern "C" int hc_residual_blocks(int model, int nn, int ne, const int32_t* conn, const double* coords, const double* mat5,
                                  const double* u, const double* p, const double* z5, const double* eqps_old,
                                  const double* Fp_old, double* R) {
  gx_ctx c;
  c.nn = nn; c.ne = ne; c.model = model; c.nsets = 1;
  c.conn.assign(conn, conn + 4 * (size_t)ne);
  c.coords.assign(coords, coords + 3 * (size_t)nn);
  gx::Material const mat = gx::make_material(mat5[0], mat5[1], mat5[2], mat5[3], mat5[4]);
  int rc = gx::build_graph_and_schedule(&c);
  if (rc) return rc;
  std::vector<double> rvec(16 * (size_t)ne);
  for (int e = 0; e < ne; ++e) {
    double X[4][3], U[4][3], P4[4], Cp[6], sg[9], eq, zu[4][3], zp[4], zpc[4];
    for (int n = 0; n < 4; ++n) {
      int const a = conn[4 * (size_t)e + n];
      for (int j = 0; j < 3; ++j) { X[n][j] = coords[3 * (size_t)a + j]; U[n][j] = u[3 * (size_t)a + j]; }
      P4[n] = p[a];
      if (z5) { for (int j = 0; j < 3; ++j) zu[n][j] = z5[5 * (size_t)a + j]; zp[n] = z5[5 * (size_t)a + 3]; zpc[n] = z5[5 * (size_t)a + 4]; }
    }
    if (model == 1) gx::cp_inverse(Fp_old + 9 * (size_t)e, Cp);
    gx::Core<double> core;
    rc = model == 0 ? gx::element_core<gx::MODEL_NEOHOOKEAN>(X, U, P4, mat, Cp, 0.0, false, sg, eq, core)
                    : gx::element_core<gx::MODEL_J2>(X, U, P4, mat, Cp, eqps_old[e], false, sg, eq, core);
    if (rc) return rc;
    double ru[12], rp[4];
    if (z5) gx::element_error_residual(core, zu, zp, zpc, ru, rp);
    else gx::element_residual(core, ru, rp);
    for (int n = 0; n < 4; ++n) { for (int i = 0; i < 3; ++i) rvec[16 * (size_t)e + 4 * n + i] = ru[3 * n + i]; rvec[16 * (size_t)e + 4 * n + 3] = rp[n]; }
  }
  return replay_residual_schedule(c, rvec.data(), R, nullptr);
}
