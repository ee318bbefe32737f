// goal_gx.hpp -- C++ host mirror of the reference's assembly-facing classes over the C-ABI.
//
// The reference is compiled C++ (serial per MPI rank); its toolchain (Trilinos + SCOREC) is absent
// here, so this header plays the role of the code a Goal maintainer would write inside
// goal_disc / goal_sol_info / goal_primal / goal_nested_adjoint to route the hot path through
// libgoal_b200.so.  Names, argument meaning and error behaviour follow the reference:
//
//   gx::Disc         goal::Disc         maps / graphs / get_lid(s)          src/goal_disc.hpp:27-93
//   gx::SolInfo      goal::SolInfo      owned/ghost R, dRdu; zero / gather  src/goal_sol_info.hpp:12-39
//   gx::States       goal::States       get/set by field name, update()     src/goal_states.hpp:13-23
//   gx::Primal       goal::Primal       compute_resid / compute_jacob       src/goal_primal.cpp:75-109
//   gx::NestedAdjoint (localize part)   compute_adjoint / localize          src/goal_nested_adjoint.cpp:163-234
//   gx::compute_error / sum_contribs    goal::compute_error / sum_contribs  src/goal_error.cpp:7-56
//   gx::Functional   goal::Functional + QoI evaluators                      src/goal_functional.cpp:30-70, goal_qoi.cpp
//   gx::set_tbcs / set_ibcs / set_*_dbcs                                    src/goal_tbcs.cpp, goal_ibcs.cpp, goal_dbcs.cpp
//   gx::add_soln, gx::get_iso_target_size                                   src/goal_disc.cpp:398-422, goal_size_field.cpp
//
// Errors: the reference calls goal::fail(), which prints and abort()s (src/goal_control.cpp:91-99).
// Here every non-zero gx_status is turned into gx::fail(), which throws std::runtime_error by default
// (set GX_FAIL_ABORTS to get the reference's abort()).
#ifndef GOAL_GX_HPP
#define GOAL_GX_HPP

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/goal_b200.h"

namespace gx {

using LO = int32_t;  // src/goal_data_types.hpp:13
using GO = int64_t;  // src/goal_data_types.hpp:14
enum EvalModes { NONE = GX_MODE_NONE, PRIMAL = GX_MODE_PRIMAL, ADJOINT = GX_MODE_ADJOINT };  // goal_eval_modes.hpp:6

[[noreturn]] inline void fail(std::string const& msg) {
#ifdef GX_FAIL_ABORTS
  std::fprintf(stderr, "GOAL FAILED: %s\n", msg.c_str());
  std::abort();
#else
  throw std::runtime_error(msg);
#endif
}

struct Material { double E, nu, K, Y, c0; };  // one per elem set (src/goal_J2.cpp:12-24)

// What Disc::build_data produces for one part (src/goal_disc.cpp:224-235).
class Disc {
 public:
  Disc(std::vector<double> coords_, std::vector<LO> conn_, std::string model_, std::vector<Material> mats,
       std::vector<LO> elem_set = {}, int device = 0)
      : coords(std::move(coords_)), conn(std::move(conn_)), model(std::move(model_)) {
    gx_desc d{};
    d.n_nodes = (LO)(coords.size() / 3);
    d.n_elems = (LO)(conn.size() / 4);
    d.conn = conn.data();
    d.coords = coords.data();
    d.elem_set = elem_set.empty() ? nullptr : elem_set.data();
    d.n_elem_sets = (LO)mats.size();
    if (model == "neohookean") d.model = GX_MODEL_NEOHOOKEAN;
    else if (model == "J2") d.model = GX_MODEL_J2;
    else fail("unknown model: " + model);  // src/goal_mechanics.cpp:118
    std::vector<double> m5;
    for (auto const& m : mats) m5.insert(m5.end(), {m.E, m.nu, m.K, m.Y, m.c0});
    d.materials = m5.data();
    d.device = device;
    if (gx_create(&d, &ctx)) fail(gx_last_error(nullptr));
    int64_t nnz_;
    if (gx_graph(ctx, &nnz_, &rowptr_, &colind_)) fail(gx_last_error(ctx));
    nnz = nnz_;
  }
  ~Disc() { gx_destroy(ctx); }
  Disc(Disc const&) = delete;
  Disc& operator=(Disc const&) = delete;

  int get_num_dims() const { return 3; }
  int get_num_eqs() const { return 4; }                               // src/goal_disc.cpp:260
  LO get_num_nodes() const { return (LO)(coords.size() / 3); }
  LO get_num_elems() const { return (LO)(conn.size() / 4); }
  LO get_lid(LO elem, int n, int eq) const { return conn[4 * (size_t)elem + n] * 4 + eq; }  // src/goal_disc.cpp:203-207
  void get_lids(LO elem, std::vector<LO>& lids) const {                // src/goal_disc.cpp:214-222
    lids.resize(16);
    for (int n = 0; n < 4; ++n) for (int eq = 0; eq < 4; ++eq) lids[4 * n + eq] = get_lid(elem, n, eq);
  }
  // ghost graph in Tpetra's local layout (src/goal_disc.cpp:307-332)
  int64_t get_nnz() const { return nnz; }
  const int64_t* get_rowptr() const { return rowptr_; }
  const LO* get_colind() const { return colind_; }

  gx_ctx* ctx = nullptr;
  std::vector<double> coords;
  std::vector<LO> conn;
  std::string model;
  int64_t nnz = 0;

 private:
  const int64_t* rowptr_ = nullptr;
  const LO* colind_ = nullptr;
};

// owned == ghost for a single part; the vectors are what Tpetra's get1dView would expose.
struct LinearObj {
  std::vector<double> R;      // [4*n_nodes]
  std::vector<double> dRdu;   // CRS values, order of Disc::get_colind
};

class SolInfo {
 public:
  explicit SolInfo(Disc* d) : disc(d) {
    ghost.R.assign(4 * (size_t)d->get_num_nodes(), 0.0);
    ghost.dRdu.assign((size_t)d->get_nnz(), 0.0);
  }
  Disc* get_disc() { return disc; }
  void zero_R() { ghost.R.assign(ghost.R.size(), 0.0); }
  void zero_all() { zero_R(); ghost.dRdu.assign(ghost.dRdu.size(), 0.0); }
  void gather_R() {}    // Export ghost->owned, ADD (src/goal_sol_info.cpp:33-35): identity on one part;
  void gather_all() {}  // with several parts this is gx_reduce_interfaces + gx_fetch_owned
  LinearObj ghost;

 private:
  Disc* disc;
};

class States {  // src/goal_states.cpp:21-57, 130-141
 public:
  explicit States(Disc* d) : disc(d) {}
  void get(const char* name, std::vector<double>& v) const {
    v.resize((size_t)disc->get_num_elems() * (is_tensor(name) ? 9 : 1));
    if (gx_get_state(disc->ctx, name, v.data())) fail(gx_last_error(disc->ctx));
  }
  void set(const char* name, std::vector<double> const& v) {
    if (v.size() != (size_t)disc->get_num_elems() * (is_tensor(name) ? 9 : 1)) fail(std::string("bad state size: ") + name);
    if (gx_set_state(disc->ctx, name, v.data())) fail(gx_last_error(disc->ctx));
  }
  void update() { if (gx_update_states(disc->ctx)) fail(gx_last_error(disc->ctx)); }

 private:
  static bool is_tensor(std::string const& n) { return n == "sigma" || n == "Fp" || n == "Fp_old"; }
  Disc* disc;
};

// The two entry points of the primal problem, minus boundary conditions (the caller applies
// set_tbcs / set_ibcs to the ghost R before gather and set_*_dbcs after, exactly as in
// src/goal_primal.cpp:81-88 and :98-106).
class Primal {
 public:
  explicit Primal(Disc* d) : disc(d), sol_info(d) {}
  SolInfo* get_sol_info() { return &sol_info; }
  void set_solution(std::vector<double> const& u, std::vector<double> const& p) {  // fields "u","p" (goal_disc.cpp:398-422)
    if (gx_set_solution(disc->ctx, u.data(), p.data())) fail(gx_last_error(disc->ctx));
  }
  void compute_resid() {  // zero_R + assemble(residual), save=true (goal_primal.cpp:49, 81-83)
    if (gx_compute_residual(disc->ctx, 1, sol_info.ghost.R.data())) fail(gx_last_error(disc->ctx));
  }
  void compute_jacob() {  // zero_all + assemble(jacobian), save=true (goal_primal.cpp:50, 98-101)
    if (gx_compute_jacobian(disc->ctx, GX_MODE_PRIMAL, 1, sol_info.ghost.R.data(), sol_info.ghost.dRdu.data()))
      fail(gx_last_error(disc->ctx));
  }

 private:
  Disc* disc;
  SolInfo sol_info;
};

class NestedAdjoint {
 public:
  explicit NestedAdjoint(Disc* nested) : disc(nested), sol_info(nested) {}
  SolInfo* get_sol_info() { return &sol_info; }
  void compute_adjoint() {  // FADT chain, ADJOINT scatter, save=false (goal_nested_adjoint.cpp:121-123, 169-172)
    if (gx_compute_jacobian(disc->ctx, GX_MODE_ADJOINT, 0, sol_info.ghost.R.data(), sol_info.ghost.dRdu.data()))
      fail(gx_last_error(disc->ctx));
  }
  void localize(std::vector<double> const& zu_diff, std::vector<double> const& zp_diff, std::vector<double> const& zp_coarse) {
    if (gx_localize_error(disc->ctx, zu_diff.data(), zp_diff.data(), zp_coarse.data(), sol_info.ghost.R.data()))
      fail(gx_last_error(disc->ctx));  // goal_nested_adjoint.cpp:224-226
  }

 private:
  Disc* disc;
  SolInfo sol_info;
};

// compute_error + Nested::set_error (src/goal_error.cpp:7-35, goal_nested.cpp:395-412); returns sum_contribs (:37-56)
inline double compute_error(Disc* nested, std::vector<double> const& u_error, std::vector<double> const& p_error,
                            std::vector<LO> const& parent, LO n_parent, std::vector<double>& e_nested,
                            std::vector<double>& e_base) {
  e_nested.resize(nested->get_num_elems());
  e_base.resize(n_parent);
  double bound = 0.0;
  if (gx_element_error(nested->ctx, u_error.data(), p_error.data(), parent.empty() ? nullptr : parent.data(), n_parent,
                       e_nested.data(), parent.empty() ? nullptr : e_base.data(), &bound))
    fail(gx_last_error(nested->ctx));
  return bound;
}

// Functional (src/goal_functional.cpp:30-70) over Mechanics::build_functional's evaluators
// (src/goal_mechanics.cpp:149-167).  type is the yaml string; compute() is the ST chain (value only),
// compute_adjoint_rhs() the FADT chain that also fills dMdu (QoI<FADT>::scatter, src/goal_qoi.cpp:63-76).
class Functional {
 public:
  Functional(Disc* d, std::string const& type, int elem_set = 0, double rho = 0.0, LO point_node = -1, int point_idx = 0)
      : disc(d) {
    q = gx_qoi{};
    if (type == "avg disp") q.type = GX_QOI_AVG_DISP;
    else if (type == "avg disp subdomain") q.type = GX_QOI_AVG_DISP_SUBDOMAIN;
    else if (type == "avg vm") q.type = GX_QOI_AVG_VM;
    else if (type == "max vm") q.type = GX_QOI_KS_VM;
    else if (type == "point wise") q.type = GX_QOI_POINT_WISE;
    else fail("unknown functional type: " + type);  // src/goal_mechanics.cpp:165
    q.elem_set = elem_set; q.rho = rho; q.point_node = point_node; q.point_idx = point_idx;
  }
  void compute() {
    q.ks_scale = 0.0;
    if (gx_functional(disc->ctx, &q, &value, nullptr)) fail(gx_last_error(disc->ctx));
  }
  void compute_adjoint_rhs(std::vector<double>& dMdu) {
    dMdu.resize(4 * (size_t)disc->get_num_nodes());
    q.ks_scale = 0.0;
    if (gx_functional(disc->ctx, &q, &value, dMdu.data())) fail(gx_last_error(disc->ctx));
  }
  double get_value() const { return value; }

 private:
  Disc* disc;
  gx_qoi q;
  double value = 0.0;
};

// Boundary terms on the device-resident result of the last compute call, in the reference's order
// (src/goal_primal.cpp:84-88, 102-106): set_tbcs, set_ibcs, [gather], set_resid_dbcs / set_jac_dbcs.
inline void set_tbcs(Disc* d, std::vector<LO> const& side_nodes, std::vector<double> const& traction) {
  if (traction.size() != side_nodes.size()) fail("set_tbcs: one traction vector per side expected");
  if (gx_apply_tbcs(d->ctx, (LO)(side_nodes.size() / 3), side_nodes.data(), traction.data())) fail(gx_last_error(d->ctx));
}
inline void set_ibcs(Disc* d, std::vector<LO> const& side_nodes, double scale, double const center[3]) {
  if (gx_apply_ibcs(d->ctx, (LO)(side_nodes.size() / 3), side_nodes.data(), scale, center)) fail(gx_last_error(d->ctx));
}
// BForce<T>::at_point behind MResidual (src/goal_bforce.cpp:58-68, goal_mechanics.cpp:132-136; error chain :204-208):
// b = the body-force expression at every element's integration point, [n_elems * 3]
inline void set_bforce(Disc* d, std::vector<double> const& b, bool error_weights = false) {
  if (b.size() != 3 * (size_t)d->get_num_elems()) fail("set_bforce: one force vector per element expected");
  if (gx_apply_bforce(d->ctx, b.data(), error_weights ? 1 : 0)) fail(gx_last_error(d->ctx));
}
inline void set_resid_dbcs(Disc* d, std::vector<LO> const& rows, std::vector<double> const& g) {
  if (gx_apply_dbcs(d->ctx, (LO)rows.size(), rows.data(), g.data(), 0)) fail(gx_last_error(d->ctx));
}
inline void set_jac_dbcs(Disc* d, std::vector<LO> const& rows, std::vector<double> const& g) {
  if (gx_apply_dbcs(d->ctx, (LO)rows.size(), rows.data(), g.data(), 1)) fail(gx_last_error(d->ctx));
}
// Disc::add_soln (src/goal_disc.cpp:398-422)
inline void add_soln(Disc* d, std::vector<double> const& du) {
  if (du.size() != 4 * (size_t)d->get_num_nodes()) fail("add_soln: bad vector size");
  if (gx_add_solution(d->ctx, du.data()) || gx_sync_solution(d->ctx)) fail(gx_last_error(d->ctx));
}
// get_iso_target_size (src/goal_size_field.cpp:140-150), single part
inline void get_iso_target_size(Disc* d, std::vector<double> const& e_elem, int target, std::vector<double>& vtx_size) {
  vtx_size.resize(d->get_num_nodes());
  double G = 0.0;
  if (gx_size_field(d->ctx, e_elem.data(), target, 1, &G, vtx_size.data(), nullptr)) fail(gx_last_error(d->ctx));
}

}  // namespace gx
#endif
