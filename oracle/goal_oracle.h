/* goal_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a plain, serial, CPU restatement of
 * the reference's (bgranzow/goal) finite-element assembly hot path.  It is the
 * checker the CUDA path is compared against; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  Nothing under
 * goal_b200/ links, imports or calls it.
 *
 * Parity pinning: the reference cannot be compiled in this environment (it needs
 * MPI + Trilinos (Sacado, MiniTensor, Tpetra) + SCOREC/core, none present, none
 * version-pinned: /root/reference/cmake/dependencies.cmake:1-19).  The oracle is
 * pinned END-TO-END against the reference's own golden functional values
 * (example/primal/{neohookean_uniaxial,J2_uniaxial,J2_traction}_3D.yaml
 * `regression:` blocks) by tests/test_oracle_goldens.py.  Element-level values
 * and the error-localisation path (goal_error / goal_*_adjoint) have no reference
 * golden: for those rows parity is UNPINNED and only oracle-internal identities
 * are checked.
 *
 * All arrays are caller-owned unless noted.  Indices: LO = int32, GO = int64
 * (reference: src/goal_data_types.hpp:13-14).  DOF layout: dof = node*4 + eq,
 * eq 0..2 = u, eq 3 = p (src/goal_disc.cpp:195-201).
 */
#ifndef GOAL_ORACLE_H
#define GOAL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct go_ctx go_ctx;

enum { GO_MODEL_NEOHOOKEAN = 0, GO_MODEL_J2 = 1 };
/* scatter modes: src/goal_eval_modes.hpp:6 */
enum { GO_MODE_NONE = 0, GO_MODE_PRIMAL = 1, GO_MODE_ADJOINT = 2 };

/* materials: per elem set {E, nu, K, Y, c0} (src/goal_J2.cpp:54-64) */
go_ctx* go_create(int n_nodes, int n_elems, const int32_t* conn /*[Ne*4]*/,
                  const double* coords /*[Nn*3]*/, const int32_t* elem_set /*[Ne] or NULL*/,
                  int n_sets, int model, const double* materials /*[n_sets*5]*/);
void go_destroy(go_ctx*);

/* CRS graph in the reference's ghost layout (src/goal_disc.cpp:307-332):
 * rows sorted & unique, local column index == ghost row LID. */
int64_t go_graph_nnz(go_ctx*);
const int64_t* go_graph_rowptr(go_ctx*); /* [4*Nn+1] */
const int32_t* go_graph_colind(go_ctx*); /* [nnz]    */

void go_set_solution(go_ctx*, const double* u /*[Nn*3]*/, const double* p /*[Nn]*/);

/* history state, AoS as apf stores it: 3x3 row-major per element
 * (src/goal_states.cpp:21-57).  names: sigma, eqps, eqps_old, Fp, Fp_old. */
double* go_state(go_ctx*, const char* name);
void go_update_states(go_ctx*); /* X_old <- X, src/goal_states.cpp:130-141 */

/* ST residual chain (src/goal_primal.cpp:43-51, goal_mechanics.cpp:97-146):
 * R += element residuals (R is NOT zeroed here; zero_R is the caller's step). */
int go_assemble_residual(go_ctx*, int save_state, double* R /*[4Nn]*/);
/* FADT chain; mode PRIMAL or ADJOINT (transposed scatter). values may be NULL. */
int go_assemble_jacobian(go_ctx*, int mode, int save_state, double* R, double* values /*[nnz]*/);
/* avg-disp functional of the current solution (src/goal_avg_disp.cpp:17-21) and,
 * if dMdu != NULL, its FAD derivative scattered like QoI<FADT>::scatter. */
double go_functional_avg_disp(go_ctx*, double* dMdu /*[4Nn] or NULL*/);

/* The reference's functionals (Mechanics::build_functional, src/goal_mechanics.cpp:149-167) evaluated behind the
 * save=false residual chain like Functional::compute (src/goal_functional.cpp:62-70); with dMdu != NULL the FADT
 * chain and QoI<FADT>::scatter (src/goal_qoi.cpp:63-76; dMdu is NOT zeroed here).  elem_set: "elem set" index of
 * the subdomain / avg-vm functionals; rho: KS parameter of "max vm"; point_node/point_idx: "point wise". */
enum { GO_QOI_AVG_DISP = 0, GO_QOI_AVG_DISP_SUBDOMAIN = 1, GO_QOI_AVG_VM = 2, GO_QOI_KS_VM = 3, GO_QOI_POINT_WISE = 4 };
double go_functional(go_ctx*, int type, int elem_set, double rho, int point_node, int point_idx, double* dMdu);

/* error chain (src/goal_mechanics.cpp:169-218) with adjoint-weighted test
 * functions (goal_displacement_adjoint.cpp:37-53, goal_pressure_adjoint.cpp:38-49) */
int go_assemble_error(go_ctx*, const double* zu_diff /*[Nn*3]*/, const double* zp_diff /*[Nn]*/,
                      const double* zp_coarse /*[Nn]*/, double* R);
/* BForce (src/goal_bforce.cpp:58-68): R_u[n][i] -= b_i w_n^i w dv with b[Ne*3] given per element; zu_diff != NULL
 * selects the adjoint-weighted test functions of the error chain (src/goal_mechanics.cpp:204-208). */
int go_apply_bforce(go_ctx*, const double* b /*[Ne*3]*/, const double* zu_diff /*[Nn*3] or NULL*/, double* R);
/* src/goal_error.cpp:7-56 and goal_nested.cpp:395-412.  u_err[Nn*3], p_err[Nn];
 * eta_elem[Ne]; parent[Ne] -> eta_parent[n_parent] (zeroed here); returns bound. */
double go_element_error(go_ctx*, const double* u_err, const double* p_err, double* eta_elem,
                        const int32_t* parent, int n_parent, double* eta_parent);

/* get_iso_target_size (src/goal_size_field.cpp:39-150), one part: vertex sizes [Nn]; returns sum_contributions */
double go_size_field(go_ctx*, const double* eta /*[Ne]*/, int target, int p_order, double* vtx_size);

/* number of elements that took the plastic branch in the last assemble call */
int64_t go_last_plastic_count(go_ctx*);
const char* go_last_error(go_ctx*);

/* reference-style per-element work only (no scatter), for CPU timing:
 * evaluates elements [e0,e1) with the FADT chain and returns a checksum. */
double go_time_jacobian_elements(go_ctx*, int64_t e0, int64_t e1, int save_state);

#ifdef __cplusplus
}
#endif
#endif
