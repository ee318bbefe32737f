/* goal_b200.h -- C-ABI of the B200 assembly path for bgranzow/goal.
 *
 * The reference has no FFI: its assembly hot path is the C++ plug-in loop
 *   goal::assemble(Evaluators const&, SolInfo*)          src/goal_assembly.hpp:17
 * driven by
 *   Primal::compute_resid / compute_jacob                src/goal_primal.cpp:75-109
 *   NestedAdjoint::compute_adjoint / localize            src/goal_nested_adjoint.cpp:163-234
 * over Disc's maps/graphs (src/goal_disc.hpp:27-93), SolInfo's owned/ghost
 * R/dRdu (src/goal_sol_info.hpp:12-39) and the States history fields
 * (src/goal_states.hpp:13-23).  Each entry point below replaces one of those
 * call sites; the replaced reference interface is cited per function and the
 * binding a Goal maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C, no C++/torch types; every function returns gx_status (0 = ok) and
 *    leaves a message for gx_last_error().  The reference aborts through
 *    goal::fail() (src/goal_control.cpp:91-99); a shim maps non-zero to fail().
 *  - LO = int32_t local index, GO = int64_t global index (src/goal_data_types.hpp:13-14).
 *  - dof = node*4 + eq, eq 0..2 = u, eq 3 = p (src/goal_disc.cpp:195-201).
 *  - "ghost" numbering = all nodes present on the part (apf::numberOverlapNodes,
 *    src/goal_disc.cpp:294); "owned" = the subset this part owns (:273).
 *  - host pointers unless a name ends in _dev.  Host buffers should be pinned
 *    (cudaHostAlloc/cudaHostRegister) for full PCIe rate; pageable memory works.
 *  - one gx_ctx per GPU / MPI rank; calls on one ctx are serialised by the caller;
 *    every entry point is synchronous on return.
 *  - there is NO CPU fallback: without a CUDA device gx_create fails.
 */
#ifndef GOAL_B200_H
#define GOAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gx_ctx gx_ctx;

typedef enum {
  GX_OK = 0,
  GX_ERR_ARG = 1,                   /* bad argument / call order */
  GX_ERR_CUDA = 2,                  /* CUDA runtime failure (message has the cudaError) */
  GX_ERR_INVERTED_ELEMENT = 3,      /* dv <= 0 (apf::getDV), element id in message */
  GX_ERR_INVERTED_DEFORMATION = 4,  /* det F <= 0, element id in message */
  GX_ERR_J2_RETURN_MAP = 5,         /* "J2: return mapping failed" src/goal_J2.cpp:119-120 */
  GX_ERR_NCCL = 6,
  GX_ERR_UNSUPPORTED = 7
} gx_status;

/* constitutive models named by north_star: src/goal_neohookean.cpp, src/goal_J2.cpp */
enum { GX_MODEL_NEOHOOKEAN = 0, GX_MODEL_J2 = 1 };
/* scatter modes, src/goal_eval_modes.hpp:6 */
enum { GX_MODE_NONE = 0, GX_MODE_PRIMAL = 1, GX_MODE_ADJOINT = 2 };
/* gx_desc.flags */
enum { GX_FLAG_NO_STABILIZATION = 1 /* mechanics: stabilization: false, src/goal_mechanics.cpp:55-56 */ };

#define GX_MAX_ELEM_SETS 16

/* What Disc + Mechanics hold for one mesh part (src/goal_disc.cpp:224-235,
 * src/goal_mechanics.cpp:41-63).  All arrays are copied; the caller may free them. */
typedef struct {
  int32_t n_nodes;          /* ghost (overlap) node count on this part */
  int32_t n_elems;          /* elements of this part (no ghost element layers) */
  const int32_t* conn;      /* [n_elems*4] ghost-local node ids, det(x1-x0,x2-x0,x3-x0) > 0 */
  const double* coords;     /* [n_nodes*3] */
  const int32_t* elem_set;  /* [n_elems] elem-set index per element, or NULL (one set) */
  int32_t n_elem_sets;      /* <= GX_MAX_ELEM_SETS */
  int32_t model;            /* GX_MODEL_* (mechanics: model) */
  const double* materials;  /* [n_elem_sets*5] E, nu, K, Y, c0 per set (src/goal_J2.cpp:12-24) */
  int32_t device;           /* CUDA device ordinal; -1 = host-only context (graph / exchange plan only, cannot compute) */
  uint32_t flags;
  /* ---- partition (all zero / NULL for a single part) ------------------------
   * PUMI remotes as the .smb stores them: for each neighbouring part the list of
   * local vertices shared with it, in an order agreed by both sides.            */
  int32_t rank, n_ranks;
  const int64_t* node_gid;      /* [n_nodes] global node id (apf::makeGlobal numbering) or NULL */
  const int32_t* node_owner;    /* [n_nodes] owning rank of each node, or NULL (= all owned) */
  int32_t n_peers;
  const int32_t* peer_rank;     /* [n_peers] */
  const int32_t* peer_offset;   /* [n_peers+1] offsets into peer_nodes */
  const int32_t* peer_nodes;    /* concatenated shared-vertex lists (ghost-local ids) */
} gx_desc;

/* Disc::build_data + Primal::build_data (src/goal_disc.cpp:224-235, goal_primal.cpp:57-60):
 * builds the ghost CRS graph, the element->nonzero scatter map and the conflict-free
 * element schedule, allocates device mirrors of R / dRdu / states. */
int gx_create(const gx_desc* desc, gx_ctx** out);
int gx_destroy(gx_ctx* ctx);
const char* gx_last_error(const gx_ctx* ctx); /* never NULL; ctx may be NULL for gx_create failures */

/* Ghost CRS graph in Tpetra's local layout after fillComplete (Disc::compute_graphs,
 * src/goal_disc.cpp:307-332): rows sorted & unique, local column index == ghost row LID.
 * Pointers stay valid until gx_destroy. nnz may exceed 2^31 -> 64-bit row offsets. */
int gx_graph(gx_ctx* ctx, int64_t* nnz, const int64_t** rowptr /*[4*n_nodes+1]*/, const int32_t** colind /*[nnz]*/);
/* Same graph without materialising colind (nnz and rowptr only). */
int gx_graph_size(gx_ctx* ctx, int64_t* nnz, int32_t* n_rows);
/* The node-level form the library stores: block row offsets [n_nodes+1] and the neighbour node of every 4x4 block,
 * sorted per row.  dof row 4a+i starts at 16*nrow[a] + i*4*(nrow[a+1]-nrow[a]); its columns are 4*ncol[..]+k. */
int gx_node_graph(gx_ctx* ctx, const int64_t** nrow, const int32_t** ncol);

/* Element -> nonzero scatter map (what sumIntoLocalValues searches for on every call,
 * src/goal_displacement.cpp:191): for element e, local nodes (n,m), the position of the
 * 4x4 node block inside node a_n's block row; value index of entry ((n,i),(m,k)) is
 *   rowptr[4*a_n+i] + 4*bpos + k.  out: [n_elems*16] uint8, user element order. */
int gx_scatter_map(gx_ctx* ctx, uint8_t* bpos);

/* Solution fields "u","p" (Disc::add_soln writes them, src/goal_disc.cpp:398-422;
 * Displacement/Pressure::gather read them, src/goal_displacement.cpp:139-156). */
int gx_set_solution(gx_ctx* ctx, const double* u /*[n_nodes*3]*/, const double* p /*[n_nodes]*/);

/* Disc::add_soln (src/goal_disc.cpp:398-422): u += du[4a+d], p += du[4a+3] on the device-resident solution, du in
 * ghost dof layout [4*n_nodes].  The reference updates owned nodes and then apf::synchronize()s the copies
 * (:420-421): on a partitioned context follow with gx_sync_solution (NCCL) or the gx_pack_solution /
 * gx_unpack_solution halves (own transport), after which non-owned entries of du no longer matter. */
int gx_add_solution(gx_ctx* ctx, const double* du /*[4*n_nodes]*/);
int gx_get_solution(gx_ctx* ctx, double* u /*[n_nodes*3]*/, double* p /*[n_nodes]*/);
int gx_sync_solution(gx_ctx* ctx);
int gx_pack_solution(gx_ctx* ctx, int peer_index, void** send_dev, int64_t* send_bytes); /* owner side */
int gx_unpack_solution(gx_ctx* ctx, int peer_index, const void* recv_dev);               /* copy side: overwrite */

/* History state, AoS as apf stores it (src/goal_states.cpp:21-57): names "sigma",
 * "eqps", "eqps_old", "Fp", "Fp_old"; tensors are 9 doubles row-major per element. */
int gx_get_state(gx_ctx* ctx, const char* name, double* out);
int gx_set_state(gx_ctx* ctx, const char* name, const double* in);
int gx_update_states(gx_ctx* ctx); /* States::update: X_old <- X  (src/goal_states.cpp:130-141) */

/* Primal::compute_resid minus BCs (src/goal_primal.cpp:81-83): zero_R + assemble(residual).
 * R_out (ghost layout, [4*n_nodes]) may be NULL to leave the result on the device. */
int gx_compute_residual(gx_ctx* ctx, int save_state, double* R_out);
/* Primal::compute_jacob / NestedAdjoint::compute_adjoint minus BCs
 * (src/goal_primal.cpp:98-101, goal_nested_adjoint.cpp:169-172): zero_all + assemble(jacobian).
 * mode PRIMAL scatters dRdu, ADJOINT its transpose (src/goal_displacement.cpp:196-214).
 * R_out / values_out ([nnz], ghost CRS order) may be NULL. */
int gx_compute_jacobian(gx_ctx* ctx, int mode, int save_state, double* R_out, double* values_out);
/* NestedAdjoint::localize minus BCs (src/goal_nested_adjoint.cpp:224-226): zero_R +
 * assemble(error chain of Mechanics::build_error, src/goal_mechanics.cpp:169-218). */
int gx_localize_error(gx_ctx* ctx, const double* zu_diff /*[n_nodes*3]*/, const double* zp_diff /*[n_nodes]*/,
                      const double* zp_coarse /*[n_nodes]*/, double* R_out);
/* compute_error + sum_contribs + Nested::set_error (src/goal_error.cpp:7-56,
 * goal_nested.cpp:395-412).  parent may be NULL.  *bound is this part's sum; across
 * parts use gx_allreduce_sum (== PCU_Add_Doubles, src/goal_error.cpp:54). */
int gx_element_error(gx_ctx* ctx, const double* u_err /*[n_nodes*3]*/, const double* p_err /*[n_nodes]*/,
                     const int32_t* parent /*[n_elems]*/, int32_t n_parent, double* eta_elem /*[n_elems]*/,
                     double* eta_parent /*[n_parent]*/, double* bound);

/* get_iso_target_size (src/goal_size_field.cpp:39-150; Nested adaptation input): from the element error
 * indicators, G = sum_e |eta_e|^(2d/(2p+d)), size_factor = (G/target)^(1/d), element size
 * clamp(size_factor |eta_e|^(-2/(2p+d)) h_e, h_e/4, 2 h_e) with h_e = sqrt(mean squared edge length), and the vertex
 * size = mean over the adjacent elements.  *G <= 0 on entry: this part's G is computed and returned (several parts:
 * PCU_Add it, then call again with the sum); vtx_size == NULL stops after G.  vtx_count != NULL returns per-vertex
 * sums and adjacent-element counts instead of means, so that parts can add both before dividing. */
int gx_size_field(gx_ctx* ctx, const double* eta_elem /*[n_elems]*/, int32_t target, int32_t p_order, double* G,
                  double* vtx_size /*[n_nodes] or NULL*/, double* vtx_count /*[n_nodes] or NULL*/);

/* ---- "next" rows of SURVEY.md 8(f): what sits either side of the assembly in the same drivers ---------------
 * Functional "avg disp" (Functional::compute src/goal_functional.cpp:62-70, AvgDisp src/goal_avg_disp.cpp:17-21):
 *   *J = sum over this part's elements of (sum_i u_i(xi_c)) w dv / 3   (PCU_Add across parts: gx_allreduce_sum)
 * dMdu_out ([4*n_nodes], ghost layout, may be NULL) receives what QoI<FADT>::scatter adds (src/goal_qoi.cpp:63-76). */
int gx_functional_avg_disp(gx_ctx* ctx, double* J, double* dMdu_out);
/* Every functional of Mechanics::build_functional (src/goal_mechanics.cpp:149-167), evaluated behind the save=false
 * residual chain like Functional::compute (src/goal_functional.cpp:41-45, 62-70):
 *   GX_QOI_AVG_DISP            "avg disp"            src/goal_avg_disp.cpp:17-21
 *   GX_QOI_AVG_DISP_SUBDOMAIN  "avg disp subdomain"  src/goal_avg_disp_subdomain.cpp:37-53   (elem_set)
 *   GX_QOI_AVG_VM              "avg vm"              src/goal_avg_vm.cpp:43-61, goal_von_mises.cpp:6-18 (elem_set)
 *   GX_QOI_KS_VM               "max vm"              src/goal_ks_vm.cpp:36-105 (rho; max/scale from the saved sigma)
 *   GX_QOI_POINT_WISE          "point wise"          src/goal_point_wise.cpp:37-56 (point_node ghost-local or -1, point_idx)
 * *J = this part's value (PCU_Add across parts: gx_allreduce_sum).  dMdu_out ([4*n_nodes] ghost layout, may be
 * NULL) receives what QoI<FADT>::scatter adds (src/goal_qoi.cpp:63-76); the derivative of the stress functionals
 * is closed-form (no FAD).  It also stays on the device: gx_dmdu_dev, and gx_reduce_interfaces(ctx, 4) ==
 * SolInfo::gather_dMdu (src/goal_sol_info.cpp:37-39).
 * "max vm": with ks_scale <= 0 on entry the part-local max / scale are computed here and returned in *q (single
 * part).  With several parts reduce gx_ks_vm_max (PCU_Max) and gx_ks_vm_scale (PCU_Add) on the host, pass them in
 * q, and take J = ks_max + log(ks_scale)/rho (KSVM::post_process). */
enum { GX_QOI_AVG_DISP = 0, GX_QOI_AVG_DISP_SUBDOMAIN = 1, GX_QOI_AVG_VM = 2, GX_QOI_KS_VM = 3, GX_QOI_POINT_WISE = 4 };
typedef struct {
  int32_t type;        /* GX_QOI_* */
  int32_t elem_set;    /* "elem set" index (subdomain / avg vm) */
  double rho;          /* "max vm" */
  double ks_max, ks_scale; /* "max vm": in (ks_scale > 0) or out */
  int32_t point_node, point_idx; /* "point wise" */
} gx_qoi;
int gx_functional(gx_ctx* ctx, gx_qoi* q, double* J, double* dMdu_out);
int gx_ks_vm_max(gx_ctx* ctx, double* max_vm);
int gx_ks_vm_scale(gx_ctx* ctx, double rho, double max_vm, double* scale);
int gx_dmdu_dev(gx_ctx* ctx, double** dMdu_dev);
int gx_fetch_dmdu(gx_ctx* ctx, double* dMdu_out);
/* Dirichlet rows on the device-resident result of the last compute call (set_resid_dbcs / set_jac_dbcs,
 * src/goal_dbcs.cpp:39-99): for each listed ghost-local dof row (which must be owned by this rank):
 * R[row] = solution - g; with_jacobian != 0 additionally zeroes the CRS row, puts 1 on the diagonal and clears
 * dMdu[row] of the last gx_functional (src/goal_dbcs.cpp:86), if there is one on the device.
 * As in the reference this runs after the interface reduction and does not eliminate columns. */
int gx_apply_dbcs(gx_ctx* ctx, int32_t n, const int32_t* rows, const double* g, int with_jacobian);
/* Traction and inward-traction boundary terms on the device-resident ghost R of the last compute call, before the
 * interface reduction (set_tbcs src/goal_tbcs.cpp:29-71, set_ibcs src/goal_ibcs.cpp:41-83; call order of
 * Primal::compute_resid/compute_jacob, src/goal_primal.cpp:84-85, 102-103): for every triangle of the side set
 * R[row(n,d)] -= T_d N_n(xi_c) w dv.  side_nodes: [n_sides*3] ghost-local vertex ids.  traction: [n_sides*3], the
 * side's expression evaluated by the host at the triangle centroid and the current time (goal::eval); the
 * inward form computes T = scale (x_c - center) itself.  Sides sharing a node are added in ascending side order
 * (the reference's loop order), without atomics. */
int gx_apply_tbcs(gx_ctx* ctx, int32_t n_sides, const int32_t* side_nodes, const double* traction);
int gx_apply_ibcs(gx_ctx* ctx, int32_t n_sides, const int32_t* side_nodes, double scale, const double center[3]);
/* BForce<T>::at_point (src/goal_bforce.cpp:58-68; wired behind MResidual, goal_mechanics.cpp:132-136 / 204-208):
 * R_u[n][i] -= b_i w_n^i w dv on the device-resident ghost R of the last pass.  b: [n_elems * 3], the body force at
 * each element's integration point (the reference evaluates a named expression there; the value is the caller's).
 * error_weights = 0: w = N_n (residual / Jacobian pass; b does not depend on u, the Jacobian is unchanged);
 * error_weights = 1: w_n^i = z_i N_n with the u_z_diff of the last gx_localize_error (error chain). */
int gx_apply_bforce(gx_ctx* ctx, const double* b, int error_weights);

/* ---- mesh parts ------------------------------------------------------------------------------
 * Structure exchange == the owned_graph Export/INSERT of Disc::compute_graphs (src/goal_disc.cpp:327-329):
 * once after gx_create on a partitioned context, every rank sends peer p the blob of gx_struct_pack(p),
 * feeds what it received to gx_struct_unpack(p), then calls gx_struct_finalize.  Any transport works
 * (MPI in the reference, torch.distributed in the tests); gx_comm_init does it over NCCL. */
int gx_num_peers(gx_ctx* ctx, int32_t* n);
int gx_struct_pack(gx_ctx* ctx, int peer_index, const void** blob, int64_t* bytes);
int gx_struct_unpack(gx_ctx* ctx, int peer_index, const void* blob, int64_t bytes);
int gx_struct_finalize(gx_ctx* ctx);
/* Owned view (Disc owned_map / owned_graph, src/goal_disc.cpp:270-290, 327-329): the nodes this rank owns
 * in ascending local id, dof-level row offsets, and GLOBAL column dof ids (4*global_node + eq) in stored
 * order: the ghost row's columns (sorted by local id) followed by columns that exist only on other parts
 * (sorted by global id). */
int gx_owned_graph(gx_ctx* ctx, int32_t* n_owned_nodes, const int32_t** owned_nodes, int64_t* nnz_owned,
                   const int64_t** rowptr /*[4*n_owned+1]*/, const int64_t** col_gid /*[nnz_owned]*/);
int gx_fetch_owned(gx_ctx* ctx, double* R_owned /*[4*n_owned]*/, double* values_owned /*[nnz_owned]*/);
/* The owned matrix in Tpetra's LOCAL layout -- what sol_info->owned->dRdu holds after owned_graph->fillComplete()
 * (src/goal_disc.cpp:327-332), so gx_fetch_owned_tpetra can fill the matrix goal::solve consumes
 * (src/goal_sol_info.cpp:41-43, src/goal_linear_solve.cpp:72-90) value for value.  Rule restated from Tpetra
 * (fillComplete -> makeColMap; Trilinos version unpinned by the reference): rows = owned dofs in owned_map order (owned
 * nodes in ascending overlap-local id, dof = node * 4 + eq); column map = the owned dofs in that same order, then the
 * remote dofs grouped by owning rank (ascending) and by global id inside a rank; every row's local column indices
 * ascending.  colmap_node_gid[n_col_nodes] lists the column map at node level (dof = 4 * position + eq);
 * rowptr[4 * n_owned + 1], colind[nnz_owned] are the local CRS arrays.  A single-part context gives the ghost
 * layout of gx_graph. */
int gx_owned_tpetra_graph(gx_ctx* ctx, int32_t* n_owned_nodes, int32_t* n_col_nodes, const int64_t** colmap_node_gid,
                          int64_t* nnz_owned, const int64_t** rowptr, const int32_t** colind);
int gx_fetch_owned_tpetra(gx_ctx* ctx, double* R_owned /*[4*n_owned]*/, double* values_owned /*[nnz_owned]*/);
/* The exchange plan of one peer (for hosts that verify or emulate the exchange). counts = {n_send, n_recv};
 * recv_cnt[s] = sender's block count of receive node s, recv_map = concatenated block maps. */
int gx_exchange_plan(gx_ctx* ctx, int peer_index, int32_t* peer_rank, int32_t counts[2], const int32_t** send_nodes,
                     const int32_t** recv_nodes, const int32_t** recv_cnt, const int32_t** recv_map);

/* SolInfo::gather_R / gather_dRdu (Tpetra Export ghost->owned, ADD; src/goal_sol_info.cpp:33-43)
 * over NCCL.  After it, rows of nodes this part owns hold the sum over all parts, rows of
 * nodes owned elsewhere are left as local partial sums.  what: bit 0 = R, bit 1 = dRdu; or what = 4 alone:
 * dMdu of the last gx_functional (gather_dMdu, src/goal_sol_info.cpp:37-39). */
int gx_comm_init(gx_ctx* ctx, const void* nccl_unique_id, size_t id_bytes);
int gx_nccl_unique_id(void* out, size_t* id_bytes); /* helper: rank 0 creates, host code broadcasts */
int gx_reduce_interfaces(gx_ctx* ctx, int what);
int gx_allreduce_sum(gx_ctx* ctx, double* x, int n);
/* Pack / unpack halves of the exchange for hosts that bring their own transport
 * (MPI, torch.distributed): interface rows of peer `p` as a flat device buffer.
 * `what` is 1 (R: 4 doubles per send node) or 2 (dRdu: the node's block rows). */
int gx_interface_bytes(gx_ctx* ctx, int peer_index, int what, int64_t* send_bytes, int64_t* recv_bytes);
int gx_pack_interface(gx_ctx* ctx, int peer_index, int what, void** send_dev);
int gx_unpack_add_interface(gx_ctx* ctx, int peer_index, int what, const void* recv_dev);

/* Results of the last compute call, device resident (for a GPU-resident consumer). */
int gx_result_dev(gx_ctx* ctx, double** R_dev, double** values_dev);
int gx_fetch(gx_ctx* ctx, double* R_out, double* values_out); /* D2H copy of the above */

/* Introspection */
int gx_plastic_count(gx_ctx* ctx, int64_t* n);  /* elements on the plastic branch in the last call */
int gx_num_colors(gx_ctx* ctx, int32_t* n);
void* gx_stream(gx_ctx* ctx);                   /* the cudaStream_t every kernel of ctx runs on */
/* Device time (ms, CUDA events on gx_stream) of the last compute call:
 * t[0] zeroing, t[1] assembly kernels, t[2] interface exchange, t[3] kernel launches counted */
int gx_last_timing(gx_ctx* ctx, double t[4]);
/* t[1] of gx_last_timing split at the boundary between the element kernel (t[0]: stress update, state save, element
 * records) and the kernels that gather the records into R / the CRS values (t[1]); {t[1], 0} for a one-kernel pass */
int gx_last_stage_timing(gx_ctx* ctx, double t[2]);
/* Tuning / cross-check switches (defaults in brackets; everything else is GX_ERR_ARG):
 *   "kernel"                 [0] owner-computes schedules; 1 = coloured element schedule for every pass (fallback, cross-check)
 *   "residual_kernel"        [0] block-reduced residual / localisation passes; 1 = element lines + node gather
 *   "overlap"                [0] 1 = R, 2 = dRdu, 3 = both: the Jacobian pass reduces the interfaces itself, hidden
 *                            behind the interior patches (== SolInfo::gather_R / gather_dRdu, goal_sol_info.cpp:33-43)
 *   "prefetch"               [256] stage B: L2 prefetch distance in patches (0 = off)
 *   "prefetch_elems"         [18944] element kernels of passes that save the state: L2 prefetch distance in elements
 *   "prefetch_elems_nosave"  [151552] the same for passes that do not save
 *   "block_size"             [128] threads per block of the coloured schedule
 *   "patch_schedule_dryrun"  builds the patch schedule on the host only (works on host-only contexts) */
int gx_set_option(gx_ctx* ctx, const char* key, int64_t value);
/* The work list of the patch-gather Jacobian pass as the device reads it (layout: goal_b200/csrc/gx_setup.cpp,
 * build_patch_schedule); dims = {patches, words per patch, record slots per patch, threads per patch}.  Valid until
 * the next call on ctx.  Host-only contexts can build it too: the CPU tests check its invariants. */
int gx_patch_schedule(gx_ctx* ctx, const uint32_t** words, int32_t dims[4]);
/* Measurement aid ("the FP64 peak must be measured on the box in the same run"): a register-only
 * DFMA kernel, 8 independent chains per thread, every SM full, timed with CUDA events on gx_stream.
 * *tflops = 2 * fused multiply-adds / time; *sm_mhz = SM clock seen by the kernel (clock64 / elapsed). */
int gx_measure_fp64_peak(gx_ctx* ctx, double* tflops, double* sm_mhz);

#ifdef __cplusplus
}
#endif
#endif
