"""ctypes binding of the C-ABI (include/goal_b200.h) -> libgoal_b200.so.

`Assembler` mirrors, call for call, what the reference's Primal / NestedAdjoint do
around goal::assemble (src/goal_primal.cpp:75-109, src/goal_nested_adjoint.cpp:163-234):
set the solution fields, compute_resid / compute_jacob, localize, states->update().
There is no CPU path: if the CUDA library is missing or no GPU is visible the
constructor raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GOAL_B200_LIB: developer switch for A/B timing of two builds of the same library (never a different backend)
LIB_PATH = os.environ.get("GOAL_B200_LIB") or os.path.join(_HERE, "libgoal_b200.so")

MODEL = {"neohookean": 0, "J2": 1}
NONE, PRIMAL, ADJOINT = 0, 1, 2

STATUS = {0: "GX_OK", 1: "GX_ERR_ARG", 2: "GX_ERR_CUDA", 3: "GX_ERR_INVERTED_ELEMENT",
          4: "GX_ERR_INVERTED_DEFORMATION", 5: "GX_ERR_J2_RETURN_MAP", 6: "GX_ERR_NCCL", 7: "GX_ERR_UNSUPPORTED"}

# every symbol include/goal_b200.h declares (tests/test_abi.py checks the .so exports them all)
SYMBOLS = [
    "gx_create", "gx_destroy", "gx_last_error", "gx_graph", "gx_graph_size", "gx_scatter_map",
    "gx_set_solution", "gx_get_state", "gx_set_state", "gx_update_states", "gx_compute_residual",
    "gx_compute_jacobian", "gx_localize_error", "gx_element_error", "gx_comm_init", "gx_nccl_unique_id",
    "gx_reduce_interfaces", "gx_allreduce_sum", "gx_interface_bytes", "gx_pack_interface",
    "gx_unpack_add_interface", "gx_result_dev", "gx_fetch", "gx_plastic_count", "gx_num_colors",
    "gx_stream", "gx_last_timing", "gx_last_stage_timing", "gx_set_option", "gx_measure_fp64_peak", "gx_apply_bforce", "gx_owned_tpetra_graph", "gx_fetch_owned_tpetra", "gx_num_peers", "gx_struct_pack", "gx_struct_unpack",
    "gx_struct_finalize", "gx_owned_graph", "gx_fetch_owned", "gx_exchange_plan", "gx_functional_avg_disp",
    "gx_apply_dbcs", "gx_node_graph", "gx_functional", "gx_ks_vm_max", "gx_ks_vm_scale", "gx_dmdu_dev",
    "gx_fetch_dmdu", "gx_apply_tbcs", "gx_apply_ibcs", "gx_add_solution", "gx_get_solution", "gx_sync_solution",
    "gx_pack_solution", "gx_unpack_solution", "gx_size_field", "gx_patch_schedule",
]

# Mechanics::build_functional types by their yaml name (src/goal_mechanics.cpp:149-167)
QOI = {"avg disp": 0, "avg disp subdomain": 1, "avg vm": 2, "max vm": 3, "point wise": 4}


class GxQoi(C.Structure):
    _fields_ = [("type", C.c_int32), ("elem_set", C.c_int32), ("rho", C.c_double), ("ks_max", C.c_double),
                ("ks_scale", C.c_double), ("point_node", C.c_int32), ("point_idx", C.c_int32)]


class GxError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"{STATUS.get(status, status)}: {msg}")
        self.status = status


class GxDesc(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32), ("n_elems", C.c_int32),
        ("conn", C.POINTER(C.c_int32)), ("coords", C.POINTER(C.c_double)),
        ("elem_set", C.POINTER(C.c_int32)), ("n_elem_sets", C.c_int32), ("model", C.c_int32),
        ("materials", C.POINTER(C.c_double)), ("device", C.c_int32), ("flags", C.c_uint32),
        ("rank", C.c_int32), ("n_ranks", C.c_int32),
        ("node_gid", C.POINTER(C.c_int64)), ("node_owner", C.POINTER(C.c_int32)),
        ("n_peers", C.c_int32), ("peer_rank", C.POINTER(C.c_int32)),
        ("peer_offset", C.POINTER(C.c_int32)), ("peer_nodes", C.POINTER(C.c_int32)),
    ]


_LIB = None


def load_library():
    """Load libgoal_b200.so; raises (never falls back) when it is not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(goal_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    dp, ip, lp, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_void_p
    L.gx_create.argtypes = [C.POINTER(GxDesc), C.POINTER(vp)]
    L.gx_destroy.argtypes = [vp]
    L.gx_last_error.restype = C.c_char_p
    L.gx_last_error.argtypes = [vp]
    L.gx_graph.argtypes = [vp, lp, C.POINTER(lp), C.POINTER(ip)]
    L.gx_graph_size.argtypes = [vp, lp, ip]
    L.gx_node_graph.argtypes = [vp, C.POINTER(lp), C.POINTER(ip)]
    L.gx_scatter_map.argtypes = [vp, C.POINTER(C.c_uint8)]
    L.gx_set_solution.argtypes = [vp, vp, vp]
    L.gx_get_state.argtypes = [vp, C.c_char_p, dp]
    L.gx_set_state.argtypes = [vp, C.c_char_p, dp]
    L.gx_update_states.argtypes = [vp]
    L.gx_compute_residual.argtypes = [vp, C.c_int, vp]
    L.gx_compute_jacobian.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.gx_localize_error.argtypes = [vp, vp, vp, vp, vp]
    L.gx_element_error.argtypes = [vp, dp, dp, ip, C.c_int32, dp, dp, dp]
    L.gx_comm_init.argtypes = [vp, vp, C.c_size_t]
    L.gx_nccl_unique_id.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.gx_reduce_interfaces.argtypes = [vp, C.c_int]
    L.gx_allreduce_sum.argtypes = [vp, dp, C.c_int]
    L.gx_interface_bytes.argtypes = [vp, C.c_int, C.c_int, lp, lp]
    L.gx_pack_interface.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.gx_unpack_add_interface.argtypes = [vp, C.c_int, C.c_int, vp]
    L.gx_result_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.gx_fetch.argtypes = [vp, vp, vp]
    L.gx_plastic_count.argtypes = [vp, lp]
    L.gx_num_colors.argtypes = [vp, ip]
    L.gx_stream.restype = vp
    L.gx_stream.argtypes = [vp]
    L.gx_last_timing.argtypes = [vp, dp]
    L.gx_last_stage_timing.argtypes = [vp, dp]
    L.gx_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.gx_measure_fp64_peak.argtypes = [vp, dp, dp]
    L.gx_apply_bforce.argtypes = [vp, dp, C.c_int]
    L.gx_owned_tpetra_graph.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(lp), lp, C.POINTER(lp), C.POINTER(ip)]
    L.gx_fetch_owned_tpetra.argtypes = [vp, C.c_void_p, C.c_void_p]
    L.gx_functional_avg_disp.argtypes = [vp, dp, vp]
    L.gx_apply_dbcs.argtypes = [vp, C.c_int32, ip, dp, C.c_int]
    L.gx_apply_tbcs.argtypes = [vp, C.c_int32, ip, dp]
    L.gx_apply_ibcs.argtypes = [vp, C.c_int32, ip, C.c_double, dp]
    L.gx_add_solution.argtypes = [vp, vp]
    L.gx_get_solution.argtypes = [vp, vp, vp]
    L.gx_sync_solution.argtypes = [vp]
    L.gx_pack_solution.argtypes = [vp, C.c_int, C.POINTER(vp), lp]
    L.gx_unpack_solution.argtypes = [vp, C.c_int, vp]
    L.gx_size_field.argtypes = [vp, dp, C.c_int32, C.c_int32, dp, dp, dp]
    L.gx_patch_schedule.argtypes = [vp, C.POINTER(C.POINTER(C.c_uint32)), ip]
    L.gx_functional.argtypes = [vp, C.POINTER(GxQoi), dp, vp]
    L.gx_ks_vm_max.argtypes = [vp, dp]
    L.gx_ks_vm_scale.argtypes = [vp, C.c_double, C.c_double, dp]
    L.gx_dmdu_dev.argtypes = [vp, C.POINTER(vp)]
    L.gx_fetch_dmdu.argtypes = [vp, vp]
    L.gx_num_peers.argtypes = [vp, ip]
    L.gx_struct_pack.argtypes = [vp, C.c_int, C.POINTER(vp), lp]
    L.gx_struct_unpack.argtypes = [vp, C.c_int, vp, C.c_int64]
    L.gx_struct_finalize.argtypes = [vp]
    L.gx_owned_graph.argtypes = [vp, ip, C.POINTER(ip), lp, C.POINTER(lp), C.POINTER(lp)]
    L.gx_fetch_owned.argtypes = [vp, vp, vp]
    L.gx_exchange_plan.argtypes = [vp, C.c_int, ip, ip, C.POINTER(ip), C.POINTER(ip), C.POINTER(ip), C.POINTER(ip)]
    _LIB = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _addr(a):
    """host address of a numpy array or a (pinned) torch tensor, or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    assert a.dtype.is_floating_point and a.element_size() == 8 and a.is_contiguous() and not a.is_cuda
    return C.c_void_p(a.data_ptr())


class Assembler:
    """One mesh part on one GPU (a gx_ctx)."""

    def __init__(self, coords, conn, model, materials, elem_set=None, device=0, partition=None, flags=0):
        self.L = load_library()
        self.coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, 3)
        self.conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, 4)
        self.nn, self.ne = len(self.coords), len(self.conn)
        self.model = model
        mats = np.ascontiguousarray(materials, dtype=np.float64).reshape(-1, 5)
        self._keep = [mats]
        d = GxDesc()
        d.flags = flags  # GX_FLAG_NO_STABILIZATION = 1
        d.n_nodes, d.n_elems = self.nn, self.ne
        d.conn, d.coords = _ip(self.conn), _dp(self.coords)
        if elem_set is not None:
            es = np.ascontiguousarray(elem_set, dtype=np.int32)
            self._keep.append(es)
            d.elem_set = _ip(es)
        d.n_elem_sets, d.model, d.materials, d.device = len(mats), MODEL[model], _dp(mats), device
        if partition is not None:
            d.rank, d.n_ranks = partition["rank"], partition["n_ranks"]
            gid = np.ascontiguousarray(partition["node_gid"], dtype=np.int64)
            own = np.ascontiguousarray(partition["node_owner"], dtype=np.int32)
            pr = np.ascontiguousarray(partition["peer_rank"], dtype=np.int32)
            po = np.ascontiguousarray(partition["peer_offset"], dtype=np.int32)
            pn = np.ascontiguousarray(partition["peer_nodes"], dtype=np.int32)
            self._keep += [gid, own, pr, po, pn]
            d.node_gid = gid.ctypes.data_as(C.POINTER(C.c_int64))
            d.node_owner, d.n_peers = _ip(own), len(pr)
            d.peer_rank, d.peer_offset, d.peer_nodes = _ip(pr), _ip(po), _ip(pn)
            self._peer_ranks = pr
        self.h = C.c_void_p()
        rc = self.L.gx_create(C.byref(d), C.byref(self.h))
        if rc:
            raise GxError(rc, self.L.gx_last_error(None).decode())
        nnz, nrows = C.c_int64(), C.c_int32()
        self._ck(self.L.gx_graph_size(self.h, C.byref(nnz), C.byref(nrows)))
        self.nnz = nnz.value
        self._rowptr = self._colind = None
        self._R = np.zeros(4 * self.nn)
        self._vals = None

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.gx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise GxError(rc, self.L.gx_last_error(self.h).decode())

    # ---- Disc::build_data products
    def _graph(self):
        if self._rowptr is None:
            nnz = C.c_int64()
            rp, ci = C.POINTER(C.c_int64)(), C.POINTER(C.c_int32)()
            self._ck(self.L.gx_graph(self.h, C.byref(nnz), C.byref(rp), C.byref(ci)))
            self._rowptr = np.ctypeslib.as_array(rp, (4 * self.nn + 1,)).copy()
            self._colind = np.ctypeslib.as_array(ci, (nnz.value,)).copy()

    @property
    def rowptr(self):
        self._graph()
        return self._rowptr

    @property
    def colind(self):
        self._graph()
        return self._colind

    def node_graph(self):
        """(nrow [Nn+1] int64, ncol [nblocks] int32): zero-copy views of the library's block-CRS graph."""
        nr, nc = C.POINTER(C.c_int64)(), C.POINTER(C.c_int32)()
        self._ck(self.L.gx_node_graph(self.h, C.byref(nr), C.byref(nc)))
        nrow = np.ctypeslib.as_array(nr, (self.nn + 1,))
        return nrow, np.ctypeslib.as_array(nc, (int(nrow[-1]),))

    def result_dev(self):
        """device addresses (ints) of R and of the CRS values of the last compute call."""
        r, v = C.c_void_p(), C.c_void_p()
        self._ck(self.L.gx_result_dev(self.h, C.byref(r), C.byref(v)))
        return r.value, v.value

    def scatter_map(self):
        b = np.zeros((self.ne, 16), dtype=np.uint8)
        self._ck(self.L.gx_scatter_map(self.h, b.ctypes.data_as(C.POINTER(C.c_uint8))))
        return b

    @property
    def num_colors(self):
        n = C.c_int32()
        self._ck(self.L.gx_num_colors(self.h, C.byref(n)))
        return n.value

    # ---- fields and state
    def set_solution(self, u, p):
        if isinstance(u, np.ndarray):
            u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
            p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1)
            assert u.size == 3 * self.nn and p.size == self.nn
        self._ck(self.L.gx_set_solution(self.h, _addr(u), _addr(p)))

    def get_state(self, name):
        out = np.zeros((self.ne, 9) if name in ("sigma", "Fp", "Fp_old") else (self.ne,))
        self._ck(self.L.gx_get_state(self.h, name.encode(), _dp(out)))
        return out

    def set_state(self, name, val):
        val = np.ascontiguousarray(val, dtype=np.float64)
        assert val.size == self.ne * (9 if name in ("sigma", "Fp", "Fp_old") else 1)
        self._ck(self.L.gx_set_state(self.h, name.encode(), _dp(val)))

    def update_states(self):
        self._ck(self.L.gx_update_states(self.h))

    # ---- the hot path
    def residual(self, save=True, out=True):
        """Primal::compute_resid minus BCs. out=False leaves R on the device."""
        self._ck(self.L.gx_compute_residual(self.h, int(save), _addr(self._R) if out is True else _addr(out) if out is not False else None))
        return self._R if out is True else out

    def jacobian(self, mode=PRIMAL, save=True, out=True, R_out=None, values_out=None):
        """Primal::compute_jacob / NestedAdjoint::compute_adjoint minus BCs -> (R, values)."""
        if out is True and R_out is None:
            if self._vals is None:
                self._vals = np.zeros(self.nnz)
            R_out, values_out = self._R, self._vals
        self._ck(self.L.gx_compute_jacobian(self.h, mode, int(save), _addr(R_out), _addr(values_out)))
        return R_out, values_out

    def localize(self, zu_diff, zp_diff, zp_coarse):
        a = [np.ascontiguousarray(x, dtype=np.float64).reshape(-1) for x in (zu_diff, zp_diff, zp_coarse)]
        self._ck(self.L.gx_localize_error(self.h, _addr(a[0]), _addr(a[1]), _addr(a[2]), _addr(self._R)))
        return self._R

    def element_error(self, u_err, p_err, parent=None, n_parent=0):
        u_err = np.ascontiguousarray(u_err, dtype=np.float64).reshape(-1)
        p_err = np.ascontiguousarray(p_err, dtype=np.float64).reshape(-1)
        eta, bound = np.zeros(self.ne), C.c_double()
        etap = None
        if parent is not None:
            parent = np.ascontiguousarray(parent, dtype=np.int32)
            etap = np.zeros(n_parent)
        self._ck(self.L.gx_element_error(self.h, _dp(u_err), _dp(p_err), None if parent is None else _ip(parent),
                                         n_parent, _dp(eta), None if etap is None else _dp(etap), C.byref(bound)))
        return eta, etap, bound.value

    def avg_disp(self, with_dMdu=False):
        """Functional "avg disp" of the current solution on the device (src/goal_avg_disp.cpp:17-21)."""
        J = C.c_double()
        d = np.zeros(4 * self.nn) if with_dMdu else None
        self._ck(self.L.gx_functional_avg_disp(self.h, C.byref(J), _addr(d)))
        return (J.value, d) if with_dMdu else J.value

    def functional(self, type, elem_set=0, rho=1.0, point=(0, 0), with_dMdu=False, ks=None):
        """Any functional of Mechanics::build_functional by its yaml `type` (src/goal_mechanics.cpp:149-167) on the
        device; with_dMdu also returns QoI<FADT>::scatter's dMdu (ghost layout).  ks = (max, scale) reduced over
        parts for "max vm" on a partitioned mesh."""
        q = GxQoi(QOI[type], elem_set, rho, 0.0 if ks is None else ks[0], 0.0 if ks is None else ks[1], point[0], point[1])
        J = C.c_double()
        d = np.zeros(4 * self.nn) if with_dMdu else None
        self._ck(self.L.gx_functional(self.h, C.byref(q), C.byref(J), _addr(d)))
        self.last_ks = (q.ks_max, q.ks_scale)
        return (J.value, d) if with_dMdu else J.value

    def ks_vm_max(self):
        v = C.c_double()
        self._ck(self.L.gx_ks_vm_max(self.h, C.byref(v)))
        return v.value

    def ks_vm_scale(self, rho, max_vm):
        v = C.c_double()
        self._ck(self.L.gx_ks_vm_scale(self.h, rho, max_vm, C.byref(v)))
        return v.value

    def fetch_dMdu(self):
        d = np.zeros(4 * self.nn)
        self._ck(self.L.gx_fetch_dmdu(self.h, _addr(d)))
        return d

    def apply_dbcs(self, rows, g, with_jacobian):
        """set_resid_dbcs / set_jac_dbcs on the device-resident result (src/goal_dbcs.cpp:39-99)."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        g = np.ascontiguousarray(g, dtype=np.float64)
        self._ck(self.L.gx_apply_dbcs(self.h, len(rows), _ip(rows), _dp(g), int(with_jacobian)))

    def add_solution(self, du):
        """Disc::add_soln (src/goal_disc.cpp:398-422): du in ghost dof layout [4 nn]."""
        du = np.ascontiguousarray(du, dtype=np.float64).reshape(-1)
        assert du.size == 4 * self.nn
        self._ck(self.L.gx_add_solution(self.h, _addr(du)))

    def get_solution(self):
        u, p = np.zeros((self.nn, 3)), np.zeros(self.nn)
        self._ck(self.L.gx_get_solution(self.h, _addr(u), _addr(p)))
        return u, p

    def sync_solution(self):
        self._ck(self.L.gx_sync_solution(self.h))

    def size_field(self, eta, target, p_order=1, G=0.0, counts=False):
        """get_iso_target_size (src/goal_size_field.cpp:39-150) -> (vertex sizes [nn], G[, counts])."""
        eta = np.ascontiguousarray(eta, dtype=np.float64)
        g = C.c_double(G)
        v = np.zeros(self.nn)
        c = np.zeros(self.nn) if counts else None
        self._ck(self.L.gx_size_field(self.h, _dp(eta), target, p_order, C.byref(g), _dp(v), None if c is None else _dp(c)))
        return (v, g.value, c) if counts else (v, g.value)

    def patch_schedule(self):
        """(words [n_patches, words_per_patch] uint32, record slots per patch, threads per patch) of the patch gather."""
        ptr, dims = C.POINTER(C.c_uint32)(), (C.c_int32 * 4)()
        self._ck(self.L.gx_patch_schedule(self.h, C.byref(ptr), dims))
        w = np.ctypeslib.as_array(ptr, (dims[0] * dims[1],)).reshape(dims[0], dims[1]).copy()
        return w, dims[2], dims[3]

    def apply_tbcs(self, sides, traction):
        """set_tbcs on the device-resident ghost R (src/goal_tbcs.cpp:29-71); traction: [n_sides, 3] or one 3-vector."""
        sides = np.ascontiguousarray(sides, dtype=np.int32).reshape(-1, 3)
        T = np.ascontiguousarray(np.broadcast_to(np.asarray(traction, dtype=np.float64), (len(sides), 3)))
        self._ck(self.L.gx_apply_tbcs(self.h, len(sides), _ip(sides), _dp(T)))

    def apply_bforce(self, b, error_weights=False):
        """BForce on the device-resident ghost R (src/goal_bforce.cpp:58-68); b: [n_elems, 3] at the element centroids."""
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
        assert b.size == 3 * self.ne
        self._ck(self.L.gx_apply_bforce(self.h, _dp(b), int(bool(error_weights))))

    def apply_ibcs(self, sides, scale, center):
        """set_ibcs (src/goal_ibcs.cpp:41-83): inward traction T = scale (x_c - center)."""
        sides = np.ascontiguousarray(sides, dtype=np.int32).reshape(-1, 3)
        c = np.ascontiguousarray(center, dtype=np.float64)
        self._ck(self.L.gx_apply_ibcs(self.h, len(sides), _ip(sides), float(scale), _dp(c)))

    def fetch(self, R=True, values=True):
        Rv = np.zeros(4 * self.nn) if R else None
        Vv = np.zeros(self.nnz) if values else None
        self._ck(self.L.gx_fetch(self.h, _addr(Rv), _addr(Vv)))
        return Rv, Vv

    # ---- introspection
    def plastic_count(self):
        n = C.c_int64()
        self._ck(self.L.gx_plastic_count(self.h, C.byref(n)))
        return n.value

    def last_timing(self):
        t = (C.c_double * 4)()
        self._ck(self.L.gx_last_timing(self.h, t))
        return dict(zero_ms=t[0], assemble_ms=t[1], exchange_ms=t[2], launches=int(t[3]))

    def last_stage_timing(self):
        """assemble_ms of last_timing() split into the element kernel and the gather kernel(s)"""
        t = (C.c_double * 2)()
        self._ck(self.L.gx_last_stage_timing(self.h, t))
        return dict(element_ms=t[0], gather_ms=t[1])

    def stream(self):
        return self.L.gx_stream(self.h)

    def measure_fp64_peak(self):
        """(TFLOP/s of a register-only DFMA kernel on this device, SM MHz it ran at): the measured FP64 roof."""
        t, f = C.c_double(0), C.c_double(0)
        self._ck(self.L.gx_measure_fp64_peak(self.h, C.byref(t), C.byref(f)))
        return t.value, f.value

    def set_option(self, key, value):
        self._ck(self.L.gx_set_option(self.h, key.encode(), int(value)))

    def csr(self, values):
        import scipy.sparse as sp
        return sp.csr_matrix((values, self.colind, self.rowptr), shape=(4 * self.nn, 4 * self.nn))

    # ---- mesh parts (SolInfo::gather_*, Disc owned graph)
    @property
    def num_peers(self):
        n = C.c_int32()
        self._ck(self.L.gx_num_peers(self.h, C.byref(n)))
        return n.value

    def struct_pack(self, peer):
        blob, nb = C.c_void_p(), C.c_int64()
        self._ck(self.L.gx_struct_pack(self.h, peer, C.byref(blob), C.byref(nb)))
        return C.string_at(blob, nb.value) if nb.value else b""

    def struct_unpack(self, peer, data):
        buf = C.create_string_buffer(data, len(data)) if data else None
        self._ck(self.L.gx_struct_unpack(self.h, peer, buf, len(data)))

    def struct_finalize(self):
        self._ck(self.L.gx_struct_finalize(self.h))
        nnz, nrows = C.c_int64(), C.c_int32()
        self._ck(self.L.gx_graph_size(self.h, C.byref(nnz), C.byref(nrows)))

    def exchange_structure(self, dist):
        """structure exchange over torch.distributed point-to-point (any backend)."""
        import torch
        ops, recv = [], {}
        peers = [self.exchange_plan_rank(p) for p in range(self.num_peers)]
        out = [torch.frombuffer(bytearray(self.struct_pack(p)) or bytearray(8), dtype=torch.int64) for p in range(self.num_peers)]
        lens_out = [torch.tensor([len(self.struct_pack(p)) // 8]) for p in range(self.num_peers)]
        lens_in = [torch.zeros(1, dtype=torch.int64) for _ in peers]
        for p, q in enumerate(peers):
            ops += [dist.P2POp(dist.isend, lens_out[p], q), dist.P2POp(dist.irecv, lens_in[p], q)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        ops = []
        for p, q in enumerate(peers):
            recv[p] = torch.zeros(max(int(lens_in[p]), 1), dtype=torch.int64)
            if int(lens_out[p]):
                ops.append(dist.P2POp(dist.isend, out[p], q))
            if int(lens_in[p]):
                ops.append(dist.P2POp(dist.irecv, recv[p], q))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for p in range(self.num_peers):
            n = int(lens_in[p])
            self.struct_unpack(p, recv[p][:n].numpy().tobytes() if n else b"")
        self.struct_finalize()

    def exchange_plan_rank(self, peer):
        r = C.c_int32()
        cnt = (C.c_int32 * 2)()
        # before finalize only the rank is meaningful; use the description we were built with
        return int(self._peer_ranks[peer])

    def exchange_plan(self, peer):
        r, cnt = C.c_int32(), (C.c_int32 * 2)()
        sn, rn, rc, rm = (C.POINTER(C.c_int32)() for _ in range(4))
        self._ck(self.L.gx_exchange_plan(self.h, peer, C.byref(r), cnt, C.byref(sn), C.byref(rn), C.byref(rc), C.byref(rm)))
        ns, nr = cnt[0], cnt[1]
        arr = lambda p, n: np.ctypeslib.as_array(p, (n,)).copy() if n else np.zeros(0, np.int32)
        recv_cnt = arr(rc, nr)
        return dict(rank=r.value, send_nodes=arr(sn, ns), recv_nodes=arr(rn, nr), recv_cnt=recv_cnt,
                    recv_map=arr(rm, int(recv_cnt.sum())))

    def owned_graph(self):
        no, nnz = C.c_int32(), C.c_int64()
        on, rp, cg = C.POINTER(C.c_int32)(), C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)()
        self._ck(self.L.gx_owned_graph(self.h, C.byref(no), C.byref(on), C.byref(nnz), C.byref(rp), C.byref(cg)))
        return dict(nodes=np.ctypeslib.as_array(on, (no.value,)).copy(),
                    rowptr=np.ctypeslib.as_array(rp, (4 * no.value + 1,)).copy(),
                    col_gid=np.ctypeslib.as_array(cg, (nnz.value,)).copy())

    def owned_tpetra_graph(self):
        """The owned matrix in Tpetra's local layout: dict(n_owned, colmap [node gids in column-map order], rowptr, colind)."""
        no, nc, nnz = C.c_int32(), C.c_int32(), C.c_int64()
        cm, rp, ci = C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)(), C.POINTER(C.c_int32)()
        self._ck(self.L.gx_owned_tpetra_graph(self.h, C.byref(no), C.byref(nc), C.byref(cm), C.byref(nnz), C.byref(rp), C.byref(ci)))
        return dict(n_owned=no.value, colmap=np.ctypeslib.as_array(cm, (nc.value,)).copy(),
                    rowptr=np.ctypeslib.as_array(rp, (4 * no.value + 1,)).copy(), colind=np.ctypeslib.as_array(ci, (nnz.value,)).copy())

    def fetch_owned_tpetra(self, values=True):
        g = self.owned_tpetra_graph()
        R = np.zeros(4 * g["n_owned"])
        V = np.zeros(len(g["colind"])) if values else None
        self._ck(self.L.gx_fetch_owned_tpetra(self.h, _addr(R), _addr(V)))
        return R, V, g

    def fetch_owned(self, values=True):
        g = self.owned_graph()
        R = np.zeros(4 * len(g["nodes"]))
        V = np.zeros(len(g["col_gid"])) if values else None
        self._ck(self.L.gx_fetch_owned(self.h, _addr(R), _addr(V)))
        return R, V, g

    def comm_init_torch(self, dist):
        """NCCL communicator for this context; the unique id travels over torch.distributed."""
        uid = C.create_string_buffer(128)
        n = C.c_size_t(128)
        if dist.get_rank() == 0:
            rc = self.L.gx_nccl_unique_id(uid, C.byref(n))
            if rc:
                raise GxError(rc, "gx_nccl_unique_id failed (libnccl not loadable?)")
        box = [uid.raw]
        dist.broadcast_object_list(box, src=0)
        self._ck(self.L.gx_comm_init(self.h, C.create_string_buffer(box[0], 128), 128))

    def reduce_interfaces(self, what=3):
        self._ck(self.L.gx_reduce_interfaces(self.h, what))

    def allreduce_sum(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1).copy()
        self._ck(self.L.gx_allreduce_sum(self.h, _dp(x), len(x)))
        return x
