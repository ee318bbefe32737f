"""ctypes binding of the CPU oracle (oracle/libgoal_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by goal_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODEL = {"neohookean": 0, "J2": 1}
NONE, PRIMAL, ADJOINT = 0, 1, 2
QOI = {"avg disp": 0, "avg disp subdomain": 1, "avg vm": 2, "max vm": 3, "point wise": 4}


def build(force=False):
    so = os.path.join(_HERE, "libgoal_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("goal_oracle.cpp", "goal_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.go_create.restype = C.c_void_p
        L.go_create.argtypes = [C.c_int, C.c_int, ip, dp, ip, C.c_int, C.c_int, dp]
        L.go_destroy.argtypes = [C.c_void_p]
        L.go_graph_nnz.restype = C.c_int64
        L.go_graph_nnz.argtypes = [C.c_void_p]
        L.go_graph_rowptr.restype = lp
        L.go_graph_rowptr.argtypes = [C.c_void_p]
        L.go_graph_colind.restype = ip
        L.go_graph_colind.argtypes = [C.c_void_p]
        L.go_set_solution.argtypes = [C.c_void_p, dp, dp]
        L.go_state.restype = dp
        L.go_state.argtypes = [C.c_void_p, C.c_char_p]
        L.go_update_states.argtypes = [C.c_void_p]
        L.go_assemble_residual.argtypes = [C.c_void_p, C.c_int, dp]
        L.go_assemble_jacobian.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, dp]
        L.go_functional_avg_disp.restype = C.c_double
        L.go_functional_avg_disp.argtypes = [C.c_void_p, dp]
        L.go_assemble_error.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.go_size_field.restype = C.c_double
        L.go_size_field.argtypes = [C.c_void_p, dp, C.c_int, C.c_int, dp]
        L.go_functional.restype = C.c_double
        L.go_functional.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, dp]
        L.go_element_error.restype = C.c_double
        L.go_element_error.argtypes = [C.c_void_p, dp, dp, dp, ip, C.c_int, dp]
        L.go_last_plastic_count.restype = C.c_int64
        L.go_last_plastic_count.argtypes = [C.c_void_p]
        L.go_last_error.restype = C.c_char_p
        L.go_last_error.argtypes = [C.c_void_p]
        L.go_apply_bforce.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.go_time_jacobian_elements.restype = C.c_double
        L.go_time_jacobian_elements.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class Oracle:
    """One mesh part in the reference's ghost (overlap) numbering."""

    def __init__(self, coords, conn, model, materials, elem_set=None):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, 3)
        self.conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, 4)
        self.nn, self.ne = len(self.coords), len(self.conn)
        self.model = model
        mats = np.ascontiguousarray(materials, dtype=np.float64).reshape(-1, 5)
        self.eset = None if elem_set is None else np.ascontiguousarray(elem_set, dtype=np.int32)
        self.L = lib()
        self.h = self.L.go_create(self.nn, self.ne, _ip(self.conn), _dp(self.coords),
                                  None if self.eset is None else _ip(self.eset),
                                  len(mats), MODEL[model], _dp(mats))
        self.nnz = self.L.go_graph_nnz(self.h)
        self.rowptr = np.ctypeslib.as_array(self.L.go_graph_rowptr(self.h), (4 * self.nn + 1,)).copy()
        self.colind = np.ctypeslib.as_array(self.L.go_graph_colind(self.h), (self.nnz,)).copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.go_destroy(self.h)
            self.h = None

    def _check(self, rc):
        if rc:
            raise RuntimeError(self.L.go_last_error(self.h).decode())

    def set_solution(self, u, p):
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
        p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1)
        assert u.size == 3 * self.nn and p.size == self.nn
        self.L.go_set_solution(self.h, _dp(u), _dp(p))

    def state(self, name):
        """numpy view of a state array (sigma/Fp/Fp_old: [Ne,9]; eqps/eqps_old: [Ne])."""
        ptr = self.L.go_state(self.h, name.encode())
        if not ptr:
            raise KeyError(name)
        shape = (self.ne, 9) if name in ("sigma", "Fp", "Fp_old") else (self.ne,)
        return np.ctypeslib.as_array(ptr, shape)

    def update_states(self):
        self.L.go_update_states(self.h)

    def residual(self, save=True, R=None):
        R = np.zeros(4 * self.nn) if R is None else R
        self._check(self.L.go_assemble_residual(self.h, int(save), _dp(R)))
        return R

    def jacobian(self, mode=PRIMAL, save=True, R=None, values=None):
        R = np.zeros(4 * self.nn) if R is None else R
        values = np.zeros(self.nnz) if values is None else values
        self._check(self.L.go_assemble_jacobian(self.h, mode, int(save), _dp(R), _dp(values)))
        return R, values

    def avg_disp(self, with_dMdu=False):
        if with_dMdu:
            d = np.zeros(4 * self.nn)
            return self.L.go_functional_avg_disp(self.h, _dp(d)), d
        return self.L.go_functional_avg_disp(self.h, None)

    def functional(self, type, elem_set=0, rho=1.0, point=(0, 0), with_dMdu=False):
        """Any of the reference's functionals by its yaml `type` (src/goal_mechanics.cpp:149-167)."""
        d = np.zeros(4 * self.nn) if with_dMdu else None
        J = self.L.go_functional(self.h, QOI[type], elem_set, rho, point[0], point[1], None if d is None else _dp(d))
        if J != J:
            raise RuntimeError(self.L.go_last_error(self.h).decode())
        return (J, d) if with_dMdu else J

    def localize(self, zu_diff, zp_diff, zp_coarse, R=None):
        R = np.zeros(4 * self.nn) if R is None else R
        a = [np.ascontiguousarray(x, dtype=np.float64).reshape(-1) for x in (zu_diff, zp_diff, zp_coarse)]
        self._check(self.L.go_assemble_error(self.h, _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(R)))
        return R

    def apply_bforce(self, b, R, zu_diff=None):
        """BForce (goal_bforce.cpp:58-68) added to the ghost R: b [Ne,3] per element; zu_diff selects the error chain's weights."""
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
        z = None if zu_diff is None else np.ascontiguousarray(zu_diff, dtype=np.float64).reshape(-1)
        self._check(self.L.go_apply_bforce(self.h, _dp(b), None if z is None else _dp(z), _dp(R)))
        return R

    def element_error(self, u_err, p_err, parent=None, n_parent=0):
        u_err = np.ascontiguousarray(u_err, dtype=np.float64).reshape(-1)
        p_err = np.ascontiguousarray(p_err, dtype=np.float64).reshape(-1)
        eta = np.zeros(self.ne)
        if parent is not None:
            parent = np.ascontiguousarray(parent, dtype=np.int32)
            etap = np.zeros(n_parent)
            b = self.L.go_element_error(self.h, _dp(u_err), _dp(p_err), _dp(eta), _ip(parent), n_parent, _dp(etap))
            return eta, etap, b
        b = self.L.go_element_error(self.h, _dp(u_err), _dp(p_err), _dp(eta), None, 0, None)
        return eta, None, b

    def size_field(self, eta, target, p_order=1):
        eta = np.ascontiguousarray(eta, dtype=np.float64)
        v = np.zeros(self.nn)
        G = self.L.go_size_field(self.h, _dp(eta), target, p_order, _dp(v))
        return v, G

    def plastic_count(self):
        return self.L.go_last_plastic_count(self.h)

    def time_jacobian_elements(self, e0, e1, save=False):
        return self.L.go_time_jacobian_elements(self.h, e0, e1, int(save))

    def csr(self, values):
        import scipy.sparse as sp
        return sp.csr_matrix((values, self.colind, self.rowptr), shape=(4 * self.nn, 4 * self.nn))
