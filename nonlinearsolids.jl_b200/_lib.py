"""ctypes binding of libnls_b200.so (the C ABI declared in include/nls_b200.h).

This is the same seam a Julia `ccall` shim binds. The library is the
product: if it is missing, or no CUDA device is present, everything here fails loudly --
there is no CPU path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NLS_B200_LIB: a developer override used by tools/ to time kernel variants built side by side
LIB_PATH = os.environ.get("NLS_B200_LIB") or os.path.join(_HERE, "libnls_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)

STATUS = {0: "NLS_OK", 1: "NLS_ERR_ARG", 2: "NLS_ERR_STATE", 3: "NLS_ERR_CUDA", 4: "NLS_ERR_NCCL",
          5: "NLS_ERR_NOGPU", 6: "NLS_ERR_UNSUPPORTED"}

MAT_LINEAR_ELASTIC_1D, MAT_EXPONENTIAL_1D, MAT_ELASTOPLASTIC_1D, MAT_NEOHOOKEAN, MAT_J2_SMALL_STRAIN = 0, 1, 2, 3, 4


class NewtonOpts(C.Structure):
    _fields_ = [("type_modified", C.c_int32), ("atol", C.c_double), ("rtol", C.c_double),
                ("dtol", C.c_double), ("maxits", C.c_int32), ("lin_rtol", C.c_double),
                ("lin_maxits", C.c_int32)]


class NewtonResult(C.Structure):
    _fields_ = [("reason", C.c_int32), ("iterations", C.c_int32), ("lin_its_total", C.c_int32),
                ("nhist", C.c_int32), ("norm_r", C.c_double), ("norm_r0", C.c_double),
                ("ms_assembly", C.c_double), ("ms_linear", C.c_double), ("ms_total", C.c_double)]


# every symbol include/nls_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
PROTOTYPES = {
    "nls_create": (C.c_int32, [C.c_int32, C.POINTER(_H)]),
    "nls_destroy": (C.c_int32, [_H]),
    "nls_last_error": (C.c_char_p, [_H]),
    "nls_version": (C.c_char_p, []),
    "nls_lagrange": (C.c_int32, [C.c_int32, C.c_int32, _dp, _dp]),
    "nls_shape_eval": (C.c_int32, [C.c_int32, _dp, _dp, C.c_double, _dp, _dp]),
    "nls_gauss_quadrature": (C.c_int32, [C.c_int32, _dp, _dp]),
    "nls_pattern_host": (C.c_int32, [C.c_int32, C.c_int32, C.c_int64, _ip, C.c_int64, C.c_int64, _lp, _ip, _lp]),
    "nls_plan_stats": (C.c_int32, [C.c_int32, C.c_int32, C.c_int64, _ip, C.c_int64, C.c_int64, _dp, C.c_int32, _lp]),
    "nls_set_mesh": (C.c_int32, [_H, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _ip, _dp]),
    "nls_set_discretization": (C.c_int32, [_H, C.c_int32, C.c_int32]),
    "nls_set_material": (C.c_int32, [_H, C.c_int32, _dp, C.c_int32]),
    "nls_set_area": (C.c_int32, [_H, C.c_double]),
    "nls_set_dirichlet": (C.c_int32, [_H, C.c_int64, _lp, _dp]),
    "nls_set_dirichlet_values": (C.c_int32, [_H, C.c_int64, _dp]),
    "nls_set_neumann": (C.c_int32, [_H, C.c_int64, _lp, _dp]),
    "nls_set_neumann_values": (C.c_int32, [_H, C.c_int64, _dp]),
    "nls_set_body_force": (C.c_int32, [_H, C.c_int32, _dp]),
    "nls_init_state": (C.c_int32, [_H]),
    "nls_commit_state": (C.c_int32, [_H, _dp]),
    "nls_get_state": (C.c_int32, [_H, C.c_char_p, _dp]),
    "nls_build_pattern": (C.c_int32, [_H]),
    "nls_pattern_size": (C.c_int32, [_H, _lp, _lp]),
    "nls_get_pattern": (C.c_int32, [_H, _lp, _ip]),
    "nls_residual": (C.c_int32, [_H, _dp, _dp]),
    "nls_jacobian": (C.c_int32, [_H, _dp, _dp]),
    "nls_residual_jacobian": (C.c_int32, [_H, _dp, _dp, _dp]),
    "nls_internal_force": (C.c_int32, [_H, _dp, _dp]),
    "nls_internal_force_jacobian": (C.c_int32, [_H, _dp, _dp]),
    "nls_set_values": (C.c_int32, [_H, _dp]),
    "nls_spmv": (C.c_int32, [_H, _dp, _dp]),
    "nls_dot": (C.c_int32, [_H, C.c_int64, _dp, _dp, _dp]),
    "nls_axpy": (C.c_int32, [_H, C.c_int64, C.c_double, _dp, _dp]),
    "nls_linear_solve": (C.c_int32, [_H, C.c_double, C.c_int32, _dp, _dp, _ip, _dp, _ip]),
    "nls_linear_solve_indefinite": (C.c_int32, [_H, C.c_double, C.c_int32, _dp, _dp, _ip, _dp, _ip]),
    "nls_newton_solve": (C.c_int32, [_H, C.POINTER(NewtonOpts), _dp, _dp, _dp, _dp, C.POINTER(NewtonResult)]),
    "nls_set_mass": (C.c_int32, [_H, C.c_double, C.c_double]),
    "nls_get_mass": (C.c_int32, [_H, _dp]),
    "nls_apply_mass": (C.c_int32, [_H, _dp, _dp]),
    "nls_dynamic_residual": (C.c_int32, [_H, _dp, _dp, _dp]),
    "nls_dynamic_jacobian": (C.c_int32, [_H, _dp, C.c_double, _dp]),
    "nls_mass_solve": (C.c_int32, [_H, C.c_double, C.c_int32, _dp, _dp, _ip, _ip]),
    "nls_newmark_solve": (C.c_int32, [_H, C.POINTER(NewtonOpts), C.c_double, C.c_double, C.c_double, _dp, _dp, _dp,
                                      _dp, _dp, _dp, _dp, _dp, C.POINTER(NewtonResult)]),
    "nls_dev_upload_u": (C.c_int32, [_H, _dp]),
    "nls_dev_download_u": (C.c_int32, [_H, _dp]),
    "nls_dev_download_r": (C.c_int32, [_H, _dp]),
    "nls_dev_download_vals": (C.c_int32, [_H, _dp]),
    "nls_dev_download_rows": (C.c_int32, [_H, C.c_int64, C.c_int64, _dp, _lp, _ip, _dp]),
    "nls_dev_assemble": (C.c_int32, [_H, C.c_int32, C.c_int32]),
    "nls_dev_spmv": (C.c_int32, [_H]),
    "nls_dev_pcg_iterations": (C.c_int32, [_H, C.c_int32]),
    "nls_dev_newton_step": (C.c_int32, [_H, C.c_double, C.c_int32, _ip, _dp]),
    "nls_dev_newton_solve": (C.c_int32, [_H, C.POINTER(NewtonOpts), _dp, C.POINTER(NewtonResult)]),
    "nls_sync": (C.c_int32, [_H]),
    "nls_timer_start": (C.c_int32, [_H]),
    "nls_timer_stop": (C.c_int32, [_H, C.POINTER(C.c_float)]),
    "nls_flush_l2": (C.c_int32, [_H]),
    "nls_kernel_launches": (C.c_int64, [_H]),
    "nls_set_option": (C.c_int32, [_H, C.c_char_p, C.c_int32]),
    "nls_debug_counters": (C.c_int32, [_H, _lp]),
    "nls_debug_counters_ext": (C.c_int32, [_H, _lp]),
    "nls_solver_prepare": (C.c_int32, [_H]),
    "nls_krylov_preconditioner": (C.c_int32, [_H]),
    "nls_assembly_scheme": (C.c_int32, [_H]),
    "nls_partition": (C.c_int32, [C.c_int32, C.c_int64, _ip, C.c_int32, _lp, C.c_int32,
                                  _lp, _lp, _lp, _lp, _lp, _lp]),
    "nls_set_halo": (C.c_int32, [_H, C.c_int64, C.c_int32, _ip, _lp, _ip, _lp, _ip]),
    "nls_comm_unique_id": (C.c_int32, [C.c_char_p, _bp]),
    "nls_comm_init": (C.c_int32, [_H, C.c_char_p, C.c_int32, C.c_int32, _bp]),
    "nls_dev_halo_exchange_u": (C.c_int32, [_H]),
    "nls_comm_path": (C.c_int32, [_H]),
    "nls_dev_barrier": (C.c_int32, [_H]),
    "nls_dev_comm_probe": (C.c_int32, [_H, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
}

_LIB = None


class NLSError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (STATUS.get(status, status), message))
        self.status = status


def lib():
    """Load libnls_b200.so; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libnls_b200.so not found at %s -- run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)   # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def nccl_lib_path():
    """Path of the libnccl the host process should share (torch's bundled copy if present)."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations or []) if spec else []:
            p = os.path.join(loc, "lib", "libnccl.so.2")
            if os.path.exists(p):
                return p
    except Exception:
        pass
    return "libnccl.so.2"


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def dptr(a):
    return a.ctypes.data_as(_dp)


def iptr(a):
    return a.ctypes.data_as(_ip)


def lptr(a):
    return a.ctypes.data_as(_lp)


class Handle:
    """Owns one nls_handle (one GPU). Thin, checked wrappers over the C ABI."""

    def __init__(self, device=0):
        self._L = lib()
        h = _H()
        rc = self._L.nls_create(C.c_int32(device), C.byref(h))
        if rc != 0:
            raise NLSError(rc, self._L.nls_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self._L.nls_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, name, *args):
        rc = getattr(self._L, name)(self.h, *args)
        if rc != 0:
            raise NLSError(rc, self._L.nls_last_error(self.h).decode())
        return rc
