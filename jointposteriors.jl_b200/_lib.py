"""ctypes binding of libjpcuda.so (include/jpcuda.h).  No torch types cross this boundary."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# JPCUDA_LIB selects an alternative build of the same library (kernel A/B experiments); default: the in-tree build
_SO = os.environ.get("JPCUDA_LIB") or os.path.join(_HERE, "libjpcuda.so")
_lib = None

JP_OK, JP_ERR_BAD_ARG, JP_ERR_NOT_PD, JP_ERR_CUDA, JP_ERR_NO_DEVICE, JP_ERR_ALLOC, JP_ERR_UNSUPPORTED, JP_ERR_COMM = range(8)
PATH_AUTO, PATH_FP64, PATH_TC = 0, 1, 2
GRID_KNOTS = 100


class JPError(RuntimeError):
    """Raised for every non-zero jp_status (the Julia shim rethrows as ErrorException)."""

    def __init__(self, status, message):
        super().__init__("libjpcuda status %d: %s" % (status, message))
        self.status = status


class NotPositiveDefinite(JPError):
    pass


class FitArgs(C.Structure):
    _fields_ = [
        ("d", C.c_int),
        ("p", C.c_int),
        ("h_transform", C.POINTER(C.c_int)),
        ("h_mu_hat", C.POINTER(C.c_double)),
        ("h_U", C.POINTER(C.c_double)),
        ("neg_min", C.c_double),
        ("path", C.c_int),
        ("node_begin", C.c_longlong),
        ("node_end", C.c_longlong),
        ("raw", C.c_int),
    ]


class SmoothCDF(C.Structure):
    """jp_smooth_cdf of include/jpcuda.h."""
    _fields_ = [
        ("beta", C.c_double * 10),
        ("theta", C.c_double * 7),
        ("phi", C.c_double * 9),
        ("mu", C.c_double),
        ("sigma", C.c_double),
        ("objective", C.c_double),
        ("grad_inf_norm", C.c_double),
        ("iterations", C.c_int),
        ("evaluations", C.c_int),
        ("converged", C.c_int),
    ]


# every symbol include/jpcuda.h declares (tests check that the .so exports exactly these)
SYMBOLS = [
    "jp_last_error", "jp_version",
    "jp_chol", "jp_try_chol", "jp_inv_upper", "jp_inv_chol", "jp_reduce_dimensions", "jp_reduce_dimensions_ldr", "jp_deduce_scale_dynamic",
    "jp_ctx_create", "jp_ctx_destroy", "jp_ctx_set_stream", "jp_ctx_sync", "jp_ctx_launch_count",
    "jp_ctx_last_kernel_ms", "jp_ctx_trace", "jp_ctx_trace_dump",
    "jp_grid_get", "jp_grid_size", "jp_grid_dim", "jp_grid_build_stats", "jp_grid_download", "jp_rule_info", "jp_rule_level_nodes",
    "jp_grid_level_cap",
    "jp_data_upload", "jp_data_adopt_device", "jp_data_free", "jp_glm_grad_hess", "jp_log_density_points", "jp_mode", "jp_mode_report",
    "jp_posterior_create", "jp_posterior_free", "jp_posterior_size",
    "jp_fit", "jp_fit_local", "jp_fit_local_sum", "jp_fit_normalise", "jp_fit_local_stats", "jp_fit_normalise_gathered",
    "jp_fit_prep_len", "jp_fit_prep_local", "jp_fit_prep_gathered", "jp_fit_coef_slab", "jp_fit_local_stats_prepared",
    "jp_comm_create", "jp_comm_ipc_handle", "jp_comm_connect_ipc", "jp_comm_connect_local", "jp_comm_bulk_bytes", "jp_comm_status",
    "jp_comm_destroy", "jp_comm_all_gather", "jp_fit_p2p", "jp_fit_p2p_check", "jp_marginal_coords_p2p",
    "jp_fit_p2p_obs", "jp_mode_p2p", "jp_get_cache",
    "jp_get_theta", "jp_get_logdens", "jp_get_density", "jp_dev_theta", "jp_dev_density", "jp_fit_path_used",
    "jp_fit_diagnostics",
    "jp_marginal_coords", "jp_marginal_values", "jp_marginal_sorted", "jp_marginal_buffer", "jp_marginal_knots_from_sort",
    "jp_marginal_local_moments", "jp_marginal_local_knots", "jp_marginal_local_knots_gathered",
    "jp_marginal_combine_gathered",
    "jp_quantile", "jp_cdf",
    "jp_marginal_smooth", "jp_marginal_smooth_keyed", "jp_smooth_objective", "jp_smooth_cdf_eval", "jp_smooth_pdf_eval", "jp_smooth_quantile_eval",
]


def lib():
    """Load libjpcuda.so.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise ImportError(
            "libjpcuda.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C jointposteriors.jl_b200/csrc`; there is no CPU fallback." % _SO)
    L = C.CDLL(_SO)
    L.jp_last_error.restype = C.c_char_p
    L.jp_grid_size.restype = C.c_longlong
    L.jp_posterior_size.restype = C.c_longlong
    L.jp_ctx_launch_count.restype = C.c_longlong
    L.jp_comm_bulk_bytes.restype = C.c_longlong
    L.jp_quantile.restype = C.c_double
    L.jp_cdf.restype = C.c_double
    L.jp_dev_theta.restype = C.c_void_p
    L.jp_dev_density.restype = C.c_void_p
    for name in ("jp_quantile", "jp_cdf"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double]
    for name in ("jp_smooth_cdf_eval", "jp_smooth_pdf_eval", "jp_smooth_quantile_eval"):
        getattr(L, name).restype = C.c_double
        getattr(L, name).argtypes = [C.c_void_p, C.c_double]
    _lib = L
    return L


def check(status):
    if status != JP_OK:
        msg = lib().jp_last_error().decode("utf-8", "replace")
        if status == JP_ERR_NOT_PD:
            raise NotPositiveDefinite(status, msg)
        raise JPError(status, msg)


def ptr(a):
    """void* of a numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def colmajor(a):
    return np.asfortranarray(np.array(a, dtype=np.float64))
