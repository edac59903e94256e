"""ctypes binding of ``liboak_b200.so`` (C ABI declared in ``include/oak_b200.h``).

This is the only place the host package touches native code.  There is **no CPU fallback**:
if the shared library is missing, or no CUDA device is visible, every compute entry point
raises ``OakNativeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OAK_B200_LIB") or os.path.join(_HERE, "liboak_b200.so")  # env: dev builds only

OAK_MAX_DEPTH = 16
DIM_RBF, DIM_BINARY, DIM_CATEGORICAL = 0, 1, 2
MEASURE_NONE, MEASURE_GAUSSIAN, MEASURE_UNIFORM, MEASURE_EMPIRICAL, MEASURE_MOG = 0, 1, 2, 3, 4
ESP_NEWTON_GIRARD, ESP_DIRECT = 0, 1
LINK_LOGIT, LINK_PROBIT = 0, 1


class OakNativeError(RuntimeError):
    pass


class DimDesc(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("column", C.c_int32),
        ("measure", C.c_int32),
        ("count", C.c_int32),
        ("rank", C.c_int32),
        ("reserved", C.c_int32),
        ("lengthscale", C.c_double),
        ("variance", C.c_double),
        ("m0", C.c_double),
        ("m1", C.c_double),
        ("v0", C.POINTER(C.c_double)),
        ("v1", C.POINTER(C.c_double)),
        ("v2", C.POINTER(C.c_double)),
    ]


class KernelDesc(C.Structure):
    _fields_ = [
        ("num_dims", C.c_int32),
        ("depth", C.c_int32),
        ("share_var_across_orders", C.c_int32),
        ("esp_algorithm", C.c_int32),
        ("variances", C.POINTER(C.c_double)),
        ("dims", C.POINTER(DimDesc)),
    ]


# the same layout for bulk packing
_DIM_DTYPE = np.dtype([("type", "<i4"), ("column", "<i4"), ("measure", "<i4"), ("count", "<i4"), ("rank", "<i4"),
                       ("reserved", "<i4"), ("lengthscale", "<f8"), ("variance", "<f8"), ("m0", "<f8"), ("m1", "<f8"),
                       ("v0", "<u8"), ("v1", "<u8"), ("v2", "<u8")], align=True)
assert _DIM_DTYPE.itemsize == C.sizeof(DimDesc)

_i64, _i32, _vp, _dp, _sz = C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); mirrors include/oak_b200.h one to one
SIGNATURES = {
    "oak_last_error": (C.c_char_p, []),
    "oak_version": (C.c_int, []),
    "oak_device_count": (C.c_int, []),
    "oak_spec_create": (C.c_int, [C.POINTER(KernelDesc), _vp, C.POINTER(_vp)]),
    "oak_spec_destroy": (C.c_int, [_vp]),
    "oak_spec_var_s_f64": (C.c_int, [_vp, _i32, C.POINTER(C.c_double), _vp]),
    "oak_component_diag_f64": (C.c_int, [_vp, C.POINTER(_i32), _i32, _vp, _i64, _dp, _vp]),
    "oak_additive_terms_f64": (C.c_int, [_dp, _i32, _i32, _i64, _dp, _vp]),
    "oak_points_bytes": (_sz, [_vp, _i64]),
    "oak_prepare_points_f64": (C.c_int, [_vp, _dp, _i64, _i64, _vp, _vp]),
    "oak_gram_f64": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, _i64, _dp, _i64, _vp]),
    "oak_gram_lower_f64": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _dp, _i64, _vp]),
    "oak_gram_lower_mirror_f64": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _dp, _i64, _dp, _i64, _vp]),
    "oak_gram_diag_f64": (C.c_int, [_vp, _vp, _i64, _dp, _vp]),
    "oak_gram_matvec_f64": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _dp, _dp, _vp]),
    "oak_component_gram_f64": (C.c_int, [_vp, C.POINTER(_i32), _i32, _vp, _i64, _vp, _i64, _dp, _i64, _vp]),
    "oak_component_predict_f64": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _i64, _vp, _i64, _dp, _dp, _vp]),
    "oak_gram_host_work_bytes": (_sz, [_vp, _i64, _i64, _i64, _i64]),
    "oak_gram_host_f64": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp]),
    "oak_gram_host_lower_work_bytes": (_sz, [_vp, _i64, _i64, _i64, _i64]),
    "oak_gram_host_lower_f64": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _i64, C.c_int, _vp, _vp]),
    "oak_gram_backward_work_bytes": (_sz, [_vp, _i64]),
    "oak_backward_grad_count": (_sz, [_vp]),
    "oak_spec_table_layout": (C.c_int, [_vp, _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "oak_backward_points_bytes": (_sz, [_vp, _i64]),
    "oak_prepare_backward_f64": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "oak_sobol_gaussian_terms_f64": (C.c_int, [_dp, _dp, _i64, C.c_double, C.c_double, C.c_double, C.c_double, _dp,
                                               _vp]),
    "oak_flow_forward_f64": (C.c_int, [_dp, _i64, _i64, C.c_double, C.c_int32, C.c_double, C.c_double, C.c_double,
                                       C.c_double, _dp, _i64, _vp]),
    "oak_flow_objective_work_bytes": (_sz, [_i64]),
    "oak_flow_objective_f64": (C.c_int, [_dp, _i64, _i64, C.c_double, C.c_int32, C.c_double, C.c_double, C.c_double,
                                         C.c_double, _dp, _vp, _vp]),
    "oak_svgp_moments_f64": (C.c_int, [_dp, _i64, C.c_int32, _i64, _dp, _dp, _dp, _dp, _dp, _vp]),
    "oak_svgp_moments_backward_f64": (C.c_int, [_dp, _i64, C.c_int32, _i64, _dp, _dp, _dp, _dp, _dp, _i64, _dp, _vp]),
    "oak_bernoulli_quadrature_f64": (C.c_int, [_dp, _dp, _dp, _i64, C.c_int32, C.c_double, _dp, _dp, C.c_int32,
                                               _dp, _dp, _dp, _dp, _vp]),
    "oak_gram_backward_rows_work_bytes": (_sz, [_vp, _i64, _i64]),
    "oak_gram_backward_rows_f64": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _dp, _i64, _dp, _dp, _i64, _vp, _vp]),
    "oak_gram_backward_f64": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _i64, _dp, _i64, _dp, _vp, _vp]),
    "oak_gram_diag_backward_f64": (C.c_int, [_vp, _vp, _vp, _i64, _dp, C.c_double, _dp, _vp, _vp]),
    "oak_column_unique_work_bytes": (_sz, [_i64]),
    "oak_column_unique_f64": (C.c_int, [_dp, _i64, _i64, _i64, _dp, _vp, _vp, _vp, _vp]),
    "oak_column_mean_f64": (C.c_int, [_dp, _i64, _i64, _i64, _dp, _vp, _vp]),
    "oak_sgpr_stats_count": (_sz, [_i64]),
    "oak_sgpr_stats_work_bytes": (_sz, [_i64, _i64]),
    "oak_sgpr_stats_f64": (C.c_int, [_vp, _vp, _i64, _vp, _dp, _i64, _i64, _dp, _vp, _vp]),
    "oak_sgpr_stats_keep_f64": (C.c_int, [_vp, _vp, _i64, _vp, _dp, _i64, _i64, _dp, _vp, _dp, _vp]),
    "oak_sgpr_finish_work_bytes": (_sz, [_i64]),
    "oak_sgpr_finish_f64": (C.c_int, [_dp, _dp, _i64, _i64, C.c_double, C.c_double, _dp, _dp, _vp, _vp]),
    "oak_sgpr_factor_count": (_sz, [_i64]),
    "oak_sgpr_factor_ld": (_i64, [_i64]),
    "oak_sgpr_lb_ld": (_i64, [_i64]),
    "oak_sgpr_factor_f64": (C.c_int, [_vp, _vp, _i64, C.c_double, C.c_int, C.c_double, _dp, _vp]),
    "oak_sgpr_stats2_work_bytes": (_sz, [_i64, _i64]),
    "oak_sgpr_stats2_f64": (C.c_int, [_vp, _vp, _i64, _dp, _vp, _dp, _i64, _i64, _dp, _vp, _dp, _vp]),
    "oak_sgpr_factor_stats_f64": (C.c_int, [_vp, _vp, _i64, C.c_double, C.c_int, C.c_double, _dp, _vp, _dp, _i64, _i64,
                                             _dp, _vp, _dp, C.c_int, _vp]),
    "oak_kmeans_work_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "oak_kmeans_center_f64": (C.c_int, [_dp, _i64, _i64, _i64, _dp, _dp, _vp, _vp]),
    "oak_kmeanspp_round_f64": (C.c_int, [_dp, _i64, _i64, _dp, _dp, _i64, _vp, _dp, _dp, _dp, _vp, _vp]),
    "oak_kmeans_lloyd_f64": (C.c_int, [_dp, _i64, _i64, _dp, _i64, _vp, _dp, _dp, _dp, _dp, C.c_int, _vp, _vp]),
    "oak_gmm1d_work_bytes": (_sz, [_i64, _i64]),
    "oak_gmm1d_estep_f64": (C.c_int, [_dp, _i64, _i64, _dp, _vp, _dp, _vp, _vp]),
    "oak_sgpr_finish2_work_bytes": (_sz, [_i64]),
    "oak_sgpr_finish2_f64": (C.c_int, [_dp, _dp, _i64, _i64, C.c_double, _dp, _dp, _dp, _vp, _vp]),
    "oak_chol_f64": (C.c_int, [_dp, _i64, _i64, _i64, _i64, C.c_int, _vp, _dp, _vp]),
    "oak_panel_gemm_work_bytes": (_sz, []),
    "oak_panel_gemm_f64": (C.c_int, [_dp, _i64, _dp, _i64, _dp, _i64, _i64, _i64, _i64, C.c_int, _dp, _dp, _vp, _vp]),
    "oak_gpr_finish_work_bytes": (_sz, [_i64]),
    "oak_gpr_finish_f64": (C.c_int, [_dp, _dp, _i64, C.c_double, _dp, _dp, _vp, _vp]),
    "oak_sobol_L_f64": (C.c_int, [_vp, _i32, _dp, _i64, _i64, C.c_double, C.c_double, _dp, _i64, _vp, _vp]),
    "oak_sobol_L_work_bytes": (_sz, [_vp, _i32, _i64]),
    "oak_sobol_quadforms_work_bytes": (_sz, [_i32, _i64]),
    "oak_sobol_quadforms_f64": (C.c_int, [_dp, _i32, _i64, _vp, _dp, _i32, _i32, _dp, _dp, _vp, _vp]),
    "oak_measure_fp64_peak": (C.c_int, [C.c_double, C.POINTER(C.c_double), _vp]),
    "oak_launch_count": (_i64, []),
}

_lib = None


def load():
    """Loads the shared library (idempotent).  Raises ``OakNativeError`` if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OakNativeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    try:
        import torch  # noqa: F401  (loads the CUDA runtime libraries first; plumbing only)
    except Exception:  # pragma: no cover
        pass
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().oak_last_error().decode()


def check(rc: int, what: str):
    if rc != 0:
        raise OakNativeError(f"{what}: {last_error()}")


def require_device():
    lib = load()
    if lib.oak_device_count() <= 0:
        raise OakNativeError("no CUDA device visible: the OAK B200 kernels have no CPU fallback")
    return lib


def launch_count() -> int:
    return int(load().oak_launch_count())


def _dptr(arr: Optional[np.ndarray]):
    if arr is None:
        return C.POINTER(C.c_double)()
    return arr.ctypes.data_as(C.POINTER(C.c_double))


class DimSpec:
    """Plain host-side record of one sub-kernel (what ``oak_dim_desc`` carries)."""

    def __init__(self, type, column, *, measure=MEASURE_NONE, lengthscale=1.0, variance=1.0,
                 m0=0.0, m1=0.0, v0=None, v1=None, v2=None, rank=0):
        self.type, self.column, self.measure = int(type), int(column), int(measure)
        self.lengthscale, self.variance = float(lengthscale), float(variance)
        self.m0, self.m1, self.rank = float(m0), float(m1), int(rank)
        f = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
        self.v0, self.v1, self.v2 = f(v0), f(v1), f(v2)
        if self.type == DIM_CATEGORICAL:
            self.count = int(self.v1.shape[0])
        elif self.type == DIM_BINARY:
            self.count = 2
        elif self.v0 is not None:
            self.count = int(self.v0.shape[0])
        else:
            self.count = 0


class Spec:
    """Owns an ``oak_spec*``.  Created per evaluation from the current parameter values."""

    def __init__(self, dims: Sequence[DimSpec], depth: int, variances, share_var=True,
                 algorithm=ESP_NEWTON_GIRARD, stream: int = 0):
        lib = require_device()
        if depth > OAK_MAX_DEPTH:
            raise OakNativeError(f"max_interaction_depth {depth} > OAK_MAX_DEPTH {OAK_MAX_DEPTH}")
        self._keep = list(dims)
        self.num_dims, self.depth = len(dims), int(depth)
        # oak_dim_desc records packed through one structured NumPy array (per-field ctypes stores cost ~6 us a dim)
        recs = np.zeros(len(dims), dtype=_DIM_DTYPE)
        ptr = lambda a: 0 if a is None else a.ctypes.data
        recs[:] = [(d.type, d.column, d.measure, d.count, d.rank, 0, d.lengthscale, d.variance, d.m0, d.m1,
                    ptr(d.v0), ptr(d.v1), ptr(d.v2)) for d in dims]
        self._recs = recs
        arr = recs.ctypes.data_as(C.POINTER(DimDesc))
        var = np.ascontiguousarray(np.asarray(variances, dtype=np.float64).reshape(-1))
        need = depth + 1 if share_var else 1
        if var.shape[0] < need:
            raise ValueError(f"need {need} order variances for depth {depth} (share_var={share_var}), got {var.shape[0]}")
        desc = KernelDesc(len(dims), int(depth), int(bool(share_var)), int(algorithm), _dptr(var), arr)
        handle = _vp()
        check(lib.oak_spec_create(C.byref(desc), _vp(stream), C.byref(handle)), "oak_spec_create")
        self.handle = handle
        self._lib = lib

    def points_bytes(self, n: int) -> int:
        return int(self._lib.oak_points_bytes(self.handle, n))

    def close(self):
        if getattr(self, "handle", None):
            self._lib.oak_spec_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
