"""ctypes binding of the C ABI declared in ``include/lqpb.h``.

The shared library ``_lqpb.so`` (built by ``python -m lqp_py_b200.build`` /
``__graft_entry__.build()``) is the only compute backend: there is no CPU or PyTorch
fallback.  If the library is missing, or a compute call is made without a CUDA
device, this module raises -- it never silently degrades.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "_lqpb.so")
LOG_CAP = 64
E_TAPE = 5          # LQPB_E_TAPE

#: every symbol include/lqpb.h declares (checked by tests/test_abi_cpu.py)
EXPORTS = (
    "lqpb_abi_version", "lqpb_last_error", "lqpb_profile_enable", "lqpb_profile_get",
    "lqpb_forward_workspace_bytes_f32", "lqpb_forward_workspace_bytes_f64",
    "lqpb_forward_f32", "lqpb_forward_f64", "lqpb_forward_warm_f32", "lqpb_forward_warm_f64",
    "lqpb_solution_status_f32", "lqpb_solution_status_f64", "lqpb_forward_async_f32", "lqpb_forward_async_f64",
    "lqpb_ctrl_bytes", "lqpb_forward_collect", "lqpb_iterate_regime_f32", "lqpb_iterate_regime_f64",
    "lqpb_backward_workspace_bytes_f32", "lqpb_backward_workspace_bytes_f64",
    "lqpb_backward_f32", "lqpb_backward_f64", "lqpb_backward_kkt_f32", "lqpb_backward_kkt_f64",
    "lqpb_forward_prep_f32", "lqpb_forward_prep_f64", "lqpb_backward_finish_f32", "lqpb_backward_finish_f64",
    "lqpb_forward_host_f32", "lqpb_forward_host_f64", "lqpb_backward_host_f32", "lqpb_backward_host_f64",
    "lqpb_unroll_snapshot_bytes_f32", "lqpb_unroll_snapshot_bytes_f64",
    "lqpb_unroll_record_f32", "lqpb_unroll_record_f64", "lqpb_unroll_forward_f32", "lqpb_unroll_forward_f64",
    "lqpb_unroll_backward_f32", "lqpb_unroll_backward_f64",
    "lqpb_unroll_scale_grad_scratch_elems_f32", "lqpb_unroll_scale_grad_scratch_elems_f64",
    "lqpb_unroll_scale_grad_f32", "lqpb_unroll_scale_grad_f64",
    "lqpb_unroll_scaled_vectors_f32", "lqpb_unroll_scaled_vectors_f64",
    "lqpb_unroll_scale_vec_grad_f32", "lqpb_unroll_scale_vec_grad_f64",
    "lqpb_unroll_colmax_f32", "lqpb_unroll_colmax_f64", "lqpb_unroll_colmax_grad_f32", "lqpb_unroll_colmax_grad_f64",
    "lqpb_lu_factor_f32", "lqpb_lu_factor_f64", "lqpb_lu_solve_f32", "lqpb_lu_solve_f64",
    "lqpb_outer_f32", "lqpb_outer_f64",
    "lqpb_dev_tc_inverse_work_bytes", "lqpb_dev_tc_inverse_f32", "lqpb_dev_stream_read",
    "lqpb_copy_mapped",
)


class Config(C.Structure):
    """``lqpb_config``"""
    _fields_ = [
        ("max_iters", C.c_int32), ("check_solved", C.c_int32), ("adaptive_rho", C.c_int32),
        ("adaptive_rho_iter", C.c_int32), ("adaptive_rho_max_iter", C.c_int32), ("scale", C.c_int32),
        ("rho_auto", C.c_int32), ("beta_auto", C.c_int32), ("verbose", C.c_int32), ("keep_operators", C.c_int32),
        ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("rho", C.c_double), ("rho_min", C.c_double),
        ("rho_max", C.c_double), ("adaptive_rho_tol", C.c_double), ("adaptive_rho_threshold", C.c_double),
        ("beta", C.c_double), ("zero_clamp", C.c_double),
    ]


class Info(C.Structure):
    """``lqpb_info``"""
    _fields_ = [
        ("iter", C.c_int32), ("status", C.c_int32), ("n_factor", C.c_int32), ("any_lb", C.c_int32),
        ("any_ub", C.c_int32), ("n_log", C.c_int32), ("log_iter", C.c_int32 * LOG_CAP),
        ("log_primal", C.c_double * LOG_CAP), ("log_dual", C.c_double * LOG_CAP),
    ]


class Profile(C.Structure):
    """``lqpb_profile``"""
    _fields_ = [
        ("scale_ms", C.c_float), ("factor_ms", C.c_float), ("iterate_ms", C.c_float), ("finalize_ms", C.c_float),
        ("bwd_factor_ms", C.c_float), ("bwd_solve_ms", C.c_float), ("bwd_grad_ms", C.c_float),
        ("iterate_launches", C.c_int32), ("factor_launches", C.c_int32), ("kernel_launches", C.c_int32),
    ]


_lib = None


class LqpbError(RuntimeError):
    pass


def lib():
    """Load ``_lqpb.so`` once and declare the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built. "
            "Run `python -m lqp_py_b200.build` (needs nvcc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, dbl, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
    L.lqpb_abi_version.restype = i32
    L.lqpb_last_error.restype = C.c_char_p
    L.lqpb_profile_enable.argtypes = [i32]
    L.lqpb_profile_enable.restype = None
    L.lqpb_profile_get.argtypes = [C.POINTER(Profile)]
    L.lqpb_profile_get.restype = None
    for sfx in ("f32", "f64"):
        f = getattr(L, f"lqpb_forward_workspace_bytes_{sfx}")
        f.argtypes, f.restype = [i32, i32, i32], sz
        f = getattr(L, f"lqpb_backward_workspace_bytes_{sfx}")
        f.argtypes, f.restype = [i32, i32, i32], sz
        f = getattr(L, f"lqpb_forward_{sfx}")
        f.argtypes = [C.POINTER(Config), i32, i32, i32] + [vp] * 6 + [vp] * 6 + [C.POINTER(Info), vp, sz, vp]
        f.restype = i32
        f = getattr(L, f"lqpb_forward_warm_{sfx}")
        f.argtypes = [C.POINTER(Config), i32, i32, i32] + [vp] * 6 + [vp] * 3 + [vp] * 6 + [C.POINTER(Info), vp, sz, vp]
        f.restype = i32
        f = getattr(L, f"lqpb_forward_async_{sfx}")
        f.argtypes = ([C.POINTER(Config), i32, i32, i32] + [vp] * 6 + [vp] * 2 + [vp] * 6 + [vp, C.POINTER(Info), vp, sz, vp,
                      C.POINTER(C.c_int32)])
        f.restype = i32
        f = getattr(L, f"lqpb_iterate_regime_{sfx}")
        f.argtypes, f.restype = [C.POINTER(Config), i32, i32, i32], i32
        f = getattr(L, f"lqpb_solution_status_{sfx}")
        f.argtypes, f.restype = [C.POINTER(Config), i32, i32, i32, vp, sz, vp, vp, vp], i32
        f = getattr(L, f"lqpb_backward_{sfx}")
        f.argtypes = [i32, i32, i32] + [vp] * 9 + [vp, dbl] + [vp] * 6 + [vp, sz, vp]
        f.restype = i32
        f = getattr(L, f"lqpb_backward_kkt_{sfx}")
        f.argtypes = [i32, i32, i32] + [vp] * 8 + [vp] * 6 + [C.POINTER(C.c_int32), vp, sz, vp]
        f.restype = i32
        f = getattr(L, f"lqpb_forward_prep_{sfx}")
        f.argtypes = ([C.POINTER(Config), i32, i32, i32] + [vp] * 6 + [vp] * 6 + [C.POINTER(Info), vp, sz, vp, sz, i32,
                      C.POINTER(C.c_int32), vp])
        f.restype = i32
        f = getattr(L, f"lqpb_backward_finish_{sfx}")
        f.argtypes = [i32, i32, i32, i32] + [vp] * 9 + [vp, dbl] + [vp] * 6 + [vp, sz, vp]
        f.restype = i32
        f = getattr(L, f"lqpb_forward_host_{sfx}")
        f.argtypes = ([C.POINTER(Config), i32, i32, i32] + [vp] * 6 + [vp] * 6 + [vp] * 6 + [vp]
                      + [C.POINTER(Info), vp, sz, vp, i32, vp, sz, i32, C.POINTER(C.c_int32)])
        f.restype = i32
        f = getattr(L, f"lqpb_backward_host_{sfx}")
        f.argtypes = ([i32, i32, i32, i32] + [vp] * 2 + [vp] * 8 + [vp, dbl] + [vp] * 6 + [vp] * 6
                      + [C.POINTER(C.c_int32), vp, sz, vp, i32, i32])
        f.restype = i32
        f = getattr(L, f"lqpb_unroll_record_{sfx}")
        f.argtypes, f.restype = [C.POINTER(Config), i32, i32, i32, i32, vp, sz] + [vp] * 4 + [vp], i32
        f = getattr(L, f"lqpb_unroll_snapshot_bytes_{sfx}")
        f.argtypes, f.restype = [i32, i32, i32], sz
        f = getattr(L, f"lqpb_unroll_forward_{sfx}")
        f.argtypes = ([C.POINTER(Config), i32, i32, i32, i32, i32] + [vp] * 6 + [vp] * 6 + [vp] * 4 + [vp, sz]
                      + [C.POINTER(C.c_int32), vp, C.POINTER(Info), vp, sz, vp])
        f.restype = i32
        f = getattr(L, f"lqpb_unroll_backward_{sfx}")
        f.argtypes = [i32] * 6 + [vp, sz, vp] + [vp] * 4 + [vp] * 4 + [vp] * 2 + [vp] * 7 + [vp] * 2 + [vp]
        f.restype = i32
        f = getattr(L, f"lqpb_unroll_scale_grad_scratch_elems_{sfx}")
        f.argtypes, f.restype = [i32, i32], sz
        f = getattr(L, f"lqpb_unroll_scale_grad_{sfx}")
        f.argtypes, f.restype = [i32, i32] + [vp] * 6 + [vp], i32
        f = getattr(L, f"lqpb_unroll_scaled_vectors_{sfx}")
        f.argtypes, f.restype = [i32, i32, i32, vp, sz] + [vp] * 7 + [vp], i32
        f = getattr(L, f"lqpb_unroll_scale_vec_grad_{sfx}")
        f.argtypes, f.restype = [i32, i32, i32, i32, dbl, i32, i32] + [vp] * 8 + [vp] * 7 + [vp] * 6 + [vp], i32
        f = getattr(L, f"lqpb_unroll_colmax_{sfx}")
        f.argtypes, f.restype = [i32, i32, vp, vp, vp], i32
        f = getattr(L, f"lqpb_unroll_colmax_grad_{sfx}")
        f.argtypes, f.restype = [i32, i32, vp, vp, vp, vp, vp], i32
        f = getattr(L, f"lqpb_lu_factor_{sfx}")
        f.argtypes, f.restype = [i32, i32, vp, vp, vp, vp], i32
        f = getattr(L, f"lqpb_lu_solve_{sfx}")
        f.argtypes, f.restype = [i32, i32, i32, vp, vp, vp, vp, i32, vp], i32
        f = getattr(L, f"lqpb_outer_{sfx}")
        f.argtypes, f.restype = [i32, i32, i32, vp, vp, vp, vp], i32
    L.lqpb_ctrl_bytes.argtypes, L.lqpb_ctrl_bytes.restype = [], sz
    L.lqpb_forward_collect.argtypes, L.lqpb_forward_collect.restype = [vp, C.POINTER(Config), C.POINTER(Info)], i32
    L.lqpb_dev_tc_inverse_work_bytes.argtypes, L.lqpb_dev_tc_inverse_work_bytes.restype = [i32, i32], sz
    L.lqpb_dev_tc_inverse_f32.argtypes, L.lqpb_dev_tc_inverse_f32.restype = [i32, i32, vp, vp, vp, vp], i32
    L.lqpb_dev_stream_read.argtypes, L.lqpb_dev_stream_read.restype = [vp, sz, i32, vp, vp], i32
    L.lqpb_copy_mapped.argtypes, L.lqpb_copy_mapped.restype = [vp, vp, sz, vp], i32
    if L.lqpb_abi_version() != 1:
        raise ImportError("lqpb ABI version mismatch: rebuild with `python -m lqp_py_b200.build --force`")
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().lqpb_last_error().decode("utf-8", "replace")
        raise LqpbError(f"{what} failed (code {rc}): {msg}")


def suffix(dtype):
    import torch
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise TypeError(f"lqp_py_b200 supports float32 and float64 tensors, got {dtype}")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def profile_enable(on=True):
    lib().lqpb_profile_enable(1 if on else 0)


def profile_get():
    pr = Profile()
    lib().lqpb_profile_get(C.byref(pr))
    return {name: getattr(pr, name) for name, _ in Profile._fields_}
