"""ctypes loader for libtealeaf_b200.so (the C-ABI declared in include/tealeaf_b200.h).

The library is the product: there is no Python or CPU fallback.  If it has not been built
(`python -m exploringsycl_b200.build` or `__graft_entry__.build()`), importing a symbol raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtealeaf_b200.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "tealeaf_b200.h")


class TeaLeafError(RuntimeError):
    pass


class TlState(C.Structure):
    _fields_ = [("geometry", C.c_int), ("density", C.c_double), ("energy", C.c_double),
                ("x_min", C.c_double), ("y_min", C.c_double), ("x_max", C.c_double),
                ("y_max", C.c_double), ("radius", C.c_double)]


class TlSolveOpts(C.Structure):
    _fields_ = [("solver", C.c_int), ("coefficient", C.c_int), ("max_iters", C.c_int),
                ("eps", C.c_double), ("presteps", C.c_int), ("ppcg_inner_steps", C.c_int),
                ("error_switch", C.c_int), ("eps_lim", C.c_double), ("check_result", C.c_int),
                ("fuse_p_into_w", C.c_int), ("batch", C.c_int)]


class TlSolveInfo(C.Structure):
    _fields_ = [("iters_a", C.c_int), ("iters_b", C.c_int), ("est_iters", C.c_int),
                ("total_iters", C.c_int), ("error", C.c_double), ("eigmin", C.c_double),
                ("eigmax", C.c_double), ("gpu_ms", C.c_double), ("kernel_launches", C.c_long)]


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i4 = C.c_int * 4
_i6 = C.c_int * 6
_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)
_vp = C.c_void_p
_i, _d = C.c_int, C.c_double

# name -> (argtypes, restype).  Every symbol include/tealeaf_b200.h declares is listed here;
# tests/test_abi.py checks the two stay in sync.
SIGNATURES = {
    "tl_last_error": ([], C.c_char_p),
    "tl_device_count": ([], _i),
    "tl_version": ([], C.c_char_p),
    "tl_chunk_create": ([C.POINTER(_vp), _i, _i, _i, _i, _i, _i4, _i, _i], _i),
    "tl_chunk_destroy": ([_vp], _i),
    "tl_chunk_dims": ([_vp, _pi, _pi, _pi, _pi], _i),
    "tl_chunk_sync": ([_vp], _i),
    "tl_field_write": ([_vp, _i, _dp], _i),
    "tl_field_read": ([_vp, _i, _dp], _i),
    "tl_array_read": ([_vp, _i, _dp], _i),
    "tl_cg_alphas": ([_vp], _pd),
    "tl_cg_betas": ([_vp], _pd),
    "tl_cheby_alphas": ([_vp], _pd),
    "tl_cheby_betas": ([_vp], _pd),
    "tl_run_set_chunk_data": ([_vp, _d, _d, _d, _d], _i),
    "tl_run_set_chunk_state": ([_vp, _i, C.POINTER(TlState)], _i),
    "tl_run_local_halos": ([_vp, _i6, _i], _i),
    "tl_run_pack_or_unpack": ([_vp, _i, _i, _i, _i, _dp], _i),
    "tl_pack_face_device": ([_vp, _i6, _i, _i, _i, _pi], _i),
    "tl_face_buffer_read": ([_vp, _i, _i, _dp, _i], _i),
    "tl_face_buffer_write": ([_vp, _i, _i, _dp, _i], _i),
    "tl_run_store_energy": ([_vp], _i),
    "tl_run_field_summary": ([_vp, _pd, _pd, _pd, _pd], _i),
    "tl_run_cg_init": ([_vp, _i, _d, _d, _pd], _i),
    "tl_run_cg_calc_w": ([_vp, _pd], _i),
    "tl_run_cg_calc_ur": ([_vp, _d, _pd], _i),
    "tl_run_cg_calc_p": ([_vp, _d], _i),
    "tl_run_cheby_init": ([_vp, _d], _i),
    "tl_run_cheby_iterate": ([_vp, _d, _d], _i),
    "tl_run_jacobi_init": ([_vp, _i, _d, _d], _i),
    "tl_run_jacobi_iterate": ([_vp, _pd], _i),
    "tl_run_ppcg_init": ([_vp, _d], _i),
    "tl_run_ppcg_inner_iteration": ([_vp, _d, _d], _i),
    "tl_run_copy_u": ([_vp], _i),
    "tl_run_calculate_residual": ([_vp], _i),
    "tl_run_calculate_2norm": ([_vp, _i, _pd], _i),
    "tl_run_finalise": ([_vp], _i),
    "tl_comms_create": ([C.POINTER(_vp), C.c_char_p, _i, _i, _i, _i], _i),
    "tl_comms_destroy": ([_vp], _i),
    "tl_comms_rank": ([_vp], _i),
    "tl_comms_size": ([_vp], _i),
    "tl_comms_barrier": ([_vp], _i),
    "tl_comms_abort": ([_vp], _i),
    "tl_comms_sum": ([_vp, _pd], _i),
    "tl_comms_min": ([_vp, _pd], _i),
    "tl_comms_send_recv": ([_vp, _dp, _dp, _i, _i, _i, _i], _i),
    "tl_comms_post": ([_vp, _dp, _i, _i, _i], _i),
    "tl_comms_recv": ([_vp, _dp, _i, _i, _i], _i),
    "tl_comms_attach_chunk": ([_vp, _vp], _i),
    "tl_decompose": ([_i, _i, _i, _i, _pi, _pi, _pi, _pi, _i4, _pi, _pi], _i),
    "tl_halo_update": ([_vp, _vp, _i6, _i], _i),
    "tl_halo_stress": ([_vp, _vp, _i, _i, _i, _i, C.POINTER(C.c_long)], _i),
    "tl_solve_opts_default": ([C.POINTER(TlSolveOpts)], None),
    "tl_cg_loop_is_fused": ([_vp, _i], _i),
    "tl_solve": ([_vp, _vp, C.POINTER(TlSolveOpts), _d, _d, C.POINTER(TlSolveInfo)], _i),
    "tl_timestep": ([_vp, _vp, C.POINTER(TlSolveOpts), _d, _d, _d, C.POINTER(TlSolveInfo)], _i),
    "tl_field_summary": ([_vp, _vp, _pd, _pd, _pd, _pd], _i),
    "tl_timestep_host": ([_vp, _vp, C.POINTER(TlSolveOpts), _d, _d, _d, _vp, _vp,
                          C.POINTER(TlSolveInfo), C.POINTER(C.c_double * 4)], _i),
    "tl_host_alloc_pinned": ([C.c_long], _vp),
    "tl_host_free_pinned": ([_vp], None),
    "tl_time_kernel": ([_vp, _i, _i, _pd], _i),
    "tl_set_tuning": ([_i, _i, _i], _i),
    "tl_set_pw_pipeline": ([_i, _i, _i], _i),
    "tl_stamps_enable": ([_vp, _i], _i),
    "tl_stamps_read": ([_vp, C.POINTER(C.c_ulonglong), _i], _i),
    "tl_kernel_launch_count": ([], C.c_long),
    "tl_timer_start": ([_vp], _i),
    "tl_timer_stop": ([_vp, _pd], _i),
}

_lib = None


def lib():
    """Load the shared library (once). Raises TeaLeafError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TeaLeafError(
                "%s is missing: build it with `python -m exploringsycl_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (args, res) in SIGNATURES.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = res
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise TeaLeafError("tealeaf_b200 error %d: %s" % (rc, lib().tl_last_error().decode()))
