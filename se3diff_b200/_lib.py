"""ctypes binding of libse3diff_b200.so (C ABI: include/se3diff_b200.h).  Fails loudly."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SE3DIFF_B200_LIB") or os.path.join(_HERE, "_lib", "libse3diff_b200.so")   # override: developer experiments only
_lock = threading.Lock()
_lib = None

f32p, f64p, i64, i32, u64, vp = C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_void_p
f32, f64 = C.c_float, C.c_double


class EmScalars(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "dt", "sqrt_abs_dt", "noise_weight", "score_weight", "rot_g", "rot_scale", "pos_beta", "pos_sqrt_beta",
        "pos_std", "tol")]


class DpmScalars(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "pos_std_t", "pos_c_x_mid", "pos_c_s_mid", "pos_std_lam", "pos_c_x_fin", "pos_c_s_fin", "rot_scale_t",
        "rot_scale_lam", "rot_g_t", "rot_g_lam", "dt_mid", "dt", "tol")]


class HeunScalars(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "churn_dt", "churn_sqrt_abs_dt", "churn_rot_g", "churn_pos_beta", "churn_pos_sqrt_beta", "step_dt",
        "hat_rot_g", "hat_rot_scale", "hat_pos_beta", "hat_pos_sqrt_beta", "hat_pos_std", "next_rot_g",
        "next_rot_scale", "next_pos_beta", "next_pos_sqrt_beta", "next_pos_std", "tol")]


class IpaShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "batch", "len", "heads", "dk", "pq", "pv", "proj_stride", "off_q", "off_k", "off_v", "off_qp", "off_kp",
        "off_vp", "hs_scalar", "hs_point", "hs_vpoint", "pair_batch")]


ABI_VERSION = 5   # SE3_ABI_VERSION of include/se3diff_b200.h this module's signature table was written against

# name -> argtypes (all return int unless listed in _RESTYPES).  Must list every symbol of the header.
SIGNATURES = {
    "se3_so3_exp": [f32p, f32p, i64, f32, vp],
    "se3_so3_exp_f64": [f64p, f64p, i64, f64, vp],
    "se3_so3_log": [f32p, f32p, i64, vp],
    "se3_so3_log_f64": [f64p, f64p, i64, vp],
    "se3_so3_angle": [f32p, f32p, f32p, f32p, i64, vp],
    "se3_so3_compose_rotvec": [f32p, f32p, f32p, i64, f32, vp],
    "se3_so3_matmul": [f32p, f32p, f32p, i64, i32, vp],
    "se3_so3_rel_log": [f32p, f32p, f32p, i64, vp],
    "se3_so3_geodesic": [f32p, f32p, f32, f32p, i64, f32, vp],
    "se3_so3_from_quat": [f32p, f32p, f32p, i64, f32, vp],
    "se3_frame_update_em": [f32p] * 12 + [i64, C.POINTER(EmScalars), vp],
    "se3_so3_update_em": [f32p] * 6 + [i64, C.POINTER(EmScalars), vp],
    "se3_frame_update_dpm_mid": [f32p] * 6 + [i64, C.POINTER(DpmScalars), vp],
    "se3_frame_update_dpm_final": [f32p] * 7 + [i64, C.POINTER(DpmScalars), vp],
    "se3_frame_heun_churn": [f32p] * 6 + [i64, C.POINTER(HeunScalars), vp],
    "se3_frame_heun_predict": [f32p] * 6 + [i64, C.POINTER(HeunScalars), vp],
    "se3_frame_heun_correct": [f32p] * 9 + [i64, C.POINTER(HeunScalars), vp],
    "se3_frame_traceback": [f32p] * 10 + [i64, C.POINTER(EmScalars), vp],
    "se3_r3_update_em": [f32p] * 6 + [i64, C.POINTER(EmScalars), vp],
    "se3_r3_update_dpm": [f32p] * 3 + [i64, C.POINTER(DpmScalars), i32, vp],
    "se3_r3_heun_churn": [f32p] * 3 + [i64, C.POINTER(HeunScalars), vp],
    "se3_r3_heun_step": [f32p] * 5 + [i64, C.POINTER(HeunScalars), vp],
    "se3_igso3_series_f32": [f32p] * 5 + [i64, i32, f32, vp],
    "se3_igso3_series_f64": [f64p] * 5 + [i64, i32, f64, vp],
    "se3_igso3_score": [f32p, f32p, f32p, i64, i32, f32, vp],
    "se3_igso3_marginal_pdf": [f32p, f32p, f32p, f32p, i64, i32, f32, vp],
    "se3_igso3_build_cdf": [f32p, i32, f64p, i32, i32, f64, i32, f32p, vp],
    "se3_igso3_build_score_scaling": [f32p, i32, f64p, i32, i32, f64, f32p, vp],
    "se3_igso3_sample": [f32p, f32p, i32, f32p, f32p, i32, f32p, f32p, u64, f32p, f32p, f32p, i64, f32, f32p, vp],
    "se3_igso3_build_cdf_index": [f32p, i32, i32, f32p, vp],
    "se3_igso3_cdf_index_floats": [i32, i32],
    "se3_ipa_attention_fwd": [f32p] * 7 + [f32, f32p, C.POINTER(IpaShape), i32, vp],
    "se3_ipa_attention_bwd": [f32p] * 7 + [f32] + [f32p] * 6 + [C.POINTER(IpaShape), vp],
    "se3_ipa_tc_workspace_bytes": [C.POINTER(IpaShape), C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
    "se3_ipa_attention_tc_fwd": [vp, i64, vp, i32, i64, f32p, f32p, vp, vp, f32p, f32p, vp, i32, vp, f32p, C.POINTER(IpaShape), vp],
    "se3_ipa_tc_packed_pair_bytes": [i32, i32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
    "se3_ipa_tc_pack_pair": [f32p, f32p, vp, vp, i32, i32, vp],
    "se3_ipa_split_perm": [i32, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
    "se3_pair_embed": [f32p, f32p, f32p, f32, f32p, f32p, vp, f32p, f32p, i64, i32, i32, i32, vp],
    "se3_pair_project": [f32p, f32p, f32, vp, vp, i32, i64, i32, i32, i32, i32, vp],
    "se3_folded_proportion": [f32p, f32p, f32p, f32p, i64, i32, f32, f32, f32, vp],
    "se3_backbone_atoms": [f32p, f32p, vp, vp, f32p, i64, i32, vp],
    "se3_physicality": [f32p, vp, f32p, i64, i32, vp],
    "se3_residual_layernorm": [f32p, vp, i32, f32p, f32p, f32p, f32, vp, i32, i64, i32, vp],
    "se3_bias_relu_project3": [f32p] * 6 + [i64, i32, vp],
    "se3_gelu_bf16": [vp, vp, i64, vp],
    "se3_debug_umma_gemm": [vp, vp, f32p, i32, i32, vp],
    "se3_last_error": [],
    "se3_abi_version": [],
    "se3_launch_count": [],
    "se3_launch_count_reset": [],
}
_RESTYPES = {"se3_ipa_tc_workspace_bytes": C.c_int64, "se3_ipa_tc_packed_pair_bytes": C.c_int64, "se3_igso3_cdf_index_floats": C.c_int64, "se3_last_error": C.c_char_p, "se3_launch_count": C.c_int64, "se3_launch_count_reset": None}


class Se3LibraryError(RuntimeError):
    pass


def lib():
    """Loads the shared library (building it with nvcc if it is absent and nvcc exists).  Raises
    Se3LibraryError when it cannot be had -- there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        # A library that does not match the sources / header on disk would be called with mismatched pointers and structs:
        # rebuild when the digest differs (a no-op when it matches), and refuse a stale binary that cannot be rebuilt.
        from .build import build, stale

        if stale():
            try:
                build()
            except Exception as e:  # noqa: BLE001
                what = "is missing" if not os.path.exists(LIB_PATH) else "was built from other sources than the ones on disk"
                raise Se3LibraryError(
                    f"libse3diff_b200.so {what} at {LIB_PATH} and could not be rebuilt ({e}); "
                    "run `python -m se3diff_b200.build` (needs nvcc)") from e
        try:
            h = C.CDLL(LIB_PATH)
        except OSError as e:
            raise Se3LibraryError(f"cannot load {LIB_PATH}: {e}") from e
        for name, args in SIGNATURES.items():
            try:
                fn = getattr(h, name)
            except AttributeError as e:
                raise Se3LibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        if h.se3_abi_version() != ABI_VERSION:
            raise Se3LibraryError(f"{LIB_PATH} reports ABI version {h.se3_abi_version()}, this package was written against {ABI_VERSION}; rebuild it")
        _lib = h
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().se3_last_error()
        raise Se3LibraryError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
