"""ctypes binding of libadept_b200.so (the C ABI declared in include/adept_b200.h).

There is no CPU fallback: if the shared library has not been built (``python -m adept_b200.build`` or
``__graft_entry__.build()``) or a call fails, a :class:`AdeptB200Error` is raised.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libadept_b200.so"

c_dp = C.c_void_p  # device pointers travel as integers
c_i, c_d, c_ll = C.c_int, C.c_double, C.c_longlong

MAX_SPECIES, MAX_DRIVERS, MAX_SUBSTEPS = 4, 8, 6


class Species(C.Structure):
    """struct adept_b200_species (include/adept_b200.h)."""

    _fields_ = [("f_in", c_dp), ("f_out", c_dp), ("f_tmp", c_dp), ("v", c_dp), ("nv", c_i), ("dv", c_d), ("k1v", c_d),
                ("charge", c_d), ("mass", c_d), ("rho_parts", c_dp), ("rho_nparts", c_i)]


class Step(C.Structure):
    """struct adept_b200_step (include/adept_b200.h); field order and types must match the header."""

    _fields_ = [
        ("batch", c_i), ("nx", c_i), ("n_species", c_i),
        ("species", Species * MAX_SPECIES),
        ("electron_species", c_i), ("collide_species", c_i), ("time_integrator", c_i), ("edfdv", c_i), ("field", c_i),
        ("dt", c_d), ("dx", c_d), ("k1x", c_d),
        ("k1x_batch", c_dp), ("ion_charge", c_dp), ("kmul", c_dp), ("kmul_stride", c_ll), ("Te", c_d),
        ("lambda_De", c_d),
        ("e_in", c_dp), ("e_out", c_dp), ("dex", c_dp), ("a", c_dp), ("prev_a", c_dp), ("djy", c_dp), ("a_out", c_dp),
        ("c_light", c_d), ("wave_on", c_i),
        ("pond", c_dp), ("rho", c_dp), ("ne_n", c_dp), ("ne_np1", c_dp),
        ("n_ex", c_i), ("ex_space", c_dp), ("ex_kx", c_dp),
        ("ex_w", c_d * MAX_DRIVERS), ("ex_a0", c_d * MAX_DRIVERS),
        ("ex_tenv", (c_d * MAX_DRIVERS) * MAX_SUBSTEPS), ("ex_wt", (c_d * MAX_DRIVERS) * MAX_SUBSTEPS),
        ("fp_on", c_i), ("krook_on", c_i), ("fp_model", c_i), ("fp_scheme", c_i), ("fp_nodrag", c_i),
        ("sg_m", c_d), ("sg_ratio", c_d),
        ("nu_fp_space", c_dp), ("nu_K_space", c_dp), ("nu_fp_time", c_d), ("nu_K_time", c_d), ("f_mx", c_dp),
        ("sync_counter", c_dp),
        ("ex_w_row", c_dp), ("ex_a0_row", c_dp), ("ex_t", c_d * MAX_SUBSTEPS),
        ("fp_sc_steps", c_i), ("fp_sc_rtol", c_d), ("fp_sc_atol", c_d),
        ("poisson_green", c_dp),
        ("diag_vlasov_dfdt", c_dp), ("diag_fp_dfdt", c_dp), ("diag_species", c_i), ("hou_li_filt", c_dp),
        ("time_row", c_dp),
    ]


class StepBwd(C.Structure):
    """struct adept_b200_step_bwd (include/adept_b200.h)."""

    _fields_ = [("f_out_bar", c_dp), ("e_out_bar", c_dp), ("f_in_bar", c_dp), ("dex_bar", c_dp), ("nu_fp_bar", c_dp),
                ("nu_K_bar", c_dp), ("scratch_f", c_dp * 3), ("scratch_row", c_dp * 2)]


class FieldPeers(C.Structure):
    """ctypes mirror of struct adept_b200_field_peers (include/adept_b200.h)."""

    _fields_ = [("n_peers", c_i), ("my_rank", c_i), ("epoch", C.c_ulonglong), ("share_in", c_dp * 8),
                ("flag_in", c_dp * 8), ("sync_counter", c_dp), ("ion_share", c_dp), ("dv", c_d), ("charge", c_d),
                ("dx", c_d), ("green", c_dp), ("rho", c_dp), ("e", c_dp), ("dex", c_dp), ("pond", c_dp),
                ("a_zero", c_dp), ("n_ex", c_i), ("ex_space", c_dp), ("ex_kx", c_dp), ("ex_w", c_d * 8),
                ("ex_a0", c_d * 8), ("ex_tenv", c_d * 8), ("ex_wt", c_d * 8)]


TIME_ROW_LEN = 104  # ADEPT_B200_TIME_ROW_LEN: tenv[6][8] | wt[6][8] | nu_fp_time | nu_K_time | ex_t[6]


# name -> argtypes; mirrors include/adept_b200.h one to one (checked by tests/test_abi.py)
SIGNATURES = {
    "adept_b200_version": [],
    "adept_b200_last_error": [],
    "adept_b200_launch_count": [],
    "adept_b200_prepare": [c_i],
    "adept_b200_profile": [c_i],
    "adept_b200_profile_report": [C.c_char_p, c_i],
    "adept_b200_vdfdx_f64": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_d, c_d, c_dp, c_dp],
    "adept_b200_vdfdx_scratch_f64": [c_dp, c_dp, c_dp, c_i, c_i, c_i, c_dp, c_d, c_d, c_dp, c_dp],
    "adept_b200_edfdv_exp_bwd_accel_f64": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_dp, c_dp, c_d, c_d, c_d, c_d, c_dp, c_dp],
    "adept_b200_moments_bwd_f64": [C.POINTER(c_dp), C.POINTER(c_d), c_i, c_i, c_i, c_dp, c_i, c_dp, c_dp],
    "adept_b200_collide_bwd_f64": [c_dp, c_dp, c_dp, c_dp, c_dp, c_i, c_i, c_i, c_dp, c_d, c_d, c_dp, c_i, c_i, c_dp],
    "adept_b200_edfdv_spline_bwd_f64": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_dp, c_dp, c_d, c_d, c_d, c_d, c_dp, c_dp, c_dp],
    "adept_b200_krook_bwd_f64": [c_dp, c_dp, c_i, c_i, c_i, c_d, c_d, c_dp, c_dp, c_dp, c_dp, c_dp],
    "adept_b200_vpush_collide_f64": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_dp, c_dp, c_d, c_d, c_d, c_d, c_dp, c_d, c_dp,
                                     c_i, c_i, c_dp],
    "adept_b200_vpush_collide_p2p_f64": [C.POINTER(c_dp), C.POINTER(c_dp), c_i, c_ll, c_i, c_i, c_dp, c_dp, c_dp, c_d,
                                         c_d, c_d, c_d, c_dp, c_d, c_dp, c_i, c_i, c_dp],
    "adept_b200_vpush_collide_p2p_staged_f64": [C.POINTER(c_dp), C.POINTER(c_dp), c_i, c_ll, c_i, c_i, c_dp, c_dp, c_dp,
                                                c_d, c_d, c_d, c_d, c_dp, c_d, c_dp, c_i, c_i, c_dp, c_dp, c_i, c_ll, c_dp],
    "adept_b200_copy2d_f64": [c_dp, c_ll, c_dp, c_ll, c_ll, c_ll, c_dp],
    "adept_b200_sum_peers_f64": [C.POINTER(c_dp), c_i, c_ll, c_dp, c_dp],
    "adept_b200_vdfdx_field_peers_f64": [c_dp, c_dp, c_i, c_i, c_dp, c_d, c_d, c_dp, c_i, C.POINTER(FieldPeers), c_dp],
    "adept_b200_save_moments_f64": [c_dp, c_dp, c_d, c_i, c_i, c_i, c_dp, c_d, c_dp, c_dp],
    "adept_b200_interp2d_f64": [c_dp, c_dp, c_d, c_i, c_i, c_dp, c_dp, c_dp, c_dp, c_i, c_i, c_dp, c_dp],
    "adept_b200_marginal_f64": [c_dp, c_dp, c_ll, c_i, c_dp, c_dp],
    "adept_b200_transpose_f64": [c_dp, c_dp, c_i, c_i, c_i, c_dp],
    "adept_b200_collide_coef_f64": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_d, c_d, c_dp, c_i, c_i, c_i, c_i, c_d, c_d,
                                    c_dp, c_dp, c_i, c_dp],
    "adept_b200_vdfdx_f32": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_d, c_d, c_dp, c_dp],
    "adept_b200_edfdv_exp_f32": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_dp, c_dp, c_d, c_d, c_d, c_d, c_dp],
    "adept_b200_collide_f32": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_d, c_d, c_dp, c_dp, c_dp, c_i, c_i, c_dp, c_dp],
    "adept_b200_abs_rfft_x_f64": [c_dp, c_dp, c_i, c_i, c_i, c_dp],
    "adept_b200_filter_x_f64": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_dp, c_dp],
    "adept_b200_vdfdx_rho_parts": [c_i, c_i, c_i],
    "adept_b200_vdfdx_rho_f64": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_d, c_d, c_dp, c_dp, c_i, c_dp],
    "adept_b200_reduce_parts_f64": [c_dp, c_i, c_ll, c_d, c_d, c_dp, c_dp, c_dp],
    "adept_b200_edfdv_exp_f64": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_dp, c_dp, c_d, c_d, c_d, c_d, c_dp],
    "adept_b200_edfdv_spline_f64": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_dp, c_dp, c_d, c_d, c_d, c_d, c_dp],
    "adept_b200_moments_f64": [c_dp, c_i, c_i, c_i, c_dp, c_d, C.POINTER(c_dp), C.POINTER(c_dp), C.POINTER(c_d), c_dp],
    "adept_b200_poisson_f64": [c_dp, c_dp, c_ll, c_dp, c_i, c_i, c_i, c_d, c_d, c_dp],
    "adept_b200_axpy_f64": [c_dp, c_dp, c_d, c_dp, c_ll, c_dp],
    "adept_b200_poisson_green_f64": [c_dp, c_dp, c_ll, c_dp, c_i, c_i, c_dp],
    "adept_b200_row_means_f64": [c_dp, c_i, c_ll, c_dp, c_dp],
    "adept_b200_field_energy_f64": [c_dp, c_dp, c_dp, c_dp, c_d, c_i, c_i, c_dp, c_dp],
    "adept_b200_ponderomotive_f64": [c_dp, c_dp, c_i, c_i, c_d, c_dp],
    "adept_b200_wave_step_f64": [c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_i, c_i, c_d, c_d, c_d, c_dp],
    "adept_b200_collide_f64": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_d, c_d, c_dp, c_dp, c_dp, c_i, c_i, c_i, c_d, c_d,
                               c_dp, c_dp],
    "adept_b200_collide_sc_f64": [c_dp, c_dp, c_i, c_i, c_i, c_dp, c_d, c_d, c_dp, c_dp, c_dp, c_i, c_i, c_i, c_d, c_d,
                                  c_dp, c_i, c_d, c_d, c_dp],
    "adept_b200_step_f64": [C.POINTER(Step), c_dp],
    "adept_b200_step_bwd_f64": [C.POINTER(Step), C.POINTER(StepBwd), c_dp],
    "adept_b200_ex_driver_f64": [c_dp, c_dp, c_i, C.POINTER(c_d), C.POINTER(c_d), C.POINTER(c_d), C.POINTER(c_d), c_ll, c_dp,
                                 c_dp],
    "adept_b200_time_row_advance": [c_dp, c_ll, c_dp, c_dp, c_dp],
}


class AdeptB200Error(RuntimeError):
    """Raised when libadept_b200.so is missing or an entry point returns a negative error code."""


_lib = None


def load() -> C.CDLL:
    """Load (once) and return the shared library, with argtypes/restype set for every exported symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise AdeptB200Error(
            f"{LIB_PATH} not found: build the CUDA extension first (python -m adept_b200.build). "
            "adept_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so and the header disagree
        fn.argtypes = argtypes
        fn.restype = {"adept_b200_last_error": C.c_char_p, "adept_b200_launch_count": c_ll}.get(name, c_i)
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().adept_b200_last_error()
        raise AdeptB200Error(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
