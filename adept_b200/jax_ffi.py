"""EXPERIMENTAL (never imported here: no jax in the build image).  JAX host side of the drop-in: ``jax.ffi`` custom calls into ``libadept_b200_xla.so`` (csrc/xla_ffi.cc) paired with
their adjoints in ``jax.custom_vjp``, and pusher classes with the constructor and call signatures of the reference
(adept/_vlasov1d/solvers/pushers/vlasov.py:63-251, fokker_planck.py:272-443) that ``B200Vlasov1D`` installs into
``VlasovMaxwell``.

jax is NOT installed in the B200 build image, so nothing in the tests, the bench or ``smoke()`` imports this module:
it is the binding a maintainer uses on a machine with ``jax[cuda12]`` and is unverified here.  The arithmetic it
reaches is the C ABI that the torch-hosted mirror (``adept_b200/pushers.py``, ``adept_b200/autodiff.py``) exercises in
the GPU tests, including the finite-difference checks of every adjoint.
"""

from __future__ import annotations

import ctypes
from pathlib import Path

try:
    import jax
    import jax.numpy as jnp
except ImportError as exc:  # fail loudly: there is no fallback path
    raise ImportError("adept_b200.jax_ffi needs jax[cuda12]; the torch-hosted path is adept_b200.pushers") from exc

_LIB = Path(__file__).resolve().parent / "libadept_b200_xla.so"
_TARGETS = ("vdfdx", "vdfdx_rho", "reduce_parts", "edfdv_exp", "edfdv_exp_bwd_accel", "edfdv_spline", "poisson",
            "collide", "collide_bwd", "vpush_collide", "save_moments")


def register() -> None:
    """Register every handler of libadept_b200_xla.so as a CUDA FFI target (python -m adept_b200.build --xla first)."""
    if not _LIB.exists():
        raise FileNotFoundError(f"{_LIB} not found: run `python -m adept_b200.build --xla` on a machine with jax")
    lib = ctypes.CDLL(str(_LIB))
    for name in _TARGETS:
        jax.ffi.register_ffi_target(f"adept_b200_{name}", jax.ffi.pycapsule(getattr(lib, f"adept_b200_xla_{name}")),
                                    platform="CUDA")


def _call(name, out_types, *arrays, **attrs):
    return jax.ffi.ffi_call(f"adept_b200_{name}", out_types, vmap_method="broadcast_all")(*arrays, **attrs)


def _like(x):
    return jax.ShapeDtypeStruct(x.shape, x.dtype)


# ---- x-advection: real circulant with a unit-modulus symbol, adjoint = the same push with dt -> -dt ----------------
def _make_vdfdx(dt: float, k1x: float):
    @jax.custom_vjp
    def vdfdx(f, v):
        return _call("vdfdx", _like(f), f, v, dt=dt, k1x=k1x)

    def fwd(f, v):
        return vdfdx(f, v), v

    def bwd(v, g):
        return _call("vdfdx", _like(g), g, v, dt=-dt, k1x=k1x), jnp.zeros_like(v)

    vdfdx.defvjp(fwd, bwd)
    return vdfdx


# ---- v-advection: adjoint w.r.t. f is the push with dt -> -dt; w.r.t. e / dex / pond through the acceleration ----------
def _make_edfdv(charge: float, mass: float, dt: float, k1v: float):
    @jax.custom_vjp
    def edfdv(f, e, dex, pond):
        return _call("edfdv_exp", _like(f), f, e, dex, pond, charge=charge, mass=mass, dt=dt, k1v=k1v)

    def fwd(f, e, dex, pond):
        return edfdv(f, e, dex, pond), (f, e, dex, pond)

    def bwd(res, g):
        f, e, dex, pond = res
        f_bar = _call("edfdv_exp", _like(g), g, e, dex, pond, charge=charge, mass=mass, dt=-dt, k1v=k1v)
        a_bar = _call("edfdv_exp_bwd_accel", _like(e), f, g, e, dex, pond, charge=charge, mass=mass, dt=dt, k1v=k1v)
        e_bar = a_bar * (charge / mass)
        return f_bar, e_bar, e_bar, a_bar * (charge * charge / (mass * mass))

    edfdv.defvjp(fwd, bwd)
    return edfdv


# ---- Fokker-Planck step (central differencing, LB / Dougherty) ------------------------------------------------------
def _make_collide(dv: float, dt: float, model: int, scheme: int):
    attrs = dict(dv=dv, dt=dt, model=model, scheme=scheme)

    @jax.custom_vjp
    def collide(f, v, nu_fp):
        zero = jnp.zeros_like(nu_fp)
        return _call("collide", _like(f), f, v, nu_fp, zero, jnp.zeros_like(v), nodrag=0, fp_on=1, krook=0, sg_m=2.0,
                     sg_ratio=0.5, **attrs)

    def fwd(f, v, nu_fp):
        out = collide(f, v, nu_fp)
        return out, (f, out, v, nu_fp)

    def bwd(res, g):
        f, out, v, nu_fp = res
        f_bar, nu_bar = _call("collide_bwd", (_like(f), _like(nu_fp)), f, out, g, v, nu_fp, **attrs)
        return f_bar, jnp.zeros_like(v), nu_bar

    collide.defvjp(fwd, bwd)
    return collide


# ---- pusher objects with the reference's signatures -----------------------------------------------------------------
class SpaceExponential:
    """Drop-in for pushers/vlasov.py:223-251: ``vdfdx(f_dict, dt) -> f_dict``."""

    def __init__(self, x, species_grids, parallel=False):
        self.species_grids = species_grids
        self.k1x = float(2.0 * jnp.pi / ((x[1] - x[0]) * len(x)))
        self._ops = {}

    def __call__(self, f_dict, dt):
        out = {}
        for name, f in f_dict.items():
            key = float(dt)
            if key not in self._ops:
                self._ops[key] = _make_vdfdx(key, self.k1x)
            out[name] = self._ops[key](f, jnp.asarray(self.species_grids[name]["v"]))
        return out


class VelocityExponential:
    """Drop-in for pushers/vlasov.py:63-91: ``edfdv(f_dict, e, pond, dt) -> f_dict``."""

    def __init__(self, species_grids, species_params, parallel=False):
        self.species_grids, self.species_params = species_grids, species_params
        self._ops = {}

    def __call__(self, f_dict, e, pond, dt):
        out = {}
        for name, f in f_dict.items():
            sp = self.species_params[name]
            key = (name, float(dt))
            if key not in self._ops:
                self._ops[key] = _make_edfdv(float(sp["charge"]), float(sp["mass"]), float(dt),
                                             float(self.species_grids[name]["kvr"][1]))
            out[name] = self._ops[key](f, e, jnp.zeros_like(e), pond)
        return out
