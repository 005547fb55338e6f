"""Host-side 1-D profile functions of the vlasov-1d deck (O(nx) work, evaluated with numpy on the host).

Mirrors the reference's ``adept/functions.py`` (EnvelopeFunction :46-103, SpaceTimeEnvelopeFunction :106-135) and
the density bases of ``adept/_vlasov1d/simulation.py:205-270``.  Only dimensionless decks are supported here (the
reference resolves unit strings such as "100um" with pint, which is not part of this hot path).
"""

from __future__ import annotations

import numpy as np


def _num(x, what):
    if isinstance(x, (int, float)) and not isinstance(x, bool):
        return float(x)
    raise ValueError(f"{what}={x!r}: adept_b200 only accepts dimensionless numbers (no unit strings)")


class EnvelopeFunction:
    """baseline + bump_height * 0.5 (tanh((x-l)/rise) - tanh((x-r)/rise)), optionally inverted (trough)."""

    def __init__(self, center, width, rise, baseline=0.0, bump_height=1.0, is_trough=False):
        self.center, self.width, self.rise = _num(center, "center"), _num(width, "width"), _num(rise, "rise")
        self.baseline, self.bump_height, self.is_trough = float(baseline), float(bump_height), bool(is_trough)

    @staticmethod
    def from_config(cfg: dict) -> "EnvelopeFunction":
        return EnvelopeFunction(
            cfg["center"], cfg["width"], cfg["rise"], cfg.get("baseline", 0.0), cfg.get("bump_height", 1.0),
            cfg.get("bump_or_trough", "bump") == "trough",
        )

    def __call__(self, x):
        left = self.center - self.width * 0.5
        right = self.center + self.width * 0.5
        env = 0.5 * (np.tanh((x - left) / self.rise) - np.tanh((x - right) / self.rise))
        if self.is_trough:
            env = 1 - env
        return self.baseline + self.bump_height * env


class SpaceTimeEnvelopeFunction:
    """time_envelope(t) * space_envelope(x)."""

    def __init__(self, time_envelope: EnvelopeFunction, space_envelope: EnvelopeFunction):
        self.time_envelope, self.space_envelope = time_envelope, space_envelope

    @staticmethod
    def from_config(cfg: dict) -> "SpaceTimeEnvelopeFunction":
        return SpaceTimeEnvelopeFunction(
            EnvelopeFunction.from_config(cfg["time"]), EnvelopeFunction.from_config(cfg["space"])
        )

    def __call__(self, x, t):
        return self.time_envelope(t) * self.space_envelope(x)


def density_profile(comp: dict, x: np.ndarray) -> np.ndarray:
    """Density component n(x) for basis uniform / sine / tanh / linear / exponential."""
    basis = comp["basis"]
    if basis == "uniform":
        base = comp.get("baseline")
        prof = (float(base) if base is not None else 1.0) * np.ones_like(x)
    elif basis == "sine":
        prof = float(comp["baseline"]) * (1.0 + float(comp["amplitude"]) * np.sin(float(comp["wavenumber"]) * x))
    elif basis == "tanh":
        prof = EnvelopeFunction.from_config(comp)(x) * np.ones_like(x)
    elif basis in ("linear", "exponential"):
        center = _num(comp["center"], "center")
        L = _num(comp["gradient scale length"], "gradient scale length")
        val = _num(comp["val at center"], "val at center")
        dens = val + (x - center) / L if basis == "linear" else val * np.exp((x - center) / L)
        prof = EnvelopeFunction.from_config(comp)(x) * dens
    else:
        raise NotImplementedError(f"Unknown density basis: {basis}")
    if float(comp.get("noise_val", 0.0)) != 0.0:
        raise NotImplementedError(
            "density noise uses jax.random in the reference; supply the perturbed profile explicitly instead"
        )
    return prof
