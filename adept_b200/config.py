"""Deck -> derived configuration for the vlasov-1d path (host side, numpy; O(nx*nv) once per run).

Mirrors what the reference does before the time loop starts:
  Grid                          adept/_vlasov1d/grid.py:37-88
  species + state construction  adept/_vlasov1d/modules.py:26-88, 190-317
  initial distribution          adept/_vlasov1d/helpers.py:37-161
The resulting ``cfg`` dict has the same keys the reference's pushers read (``cfg["grid"]["species_grids"]`` ...), so
the operator classes in :mod:`adept_b200.pushers` accept either this dict or one produced by the reference itself.
"""

from __future__ import annotations

import math
from copy import deepcopy

import numpy as np
from scipy.special import gamma

from .functions import density_profile


class Grid:
    """Configuration-space grid (x, t and their Fourier duals); attribute names follow grid.py."""

    def __init__(self, xmin, xmax, nx, tmin, tmax_requested, dt_requested, should_override_dt_for_em_waves=False,
                 beta=1.0):
        self.xmin, self.xmax, self.nx, self.tmin = float(xmin), float(xmax), int(nx), float(tmin)
        self.dx = (self.xmax - self.xmin) / self.nx
        if should_override_dt_for_em_waves:
            c_light = 1.0 / beta
            self.dt = min(dt_requested, float(0.95 * self.dx / c_light))
        else:
            self.dt = dt_requested
        self.nt = int(tmax_requested / self.dt + 1)
        self.tmax = self.dt * self.nt
        self.max_steps = min(self.nt + 4, int(1e8))
        self.x = np.linspace(self.xmin + self.dx / 2, self.xmax - self.dx / 2, self.nx)
        self.t = np.linspace(0, self.tmax, self.nt)
        self.kx = np.fft.fftfreq(self.nx, d=self.dx) * 2.0 * np.pi
        self.kxr = np.fft.rfftfreq(self.nx, d=self.dx) * 2.0 * np.pi
        self.one_over_kx = np.zeros(self.nx)
        self.one_over_kx[1:] = 1.0 / self.kx[1:]
        self.one_over_kxr = np.zeros(len(self.kxr))
        self.one_over_kxr[1:] = 1.0 / self.kxr[1:]
        self.x_a = np.concatenate([[self.x[0] - self.dx], self.x, [self.x[-1] + self.dx]])

    def asdict(self):
        return dict(self.__dict__)


def velocity_grid(vmin, vmax, nv):
    dv = (vmax - vmin) / nv
    v = np.linspace(vmin + dv / 2.0, vmax - dv / 2.0, nv)
    kv = np.fft.fftfreq(nv, d=dv) * 2.0 * np.pi
    kvr = np.fft.rfftfreq(nv, d=dv) * 2.0 * np.pi
    ookv, ookvr = np.zeros(nv), np.zeros(len(kvr))
    ookv[1:] = 1.0 / kv[1:]
    ookvr[1:] = 1.0 / kvr[1:]
    return {"v": v, "dv": dv, "nv": nv, "vmax": vmax, "vmin": vmin, "kv": kv, "kvr": kvr, "one_over_kv": ookv,
            "one_over_kvr": ookvr}


def initialize_supergaussian(nx, nv, v0, order, T0, mass, vmax, vmin, n_prof):
    """f[nx, nv] = n(x) * exp(-|(v - v0)/(alpha vth)|^m) / (sum * dv), alpha = sqrt(3 G(3/m)/G(5/m))."""
    dv = (vmax - vmin) / nv
    vax = np.linspace(vmin + dv / 2.0, vmax - dv / 2.0, nv)
    alpha = np.sqrt(3.0 * gamma(3.0 / order) / gamma(5.0 / order))
    g = np.exp(-(np.power(np.abs((vax[None, :] - v0) / (alpha * np.sqrt(T0 / mass))), order)))
    f = np.repeat(g, nx, axis=0)
    f = f / np.sum(f, axis=1)[:, None] / dv
    if n_prof.size > 1:
        f = n_prof[:, None] * f
    return f, vax


def speed_of_light_norm(units: dict | None) -> float:
    """c / v0 with v0 = sqrt(T0/m0): m0 = m_e for the electron Debye normalisation (normalization.py:100-121), or
    A m_p for ``units.reference: ion`` (ion_debye_normalization, normalization.py:124-152)."""
    if not units:
        return 1.0
    s = str(units["normalizing_temperature"]).strip()
    ref = units.get("reference", "electron")
    if ref not in ("electron", "ion"):
        raise NotImplementedError(f"adept_b200: units.reference={ref!r} (electron or ion)")
    rest_energy_eV = 510998.95 if ref == "electron" else float(units.get("A", 1.0)) * 938272088.16
    scale = 1.0
    if s.endswith("keV"):
        scale, s = 1.0e3, s[:-3]
    elif s.endswith("eV"):
        s = s[:-2]
    else:
        raise ValueError(f"cannot parse normalizing_temperature={units['normalizing_temperature']!r} (eV / keV only)")
    return 1.0 / math.sqrt(float(s) * scale / rest_energy_eV)


def species_list(cfg: dict) -> list[dict]:
    terms, gin = cfg["terms"], cfg["grid"]
    if terms.get("species"):
        out = []
        for s in terms["species"]:
            s = dict(s)
            s["vmax"] = float(s["vmax"])
            s["vmin"] = float(s["vmin"]) if s.get("vmin") is not None else -s["vmax"]
            out.append(s)
        return out
    comps = [k for k in cfg["density"].keys() if k.startswith("species-")]
    if not comps:
        raise ValueError("No density components found (expected keys starting with 'species-')")
    vmax = float(gin["vmax"])
    vmin = float(gin["vmin"]) if gin.get("vmin") is not None else -vmax
    return [{"name": "electron", "charge": -1.0, "mass": 1.0, "vmax": vmax, "vmin": vmin, "nv": gin["nv"],
             "density_components": comps}]


def build_cfg(deck: dict) -> tuple[dict, Grid]:
    """Return (cfg, grid): the deck completed with every derived quantity the operators read."""
    cfg = deepcopy(deck)
    gin = cfg["grid"]
    c_norm = speed_of_light_norm(cfg.get("units"))
    beta = 1.0 / c_norm
    has_ey = len(cfg.get("drivers", {}).get("ey", {})) > 0
    grid = Grid(gin["xmin"], gin["xmax"], gin["nx"], gin.get("tmin", 0.0), gin["tmax"], gin["dt"], has_ey, beta)
    g = {**gin, **grid.asdict(), "beta": beta}
    g["species_grids"], g["species_params"], g["species_distributions"] = {}, {}, {}
    n_total = np.zeros(grid.nx)
    species = species_list(cfg)
    for s in species:
        name, nv, mass = s["name"], int(s["nv"]), float(s["mass"])
        n_s, f_s, T0_first = np.zeros(grid.nx), np.zeros((grid.nx, nv)), None
        for cname in s["density_components"]:
            comp = cfg["density"][cname]
            nprof = np.array(density_profile(comp, grid.x))
            n_s += nprof
            tmp, _ = initialize_supergaussian(grid.nx, nv, float(comp["v0"]), float(comp.get("m", 2.0)),
                                              float(comp["T0"]), mass, s["vmax"], s["vmin"], nprof)
            f_s += tmp
            T0_first = float(comp["T0"]) if T0_first is None else T0_first
        g["species_grids"][name] = velocity_grid(s["vmin"], s["vmax"], nv)
        g["species_params"][name] = {"charge": float(s["charge"]), "mass": mass,
                                     "charge_to_mass": float(s["charge"]) / mass, "T0": T0_first}
        g["species_distributions"][name] = (n_s, f_s, g["species_grids"][name]["v"])
        n_total += n_s
    g["n_prof_total"] = n_total
    g["ion_charge"] = np.zeros_like(n_total) if len(species) > 1 else n_total.copy()
    if len(species) == 1 and "electron" in g["species_grids"]:
        for k in ("v", "kv", "kvr", "one_over_kv", "one_over_kvr"):
            g[k] = g["species_grids"]["electron"][k]
    cfg["grid"] = g
    cfg.setdefault("diagnostics", {"diag-vlasov-dfdt": False, "diag-fp-dfdt": False})
    cfg.setdefault("drivers", {"ex": {}, "ey": {}})
    return cfg, grid
