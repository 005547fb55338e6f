"""Run driver for ``solver: vlasov-1d`` decks on the B200 path.

Plays the role of ``BaseVlasov1D`` + the diffrax loop (adept/_vlasov1d/modules.py:91-358, adept/_base_.py:30-41):
the deck is completed by :func:`adept_b200.config.build_cfg`, the state dict of the reference is allocated on the
GPU, and ``y_{n+1} = VlasovMaxwell(t_n, y_n)`` is iterated.  Saves follow diffrax's semantics for the reference's
``Stepper(Euler)``: the state is linearly interpolated between y_n and y_{n+1} at each requested save time.
"""

from __future__ import annotations

import numpy as np
import torch

from .config import build_cfg
from .vector_field import VlasovMaxwell


def default_scalars(cfg, y):
    """Scalar time series saved at every step by the reference (storage.py:286-327); torch, on device."""
    g = cfg["grid"]
    s = {}
    ke = 0.0
    for name, sg in g["species_grids"].items():
        f = y[name]
        v = torch.as_tensor(sg["v"], device=f.device)[None, :]
        dv, mass = sg["dv"], g["species_params"][name]["mass"]

        def mm(inp):
            return torch.mean(torch.sum(inp, dim=-1) * dv)

        s[f"mean_P_{name}"] = mm(f * v**2.0)
        s[f"mean_j_{name}"] = mm(f * v)
        s[f"mean_n_{name}"] = mm(f)
        s[f"mean_q_{name}"] = mm(f * v**3.0)
        af = torch.abs(f)
        s[f"mean_-flogf_{name}"] = mm(-torch.log(af) * af)
        s[f"mean_f2_{name}"] = mm(f * f)
        ke = ke + 0.5 * mass * s[f"mean_P_{name}"]
    s["mean_de2"] = torch.mean(y["de"] ** 2.0)
    s["mean_e2"] = torch.mean(y["e"] ** 2.0)
    a2 = y["a"] ** 2.0
    s["mean_pond"] = torch.mean(-0.5 * (a2[2:] - a2[:-2]) / (2.0 * g["dx"]))
    s["mean_kinetic_energy"] = ke
    s["mean_field_energy"] = 0.5 * s["mean_e2"]
    s["mean_total_energy"] = ke + 0.5 * s["mean_e2"]
    return s


class Vlasov1D:
    """``sim = Vlasov1D(deck); y = sim.run(nsteps)``; ``sim.state`` holds the reference's state dict on the GPU."""

    def __init__(self, deck: dict, device="cuda"):
        if not torch.cuda.is_available():
            from ._lib import AdeptB200Error

            raise AdeptB200Error("adept_b200 needs a CUDA device: there is no CPU implementation of the time step")
        self.device = device
        self.cfg, self.grid = build_cfg(deck)
        self.vector_field = VlasovMaxwell(self.cfg, self.grid, device=device)
        self.state = self.init_state()
        self.t = 0.0
        self.step_index = 0

    def init_state(self):
        """modules.py:279-317."""
        g = self.cfg["grid"]
        dev = self.device
        state = {name: torch.as_tensor(d[1], device=dev).contiguous() for name, d in g["species_distributions"].items()}
        ref = "electron" if "electron" in state else next(iter(state))
        for k in ("e", "de"):
            state[k] = torch.zeros(g["nx"], dtype=torch.float64, device=dev)
        for k in ("a", "da", "prev_a"):
            state[k] = torch.zeros(g["nx"] + 2, dtype=torch.float64, device=dev)
        for k in ("diag-vlasov-dfdt", "diag-fp-dfdt"):
            if self.cfg["diagnostics"].get(k, False):
                state[k] = torch.zeros_like(state[ref])
        return state

    def step(self):
        self.state = self.vector_field(self.t, self.state, None)
        self.step_index += 1
        self.t = self.step_index * self.grid.dt
        return self.state

    def run(self, nsteps=None, save=None):
        """Advance ``nsteps`` (default: the deck's nt).  ``save``: name -> (times, fn(cfg, y)); returns
        (state, {name: [fn outputs]})."""
        nsteps = self.grid.nt if nsteps is None else nsteps
        save = save or {}
        out = {k: [] for k in save}
        cursor = {k: 0 for k in save}
        dt = self.grid.dt
        for _ in range(nsteps):
            t0, t1 = self.t, (self.step_index + 1) * dt
            y0 = self.state
            y1 = self.step()
            for k, (ts, fn) in save.items():
                while cursor[k] < len(ts) and ts[cursor[k]] <= t1 + 1e-12 * max(1.0, abs(t1)):
                    w = (ts[cursor[k]] - t0) / (t1 - t0)
                    yi = {kk: y0[kk] + w * (y1[kk] - y0[kk]) for kk in y1}
                    out[k].append(fn(self.cfg, yi))
                    cursor[k] += 1
        return self.state, out


def save_axis(tcfg: dict, grid) -> np.ndarray:
    """Save times of one save block (storage.py:203-219, modules.py:166-181 defaults)."""
    return np.linspace(float(tcfg.get("tmin", grid.tmin)), float(tcfg.get("tmax", grid.tmax)), int(tcfg["nt"]))
