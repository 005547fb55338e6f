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


def _lerp(y0, y1, w, key):
    return y0[key] if y1 is None else y0[key] + w * (y1[key] - y0[key])


def _species_moments(cfg, y0, y1, w, cache):
    """{name: [6, nx] tensor} = dv * sum_v {f, f v, f v^2, f v^3, -|f| log|f|, f^2} of the state interpolated between
    y0 and y1 (one fused pass per species; the interpolated f is never materialised)."""
    from . import ops

    out = {}
    for name, sg in cfg["grid"]["species_grids"].items():
        f0 = y0[name]
        key = (name, str(f0.device), len(sg["v"]), float(sg["v"][0]), float(sg["dv"]))
        if key not in cache:
            cache[key] = torch.as_tensor(np.array(sg["v"], dtype=np.float64), device=f0.device)
        out[name] = ops.save_moments(f0, cache[key], float(sg["dv"]), None if y1 is None else y1[name], w)
    return out


_V_CACHE = {}


def default_scalars(cfg, y0, y1=None, w=0.0):
    """Scalar time series saved at every step by the reference (storage.py:286-327) for the state interpolated
    linearly between y0 and y1 (weight w); device tensors."""
    from . import ops

    g = cfg["grid"]
    s = {}
    ke = 0.0
    for name, m in _species_moments(cfg, y0, y1, w, _V_CACHE).items():
        mass = g["species_params"][name]["mass"]
        mean = ops.row_means(m.reshape(6, -1))  # the jnp.mean over x of storage.py:306-323, one launch
        s[f"mean_n_{name}"], s[f"mean_j_{name}"], s[f"mean_P_{name}"] = mean[0], mean[1], mean[2]
        s[f"mean_q_{name}"], s[f"mean_-flogf_{name}"], s[f"mean_f2_{name}"] = mean[3], mean[4], mean[5]
        ke = ke + 0.5 * mass * mean[2]
    e2 = ops.field_energy(y0["e"], y0["de"], None if y1 is None else y1["e"], None if y1 is None else y1["de"], w)
    s["mean_e2"], s["mean_de2"] = e2[..., 0], e2[..., 1]
    a2 = _lerp(y0, y1, w, "a") ** 2.0
    s["mean_pond"] = torch.mean(-0.5 * (a2[..., 2:] - a2[..., :-2]) / (2.0 * g["dx"]))
    s["mean_kinetic_energy"] = ke
    s["mean_field_energy"] = 0.5 * s["mean_e2"]
    s["mean_total_energy"] = ke + 0.5 * s["mean_e2"]
    return s


def field_moments(cfg, y0, y1=None, w=0.0):
    """Per-species x-profiles n, j, v, p, q, -flogf, f^2 and the fields (storage.py:119-162).  p and q are central
    moments about the local mean velocity, obtained from the raw moments of the fused pass."""
    res = {}
    for name, m in _species_moments(cfg, y0, y1, w, _V_CACHE).items():
        n, j, m2, m3 = m[0], m[1], m[2], m[3]
        u = j / n
        res[name] = {"n": n, "j": j, "v": u, "p": m2 - 2.0 * u * j + u * u * n,
                     "q": m3 - 3.0 * u * m2 + 3.0 * u * u * j - u**3 * n, "-flogf": m[4], "f^2": m[5]}
    for k in ("e", "de", "a", "prev_a"):
        res[k] = _lerp(y0, y1, w, k)
    a2 = res["a"] ** 2.0
    res["pond"] = -0.5 * (a2[..., 2:] - a2[..., :-2]) / (2.0 * cfg["grid"]["dx"])
    return res


def dist_save(name, xax=None, vax=None, kxax=None):
    """Save function for a species' distribution (get_dist_save_func, storage.py:165-190): the full f for a {t} block,
    f interpolated linearly on the mesh ``xax x vax`` for a {t, x, v} block, or |rfft_x f| interpolated on
    ``kxax x vax`` for a {t, kx, v} block (NaN outside the grid, as interpax does with extrapolation off).  The
    spectrum block uses the one-sided axis 2 pi rfftfreq(nx, dx) that matches rfft's rows (the reference passes the
    two-sided nx-long axis, which does not fit the array: storage.py:259, 271)."""
    if xax is not None and kxax is not None:
        raise ValueError("dist_save: give xax or kxax, not both")
    if (xax is None and kxax is None) != (vax is None):
        raise ValueError("dist_save: give (xax or kxax) together with vax, or none of them")
    tables = {}

    def fn(cfg, y0, y1=None, w=0.0):
        f0 = y0[name]
        if vax is None:
            return f0.clone() if y1 is None else f0 + w * (y1[name] - f0)
        from . import ops

        key = str(f0.device)
        if key not in tables:
            sg = cfg["grid"]["species_grids"][name]
            first = cfg["grid"]["x"] if kxax is None else cfg["grid"]["kxr"]
            tables[key] = tuple(torch.as_tensor(np.array(a, dtype=np.float64), device=f0.device)
                                for a in (first, sg["v"], xax if kxax is None else kxax, vax))
        x, v, xq, vq = tables[key]
        if kxax is None:
            return ops.interp2d(f0, x, v, xq, vq, None if y1 is None else y1[name], w)
        f = f0 if y1 is None else f0 + w * (y1[name] - f0)  # the modulus is not linear: interpolate first
        return ops.interp2d(ops.abs_rfft_x(f.contiguous()), x, v, xq, vq)

    return fn


def _dim_axis(key, c):
    """_add_dim_axes (storage.py:203-219): cell-centred for x, end-point inclusive otherwise."""
    lo, hi, n = float(c[f"{key}min"]), float(c[f"{key}max"]), int(c[f"n{key}"])
    d = (hi - lo) / n if key == "x" else 0.0
    return np.linspace(lo + d / 2.0, hi - d / 2.0, n)


def save_functions(cfg, grid):
    """The deck's ``save:`` block as {key: (times, fn)} for :meth:`Vlasov1D.run` (get_save_quantities,
    storage.py:222-283): ``fields*`` -> field_moments, ``<species>: {label: {t[, x, v | kx, v]}}`` ->
    "<species>.<label>" distribution saves, dfdt diagnostics, plus the always-on ``default`` scalars at every step."""
    out = {}
    species = list(cfg["grid"]["species_grids"].keys())
    for key, sc in (cfg.get("save") or {}).items():
        if key.startswith("fields"):
            out[key] = (save_axis(sc["t"], grid), field_moments)
        elif key in species or key in ("diag-vlasov-dfdt", "diag-fp-dfdt"):
            blocks = sc.items() if key in species else [(None, sc)]
            for label, lc in blocks:
                dims = set(k for k in lc if k in ("t", "x", "v", "kx"))
                if dims == {"t"}:
                    fn = dist_save(key)
                elif dims == {"t", "x", "v"}:
                    fn = dist_save(key, xax=_dim_axis("x", lc["x"]), vax=_dim_axis("v", lc["v"]))
                elif dims == {"t", "kx", "v"}:
                    fn = dist_save(key, kxax=_dim_axis("kx", lc["kx"]), vax=_dim_axis("v", lc["v"]))
                else:
                    raise NotImplementedError(f"save block {key}/{label}: dimensions {sorted(dims)}")
                if key not in species:  # the diagnostics live on the electron grid (storage.py:262-270)
                    fn = _on_species_grid(fn, key, "electron" if "electron" in species else species[0])
                out[key if label is None else f"{key}.{label}"] = (save_axis(lc["t"], grid), fn)
        else:
            raise NotImplementedError(f"Unknown save type: {key}")
    out["default"] = (np.asarray(grid.t, dtype=np.float64), default_scalars)
    return out


def _on_species_grid(fn, key, species):
    """A distribution save function applied to a state entry that is not a species (dfdt diagnostics)."""

    def wrapped(cfg, y0, y1=None, w=0.0):
        cfg2 = dict(cfg)
        cfg2["grid"] = dict(cfg["grid"])
        cfg2["grid"]["species_grids"] = {key: cfg["grid"]["species_grids"][species]}
        return fn(cfg2, y0, y1, w)

    return wrapped


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

    def graph_stepper(self, max_steps: int) -> "GraphStepper":
        """CUDA-graph replay of the next ``max_steps`` steps (see GraphStepper); the first call costs one eager step."""
        return GraphStepper(self, max_steps)

    def graph_steps_per_second(self, nsteps: int) -> float:
        """Steps per second of graph replay over ``nsteps`` (even) steps, timed with CUDA events (bench helper)."""
        nsteps += nsteps % 2
        gs = GraphStepper(self, 2 * nsteps + 2)
        gs.run(nsteps)  # warm replay
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gs.run(nsteps)
        e1.record()
        torch.cuda.synchronize()
        return nsteps / (e0.elapsed_time(e1) * 1e-3)

    def run(self, nsteps=None, save=None):
        """Advance ``nsteps`` (default: the deck's nt).  ``save``: name -> (times, fn(cfg, y0, y1, w)) where the
        saved state is y0 + w (y1 - y0) (diffrax's linear dense output; see default_scalars / field_moments); returns
        (state, {name: [fn outputs]})."""
        nsteps = self.grid.nt if nsteps is None else nsteps
        if save == "deck":  # the deck's own save: block, as the reference wires it (storage.py:222-283)
            save = save_functions(self.cfg, self.grid)
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
                    out[k].append(fn(self.cfg, y0, y1, w))  # save functions interpolate on the fly
                    cursor[k] += 1
        return self.state, out


class GraphStepper:
    """CUDA-graph replay of the native time step for decks whose host-issue time dominates (small grids: C1, C2, C5).

    One replay advances TWO steps (state buffers A -> B -> A), each preceded by ``adept_b200_time_row_advance``: the
    time factors of the drivers and collision profiles of every step of the run live in a device table, the kernels
    read the current row through ``adept_b200_step.time_row``, so the same captured graph serves every step.  Decks
    with Ey drivers or a live transverse field are refused (their per-step inputs are evaluated by host code)."""

    def __init__(self, sim: "Vlasov1D", max_steps: int):
        import ctypes as C

        from . import _lib

        vm = sim.vector_field
        if vm.native is None or vm.has_ey:
            raise NotImplementedError("GraphStepper: needs the native step and no Ey driver")
        y = sim.state
        if bool(torch.any(y["a"] != 0)) or bool(torch.any(y["prev_a"] != 0)):
            raise NotImplementedError("GraphStepper: the transverse field must be zero (wave update skipped)")
        vm._a_live = False
        self.sim, self.lib, self.nat = sim, _lib.load(), vm.native
        nat = self.nat
        dev = y[nat.names[0]].device
        f0 = y[nat.names[0]]
        batch = f0.shape[0] if f0.dim() == 3 else 1
        sim.step()  # one eager step: builds the static descriptor, twiddle tables and kernel attributes
        y = sim.state
        static = nat._static_step(y, dev, batch, False)
        nsub = len(nat.dt_array)
        # ping-pong state: every tensor of the state dict twice
        self.buf = []
        for _ in range(2):
            b = {k: torch.empty_like(v) for k, v in y.items() if k not in ("a", "prev_a", "da", "de")}
            b["dex"] = torch.zeros((nsub,) + tuple(y["e"].shape), dtype=torch.float64, device=dev)
            self.buf.append(b)
        for k, v in y.items():
            if k in self.buf[0]:
                self.buf[0][k].copy_(v)
        self.fixed = {"a": y["a"], "prev_a": y["a"], "da": vm._zeros_a}
        self.row = torch.zeros(_lib.TIME_ROW_LEN, dtype=torch.float64, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.max_steps = int(max_steps)
        t0, dt = sim.step_index, sim.grid.dt
        self.first_step = t0
        self.table = torch.as_tensor(np.stack([nat.time_row((t0 + k) * dt) for k in range(self.max_steps)]), device=dev)
        self.steps_done = 0
        self.sts = []
        for src, dst in ((0, 1), (1, 0)):
            st = _lib.Step.from_buffer_copy(static)
            a, b = self.buf[src], self.buf[dst]
            for k, name in enumerate(nat.names):
                st.species[k].f_in, st.species[k].f_out = a[name].data_ptr(), b[name].data_ptr()
            st.e_in, st.e_out, st.dex = a["e"].data_ptr(), b["e"].data_ptr(), b["dex"].data_ptr()
            st.a, st.prev_a = y["a"].data_ptr(), y["prev_a"].data_ptr()
            st.wave_on = 0
            st.diag_vlasov_dfdt = b["diag-vlasov-dfdt"].data_ptr() if "diag-vlasov-dfdt" in b else None
            st.diag_fp_dfdt = b["diag-fp-dfdt"].data_ptr() if "diag-fp-dfdt" in b else None
            st.time_row = self.row.data_ptr()
            self.sts.append(st)
        self._C = C
        # eager pass of both descriptors (warms every launch path with the device time row), then rewind
        saved = {k: v.clone() for k, v in self.buf[0].items()}
        self._enqueue_pair()
        torch.cuda.synchronize()
        for k, v in saved.items():
            self.buf[0][k].copy_(v)
        self.counter.zero_()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._enqueue_pair()
        self.counter.zero_()

    def _enqueue_pair(self):
        from . import _lib

        C, lib = self._C, self.lib
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        before = lib.adept_b200_launch_count()
        for st in self.sts:
            _lib.check(lib.adept_b200_time_row_advance(self.table.data_ptr(), self.max_steps, self.counter.data_ptr(),
                                                       self.row.data_ptr(), stream), "time_row_advance")
            _lib.check(lib.adept_b200_step_f64(C.byref(st), stream), "step")
        self.launches_per_replay = lib.adept_b200_launch_count() - before

    def run(self, nsteps: int):
        """Advance an even number of steps by graph replay; returns the state dict (views of the A buffers)."""
        if nsteps % 2 or nsteps < 0 or self.steps_done + nsteps > self.max_steps:
            raise ValueError("GraphStepper.run: nsteps must be even and within the table built at construction")
        for _ in range(nsteps // 2):
            self.graph.replay()
        self.steps_done += nsteps
        sim, a = self.sim, self.buf[0]
        nat = self.nat
        state = {k: v for k, v in a.items() if k != "dex"}
        state["de"] = a["dex"][sim.vector_field.vpfp.dex_save]
        state.update(self.fixed)
        sim.state = state
        sim.step_index = self.first_step + self.steps_done
        sim.t = sim.step_index * sim.grid.dt
        from . import ops

        ops.LAUNCHES += self.launches_per_replay * (nsteps // 2)
        return state


def save_axis(tcfg: dict, grid) -> np.ndarray:
    """Save times of one save block (storage.py:203-219, modules.py:166-181 defaults)."""
    return np.linspace(float(tcfg.get("tmin", grid.tmin)), float(tcfg.get("tmax", grid.tmax)), int(tcfg["nt"]))
