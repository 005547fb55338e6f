"""``solver: vlasov-1d2v`` on the B200 path: f(x, v_par, v_perp) in cylindrical velocity geometry.

Mirrors adept/_vlasov1d2v/: ``VelocityExponential2V`` / ``SpaceExponential2V`` (solvers/pushers/vlasov.py:12-56),
``Collisions`` with marginal-moment coefficients (solvers/pushers/fokker_planck.py:31-140), the integrators,
``VlasovPoissonFokkerPlanck`` and ``VlasovMaxwell2V`` (solvers/vector_field.py:19-239), ``BaseVlasov1D2V``'s derived
quantities and state (modules.py:44-146).  Same class names, constructor dicts and call signatures; tensors are float64
CUDA tensors ``[nx, nv, nvperp]``.

Every operator is one of the library's kernels -- the perpendicular axis is a spectator, so the pencil kernels of the 1-D
solver serve the extra axis unchanged:

* x-advection: the x-pencil kernel on ``[nx, nv * nvperp]`` with the velocity table repeated along v_perp;
* v_par-advection: the same strided-pencil kernel with one ``[nv, nvperp]`` member per x: the transform runs along nv,
  the per-member wavenumber is ``kv_1 * accel[x]`` and the "velocity" table is ones, so the phase of mode m is
  ``m kv_1 accel[x] dt`` for every v_perp column;
* field solve: the 1-D ``ElectricFieldSolver`` fed with the marginals (``adept_b200_marginal_f64``);
* collisions: ``adept_b200_collide_coef_f64`` twice -- on the marginal rows (records vbar, beta after the self-consistent
  Newton refinement), then on the ``nx * nvperp`` slice rows laid out contiguously by ``adept_b200_transpose_f64``.

Not built: the ``cylindrical_landau`` operator (fokker_planck.py:143-403) -- raises NotImplementedError.
"""

from __future__ import annotations

import numpy as np
import torch

from . import ops, pushers
from .config import build_cfg
from .functions import SpaceTimeEnvelopeFunction


def perp_grid(nvperp: int, vperp_max: float):
    """helpers.py:20-29."""
    dvperp = vperp_max / nvperp
    vperp = np.linspace(dvperp / 2.0, vperp_max - dvperp / 2.0, nvperp)
    return vperp, dvperp, 2.0 * np.pi * vperp * dvperp


def build_cfg_2v(deck: dict):
    """(cfg, grid) with the 2V quantities added (modules.py:44-118, helpers.py:32-91): per species ``vperp, dvperp,
    nvperp, vperp_max, wperp`` and f = F_1d(x, v_par) M(v_perp), M a Maxwellian at each component's T0 normalised to
    sum_j M_j w_j = 1."""
    cfg, grid = build_cfg(deck)
    g = cfg["grid"]
    nvperp, vperp_max = int(g["nvperp"]), float(g["vperp_max"])
    vperp, dvperp, wperp = perp_grid(nvperp, vperp_max)
    from .config import density_profile, initialize_supergaussian, species_list

    dists = {}
    for s in species_list(cfg):
        name, nv, mass = s["name"], int(s["nv"]), float(s["mass"])
        f_s = np.zeros((grid.nx, nv, nvperp))
        for cname in s["density_components"]:
            comp = cfg["density"][cname]
            nprof = np.array(density_profile(comp, grid.x))
            tmp, _ = initialize_supergaussian(grid.nx, nv, float(comp["v0"]), float(comp.get("m", 2.0)),
                                              float(comp["T0"]), mass, s["vmax"], s["vmin"], nprof)
            m_perp = np.exp(-(vperp**2.0) / (2.0 * float(comp["T0"]) / mass))
            m_perp = m_perp / np.sum(m_perp * wperp)
            f_s += tmp[:, :, None] * m_perp[None, None, :]
        n_s, _, v_ax = g["species_distributions"][name]
        dists[name] = (n_s, f_s, v_ax, vperp)
        g["species_grids"][name].update(vperp=vperp, dvperp=dvperp, nvperp=nvperp, vperp_max=vperp_max, wperp=wperp)
    g["species_distributions"] = dists
    d = cfg["diagnostics"]
    d.setdefault("diag-vlasov-cumulative", False)
    d.setdefault("diag-fp-cumulative", False)
    return cfg, grid


class _Tables:
    def __init__(self):
        self.t = {}

    def get(self, key, make, device):
        k = (key, str(device))
        if k not in self.t:
            self.t[k] = torch.as_tensor(np.ascontiguousarray(make(), dtype=np.float64), device=device)
        return self.t[k]


class SpaceExponential2V:
    """pushers/vlasov.py:41-56: x-advection of every species, v_perp a spectator."""

    def __init__(self, x, species_grids):
        self.k1x = float(2.0 * np.pi / (len(x) * (x[1] - x[0])))
        self.species_grids = species_grids
        self._tab = _Tables()

    def __call__(self, f_dict, dt):
        out = {}
        for name, f in f_dict.items():
            sg = self.species_grids[name]
            nx, nv, npp = f.shape
            vrep = self._tab.get(("vrep", name), lambda sg=sg, npp=npp: np.repeat(np.asarray(sg["v"]), npp), f.device)
            out[name] = ops.vdfdx(f.reshape(nx, nv * npp), vrep, dt, self.k1x).reshape(nx, nv, npp)
        return out


class VelocityExponential2V:
    """pushers/vlasov.py:12-37: spectral v_par-advection under the electric and ponderomotive forces."""

    def __init__(self, species_grids, species_params):
        self.species_grids, self.species_params = species_grids, species_params
        self._tab = _Tables()

    def __call__(self, f_dict, e, pond, dt):
        out = {}
        for name, f in f_dict.items():
            q, m = self.species_params[name]["charge"], self.species_params[name]["mass"]
            k1v = float(self.species_grids[name]["kvr"][1])
            force = q * e + (q**2 / m) * pond
            accel = force / m
            ones = self._tab.get(("ones", f.shape[-1]), lambda n=f.shape[-1]: np.ones(n), f.device)
            # one [nv, nvperp] member per x: transform along nv, phase increment of member x = kv_1 accel[x] dt
            out[name] = ops.vdfdx(f, ones, dt, 0.0, k1x_batch=(accel * k1v).contiguous())
        return out


class Collisions:
    """pushers/fokker_planck.py:31-140: v_par drift-diffusion collisions with marginal-moment coefficients."""

    def __init__(self, cfg):
        self.cfg = cfg
        fp_type = cfg["terms"]["fokker_planck"]["type"].casefold()
        if fp_type == "cylindrical_landau":
            raise NotImplementedError("adept_b200: the cylindrical_landau operator of vlasov-1d2v is not built")
        if fp_type not in ("dougherty", "dougherty_nodrag", "lenard_bernstein"):
            raise NotImplementedError(f"Unknown Fokker-Planck type for vlasov-1d2v: {fp_type}")
        if cfg["terms"]["krook"]["is_on"]:
            raise NotImplementedError("Krook is not implemented for vlasov-1d2v")
        self.c1 = pushers.Collisions(cfg)  # model / scheme / nodrag / self-consistent-beta controls, v, dv
        self.wperp = np.asarray(cfg["grid"]["species_grids"]["electron"]["wperp"], dtype=np.float64)
        self._tab = _Tables()

    def marginal(self, f):
        return ops.marginal(f, self._tab.get("wperp", lambda: self.wperp, f.device))

    def __call__(self, nu_fp, nu_K, f, dt):
        if isinstance(f, dict):
            return {k: (self._apply(nu_fp, fs, dt) if k == "electron" else fs) for k, fs in f.items()}
        return self._apply(nu_fp, f, dt)

    def _apply(self, nu_fp, f, dt):
        if not self.cfg["terms"]["fokker_planck"]["is_on"]:
            return f
        c1 = self.c1
        nx, nv, npp = f.shape
        dev = f.device
        nu = nu_fp if nu_fp is not None else torch.zeros(nx, dtype=torch.float64, device=dev)
        v = self._tab.get("v", lambda: c1.v, dev)
        kw = dict(model=c1.model, scheme=c1.scheme, nodrag=c1.nodrag, sc_steps=c1.sc_steps, sc_rtol=c1.sc_rtol,
                  sc_atol=c1.sc_atol)
        F = self.marginal(f)
        coef = torch.empty((nx, 2), dtype=torch.float64, device=dev)
        ops.collide_coef(F, v, c1.dv, float(dt), nu, coef_out=coef, **kw)  # (vbar, beta) of every marginal row
        ft = ops.transpose_last2(f)                                          # [nx, nvperp, nv]
        ft_new = ops.collide_coef(ft.reshape(nx * npp, nv), v, c1.dv, float(dt), nu, coef_in=coef, coef_div=npp, **kw)
        return ops.transpose_last2(ft_new.reshape(nx, npp, nv))


class TimeIntegrator:
    """vector_field.py:19-38: the 1-D field solver fed with marginals, the 2V pushers."""

    def __init__(self, cfg, grid):
        self.field_solve = pushers.ElectricFieldSolver(cfg, grid)
        self.species_grids = cfg["grid"]["species_grids"]
        self.species_params = cfg["grid"]["species_params"]
        if cfg["terms"]["edfdv"] != "exponential":
            raise NotImplementedError("vlasov-1d2v supports edfdv: exponential only")
        self.edfdv = VelocityExponential2V(self.species_grids, self.species_params)
        self.vdfdx = SpaceExponential2V(grid.x, self.species_grids)
        self._tab = _Tables()

    def marginals(self, f_dict):
        return {name: ops.marginal(f, self._tab.get(("w", name), lambda n=name: self.species_grids[n]["wperp"], f.device))
                for name, f in f_dict.items()}


class LeapfrogIntegrator(TimeIntegrator):
    """vector_field.py:41-58."""

    def __init__(self, cfg, grid):
        super().__init__(cfg, grid)
        self.dt = grid.dt
        self.dt_array = self.dt * np.array([0.0, 1.0])

    def __call__(self, f_dict, a, dex_array, prev_ex):
        f_after_v = self.vdfdx(f_dict, dt=self.dt)
        f_for_field = f_dict if self.field_solve.hampere else f_after_v
        if self.field_solve.hampere:
            raise NotImplementedError("adept_b200 vlasov-1d2v: field = hampere")
        pond, e = self.field_solve(f_dict=self.marginals(f_for_field), a=a, prev_ex=prev_ex, dt=self.dt)
        return e, self.edfdv(f_after_v, e=e + dex_array[0], pond=pond, dt=self.dt)


class SixthOrderHamIntegrator(TimeIntegrator):
    """vector_field.py:61-113."""

    def __init__(self, cfg, grid):
        super().__init__(cfg, grid)
        from .vector_field import SixthOrderHamIntegrator as S1

        self.dt = grid.dt
        s = S1(_cfg_1d(cfg), grid)  # coefficients exactly as the 1-D integrator computes them
        self.a1, self.a2, self.a3, self.D1, self.D2, self.D3, self.dt_array = s.a1, s.a2, s.a3, s.D1, s.D2, s.D3, s.dt_array

    def __call__(self, f_dict, a, dex_array, prev_ex):
        drifts = [self.a1, self.a2, self.a3, self.a2, self.a1]
        kicks = [self.D1, self.D2, self.D3, self.D3, self.D2, self.D1]
        pond, e = self.field_solve(f_dict=self.marginals(f_dict), a=a, prev_ex=None, dt=None)
        f_dict = self.edfdv(f_dict, e=dex_array[0] + e, pond=pond, dt=kicks[0] * self.dt)
        for i, drift in enumerate(drifts):
            f_dict = self.vdfdx(f_dict, dt=drift * self.dt)
            pond, e = self.field_solve(f_dict=self.marginals(f_dict), a=a, prev_ex=None, dt=None)
            f_dict = self.edfdv(f_dict, e=dex_array[i + 1] + e, pond=pond, dt=kicks[i + 1] * self.dt)
        return e, f_dict


def _cfg_1d(cfg):
    c = dict(cfg)
    c["diagnostics"] = {"diag-vlasov-dfdt": False, "diag-fp-dfdt": False}
    return c


class VlasovPoissonFokkerPlanck:
    """vector_field.py:116-172: integrator -> collisions -> per-step increments of the marginal diagnostics."""

    def __init__(self, cfg, grid):
        self.dt = grid.dt
        if cfg["terms"]["time"] == "sixth":
            self.vlasov_poisson, self.dex_save = SixthOrderHamIntegrator(cfg, grid), 3
        elif cfg["terms"]["time"] == "leapfrog":
            self.vlasov_poisson, self.dex_save = LeapfrogIntegrator(cfg, grid), 0
        else:
            raise NotImplementedError
        self.fp = Collisions(cfg)
        self.vlasov_cumulative = cfg["diagnostics"]["diag-vlasov-cumulative"]
        self.fp_cumulative = cfg["diagnostics"]["diag-fp-cumulative"]

    def __call__(self, f_dict, a, prev_ex, dex_array, nu_fp, nu_K):
        e, f_vlasov = self.vlasov_poisson(f_dict, a, dex_array, prev_ex)
        f_fp = self.fp(nu_fp, nu_K, f_vlasov, dt=self.dt)
        diags = {}
        ref = "electron" if "electron" in f_dict else next(iter(f_dict))
        if self.vlasov_cumulative or self.fp_cumulative:
            marg = self.vlasov_poisson.marginals
            m_v = marg({ref: f_vlasov[ref]})[ref]
            if self.vlasov_cumulative:
                diags["diag-vlasov-cumulative"] = m_v - marg({ref: f_dict[ref]})[ref]
            if self.fp_cumulative:
                diags["diag-fp-cumulative"] = marg({ref: f_fp[ref]})[ref] - m_v
        return e, f_fp, diags


class VlasovMaxwell2V:
    """vector_field.py:175-239: one full vlasov-1d2v step y -> y'."""

    def __init__(self, cfg, grid, drivers=None, nu_fp_prof=None, nu_K_prof=None, device="cuda"):
        self.cfg, self.grid, self.device = cfg, grid, device
        self.vpfp = VlasovPoissonFokkerPlanck(cfg, grid)
        c = 1.0 / cfg["grid"]["beta"]
        self.wave_solver = pushers.WaveSolver(c=c, dx=grid.dx, dt=grid.dt)
        self.dt = grid.dt
        dcfg = cfg.get("drivers", {"ex": {}, "ey": {}})
        if drivers is None:
            ex = [pushers.EMDriver.from_config(d, c) for d in dcfg.get("ex", {}).values()]
            ey = [pushers.EMDriver.from_config(d, c) for d in dcfg.get("ey", {}).values()]
        else:
            ex, ey = drivers["ex"], drivers["ey"]
        self.ey_driver = pushers.TransverseCurrentSourceDriver(grid.x_a, drivers=ey, c=c, device=device)
        self.ex_driver = pushers.LongitudinalElectricFieldDriver(grid.x, drivers=ex, device=device)
        fpc = cfg["terms"]["fokker_planck"]
        self.fp_on = bool(fpc["is_on"])
        self.nu_fp_prof = nu_fp_prof if nu_fp_prof is not None else (
            SpaceTimeEnvelopeFunction.from_config(fpc) if self.fp_on else None)
        self._x = np.asarray(grid.x)
        self._tab = _Tables()

    def compute_electron_charge_density(self, f_dict):
        if "electron" not in f_dict:
            return torch.zeros(self.grid.nx, dtype=torch.float64, device=self.device)
        sg = self.cfg["grid"]["species_grids"]["electron"]
        q = self.cfg["grid"]["species_params"]["electron"]["charge"]
        F = ops.marginal(f_dict["electron"], self._tab.get("w", lambda: sg["wperp"], f_dict["electron"].device))
        out = torch.empty(F.shape[:-1], dtype=torch.float64, device=F.device)
        ops.moments(F, None, float(sg["dv"]), (out, None, None), scale_b=(q, 1.0, 1.0))
        return out

    def __call__(self, t, y, args=None):
        t = float(t)
        dt_array = self.vpfp.vlasov_poisson.dt_array
        dex = [self.ex_driver(t + float(d), args) for d in dt_array]
        djy = self.ey_driver(t + float(dt_array[1]), args)
        nu_fp = None
        if self.fp_on:
            nu_fp = self._tab.get("nu_space", lambda: self.nu_fp_prof.space_envelope(self._x) * np.ones_like(self._x),
                                  self.device) * float(self.nu_fp_prof.time_envelope(t))
        f_dict = {k: v for k, v in y.items() if k in self.cfg["grid"]["species_grids"]}
        ne_n = self.compute_electron_charge_density(f_dict)
        e, f_new, diags = self.vpfp(f_dict=f_dict, a=y["a"], prev_ex=y["e"], dex_array=dex, nu_fp=nu_fp, nu_K=None)
        ne_np1 = self.compute_electron_charge_density(f_new)
        a = self.wave_solver(a=y["a"], aold=y["prev_a"], djy_array=djy, electron_density_n=ne_n,
                             electron_density_np1=ne_np1)
        result = {"a": a["a"], "prev_a": a["prev_a"], "da": djy, "de": dex[self.vpfp.dex_save], "e": e}
        result.update(f_new)
        for key, inc in diags.items():  # running time integrals of the per-step increments
            result[key] = y[key] + inc
        return result


class Vlasov1D2V:
    """``sim = Vlasov1D2V(deck); sim.run(nsteps)``: BaseVlasov1D2V + the fixed-step loop (modules.py:26-160)."""

    def __init__(self, deck: dict, device="cuda"):
        if not torch.cuda.is_available():
            from ._lib import AdeptB200Error

            raise AdeptB200Error("adept_b200 needs a CUDA device: there is no CPU implementation of the time step")
        self.device = device
        self.cfg, self.grid = build_cfg_2v(deck)
        self.vector_field = VlasovMaxwell2V(self.cfg, self.grid, device=device)
        self.state = self.init_state()
        self.t, self.step_index = 0.0, 0

    def init_state(self):
        g, dev = self.cfg["grid"], self.device
        state = {name: torch.as_tensor(d[1], device=dev).contiguous() for name, d in g["species_distributions"].items()}
        ref = "electron" if "electron" in state else next(iter(state))
        for k in ("e", "de"):
            state[k] = torch.zeros(g["nx"], dtype=torch.float64, device=dev)
        for k in ("a", "da", "prev_a"):
            state[k] = torch.zeros(g["nx"] + 2, dtype=torch.float64, device=dev)
        nv = state[ref].shape[1]
        for k in ("diag-vlasov-cumulative", "diag-fp-cumulative"):
            if self.cfg["diagnostics"].get(k, False):
                state[k] = torch.zeros((g["nx"], nv), dtype=torch.float64, device=dev)
        return state

    def step(self):
        self.state = self.vector_field(self.t, self.state, None)
        self.step_index += 1
        self.t = self.step_index * self.grid.dt
        return self.state

    def run(self, nsteps):
        for _ in range(nsteps):
            self.step()
        return self.state
