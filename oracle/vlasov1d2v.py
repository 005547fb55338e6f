"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's ``vlasov-1d2v`` time step (adept/_vlasov1d2v/), the
sibling solver SURVEY.md 8f ranks 4th: f(x, v_par, v_perp) in cylindrical velocity geometry, advection along x and
v_par with v_perp as a spectator, the 1-D field machinery fed with the v_perp marginals, and the v_par-only
Fokker-Planck operator whose coefficients come from the marginal.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this file; nothing under adept_b200/ does.

Built on oracle/vlasov1d.py (field solvers, drivers, operator assembly).  Pins: the reference's own 1D-limit identity
(tests/test_vlasov1d2v/test_1d_limit.py: the marginal of a v_perp-separable run equals the vlasov-1d solve) and the
telescoping identity of the cumulative diagnostics hold for this restatement (tests/test_oracle_1d2v.py); post-step
arrays are parity-unpinned against real JAX output, as for the 1-D oracle.  Not restated: the cylindrical_landau
operator (fokker_planck.py:143-403).
"""

from copy import deepcopy

import numpy as np

from . import vlasov1d as O


def perp_grid(nvperp, vperp_max):
    """helpers.py:20-29: cell-centred v_perp, spacing, weights w = 2 pi v_perp dv_perp."""
    dvperp = vperp_max / nvperp
    vperp = np.linspace(dvperp / 2.0, vperp_max - dvperp / 2.0, nvperp)
    return vperp, dvperp, 2.0 * np.pi * vperp * dvperp


def build_cfg(cfg_in: dict) -> dict:
    """modules.py:44-118 + helpers.py:32-91: the 1-D derived quantities plus the perpendicular grid; every species'
    f = F_1d(x, v_par) M(v_perp) with M a Maxwellian at the component's T0 normalised to sum_j M_j w_j = 1."""
    cfg = O.build_cfg(cfg_in)
    g = cfg["grid"]
    nvperp, vperp_max = int(g["nvperp"]), float(g["vperp_max"])
    vperp, dvperp, wperp = perp_grid(nvperp, vperp_max)
    species = cfg["terms"].get("species") or [{"name": "electron", "mass": 1.0, "nv": g["nv"],
                                                "vmax": float(g["vmax"]),
                                                "vmin": float(g["vmin"]) if g.get("vmin") is not None else -float(g["vmax"]),
                                                "density_components": [k for k in cfg["density"] if k.startswith("species-")]}]
    dists = {}
    for s in species:
        name, nv, mass = s["name"], int(s["nv"]), float(s["mass"])
        vmax = float(s["vmax"])
        vmin = float(s["vmin"]) if s.get("vmin") is not None else -vmax
        f_s = np.zeros((g["nx"], nv, nvperp))
        for cname in s["density_components"]:
            comp = cfg["density"][cname]
            nprof = np.array(O.density_profile(comp, g["x"]))
            tmp, _ = O.supergaussian_f0(g["nx"], nv, float(comp["v0"]), float(comp.get("m", 2.0)), float(comp["T0"]),
                                        mass, vmax, vmin, nprof)
            m_perp = np.exp(-(vperp**2.0) / (2.0 * float(comp["T0"]) / mass))
            m_perp = m_perp / np.sum(m_perp * wperp)
            f_s += tmp[:, :, None] * m_perp[None, None, :]
        n_s, _, v_ax = g["species_distributions"][name]
        dists[name] = (n_s, f_s, v_ax, vperp)
        g["species_grids"][name].update(vperp=vperp, dvperp=dvperp, nvperp=nvperp, vperp_max=vperp_max, wperp=wperp)
    g["species_distributions"] = dists
    d = cfg.setdefault("diagnostics", {})
    d.setdefault("diag-vlasov-cumulative", False)
    d.setdefault("diag-fp-cumulative", False)
    return cfg


def init_state(cfg: dict) -> dict:
    """modules.py:120-146."""
    g = cfg["grid"]
    state = {name: np.array(d[1]) for name, d in g["species_distributions"].items()}
    ref = "electron" if "electron" in state else next(iter(state))
    for k in ("e", "de"):
        state[k] = np.zeros(g["nx"])
    for k in ("a", "da", "prev_a"):
        state[k] = np.zeros(g["nx"] + 2)
    nv = state[ref].shape[1]
    for k in ("diag-vlasov-cumulative", "diag-fp-cumulative"):
        if cfg["diagnostics"].get(k, False):
            state[k] = np.zeros((g["nx"], nv))
    return state


def marginal(f, wperp):
    """einsum("xvp,p->xv") (vector_field.py:36-38)."""
    return np.einsum("xvp,p->xv", f, wperp)


def space_exponential_2v(f, kx_real, v, dt):
    """pushers/vlasov.py:41-56."""
    vdt = v * dt
    return np.real(np.fft.irfft(np.exp(-1j * kx_real[:, None, None] * vdt[None, :, None]) * np.fft.rfft(f, axis=0),
                                n=f.shape[0], axis=0))


def velocity_exponential_2v(f, kv_real, e, pond, dt, q, m):
    """pushers/vlasov.py:12-37."""
    accel = O.accel_from_fields(e, pond, q, m)
    return np.real(np.fft.irfft(np.exp(-1j * kv_real[None, :, None] * dt * accel[:, None, None]) * np.fft.rfft(f, axis=1),
                                n=f.shape[1], axis=1))


class Collisions2V:
    """pushers/fokker_planck.py:31-140 (dougherty, dougherty_nodrag, lenard_bernstein; electron species only)."""

    def __init__(self, cfg):
        self.cfg = cfg
        fp_type = cfg["terms"]["fokker_planck"]["type"].casefold()
        if fp_type not in ("dougherty", "dougherty_nodrag", "lenard_bernstein"):
            raise NotImplementedError(f"Unknown Fokker-Planck type for vlasov-1d2v: {fp_type}")
        if cfg["terms"]["krook"]["is_on"]:
            raise NotImplementedError("Krook is not implemented for vlasov-1d2v")
        self.c1 = O.Collisions(cfg)  # model, sc controls, v, dv of the electron grid
        self.wperp = np.asarray(cfg["grid"]["species_grids"]["electron"]["wperp"])
        self.nodrag = fp_type == "dougherty_nodrag"

    def __call__(self, nu_fp, f, dt):
        if isinstance(f, dict):
            return {k: (self._apply(nu_fp, fs, dt) if k == "electron" else fs) for k, fs in f.items()}
        return self._apply(nu_fp, f, dt)

    def _apply(self, nu_fp, f, dt):
        if not self.cfg["terms"]["fokker_planck"]["is_on"]:
            return f
        c1 = self.c1
        v, dv = c1.v, c1.dv
        nu = nu_fp if nu_fp is not None else np.zeros(f.shape[0])
        F = marginal(f, self.wperp)
        vbar, beta, C_edge, D = c1.moments_beta(F)
        if self.nodrag:
            C_edge = np.zeros_like(C_edge)
        ft = np.transpose(f, (0, 2, 1))
        ft_new = np.empty_like(ft)
        for i in range(f.shape[0]):
            for j in range(ft.shape[1]):
                ft_new[i, j] = O.solve_one_x(c1.scheme, C_edge[i], D[i], nu[i], ft[i, j], dt, dv)
        if self.nodrag:
            n_sl = np.sum(ft, axis=-1) * dv
            f_mx = np.exp(-beta[:, None, None] * (v[None, None, :] - vbar[:, None, None]) ** 2)
            f_mx = f_mx * (n_sl / (np.sum(f_mx, axis=-1) * dv))[..., None]
            DfM = D[:, None, None] * f_mx
            lap = np.zeros_like(DfM)
            lap[..., 1:-1] = (DfM[..., 2:] - 2.0 * DfM[..., 1:-1] + DfM[..., :-2]) / dv**2
            lap[..., 0] = (DfM[..., 1] - DfM[..., 0]) / dv**2
            lap[..., -1] = (DfM[..., -2] - DfM[..., -1]) / dv**2
            ft_new = ft_new - dt * nu[:, None, None] * lap
        return np.transpose(ft_new, (0, 2, 1))


class VlasovMaxwell2V:
    """solvers/vector_field.py:19-239 in one class."""

    def __init__(self, cfg):
        self.cfg, self.g = cfg, cfg["grid"]
        g = self.g
        self.dt = g["dt"]
        self.sg, self.sp = g["species_grids"], g["species_params"]
        if cfg["terms"]["edfdv"] != "exponential":
            raise NotImplementedError("vlasov-1d2v supports edfdv: exponential only")
        self.v1 = O.VlasovMaxwell(_cfg_1d(cfg))  # field solver, drivers, profiles, sixth-order coefficients
        self.time = cfg["terms"]["time"]
        self.dt_array, self.dex_save = self.v1.dt_array, self.v1.dex_save
        self.fp = Collisions2V(cfg)
        self.vlasov_cum = cfg["diagnostics"].get("diag-vlasov-cumulative", False)
        self.fp_cum = cfg["diagnostics"].get("diag-fp-cumulative", False)

    def marginals(self, f_dict):
        return {k: marginal(f, self.sg[k]["wperp"]) for k, f in f_dict.items()}

    def vdfdx(self, f_dict, dt):
        return {k: space_exponential_2v(f, self.g["kxr"], self.sg[k]["v"], dt) for k, f in f_dict.items()}

    def edfdv(self, f_dict, e, pond, dt):
        return {k: velocity_exponential_2v(f, self.sg[k]["kvr"], e, pond, dt, self.sp[k]["charge"], self.sp[k]["mass"])
                for k, f in f_dict.items()}

    def leapfrog(self, f_dict, a, dex, prev_ex):
        f_after = self.vdfdx(f_dict, self.dt)
        f_for_field = f_dict if self.v1.field_solve.hampere else f_after
        pond, e = self.v1.field_solve(self.marginals(f_for_field), a, prev_ex, self.dt)
        return e, self.edfdv(f_after, e + dex[0], pond, self.dt)

    def sixth(self, f_dict, a, dex, prev_ex):
        v1, dt = self.v1, self.dt
        kicks, drifts = [v1.D1, v1.D2, v1.D3, v1.D3, v1.D2, v1.D1], [v1.a1, v1.a2, v1.a3, v1.a2, v1.a1]
        pond, e = v1.field_solve(self.marginals(f_dict), a, None, None)
        f_dict = self.edfdv(f_dict, dex[0] + e, pond, kicks[0] * dt)
        for i, d in enumerate(drifts):
            f_dict = self.vdfdx(f_dict, d * dt)
            pond, e = v1.field_solve(self.marginals(f_dict), a, None, None)
            f_dict = self.edfdv(f_dict, dex[i + 1] + e, pond, kicks[i + 1] * dt)
        return e, f_dict

    def electron_charge_density(self, f_dict):
        cd = np.zeros_like(self.g["x"])
        if "electron" in f_dict:
            sg = self.sg["electron"]
            cd += self.sp["electron"]["charge"] * np.sum(marginal(f_dict["electron"], sg["wperp"]), axis=1) * sg["dv"]
        return cd

    def __call__(self, t, y, args=None):
        g, v1 = self.g, self.v1
        dex = [O.ex_driver_field(v1.drivers_ex, g["x"], t + d) for d in self.dt_array]
        djy = O.ey_driver_source(v1.drivers_ey, g["x_a"], t + self.dt_array[1], v1.c)
        nu_fp = v1.nu_fp_prof(g["x"], t) if self.cfg["terms"]["fokker_planck"]["is_on"] else None
        f_dict = {k: v for k, v in y.items() if k in self.sg}
        n_n = self.electron_charge_density(f_dict)
        integ = self.sixth if self.time == "sixth" else self.leapfrog
        e, f_vlasov = integ(f_dict, y["a"], dex, y["e"])
        f_fp = self.fp(nu_fp, f_vlasov, self.dt)
        n_np1 = self.electron_charge_density(f_fp)
        a = O.wave_solver(y["a"], y["prev_a"], djy, -0.5 * (n_n + n_np1), v1.c, g["dx"], self.dt)
        result = {"a": a["a"], "prev_a": a["prev_a"], "da": djy, "de": dex[self.dex_save], "e": e}
        result.update(f_fp)
        ref = "electron" if "electron" in f_dict else next(iter(f_dict))
        w = self.sg[ref]["wperp"]
        if self.vlasov_cum:
            result["diag-vlasov-cumulative"] = y["diag-vlasov-cumulative"] + (marginal(f_vlasov[ref], w) - marginal(f_dict[ref], w))
        if self.fp_cum:
            result["diag-fp-cumulative"] = y["diag-fp-cumulative"] + (marginal(f_fp[ref], w) - marginal(f_vlasov[ref], w))
        return result


def _cfg_1d(cfg):
    """The same completed deck as the 1-D classes read it (the dfdt diagnostics of the 1-D step do not apply here)."""
    c = dict(cfg)
    c["diagnostics"] = {"diag-vlasov-dfdt": False, "diag-fp-dfdt": False}
    return c


def run(cfg, nsteps, save_every=None):
    """Fixed-step loop y_{n+1} = vf(t_n, y_n); returns (final state, [states every save_every steps incl. the first])."""
    vf = VlasovMaxwell2V(cfg)
    y = init_state(cfg)
    dt = cfg["grid"]["dt"]
    kept = [deepcopy(y)] if save_every else []
    for n in range(nsteps):
        y = vf(n * dt, y, None)
        if save_every and (n + 1) % save_every == 0:
            kept.append(deepcopy(y))
    return y, kept
