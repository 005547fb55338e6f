"""numpy restatement of the reference ``vlasov-1d`` step (TEST INFRASTRUCTURE ONLY).

Every function cites the reference file:line (relative to /root/reference) that it
follows.  Arithmetic is fp64 numpy; FFTs use numpy's pocketfft (same rfft/irfft
conventions as ``jnp.fft``); the tridiagonal solve uses LAPACK ``dgtsv`` (the routine
behind ``jax.lax.linalg.tridiagonal_solve`` on CPU).

Parity status: the JAX implementation itself cannot be run in this image (no jax), so
post-step arrays are "parity unpinned" against real JAX; this restatement is pinned by
the reference's golden grid/f0 vectors and its known-answer tests (see tests/).
"""

from __future__ import annotations

import math
from copy import deepcopy

import numpy as np
from scipy.linalg import lapack
from scipy.special import gamma, gammaln

# --------------------------------------------------------------------------------------
# grids and initial state
# --------------------------------------------------------------------------------------


def make_grid(xmin, xmax, nx, tmin, tmax_requested, dt_requested, override_dt_for_em=False, beta=1.0):
    """adept/_vlasov1d/grid.py:37-88 (Grid.__init__)."""
    g = {"xmin": float(xmin), "xmax": float(xmax), "nx": int(nx), "tmin": float(tmin)}
    dx = (xmax - xmin) / nx
    g["dx"] = dx
    if override_dt_for_em:
        c_light = 1.0 / beta
        dt = min(dt_requested, float(0.95 * dx / c_light))
    else:
        dt = dt_requested
    g["dt"] = dt
    nt = int(tmax_requested / dt + 1)
    g["nt"] = nt
    g["tmax"] = dt * nt
    g["max_steps"] = min(nt + 4, int(1e8))
    g["x"] = np.linspace(xmin + dx / 2, xmax - dx / 2, nx)
    g["t"] = np.linspace(0, g["tmax"], nt)
    g["kx"] = np.fft.fftfreq(nx, d=dx) * 2.0 * np.pi
    g["kxr"] = np.fft.rfftfreq(nx, d=dx) * 2.0 * np.pi
    one_over_kx = np.zeros(nx)
    one_over_kx[1:] = 1.0 / g["kx"][1:]
    g["one_over_kx"] = one_over_kx
    one_over_kxr = np.zeros(len(g["kxr"]))
    one_over_kxr[1:] = 1.0 / g["kxr"][1:]
    g["one_over_kxr"] = one_over_kxr
    g["x_a"] = np.concatenate([[g["x"][0] - dx], g["x"], [g["x"][-1] + dx]])
    return g


def species_grid(vmin, vmax, nv):
    """adept/_vlasov1d/modules.py:222-247 (per-species velocity grid + Fourier duals)."""
    dv = (vmax - vmin) / nv
    v = np.linspace(vmin + dv / 2.0, vmax - dv / 2.0, nv)
    kv = np.fft.fftfreq(nv, d=dv) * 2.0 * np.pi
    kvr = np.fft.rfftfreq(nv, d=dv) * 2.0 * np.pi
    one_over_kv = np.zeros(nv)
    one_over_kv[1:] = 1.0 / kv[1:]
    one_over_kvr = np.zeros(len(kvr))
    one_over_kvr[1:] = 1.0 / kvr[1:]
    return {
        "v": v,
        "dv": dv,
        "nv": nv,
        "vmax": vmax,
        "vmin": vmin,
        "kv": kv,
        "kvr": kvr,
        "one_over_kv": one_over_kv,
        "one_over_kvr": one_over_kvr,
    }


def supergaussian_f0(nx, nv, v0=0.0, m=2.0, T0=1.0, mass=1.0, vmax=6.0, vmin=None, n_prof=np.ones(1)):
    """adept/_vlasov1d/helpers.py:37-93 (_initialize_supergaussian_distribution_)."""
    if vmin is None:
        vmin = -vmax
    dv = (vmax - vmin) / nv
    vax = np.linspace(vmin + dv / 2.0, vmax - dv / 2.0, nv)
    v_thermal = np.sqrt(T0 / mass)
    alpha = np.sqrt(3.0 * gamma(3.0 / m) / gamma(5.0 / m))
    single = -(np.power(np.abs((vax[None, :] - v0) / (alpha * v_thermal)), m))
    single = np.exp(single)
    f = np.repeat(single, nx, axis=0)
    f = f / np.sum(f, axis=1)[:, None] / dv
    if n_prof.size > 1:
        f = n_prof[:, None] * f
    return f, vax


class Envelope:
    """adept/functions.py:46-80 (EnvelopeFunction)."""

    def __init__(self, center, width, rise, baseline=0.0, bump_height=1.0, is_trough=False):
        self.center, self.width, self.rise = float(center), float(width), float(rise)
        self.baseline, self.bump_height, self.is_trough = float(baseline), float(bump_height), bool(is_trough)

    @staticmethod
    def from_config(c):
        """adept/functions.py:82-103; plain floats only (no pint strings in the oracle)."""
        return Envelope(
            c["center"],
            c["width"],
            c["rise"],
            c.get("baseline", 0.0),
            c.get("bump_height", 1.0),
            c.get("bump_or_trough", "bump") == "trough",
        )

    def __call__(self, x):
        left = self.center - self.width * 0.5
        right = self.center + self.width * 0.5
        env = 0.5 * (np.tanh((x - left) / self.rise) - np.tanh((x - right) / self.rise))
        if self.is_trough:
            env = 1 - env
        return self.baseline + self.bump_height * env


class SpaceTimeEnvelope:
    """adept/functions.py:106-135 (time_envelope(t) * space_envelope(x))."""

    def __init__(self, time_env, space_env):
        self.time_envelope, self.space_envelope = time_env, space_env

    @staticmethod
    def from_config(c):
        return SpaceTimeEnvelope(Envelope.from_config(c["time"]), Envelope.from_config(c["space"]))

    def __call__(self, x, t):
        return self.time_envelope(t) * self.space_envelope(x)


def density_profile(comp: dict, x: np.ndarray) -> np.ndarray:
    """adept/_vlasov1d/simulation.py:205-270 (SubspeciesDensityProfile); noise_val must be 0."""
    basis = comp["basis"]
    if basis == "uniform":
        base = comp.get("baseline")
        prof = (float(base) if base is not None else 1.0) * np.ones_like(x)
    elif basis == "sine":
        prof = float(comp["baseline"]) * (1.0 + float(comp["amplitude"]) * np.sin(float(comp["wavenumber"]) * x))
    elif basis == "tanh":
        prof = Envelope.from_config(comp)(x) * np.ones_like(x)
    elif basis == "linear":
        prof = Envelope.from_config(comp)(x) * (
            float(comp["val at center"]) + (x - float(comp["center"])) / float(comp["gradient scale length"])
        )
    elif basis == "exponential":
        prof = Envelope.from_config(comp)(x) * (
            float(comp["val at center"]) * np.exp((x - float(comp["center"])) / float(comp["gradient scale length"]))
        )
    else:
        raise NotImplementedError(basis)
    if float(comp.get("noise_val", 0.0)) != 0.0:
        raise NotImplementedError("oracle: jax.random noise is not reproducible without jax; use noise_val=0")
    return prof * (1.0 + np.zeros_like(prof))


def electron_beta(normalizing_temperature: str, reference: str = "electron", A: float = 1.0) -> float:
    """beta = v0/c with v0 = sqrt(T0/m0), m0 = m_e (adept/normalization.py:100-121) or A m_p for units.reference = ion
    (normalization.py:124-152); modules.py:55."""
    s = normalizing_temperature.strip()
    scale = 1.0
    if s.endswith("keV"):
        scale, s = 1.0e3, s[:-3]
    elif s.endswith("eV"):
        s = s[:-2]
    else:
        raise ValueError(f"oracle only parses eV/keV temperatures, got {normalizing_temperature}")
    T_eV = float(s) * scale
    me_c2_eV = 510998.95  # CODATA m_e c^2 (pint's registry value to 8 s.f.)
    mp_c2_eV = 938272088.16
    return math.sqrt(T_eV / (me_c2_eV if reference == "electron" else A * mp_c2_eV))


def build_cfg(cfg_in: dict) -> dict:
    """Restates sim_from_config + get_derived/solver_quantities + init_state.

    adept/_vlasov1d/modules.py:26-88 (species + grid), :190-277 (species_grids/params,
    ion_charge), :279-317 (state dict).  Only dimensionless decks (no pint strings).
    """
    cfg = deepcopy(cfg_in)
    units = cfg.get("units")
    beta = electron_beta(units["normalizing_temperature"], units.get("reference", "electron"),
                         float(units.get("A", 1.0))) if units else 1.0
    gin = cfg["grid"]
    has_ey = len(cfg["drivers"].get("ey", {})) > 0
    grid = make_grid(gin["xmin"], gin["xmax"], gin["nx"], gin.get("tmin", 0.0), gin["tmax"], gin["dt"], has_ey, beta)
    grid["beta"] = beta

    terms = cfg["terms"]
    if terms.get("species"):
        species = [dict(s) for s in terms["species"]]
        for s in species:
            s["vmax"] = float(s["vmax"])
            s["vmin"] = float(s["vmin"]) if s.get("vmin") is not None else -s["vmax"]
    else:
        comps = [k for k in cfg["density"].keys() if k.startswith("species-")]
        if not comps:
            raise ValueError("No density components found (expected keys starting with 'species-')")
        vmax = float(gin["vmax"])
        vmin = float(gin["vmin"]) if gin.get("vmin") is not None else -vmax
        species = [
            {
                "name": "electron",
                "charge": -1.0,
                "mass": 1.0,
                "vmax": vmax,
                "vmin": vmin,
                "nv": gin["nv"],
                "density_components": comps,
            }
        ]

    grid["species_grids"], grid["species_params"], grid["species_distributions"] = {}, {}, {}
    n_prof_total = np.zeros(grid["nx"])
    for s in species:
        name, nv, mass = s["name"], int(s["nv"]), float(s["mass"])
        n_prof_s = np.zeros(grid["nx"])
        f_s = np.zeros((grid["nx"], nv))
        first_T0 = None
        for cname in s["density_components"]:
            comp = cfg["density"][cname]
            nprof = np.array(density_profile(comp, grid["x"]))
            n_prof_s += nprof
            tmp, _ = supergaussian_f0(
                grid["nx"], nv, float(comp["v0"]), float(comp.get("m", 2.0)), float(comp["T0"]), mass, s["vmax"],
                s["vmin"], nprof,
            )
            f_s += tmp
            if first_T0 is None:
                first_T0 = float(comp["T0"])
        sg = species_grid(s["vmin"], s["vmax"], nv)
        grid["species_grids"][name] = sg
        grid["species_params"][name] = {
            "charge": float(s["charge"]),
            "mass": mass,
            "charge_to_mass": float(s["charge"]) / mass,
            "T0": first_T0,
        }
        grid["species_distributions"][name] = (n_prof_s, f_s, sg["v"])
        n_prof_total += n_prof_s
    grid["n_prof_total"] = n_prof_total
    grid["ion_charge"] = np.zeros_like(n_prof_total) if len(species) > 1 else n_prof_total.copy()
    if len(species) == 1 and "electron" in grid["species_grids"]:  # modules.py:269-275 (grid-level aliases)
        for k in ("v", "kv", "kvr", "one_over_kv", "one_over_kvr"):
            if k in grid["species_grids"]["electron"]:
                grid[k] = grid["species_grids"]["electron"][k]
    cfg["grid"] = {**gin, **grid}
    cfg.setdefault("diagnostics", {"diag-vlasov-dfdt": False, "diag-fp-dfdt": False})
    return cfg


def init_state(cfg: dict) -> dict:
    """adept/_vlasov1d/modules.py:279-317."""
    g = cfg["grid"]
    state = {name: np.array(d[1]) for name, d in g["species_distributions"].items()}
    ref = "electron" if "electron" in state else next(iter(state))
    for k in ["e", "de"]:
        state[k] = np.zeros(g["nx"])
    for k in ["a", "da", "prev_a"]:
        state[k] = np.zeros(g["nx"] + 2)
    for k in ["diag-vlasov-dfdt", "diag-fp-dfdt"]:
        if cfg["diagnostics"].get(k, False):
            state[k] = np.zeros_like(state[ref])
    return state


# --------------------------------------------------------------------------------------
# Vlasov pushers  (adept/_vlasov1d/solvers/pushers/vlasov.py)
# --------------------------------------------------------------------------------------

# bench.py's reference arm sets this to the host core count: the batched 1-D transforms then run through
# scipy.fft (pocketfft, ``workers=``); 1 keeps numpy's single-threaded pocketfft (used by every parity test).
FFT_WORKERS = 1


def _rfft(a, axis):
    if FFT_WORKERS > 1:
        import scipy.fft

        return scipy.fft.rfft(a, axis=axis, workers=FFT_WORKERS)
    return np.fft.rfft(a, axis=axis)


def _irfft(a, axis):
    if FFT_WORKERS > 1:
        import scipy.fft

        return scipy.fft.irfft(a, axis=axis, workers=FFT_WORKERS)
    return np.fft.irfft(a, axis=axis)



def space_exponential(f, kx_real, v, dt):
    """vlasov.py:234-251: irfft(exp(-i kx (v dt)) rfft(f, axis=0), axis=0)."""
    vdt = v * dt
    return np.real(_irfft(np.exp(-1j * kx_real[:, None] * vdt[None, :]) * _rfft(f, axis=0), axis=0))


def accel_from_fields(e, pond, q, m):
    """vlasov.py:81-84: force = q e + (q^2/m) pond; accel = force / m."""
    force = q * e + (q**2 / m) * pond
    return force / m


def velocity_exponential(f, kv_real, e, pond, dt, q, m):
    """vlasov.py:74-91."""
    accel = accel_from_fields(e, pond, q, m)
    return np.real(
        _irfft(np.exp(-1j * kv_real[None, :] * dt * accel[:, None]) * _rfft(f, axis=1), axis=1)
    )


def uniform_cubic_interp(f, shift, dv):
    """vlasov.py:106-148 (_uniform_cubic_interp)."""
    _, nv = f.shape
    if nv < 2:
        raise ValueError("cubic interpolation requires at least two velocity cells")
    scaled_shift = shift / np.asarray(dv, dtype=f.dtype)
    vidx = np.arange(nv, dtype=np.int32)[None, :]
    row_offset = np.floor(-scaled_shift).astype(np.int32)[:, None]
    left = np.clip(vidx + row_offset, 0, nv - 2)
    query_index = vidx - scaled_shift[:, None]
    t = np.clip(query_index - left, 0.0, 1.0)
    fm1 = np.take_along_axis(f, np.clip(left - 1, 0, nv - 1), axis=1)
    f0 = np.take_along_axis(f, left, axis=1)
    f1 = np.take_along_axis(f, left + 1, axis=1)
    f2 = np.take_along_axis(f, np.clip(left + 2, 0, nv - 1), axis=1)
    m0 = np.where(left == 0, f1 - f0, 0.5 * (f1 - fm1))
    m1 = np.where(left == nv - 2, f1 - f0, 0.5 * (f2 - f0))
    t2 = t * t
    t3 = t2 * t
    interpolated = (
        (2.0 * t3 - 3.0 * t2 + 1.0) * f0 + (t3 - 2.0 * t2 + t) * m0 + (-2.0 * t3 + 3.0 * t2) * f1 + (t3 - t2) * m1
    )
    outside = (query_index < 0.0) | (query_index > nv - 1)
    return np.where(outside, np.asarray(1.0e-30, dtype=f.dtype), interpolated)


def velocity_cubic_spline(f, dv, e, pond, dt, q, m):
    """vlasov.py:162-172."""
    accel = accel_from_fields(e, pond, q, m)
    return uniform_cubic_interp(f, accel * dt, dv)


def hou_li_filter(f, nx, alpha, order):
    """vlasov.py:209-220."""
    j_x = np.arange(nx // 2 + 1)
    eta_x = j_x / (nx // 2)
    filt = np.exp(-alpha * eta_x ** (2 * order))
    return np.real(np.fft.irfft(filt[:, None] * np.fft.rfft(f, axis=0), axis=0))


# --------------------------------------------------------------------------------------
# field solvers (adept/_vlasov1d/solvers/pushers/field.py)
# --------------------------------------------------------------------------------------


def charge_density(f_dict, species_grids, species_params, static=None):
    """field.py:186-208."""
    rho = np.zeros_like(next(iter(f_dict.values()))[:, 0])
    for name, f_s in f_dict.items():
        q_s = species_params[name]["charge"]
        dv_s = species_grids[name]["dv"]
        n_s = np.sum(f_s, axis=1) * dv_s
        rho = rho + q_s * n_s
    if static is not None:
        rho = rho + static
    return rho


def poisson(rho, one_over_kx):
    """field.py:221-224."""
    return np.real(np.fft.ifft(-1j * one_over_kx * np.fft.fft(rho)))


def boltzmann_poisson(rho, kx, Te, lambda_De=None):
    """field.py:293-298."""
    rho_0 = np.mean(rho)
    lambda_sq = Te / rho_0 if lambda_De is None else lambda_De**2
    kernel = kx * (Te / rho_0) / (1.0 + lambda_sq * kx**2)
    return np.real(np.fft.ifft(-1j * kernel * np.fft.fft(rho)))


def current_density(f_dict, species_grids, species_params):
    """field.py:319-340."""
    j = np.zeros_like(next(iter(f_dict.values()))[:, 0])
    for name, f_s in f_dict.items():
        q_s = species_params[name]["charge"]
        v_s = species_grids[name]["v"]
        dv_s = species_grids[name]["dv"]
        j_s = np.sum(v_s[None, :] * f_s, axis=1) * dv_s
        j = j + q_s * j_s
    return j


def ampere(f_dict, species_grids, species_params, prev_ex, dt):
    """field.py:342-354."""
    return prev_ex - dt * current_density(f_dict, species_grids, species_params)


def hampere(f, kx, one_over_kx, v, dv, charge, prev_ex, dt):
    """field.py:395-419."""
    one_over_ikx = one_over_kx / 1j
    prev_ek = np.fft.fft(prev_ex, axis=0)
    fk = np.fft.fft(f, axis=0)
    new_ek = prev_ek + charge * one_over_ikx * np.sum(
        fk * (np.exp(-1j * kx[:, None] * dt * v[None, :]) - 1), axis=1
    ) * dv
    return np.real(np.fft.ifft(new_ek))


def ponderomotive(a, dx):
    """field.py:495: -0.5 * gradient(a**2, dx)[1:-1]."""
    return -0.5 * np.gradient(a**2.0, dx)[1:-1]


class ElectricFieldSolver:
    """field.py:422-497."""

    def __init__(self, cfg):
        g = cfg["grid"]
        self.cfg, self.g = cfg, g
        self.kind = cfg["terms"]["field"]
        time = cfg["terms"]["time"]
        if self.kind in ("ampere", "hampere") and time != "leapfrog":
            raise NotImplementedError(f"ampere + {time} has not yet been implemented")
        if self.kind == "hampere" and len(g["species_grids"]) > 1:
            raise NotImplementedError("HampereSolver currently only supports single-species simulations.")
        if self.kind not in ("poisson", "poisson-boltzmann", "ampere", "hampere"):
            raise NotImplementedError(self.kind)
        self.hampere = self.kind == "hampere"
        self.dx = g["dx"]

    def __call__(self, f_dict, a, prev_ex, dt):
        g = self.g
        pond = ponderomotive(a, self.dx)
        if self.kind == "poisson":
            rho = charge_density(f_dict, g["species_grids"], g["species_params"], g.get("ion_charge"))
            e = poisson(rho, g["one_over_kx"])
        elif self.kind == "poisson-boltzmann":
            bz = self.cfg["terms"].get("boltzmann_electrons") or {}
            rho = charge_density(f_dict, g["species_grids"], g["species_params"], None)
            e = boltzmann_poisson(rho, g["kx"], bz.get("Te", 1.0), bz.get("lambda_De"))
        elif self.kind == "ampere":
            e = ampere(f_dict, g["species_grids"], g["species_params"], prev_ex, dt)
        else:
            name = next(iter(g["species_grids"]))
            sg = g["species_grids"][name]
            e = hampere(
                next(iter(f_dict.values())), g["kx"], g["one_over_kx"], sg["v"], sg["dv"],
                g["species_params"][name]["charge"], prev_ex, dt,
            )
        return pond, e


def wave_solver(a, aold, djy, electron_density, c, dx, dt):
    """field.py:97-157 (WaveSolver incl. apply_2nd_order_abc)."""
    if not c > 0:
        return {"a": a, "prev_a": aold}
    c_sq = c**2.0
    c_over_dx = c / dx
    const = c_over_dx * dt
    one_over_const = 1.0 / dt / c_over_dx
    d2dx2 = (a[:-2] - 2.0 * a[1:-1] + a[2:]) / dx**2.0
    anew = 2.0 * a[1:-1] - aold[1:-1] + dt**2.0 * (c_sq * d2dx2 - electron_density * a[1:-1] + djy[1:-1])
    coeff = -1.0 / (one_over_const + 2.0 + const)
    a_left = (one_over_const - 2.0 + const) * (anew[1] + aold[0])
    a_left += 2.0 * (const - one_over_const) * (a[0] + a[2] - anew[0] - aold[1])
    a_left -= 4.0 * (one_over_const + const) * a[1]
    a_left *= coeff
    a_left -= aold[2]
    a_right = (one_over_const - 2.0 + const) * (anew[-2] + aold[-1])
    a_right += 2.0 * (const - one_over_const) * (a[-1] + a[-3] - anew[-1] - aold[-2])
    a_right -= 4.0 * (one_over_const + const) * a[-2]
    a_right *= coeff
    a_right -= aold[-3]
    return {"a": np.concatenate([[a_left], anew, [a_right]]), "prev_a": a}


# --------------------------------------------------------------------------------------
# drivers (adept/_vlasov1d/solvers/pushers/field.py:13-91, simulation.py:37-93)
# --------------------------------------------------------------------------------------


class EMDriver:
    def __init__(self, a0, k0, w0, dw0, envelope, is_point_source=False):
        self.a0, self.k0, self.w0, self.dw0 = float(a0), float(k0), float(w0), float(dw0)
        self.envelope, self.is_point_source = envelope, is_point_source

    @staticmethod
    def from_config(c, c_light):
        """simulation.py:47-63 (AKW params only)."""
        p = c["params"]
        k0, w0 = p.get("k0"), p.get("w0")
        if k0 is None and w0 is None:
            raise ValueError("You must specify at least one of k0 or w0.")
        if k0 is None:
            k0 = w0 / c_light
        return EMDriver(
            p["a0"], k0, w0, p.get("dw0", 0.0), SpaceTimeEnvelope.from_config(c["envelope"]),
            c.get("source_type", "extended") == "point",
        )


class StochasticDriver:
    """simulation.py:95-148: band-limited Ornstein-Uhlenbeck forcing of Ex.  The reference draws the realisation with
    numpy's Generator (not jax.random), so the same seed gives the same amplitude series here."""

    def __init__(self, scfg, xmin, xmax, tmin, tmax):
        length = xmax - xmin
        modes = np.asarray(scfg.get("modes", [1]), dtype=np.float64)
        n_modes = len(modes)
        amplitude, tau = scfg["amplitude"], scfg["tau"]
        dt_update = scfg["dt_update"] if scfg.get("dt_update") is not None else tau / 10.0
        dt_update = min(dt_update, tau / 2.0)
        nt = int(np.ceil((tmax - tmin) / dt_update)) + 2
        rng = np.random.default_rng(scfg.get("seed", 42))
        theta = np.exp(-dt_update / tau)
        kick = amplitude * np.sqrt(1.0 - theta**2)
        amps = np.zeros((nt, n_modes), dtype=np.complex128)
        amps[0] = amplitude * (rng.standard_normal(n_modes) + 1j * rng.standard_normal(n_modes)) / np.sqrt(2.0)
        for it in range(1, nt):
            xi = (rng.standard_normal(n_modes) + 1j * rng.standard_normal(n_modes)) / np.sqrt(2.0)
            amps[it] = theta * amps[it - 1] + kick * xi
        self.t_grid = tmin + dt_update * np.arange(nt)
        self.amp_real, self.amp_imag = amps.real.copy(), amps.imag.copy()
        self.k_modes = 2.0 * np.pi * modes / length

    def __call__(self, t, x):
        ar = np.array([np.interp(t, self.t_grid, col) for col in self.amp_real.T])
        ai = np.array([np.interp(t, self.t_grid, col) for col in self.amp_imag.T])
        phase = self.k_modes[:, None] * x[None, :]
        return np.sum(ar[:, None] * np.cos(phase) - ai[:, None] * np.sin(phase), axis=0)


def ex_driver_field(drivers, x, t):
    """field.py:21-33."""
    total = np.zeros_like(x)
    for d in drivers:
        factor = d.envelope(x, t)
        total += factor * (d.w0 + d.dw0) * d.a0 * np.sin(d.k0 * x - (d.w0 + d.dw0) * t)
    return total


def ey_driver_source(drivers, x_a, t, c):
    """field.py:53-91."""
    total = np.zeros_like(x_a)
    dx = float(x_a[1] - x_a[0])
    for d in drivers:
        w_total = d.w0 + d.dw0
        if d.is_point_source:
            i0 = np.argmin(np.abs(x_a - d.envelope.space_envelope.center))
            mask = np.zeros_like(x_a)
            mask[i0] = 1.0
            F0 = 2.0 * w_total * c * d.a0
            total += (F0 / dx) * d.envelope.time_envelope(t) * mask * np.sin(w_total * t)
        else:
            factor = d.envelope(x_a, t)
            total += -factor * w_total**2 * d.a0 * np.sin(d.k0 * x_a - w_total * t)
    return total


# --------------------------------------------------------------------------------------
# Fokker-Planck (adept/driftdiffusion.py, adept/_vlasov1d/solvers/pushers/fokker_planck.py)
# --------------------------------------------------------------------------------------


def chang_cooper_delta(w):
    """driftdiffusion.py:77-103."""
    w = np.asarray(w, dtype=np.float64)
    small = np.abs(w) < 1.0e-8
    w_safe = np.where(small, 1.0, w)
    delta_small = 0.5 - w / 12.0 + w**3 / 720.0
    with np.errstate(over="ignore"):
        delta_full = 1.0 / w_safe - 1.0 / np.expm1(w_safe)
    return np.where(small, delta_small, delta_full)


def discrete_temperature(f, v, dv, vbar=None):
    """driftdiffusion.py:106-137 (non-spherical branch)."""
    v_shifted = v if vbar is None else (v - vbar[..., None])
    vsq = v_shifted**2
    v2_moment = np.sum(f * vsq * dv, axis=-1)
    norm = np.sum(f * dv, axis=-1)
    return v2_moment / norm


def newton_root_find(residual_and_slope, y0, rtol, atol, max_steps):
    """optimistix 0.1.0 `optx.root_find(fn, optx.Newton(rtol, atol), y0, max_steps=..., throw=False)` for a scalar
    unknown, vectorised over rows the way `vmap` of the reference's `lax.while_loop` behaves (a finished row keeps its
    value).  optimistix is absent from /root/reference and from this image; this restates its published algorithm
    (`_solver/newton_chord.py`: `diff = J^-1 f(y)`, `y <- y - diff`; Cauchy termination, checked BEFORE every step,
    `|diff| < atol + rtol |y_new|` and `|f(y_old)| < atol`; the last iterate is returned when max_steps runs out).
    PARITY UNPINNED for the termination rule (no reference test stores a refined beta); call sites
    driftdiffusion.py:223-232 and fokker_planck.py:205-207.
    """
    y = np.array(y0, dtype=np.float64, copy=True)
    diff = np.full_like(y, np.inf)
    fprev = np.full_like(y, np.inf)
    done = np.zeros(y.shape, dtype=bool)
    for _ in range(int(max_steps)):
        done = done | ((np.abs(diff) < atol + rtol * np.abs(y)) & (np.abs(fprev) < atol))
        fx, slope = residual_and_slope(y)
        d = fx / slope
        y = np.where(done, y, y - d)
        diff = np.where(done, diff, d)
        fprev = np.where(done, fprev, fx)
    return y


def find_self_consistent_beta(f, v, dv, vbar=None, rtol=1e-8, atol=1e-12, max_steps=3):
    """driftdiffusion.py:161-283 (m = 2, non-spherical): beta* such that exp(-beta (v - vbar)^2) has the discrete
    temperature of f.  The slope is the derivative of `discrete_temperature(f_model)` w.r.t. beta that jax.linearize
    forms: d(v2/norm) = (dv2 norm - v2 dnorm) / norm^2 with df_model = -(v - vbar)^2 f_model."""
    T_target = discrete_temperature(f, v, dv, vbar)
    beta_init = 1.0 / (2.0 * T_target)
    if max_steps == 0:
        return beta_init
    vs = v[None, :] if vbar is None else (v[None, :] - vbar[:, None])
    sq = vs**2

    def residual_and_slope(beta):
        fm = np.exp(-beta[:, None] * sq)
        norm = np.sum(fm * dv, axis=-1)
        v2 = np.sum(fm * sq * dv, axis=-1)
        dnorm = -np.sum(fm * sq * dv, axis=-1)
        dv2 = -np.sum(fm * sq * sq * dv, axis=-1)
        return v2 / norm - T_target, (dv2 * norm - v2 * dnorm) / norm**2

    return newton_root_find(residual_and_slope, beta_init, rtol, atol, max_steps)


def chang_cooper_delta_prime(w):
    """d delta / d w of chang_cooper_delta as autodiff sees it (driftdiffusion.py:96-103): the derivative of the
    selected branch."""
    w = np.asarray(w, dtype=np.float64)
    small = np.abs(w) < 1.0e-8
    w_safe = np.where(small, 1.0, w)
    with np.errstate(over="ignore", invalid="ignore"):
        em = np.expm1(w_safe)
        full = -1.0 / w_safe**2 + np.exp(w_safe) / em**2
    full = np.where(np.isfinite(full), full, -1.0 / w_safe**2)  # exp overflow: 1/expm1 and its slope vanish
    return np.where(small, -1.0 / 12.0 + w**2 / 240.0, full)


def supergaussian_beta(f, v, m, rtol=1e-8, atol=1e-12, max_steps=3):
    """SuperGaussianDougherty.compute_beta, fokker_planck.py:139-210: continuum closure, then Newton on the discrete
    energy-flux condition h(beta) = sum_e v_e (w ftilde(w) + df), w = beta dpsi."""
    vbar = np.sum(f * v, axis=-1) / np.sum(f, axis=-1)
    psi = np.abs(v[None, :] - vbar[:, None]) ** m
    beta_init = np.sum(f, axis=-1) / (m * np.sum(f * psi, axis=-1))
    if max_steps == 0:
        return beta_init
    v_edge = 0.5 * (v[1:] + v[:-1])
    d_psi = psi[:, 1:] - psi[:, :-1]
    d_f = f[:, 1:] - f[:, :-1]
    fl, fr = f[:, :-1], f[:, 1:]

    def residual_and_slope(beta):
        w = beta[:, None] * d_psi
        delta = chang_cooper_delta(w)
        f_tilde = delta * fl + (1.0 - delta) * fr
        h = np.sum(v_edge * (w * f_tilde + d_f), axis=-1)
        dh = np.sum(v_edge * d_psi * (f_tilde + w * chang_cooper_delta_prime(w) * (fl - fr)), axis=-1)
        return h, dh

    return newton_root_find(residual_and_slope, beta_init, rtol, atol, max_steps)


def central_operator(C_edge, D, nu, dt, dv):
    """driftdiffusion.py:563-600 (CentralDifferencing.get_operator); one row."""
    nv = C_edge.shape[-1] + 1
    nu_full = np.broadcast_to(nu, (nv,))
    bare_diag = np.zeros(nv)
    bare_diag[:-1] += (C_edge / 2.0 - D / dv) / dv
    bare_diag[1:] += -(C_edge / 2.0 + D / dv) / dv
    bare_upper = (C_edge / 2.0 + D / dv) / dv
    bare_lower = (-C_edge / 2.0 + D / dv) / dv
    diag = 1.0 - dt * nu_full * bare_diag
    upper = -dt * nu_full[:-1] * bare_upper
    lower = -dt * nu_full[1:] * bare_lower
    return diag, upper, lower


def chang_cooper_operator(C_edge, D, nu, dt, dv):
    """driftdiffusion.py:614-658 (ChangCooper.get_operator); one row."""
    nv = C_edge.shape[-1] + 1
    nu_full = np.broadcast_to(nu, (nv,))
    safe_D = np.maximum(D, 1.0e-30)
    w = C_edge * dv / safe_D
    delta = chang_cooper_delta(w)
    alpha = -C_edge * delta + safe_D / dv
    beta = -C_edge * (1.0 - delta) - safe_D / dv
    bare_diag = np.zeros(nv)
    bare_diag[:-1] += -alpha / dv
    bare_diag[1:] += beta / dv
    bare_upper = -beta / dv
    bare_lower = alpha / dv
    diag = 1.0 - dt * nu_full * bare_diag
    upper = -dt * nu_full[:-1] * bare_upper
    lower = -dt * nu_full[1:] * bare_lower
    return diag, upper, lower


def tridiag_mv(diag, upper, lower, x):
    """lineax 0.1.0 TridiagonalLinearOperator.mv: (diag*x) .at[:-1].add(upper*x[1:]) .at[1:].add(lower*x[:-1])."""
    a = upper * x[1:]
    b = diag * x
    c = lower * x[:-1]
    b = b.copy()
    b[:-1] += a
    b[1:] += c
    return b


def tridiag_solve(lower, diag, upper, rhs):
    """jax.lax.linalg.tridiagonal_solve on CPU == LAPACK ?gtsv (partial pivoting)."""
    _, _, _, x, info = lapack.dgtsv(lower, diag, upper, rhs)
    if info != 0:
        raise np.linalg.LinAlgError(f"dgtsv info={info}")
    return x


def solve_one_x(scheme, C_edge, D, nu, f_v, dt, dv):
    """fokker_planck.py:368-376 (_solve_one_x): delta formulation."""
    op = central_operator if scheme == "central" else chang_cooper_operator
    diag, upper, lower = op(C_edge, D, nu, dt, dv)
    rhs = f_v - tridiag_mv(diag, upper, lower, f_v)
    delta = tridiag_solve(lower, diag, upper, rhs)
    return f_v + delta


FP_TYPES = {
    # fokker_planck.py:314-342: type -> (model, scheme)
    "lenard_bernstein": ("lb", "central"),
    "chang_cooper": ("lb", "cc"),
    "lenard_bernstein_chang_cooper": ("lb", "cc"),
    "chang_cooper_dougherty": ("dougherty", "cc"),
    "dougherty_chang_cooper": ("dougherty", "cc"),
    "dougherty": ("dougherty", "central"),
    "super_gaussian": ("sg", "cc"),
    "super_gaussian_chang_cooper": ("sg", "cc"),
    "dougherty_nodrag": ("dougherty", "central"),
}


def krook_fmx(v, dv, T0=1.0, mass=1.0):
    """fokker_planck.py:456-465."""
    f_mx = np.exp(-(v[None, :] ** 2.0) / (2.0 * T0 / mass))
    return f_mx / np.sum(f_mx, axis=1)[:, None] / dv


def krook(nu_K, f_xv, dt, f_mx, dv):
    """fokker_planck.py:471-484."""
    nu_Kxdt = dt * nu_K[:, None]
    exp_nuKxdt = np.exp(-nu_Kxdt)
    n_prof = np.sum(f_xv, axis=1) * dv
    return f_xv * exp_nuKxdt + n_prof[:, None] * f_mx * (1.0 - exp_nuKxdt)


class Collisions:
    """fokker_planck.py:272-443."""

    def __init__(self, cfg):
        self.cfg = cfg
        sg = cfg["grid"]["species_grids"]
        self.ref_species = "electron" if "electron" in sg else next(iter(sg))
        fp_cfg = cfg["terms"]["fokker_planck"]
        fp_type = fp_cfg.get("type", "").casefold()
        self.fp_on = fp_cfg["is_on"]
        if self.fp_on or fp_type:
            if fp_type not in FP_TYPES:
                raise NotImplementedError(f"Unknown Fokker-Planck type: {fp_type}")
            self.model, self.scheme = FP_TYPES[fp_type]
        else:
            self.model, self.scheme = "lb", "central"
        self.nodrag = fp_type == "dougherty_nodrag"
        self.m = float(fp_cfg.get("m", 2.0))
        sc = fp_cfg.get("self_consistent_beta", {})
        self.sc_max_steps = sc.get("max_steps", 3) if sc.get("enabled", False) else 0
        self.sc_rtol, self.sc_atol = sc.get("rtol", 1e-8), sc.get("atol", 1e-12)
        self.v = np.asarray(sg[self.ref_species]["v"])
        self.dv = sg[self.ref_species]["dv"]
        self.krook_on = cfg["terms"]["krook"]["is_on"]
        params = cfg["grid"].get("species_params", {}).get(self.ref_species, {})
        self.f_mx = krook_fmx(self.v, self.dv, params.get("T0", 1.0), params.get("mass", 1.0))

    def __call__(self, nu_fp, nu_K, f, dt):
        if isinstance(f, dict):
            return {k: (self._apply(nu_fp, nu_K, fs, dt) if k == self.ref_species else fs) for k, fs in f.items()}
        return self._apply(nu_fp, nu_K, f, dt)

    def moments_beta(self, f):
        """fokker_planck.py:391-410; returns (vbar or None, beta, C_edge, D)."""
        v, dv = self.v, self.dv
        v_edge = 0.5 * (v[1:] + v[:-1])
        if self.model == "sg":
            # fokker_planck.py:187-193, 211-255
            m = self.m
            vbar = np.sum(f * v, axis=-1) / np.sum(f, axis=-1)
            psi = np.abs(v[None, :] - vbar[:, None]) ** m
            beta = supergaussian_beta(f, v, m, self.sc_rtol, self.sc_atol, self.sc_max_steps)
            D = beta ** (-2.0 / m) * np.exp(gammaln(3.0 / m) - gammaln(1.0 / m))
            phi = beta[:, None] * psi
            C_edge = D[:, None] * (phi[:, 1:] - phi[:, :-1]) / dv
            return vbar, beta, C_edge, D
        vbar = np.sum(f * v, axis=-1) / np.sum(f, axis=-1) if self.model == "dougherty" else None
        beta = find_self_consistent_beta(f, v, dv, vbar, self.sc_rtol, self.sc_atol, self.sc_max_steps)
        D = 1.0 / (2.0 * beta)
        v_eff = v_edge[None, :] if vbar is None else (v_edge[None, :] - vbar[:, None])
        C_edge = 2.0 * beta[:, None] * D[:, None] * v_eff
        return vbar, beta, C_edge, D

    def _apply(self, nu_fp, nu_K, f, dt):
        nx = f.shape[0]
        nu_fp_in = nu_fp if nu_fp is not None else np.zeros(nx)
        nu_K_in = nu_K if nu_K is not None else np.zeros(nx)
        v, dv = self.v, self.dv
        if self.fp_on:
            vbar, beta, C_edge, D = self.moments_beta(f)
            if self.nodrag:
                C_edge = np.zeros_like(C_edge)
            f_new = np.empty_like(f)
            for i in range(nx):
                f_new[i] = solve_one_x(self.scheme, C_edge[i], D[i], nu_fp_in[i], f[i], dt, dv)
            if self.nodrag:
                n_prof = np.sum(f, axis=-1) * dv
                f_mx = np.exp(-beta[:, None] * (v[None, :] - vbar[:, None]) ** 2)
                f_mx = f_mx * (n_prof / (np.sum(f_mx, axis=-1) * dv))[:, None]
                DfM = D[:, None] * f_mx
                lap = np.zeros_like(DfM)
                lap[:, 1:-1] = (DfM[:, 2:] - 2.0 * DfM[:, 1:-1] + DfM[:, :-2]) / dv**2
                lap[:, 0] = (DfM[:, 1] - DfM[:, 0]) / dv**2
                lap[:, -1] = (DfM[:, -2] - DfM[:, -1]) / dv**2
                f_new = f_new - dt * nu_fp_in[:, None] * lap
            f = f_new
        if self.krook_on:
            f = krook(nu_K_in, f, dt, self.f_mx, dv)
        return f


# --------------------------------------------------------------------------------------
# integrators + vector field (adept/_vlasov1d/solvers/vector_field.py)
# --------------------------------------------------------------------------------------

SIXTH = dict(
    a1=0.168735950563437422448196,
    a2=0.377851589220928303880766,
    a3=-0.093175079568731452657924,
    b1=0.049086460976116245491441,
    b2=0.264177609888976700200146,
    b3=0.186735929134907054308413,
    c1=-0.000069728715055305084099,
    c2=-0.000625704827430047189169,
    c3=-0.002213085124045325561636,
    d2=-2.916600457689847816445691e-6,
    d3=3.048480261700038788680723e-5,
    e3=4.985549387875068121593988e-7,
)


def sixth_coefficients(dt):
    """vector_field.py:118-144: returns (a1,a2,a3,D1,D2,D3,dt_array)."""
    c = SIXTH
    a1, a2, a3 = c["a1"], c["a2"], c["a3"]
    D1 = c["b1"] + 2.0 * c["c1"] * dt**2.0
    D2 = c["b2"] + 2.0 * c["c2"] * dt**2.0 + 4.0 * c["d2"] * dt**4.0
    D3 = c["b3"] + 2.0 * c["c3"] * dt**2.0 + 4.0 * c["d3"] * dt**4.0 - 8.0 * c["e3"] * dt**6.0
    dt_array = dt * np.array([0.0, a1, a1 + a2, a1 + a2 + a3, a1 + a2 + a3 + a2, a1 + a2 + a3 + a2 + a1])
    return a1, a2, a3, D1, D2, D3, dt_array


class VlasovMaxwell:
    """vector_field.py:19-361: TimeIntegrator + Leapfrog/Sixth + VPFP + VlasovMaxwell, in one class."""

    def __init__(self, cfg, drivers_ex=None, drivers_ey=None, nu_fp_prof=None, nu_K_prof=None):
        self.cfg, self.g = cfg, cfg["grid"]
        g = self.g
        self.dt = g["dt"]
        self.sg, self.sp = g["species_grids"], g["species_params"]
        self.field_solve = ElectricFieldSolver(cfg)
        self.edfdv_kind = cfg["terms"]["edfdv"]
        if self.edfdv_kind not in ("exponential", "cubic-spline"):
            raise NotImplementedError(f"{self.edfdv_kind} has not been implemented")
        self.time = cfg["terms"]["time"]
        if self.time == "sixth":
            (self.a1, self.a2, self.a3, self.D1, self.D2, self.D3, self.dt_array) = sixth_coefficients(self.dt)
            self.dex_save = 3
        elif self.time == "leapfrog":
            self.dt_array = self.dt * np.array([0.0, 1.0])
            self.dex_save = 0
        else:
            raise NotImplementedError
        self.fp = Collisions(cfg)
        diag = cfg.get("diagnostics", {})
        self.vlasov_dfdt = diag.get("diag-vlasov-dfdt", False)
        self.fp_dfdt = diag.get("diag-fp-dfdt", False)
        hl = cfg["terms"].get("hou_li_filter", {"is_on": False})
        self.hou_li = hl if hl.get("is_on", False) else None
        beta = g.get("beta", 1.0)
        self.c = 1.0 / beta
        c_light = self.c
        dcfg = cfg.get("drivers", {"ex": {}, "ey": {}})
        self.drivers_ex = drivers_ex if drivers_ex is not None else [
            EMDriver.from_config(d, c_light) for d in dcfg.get("ex", {}).values()
        ]
        self.drivers_ey = drivers_ey if drivers_ey is not None else [
            EMDriver.from_config(d, c_light) for d in dcfg.get("ey", {}).values()
        ]
        scfg = dcfg.get("ex_stochastic")  # simulation.py:160-168
        self.ex_stochastic = StochasticDriver(scfg, g["xmin"], g["xmax"], g["tmin"], g["tmax"]) if scfg else None
        fpc, kc = cfg["terms"]["fokker_planck"], cfg["terms"]["krook"]
        self.nu_fp_prof = nu_fp_prof if nu_fp_prof is not None else (
            SpaceTimeEnvelope.from_config(fpc) if fpc["is_on"] else None
        )
        self.nu_K_prof = nu_K_prof if nu_K_prof is not None else (
            SpaceTimeEnvelope.from_config(kc) if kc["is_on"] else None
        )

    # pushers ---------------------------------------------------------------------------
    def vdfdx(self, f_dict, dt):
        return {k: space_exponential(f, self.g["kxr"], self.sg[k]["v"], dt) for k, f in f_dict.items()}

    def edfdv(self, f_dict, e, pond, dt):
        out = {}
        for k, f in f_dict.items():
            q, m = self.sp[k]["charge"], self.sp[k]["mass"]
            if self.edfdv_kind == "exponential":
                out[k] = velocity_exponential(f, self.sg[k]["kvr"], e, pond, dt, q, m)
            else:
                out[k] = velocity_cubic_spline(f, self.sg[k]["dv"], e, pond, dt, q, m)
        return out

    # integrators -----------------------------------------------------------------------
    def leapfrog(self, f_dict, a, dex, prev_ex):
        """vector_field.py:75-95."""
        f_after_v = self.vdfdx(f_dict, self.dt)
        f_for_field = f_dict if self.field_solve.hampere else f_after_v
        pond, e = self.field_solve(f_for_field, a, prev_ex, self.dt)
        f_out = self.edfdv(f_after_v, e + dex[0], pond, self.dt)
        return e, f_out

    def sixth(self, f_dict, a, dex, prev_ex):
        """vector_field.py:146-186."""
        dt = self.dt
        seq_D = [self.D1, self.D2, self.D3, self.D3, self.D2, self.D1]
        seq_a = [self.a1, self.a2, self.a3, self.a2, self.a1]
        e = None
        for i in range(6):
            pond, e = self.field_solve(f_dict, a, None, None)
            f_dict = self.edfdv(f_dict, dex[i] + e, pond, seq_D[i] * dt)
            if i < 5:
                f_dict = self.vdfdx(f_dict, seq_a[i] * dt)
        return e, f_dict

    def vpfp(self, f_dict, a, prev_ex, dex, nu_fp, nu_K):
        """vector_field.py:232-253."""
        integ = self.sixth if self.time == "sixth" else self.leapfrog
        e, f_vlasov = integ(f_dict, a, dex, prev_ex)
        f_fp = self.fp(nu_fp, nu_K, f_vlasov, self.dt)
        if self.hou_li is not None:
            f_fp = {
                k: hou_li_filter(f, self.g["nx"], self.hou_li.get("alpha", 36.0), self.hou_li.get("order", 36))
                for k, f in f_fp.items()
            }
        diags = {}
        ref = "electron" if "electron" in f_dict else next(iter(f_dict))
        if self.vlasov_dfdt:
            diags["diag-vlasov-dfdt"] = (f_vlasov[ref] - f_dict[ref]) / self.dt
        if self.fp_dfdt:
            diags["diag-fp-dfdt"] = (f_fp[ref] - f_vlasov[ref]) / self.dt
        return e, f_fp, diags

    def electron_charge_density(self, f_dict):
        """vector_field.py:297-306."""
        cd = np.zeros_like(self.g["x"])
        if "electron" in f_dict:
            cd += self.sp["electron"]["charge"] * np.sum(f_dict["electron"], axis=1) * self.sg["electron"]["dv"]
        return cd

    def __call__(self, t, y, args=None):
        """vector_field.py:308-361."""
        g = self.g
        dex = [ex_driver_field(self.drivers_ex, g["x"], t + d) for d in self.dt_array]
        if self.ex_stochastic is not None:  # vector_field.py:290-295
            dex = [de + self.ex_stochastic(t + d, g["x"]) for de, d in zip(dex, self.dt_array)]
        djy = ey_driver_source(self.drivers_ey, g["x_a"], t + self.dt_array[1], self.c)
        nu_fp = self.nu_fp_prof(g["x"], t) if self.cfg["terms"]["fokker_planck"]["is_on"] else None
        nu_K = self.nu_K_prof(g["x"], t) if self.cfg["terms"]["krook"]["is_on"] else None
        f_dict = {k: v for k, v in y.items() if k in self.sg}
        n_n = self.electron_charge_density(f_dict)
        e, f_new, diags = self.vpfp(f_dict, y["a"], y["e"], dex, nu_fp, nu_K)
        n_np1 = self.electron_charge_density(f_new)
        a = wave_solver(y["a"], y["prev_a"], djy, -0.5 * (n_n + n_np1), self.c, g["dx"], self.dt)
        result = {"a": a["a"], "prev_a": a["prev_a"], "da": djy, "de": dex[self.dex_save], "e": e}
        result.update(f_new)
        result.update(diags)
        return result


# --------------------------------------------------------------------------------------
# save functions + time loop (adept/_vlasov1d/storage.py, adept/_base_.py:30-41, modules.py:338-358)
# --------------------------------------------------------------------------------------


def default_scalars(cfg, y):
    """storage.py:286-327 (get_default_save_func)."""
    g = cfg["grid"]
    s = {}
    ke = 0.0
    for name, sg in g["species_grids"].items():
        v, dv = sg["v"][None, :], sg["dv"]
        mass = g["species_params"][name]["mass"]
        f = y[name]

        def mm(inp):
            return np.mean(np.sum(inp, axis=1) * dv)

        s[f"mean_P_{name}"] = mm(f * v**2.0)
        s[f"mean_j_{name}"] = mm(f * v)
        s[f"mean_n_{name}"] = mm(f)
        s[f"mean_q_{name}"] = mm(f * v**3.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            s[f"mean_-flogf_{name}"] = mm(-np.log(np.abs(f)) * np.abs(f))
        s[f"mean_f2_{name}"] = mm(f * f)
        ke += 0.5 * mass * s[f"mean_P_{name}"]
    s["mean_de2"] = np.mean(y["de"] ** 2.0)
    s["mean_e2"] = np.mean(y["e"] ** 2.0)
    s["mean_pond"] = np.mean(ponderomotive(y["a"], g["dx"]))
    s["mean_kinetic_energy"] = ke
    s["mean_field_energy"] = 0.5 * s["mean_e2"]
    s["mean_total_energy"] = ke + 0.5 * s["mean_e2"]
    return s


def field_moments(cfg, y):
    """storage.py:119-162 (get_field_save_func)."""
    g = cfg["grid"]
    res = {}
    for name, sg in g["species_grids"].items():
        v, dv = sg["v"], sg["dv"]
        f = y[name]

        def mom(inp):
            return np.sum(inp, axis=1) * dv

        m = {}
        m["n"] = mom(f)
        m["j"] = mom(f * v[None, :])
        m["v"] = m["j"] / m["n"]
        vm = v[None, :] - m["v"][:, None]
        m["p"] = mom(f * vm**2.0)
        m["q"] = mom(f * vm**3.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            m["-flogf"] = mom(-np.abs(f) * np.log(np.abs(f)))
        m["f^2"] = mom(f * f)
        res[name] = m
    res["e"], res["de"], res["a"], res["prev_a"] = y["e"], y["de"], y["a"], y["prev_a"]
    res["pond"] = ponderomotive(y["a"], g["dx"])
    return res


def interp2d_linear(xq, yq, x, y, f):
    """interpax.interp2d(xq, yq, x, y, f, method="linear") with its defaults (extrap=False -> NaN outside the grid,
    no period), as called by get_dist_save_func (adept/_vlasov1d/storage.py:173-188).  interpax 0.3.12 is not installed
    here (PARITY UNPINNED for this function): restated from its published algorithm -- i = clip(searchsorted(x, xq,
    side="right"), 1, nx-1), same for j, bilinear blend of the four surrounding nodes divided by the cell area, queries
    outside [x[0], x[-1]] x [y[0], y[-1]] replaced by NaN."""
    xq, yq, x, y = (np.asarray(a, dtype=float) for a in (xq, yq, x, y))
    i = np.clip(np.searchsorted(x, xq, side="right"), 1, len(x) - 1)
    j = np.clip(np.searchsorted(y, yq, side="right"), 1, len(y) - 1)
    f00, f01, f10, f11 = f[i - 1, j - 1], f[i - 1, j], f[i, j - 1], f[i, j]
    x0, x1, y0, y1 = x[i - 1], x[i], y[j - 1], y[j]
    dx0, dx1, dy0, dy1 = xq - x0, x1 - xq, yq - y0, y1 - yq
    fq = (dx1 * (f00 * dy1 + f01 * dy0) + dx0 * (f10 * dy1 + f11 * dy0)) / ((x1 - x0) * (y1 - y0))
    outside = (xq < x[0]) | (xq > x[-1]) | (yq < y[0]) | (yq > y[-1])
    return np.where(outside, np.nan, fq)


def dist_save_xv(f, x, v, xax, vax):
    """get_dist_save_func for a {t, x, v} save block (storage.py:173-181): f interpolated on meshgrid(xax, vax, "ij")."""
    xq, vq = np.meshgrid(xax, vax, indexing="ij")
    return interp2d_linear(xq.ravel(), vq.ravel(), x, v, f).reshape(xq.shape)


def dist_save_kxv(f, kxr, v, kxax, vax):
    """get_dist_save_func for a {t, kx, v} save block (storage.py:183-190): abs(rfft(f, axis=x)) interpolated on
    meshgrid(kxax, vax, "ij").  The reference hands interp2d the nx-long two-sided axis cfg["grid"]["kx"] for an array
    with nx/2 + 1 rows (storage.py:259, 271), which interpax rejects; the one-sided axis kxr = 2 pi rfftfreq(nx, dx)
    that matches rfft's rows is used here (PARITY UNPINNED: the reference path cannot produce an array)."""
    fk = np.abs(np.fft.rfft(f, axis=0))
    kq, vq = np.meshgrid(kxax, vax, indexing="ij")
    return interp2d_linear(kq.ravel(), vq.ravel(), kxr, v, fk).reshape(kq.shape)


def save_axis(tcfg, grid):
    """storage.py:203-219 (_add_dim_axes for 't') + modules.py:166-181 defaults."""
    tmin = float(tcfg.get("tmin", grid["tmin"]))
    tmax = float(tcfg.get("tmax", grid["tmax"]))
    return np.linspace(tmin, tmax, int(tcfg["nt"]))


def run(cfg, nsteps=None, save=None, vf=None):
    """Fixed-step loop: y_{n+1} = vf(t_n, y_n) (adept/_base_.py:37-41), saves by linear interpolation
    between y_n and y_{n+1} (diffrax Euler dense output).

    ``save``: dict name -> (ts array, fn(cfg, y)).  Returns (final state, saved dict of lists).
    """
    g = cfg["grid"]
    vf = vf if vf is not None else VlasovMaxwell(cfg)
    y = init_state(cfg)
    dt = g["dt"]
    nsteps = g["nt"] if nsteps is None else nsteps
    save = save or {}
    out = {k: [] for k in save}
    cursor = {k: 0 for k in save}
    for n in range(nsteps):
        t0, t1 = n * dt, (n + 1) * dt
        y1 = vf(t0, y, None)
        for k, (ts, fn) in save.items():
            while cursor[k] < len(ts) and ts[cursor[k]] <= t1 + 1e-12 * max(1.0, abs(t1)):
                tq = ts[cursor[k]]
                w = (tq - t0) / (t1 - t0)
                yi = {kk: y[kk] + w * (y1[kk] - y[kk]) for kk in y1}
                out[k].append(fn(cfg, yi))
                cursor[k] += 1
        y = y1
    return y, out
