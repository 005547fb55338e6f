"""Drop-in operator objects for the vlasov-1d path, same names / constructor arguments / call signatures as the
reference's pushers, backed by the sm_100a kernels of libadept_b200.so.

  SpaceExponential, VelocityExponential, VelocityCubicSpline, HouLiFilter   adept/_vlasov1d/solvers/pushers/vlasov.py
  ElectricFieldSolver (Poisson / Boltzmann-Poisson / Ampere), WaveSolver,
  LongitudinalElectricFieldDriver, TransverseCurrentSourceDriver            adept/_vlasov1d/solvers/pushers/field.py
  Collisions (+ Krook)                                                      adept/_vlasov1d/solvers/pushers/fokker_planck.py

Distributions are float64 CUDA tensors ``f[nx, nv]`` (or ``[batch, nx, nv]``); 1-D fields are CUDA tensors.  Grid
arrays may be numpy arrays or tensors.  There is no CPU path: calling with CPU tensors raises AdeptB200Error.
"""

from __future__ import annotations

import math

import numpy as np
import torch
from scipy.special import gammaln

from . import ops
from ._lib import AdeptB200Error


def _np(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


def _dev(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=device)


def _device_of(f_dict):
    return next(iter(f_dict.values())).device


class _DeviceCache:
    """Keeps device copies of small host tables (v grids, 1/kx, ...) per CUDA device."""

    def __init__(self):
        self._c = {}

    def get(self, key, host_array, device):
        k = (key, str(device))
        if k not in self._c:
            self._c[k] = _dev(host_array, device)
        return self._c[k]


# ------------------------------------------------------------------------------------------------ Vlasov pushers
class SpaceExponential:
    """x-advection f <- irfft(exp(-i kx v dt) rfft(f, axis=0), axis=0); vlasov.py:223-251."""

    def __init__(self, x, species_grids, parallel=False):
        xh = _np(x)
        self.nx = len(xh)
        self.kx_real = np.fft.rfftfreq(self.nx, d=float(xh[1] - xh[0])) * 2 * np.pi
        self.k1x = float(self.kx_real[1])
        self.species_grids = species_grids
        self.parallel = parallel
        self._cache = _DeviceCache()
        self._parts = {}

    def push(self, f, v, dt, out=None):
        return ops.vdfdx(f, v, dt, self.k1x, out=out)

    def __call__(self, f_dict, dt, out=None):
        result = {}
        for name, f in f_dict.items():
            v = self._cache.get(name, self.species_grids[name]["v"], f.device)
            result[name] = self.push(f, v, float(dt), None if out is None else out.get(name))
        return result

    def push_with_rho(self, f_dict, dt, out=None):
        """Same push, fused with the first stage of the velocity sum of the result: returns (f_dict, parts_dict) where
        parts_dict[name] is the per-CTA partial-sum scratch that ``ElectricFieldSolver(..., rho_parts=...)`` finishes
        (SpaceExponential followed by compute_charge_density, vlasov.py:234-251 + field.py:197-208)."""
        result, parts = {}, {}
        for name, f in f_dict.items():
            v = self._cache.get(name, self.species_grids[name]["v"], f.device)
            key = (name, tuple(f.shape), str(f.device))
            if key not in self._parts:
                n = f.numel() // f.shape[-1]
                self._parts[key] = torch.empty((ops.vdfdx_rho_parts(f), n), dtype=torch.float64, device=f.device)
            parts[name] = self._parts[key]
            result[name] = ops.vdfdx_rho(f, v, float(dt), self.k1x, parts[name],
                                         out=None if out is None else out.get(name))
        return result, parts


class _VelocityPusher:
    def __init__(self, species_grids, species_params, parallel=False):
        self.species_grids = species_grids
        self.species_params = species_params
        self.parallel = parallel


class VelocityExponential(_VelocityPusher):
    """Spectral v-advection under q e + (q^2/m) pond; vlasov.py:63-103."""

    def __call__(self, f_dict, e, pond, dt, dex=None, out=None):
        result = {}
        for name, f in f_dict.items():
            kvr = self.species_grids[name]["kvr"]
            k1v = float(kvr[1])
            q, m = self.species_params[name]["charge"], self.species_params[name]["mass"]
            result[name] = ops.edfdv_exp(f, e, pond, q, m, float(dt), k1v, dex=dex,
                                         out=None if out is None else out.get(name))
        return result

    push = __call__


class VelocityCubicSpline(_VelocityPusher):
    """Semi-Lagrangian v-advection with the local cubic Hermite stencil; vlasov.py:106-184."""

    def __call__(self, f_dict, e, pond, dt, dex=None, out=None):
        result = {}
        for name, f in f_dict.items():
            dv = float(self.species_grids[name]["dv"])
            q, m = self.species_params[name]["charge"], self.species_params[name]["mass"]
            result[name] = ops.edfdv_spline(f, e, pond, q, m, float(dt), dv, dex=dex,
                                            out=None if out is None else out.get(name))
        return result

    push = __call__


# ------------------------------------------------------------------------------------------------ field solve
class ElectricFieldSolver:
    """(pond, e) = field_solve(f_dict, a, prev_ex, dt); field.py:422-497.

    ``poisson`` / ``poisson-boltzmann`` / ``ampere`` / ``hampere`` (single species, leapfrog).
    """

    def __init__(self, cfg: dict, grid):
        g = cfg["grid"]
        self.species_grids = g["species_grids"]
        self.species_params = g["species_params"]
        self.static_charge_density = g.get("ion_charge")
        self.kind = cfg["terms"]["field"]
        time = cfg["terms"]["time"]
        self.hampere = False
        if self.kind == "poisson":
            self.kmul, self.mode = _np(grid.one_over_kx), 0
        elif self.kind == "poisson-boltzmann":
            bz = cfg["terms"].get("boltzmann_electrons") or {}
            self.kmul, self.mode = _np(grid.kx), 1
            self.Te = float(bz.get("Te", 1.0))
            lam = bz.get("lambda_De")
            self.lambda_De = -1.0 if lam is None else float(lam)
        elif self.kind == "ampere":
            if time != "leapfrog":
                raise NotImplementedError(f"ampere + {time} has not yet been implemented")
        elif self.kind == "hampere":
            if time != "leapfrog":
                raise NotImplementedError(f"hampere + {time} has not yet been implemented")
            if len(self.species_grids) > 1:
                raise NotImplementedError(
                    "HampereSolver currently only supports single-species simulations. "
                    "For multi-species, use 'ampere' or 'poisson' field solver instead."
                )
            self.hampere = True
            self.kmul, self.mode = _np(grid.one_over_kx), 0
            self.kx = _np(grid.kx)
        else:
            raise NotImplementedError("Field Solver: <" + str(self.kind) + "> has not yet been implemented")
        self.dx = float(grid.dx)
        self._cache = _DeviceCache()

    def compute_charge_density(self, f_dict, out=None, rho_parts=None):
        """rho = sum_s q_s dv_s sum_v f_s (+ static ion background); field.py:186-208.

        ``rho_parts`` (from ``SpaceExponential.push_with_rho``) supplies the velocity sums already accumulated by the
        x-push kernel, so f is not read again."""
        dev = _device_of(f_dict)
        rho = None
        if self.static_charge_density is not None and self.kind == "poisson":
            # (0 + q n) + static == static + q n: the background rides along as the base of the first reduction
            rho = self._cache.get("ion", self.static_charge_density, dev)
            if next(iter(f_dict.values())).dim() == 3:
                rho = rho.expand(next(iter(f_dict.values())).shape[0], -1).contiguous()
        for name, f in f_dict.items():
            q, dv = self.species_params[name]["charge"], float(self.species_grids[name]["dv"])
            new = torch.empty(f.shape[:-1], dtype=torch.float64, device=dev) if out is None else out
            if rho_parts is not None:
                ops.reduce_parts(rho_parts[name], dv, q, base=None if rho is None else rho.reshape(-1),
                                 out=new.reshape(-1))
            else:
                ops.moments(f, None, dv, (new, None, None), bases=(rho, None, None), scale_b=(q, 1.0, 1.0))
            rho = new
        return rho

    def compute_current_density(self, f_dict):
        """j = sum_s q_s dv_s sum_v v f_s; field.py:319-340."""
        dev = _device_of(f_dict)
        j = None
        for name, f in f_dict.items():
            q, dv = self.species_params[name]["charge"], float(self.species_grids[name]["dv"])
            v = self._cache.get("v-" + name, self.species_grids[name]["v"], dev)
            new = torch.empty(f.shape[:-1], dtype=torch.float64, device=dev)
            ops.moments(f, v, dv, (None, new, None), bases=(None, j, None), scale_b=(1.0, q, 1.0))
            j = new
        return j

    def solve_hampere(self, f_dict, a, prev_ex, dt, rho_parts):
        """HampereSolver.__call__ (field.py:395-419) without a 2-D FFT of f.

        E_k += q/(i k) dv sum_v f_k(v) (e^{-i k v dt} - 1) is, mode by mode, the Poisson integral of the change of
        charge density under the x-advection, q dv (sum_v f* - sum_v f): f* is what the x-push computes anyway and its
        velocity sum comes out of that kernel (``rho_parts``).  Only the Nyquist mode differs -- irfft drops the
        imaginary part of its phase while the full complex transform of the reference keeps it -- and is added
        explicitly: (-1)^x / nx * q / k_N * dv sum_x (-1)^x sum_v f(x, v) (-sin(k_N v dt))."""
        name = next(iter(f_dict))
        f = f_dict[name]
        dev = f.device
        q, dv = self.species_params[name]["charge"], float(self.species_grids[name]["dv"])
        nx = f.shape[-2]
        k_ny = float(self.kx[nx // 2])
        w = self._cache.get(("hampere-w", float(dt)), -np.sin(k_ny * float(dt) * _np(self.species_grids[name]["v"])), dev)
        n0 = torch.empty(f.shape[:-1], dtype=torch.float64, device=dev)
        s1 = torch.empty(f.shape[:-1], dtype=torch.float64, device=dev)
        ops.moments(f, w, dv, (n0, s1, None), scale_b=(-q, 1.0, 1.0))  # n0 = -q dv sum f ; s1 = dv sum f w
        drho = ops.reduce_parts(rho_parts[name], dv, q, base=n0.reshape(-1))  # q dv (sum f* - sum f)
        de = ops.poisson(drho.reshape(n0.shape), self._cache.get("kmul", self.kmul, dev))
        sign = self._cache.get(("sign", nx), (-1.0) ** np.arange(nx), dev)
        s_ny = torch.sum(sign * s1, dim=-1, keepdim=True)
        e = prev_ex + de + sign * (q * float(self.kmul[nx // 2]) / nx) * s_ny
        return ops.ponderomotive(a, self.dx), e

    @property
    def wants_rho(self):
        """True when the field solve consumes the charge density (so the x-push should accumulate it)."""
        return self.kind != "ampere"

    def __call__(self, f_dict, a, prev_ex, dt, rho_parts=None):
        dev = _device_of(f_dict)
        pond = ops.ponderomotive(a, self.dx)
        if self.kind == "ampere":
            e = ops.axpy(prev_ex, self.compute_current_density(f_dict), -float(dt))
        else:
            rho = self.compute_charge_density(f_dict, rho_parts=rho_parts)
            kmul = self._cache.get("kmul", self.kmul, dev)
            if self.mode == 0:
                e = ops.poisson(rho, kmul)
            else:
                e = ops.poisson(rho, kmul, mode=1, Te=self.Te, lambda_De=self.lambda_De)
        return pond, e


class WaveSolver:
    """Leap-frog wave equation for a[nx+2] with 2nd-order absorbing boundaries; field.py:94-157."""

    def __init__(self, c, dx, dt):
        self.c, self.dx, self.dt = float(c), float(dx), float(dt)

    def __call__(self, a, aold, djy_array, electron_density_n=None, electron_density_np1=None):
        """electron_density_{n,np1}: electron CHARGE densities before/after the step (the kernel forms
        -0.5 (rho_n + rho_np1) itself, vector_field.py:346); None means zero density."""
        if self.c > 0:
            anew = ops.wave_step(a, aold, djy_array, electron_density_n, electron_density_np1, self.c, self.dx,
                                 self.dt)
            return {"a": anew, "prev_a": a}
        return {"a": a, "prev_a": aold}


# ------------------------------------------------------------------------------------------------ drivers (O(nx), torch)
class EMDriver:
    """a0, k0, w0, dw0 + space-time envelope; simulation.py:37-93 (a0/k0/w0 parametrisation)."""

    def __init__(self, a0, k0, w0, dw0, envelope, is_point_source=False):
        self.a0, self.k0, self.w0, self.dw0 = float(a0), float(k0), float(w0), float(dw0)
        self.envelope = envelope
        self.is_point_source = bool(is_point_source)

    def phase(self, t):
        """Temporal phase of the carrier, (w0 + dw0) t (field.py:31)."""
        return (self.w0 + self.dw0) * t

    @staticmethod
    def from_config(cfg: dict, c_light: float) -> "EMDriver":
        from .functions import SpaceTimeEnvelopeFunction

        p = cfg["params"]
        if "intensity" in p:
            raise NotImplementedError("adept_b200: intensity/wavelength drivers need pint; give a0/k0/w0 instead")
        k0, w0 = p.get("k0"), p.get("w0")
        if k0 is None and w0 is None:
            raise ValueError("You must specify at least one of k0 or w0.")
        if k0 is None:
            k0 = w0 / c_light
        elif w0 is None:
            w0 = c_light * k0
        return EMDriver(p["a0"], k0, w0, p.get("dw0", 0.0), SpaceTimeEnvelopeFunction.from_config(cfg["envelope"]),
                        cfg.get("source_type", "extended") == "point")


class StochasticDriver:
    """Ornstein-Uhlenbeck amplitudes of a few box modes, drawn once on the host with numpy's Generator exactly as the
    reference does (simulation.py:95-148), linearly interpolated in time."""

    def __init__(self, scfg: dict, grid):
        modes = [int(m) for m in scfg.get("modes", [1])]
        if not modes or any(m < 1 for m in modes):  # datamodel.py:177-182
            raise ValueError(f"stochastic driver modes must be positive integers, got {modes}")
        amplitude, tau = float(scfg["amplitude"]), float(scfg["tau"])
        if tau <= 0:  # datamodel.py:184-189
            raise ValueError(f"stochastic driver correlation time tau must be > 0, got {tau}")
        dt_update = float(scfg["dt_update"]) if scfg.get("dt_update") is not None else tau / 10.0
        dt_update = min(dt_update, tau / 2.0)
        nt = int(np.ceil((grid.tmax - grid.tmin) / dt_update)) + 2
        rng = np.random.default_rng(int(scfg.get("seed", 42)))
        theta = np.exp(-dt_update / tau)
        kick = amplitude * np.sqrt(1.0 - theta**2)
        n = len(modes)
        amps = np.zeros((nt, n), dtype=np.complex128)
        amps[0] = amplitude * (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2.0)
        for it in range(1, nt):
            xi = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2.0)
            amps[it] = theta * amps[it - 1] + kick * xi
        self.t_grid = grid.tmin + dt_update * np.arange(nt)
        self.amp_real, self.amp_imag = amps.real.copy(), amps.imag.copy()
        self.k_modes = 2.0 * np.pi * np.asarray(modes, dtype=np.float64) / (grid.xmax - grid.xmin)

    def modes(self):
        """One pseudo driver per mode for LongitudinalElectricFieldDriver / the native step: the mode's field
        ar cos(kx) - ai sin(kx) is A sin(kx - phase) with A = hypot(ar, ai), phase = -atan2(ar, -ai)."""
        return [_StochasticMode(self, j) for j in range(len(self.k_modes))]


class _Ones:
    def __call__(self, x):
        return np.ones_like(np.asarray(x, dtype=np.float64))


class _StochasticMode:
    """Quacks like EMDriver (a0 = w0 = 1): time_envelope(t) is the mode's modulus, phase(t) minus its argument."""

    is_point_source = False

    def __init__(self, parent, j):
        self.a0, self.w0, self.dw0, self.k0 = 1.0, 1.0, 0.0, float(parent.k_modes[j])
        self._p, self._j = parent, j
        self.envelope = self
        self.space_envelope = _Ones()

    def _amp(self, t):
        p, j = self._p, self._j
        return float(np.interp(t, p.t_grid, p.amp_real[:, j])), float(np.interp(t, p.t_grid, p.amp_imag[:, j]))

    def time_envelope(self, t):
        ar, ai = self._amp(t)
        return math.hypot(ar, ai)

    def phase(self, t):
        ar, ai = self._amp(t)
        return -math.atan2(ar, -ai)

    def __call__(self, x, t):  # envelope(x, t)
        return self.time_envelope(t) * np.ones_like(np.asarray(x, dtype=np.float64))


class LongitudinalElectricFieldDriver:
    """E_D(x, t) = sum env(x,t) (w0+dw0) a0 sin(k0 x - (w0+dw0) t); field.py:13-33.  Returns a CUDA tensor."""

    def __init__(self, xax, drivers, device="cuda"):
        self.xax = _np(xax)
        self.drivers = list(drivers)
        self.device = device
        # the spatial factors are time independent: keep them on the device
        self._space = [_dev(d.envelope.space_envelope(self.xax), device) for d in self.drivers]
        self._kx = [_dev(d.k0 * self.xax, device) for d in self.drivers]
        self._zero = torch.zeros(len(self.xax), dtype=torch.float64, device=device)

    def __call__(self, t, args=None):
        total = self._zero.clone()
        for d, sp, kx in zip(self.drivers, self._space, self._kx):
            w = d.w0 + d.dw0
            amp = float(d.envelope.time_envelope(t)) * w * d.a0
            total += torch.sin(kx - d.phase(t)) * sp * amp
        return total

    def host(self, t):
        """Same field evaluated on the host (numpy), used to pre-tabulate many steps at once."""
        total = np.zeros_like(self.xax)
        for d in self.drivers:
            w = d.w0 + d.dw0
            total += d.envelope(self.xax, t) * w * d.a0 * np.sin(d.k0 * self.xax - d.phase(t))
        return total


class TransverseCurrentSourceDriver:
    """Source of the transverse wave equation (extended or point sources); field.py:36-91."""

    def __init__(self, xax, drivers, c=0.0, device="cuda"):
        self.xax = _np(xax)
        self.drivers = list(drivers)
        self.c = float(c)
        self.device = device
        self.dx = float(self.xax[1] - self.xax[0])

    def host(self, t):
        total = np.zeros_like(self.xax)
        for d in self.drivers:
            w = d.w0 + d.dw0
            if d.is_point_source:
                i0 = int(np.argmin(np.abs(self.xax - d.envelope.space_envelope.center)))
                mask = np.zeros_like(self.xax)
                mask[i0] = 1.0
                F0 = 2.0 * w * self.c * d.a0
                total += (F0 / self.dx) * d.envelope.time_envelope(t) * mask * np.sin(w * t)
            else:
                total += -d.envelope(self.xax, t) * w**2 * d.a0 * np.sin(d.k0 * self.xax - w * t)
        return total

    def __call__(self, t, args=None):
        return _dev(self.host(t), self.device)


# ------------------------------------------------------------------------------------------------ collisions
_FP_TYPES = {
    "lenard_bernstein": (0, 0),
    "chang_cooper": (0, 1),
    "lenard_bernstein_chang_cooper": (0, 1),
    "chang_cooper_dougherty": (1, 1),
    "dougherty_chang_cooper": (1, 1),
    "dougherty": (1, 0),
    "super_gaussian": (2, 1),
    "super_gaussian_chang_cooper": (2, 1),
    "dougherty_nodrag": (1, 0),
}


def _collision_species(cfg) -> str:
    sg = cfg["grid"]["species_grids"]
    return "electron" if "electron" in sg else next(iter(sg))


class Collisions:
    """f <- Krook(FokkerPlanck(f)); fokker_planck.py:272-484.  ``__call__(nu_fp, nu_K, f, dt)`` accepts a dict of
    species (only the reference species is collided) or a bare tensor, and ``None`` frequencies."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.ref_species = _collision_species(cfg)
        fp_cfg = cfg["terms"]["fokker_planck"]
        fp_type = fp_cfg.get("type", "").casefold()
        self.fp_on = bool(fp_cfg["is_on"])
        if fp_type in _FP_TYPES:
            self.model, self.scheme = _FP_TYPES[fp_type]
        elif self.fp_on or fp_type:
            raise NotImplementedError(f"Unknown Fokker-Planck type: {fp_type}")
        else:
            self.model, self.scheme = 1, 0
        self.nodrag = fp_type == "dougherty_nodrag"
        self.m = float(fp_cfg.get("m", 2.0))
        self.sg_ratio = float(math.exp(gammaln(3.0 / self.m) - gammaln(1.0 / self.m)))
        sc = fp_cfg.get("self_consistent_beta", {})
        self.sc_steps = int(sc.get("max_steps", 3)) if sc.get("enabled", False) else 0  # fokker_planck.py:296-301
        self.sc_rtol, self.sc_atol = float(sc.get("rtol", 1e-8)), float(sc.get("atol", 1e-12))
        sg = cfg["grid"]["species_grids"][self.ref_species]
        self.v = _np(sg["v"])
        self.dv = float(sg["dv"])
        self.krook_on = bool(cfg["terms"]["krook"]["is_on"])
        params = cfg["grid"].get("species_params", {}).get(self.ref_species, {})
        T0, mass = params.get("T0", 1.0), params.get("mass", 1.0)
        f_mx = np.exp(-(self.v[None, :] ** 2.0) / (2.0 * T0 / mass))
        self.f_mx = (f_mx / np.sum(f_mx, axis=1)[:, None] / self.dv)[0]
        self._cache = _DeviceCache()

    def __call__(self, nu_fp, nu_K, f, dt, n_out=None, out=None):
        if isinstance(f, dict):
            return {k: (self._apply(nu_fp, nu_K, fs, dt, n_out, None if out is None else out.get(k))
                        if k == self.ref_species else fs) for k, fs in f.items()}
        return self._apply(nu_fp, nu_K, f, dt, n_out, out)

    def _apply(self, nu_fp, nu_K, f, dt, n_out=None, out=None):
        use_fp = self.fp_on and nu_fp is not None
        use_k = self.krook_on and nu_K is not None
        if self.fp_on and nu_fp is None:
            nu_fp, use_fp = torch.zeros(f.shape[:-1], dtype=torch.float64, device=f.device), True
        if self.krook_on and nu_K is None:
            nu_K, use_k = torch.zeros(f.shape[:-1], dtype=torch.float64, device=f.device), True
        if not use_fp and not use_k and n_out is None:
            return f
        v = self._cache.get("v", self.v, f.device)
        f_mx = self._cache.get("f_mx", self.f_mx, f.device)
        return ops.collide(f, v, self.dv, float(dt), nu_fp=nu_fp if use_fp else None, nu_K=nu_K if use_k else None,
                           f_mx=f_mx, model=self.model, scheme=self.scheme, nodrag=self.nodrag, sg_m=self.m,
                           sg_ratio=self.sg_ratio, n_out=n_out, out=out, sc_steps=self.sc_steps, sc_rtol=self.sc_rtol,
                           sc_atol=self.sc_atol)


class HouLiFilter:
    """x-only Hou-Li spectral filter sigma_j = exp(-alpha (j / (nx/2))^(2 order)); vlasov.py:187-220."""

    def __init__(self, nx: int, alpha: float, order: int):
        j_x = np.arange(nx // 2 + 1)
        eta_x = j_x / (nx // 2)
        self.filter_x = np.exp(-alpha * eta_x ** (2 * order))
        self._cache = _DeviceCache()

    def __call__(self, f_dict: dict) -> dict:
        result = {}
        for name, f in f_dict.items():
            filt = self._cache.get("filt", self.filter_x, f.device)
            zeros = self._cache.get(("z", f.shape[-1]), np.zeros(f.shape[-1]), f.device)
            result[name] = ops.filter_x(f, filt, zeros)
        return result


__all__ = ["SpaceExponential", "VelocityExponential", "VelocityCubicSpline", "ElectricFieldSolver", "WaveSolver",
           "EMDriver", "LongitudinalElectricFieldDriver", "TransverseCurrentSourceDriver", "Collisions",
           "HouLiFilter", "AdeptB200Error"]
