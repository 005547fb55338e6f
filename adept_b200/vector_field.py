"""Step composition for the vlasov-1d path: integrators, Vlasov-Poisson-Fokker-Planck step and the VlasovMaxwell
vector field, with the same class names, constructor arguments and call signatures as the reference's
``adept/_vlasov1d/solvers/vector_field.py`` (:19-361).  The state is a dict of CUDA tensors with the reference's keys
(species names, ``e, de [nx]``, ``a, da, prev_a [nx+2]``, optional ``diag-*``); ``VlasovMaxwell.__call__(t, y, args)``
is a map y -> y' exactly like the reference's (adept/_base_.py:37-41).
"""

from __future__ import annotations

import os

import numpy as np
import torch

import ctypes as C

from . import _lib, ops, pushers


def _is_parallel(parallel, axis: str) -> bool:
    if not parallel:
        return False
    return axis in parallel


class TimeIntegrator:
    """Shared field solver and Vlasov pushers; vector_field.py:19-52."""

    def __init__(self, cfg: dict, grid):
        self.field_solve = pushers.ElectricFieldSolver(cfg, grid)
        self.species_grids = cfg["grid"]["species_grids"]
        self.species_params = cfg["grid"]["species_params"]
        parallel = cfg["grid"].get("parallel", False)
        self.edfdv = self.get_edfdv(cfg, parallel)
        self.vdfdx = pushers.SpaceExponential(grid.x, self.species_grids, parallel=_is_parallel(parallel, "v"))

    def get_edfdv(self, cfg: dict, parallel):
        kind = cfg["terms"]["edfdv"]
        if kind == "exponential":
            return pushers.VelocityExponential(self.species_grids, self.species_params,
                                               parallel=_is_parallel(parallel, "x"))
        if kind == "cubic-spline":
            return pushers.VelocityCubicSpline(self.species_grids, self.species_params,
                                               parallel=_is_parallel(parallel, "x"))
        raise NotImplementedError(f"{kind} has not been implemented")


class LeapfrogIntegrator(TimeIntegrator):
    """x-push, field solve on f*, v-push; vector_field.py:55-95."""

    def __init__(self, cfg: dict, grid):
        super().__init__(cfg, grid)
        self.dt = grid.dt
        self.dt_array = self.dt * np.array([0.0, 1.0])

    def __call__(self, f_dict, a, dex_array, prev_ex):
        if self.field_solve.hampere:
            f_after_v, parts = self.vdfdx.push_with_rho(f_dict, dt=self.dt)
            pond, e = self.field_solve.solve_hampere(f_dict, a, prev_ex, self.dt, parts)
        elif self.field_solve.wants_rho:
            # the x-push kernel accumulates sum_v f* on the fly: the field solve does not read f* again
            f_after_v, parts = self.vdfdx.push_with_rho(f_dict, dt=self.dt)
            pond, e = self.field_solve(f_dict=f_after_v, a=a, prev_ex=prev_ex, dt=self.dt, rho_parts=parts)
        else:
            f_after_v = self.vdfdx(f_dict, dt=self.dt)
            pond, e = self.field_solve(f_dict=f_after_v, a=a, prev_ex=prev_ex, dt=self.dt)
        # e + dex[0] is formed inside the push kernel (same rounding as the reference's explicit sum)
        f_out = self.edfdv(f_after_v, e=e, pond=pond, dt=self.dt, dex=dex_array[0])
        return e, f_out


class SixthOrderHamIntegrator(TimeIntegrator):
    """6th-order Hamiltonian splitting: 6 x (field solve + v-push) interleaved with 5 x-pushes; :98-186."""

    def __init__(self, cfg: dict, grid):
        super().__init__(cfg, grid)
        self.dt = grid.dt
        self.a1 = 0.168735950563437422448196
        self.a2 = 0.377851589220928303880766
        self.a3 = -0.093175079568731452657924
        b1 = 0.049086460976116245491441
        b2 = 0.264177609888976700200146
        b3 = 0.186735929134907054308413
        c1 = -0.000069728715055305084099
        c2 = -0.000625704827430047189169
        c3 = -0.002213085124045325561636
        d2 = -2.916600457689847816445691e-6
        d3 = 3.048480261700038788680723e-5
        e3 = 4.985549387875068121593988e-7
        self.D1 = b1 + 2.0 * c1 * self.dt**2.0
        self.D2 = b2 + 2.0 * c2 * self.dt**2.0 + 4.0 * d2 * self.dt**4.0
        self.D3 = b3 + 2.0 * c3 * self.dt**2.0 + 4.0 * d3 * self.dt**4.0 - 8.0 * e3 * self.dt**6.0
        self.dt_array = self.dt * np.array(
            [
                0.0,
                self.a1,
                self.a1 + self.a2,
                self.a1 + self.a2 + self.a3,
                self.a1 + self.a2 + self.a3 + self.a2,
                self.a1 + self.a2 + self.a3 + self.a2 + self.a1,
            ]
        )

    def __call__(self, f_dict, a, dex_array, prev_ex):
        Ds = (self.D1, self.D2, self.D3, self.D3, self.D2, self.D1)
        As = (self.a1, self.a2, self.a3, self.a2, self.a1)
        e, parts = None, None
        fuse = self.field_solve.wants_rho
        for i in range(6):
            pond, e = self.field_solve(f_dict=f_dict, a=a, prev_ex=None, dt=None, rho_parts=parts)
            f_dict = self.edfdv(f_dict, e=e, pond=pond, dt=Ds[i] * self.dt, dex=dex_array[i])
            if i < 5:
                if fuse:  # the next field solve's density comes out of this x-push
                    f_dict, parts = self.vdfdx.push_with_rho(f_dict, dt=As[i] * self.dt)
                else:
                    f_dict = self.vdfdx(f_dict, dt=As[i] * self.dt)
        return e, f_dict


class VlasovPoissonFokkerPlanck:
    """integrator -> collisions -> (filter) -> dfdt diagnostics; vector_field.py:189-253."""

    def __init__(self, cfg: dict, grid):
        self.dt = grid.dt
        if cfg["terms"]["time"] == "sixth":
            self.vlasov_poisson = SixthOrderHamIntegrator(cfg, grid)
            self.dex_save = 3
        elif cfg["terms"]["time"] == "leapfrog":
            self.vlasov_poisson = LeapfrogIntegrator(cfg, grid)
            self.dex_save = 0
        else:
            raise NotImplementedError
        self.fp = pushers.Collisions(cfg=cfg)
        self.vlasov_dfdt = cfg["diagnostics"]["diag-vlasov-dfdt"]
        self.fp_dfdt = cfg["diagnostics"]["diag-fp-dfdt"]
        hl = cfg["terms"].get("hou_li_filter")
        self.hou_li_filter_on = bool(hl and hl.get("is_on", False))
        if self.hou_li_filter_on:
            self.hou_li_filter = pushers.HouLiFilter(nx=cfg["grid"]["nx"], alpha=hl["alpha"], order=hl["order"])

    def __call__(self, f_dict, a, prev_ex, dex_array, nu_fp, nu_K, n_out=None):
        e, f_vlasov = self.vlasov_poisson(f_dict, a, dex_array, prev_ex)
        f_fp = self.fp(nu_fp, nu_K, f_vlasov, dt=self.dt, n_out=n_out)
        if self.hou_li_filter_on:
            f_fp = self.hou_li_filter(f_fp)
        diags = {}
        ref_species = "electron" if "electron" in f_dict else next(iter(f_dict))
        if self.vlasov_dfdt:
            diags["diag-vlasov-dfdt"] = (f_vlasov[ref_species] - f_dict[ref_species]) / self.dt
        if self.fp_dfdt:
            diags["diag-fp-dfdt"] = (f_fp[ref_species] - f_vlasov[ref_species]) / self.dt
        return e, f_fp, diags


class NativeStep:
    """Fills ``struct adept_b200_step`` and calls ``adept_b200_step_f64``: the whole step is enqueued by native code.

    Built once per VlasovMaxwell; per step only the O(1) time factors (driver envelopes / phases, collision-frequency
    time envelopes) are evaluated on the host.  The dfdt diagnostics and the Hou-Li filter are part of the native
    step; VlasovMaxwell composes the step from the operator objects only for the Hamiltonian Ampere solver and when
    the caller supplies the per-step inputs itself."""

    FIELD = {"poisson": 0, "poisson-boltzmann": 1, "ampere": 2}

    def __init__(self, vm):
        self.vm = vm
        cfg, grid = vm.cfg, vm.grid
        self.cfg = cfg
        self.names = list(cfg["grid"]["species_grids"].keys())
        integ = vm.vpfp.vlasov_poisson
        self.sixth = vm.vpfp.dex_save == 3
        self.dt_array = [float(d) for d in integ.dt_array] if self.sixth else [0.0]
        self.dt_a1 = float(integ.dt_array[1])
        self.edfdv = 0 if cfg["terms"]["edfdv"] == "exponential" else 1
        self.field = self.FIELD[cfg["terms"]["field"]]
        self.dev = {}
        self.scratch = {}
        self.static = {}

    @staticmethod
    def supported(vm) -> bool:
        cfg = vm.cfg
        return (cfg["terms"]["field"] in NativeStep.FIELD and cfg["terms"]["edfdv"] in ("exponential", "cubic-spline")
                and len(cfg["grid"]["species_grids"]) <= _lib.MAX_SPECIES
                and len(vm.ex_driver.drivers) <= _lib.MAX_DRIVERS)

    def _table(self, key, make, device):
        """Device-resident constant table, built once from ``make()`` (a host array)."""
        k = (key, str(device))
        if k not in self.dev:
            self.dev[k] = torch.as_tensor(np.array(make(), dtype=np.float64, order="C"), device=device)
        return self.dev[k]

    def _scratch(self, key, shape, device, zero=False, dtype=torch.float64):
        k = (key, tuple(shape), str(device))
        if k not in self.scratch:
            self.scratch[k] = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
        return self.scratch[k]

    def _static_step(self, y, dev, batch, wave_on):
        """struct adept_b200_step with every field that does not change from step to step (tables, scratch, physics
        switches), built once per (device, batch, wave_on, shapes) and reused: per step only pointers to the state and
        the O(1) time factors are rewritten."""
        shapes = tuple(tuple(y[name].shape) for name in self.names)
        key = (str(dev), batch, bool(wave_on), shapes)
        if key in self.static:
            return self.static[key]
        vm, cfg, grid = self.vm, self.cfg, self.vm.grid
        g = cfg["grid"]
        nx = int(g["nx"])
        n = batch * nx
        st = _lib.Step()
        st.batch, st.nx, st.n_species = batch, nx, len(self.names)
        for k, name in enumerate(self.names):
            sg, sp = g["species_grids"][name], g["species_params"][name]
            f = y[name]
            s = st.species[k]
            if self.edfdv == 1 or ops.is_long_mixed_nx(nx, int(f.shape[-1])):  # spline stencil / long-pencil scratch
                s.f_tmp = self._scratch(("tmp", name), f.shape, dev).data_ptr()
            s.v = self._table(("v", name), lambda sg=sg: sg["v"], dev).data_ptr()
            s.nv, s.dv, s.k1v = int(f.shape[-1]), float(sg["dv"]), float(sg["kvr"][1])
            s.charge, s.mass = float(sp["charge"]), float(sp["mass"])
            nparts = ops.vdfdx_rho_parts(f)
            s.rho_parts = self._scratch(("parts", name), (nparts, n), dev).data_ptr()
            s.rho_nparts = nparts
        st.electron_species = self.names.index("electron") if "electron" in self.names else -1
        st.collide_species = self.names.index(vm.vpfp.fp.ref_species)
        st.time_integrator, st.edfdv, st.field = int(self.sixth), self.edfdv, self.field
        st.dt, st.dx = float(grid.dt), float(grid.dx)
        st.k1x = float(vm.vpfp.vlasov_poisson.vdfdx.k1x)
        fs = vm.vpfp.vlasov_poisson.field_solve
        if self.field == 0:
            if fs.static_charge_density is not None:
                st.ion_charge = self._table(("ion", batch), lambda: np.broadcast_to(
                    np.asarray(fs.static_charge_density, dtype=np.float64), (batch, nx)), dev).data_ptr()
            st.kmul = self._table("kmul", lambda: fs.kmul, dev).data_ptr()
            if batch == 1 and os.environ.get("ADEPT_B200_FIELD_TAIL", "1") != "0":
                # Green's function of E = Re ifft(-i (1/kx) fft(rho)) (field.py:221-224): the x-advection launch of a
                # large one-species grid solves the field in its tail as a circular convolution with it
                st.poisson_green = self._table("green", lambda: np.real(np.fft.ifft(
                    -1j * np.asarray(fs.kmul, dtype=np.float64))), dev).data_ptr()
        elif self.field == 1:
            st.kmul = self._table("kmul", lambda: fs.kmul, dev).data_ptr()
            st.Te, st.lambda_De = fs.Te, fs.lambda_De
        st.kmul_stride = 0
        st.c_light, st.wave_on = float(vm.c), int(bool(wave_on))
        st.pond = self._scratch("pond", (n,), dev).data_ptr()
        st.rho = self._scratch("rho", (n,), dev).data_ptr()
        if wave_on:
            st.ne_n = self._scratch("ne_n", (n,), dev).data_ptr()
            st.ne_np1 = self._scratch("ne_np1", (n,), dev).data_ptr()
        # longitudinal drivers: space factors resident on the device, time factors filled in per step
        drivers = vm.ex_driver.drivers
        st.n_ex = len(drivers)
        if drivers:
            x = vm._x
            st.ex_space = self._table(("ex_space", batch), lambda: np.stack(
                [np.broadcast_to(d.envelope.space_envelope(x), (batch, nx)).reshape(-1) for d in drivers]), dev).data_ptr()
            st.ex_kx = self._table(("ex_kx", batch), lambda: np.stack(
                [np.broadcast_to(d.k0 * x, (batch, nx)).reshape(-1) for d in drivers]), dev).data_ptr()
            for j, d in enumerate(drivers):
                st.ex_w[j], st.ex_a0[j] = d.w0 + d.dw0, d.a0
        # collisions
        fp = vm.vpfp.fp
        st.fp_on, st.krook_on = int(vm.fp_on), int(vm.krook_on)
        st.fp_model, st.fp_scheme, st.fp_nodrag = fp.model, fp.scheme, int(fp.nodrag)
        st.sg_m, st.sg_ratio = fp.m, fp.sg_ratio
        st.fp_sc_steps, st.fp_sc_rtol, st.fp_sc_atol = fp.sc_steps, fp.sc_rtol, fp.sc_atol
        if vm.fp_on:
            st.nu_fp_space = self._table(("nu_fp", batch), lambda: np.broadcast_to(
                vm.nu_fp_prof.space_envelope(vm._x) * np.ones(nx), (batch, nx)), dev).data_ptr()
        if vm.krook_on:
            st.nu_K_space = self._table(("nu_K", batch), lambda: np.broadcast_to(
                vm.nu_K_prof.space_envelope(vm._x) * np.ones(nx), (batch, nx)), dev).data_ptr()
        st.f_mx = self._table("f_mx", lambda: fp.f_mx, dev).data_ptr()
        if vm.vpfp.hou_li_filter_on:
            st.hou_li_filt = self._table("hou_li", lambda: vm.vpfp.hou_li_filter.filter_x, dev).data_ptr()
        ref = "electron" if "electron" in self.names else self.names[0]  # vector_field.py:243-244
        st.diag_species = self.names.index(ref)
        # one device word per integration, zeroed once; the library resets it after every use
        st.sync_counter = self._scratch("sync_counter", (4,), dev, zero=True, dtype=torch.int32).data_ptr()
        self.static[key] = st
        return st

    def time_row(self, t):
        """The O(1) time factors of one step starting at time t as a host row of ``_lib.TIME_ROW_LEN`` doubles
        (layout of include/adept_b200.h: tenv[s][d], wt[s][d], nu_fp_time, nu_K_time, ex_t[s]) -- the same closed forms
        ``__call__`` evaluates per step (field.py:21-33, functions.py:72-80,112-118)."""
        vm = self.vm
        row = np.zeros(_lib.TIME_ROW_LEN, dtype=np.float64)
        for j, d in enumerate(vm.ex_driver.drivers):
            for i, dti in enumerate(self.dt_array):
                ti = t + dti
                row[8 * i + j] = float(d.envelope.time_envelope(ti))
                row[48 + 8 * i + j] = d.phase(ti)
        if vm.fp_on:
            row[96] = float(vm.nu_fp_prof.time_envelope(t))
        if vm.krook_on:
            row[97] = float(vm.nu_K_prof.time_envelope(t))
        for i, dti in enumerate(self.dt_array):
            row[98 + i] = t + dti
        return row

    def __call__(self, t, y, wave_on):
        vm = self.vm
        f0 = y[self.names[0]]
        dev = f0.device
        batch = f0.shape[0] if f0.dim() == 3 else 1
        st = self._static_step(y, dev, batch, wave_on)
        new = {}
        for k, name in enumerate(self.names):
            f = y[name]
            if not (f.is_cuda and f.dtype == torch.float64 and f.is_contiguous()):
                raise _lib.AdeptB200Error(f"state['{name}'] must be a contiguous float64 CUDA tensor")
            out = torch.empty_like(f)
            new[name] = out
            st.species[k].f_in, st.species[k].f_out = f.data_ptr(), out.data_ptr()
        e_out = torch.empty_like(y["e"])
        st.e_in, st.e_out = y["e"].data_ptr(), e_out.data_ptr()
        dex = torch.empty((len(self.dt_array),) + tuple(y["e"].shape), dtype=torch.float64, device=dev)
        st.dex = dex.data_ptr()
        st.a, st.prev_a = y["a"].data_ptr(), y["prev_a"].data_ptr()
        if wave_on:
            djy = vm.ey_driver(t + self.dt_a1, None) if vm.has_ey else vm._zeros_a
            a_out = torch.empty_like(y["a"])
            st.djy, st.a_out = djy.data_ptr(), a_out.data_ptr()
        else:
            djy, a_out = vm._zeros_a, None
        for j, d in enumerate(vm.ex_driver.drivers):
            w = d.w0 + d.dw0
            for i, dti in enumerate(self.dt_array):
                ti = t + dti
                st.ex_tenv[i][j] = float(d.envelope.time_envelope(ti))
                st.ex_wt[i][j] = d.phase(ti)
        if vm.fp_on:
            st.nu_fp_time = float(vm.nu_fp_prof.time_envelope(t))
        if vm.krook_on:
            st.nu_K_time = float(vm.nu_K_prof.time_envelope(t))
        diags = {}
        ref_f = y[self.names[st.diag_species]]
        for on, key, field in ((vm.vpfp.vlasov_dfdt, "diag-vlasov-dfdt", "diag_vlasov_dfdt"),
                               (vm.vpfp.fp_dfdt, "diag-fp-dfdt", "diag_fp_dfdt")):
            if on:
                diags[key] = torch.empty_like(ref_f)
                setattr(st, field, diags[key].data_ptr())
        lib = _lib.load()
        before = lib.adept_b200_launch_count()
        rc = lib.adept_b200_step_f64(C.byref(st), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "step")
        ops.LAUNCHES += lib.adept_b200_launch_count() - before  # kernels the native step enqueued (exact)
        result = {"a": a_out if wave_on else y["a"], "prev_a": y["a"], "da": djy,
                  "de": dex[vm.vpfp.dex_save], "e": e_out}
        result.update(new)
        result.update(diags)
        return result

    def vjp(self, t, y, f_out_bar, e_out_bar=None, want=("dex", "nu_fp", "nu_K")):
        """Reverse mode through one native step (adept_b200_step_bwd_f64): re-runs the forward step from ``y`` at time
        ``t`` (its field arrays feed the backward pass), then returns ``(y_new, bars)`` with bars = {"f": cotangent of
        the input distribution, "dex": cotangent of the driver field [nx], "nu_fp" / "nu_K": cotangents of nu(x, t)}.
        One species, leapfrog, poisson, LB / Dougherty (+ Krook), no transverse wave -- anything else raises."""
        y_new = self(t, y, False)
        name = self.names[0]
        f = y[name]
        dev = f.device
        batch = f.shape[0] if f.dim() == 3 else 1
        st = self._static_step(y, dev, batch, False)  # still holds this step's pointers and time factors
        n = batch * int(self.cfg["grid"]["nx"])
        bw = _lib.StepBwd()
        g = f_out_bar.contiguous()
        f_in_bar = torch.empty_like(f)
        bars = {"f": f_in_bar}
        bw.f_out_bar, bw.f_in_bar = g.data_ptr(), f_in_bar.data_ptr()
        if e_out_bar is not None:
            eb = e_out_bar.contiguous()
            bw.e_out_bar = eb.data_ptr()
        for key, field, on in (("dex", "dex_bar", True), ("nu_fp", "nu_fp_bar", self.vm.fp_on),
                               ("nu_K", "nu_K_bar", self.vm.krook_on)):
            if key in want and on:
                bars[key] = torch.zeros(y["e"].shape, dtype=torch.float64, device=dev)
                setattr(bw, field, bars[key].data_ptr())
        for i in range(3):
            bw.scratch_f[i] = self._scratch(("bwd_f", i), f.shape, dev).data_ptr()
        for i in range(2):
            bw.scratch_row[i] = self._scratch(("bwd_row", i), (n,), dev).data_ptr()
        lib = _lib.load()
        before = lib.adept_b200_launch_count()
        rc = lib.adept_b200_step_bwd_f64(C.byref(st), C.byref(bw), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "step_bwd")
        ops.LAUNCHES += lib.adept_b200_launch_count() - before
        return y_new, bars


class VlasovMaxwell:
    """One full vlasov-1d step y -> y'; vector_field.py:256-361."""

    def __init__(self, cfg, grid, drivers=None, nu_fp_prof=None, nu_K_prof=None, device="cuda"):
        from .functions import SpaceTimeEnvelopeFunction

        self.cfg, self.grid, self.device = cfg, grid, device
        self.vpfp = VlasovPoissonFokkerPlanck(cfg, grid)
        beta = cfg["grid"]["beta"]
        c = 1.0 / beta
        self.c = c
        self.wave_solver = pushers.WaveSolver(c=c, dx=grid.dx, dt=grid.dt)
        self.dt = grid.dt
        dcfg = cfg.get("drivers", {"ex": {}, "ey": {}})
        if drivers is None:
            ex = [pushers.EMDriver.from_config(d, c) for d in dcfg.get("ex", {}).values()]
            ey = [pushers.EMDriver.from_config(d, c) for d in dcfg.get("ey", {}).values()]
            if dcfg.get("ex_stochastic") is not None:  # simulation.py:160-168, vector_field.py:290-295
                ex = ex + pushers.StochasticDriver(dcfg["ex_stochastic"], grid).modes()
        else:
            ex, ey = drivers["ex"], drivers["ey"]
        self.ey_driver = pushers.TransverseCurrentSourceDriver(grid.x_a, drivers=ey, c=c, device=device)
        self.ex_driver = pushers.LongitudinalElectricFieldDriver(grid.x, drivers=ex, device=device)
        self.has_ey = len(ey) > 0
        fpc, kc = cfg["terms"]["fokker_planck"], cfg["terms"]["krook"]
        self.fp_on, self.krook_on = bool(fpc["is_on"]), bool(kc["is_on"])
        self.nu_fp_prof = nu_fp_prof if nu_fp_prof is not None else (
            SpaceTimeEnvelopeFunction.from_config(fpc) if self.fp_on else None)
        self.nu_K_prof = nu_K_prof if nu_K_prof is not None else (
            SpaceTimeEnvelopeFunction.from_config(kc) if self.krook_on else None)
        self._x = np.asarray(grid.x)
        self._dev_cache = {}
        self._zeros_a = torch.zeros(grid.nx + 2, dtype=torch.float64, device=device)
        self._a_live = None
        self.native = NativeStep(self) if NativeStep.supported(self) else None

    def _dev(self, arr):
        return torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64), device=self.device)

    def _check_a_live(self, y):
        """Whether the wave equation has to run: an Ey driver, or a nonzero a / prev_a.  The answer is cached only while
        the caller feeds back the very tensors the previous step returned; a state that comes from elsewhere (restart,
        another trajectory through the same object) is inspected again."""
        seen = getattr(self, "_a_seen", None)
        if self._a_live is None or seen is None or y["a"] is not seen[0] or y["prev_a"] is not seen[1]:
            self._a_live = self.has_ey or bool(torch.any(y["a"] != 0)) or bool(torch.any(y["prev_a"] != 0))
            self._a_seen = (y["a"], y["prev_a"])

    def total_dex(self, t, args=None):
        return self.ex_driver(t, args)

    def _nu(self, prof, t, key):
        """nu(x, t) = time_env(t) * space_env(x): the space factor lives on the device, the time factor is a scalar."""
        if key not in self._dev_cache:
            self._dev_cache[key] = self._dev(prof.space_envelope(self._x))
        return self._dev_cache[key] * float(prof.time_envelope(t))

    def host_inputs(self, t):
        """Everything one step needs besides the state, evaluated on the host (numpy, O(nx)): the driver fields at the
        substep times and the collision-frequency profiles (vector_field.py:319-331).  A caller that keeps time on the
        host copies these to the device and passes them as ``args`` to ``__call__``."""
        t = float(t)
        dt_array = self.vpfp.vlasov_poisson.dt_array
        n_dex = 1 if self.vpfp.dex_save == 0 else len(dt_array)
        out = {"dex": np.stack([self.ex_driver.host(t + float(d)) for d in dt_array[:n_dex]])}
        out["djy"] = self.ey_driver.host(t + float(dt_array[1]))
        if self.fp_on:
            out["nu_fp"] = self.nu_fp_prof(self._x, t) * np.ones_like(self._x)
        if self.krook_on:
            out["nu_K"] = self.nu_K_prof(self._x, t) * np.ones_like(self._x)
        return out

    def compute_electron_charge_density(self, f_dict):
        """q_e sum_v f_e dv (zeros if there is no electron species); vector_field.py:297-306."""
        if "electron" not in f_dict:
            return torch.zeros(self.grid.nx, dtype=torch.float64, device=self.device)
        f = f_dict["electron"]
        dv = float(self.cfg["grid"]["species_grids"]["electron"]["dv"])
        q = self.cfg["grid"]["species_params"]["electron"]["charge"]
        out = torch.empty(f.shape[:-1], dtype=torch.float64, device=f.device)
        # reference: 0 + q * (sum * dv)
        ops.moments(f, None, dv, (out, None, None), scale_b=(q, 1.0, 1.0))
        return out

    def __call__(self, t, y, args=None):
        t = float(t)
        if self.native is not None and (args is None or "dex" not in args):
            # With no Ey driver and a == prev_a == 0 the wave update returns exactly 0 whatever the density is
            # (field.py:149-153), so the two density reductions and the wave kernel are skipped.  Checked once.
            self._check_a_live(y)
            out = self.native(t, y, self._a_live)
            self._a_seen = (out["a"], out["prev_a"])
            return out
        dt_array = self.vpfp.vlasov_poisson.dt_array
        n_dex = 1 if self.vpfp.dex_save == 0 else len(dt_array)  # leapfrog only ever reads dex[0]
        if args is not None and "dex" in args:
            # per-step inputs supplied by the caller as device tensors (see host_inputs): dex [n_dex, nx], ...
            dex = list(args["dex"])
            djy = args["djy"] if self.has_ey else self._zeros_a
            nu_fp = args.get("nu_fp") if self.fp_on else None
            nu_K = args.get("nu_K") if self.krook_on else None
        else:
            dex = [self.total_dex(t + float(d), args) for d in dt_array[:n_dex]]
            djy = self.ey_driver(t + float(dt_array[1]), args) if self.has_ey else self._zeros_a
            nu_fp = self._nu(self.nu_fp_prof, t, "fp") if self.fp_on else None
            nu_K = self._nu(self.nu_K_prof, t, "K") if self.krook_on else None
        species = self.cfg["grid"]["species_grids"]
        f_dict = {k: v for k, v in y.items() if k in species}

        # With no Ey driver and a == prev_a == 0 the wave update returns exactly 0 whatever the density is
        # (field.py:149-153), so the two density reductions and the wave kernel are skipped.  Checked once.
        self._check_a_live(y)
        need_wave = self._a_live
        if not need_wave:
            e, f_new, diags = self.vpfp(f_dict=f_dict, a=y["a"], prev_ex=y["e"], dex_array=dex, nu_fp=nu_fp,
                                        nu_K=nu_K)
            result = {"a": y["a"], "prev_a": y["a"], "da": djy, "de": dex[self.vpfp.dex_save], "e": e}
            result.update(f_new)
            result.update(diags)
            self._a_seen = (result["a"], result["prev_a"])
            return result
        ne_n = self.compute_electron_charge_density(f_dict) if need_wave else None
        e, f_new, diags = self.vpfp(f_dict=f_dict, a=y["a"], prev_ex=y["e"], dex_array=dex, nu_fp=nu_fp, nu_K=nu_K)
        ne_np1 = self.compute_electron_charge_density(f_new) if need_wave else None
        a = self.wave_solver(a=y["a"], aold=y["prev_a"], djy_array=djy, electron_density_n=ne_n,
                             electron_density_np1=ne_np1)
        result = {"a": a["a"], "prev_a": a["prev_a"], "da": djy, "de": dex[self.vpfp.dex_save], "e": e}
        result.update(f_new)
        result.update(diags)
        self._a_seen = (result["a"], result["prev_a"])
        return result
