"""Ensembles of independent ``vlasov-1d`` runs advanced as ONE batched problem (BASELINE.json configs[3]: parameter scans
over k lambda_D and drive amplitude, training batches).

The reference runs such scans as separate processes (one MLflow run per parameter point, adept/_base_.py:359-429) or
under ``jax.vmap``; here the members share every kernel launch: the distribution is ``f[batch, nx, nv]`` and the
per-member differences travel as per-row device tables (box length -> ``k1x_batch`` and ``1/kx``; density profile ->
initial ``f`` and ion background; driver wavenumber / frequency / amplitude / envelope -> ``ex_kx, ex_w_row,
ex_a0_row, ex_space``; collision-frequency profile -> ``nu_fp_space``).  Members must agree on everything that is a
scalar of the step: grid sizes, ``dt``, velocity grids, the ``terms`` block, and the *time* envelopes of drivers and
collision frequencies.  Across GPUs an ensemble shards by members with no communication (one ``EnsembleVlasov1D`` per
rank on its slice of the deck list).

Every member's arithmetic is the single-run arithmetic (same kernels, same rounding), so ``member_state(i)`` equals the
state of ``Vlasov1D(decks[i])`` after the same number of steps; tests/test_gpu_ensemble.py checks that and the oracle.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, ops
from .config import build_cfg
from .vector_field import NativeStep, VlasovMaxwell


def _same(values, what):
    first = values[0]
    for v in values[1:]:
        if v != first:
            raise _lib.AdeptB200Error(f"ensemble members must share {what}: {first!r} != {v!r}")
    return first


class EnsembleVlasov1D:
    """``ens = EnsembleVlasov1D(decks); ens.run(nsteps); ens.state`` -- state tensors carry a leading member axis."""

    def __init__(self, decks, device="cuda"):
        if not torch.cuda.is_available():
            raise _lib.AdeptB200Error("adept_b200 needs a CUDA device: there is no CPU implementation of the time step")
        if len(decks) < 1:
            raise ValueError("empty ensemble")
        self.device = torch.device(device)
        built = [build_cfg(d) for d in decks]
        self.cfgs, self.grids = [b[0] for b in built], [b[1] for b in built]
        self.vms = [VlasovMaxwell(cfg, grid, device=device) for cfg, grid in built]
        vm0, cfg0 = self.vms[0], self.cfgs[0]
        B = self.B = len(decks)
        # ---- what must agree --------------------------------------------------------------------------------------
        self.nx = _same([int(c["grid"]["nx"]) for c in self.cfgs], "grid.nx")
        self.dt = _same([float(g.dt) for g in self.grids], "grid.dt")
        _same([(c["terms"]["time"], c["terms"]["edfdv"], c["terms"]["field"]) for c in self.cfgs], "terms")
        # everything _static_step takes from member 0: operator type, super-Gaussian exponent, the self-consistent beta
        # controls and the Krook Maxwellian (reference species T0 / mass); dx only enters through pond (a == 0 here)
        _same([(vm.fp_on, vm.krook_on, vm.vpfp.fp.model, vm.vpfp.fp.scheme, vm.vpfp.fp.nodrag, vm.vpfp.fp.m,
                vm.vpfp.fp.sc_steps, vm.vpfp.fp.sc_rtol, vm.vpfp.fp.sc_atol,
                hash(np.asarray(vm.vpfp.fp.f_mx, dtype=np.float64).tobytes())) for vm in self.vms],
              "the collision operators (type, m, self_consistent_beta, Krook Maxwellian)")
        self.names = _same([list(c["grid"]["species_grids"].keys()) for c in self.cfgs], "the species list")
        for name in self.names:
            _same([(len(c["grid"]["species_grids"][name]["v"]), float(c["grid"]["species_grids"][name]["v"][0]),
                    float(c["grid"]["species_grids"][name]["dv"])) for c in self.cfgs], f"the velocity grid of {name}")
            _same([(c["grid"]["species_params"][name]["charge"], c["grid"]["species_params"][name]["mass"])
                   for c in self.cfgs], f"charge and mass of {name}")
        if any(c.get("drivers", {}).get("ex_stochastic") is not None for c in self.cfgs):
            raise NotImplementedError("ensembles with the stochastic Ex driver are not implemented")
        if any(vm.has_ey for vm in self.vms):
            raise NotImplementedError("ensembles with transverse (Ey) drivers are not implemented")
        if (not all(NativeStep.supported(vm) for vm in self.vms) or cfg0["terms"]["field"] not in ("poisson",)
                or any(vm.vpfp.vlasov_dfdt or vm.vpfp.fp_dfdt or vm.vpfp.hou_li_filter_on for vm in self.vms)):
            raise NotImplementedError("ensembles need field=poisson and no dfdt diagnostics / Hou-Li filter")
        self.n_ex = _same([len(vm.ex_driver.drivers) for vm in self.vms], "the number of Ex drivers")
        if self.n_ex > _lib.MAX_DRIVERS:
            raise NotImplementedError(f"at most {_lib.MAX_DRIVERS} Ex drivers")
        for j in range(self.n_ex):
            _same([self._tkey(vm.ex_driver.drivers[j].envelope.time_envelope) for vm in self.vms],
                  f"the time envelope of Ex driver {j}")
        if vm0.fp_on:
            _same([self._tkey(vm.nu_fp_prof.time_envelope) for vm in self.vms], "the time envelope of nu_fp")
        if vm0.krook_on:
            _same([self._tkey(vm.nu_K_prof.time_envelope) for vm in self.vms], "the time envelope of nu_K")
        integ = vm0.vpfp.vlasov_poisson
        self.sixth = vm0.vpfp.dex_save == 3
        self.dt_array = [float(d) for d in integ.dt_array] if self.sixth else [0.0]
        self.edfdv = 0 if cfg0["terms"]["edfdv"] == "exponential" else 1
        # ---- state and per-member tables ------------------------------------------------------------------------------
        dev = self.device
        tt = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev)  # noqa: E731
        nx = self.nx
        self.state = {}
        for name in self.names:
            self.state[name] = tt(np.stack([np.asarray(c["grid"]["species_distributions"][name][1]) for c in self.cfgs]))
        for k in ("e", "de"):
            self.state[k] = torch.zeros((B, nx), dtype=torch.float64, device=dev)
        for k in ("a", "da", "prev_a"):
            self.state[k] = torch.zeros((B, nx + 2), dtype=torch.float64, device=dev)
        xs = [np.asarray(g.x) for g in self.grids]
        fsolve = [vm.vpfp.vlasov_poisson.field_solve for vm in self.vms]
        self.tab = {
            "k1x": tt([float(vm.vpfp.vlasov_poisson.vdfdx.k1x) for vm in self.vms]),
            "kmul": tt(np.stack([np.asarray(fs.kmul) for fs in fsolve])),
            "ion": tt(np.stack([np.broadcast_to(np.asarray(fs.static_charge_density, dtype=np.float64), (nx,))
                                if fs.static_charge_density is not None else np.zeros(nx) for fs in fsolve])),
            "f_mx": tt(vm0.vpfp.fp.f_mx),
        }
        for name in self.names:
            self.tab["v", name] = tt(cfg0["grid"]["species_grids"][name]["v"])
        if self.n_ex:
            drv = [[vm.ex_driver.drivers[j] for vm in self.vms] for j in range(self.n_ex)]
            self.tab["ex_space"] = tt(np.stack([np.concatenate([d.envelope.space_envelope(x) * np.ones(nx)
                                                                for d, x in zip(row, xs)]) for row in drv]))
            self.tab["ex_kx"] = tt(np.stack([np.concatenate([d.k0 * x for d, x in zip(row, xs)]) for row in drv]))
            self.tab["ex_w"] = tt(np.stack([np.concatenate([np.full(nx, d.w0 + d.dw0) for d in row]) for row in drv]))
            self.tab["ex_a0"] = tt(np.stack([np.concatenate([np.full(nx, d.a0) for d in row]) for row in drv]))
        if vm0.fp_on:
            self.tab["nu_fp"] = tt(np.concatenate([vm.nu_fp_prof.space_envelope(x) * np.ones(nx)
                                                   for vm, x in zip(self.vms, xs)]))
        if vm0.krook_on:
            self.tab["nu_K"] = tt(np.concatenate([vm.nu_K_prof.space_envelope(x) * np.ones(nx)
                                                  for vm, x in zip(self.vms, xs)]))
        self.scratch = {}
        self._st = None
        self.t, self.step_index = 0.0, 0

    @staticmethod
    def _tkey(env):
        return tuple(sorted((k, v) for k, v in vars(env).items() if isinstance(v, (int, float, bool))))

    def _scratch(self, key, shape):
        k = (key, tuple(shape))
        if k not in self.scratch:
            self.scratch[k] = torch.empty(shape, dtype=torch.float64, device=self.device)
        return self.scratch[k]

    def _static_step(self):
        """struct adept_b200_step with everything that does not change from step to step (built once)."""
        vm0, cfg0, B, nx = self.vms[0], self.cfgs[0], self.B, self.nx
        g0 = cfg0["grid"]
        n = B * nx
        st = _lib.Step()
        st.batch, st.nx, st.n_species = B, nx, len(self.names)
        for k, name in enumerate(self.names):
            sg, sp = g0["species_grids"][name], g0["species_params"][name]
            f = self.state[name]
            s = st.species[k]
            if self.edfdv == 1:
                s.f_tmp = self._scratch(("tmp", name), f.shape).data_ptr()
            s.v = self.tab["v", name].data_ptr()
            s.nv, s.dv, s.k1v = int(f.shape[-1]), float(sg["dv"]), float(sg["kvr"][1])
            s.charge, s.mass = float(sp["charge"]), float(sp["mass"])
            nparts = ops.vdfdx_rho_parts(f)
            s.rho_parts, s.rho_nparts = self._scratch(("parts", name), (nparts, n)).data_ptr(), nparts
        st.electron_species = self.names.index("electron") if "electron" in self.names else -1
        st.collide_species = self.names.index(vm0.vpfp.fp.ref_species)
        st.time_integrator, st.edfdv, st.field = int(self.sixth), self.edfdv, 0
        st.dt, st.dx, st.k1x = self.dt, float(self.grids[0].dx), float(self.tab["k1x"][0])
        st.k1x_batch = self.tab["k1x"].data_ptr()
        st.ion_charge = self.tab["ion"].data_ptr()
        st.kmul, st.kmul_stride = self.tab["kmul"].data_ptr(), nx
        st.c_light, st.wave_on = float(vm0.c), 0
        st.pond, st.rho = self._scratch("pond", (n,)).data_ptr(), self._scratch("rho", (n,)).data_ptr()
        st.n_ex = self.n_ex
        if self.n_ex:
            st.ex_space, st.ex_kx = self.tab["ex_space"].data_ptr(), self.tab["ex_kx"].data_ptr()
            st.ex_w_row, st.ex_a0_row = self.tab["ex_w"].data_ptr(), self.tab["ex_a0"].data_ptr()
        fp = vm0.vpfp.fp
        st.fp_on, st.krook_on = int(vm0.fp_on), int(vm0.krook_on)
        st.fp_model, st.fp_scheme, st.fp_nodrag = fp.model, fp.scheme, int(fp.nodrag)
        st.sg_m, st.sg_ratio = fp.m, fp.sg_ratio
        st.fp_sc_steps, st.fp_sc_rtol, st.fp_sc_atol = fp.sc_steps, fp.sc_rtol, fp.sc_atol
        if vm0.fp_on:
            st.nu_fp_space = self.tab["nu_fp"].data_ptr()
        if vm0.krook_on:
            st.nu_K_space = self.tab["nu_K"].data_ptr()
        st.f_mx = self.tab["f_mx"].data_ptr()
        return st

    def step(self):
        """One step of every member: a single ``adept_b200_step_f64`` call with ``batch = len(decks)``."""
        vm0, B, nx, t = self.vms[0], self.B, self.nx, self.t
        if self._st is None:
            self._st = self._static_step()
        st = self._st
        new = {}
        for k, name in enumerate(self.names):
            f = self.state[name]
            out = torch.empty_like(f)
            new[name] = out
            st.species[k].f_in, st.species[k].f_out = f.data_ptr(), out.data_ptr()
        e_out = torch.empty_like(self.state["e"])
        st.e_in, st.e_out = self.state["e"].data_ptr(), e_out.data_ptr()
        dex = torch.empty((len(self.dt_array), B, nx), dtype=torch.float64, device=self.device)
        st.dex = dex.data_ptr()
        st.a, st.prev_a = self.state["a"].data_ptr(), self.state["prev_a"].data_ptr()
        for i, dti in enumerate(self.dt_array):
            st.ex_t[i] = t + dti
            for j, d in enumerate(vm0.ex_driver.drivers):
                st.ex_tenv[i][j] = float(d.envelope.time_envelope(t + dti))
        if vm0.fp_on:
            st.nu_fp_time = float(vm0.nu_fp_prof.time_envelope(t))
        if vm0.krook_on:
            st.nu_K_time = float(vm0.nu_K_prof.time_envelope(t))
        lib = _lib.load()
        before = lib.adept_b200_launch_count()
        rc = lib.adept_b200_step_f64(C.byref(st), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "ensemble step")
        ops.LAUNCHES += lib.adept_b200_launch_count() - before
        self.state = {"a": self.state["a"], "prev_a": self.state["a"], "da": self.state["da"],
                      "de": dex[vm0.vpfp.dex_save], "e": e_out, **new}
        self.step_index += 1
        self.t = self.step_index * self.dt
        return self.state

    def run(self, nsteps):
        for _ in range(nsteps):
            self.step()
        return self.state

    def member_state(self, i):
        """State dict of member ``i`` (views into the batched tensors), keyed like ``Vlasov1D.state``."""
        return {k: v[i] for k, v in self.state.items()}


# ---------------------------------------------------------------------------------------------- sharding across GPUs
def member_slice(n_members: int, rank: int, world: int) -> slice:
    """Members owned by ``rank``: contiguous blocks whose sizes differ by at most one (the first ``n % world`` ranks
    take one more).  Independent members: no data-path collective (ensembles)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(n_members), world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def run_sharded_ensemble(decks, nsteps, diagnostics, make=None, group=None):
    """Advance ``decks`` split over the ranks of ``torch.distributed`` (one process per GPU; works unsharded when the
    process group is not initialised).  Each rank builds ``make(decks[member_slice(...)])`` (default:
    ``EnsembleVlasov1D``), runs ``nsteps`` and evaluates ``diagnostics(ens) -> float64 tensor [members_here, ...]``;
    the only communication is the final gather of those small per-member results, returned on every rank in deck order
    (``all_gather`` on padded blocks, so it runs on NCCL and on gloo).  A rank whose slice is empty takes part in the
    gather only."""
    import torch.distributed as dist

    sharded = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if sharded else 1
    rank = dist.get_rank(group) if sharded else 0
    sl = member_slice(len(decks), rank, world)
    mine = list(decks[sl])
    out = None
    if mine:
        ens = (make or EnsembleVlasov1D)(mine)
        ens.run(nsteps)
        out = diagnostics(ens)
        if out.shape[0] != len(mine):
            raise ValueError("diagnostics must return one row per member of this rank")
    if not sharded:
        return out
    counts = [member_slice(len(decks), r, world) for r in range(world)]
    counts = [c.stop - c.start for c in counts]
    # shape of one member's result: agreed through rank 0's (every rank with members has the same trailing shape)
    trailing = [list(out.shape[1:])] if out is not None else [None]
    shapes = [None] * world
    dist.all_gather_object(shapes, trailing[0], group=group)
    tshape = next(s for s in shapes if s is not None)
    dev = out.device if out is not None else (torch.device("cuda", torch.cuda.current_device())
                                              if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    pad = torch.zeros((max(counts),) + tuple(tshape), dtype=torch.float64, device=dev)
    if out is not None:
        pad[: out.shape[0]] = out.to(torch.float64)
    blocks = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(blocks, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(blocks, counts)], dim=0)
