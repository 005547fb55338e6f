"""Single oversized grid sharded over the GPUs of one node.

Mirrors the reference's ``grid.parallel: ["x", "v"]`` decomposition (docs/source/solvers/vlasov1d/overview.md:155-177;
shard_map call sites pushers/vlasov.py:95-101,245-248, fokker_planck.py:435-441): the x-advection runs on a v-sharded
distribution ``f[nx, nv/P]`` (x-pencils are local), the v-advection and the collisions on an x-sharded one
``f[nx/P, nv]`` (v-rows are local).  Where XLA inserts the collectives implicitly, they are explicit here:

* ``all_reduce(SUM)`` of the nx-length partial velocity sums that the x-push kernel accumulates (32 KiB at nx=4096),
  after which every rank solves the O(nx) field equation redundantly;
* one ``all_to_all`` transpose v-sharded -> x-sharded before the v-push and one back after the collisions; each moves
  ``nx nv 8 (P-1)/P^2`` bytes per rank.  The send side of the forward transpose and the receive side of the backward
  one need no packing (row blocks of a v-shard are contiguous); the other two sides are one local permute-copy each.

On NVLink-connected GPUs the two transposes do not exist as separate operations (``transpose="p2p"``, the default
when the shapes allow it): every buffer stays v-sharded, and the fused v-advection + collision kernel of a rank reads
each cell of its rows straight from the rank that owns the column and writes the result straight back, through peer
memory mapped with ``torch.distributed._symmetric_memory`` (``adept_b200_vpush_collide_p2p_f64``).  A row touches one
contiguous ``8 nv/P``-byte segment per peer, so the link sees large transfers that overlap the kernel's math; the
x-advection stays local.  The rho all-reduce after the x-advection and a one-word all-reduce after the v-advection
order the ranks.  The all-to-all path (``transpose="nccl"``) remains for the other shapes and operators.

One process per GPU, ``torch.distributed`` (NCCL on the B200 box; the same code runs over gloo on CPU tensors when a
CPU operator table is injected -- that is how tests/test_sharded.py covers the N > 1 logic without a GPU).  State between
steps is kept v-sharded.  Scope: leapfrog, poisson, exponential or cubic-spline v-push, Fokker-Planck + Krook, no
transverse wave (``a == 0``); anything else raises.  Requires ``nx % P == 0`` and every species' ``nv % P == 0``
(overview.md:176-177).
"""

from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from ._lib import AdeptB200Error


class CudaOps:
    """Local operators backed by libadept_b200.so (the product path)."""

    def __init__(self):
        from . import ops

        self.ops = ops
        self._parts = {}

    def vdfdx_rowsum(self, f, v, dt, k1x):
        """(x-pushed f, sum over the local columns of the result [nx])."""
        ops = self.ops
        key = tuple(f.shape)
        if key not in self._parts:
            self._parts[key] = torch.empty((ops.vdfdx_rho_parts(f), f.shape[0]), dtype=torch.float64, device=f.device)
        out = ops.vdfdx_rho(f, v, dt, k1x, self._parts[key])
        return out, ops.reduce_parts(self._parts[key], 1.0, 1.0)

    def rho_from_sum(self, total, dv, q, base):
        return self.ops.reduce_parts(total.reshape(1, -1), dv, q, base=base)

    def poisson(self, rho, one_over_kx):
        return self.ops.poisson(rho, one_over_kx)

    def edfdv(self, kind, f, e, dex, q, m, dt, k1v, dv):
        if kind == "exponential":
            return self.ops.edfdv_exp(f, e, None, q, m, dt, k1v, dex=dex)
        return self.ops.edfdv_spline(f, e, None, q, m, dt, dv, dex=dex)

    def collide(self, f, v, dv, dt, nu_fp, nu_K, f_mx, model, scheme, nodrag, sg_m, sg_ratio):
        return self.ops.collide(f, v, dv, dt, nu_fp=nu_fp, nu_K=nu_K, f_mx=f_mx, model=model, scheme=scheme,
                                nodrag=nodrag, sg_m=sg_m, sg_ratio=sg_ratio)


class ShardedVlasov1D:
    """``sim = ShardedVlasov1D(deck); sim.step()``; rank r owns velocity columns ``[r nv/P, (r+1) nv/P)``."""

    def __init__(self, deck: dict, group=None, device=None, local_ops=None, transpose="auto"):
        from .config import build_cfg
        from .functions import SpaceTimeEnvelopeFunction
        from .pushers import Collisions, EMDriver

        if not dist.is_initialized():
            raise AdeptB200Error("ShardedVlasov1D needs an initialised torch.distributed process group")
        self.group = group
        self.P, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if local_ops is None:
            if not torch.cuda.is_available():
                raise AdeptB200Error("adept_b200 needs a CUDA device: there is no CPU implementation of the time step")
            local_ops = CudaOps()
            device = device or torch.device("cuda", torch.cuda.current_device())
        self.lops = local_ops
        self.device = device or torch.device("cpu")
        self.cfg, self.grid = build_cfg(deck)
        cfg, t = self.cfg, self.cfg["terms"]
        if t["time"] != "leapfrog" or t["field"] != "poisson":
            raise NotImplementedError("sharded grid: only time=leapfrog with field=poisson is implemented")
        if t["edfdv"] not in ("exponential", "cubic-spline"):
            raise NotImplementedError(f"{t['edfdv']} has not been implemented")
        if cfg["drivers"].get("ey") or cfg["diagnostics"].get("diag-vlasov-dfdt") or cfg["diagnostics"].get("diag-fp-dfdt"):
            raise NotImplementedError("sharded grid: Ey drivers and dfdt diagnostics are not implemented")
        hl = t.get("hou_li_filter")
        if hl and hl.get("is_on", False):
            raise NotImplementedError("sharded grid: terms.hou_li_filter is not implemented (it would be dropped)")
        if cfg["drivers"].get("ex_stochastic") is not None:
            raise NotImplementedError("sharded grid: drivers.ex_stochastic is not implemented (it would be dropped)")
        g = cfg["grid"]
        self.nx = int(g["nx"])
        if self.nx % self.P:
            raise AdeptB200Error(f"nx={self.nx} is not divisible by the number of ranks {self.P}")
        self.nxp = self.nx // self.P
        self.names = list(g["species_grids"].keys())
        self.nvp = {}
        for name in self.names:
            nv = int(g["species_grids"][name]["nv"])
            if nv % self.P:
                raise AdeptB200Error(f"nv={nv} of species {name} is not divisible by the number of ranks {self.P}")
            self.nvp[name] = nv // self.P
        x = np.asarray(self.grid.x)
        self.k1x = float(2 * np.pi * np.fft.rfftfreq(self.nx, d=float(x[1] - x[0]))[1])
        dev = self.device
        tt = lambda a: torch.as_tensor(np.array(a, dtype=np.float64), device=dev)  # noqa: E731
        self.v_full = {n: tt(g["species_grids"][n]["v"]) for n in self.names}
        self.v_loc = {n: self.v_full[n][self.rank * self.nvp[n]:(self.rank + 1) * self.nvp[n]].contiguous()
                      for n in self.names}
        self.one_over_kx = tt(self.grid.one_over_kx)
        self.ion = None if g.get("ion_charge") is None else tt(g["ion_charge"])
        c = 1.0 / g["beta"]
        self.ex = [EMDriver.from_config(d, c) for d in cfg["drivers"].get("ex", {}).values()]
        self.x = x
        rows = slice(self.rank * self.nxp, (self.rank + 1) * self.nxp)
        self.rows = rows
        self.coll = Collisions(cfg)
        if self.coll.sc_steps:
            raise NotImplementedError("sharded runs do not take terms.fokker_planck.self_consistent_beta yet")
        self.fp_on, self.krook_on = bool(t["fokker_planck"]["is_on"]), bool(t["krook"]["is_on"])
        self.nu_fp_prof = SpaceTimeEnvelopeFunction.from_config(t["fokker_planck"]) if self.fp_on else None
        self.nu_K_prof = SpaceTimeEnvelopeFunction.from_config(t["krook"]) if self.krook_on else None
        self.f_mx = tt(self.coll.f_mx)
        # v-sharded initial state
        self.state = {}
        for n in self.names:
            f0 = np.asarray(g["species_distributions"][n][1])
            self.state[n] = tt(f0[:, self.rank * self.nvp[n]:(self.rank + 1) * self.nvp[n]])
        self.state["e"] = torch.zeros(self.nx, dtype=torch.float64, device=dev)
        self.state["de"] = torch.zeros(self.nx, dtype=torch.float64, device=dev)
        self.t, self.step_index = 0.0, 0
        if transpose not in ("auto", "nccl", "p2p"):
            raise ValueError(f"transpose={transpose!r}")
        self.p2p = None
        if transpose != "nccl" and isinstance(local_ops, CudaOps):
            why = self._p2p_unsupported()
            if why is None:
                self._setup_p2p()
            elif transpose == "p2p":
                raise AdeptB200Error(f"transpose='p2p' is not available here: {why}")

    # ---- peer-memory transposes -----------------------------------------------------------------------------------
    def _p2p_unsupported(self):
        """None when the fused-transpose kernels cover this deck, else the reason."""
        t = self.cfg["terms"]
        P, nx = self.P, self.nx
        if len(self.names) != 1:
            return "more than one species"
        nv = self.nvp[self.names[0]] * P
        if P & (P - 1) or P > 8:
            return "number of ranks is not a power of two <= 8"
        if t["edfdv"] != "exponential" or not self.fp_on or self.krook_on:
            return "needs the spectral v-push with Fokker-Planck collisions and no Krook operator"
        if self.coll.model not in (0, 1) or self.coll.nodrag or self.coll.sc_steps:
            return "needs Lenard-Bernstein / Dougherty collisions (central or Chang-Cooper), no self-consistent beta"
        if nx & (nx - 1) or not 256 <= nx <= 4096 or (nv // P) % 4:
            return "x-advection shape is not handled by the TMA kernel"
        if nv & (nv - 1) or not 512 <= nv <= 8192 or (nx // P) % 2:
            return "v-advection shape is not handled by the fused kernel"
        return None

    def _setup_p2p(self):
        import torch.distributed._symmetric_memory as symm

        name = self.names[0]
        nvp, nv = self.nvp[name], self.nvp[name] * self.P
        group = self.group if self.group is not None else dist.group.WORLD
        f_vs = symm.empty((self.nx, nvp), dtype=torch.float64, device=self.device)   # state: all x, my columns
        f_st = symm.empty((self.nx, nvp), dtype=torch.float64, device=self.device)   # f* after the x-push, same layout
        share = symm.empty((self.nx,), dtype=torch.float64, device=self.device)      # my share of the charge density
        h_vs, h_st = symm.rendezvous(f_vs, group), symm.rendezvous(f_st, group)
        h_sh = symm.rendezvous(share, group)
        share.zero_()
        # inboxes of the field tail (adept_b200_vdfdx_field_peers_f64): every rank's share of rho lands here, double-
        # buffered by the parity of the step count; one flag per sending rank
        inbox = symm.empty((2 * self.P * self.nx,), dtype=torch.float64, device=self.device)
        flags = symm.empty((8,), dtype=torch.int64, device=self.device)
        h_in, h_fl = symm.rendezvous(inbox, group), symm.rendezvous(flags, group)
        inbox.zero_()
        flags.zero_()
        f_vs.copy_(self.state[name])
        self.state[name] = f_vs
        nparts = self.lops.ops.vdfdx_rho_parts(f_vs)
        tt = lambda a: torch.as_tensor(np.array(a, dtype=np.float64), device=self.device)  # noqa: E731
        self.p2p = {
            # per-step inputs stay on the device: space factors here, O(1) time factors from the host each step
            "ex_space": tt(np.stack([d.envelope.space_envelope(self.x) * np.ones_like(self.x) for d in self.ex]))
            if self.ex else None,
            "ex_kx": tt(np.stack([d.k0 * self.x for d in self.ex])) if self.ex else None,
            "dex": torch.zeros(self.nx, dtype=torch.float64, device=self.device),
            # ion / P: every rank adds its share, the all-reduce of the scaled partial densities is rho itself
            "ion_share": None if self.ion is None else (self.ion / self.P).contiguous(),
            "nu_fp_space": tt(self.nu_fp_prof.space_envelope(self.x[self.rows]) * np.ones(self.nxp)),
            "green": tt(np.real(np.fft.ifft(-1j * np.asarray(self.grid.one_over_kx, dtype=np.float64)))),
            "f_vs": f_vs, "f_st": f_st, "vs_ptrs": list(h_vs.buffer_ptrs), "st_ptrs": list(h_st.buffer_ptrs),
            "handles": (h_vs, h_st, h_sh), "nv": nv, "share": share, "share_ptrs": list(h_sh.buffer_ptrs),
            # ADEPT_B200_SHARDED_SYNC=nccl keeps the two NCCL all-reduces (A/B timing); default: symmetric-memory barriers
            # + a peer-memory sum in rank order (no collective library call in the step)
            "symm_sync": os.environ.get("ADEPT_B200_SHARDED_SYNC", "symm") != "nccl",
            "parts": torch.zeros((nparts, self.nx), dtype=torch.float64, device=self.device),
            "token": torch.zeros(1, dtype=torch.float64, device=self.device),
        }
        # ADEPT_B200_SHARDED_CE=c (chunks; default 0 = off): copy engines gather the rank's rows from the owning ranks into
        # an x-sharded stage chunk by chunk on a second stream while the v-row kernel works on the previous chunk.
        # Measured on 2 GPUs (r02n): 290 us per step with 4 chunks, 459 us with 8, against 218 us with the peer transfers
        # inside the kernel -- a chunk of 512 rows is less than one wave of CTAs, and each chunk costs P copy calls.
        ce = int(os.environ.get("ADEPT_B200_SHARDED_CE", "0"))
        while ce > 1 and (self.nxp % ce or (self.nxp // ce) & 1):
            ce //= 2
        self.p2p["ce_chunks"] = ce if (ce >= 1 and nv >= 2048 and self.nxp % 2 == 0) else 0
        if self.p2p["ce_chunks"]:
            self.p2p["copy_stream"] = torch.cuda.Stream(device=self.device)
            self.p2p["ev_start"] = torch.cuda.Event()
            self.p2p["ev_in"] = [torch.cuda.Event() for _ in range(ce)]
        # ADEPT_B200_SHARDED_MOVERS=n (default 0 = off; measured slower: 32 movers 203 us against 120 us for the v-row
        # kernel on 2 GPUs -- the two-slot pipelines hold too few bytes in flight): mover CTAs of the v-row kernel gather
        nm = int(os.environ.get("ADEPT_B200_SHARDED_MOVERS", "0"))
        while nm > 0 and (self.nxp % nm or nm & 1):
            nm //= 2
        if self.p2p["ce_chunks"]:
            nm = 0
            self.p2p["stage"] = torch.empty((self.nxp, nv), dtype=torch.float64, device=self.device)
            self.p2p["round_ctr"], self.p2p["n_movers"] = None, 0
        elif nm >= 2 and nv >= 2048:
            self.p2p["stage"] = torch.empty((self.nxp, nv), dtype=torch.float64, device=self.device)
            self.p2p["round_ctr"] = torch.zeros(self.nxp // nm, dtype=torch.int32, device=self.device)
            self.p2p["n_movers"] = nm
        else:
            self.p2p["stage"], self.p2p["round_ctr"], self.p2p["n_movers"] = None, None, 0
        # ADEPT_B200_SHARDED_TAIL=0 keeps the separate reduce / exchange / Poisson launches (A/B timing)
        tail_ok = (self.nx in (1024, 2048, 4096) and self.p2p["symm_sync"]
                   and os.environ.get("ADEPT_B200_SHARDED_TAIL", "1") != "0" and nvp % 4 == 0)
        g = self.cfg["grid"]
        self.p2p["tail"] = None if not tail_ok else {
            "rank": self.rank, "epoch": 0, "share_ptrs": list(h_in.buffer_ptrs), "flag_ptrs": list(h_fl.buffer_ptrs),
            "handles": (h_in, h_fl), "inbox": inbox, "flags": flags,
            "counter": torch.zeros(4, dtype=torch.int32, device=self.device),
            "ion_share": self.p2p["ion_share"], "dv": float(g["species_grids"][name]["dv"]),
            "charge": float(g["species_params"][name]["charge"]), "dx": float(g["dx"]), "green": self.p2p["green"],
            "rho": torch.zeros(self.nx, dtype=torch.float64, device=self.device),
            "e": torch.zeros(self.nx, dtype=torch.float64, device=self.device),
            "dex": torch.zeros(self.nx, dtype=torch.float64, device=self.device),
            "pond": torch.zeros(self.nx, dtype=torch.float64, device=self.device),
            "a_zero": torch.zeros(self.nx + 2, dtype=torch.float64, device=self.device),
            "ex_space": self.p2p["ex_space"], "ex_kx": self.p2p["ex_kx"],
            "ex_w": [d.w0 + d.dw0 for d in self.ex], "ex_a0": [d.a0 for d in self.ex], "ex_tenv": [], "ex_wt": [],
        }
        dist.barrier(group=self.group)  # every rank's buffers are mapped and initialised before anyone stores into them

    def _step_p2p(self):
        """One leapfrog step with the transposes fused into the v-row kernel's loads and stores (module docstring)."""
        g, dt, t = self.cfg["grid"], float(self.grid.dt), self.t
        ops, pp, n = self.lops.ops, self.p2p, self.names[0]
        sg, sp = g["species_grids"][n], g["species_params"][n]
        if pp["tail"] is not None:
            try:
                return self._step_p2p_tail()
            except AdeptB200Error:
                if pp["tail"]["epoch"] != 1:  # only the very first call may decline (shape / residency); nothing ran yet
                    raise
                pp["tail"] = None
        # driver field and collision frequency: same closed forms and rounding order as the reference (field.py:21-26,
        # functions.py:112-118), evaluated on the device from resident space factors -- no host-device copy per step
        dex = torch.empty(self.nx, dtype=torch.float64, device=self.device)
        if self.ex:  # one launch (adept_b200_ex_driver_f64), time factors from the host like the native step
            ws = [d.w0 + d.dw0 for d in self.ex]
            ops.ex_driver(pp["ex_space"], pp["ex_kx"], ws, [d.a0 for d in self.ex],
                          [float(d.envelope.time_envelope(t)) for d in self.ex], [d.phase(t) for d in self.ex], out=dex)
        else:
            dex.zero_()
        # 1. x-push on my columns, purely local (its 32-byte row pieces would waste the link)
        ops.vdfdx_rho(pp["f_vs"], self.v_loc[n], dt, self.k1x, pp["parts"], out=pp["f_st"])
        # my share of rho = ion / P + q dv (my partial velocity sums); the all-reduce finishes the charge density
        if pp["symm_sync"]:
            ops.reduce_parts(pp["parts"], float(sg["dv"]), float(sp["charge"]), base=pp["ion_share"], out=pp["share"])
            pp["handles"][2].barrier(channel=0)  # every rank's share is written (and its x-push has completed)
            rho = ops.sum_peers(pp["share_ptrs"], self.nx, torch.empty(self.nx, dtype=torch.float64, device=self.device))
        else:
            rho = ops.reduce_parts(pp["parts"], float(sg["dv"]), float(sp["charge"]), base=pp["ion_share"])
            dist.all_reduce(rho, op=dist.ReduceOp.SUM, group=self.group)  # also: every rank's x-push has completed
        e = ops.poisson_green(rho, pp["green"])  # nx/32 CTAs instead of one 4096-point FFT in a single CTA
        e_loc, dex_loc = e[self.rows].contiguous(), dex[self.rows].contiguous()
        # 2. v-push + collisions on my rows: cells come from and go back to the ranks that own their columns
        nu_fp = float(self.nu_fp_prof.time_envelope(t)) * pp["nu_fp_space"]
        ops.vpush_collide_p2p(pp["st_ptrs"], pp["vs_ptrs"], self.rank * self.nxp, self.nxp, pp["nv"], e_loc, None,
                              float(sp["charge"]), float(sp["mass"]), dt, float(sg["kvr"][1]), self.v_full[n],
                              float(sg["dv"]), nu_fp, model=self.coll.model, dex=dex_loc, scheme=self.coll.scheme,
                              stage=pp["stage"] if pp["n_movers"] else None, round_counters=pp["round_ctr"],
                              n_movers=pp["n_movers"])
        if pp["symm_sync"]:
            pp["handles"][2].barrier(channel=1)  # every rank's stores into my columns (and its reads of the shares) are done
        else:
            dist.all_reduce(pp["token"], group=self.group)  # every rank's stores into my columns have completed
        self.state["e"], self.state["de"] = e, dex
        self.step_index += 1
        self.t = self.step_index * dt
        return self.state

    def _step_p2p_tail(self):
        """The same step as two launches + one barrier: x-push with the peer exchange of the charge density and the field
        solve in its tail, fused v-row kernel over peer memory, symmetric-memory barrier."""
        g, dt, t = self.cfg["grid"], float(self.grid.dt), self.t
        ops, pp, n = self.lops.ops, self.p2p, self.names[0]
        sg, sp, tl = g["species_grids"][n], g["species_params"][n], self.p2p["tail"]
        tl["epoch"] += 1
        tl["ex_tenv"] = [float(d.envelope.time_envelope(t)) for d in self.ex]
        tl["ex_wt"] = [d.phase(t) for d in self.ex]
        if not self.ex:
            tl["dex"].zero_()
        ops.vdfdx_field_peers(pp["f_vs"], self.v_loc[n], dt, self.k1x, pp["parts"], pp["f_st"], tl)
        e, dex = tl["e"], tl["dex"]
        nu_fp = float(self.nu_fp_prof.time_envelope(t)) * pp["nu_fp_space"]
        if pp["ce_chunks"]:
            self._vpush_ce(e[self.rows], dex[self.rows], nu_fp)
        else:
            ops.vpush_collide_p2p(pp["st_ptrs"], pp["vs_ptrs"], self.rank * self.nxp, self.nxp, pp["nv"], e[self.rows],
                                  None, float(sp["charge"]), float(sp["mass"]), dt, float(sg["kvr"][1]), self.v_full[n],
                                  float(sg["dv"]), nu_fp, model=self.coll.model, dex=dex[self.rows],
                                  scheme=self.coll.scheme, stage=pp["stage"], round_counters=pp["round_ctr"],
                                  n_movers=pp["n_movers"])
        pp["handles"][2].barrier(channel=1)  # every rank's stores into my columns are done
        self.state["e"], self.state["de"] = e, dex
        self.step_index += 1
        self.t = self.step_index * dt
        return self.state

    def _vpush_ce(self, e_loc, dex_loc, nu_fp):
        """v-row kernel over my rows in chunks; the copy engines gather chunk c + 1 (one strided copy per owning rank, over
        NVLink for the others) while the SMs work on chunk c.  Results leave the kernel as TMA stores to the owners."""
        g, dt = self.cfg["grid"], float(self.grid.dt)
        ops, pp, n = self.lops.ops, self.p2p, self.names[0]
        sg, sp = g["species_grids"][n], g["species_params"][n]
        C_, nv, nvp = pp["ce_chunks"], pp["nv"], self.nvp[n]
        nc = self.nxp // C_
        main, cs = torch.cuda.current_stream(), pp["copy_stream"]
        pp["ev_start"].record(main)  # the x-push of every rank has completed (flag exchange in its tail)
        cs.wait_event(pp["ev_start"])
        stage = pp["stage"]
        with torch.cuda.stream(cs):
            for c in range(C_):
                r0 = self.rank * self.nxp + c * nc
                for dj in range(self.P):  # start with the next rank: the ranks' reads spread over the links
                    j = (self.rank + 1 + dj) % self.P
                    ops.copy2d(stage.data_ptr() + (c * nc * nv + j * nvp) * 8, nv, int(pp["st_ptrs"][j]) + r0 * nvp * 8,
                               nvp, nvp, nc)
                pp["ev_in"][c].record(cs)
        for c in range(C_):
            main.wait_event(pp["ev_in"][c])
            rows = slice(c * nc, (c + 1) * nc)
            ops.vpush_collide_p2p(pp["st_ptrs"], pp["vs_ptrs"], self.rank * self.nxp + c * nc, nc, nv, e_loc[rows], None,
                                  float(sp["charge"]), float(sp["mass"]), dt, float(sg["kvr"][1]), self.v_full[n],
                                  float(sg["dv"]), nu_fp[rows], model=self.coll.model, dex=dex_loc[rows],
                                  scheme=self.coll.scheme, stage=stage[rows], n_movers=0, nx_global=self.nx)

    # ---- layout changes -------------------------------------------------------------------------------------------
    def to_x_sharded(self, f_vs):
        """[nx, nv/P] (all x, my columns) -> [nx/P, nv] (my rows, all v): one all-to-all + local unpack."""
        P, nxp, nvp = self.P, self.nxp, f_vs.shape[1]
        recv = torch.empty((P, nxp, nvp), dtype=f_vs.dtype, device=f_vs.device)
        dist.all_to_all_single(recv, f_vs.reshape(P, nxp, nvp), group=self.group)  # row blocks are contiguous
        return recv.permute(1, 0, 2).reshape(nxp, P * nvp).contiguous()

    def to_v_sharded(self, f_xs):
        """[nx/P, nv] -> [nx, nv/P]: local pack + one all-to-all (the receive side is already in place)."""
        P, nxp = self.P, self.nxp
        nvp = f_xs.shape[1] // P
        send = f_xs.reshape(nxp, P, nvp).permute(1, 0, 2).contiguous()
        recv = torch.empty((P, nxp, nvp), dtype=f_xs.dtype, device=f_xs.device)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv.reshape(P * nxp, nvp)

    def gather_full(self, name):
        """Full f[nx, nv] on every rank (diagnostics / tests)."""
        parts = [torch.empty_like(self.state[name]) for _ in range(self.P)]
        dist.all_gather(parts, self.state[name].contiguous(), group=self.group)
        return torch.cat(parts, dim=1)

    # ---- one leapfrog step (vector_field.py:87-95 + :232-253, collectives explicit) ----------------------------------
    def _dex(self, t):
        total = np.zeros_like(self.x)
        for d in self.ex:
            w = d.w0 + d.dw0
            total += d.envelope(self.x, t) * w * d.a0 * np.sin(d.k0 * self.x - w * t)
        return torch.as_tensor(total, device=self.device)

    def step(self):
        if self.p2p is not None:
            return self._step_p2p()
        g, dt, t = self.cfg["grid"], float(self.grid.dt), self.t
        lops = self.lops
        dex = self._dex(t)
        # 1. x-push on the v-shard; partial charge density of the result
        rho = self.ion
        fstar = {}
        for n in self.names:
            sg, sp = g["species_grids"][n], g["species_params"][n]
            fstar[n], rowsum = lops.vdfdx_rowsum(self.state[n], self.v_loc[n], dt, self.k1x)
            dist.all_reduce(rowsum, op=dist.ReduceOp.SUM, group=self.group)  # nx doubles
            rho = lops.rho_from_sum(rowsum, float(sg["dv"]), float(sp["charge"]), rho)
        # 2. field solve, replicated (O(nx))
        e = lops.poisson(rho, self.one_over_kx)
        e_loc, dex_loc = e[self.rows].contiguous(), dex[self.rows].contiguous()
        # 3. transpose, v-push and collisions on my rows, transpose back
        new = {}
        for n in self.names:
            sg, sp = g["species_grids"][n], g["species_params"][n]
            rows = self.to_x_sharded(fstar[n])
            rows = lops.edfdv(self.cfg["terms"]["edfdv"], rows, e_loc, dex_loc, float(sp["charge"]), float(sp["mass"]),
                              dt, float(sg["kvr"][1]), float(sg["dv"]))
            if n == self.coll.ref_species and (self.fp_on or self.krook_on):
                xl = self.x[self.rows]
                nu_fp = nu_K = None
                if self.fp_on:
                    nu_fp = torch.as_tensor(self.nu_fp_prof(xl, t) * np.ones_like(xl), device=self.device)
                if self.krook_on:
                    nu_K = torch.as_tensor(self.nu_K_prof(xl, t) * np.ones_like(xl), device=self.device)
                c = self.coll
                rows = lops.collide(rows, self.v_full[n], float(sg["dv"]), dt, nu_fp, nu_K, self.f_mx, c.model,
                                    c.scheme, c.nodrag, c.m, c.sg_ratio)
            new[n] = self.to_v_sharded(rows)
        self.state.update(new)
        self.state["e"], self.state["de"] = e, dex
        self.step_index += 1
        self.t = self.step_index * dt
        return self.state
