#!/usr/bin/env python
"""bench.py -- phase-space cell-updates/s of the full fp64 vlasov-1d step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--nx 4096] [--nv 4096]

Workload (BASELINE.json configs[2], "C3"): one species, nx = nv = 4096, fp64, leapfrog + spectral x/v pushes +
Poisson + Dougherty Fokker-Planck collisions, driven electron plasma wave (the configs/vlasov-1d/epw.yaml driver),
synthetic Maxwellian + 1e-2 cos(k0 x) perturbation.  With N > 1 every rank advances its own independent grid of that
size (ensemble members shard with no communication: weak scaling); value = total cells advanced / max-over-ranks time.

One JSON line is printed by rank 0 (see the task contract): value (device-resident state), e2e (per-step host inputs
copied from pinned memory + per-step diagnostic read-back), roofline of the dominant kernel, cpu_baseline (the numpy
oracle timed on the host cores, rank 0 / N=1 only).  ``--impl reference`` times the oracle alone.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "phase-space cell-updates/s per fp64 vlasov-1d step"
UNIT = "cell-updates/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def c3_deck(nx: int, nv: int, steps_hint: int = 10000) -> dict:
    """BASELINE.json configs[2] built from the stock epw.yaml physics (SURVEY.md 8d synthetic inputs)."""
    env = lambda base: {"baseline": base, "bump_or_trough": "bump", "center": 0.0, "rise": 25.0, "slope": 0.0,  # noqa: E731
                        "bump_height": 0.0, "width": 100000.0}
    return {
        "units": {"normalizing_temperature": "2000eV", "normalizing_density": "1.5e21/cc"},
        "density": {"quasineutrality": True,
                    "species-background": {"noise_seed": 420, "noise_type": "gaussian", "noise_val": 0.0, "v0": 0.0,
                                           "T0": 1.0, "m": 2.0, "basis": "sine", "baseline": 1.0, "amplitude": 1.0e-2,
                                           "wavenumber": 0.3}},
        "grid": {"dt": 0.1, "nv": nv, "nx": nx, "tmin": 0.0, "tmax": 0.1 * steps_hint, "vmax": 6.4,
                 "xmax": 2 * np.pi / 0.3, "xmin": 0.0},
        "save": {}, "solver": "vlasov-1d", "mlflow": {"experiment": "bench", "run": "c3"},
        "drivers": {"ex": {"0": {"params": {"a0": 1.0e-2, "k0": 0.3, "w0": 1.1598, "dw0": 0.0},
                                 "envelope": {"time": {"center": 40.0, "rise": 5.0, "width": 30.0},
                                              "space": {"center": 0.0, "rise": 10.0, "width": 4000000.0}}}},
                    "ey": {}},
        "diagnostics": {"diag-vlasov-dfdt": False, "diag-fp-dfdt": False},
        "terms": {"field": "poisson", "edfdv": "exponential", "time": "leapfrog",
                  "fokker_planck": {"is_on": True, "type": "Dougherty", "time": env(1.0e-5), "space": env(1.0)},
                  "krook": {"is_on": False, "time": env(1.0), "space": env(1.0)}},
    }


def workload_name(nx, nv):
    return (f"C3 vlasov-1d {nx}x{nv} fp64: leapfrog + spectral vdfdx/edfdv + Poisson + Dougherty FP, driven EPW "
            "(BASELINE.json configs[2])")


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML (nvidia-ml-py) every 5 ms on a host thread while the timed
    region runs (nvidia-smi -lms block-buffers its pipe, so short runs would see no samples)."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index, self.sm, self.mask, self.smax = index, [], 0, None
        self._stop = threading.Event()
        self._thread = None
        self.err = None

    def _sample(self):
        nv, h = self._nv, self._h
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
        try:
            self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:  # older bindings
            self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))

    def _run(self):
        try:
            while not self._stop.is_set():
                self._sample()
                self._stop.wait(0.005)
        except Exception as exc:  # never fail the bench because of the sampler
            self.err = repr(exc)

    def start(self):
        """NVML is initialised here (slow, ~0.1 s), outside the timed region; sampling then runs on a host thread."""
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv, self._h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:
            self.err = repr(exc)
            return
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2.0)
        if self.err or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": [f"nvml unavailable: {self.err}"], "samples": 0}
        reasons = [n for n, bit in self.REASONS if self.mask & bit]
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.smax, "reasons": reasons,
                "samples": len(self.sm)}


# ------------------------------------------------------------------------------------------------ reference arm (CPU oracle)
def time_oracle(nx, nv, steps, warmup, workers):
    from oracle import vlasov1d as O

    O.FFT_WORKERS = workers
    try:  # torchrun exports OMP_NUM_THREADS=1: give BLAS/OpenMP pools under numpy/scipy the host cores back
        import threadpoolctl

        threadpoolctl.threadpool_limits(limits=workers)
    except Exception:
        pass
    cfg = O.build_cfg(c3_deck(nx, nv))
    vf = O.VlasovMaxwell(cfg)
    y = O.init_state(cfg)
    dt = cfg["grid"]["dt"]
    t = 30.0  # driver on
    for _ in range(warmup):
        y = vf(t, y, None)
        t += dt
    t0 = time.perf_counter()
    for _ in range(steps):
        y = vf(t, y, None)
        t += dt
    el = time.perf_counter() - t0
    return nx * nv * steps / el, el / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    # bounded sample: the full C3 grid for a few steps (each oracle step is seconds of CPU work)
    steps = max(1, min(args.steps, 3))
    warmup = 1
    val, sec = time_oracle(args.nx, args.nv, steps, warmup, workers)
    sample = (f"{steps} full steps of the {args.nx}x{args.nv} workload after {warmup} warm-up; numpy/scipy restatement "
              f"of the reference (jax is not installable offline), scipy.fft workers={workers}, LAPACK dgtsv row loop")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.nx, args.nv), "nx": args.nx, "nv": args.nv},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ library-FFT stand-in
def time_cufft_stand_in(sim, nx, nv, steps=10):
    """SURVEY.md 8d "Reference GPU path": jax[cuda12] is not installable offline, so the stand-in for XLA's lowering of
    the reference step is the same operators composed from torch.fft (cuFFT, fp64) on this GPU: x-advection
    (vlasov.py:236-238), charge density + spectral Poisson (field.py:197-224), spectral v-advection (vlasov.py:83-90).
    There is no batched tridiagonal solver in torch, so the collision step is LEFT OUT: this times two of the three
    operator applications of the step and is a lower bound on the library time.  A reported baseline only; nothing in
    the library or in the parity tests uses torch.fft."""
    import torch

    cfg = sim.cfg
    g = cfg["grid"]
    name = next(iter(g["species_grids"]))
    sg, sp = g["species_grids"][name], g["species_params"][name]
    dev = sim.state[name].device
    f = sim.state[name].clone()
    t64 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev)  # noqa: E731
    v, kxr, kvr, ook = t64(sg["v"]), t64(g["kxr"]), t64(sg["kvr"]), t64(g["one_over_kx"])
    ion = t64(g["ion_charge"])
    dt, dv, q, m = float(g["dt"]), float(sg["dv"]), float(sp["charge"]), float(sp["mass"])
    xphase = torch.exp(-1j * kxr[:, None] * v[None, :] * dt)  # static table, kept resident like XLA would constant-fold

    def step(f):
        f = torch.fft.irfft(torch.fft.rfft(f, dim=0) * xphase, n=nx, dim=0)
        rho = ion + q * (f.sum(dim=1) * dv)
        e = torch.real(torch.fft.ifft(-1j * ook * torch.fft.fft(rho)))
        accel = (q * e) / m
        return torch.fft.irfft(torch.fft.rfft(f, dim=1) * torch.exp(-1j * kvr[None, :] * dt * accel[:, None]), n=nv, dim=1)

    for _ in range(3):
        f = step(f)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        f = step(f)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / steps
    assert bool(torch.isfinite(f).all())
    del f, xphase
    torch.cuda.empty_cache()
    return {"value": nx * nv / sec, "unit": UNIT, "ms_per_step": sec * 1e3, "kind": "torch.fft (cuFFT) fp64 composition",
            "what": "x-advection + charge density + spectral Poisson + spectral v-advection on this GPU, collision "
                    "step left out (no batched tridiagonal solver in torch): 2 of the 3 operator applications, a lower "
                    "bound on the library time; stand-in for jax[cuda12] (not installable offline)",
            "steps": steps}


# ------------------------------------------------------------------------------------------------ B200 arm
# algorithmic bytes per cell and launch: one fp64 read + one fp64 write of f per operator application;
# the fused v-push + collision kernel performs two operator applications per launch (it moves 16 B/cell)
FULL_PASS = {"vdfdx": 16.0, "vdfdx_tma": 16.0, "vdfdx_tma_field": 16.0, "vdfdx_dual": 16.0, "vdfdx_dual_field": 16.0, "edfdv_exp": 16.0, "edfdv_spline": 16.0, "collide": 16.0,
             "vpush_collide": 32.0}


def profile_report(lib):
    import ctypes as C

    buf = C.create_string_buffer(1 << 16)
    lib.adept_b200_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, count, ms = line.split()
        out[name] = (int(count), float(ms))
    return out



# ------------------------------------------------------------------------------------------------ extras (N > 1)
def c4_decks(n_members: int, nx: int = 64, nv: int = 512):
    """BASELINE.json configs[3]: members scan k0 in linspace(0.2, 0.4) x a0 in logspace(-4, -1); every
    member has its own box length 2 pi / k0 (so its own kx) and a driver matched to the Bohm-Gross frequency."""
    nk = max(d for d in range(1, int(np.sqrt(n_members)) + 1) if n_members % d == 0)  # 1024 -> 32 x 32, 128 -> 8 x 16
    na = n_members // nk
    decks = []
    for k0 in np.linspace(0.2, 0.4, nk):
        for a0 in np.logspace(-4, -1, na):
            d = c3_deck(nx, nv)
            d["grid"]["xmax"] = float(2 * np.pi / k0)
            d["density"]["species-background"]["wavenumber"] = float(k0)
            d["drivers"]["ex"]["0"]["params"].update(k0=float(k0), a0=float(a0), w0=float(np.sqrt(1 + 3 * k0**2)))
            decks.append(d)
    return decks


def oracle_steps(deck, nsteps, t_start=30.0):
    """The numpy oracle advanced `nsteps` from t_start (checker only)."""
    import yaml

    from oracle import vlasov1d as O

    cfg = O.build_cfg(yaml.safe_load(yaml.safe_dump(deck)))
    vf = O.VlasovMaxwell(cfg)
    y = O.init_state(cfg)
    dt = cfg["grid"]["dt"]
    i0 = int(round(t_start / dt))
    t = t_start
    for i in range(nsteps):
        y = vf(t, y, None)
        t = (i0 + i + 1) * dt
    return y


def run_sharded_extra(world, rank, nx, nv, K, cells_n1_ms, max_over_ranks, barrier):
    """BASELINE.json configs[2], second half: the SAME single nx x nv grid sharded along v over the N ranks (the
    reference's grid.parallel: ["x", "v"]; pushers/vlasov.py:95-101,245-248, fokker_planck.py:435-441): strong scaling.
    Timed like the main figure (K steps between barriers, CUDA events, max over ranks); parity: 3 steps of a 1024 x
    2048 sharded grid gathered on rank 0 and compared with the oracle."""
    import ctypes

    import torch
    import torch.distributed as dist

    from adept_b200 import _lib
    from adept_b200.sharded import ShardedVlasov1D

    out = {"what": f"ONE {nx}x{nv} grid sharded along v over {world} GPUs (strong scaling), transposes fused into the "
                   "v-row kernel over NVLink peer memory"}
    # ---- parity first (small grid, 3 steps, against the oracle on rank 0) ----------------------------------------
    pnx, pnv, psteps = 1024, 2048, 3
    pdeck = c3_deck(pnx, pnv)
    sim = ShardedVlasov1D(c3_deck(pnx, pnv))
    sim.t, sim.step_index = 30.0, 300
    for _ in range(psteps):
        sim.step()
    name = sim.names[0]
    full = sim.gather_full(name).cpu().numpy()
    e_gpu = sim.state["e"].cpu().numpy()
    out["parity_mode"] = "p2p" if sim.p2p is not None else "nccl"
    del sim
    if rank == 0:
        y = oracle_steps(pdeck, psteps)
        out["parity_rel_l2"] = float(np.linalg.norm(full - y[name]) / np.linalg.norm(y[name]))
        out["parity_e_max_abs"] = float(np.max(np.abs(e_gpu - y["e"])))
        out["parity_what"] = f"{psteps} steps of a {pnx}x{pnv} grid sharded over {world} ranks vs the numpy oracle"
    barrier()
    # ---- timing on the full grid ----------------------------------------------------------------------------------
    sim = ShardedVlasov1D(c3_deck(nx, nv))
    sim.t, sim.step_index = 30.0, 300
    out["transpose"] = "p2p" if sim.p2p is not None else "nccl"
    for _ in range(30):  # NCCL and the symmetric-memory rendezvous keep initialising lazily for tens of steps
        sim.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        sim.step()
    e1.record()
    barrier()
    el = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    lib = _lib.load()
    lib.adept_b200_profile(1)
    for _ in range(K):
        sim.step()
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.adept_b200_profile_report(buf, len(buf))
    lib.adept_b200_profile(0)
    kern = {}
    for ln in buf.value.decode().splitlines():
        nm, cnt, ms = ln.split()
        kern[nm] = round(float(ms) / int(cnt) * 1e3, 1)
    barrier()
    del sim
    out.update({"ms_per_step": el / K * 1e3, "value": nx * nv * K / el, "unit": UNIT, "steps": K,
                "strong_scaling_vs_n1": (cells_n1_ms / (el / K * 1e3)) if cells_n1_ms else None,
                "n1_ms_per_step": cells_n1_ms, "rank0_kernel_us": kern})
    return out


def run_c4_extra(world, rank, K, max_over_ranks, barrier, n_members=1024):
    """BASELINE.json configs[3]: 1024 independent 64 x 512 runs scanning k lambda_D and drive amplitude, members sharded
    over the ranks with no data-path collective.  Parity: members 0 and -1 of rank 0's slice against the
    oracle after 3 steps."""
    import torch

    from adept_b200.ensemble import EnsembleVlasov1D, member_slice

    nx, nv = 64, 512
    decks = c4_decks(n_members, nx, nv)
    sl = member_slice(n_members, rank, world)
    mine = decks[sl]
    ens = EnsembleVlasov1D(mine)
    ens.t, ens.step_index = 30.0, 300
    out = {"what": f"{n_members} independent {nx}x{nv} leapfrog + Dougherty runs (k0 x a0 scan), "
                   f"{len(mine)} members per GPU, no collective", "members": n_members, "members_per_gpu": len(mine)}
    for _ in range(3):
        ens.step()
    if rank == 0:
        errs = []
        for i in (0, len(mine) - 1):
            y = oracle_steps(mine[i], 3)
            got = ens.member_state(i)["electron"].cpu().numpy()
            errs.append(float(np.linalg.norm(got - y["electron"]) / np.linalg.norm(y["electron"])))
        out["parity_rel_l2"] = max(errs)
        out["parity_what"] = "members 0 and last of rank 0 after 3 steps vs the numpy oracle"
    for _ in range(5):
        ens.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        ens.step()
    e1.record()
    barrier()
    el = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    cells = n_members * nx * nv
    out.update({"us_per_step": el / K * 1e6, "value": cells * K / el, "unit": UNIT, "steps": K,
                "frac_of_48B_roofline_per_gpu": None})
    del ens
    return out, cells * K / el



def _time_steps(step, nsteps, warm=10):
    import torch

    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nsteps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / nsteps


def run_single_gpu_extras(nx, nv, peak):
    """N = 1 only: per-step times of the other BASELINE configs (parity-test cases, reported beside the contract line):
    C1 = the stock configs/vlasov-1d/epw.yaml (32 x 256, sixth + cubic-spline + Dougherty + both dfdt diagnostics),
    C2 = 64 x 512 leapfrog + Dougherty, C3 with `time: sixth` (192 B/cell), C4's per-GPU share (128
    members of 64 x 512)."""
    import torch
    import yaml

    from adept_b200.ensemble import EnsembleVlasov1D
    from adept_b200.module import Vlasov1D

    out = {}
    with open(ROOT / "tests" / "golden" / "epw.yaml") as fh:
        c1 = yaml.safe_load(fh)
    for key, deck, n in (("c1_epw_32x256_sixth_spline", c1, 300), ("c2_64x512_leapfrog_fp", c3_deck(64, 512), 300)):
        sim = Vlasov1D(deck)
        sim.t, sim.step_index = 30.0, 300
        sec = _time_steps(sim.step, n)
        g = sim.cfg["grid"]
        entry = {"us_per_step": sec * 1e6, "cells": int(g["nx"]) * int(g["nv"]), "how": "Vlasov1D.step(), host-issued"}
        try:
            gsec = sim.graph_steps_per_second(n)  # CUDA-graph replay of the native step (device-resident time tables)
            entry["us_per_step_cuda_graph"] = 1e6 / gsec
        except Exception as exc:
            entry["us_per_step_cuda_graph"] = None
            entry["graph_unavailable"] = f"{type(exc).__name__}: {exc}"[:200]
        out[key] = entry
        del sim
    d6 = c3_deck(nx, nv)
    d6["terms"]["time"] = "sixth"
    sim = Vlasov1D(d6)
    sim.t, sim.step_index = 30.0, 300
    sec = _time_steps(sim.step, 10, warm=3)
    out["c3_sixth"] = {"ms_per_step": sec * 1e3, "value": nx * nv / sec, "unit": UNIT,
                       "frac_of_192B_roofline": 192.0 * nx * nv / sec / 1e9 / peak,
                       "what": f"{nx}x{nv} sixth-order Hamiltonian splitting (5 x-pushes, 6 v-pushes, 6 field solves) + "
                               "Dougherty: 12 operator applications = 192 B/cell"}
    del sim
    torch.cuda.empty_cache()
    ens = EnsembleVlasov1D(c4_decks(128))
    ens.t, ens.step_index = 30.0, 300
    sec = _time_steps(ens.step, 100)
    cells = 128 * 64 * 512
    out["c4_share_128_members"] = {"us_per_step": sec * 1e6, "value": cells / sec, "unit": UNIT,
                                   "frac_of_48B_roofline": 48.0 * cells / sec / 1e9 / peak}
    del ens
    torch.cuda.empty_cache()
    # the one deck the reference publishes a number for (configs/vlasov-1d/iaw-turbulence-big-bench.yaml: nx = 17280,
    # nv = 2048, sixth-order, cubic-spline, Boltzmann electrons, stochastic forcing; "~45 ms/step ... on 4x A100-40GB",
    # configs/vlasov-1d/run-iaw-big.sbatch:5-6), on this one GPU
    sys.path.insert(0, str(ROOT / "tools"))
    from bigbench_probe import deck as big_deck

    for tkind, nst in (("sixth", 5), ("leapfrog", 10)):
        sim = Vlasov1D(big_deck(tkind))
        sec = _time_steps(sim.step, nst, warm=2)
        finite = bool(torch.isfinite(sim.state["ion"]).all())
        out[f"iaw_big_bench_17280x2048_{tkind}"] = {
            "ms_per_step": sec * 1e3, "value": 17280 * 2048 / sec, "unit": UNIT, "finite": finite,
            "reference_published_ms_per_step": 45.0 if tkind == "sixth" else None,
            "reference_hardware": "4 x A100-40GB (run-iaw-big.sbatch:5-6)" if tkind == "sixth" else None}
        del sim
        torch.cuda.empty_cache()
    # single-precision operators (explicit extra; 8 B/cell per operator application)
    from adept_b200 import ops as _ops

    f32 = torch.rand((nx, nv), dtype=torch.float32, device="cuda") + 0.5
    g32 = torch.empty_like(f32)
    vv = torch.linspace(-6.4, 6.4, nv, dtype=torch.float64, device="cuda")
    ee = torch.full((nx,), 1e-2, dtype=torch.float64, device="cuda")
    nu = torch.full((nx,), 1e-5, dtype=torch.float64, device="cuda")
    f32_ops = {}
    for nm, fn in (("vdfdx_f32", lambda: _ops.vdfdx_f32(f32, vv, 0.1, 0.3, out=g32)),
                   ("edfdv_exp_f32", lambda: _ops.edfdv_exp_f32(f32, ee, None, -1.0, 1.0, 0.1, 0.49, out=g32)),
                   ("collide_f32", lambda: _ops.collide_f32(f32, vv, 12.8 / nv, 0.1, nu_fp=nu, out=g32))):
        sec = _time_steps(fn, 10, warm=3)
        f32_ops[nm] = {"us": sec * 1e6, "frac_of_8B_roofline": 8.0 * nx * nv / sec / 1e9 / peak}
    out["f32_operators_4096x4096"] = f32_ops
    return out


def bind_to_gpu_numa_node(torch, local):
    """Pin this rank to the CPU cores NVML reports as local to its GPU, before any pinned host buffer is allocated: the
    pages of the e2e arm's host buffers then sit on the GPU's own NUMA node.  With 8 ranks on one host the unbound run
    lost half of its end-to-end rate to cross-socket copies (round 1: e2e scaling 0.52 at N = 8).  Returns the number
    of cores bound to, or None when NVML / the affinity call is not available (nothing changes then)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:  # noqa: BLE001 -- a missing NVML or a restricted cpuset must not stop the bench
        pass
    return None


def run_b200(args):
    import torch
    import torch.distributed as dist

    from adept_b200 import _lib, ops
    from adept_b200._lib import AdeptB200Error
    from adept_b200.module import Vlasov1D

    if not torch.cuda.is_available():
        raise AdeptB200Error("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa_cores = bind_to_gpu_numa_node(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    import ctypes

    C_sizeof_step = ctypes.sizeof(_lib.Step)  # kernel-argument bytes the host sends per step

    nx, nv, K, W = args.nx, args.nv, args.steps, args.warmup
    cells = nx * nv
    sim = Vlasov1D(c3_deck(nx, nv))
    dt = sim.grid.dt
    t_start = 30.0  # inside the driver's flat top: every term of the step is active
    sim.t, sim.step_index = t_start, int(round(t_start / dt))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        tns = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    # ---- value: state resident in HBM, K steps back to back ------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        sim.step()
    barrier()
    sampler.sm.clear()  # keep only samples taken during the timed region
    sampler.mask = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        sim.step()
    ev1.record()
    barrier()
    elapsed = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    value = world * cells * K / elapsed

    # ---- the same K steps again with every kernel launch bracketed by CUDA events on the launching stream ----------
    # (the event records cost ~2.5 us of stream time per launch, ~9 % of this step, so they are kept out of `value`;
    # kernel durations themselves are unaffected)
    lib.adept_b200_profile(1)
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    pv0.record()
    for _ in range(K):
        sim.step()
    pv1.record()
    barrier()
    elapsed_prof = pv0.elapsed_time(pv1) * 1e-3
    prof = profile_report(lib)
    lib.adept_b200_profile(0)
    launches = sum(c for c, _ in prof.values())

    per_kernel = {name: {"launches_per_step": c / K, "avg_us": ms / c * 1e3, "share_of_step": ms * 1e-3 / elapsed_prof}
                  for name, (c, ms) in prof.items()}
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback 6.65 TB/s (B200_PROFILING.md)"
    full_pass = {k: v for k, v in per_kernel.items() if k in FULL_PASS}
    dom = max(full_pass, key=lambda k: full_pass[k]["share_of_step"])
    alg_bytes = FULL_PASS[dom] * cells
    achieved = alg_bytes / (full_pass[dom]["avg_us"] * 1e-6) / 1e9
    traffic_path = ROOT / "profiles" / "dram_traffic.json"
    traffic = None
    traffic_src = None
    if traffic_path.exists():
        tj = json.loads(traffic_path.read_text())
        traffic, traffic_src = tj.get(dom), f"ncu run {tj.get('_ncu_run')} (profiles/dram_traffic.json), not measured in this run"
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "operator_applications_per_launch": FULL_PASS[dom] / 16.0,
                "kernel_timing": "CUDA events around every launch in a second pass of the same K steps",
                "ms_per_step_with_events": elapsed_prof / K * 1e3,
                "step_frac_of_48B_roofline": (48.0 * cells / (elapsed / K) / 1e9) / peak,
                # the same figure for every full-pass kernel of the step (the two big kernels are within 1 % of each
                # other, so which one is "dominant" can flip from run to run)
                "per_kernel_frac": {k: FULL_PASS[k] * cells / (v["avg_us"] * 1e-6) / 1e9 / peak
                                    for k, v in full_pass.items()}}

    # ---- e2e: a K-step run through the public API starting and ending in HOST memory -------------------------------
    # timed region: H2D of the initial distribution from pinned memory, K x (step + D2H of the two field-energy
    # scalars the reference's default save logs every step, storage.py:316-317, into a pinned host ring), D2H of the
    # final distribution.  The per-step read-backs are asynchronous copies on the compute stream, as the reference's
    # own loop keeps its saves on the device until the solve returns (adept/_base_.py:413-429); the host waits once,
    # at the end of the run.
    name = next(iter(sim.cfg["grid"]["species_grids"]))
    f_host = torch.empty((nx, nv), dtype=torch.float64).pin_memory()
    f_host.copy_(sim.state[name])
    f_back = torch.empty((nx, nv), dtype=torch.float64).pin_memory()
    n_ring = max(K, W, 3)
    diag_host = torch.zeros((n_ring, 2), dtype=torch.float64).pin_memory()

    mom_host = torch.zeros((n_ring, 6), dtype=torch.float64).pin_memory()
    sgrid = sim.cfg["grid"]["species_grids"][name]
    v_dev = torch.as_tensor(np.array(sgrid["v"], dtype=np.float64), device="cuda")
    mom_dev = torch.empty((6, nx), dtype=torch.float64, device="cuda")

    def e2e_run(nsteps, with_save):
        sim.state[name] = f_host.to("cuda", non_blocking=True)
        for i in range(nsteps):
            st = sim.step()
            # the kernel writes the two scalars straight into the pinned host ring (mapped memory): the per-step
            # device-to-host transfer without a separate copy operation on the stream
            ops.field_energy(st["e"], st["de"], out=diag_host[i])
            if with_save:
                # the reference's always-on default save (storage.py:282, 306-323): mean_x sum_v {f, f v, f v^2, f v^3,
                # -|f| log|f|, f^2} dv of the new state, one more pass over f every step, read back per step
                ops.save_moments(st[name], v_dev, float(sgrid["dv"]), out=mom_dev)
                ops.row_means(mom_dev, out=mom_host[i])  # mean over x, written straight into the pinned host ring
        f_back.copy_(sim.state[name], non_blocking=True)
        torch.cuda.synchronize()
        return float(diag_host[nsteps - 1, 0])

    e2e_res = {}
    for with_save in (False, True):
        e2e_run(max(W, 3), with_save)
        barrier()
        t0 = time.perf_counter()
        last_e2 = e2e_run(K, with_save)
        e2e_res[with_save] = max_over_ranks(time.perf_counter() - t0)
        assert np.isfinite(last_e2) and np.all(np.isfinite(diag_host[:K].numpy()))
    assert np.all(np.isfinite(mom_host[:K].numpy())) and abs(float(mom_host[K - 1, 0]) - 1.0) < 1e-6  # mean density
    e2e_elapsed = e2e_res[True]  # headline: with the default save, as the reference always runs
    e2e_value = world * cells * K / e2e_elapsed
    h2d_bytes = f_host.numel() * 8 / K + C_sizeof_step  # initial state amortised over the run + the step descriptor
    d2h_bytes = 16 + 48 + f_back.numel() * 8 / K

    # ---- N > 1 only: the two BASELINE configs that need several GPUs, measured in the same job -------------------
    extra = {}
    if world > 1 and not args.no_extras:
        del f_host, f_back
        sim.state = None
        torch.cuda.empty_cache()
        try:
            extra["sharded"] = run_sharded_extra(world, rank, nx, nv, K, elapsed / K * 1e3, max_over_ranks, barrier)
        except Exception as exc:  # an extra must never take the contract line down
            extra["sharded"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
        try:
            c4, c4_value = run_c4_extra(world, rank, K, max_over_ranks, barrier)
            c4["frac_of_48B_roofline_per_gpu"] = 48.0 * c4_value / world / 1e9 / peak
            extra["ensemble_c4"] = c4
        except Exception as exc:
            extra["ensemble_c4"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    if world == 1 and not args.no_extras:
        try:
            extra.update(run_single_gpu_extras(nx, nv, peak))
        except Exception as exc:
            extra["single_gpu_extras"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
    gpu_library_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            gpu_library_baseline = time_cufft_stand_in(sim, nx, nv)
        except Exception as exc:  # a reported baseline must never take the bench line down (e.g. cuFFT plan memory)
            gpu_library_baseline = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        workers = os.cpu_count() or 1
        cval, csec = time_oracle(nx, nv, 2, 1, workers)
        cpu_baseline = {"value": cval, "unit": UNIT, "cores": workers, "kind": "port",
                        "sample": f"2 full {nx}x{nv} steps after 1 warm-up ({csec:.2f} s/step); numpy/scipy oracle, "
                                  f"scipy.fft workers={workers}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": elapsed / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(nx, nv), "nx": nx, "nv": nv, "members_per_gpu": 1,
                   "parallelism": f"ensemble x{world} (independent grids, no collective)",
                   "l2": f"working set {2 * cells * 8 / 2**20:.0f} MiB of f (in + out) per operator > 126 MB L2",
                   "t_start": t_start},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": e2e_elapsed / K * 1e3, "host_cores_bound_to_gpu_numa_node": numa_cores,
                "without_default_save": {"value": world * cells * K / e2e_res[False],
                                         "ms_per_step": e2e_res[False] / K * 1e3},
                "what": f"Vlasov1D.step() public API, {K}-step run from and to pinned HOST memory: H2D of f0 and D2H "
                        "of the final f inside the timed region (amortised per step), per-step D2H of mean_e2/mean_de2 "
                        "(written by the field-energy kernel straight into a pinned, mapped host ring) and of the six "
                        "default-save moments of the new state (one more pass over f per step, storage.py:306-323), "
                        "one host wait at the end of the run; without_default_save = the same without the moment pass"},
        "roofline": roofline, "kernels": per_kernel, "gpu_launches": launches, "clocks": clocks,
        "cpu_baseline": cpu_baseline, "gpu_library_baseline": gpu_library_baseline, "extra": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--nv", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sharded-grid / C4 / small-deck extras")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
