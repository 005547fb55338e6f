"""N > 1 host logic of the v-sharded single grid (adept_b200/sharded.py) on CPU: world_size-2 gloo processes run the
sharded leapfrog step with the numpy oracle injected as the local operator table, and the gathered result must equal the
oracle's own single-process step.  This covers the partitioning, both all-to-all transposes and the moment all-reduce;
the CUDA kernels themselves are covered by the -m gpu tests."""

import os
import socket
from copy import deepcopy
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import yaml

from oracle import vlasov1d as O

GOLD = Path(__file__).parent / "golden"


class OracleOps:
    """CPU stand-in for sharded.CudaOps (test only): same methods, numpy oracle arithmetic."""

    def __init__(self, cfg):
        self.coll = O.Collisions(cfg)

    @staticmethod
    def _t(a):
        return torch.as_tensor(np.ascontiguousarray(a))

    def vdfdx_rowsum(self, f, v, dt, k1x):
        nx = f.shape[0]
        out = O.space_exponential(f.numpy(), k1x * np.arange(nx // 2 + 1), v.numpy(), dt)
        return self._t(out), self._t(out.sum(axis=1))

    def rho_from_sum(self, total, dv, q, base):
        term = q * (total.numpy() * dv)
        return self._t(term if base is None else base.numpy() + term)

    def poisson(self, rho, one_over_kx):
        return self._t(O.poisson(rho.numpy(), one_over_kx.numpy()))

    def edfdv(self, kind, f, e, dex, q, m, dt, k1v, dv):
        nv = f.shape[1]
        ee = e.numpy() + dex.numpy()
        if kind == "exponential":
            return self._t(O.velocity_exponential(f.numpy(), k1v * np.arange(nv // 2 + 1), ee, np.zeros_like(ee), dt, q, m))
        return self._t(O.velocity_cubic_spline(f.numpy(), dv, ee, np.zeros_like(ee), dt, q, m))

    def collide(self, f, v, dv, dt, nu_fp, nu_K, f_mx, model, scheme, nodrag, sg_m, sg_ratio):
        return self._t(self.coll._apply(None if nu_fp is None else nu_fp.numpy(), None if nu_K is None else nu_K.numpy(),
                                        f.numpy(), dt))


def deck(edfdv="exponential", krook=False):
    with open(GOLD / "epw.yaml") as fh:
        d = yaml.safe_load(fh)
    d["grid"].update(nx=32, nv=64)
    d["terms"].update(time="leapfrog", edfdv=edfdv)
    d["terms"]["fokker_planck"]["time"]["baseline"] = 1.0e-2
    if krook:
        d["terms"]["krook"]["is_on"] = True
        d["terms"]["krook"]["time"]["baseline"] = 1.0e-2
    d["diagnostics"] = {"diag-vlasov-dfdt": False, "diag-fp-dfdt": False}
    return d


def _worker(rank, world, port, dk, nsteps, out_path):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from adept_b200.sharded import ShardedVlasov1D

        cfg = O.build_cfg(deepcopy(dk))
        sim = ShardedVlasov1D(dk, local_ops=OracleOps(cfg), device=torch.device("cpu"))
        sim.t, sim.step_index = 30.0, 300  # driver on
        for _ in range(nsteps):
            sim.step()
        full = sim.gather_full("electron").numpy()
        if rank == 0:
            np.savez(out_path, f=full, e=sim.state["e"].numpy(), de=sim.state["de"].numpy())
        # layout round trip: to_x_sharded followed by to_v_sharded is the identity
        back = sim.to_v_sharded(sim.to_x_sharded(sim.state["electron"]))
        assert torch.equal(back, sim.state["electron"])
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("edfdv,krook", [("exponential", False), ("cubic-spline", True)])
def test_sharded_step_equals_single_process(tmp_path, edfdv, krook):
    dk, nsteps = deck(edfdv, krook), 3
    out = tmp_path / "sharded.npz"
    mp.spawn(_worker, args=(2, _free_port(), dk, nsteps, str(out)), nprocs=2, join=True)
    got = np.load(out)
    cfg = O.build_cfg(deepcopy(dk))
    vf = O.VlasovMaxwell(cfg)
    y = O.init_state(cfg)
    t = 30.0
    for i in range(nsteps):
        y = vf(t, y, None)
        t = (300 + i + 1) * cfg["grid"]["dt"]
    rel = np.linalg.norm(got["f"] - y["electron"]) / np.linalg.norm(y["electron"])
    assert rel <= 1e-13, rel
    # E is the integral of rho = 1 - n_e (an O(1) cancellation): the all-reduce order moves it by a few 1e-16 absolute
    np.testing.assert_allclose(got["e"], y["e"], rtol=0, atol=5e-15)
    np.testing.assert_allclose(got["de"], y["de"], rtol=1e-13, atol=1e-18)


def test_sharded_rejects_indivisible_grid(tmp_path):
    dk = deck()
    dk["grid"]["nx"] = 34  # not divisible by 4; also not a power of two, but divisibility is checked first

    def run():
        mp.spawn(_bad_worker, args=(4, _free_port(), dk), nprocs=4, join=True)

    run()


def _bad_worker(rank, world, port, dk):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from adept_b200._lib import AdeptB200Error
        from adept_b200.sharded import ShardedVlasov1D

        with pytest.raises(AdeptB200Error, match="not divisible"):
            ShardedVlasov1D(dk, local_ops=object(), device=torch.device("cpu"))
    finally:
        dist.destroy_process_group()


def _reject_worker(rank, world, port, dk, match):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from adept_b200.sharded import ShardedVlasov1D

        with pytest.raises(NotImplementedError, match=match):
            ShardedVlasov1D(dk, local_ops=object(), device=torch.device("cpu"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("feature", ["hou_li_filter", "ex_stochastic"])
def test_sharded_rejects_features_it_would_drop(feature):
    """Decks whose Hou-Li filter or stochastic Ex driver the sharded step does not apply must be refused, not run
    with the feature silently missing (the single-GPU Vlasov1D and the reference do apply them)."""
    dk = deck()
    if feature == "hou_li_filter":
        dk["terms"]["hou_li_filter"] = {"is_on": True, "alpha": 36.0, "order": 36}
    else:
        dk["drivers"]["ex_stochastic"] = {"n_modes": 2, "k_min": 0.3, "k_max": 0.6, "amplitude": 1e-3, "tau": 5.0,
                                          "seed": 1}
    mp.spawn(_reject_worker, args=(2, _free_port(), dk, feature), nprocs=2, join=True)


# ------------------------------------------------------------------------------ ensembles sharded by members (C4)
def test_member_slice_covers_every_member_once():
    from adept_b200.ensemble import member_slice

    for n in (0, 1, 7, 8, 121, 1024):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                s = member_slice(n, r, world)
                seen += list(range(s.start, s.stop))
            assert seen == list(range(n))
            sizes = [member_slice(n, r, world).stop - member_slice(n, r, world).start for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        member_slice(4, 2, 2)


class _FakeEnsemble:
    """Stands in for EnsembleVlasov1D on a CPU-only box: a 'member' is a number, a step adds its own value."""

    def __init__(self, decks):
        self.x = torch.tensor([float(d) for d in decks], dtype=torch.float64)
        self.acc = torch.zeros_like(self.x)

    def run(self, nsteps):
        self.acc = self.acc + nsteps * self.x


def _ensemble_worker(rank, world, port, n_members, q):
    import torch.distributed as dist

    from adept_b200.ensemble import run_sharded_ensemble

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        decks = list(range(1, n_members + 1))
        out = run_sharded_ensemble(decks, 3, lambda e: torch.stack([e.acc, e.x], dim=1), make=_FakeEnsemble)
        q.put((rank, out.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_members", [5, 1])
def test_sharded_ensemble_gathers_in_deck_order_gloo(n_members):
    """world_size 2 over gloo: uneven split (3 + 2 members) and a rank with an empty slice (1 member)."""
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 7 * n_members
    procs = [ctx.Process(target=_ensemble_worker, args=(r, world, port, n_members, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([3.0 * np.arange(1, n_members + 1), 1.0 * np.arange(1, n_members + 1)], axis=1)
    for r in range(world):
        np.testing.assert_array_equal(got[r], want)
