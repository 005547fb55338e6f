"""BASELINE.json configs[3] in miniature: an ensemble of independent vlasov-1d runs that scan the wavenumber (box
length), the driver frequency and the drive amplitude, advanced as one batched problem.  Every member must equal the
single run of its own deck (same kernels, same arithmetic) and the oracle."""

from copy import deepcopy

import numpy as np
import pytest
import torch

from oracle import vlasov1d as O
from test_gpu_step import c2_deck, rel_l2

pytestmark = pytest.mark.gpu


def member_deck(k0, a0, w0, time="leapfrog", edfdv="exponential", nx=64, nv=512):
    d = c2_deck()
    d["grid"].update(nx=nx, nv=nv, xmax=2 * np.pi / k0)
    d["density"]["species-background"]["wavenumber"] = k0
    d["drivers"]["ex"]["0"]["params"].update(k0=k0, a0=a0, w0=w0)
    d["terms"].update(time=time, edfdv=edfdv)
    return d


SCAN = [(0.26, 1.0e-3, 1.12), (0.30, 1.0e-2, 1.1598), (0.34, 3.0e-2, 1.21), (0.38, 1.0e-4, 1.27)]


@pytest.mark.parametrize("time,edfdv", [("leapfrog", "exponential"), ("sixth", "cubic-spline")])
def test_ensemble_members_match_single_runs_and_oracle(time, edfdv):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200.ensemble import EnsembleVlasov1D
    from adept_b200.module import Vlasov1D

    decks = [member_deck(k0, a0, w0, time, edfdv) for k0, a0, w0 in SCAN]
    ens = EnsembleVlasov1D(deepcopy(decks))
    dt = ens.dt
    t0, nsteps = 30.0, 6
    ens.t, ens.step_index = t0, int(round(t0 / dt))
    ens.run(nsteps)
    for i, dk in enumerate(decks):
        sim = Vlasov1D(deepcopy(dk))
        sim.t, sim.step_index = t0, int(round(t0 / dt))
        for _ in range(nsteps):
            sim.step()
        got = ens.member_state(i)
        for key in ("electron", "e", "de"):
            a, b = got[key].cpu().numpy(), sim.state[key].cpu().numpy()
            assert rel_l2(a, b) <= 1e-13, (i, key, rel_l2(a, b))
    # member 2 against the oracle, stepping freely from the same start
    cfg = O.build_cfg(deepcopy(decks[2]))
    vf = O.VlasovMaxwell(cfg)
    y = O.init_state(cfg)
    for n in range(nsteps):
        y = vf(t0 + n * dt, y, None)
    got = ens.member_state(2)
    assert rel_l2(got["electron"].cpu().numpy(), y["electron"]) <= 1e-12
    assert np.max(np.abs(got["e"].cpu().numpy() - y["e"])) <= 1e-12 * np.max(np.abs(y["e"])) + 1e-13


def test_ensemble_rejects_members_that_differ_in_step_scalars():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200._lib import AdeptB200Error
    from adept_b200.ensemble import EnsembleVlasov1D

    a, b = member_deck(0.3, 1e-2, 1.16), member_deck(0.3, 1e-2, 1.16)
    b["grid"]["dt"] = 0.05
    with pytest.raises(AdeptB200Error):
        EnsembleVlasov1D([a, b])
