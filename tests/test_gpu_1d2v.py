"""GPU parity of ``solver: vlasov-1d2v`` (adept_b200/vlasov1d2v.py) against its numpy oracle (oracle/vlasov1d2v.py), and
the reference's own identities on the GPU path (tests/test_vlasov1d2v/test_1d_limit.py): the marginal of a
v_perp-separable run equals the vlasov-1d run, the cumulative diagnostics telescope to F(t) - F(0)."""

from copy import deepcopy

import numpy as np
import pytest
import torch

from oracle import vlasov1d as O
from oracle import vlasov1d2v as O2
from test_oracle_1d2v import base_config

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200 import ops as _ops

    return _ops


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device="cuda")


def test_marginal_and_transpose_kernels(ops):
    rng = np.random.default_rng(0)
    f = rng.standard_normal((6, 40, 10))
    w = rng.random(10)
    np.testing.assert_allclose(ops.marginal(dev(f), dev(w)).cpu().numpy(), np.einsum("xvp,p->xv", f, w), rtol=0,
                               atol=1e-14)
    for shape in ((6, 40, 10), (3, 33, 65), (1, 128, 32)):
        a = rng.standard_normal(shape)
        assert np.array_equal(ops.transpose_last2(dev(a)).cpu().numpy(), np.swapaxes(a, -1, -2))


@pytest.mark.parametrize("fp_type,time,nx,nv,nvperp", [("dougherty", "leapfrog", 16, 64, 8),
                                                       ("dougherty", "sixth", 32, 256, 16),
                                                       ("dougherty_nodrag", "sixth", 16, 128, 4),
                                                       ("lenard_bernstein", "leapfrog", 8, 512, 6),
                                                       ("dougherty", "leapfrog", 256, 64, 8)])
def test_2v_step_by_step_parity(fp_type, time, nx, nv, nvperp):
    """Every state entry of 8 steps against the oracle (both sides start each step from the oracle's state), <= 1e-12."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200.vlasov1d2v import Vlasov1D2V

    deck = base_config(nx=nx, nv=nv, nvperp=nvperp, fp_type=fp_type, time=time, nu=1e-2)
    sim = Vlasov1D2V(deepcopy(deck))
    cfg = O2.build_cfg(deck)
    vf = O2.VlasovMaxwell2V(cfg)
    y = O2.init_state(cfg)
    dt = cfg["grid"]["dt"]
    np.testing.assert_allclose(sim.state["electron"].cpu().numpy(), y["electron"], rtol=0, atol=1e-15)
    worst = {}
    for n in range(8):
        t = 2.0 + n * dt  # inside the driver's flat top
        state = {k: dev(v) for k, v in y.items()}
        y_gpu = sim.vector_field(t, state, None)
        y = vf(t, y, None)
        assert set(y_gpu) == set(y)
        for k in y:
            g = y_gpu[k].cpu().numpy()
            if k == "e":
                err = np.max(np.abs(g - y[k])) / max(np.max(np.abs(y[k])), 1e-3)
            elif k.startswith("diag-"):
                err = np.max(np.abs(g - y[k])) / np.max(np.abs(y["electron"]))
            else:
                err = rel_l2(g, y[k])
            worst[k] = max(worst.get(k, 0.0), err)
    for k, err in worst.items():
        assert err <= 1e-12, (k, err)
    assert np.max(np.abs(y["e"])) > 0


@pytest.mark.parametrize("fp_type", ["dougherty", "dougherty_nodrag"])
def test_gpu_marginal_matches_the_1d_gpu_run(fp_type):
    """tests/test_vlasov1d2v/test_1d_limit.py:21-46 on the GPU path (reduced size): 60 driven sixth-order steps; the 2V
    marginal reproduces the vlasov-1d run to 1e-9 of its maximum, the field to 1e-8."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200.module import Vlasov1D
    from adept_b200.vlasov1d2v import Vlasov1D2V

    deck = base_config(nx=32, nv=256, nvperp=16, fp_type=fp_type)
    s2, s1 = Vlasov1D2V(deepcopy(deck)), Vlasov1D(deepcopy(deck))
    w = dev(s2.cfg["grid"]["species_grids"]["electron"]["wperp"])
    emax = 0.0
    from adept_b200 import ops as _ops

    for _ in range(60):
        s2.step(), s1.step()
        emax = max(emax, float(s1.state["e"].abs().max()))
        assert float((s2.state["e"] - s1.state["e"]).abs().max()) <= 1e-8 * max(emax, 1e-12)
    assert emax > 1e-4, "driver did not couple"
    F = _ops.marginal(s2.state["electron"], w)
    f1 = s1.state["electron"]
    assert float((F - f1).abs().max() / f1.abs().max()) < 1e-9


def test_gpu_cumulative_diags_telescope():
    """tests/test_vlasov1d2v/test_1d_limit.py:49-85 on the GPU path."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200 import ops as _ops
    from adept_b200.vlasov1d2v import Vlasov1D2V

    sim = Vlasov1D2V(base_config(nx=16, nv=128, nvperp=8, nu=1e-2))
    g = sim.cfg["grid"]["species_grids"]["electron"]
    w = dev(g["wperp"])
    F0 = _ops.marginal(sim.state["electron"], w).clone()
    sim.run(40)
    F1 = _ops.marginal(sim.state["electron"], w)
    acc = sim.state["diag-vlasov-cumulative"] + sim.state["diag-fp-cumulative"]
    scale = float((F1 - F0).abs().max())
    assert scale > 0
    assert float((acc - (F1 - F0)).abs().max()) / scale < 1e-10
    assert float((sim.state["diag-fp-cumulative"].sum(dim=-1) * float(g["dv"])).abs().max()) < 1e-10


def test_cylindrical_landau_is_refused():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200.vlasov1d2v import Vlasov1D2V

    deck = base_config()
    deck["terms"]["fokker_planck"]["type"] = "cylindrical_landau"
    with pytest.raises(NotImplementedError, match="cylindrical_landau"):
        Vlasov1D2V(deck)
