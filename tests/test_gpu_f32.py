"""Single-precision entry points (adept_b200_*_f32) against the fp64 numpy oracle on the same inputs.

Bar (BASELINE.json north_star): per-application relative L2 <= 1e-5 in fp32.  The reference never runs in fp32
(adept/_base_.py:287-292 switches x64 on), so these are the explicit extra SURVEY.md 8b names; the comparison partner is
the fp64 oracle evaluated on the float32-rounded input.
"""

import numpy as np
import pytest
import torch

from oracle import vlasov1d as O
from test_gpu_ops import _fp_cfg, dev, host, make_f, rel_l2

pytestmark = pytest.mark.gpu
RTOL32 = 1e-5


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200 import ops as _ops

    return _ops


def dev32(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device="cuda")


@pytest.mark.parametrize("nx,nv", [(16, 64), (64, 512), (256, 64), (1024, 64), (4096, 64), (8192, 8)])
@pytest.mark.parametrize("noise", [0.0, 0.05])
def test_vdfdx_f32(ops, nx, nv, noise):
    f, x, v, dx, dv = make_f(nx, nv, seed=nx + nv, noise=noise)
    f = f.astype(np.float32).astype(np.float64)
    kxr = np.fft.rfftfreq(nx, d=dx) * 2 * np.pi
    ref = O.space_exponential(f, kxr, v, 0.1)
    out = host(ops.vdfdx_f32(dev32(f), dev(v), 0.1, kxr[1])).astype(np.float64)
    assert rel_l2(out, ref) <= RTOL32


@pytest.mark.parametrize("nx,nv", [(8, 16), (64, 512), (32, 1024), (16, 4096), (4, 8192)])
@pytest.mark.parametrize("noise", [0.0, 0.05])
def test_edfdv_exp_f32(ops, nx, nv, noise):
    f, x, v, dx, dv = make_f(nx, nv, seed=nx * nv, noise=noise)
    f = f.astype(np.float32).astype(np.float64)
    rng = np.random.default_rng(nv)
    e, dex, pond = 0.3 * rng.standard_normal(nx), 0.05 * rng.standard_normal(nx), 0.01 * rng.standard_normal(nx)
    kvr = np.fft.rfftfreq(nv, d=dv) * 2 * np.pi
    ref = O.velocity_exponential(f, kvr, e + dex, pond, 0.1, -1.0, 1.0)
    out = host(ops.edfdv_exp_f32(dev32(f), dev(e), dev(pond), -1.0, 1.0, 0.1, kvr[1], dex=dev(dex))).astype(np.float64)
    assert rel_l2(out, ref) <= RTOL32
    # in place
    fd = dev32(f)
    ops.edfdv_exp_f32(fd, dev(e), dev(pond), -1.0, 1.0, 0.1, kvr[1], dex=dev(dex), out=fd)
    assert rel_l2(host(fd).astype(np.float64), ref) <= RTOL32


@pytest.mark.parametrize("fp_type,model,scheme", [("lenard_bernstein", 0, 0), ("dougherty", 1, 0),
                                                  ("chang_cooper", 0, 1), ("chang_cooper_dougherty", 1, 1)])
@pytest.mark.parametrize("nx,nv", [(8, 64), (16, 512), (6, 4096)])
def test_collide_f32(ops, fp_type, model, scheme, nx, nv):
    coll = O.Collisions(_fp_cfg(nv, 6.4, fp_type, krook=True))
    f, x, v, dx, dv = make_f(nx, nv, seed=nv, noise=0.0)
    f = (f * (1 + 0.05 * np.sin(7 * v))[None, :]).astype(np.float32).astype(np.float64)
    # production regime (weak collisions) and a moderately collisional one; fp32 carries cond(I - dt nu L) * 6e-8
    for nu_max, tol in ((1e-4, RTOL32), (1e-2 * (dv / 0.025) ** 2, RTOL32)):
        nu = np.linspace(0.2, 1.0, nx) * nu_max
        nuK = np.linspace(0.5, 0.1, nx) * 1e-2
        ref = coll(nu, nuK, f, 0.1)
        nout = torch.empty(nx, dtype=torch.float32, device="cuda")
        out = host(ops.collide_f32(dev32(f), dev(v), dv, 0.1, nu_fp=dev(nu), nu_K=dev(nuK), f_mx=dev(coll.f_mx),
                                   model=model, scheme=scheme, n_out=nout)).astype(np.float64)
        assert rel_l2(out, ref) <= tol, (nu_max, rel_l2(out, ref))
        np.testing.assert_allclose(host(nout), np.sum(ref, axis=1) * dv, rtol=2e-5)


def test_f32_leapfrog_step_composition(ops):
    """A leapfrog + Dougherty step composed from the _f32 operators (field solve in fp64 on the nx-long density) stays
    within 1e-5 of the fp64 oracle step for 20 steps of the C2 deck (errors do not accumulate beyond the bar)."""
    import yaml
    from pathlib import Path

    with open(Path(__file__).parent / "golden" / "epw.yaml") as fh:
        deck = yaml.safe_load(fh)
    deck["grid"].update(nx=64, nv=512)
    deck["terms"].update(time="leapfrog", edfdv="exponential")
    deck["terms"]["fokker_planck"]["time"]["baseline"] = 1.0e-3
    deck["diagnostics"] = {"diag-vlasov-dfdt": False, "diag-fp-dfdt": False}
    cfg = O.build_cfg(deck)
    vf = O.VlasovMaxwell(cfg)
    y = O.init_state(cfg)
    g = cfg["grid"]
    sg = g["species_grids"]["electron"]
    v, dv, dt = dev(np.array(sg["v"])), float(sg["dv"]), float(g["dt"])
    k1x, k1v = float(g["kxr"][1]), float(sg["kvr"][1])
    ook = dev(np.array(g["one_over_kx"]))
    ion = dev(np.array(g["ion_charge"]))
    f = dev32(y["electron"])
    t0 = 30.0
    i0 = int(round(t0 / dt))
    for n in range(20):
        t = (i0 + n) * dt
        y = vf(t, y, None)
        fs = ops.vdfdx_f32(f, v, dt, k1x)
        rho = ion - fs.double().sum(dim=1) * dv
        e = ops.poisson(rho.contiguous(), ook)
        dex = dev(y["de"])
        f2 = ops.edfdv_exp_f32(fs, e, None, -1.0, 1.0, dt, k1v, dex=dex)
        nu = dev(vf.nu_fp_prof(np.asarray(g["x"]), t) * np.ones(g["nx"])) if hasattr(vf, "nu_fp_prof") else None
        f = ops.collide_f32(f2, v, dv, dt, nu_fp=nu, model=1, scheme=0)
    assert rel_l2(host(f).astype(np.float64), y["electron"]) <= RTOL32
    assert np.max(np.abs(host(e) - y["e"])) <= 1e-4 * max(np.max(np.abs(y["e"])), 1e-3)
