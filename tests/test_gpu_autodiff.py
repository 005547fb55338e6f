"""Backward (adjoint) entry points checked by finite differences (north_star: "a custom_vjp whose backward is the adjoint
of the same operators ... checked by finite differences"), operator by operator and through several leapfrog steps
(BASELINE.json configs[4] at reduced length: gradient of the final field energy w.r.t. the drive amplitude)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ad():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200 import autodiff

    return autodiff


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device="cuda")


def setup(nx=32, nv=64, seed=0):
    rng = np.random.default_rng(seed)
    vmax, xmax = 6.4, 2 * np.pi / 0.3
    dv, dx = 2 * vmax / nv, xmax / nx
    v = np.linspace(-vmax + dv / 2, vmax - dv / 2, nv)
    x = np.linspace(dx / 2, xmax - dx / 2, nx)
    f = (1 + 0.05 * np.cos(0.3 * x))[:, None] * np.exp(-((v - 0.2) ** 2) / 2)[None, :] / np.sqrt(2 * np.pi)
    f = f * (1 + 0.01 * rng.standard_normal((nx, nv)))
    kx = 2 * np.pi * np.fft.fftfreq(nx, d=dx)
    ook = np.zeros(nx)
    ook[1:] = 1.0 / kx[1:]
    p = dict(v=dev(v), dv=dv, dt=0.1, k1x=2 * np.pi / xmax, k1v=2 * np.pi / (nv * dv), q=-1.0, m=1.0,
             one_over_kx=dev(ook), ion=dev(np.ones(nx)), fp_model=1)
    return f, x, v, p, rng


def fd_check(fn, inputs, idx, rng, h=1e-6, rtol=2e-6):
    """Directional derivative of sum(fn(*inputs) * W) along a random direction of inputs[idx]: autograd vs central FD."""
    ins = [t.clone().requires_grad_(i == idx) if isinstance(t, torch.Tensor) else t for i, t in enumerate(inputs)]
    out = fn(*ins)
    w = dev(rng.standard_normal(tuple(out.shape)))
    (out * w).sum().backward()
    g = ins[idx].grad
    d = dev(rng.standard_normal(tuple(ins[idx].shape)))
    d = d * ins[idx].detach().abs().mean()  # perturbation on the scale of the input
    ad_val = float((g * d).sum())

    def val(sign):
        pert = [t.detach() + sign * h * d if i == idx else t for i, t in enumerate(inputs)]
        return float((fn(*pert) * w).sum())

    fd_val = (val(+1) - val(-1)) / (2 * h)
    assert abs(ad_val - fd_val) <= rtol * max(abs(fd_val), abs(ad_val), 1e-12), (ad_val, fd_val)


def test_vdfdx_vjp(ad):
    f, x, v, p, rng = setup()
    fd_check(lambda ff: ad.vdfdx(ff, p["v"], 0.37, p["k1x"]), [dev(f)], 0, rng)


@pytest.mark.parametrize("nx,nv", [(8, 16), (32, 64), (4, 512), (2, 4096)])
def test_edfdv_exp_vjp_f_and_e(ad, nx, nv):
    f, x, v, p, rng = setup(nx, nv, seed=nv)
    e = dev(0.3 * rng.standard_normal(nx))
    fn = lambda ff, ee: ad.edfdv_exp(ff, ee, -1.0, 1.0, 0.1, p["k1v"])  # noqa: E731
    fd_check(fn, [dev(f), e], 0, rng)
    fd_check(fn, [dev(f), e], 1, rng)


def test_charge_density_and_poisson_vjp(ad):
    f, x, v, p, rng = setup()
    fd_check(lambda ff: ad.charge_density(ff, p["dv"], p["q"], p["ion"]), [dev(f)], 0, rng)
    fd_check(lambda r: ad.poisson(r, p["one_over_kx"]), [dev(rng.standard_normal(32))], 0, rng)


@pytest.mark.parametrize("model", [0, 1])
@pytest.mark.parametrize("nx,nv,nu0", [(8, 64, 0.5), (4, 512, 2.0), (3, 96, 0.05)])
def test_collide_vjp_f_and_nu(ad, model, nx, nv, nu0):
    f, x, v, p, rng = setup(nx, nv, seed=7)
    nu = dev(nu0 * (1 + 0.3 * rng.random(nx)))
    fn = lambda ff, nn: ad.collide_fp(ff, nn, p["v"], p["dv"], 0.1, model)  # noqa: E731
    fd_check(fn, [dev(f), nu], 0, rng, rtol=5e-6)
    fd_check(fn, [dev(f), nu], 1, rng, rtol=5e-6)


def test_gradient_of_final_field_energy_wrt_drive_amplitude(ad):
    """configs[4] in miniature: d/d(a0) of 0.5 mean(e^2) after nsteps driven leapfrog + Dougherty steps."""
    nx, nv, nsteps = 32, 64, 12
    f0, x, v, p, rng = setup(nx, nv, seed=3)
    w0, k0 = 1.1598, 0.3
    nu = dev(np.full(nx, 1e-2))

    def loss(a0, fin):
        f, e, t = fin, None, 0.0
        for _ in range(nsteps):
            dex = a0 * w0 * torch.sin(dev(k0 * x) - w0 * t)
            f, e = ad.leapfrog_step(f, dex, nu, p)
            t += p["dt"]
        return 0.5 * torch.mean(e**2.0)

    a0 = torch.tensor(1.0e-2, dtype=torch.float64, device="cuda", requires_grad=True)
    fin = dev(f0).requires_grad_(True)
    L = loss(a0, fin)
    L.backward()
    h = 1e-6
    fd = (float(loss(a0.detach() + h, fin.detach())) - float(loss(a0.detach() - h, fin.detach()))) / (2 * h)
    assert abs(float(a0.grad) - fd) <= 1e-6 * abs(fd), (float(a0.grad), fd)
    # and along a random direction of the initial distribution
    d = dev(rng.standard_normal((nx, nv))) * fin.detach().abs().mean()
    fdf = (float(loss(a0.detach(), fin.detach() + h * d)) - float(loss(a0.detach(), fin.detach() - h * d))) / (2 * h)
    adf = float((fin.grad * d).sum())
    assert abs(adf - fdf) <= 2e-6 * max(abs(fdf), 1e-14), (adf, fdf)


def test_c5_gradient_through_2000_steps(ad):
    """BASELINE.json configs[4] at full length: C2-sized grid (64 x 512), 2000 driven leapfrog + Dougherty steps,
    d/d(a0) of the final field energy 0.5 mean(e^2) by reverse mode through the CUDA adjoints vs central differences
    of the same forward run."""
    nx, nv, nsteps = 64, 512, 2000
    f0, x, v, p, rng = setup(nx, nv, seed=5)
    w0, k0 = 1.1598, 0.3
    nu = dev(np.full(nx, 1e-3))
    kx = dev(k0 * x)

    def loss(a0, fin):
        f, e, t = fin, None, 0.0
        for i in range(nsteps):
            dex = a0 * w0 * torch.sin(kx - w0 * t)
            f, e = ad.leapfrog_step(f, dex, nu, p)
            t = (i + 1) * p["dt"]
        return 0.5 * torch.mean(e**2.0)

    a0 = torch.tensor(1.0e-3, dtype=torch.float64, device="cuda", requires_grad=True)
    L = loss(a0, dev(f0))
    L.backward()
    g = float(a0.grad)
    h = 1e-7
    with torch.no_grad():
        fd = (float(loss(a0.detach() + h, dev(f0))) - float(loss(a0.detach() - h, dev(f0)))) / (2 * h)
    assert np.isfinite(g) and abs(g) > 0
    assert abs(g - fd) <= 1e-5 * abs(fd), (g, fd, float(L))
