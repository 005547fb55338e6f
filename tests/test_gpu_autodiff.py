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


def test_spline_vjp_matches_oracle_jacobian_and_fd(ad):
    """The reference's tests/test_vlasov1d/test_velocity_cubic_spline.py:49-70 restated: gradients of
    sum(interp(f, shift) * W) w.r.t. f and the shift.  Here the comparison partner is the numpy oracle differentiated
    numerically (central differences in f are exact for a function linear in f; in the shift they converge as h^2), on
    the reference's own grid, shifts and tolerances (2e-11 for f; 1e-7 for the shift, limited by the FD step)."""
    from oracle import vlasov1d as O

    nx, nv = 5, 48
    vmin, vmax = -3.0, 7.0
    dv = (vmax - vmin) / nv
    rng = np.random.default_rng(1)
    f = rng.standard_normal((nx, nv))
    W = rng.standard_normal((nx, nv))
    shift = dv * np.array([-1.37, -0.22, 0.19, 0.83, 2.41])
    q, m, dt = -1.0, 1.0, 0.1
    e = shift / dt * m / q  # accel * dt = shift
    ft, et = dev(f).requires_grad_(True), dev(e).requires_grad_(True)
    out = ad.edfdv_spline(ft, et, q, m, dt, dv)
    np.testing.assert_allclose(out.detach().cpu().numpy(), O.uniform_cubic_interp(f, shift, dv), rtol=2e-12, atol=2e-12)
    (out * dev(W)).sum().backward()
    # d/df: the operator is linear in f, so its transpose applied to W is exact through unit perturbations
    gf = np.zeros_like(f)
    for i in range(nx):
        for j in range(nv):
            d = np.zeros_like(f)
            d[i, j] = 1.0
            gf[i, j] = np.sum((O.uniform_cubic_interp(f + d, shift, dv) - O.uniform_cubic_interp(f - d, shift, dv)) * W) / 2
    np.testing.assert_allclose(ft.grad.cpu().numpy(), gf, rtol=2e-11, atol=2e-11)
    h = 1e-6 * dv
    gs = np.array([np.sum((O.uniform_cubic_interp(f, shift + h * np.eye(nx)[i], dv)
                           - O.uniform_cubic_interp(f, shift - h * np.eye(nx)[i], dv)) * W) / (2 * h) for i in range(nx)])
    ge = et.grad.cpu().numpy() / (q / m * dt)  # chain: shift = (q/m) e dt
    np.testing.assert_allclose(ge, gs, rtol=1e-6, atol=1e-6 * np.max(np.abs(gs)))


@pytest.mark.parametrize("nx,nv", [(8, 64), (3, 384), (2, 4096)])
def test_spline_vjp_fd(ad, nx, nv):
    f, x, v, p, rng = setup(nx, nv, seed=11)
    e = dev(0.5 * rng.standard_normal(nx))
    fn = lambda ff, ee: ad.edfdv_spline(ff, ee, -1.0, 1.0, 0.1, p["dv"])  # noqa: E731
    fd_check(fn, [dev(f), e], 0, rng)
    fd_check(fn, [dev(f), e], 1, rng, h=1e-7, rtol=2e-5)


@pytest.mark.parametrize("model", [0, 1])
@pytest.mark.parametrize("nx,nv,nu0", [(8, 64, 0.5), (4, 512, 2.0), (3, 96, 0.05)])
def test_collide_chang_cooper_vjp_f_and_nu(ad, model, nx, nv, nu0):
    f, x, v, p, rng = setup(nx, nv, seed=9)
    nu = dev(nu0 * (1 + 0.3 * rng.random(nx)))
    fn = lambda ff, nn: ad.collide_fp(ff, nn, p["v"], p["dv"], 0.1, model, 1)  # noqa: E731
    fd_check(fn, [dev(f), nu], 0, rng, rtol=5e-6)
    fd_check(fn, [dev(f), nu], 1, rng, rtol=5e-6)


def test_krook_vjp_f_and_nu(ad):
    f, x, v, p, rng = setup(8, 64, seed=13)
    nuK = dev(0.3 * (1 + rng.random(8)))
    vv = v
    f_mx = dev(np.exp(-vv**2 / 2) / (np.sum(np.exp(-vv**2 / 2)) * p["dv"]))
    fn = lambda ff, nn: ad.krook(ff, nn, p["v"], p["dv"], 0.1, f_mx)  # noqa: E731
    fd_check(fn, [dev(f), nuK], 0, rng)
    fd_check(fn, [dev(f), nuK], 1, rng)


def test_sixth_spline_dougherty_gradient_through_200_steps(ad):
    """The stock configs/vlasov-1d/epw.yaml composition (sixth-order splitting + cubic-spline v-advection + Dougherty
    collisions, 32 x 256): d/d(a0) of the final field energy through 200 steps, reverse mode through the CUDA adjoints
    vs central differences of the same forward run."""
    nx, nv, nsteps = 32, 256, 200
    f0, x, v, p, rng = setup(nx, nv, seed=17)
    p = dict(p, edfdv="cubic-spline", fp_model=1, fp_scheme=0)
    w0, k0 = 1.1598, 0.3
    nu = dev(np.full(nx, 1e-3))
    kx = dev(k0 * x)
    offs = ad.sixth_substep_times(p["dt"])

    def loss(a0, fin):
        f, e = fin, None
        for i in range(nsteps):
            t = i * p["dt"]
            dex = [a0 * w0 * torch.sin(kx - w0 * (t + o)) for o in offs]
            f, e = ad.sixth_step(f, dex, nu, p)
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


def test_sixth_step_forward_matches_native_step(ad):
    """The differentiable sixth-order composition is the same map as the native step (and hence the oracle)."""
    import yaml
    from pathlib import Path

    from adept_b200.module import Vlasov1D

    with open(Path(__file__).parent / "golden" / "epw.yaml") as fh:
        deck = yaml.safe_load(fh)
    deck["diagnostics"] = {"diag-vlasov-dfdt": False, "diag-fp-dfdt": False}
    deck["drivers"]["ex"] = {}
    deck["density"]["species-background"].update(basis="sine", baseline=1.0, amplitude=1.0e-2, wavenumber=0.3)
    sim = Vlasov1D(deck)
    g = sim.cfg["grid"]
    sg, sp = g["species_grids"]["electron"], g["species_params"]["electron"]
    fs = sim.vector_field.vpfp.vlasov_poisson.field_solve
    p = dict(v=dev(np.array(sg["v"])), dv=float(sg["dv"]), dt=float(sim.grid.dt), k1x=float(g["kxr"][1]),
             k1v=float(sg["kvr"][1]), q=float(sp["charge"]), m=float(sp["mass"]), one_over_kx=dev(np.array(fs.kmul)),
             ion=dev(np.array(fs.static_charge_density)), fp_model=1, fp_scheme=0, edfdv="cubic-spline")
    nu = dev(sim.vector_field.nu_fp_prof(np.asarray(sim.grid.x), 0.0) * np.ones(g["nx"]))
    f = sim.state["electron"].clone()
    zeros = [torch.zeros(g["nx"], dtype=torch.float64, device="cuda")] * 6
    f1, e1 = ad.sixth_step(f, zeros, nu, p)
    y = sim.step()
    assert float((f1 - y["electron"]).norm() / y["electron"].norm()) <= 1e-13
    assert float((e1 - y["e"]).abs().max()) <= 1e-13 * max(float(y["e"].abs().max()), 1e-3)


@pytest.mark.parametrize("variant", ["exp-dougherty", "spline-cc-krook"])
def test_native_whole_step_backward(ad, variant):
    """adept_b200_step_bwd_f64 (the _bwd rule of a custom_vjp around the whole native step): the gradient of the final
    field energy w.r.t. the drive amplitude and the initial distribution through 40 driven leapfrog steps of a
    Vlasov1D deck, reverse sweep step by step, against central differences of the same native forward run."""
    import yaml
    from copy import deepcopy
    from pathlib import Path

    from adept_b200.module import Vlasov1D

    with open(Path(__file__).parent / "golden" / "epw.yaml") as fh:
        deck = yaml.safe_load(fh)
    deck["grid"].update(nx=32, nv=128)
    deck["terms"].update(time="leapfrog", edfdv="exponential")
    deck["diagnostics"] = {"diag-vlasov-dfdt": False, "diag-fp-dfdt": False}
    deck["terms"]["fokker_planck"]["time"]["baseline"] = 1.0e-2
    if variant == "spline-cc-krook":
        deck["terms"].update(edfdv="cubic-spline")
        deck["terms"]["fokker_planck"]["type"] = "chang_cooper_dougherty"
        deck["terms"]["krook"]["is_on"] = True
        deck["terms"]["krook"]["time"]["baseline"] = 1e-2
    a0_ref = float(deck["drivers"]["ex"]["0"]["params"]["a0"])
    nsteps, t0 = 40, 30.0

    def forward(a0, f_init=None, keep=False):
        d = deepcopy(deck)
        d["drivers"]["ex"]["0"]["params"]["a0"] = a0
        sim = Vlasov1D(d)
        i0 = int(round(t0 / sim.grid.dt))
        sim.t, sim.step_index = i0 * sim.grid.dt, i0
        if f_init is not None:
            sim.state["electron"] = f_init.clone()
        states = []
        for _ in range(nsteps):
            if keep:
                states.append((sim.t, dict(sim.state)))
            sim.step()
        return sim, states, 0.5 * float(torch.mean(sim.state["e"] ** 2.0))

    sim, states, L = forward(a0_ref, keep=True)
    f_init = states[0][1]["electron"].clone()
    nat = sim.vector_field.native
    nx = sim.cfg["grid"]["nx"]
    f_bar = torch.zeros_like(sim.state["electron"])
    e_bar = sim.state["e"] / nx  # d(0.5 mean(e^2)) / de
    a0_bar = 0.0
    for t, y in reversed(states):
        y_new, bars = nat.vjp(t, y, f_bar, e_bar)
        a0_bar += float((bars["dex"] * y_new["de"]).sum()) / a0_ref  # dex is linear in a0; de is the driver field used
        f_bar, e_bar = bars["f"], None
    h = 1e-6 * a0_ref
    fd = (forward(a0_ref + h)[2] - forward(a0_ref - h)[2]) / (2 * h)
    assert abs(a0_bar - fd) <= 2e-5 * abs(fd), (a0_bar, fd)
    rng = np.random.default_rng(3)
    d = dev(rng.standard_normal(tuple(f_init.shape))) * f_init.abs().mean()
    hf = 1e-6
    fdf = (forward(a0_ref, f_init + hf * d)[2] - forward(a0_ref, f_init - hf * d)[2]) / (2 * hf)
    adf = float((f_bar * d).sum())
    assert abs(adf - fdf) <= 2e-5 * max(abs(fdf), 1e-16), (adf, fdf)
