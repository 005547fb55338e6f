"""GPU parity tests: every C-ABI operator of libadept_b200.so against the numpy oracle on the same inputs.

Tolerance (BASELINE.json north_star): per-application relative L2 <= 1e-12 in fp64; the cubic stencil's
integer-cell shifts are bit-exact (reference test_velocity_cubic_spline.py:35-46).
"""

import numpy as np
import pytest
import torch

from oracle import vlasov1d as O

pytestmark = pytest.mark.gpu

RTOL = 1e-12


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


def host(t):
    return t.detach().cpu().numpy()


def make_f(nx, nv, vmax=6.4, seed=0, noise=0.0, xmax=20.94):
    rng = np.random.default_rng(seed)
    dv = 2 * vmax / nv
    v = np.linspace(-vmax + dv / 2, vmax - dv / 2, nv)
    dx = xmax / nx
    x = np.linspace(dx / 2, xmax - dx / 2, nx)
    k0 = 2 * np.pi / xmax
    f = (1 + 0.3 * np.cos(k0 * x) + 0.1 * np.sin(3 * k0 * x))[:, None] * np.exp(-((v - 0.3) ** 2) / 2)[None, :]
    f = f / (np.sum(np.exp(-(v**2) / 2)) * dv)
    if noise:
        f = f + noise * rng.standard_normal((nx, nv))
    return f, x, v, dx, dv


# ---------------------------------------------------------------------------------------------------- x-advection
@pytest.mark.parametrize("nx,nv", [(2, 8), (4, 16), (8, 1024), (16, 64), (32, 256), (64, 512), (128, 32), (256, 64),
                                   (512, 32), (1024, 64), (2048, 16), (4096, 64), (8192, 8)])
@pytest.mark.parametrize("noise", [0.0, 0.05])
def test_vdfdx_matches_oracle(ops, nx, nv, noise):
    f, x, v, dx, dv = make_f(nx, nv, seed=nx + nv, noise=noise)
    kxr = np.fft.rfftfreq(nx, d=dx) * 2 * np.pi
    dt = 0.1
    ref = O.space_exponential(f, kxr, v, dt)
    out = host(ops.vdfdx(dev(f), dev(v), dt, kxr[1]))
    assert rel_l2(out, ref) <= RTOL


def test_vdfdx_inplace_negative_dt_and_batch(ops):
    nx, nv, B = 64, 128, 3
    fs, k1s, refs = [], [], []
    for b in range(B):
        f, x, v, dx, dv = make_f(nx, nv, seed=b, noise=0.01, xmax=20.0 + 3 * b)
        kxr = np.fft.rfftfreq(nx, d=dx) * 2 * np.pi
        fs.append(f)
        k1s.append(kxr[1])
        refs.append(O.space_exponential(f, kxr, v, -0.37))
    fd = dev(np.stack(fs))
    ops.vdfdx(fd, dev(v), -0.37, 0.0, out=fd, k1x_batch=dev(np.array(k1s)))
    assert rel_l2(host(fd), np.stack(refs)) <= RTOL


@pytest.mark.parametrize("B,nx,nv", [(1, 256, 64), (3, 512, 36), (1, 4096, 128), (2, 1024, 8), (2, 64, 32), (1, 32, 6)])
def test_vdfdx_rho_fused_moment(ops, B, nx, nv):
    """x-advection fused with the charge-density velocity sum (TMA path for nx >= 256 and nv % 4 == 0, else fallback)."""
    fs, k1s, refs = [], [], []
    for b in range(B):
        f, x, v, dx, dv = make_f(nx, nv, seed=7 * b + nv, noise=0.02, xmax=20.0 + 2 * b)
        kxr = np.fft.rfftfreq(nx, d=dx) * 2 * np.pi
        fs.append(f)
        k1s.append(kxr[1])
        refs.append(O.space_exponential(f, kxr, v, 0.23))
    fd = dev(np.stack(fs))
    nparts = ops.vdfdx_rho_parts(fd)
    parts = torch.full((nparts + 1, B * nx), 7.0, dtype=torch.float64, device="cuda")  # stale contents must not leak
    out = ops.vdfdx_rho(fd, dev(v), 0.23, 0.0, parts, k1x_batch=dev(np.array(k1s)))
    ref = np.stack(refs)
    assert rel_l2(host(out), ref) <= RTOL
    ion = np.linspace(0.9, 1.1, B * nx)
    rho = host(ops.reduce_parts(parts, dv, -1.0, base=dev(ion)))
    rho_ref = ion + -1.0 * (np.sum(ref, axis=-1).ravel() * dv)
    np.testing.assert_allclose(rho, rho_ref, rtol=0, atol=1e-13 * np.max(np.abs(rho_ref)))
    # in place, no moment: the plain entry point takes the same TMA path
    ops.vdfdx(fd, dev(v), 0.23, 0.0, out=fd, k1x_batch=dev(np.array(k1s)))
    assert rel_l2(host(fd), ref) <= RTOL


@pytest.mark.parametrize("nx,nv", [(32, 16), (64, 6), (512, 64), (4096, 32)])
def test_hou_li_filter_matches_oracle(ops, nx, nv):
    """HouLiFilter (vlasov.py:209-220) through the x-advection kernels (direct and TMA path)."""
    from adept_b200.pushers import HouLiFilter

    f, x, v, dx, dv = make_f(nx, nv, seed=nx, noise=0.05)
    ref = O.hou_li_filter(f, nx, 36.0, 4)
    out = host(HouLiFilter(nx, 36.0, 4)({"electron": dev(f)})["electron"])
    assert rel_l2(out, ref) <= RTOL


def test_vdfdx_exact_characteristic_shift(ops):
    """reference test_multispecies_pushers.py:16-64 (sinusoid, nv=2 instead of 1: nv must be even)."""
    Lx = 2 * np.pi
    for nx in [16, 32]:
        dx = Lx / nx
        x = np.linspace(0, Lx - dx, nx)
        v = np.array([0.5, -0.25])
        f = np.sin(2 * x)[:, None] * np.ones((1, 2))
        out = host(ops.vdfdx(dev(f), dev(v), 0.01, 2 * np.pi / (nx * dx)))
        exact = np.sin(2 * x[:, None] - 2 * v[None, :] * 0.01)
        assert np.sqrt(np.mean((out - exact) ** 2)) < 1e-12


# ---------------------------------------------------------------------------------------------------- v-advection
@pytest.mark.parametrize("nx,nv", [(2, 2), (4, 8), (32, 16), (32, 256), (64, 512), (8, 1024), (16, 2048), (8, 4096),
                                   (4, 8192), (6, 64)])
@pytest.mark.parametrize("noise", [0.0, 0.05])
def test_edfdv_exp_matches_oracle(ops, nx, nv, noise):
    f, x, v, dx, dv = make_f(nx, nv, seed=nx * 3 + nv, noise=noise)
    rng = np.random.default_rng(nx)
    e, dex, pond = 0.3 * rng.standard_normal(nx), 0.01 * rng.standard_normal(nx), 0.02 * rng.standard_normal(nx)
    kvr = np.fft.rfftfreq(nv, d=dv) * 2 * np.pi
    q, m, dt = -1.0, 1.0, 0.1
    ref = O.velocity_exponential(f, kvr, e + dex, pond, dt, q, m)
    out = host(ops.edfdv_exp(dev(f), dev(e), dev(pond), q, m, dt, kvr[1], dex=dev(dex)))
    assert rel_l2(out, ref) <= RTOL
    # ion-like species, no pond / dex
    q, m = 10.0, 18360.0
    ref = O.velocity_exponential(f, kvr, e, np.zeros(nx), dt, q, m)
    out = host(ops.edfdv_exp(dev(f), dev(e), None, q, m, dt, kvr[1]))
    assert rel_l2(out, ref) <= RTOL


def test_edfdv_exp_exact_characteristic_shift(ops):
    """reference test_multispecies_pushers.py:67-140."""
    vmax = 2 * np.pi
    for nv in [16, 32]:
        dv = 2.0 * vmax / nv
        v = np.linspace(-vmax + dv / 2, vmax - dv / 2, nv)
        k = np.pi / vmax
        f = np.sin(k * v)[None, :] * np.ones((2, 1))
        e = np.array([0.5, 0.5])
        for q, m in [(-1.0, 1.0), (1.0, 1836.0)]:
            out = host(ops.edfdv_exp(dev(f), dev(e), None, q, m, 0.01, 2 * np.pi / (nv * dv)))
            exact = np.sin(k * (v - (q / m) * 0.5 * 0.01))[None, :]
            assert np.sqrt(np.mean((out - exact) ** 2)) < 1e-12


@pytest.mark.parametrize("nx,nv,vmin,vmax", [(8, 64, -6.4, 6.4), (8, 64, -4.0, 8.0), (32, 256, -6.4, 6.4),
                                             (64, 4096, -6.4, 6.4), (5, 40, -0.5, 1.0)])
def test_edfdv_spline_matches_oracle(ops, nx, nv, vmin, vmax):
    rng = np.random.default_rng(nv)
    dv = (vmax - vmin) / nv
    f = rng.standard_normal((nx, nv))
    base = np.array([-70.0, -2.17, -0.37, 0.0, 0.25, 1.13, 3.4, 70.0])
    shift_cells = np.resize(base, nx) + 0.01 * rng.standard_normal(nx) * (np.resize(base, nx) != 0)
    q, m, dt = -1.0, 1.0, 0.17
    e = shift_cells * dv / dt * m / q
    pond = np.zeros(nx)
    ref = O.velocity_cubic_spline(f, dv, e, pond, dt, q, m)
    out = host(ops.edfdv_spline(dev(f), dev(e), None, q, m, dt, dv))
    np.testing.assert_allclose(out, ref, rtol=2e-12, atol=2e-12)
    # with pond + dex and a different species
    q, m = 1.0, 4.0
    e2, dex, pond = rng.standard_normal(nx), 0.1 * rng.standard_normal(nx), 0.05 * rng.standard_normal(nx)
    ref = O.velocity_cubic_spline(f, dv, e2 + dex, pond, dt, q, m)
    out = host(ops.edfdv_spline(dev(f), dev(e2), dev(pond), q, m, dt, dv, dex=dev(dex)))
    np.testing.assert_allclose(out, ref, rtol=2e-12, atol=2e-12)


def test_edfdv_spline_integer_shifts_bit_exact(ops):
    """reference test_velocity_cubic_spline.py:35-46."""
    nx, nv, dv = 3, 16, 0.25
    f = np.arange(nx * nv, dtype=np.float64).reshape(nx, nv)
    shift = dv * np.array([1.0, -1.0, 0.0])
    out = host(ops.edfdv_spline(dev(f), dev(shift), None, 1.0, 1.0, 1.0, dv))  # accel*dt = e
    expected = np.empty((nx, nv))
    expected[0] = np.concatenate(([1.0e-30], f[0, :-1]))
    expected[1] = np.concatenate((f[1, 1:], [1.0e-30]))
    expected[2] = f[2]
    np.testing.assert_array_equal(out, expected)


# ---------------------------------------------------------------------------------------------------- moments / fields
@pytest.mark.parametrize("nx,nv", [(32, 256), (7, 33), (64, 4096)])
def test_moments_match_numpy(ops, nx, nv):
    f, x, v, dx, dv = make_f(nx, nv, seed=5, noise=0.01)
    ion = np.linspace(0.9, 1.1, nx)
    fd, vd = dev(f), dev(v)
    rho, j, p2 = (torch.empty(nx, dtype=torch.float64, device="cuda") for _ in range(3))
    ops.moments(fd, vd, dv, (rho, j, p2), bases=(dev(ion), None, None), scale_b=(-1.0, -1.0, 1.0))
    np.testing.assert_allclose(host(rho), -1.0 * (np.sum(f, 1) * dv) + ion, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(host(j), -1.0 * (np.sum(f * v, 1) * dv), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(host(p2), np.sum(f * v * v, 1) * dv, rtol=1e-13)


@pytest.mark.parametrize("nx", [2, 8, 32, 64, 256, 2048, 4096, 8192])
def test_poisson_matches_oracle(ops, nx):
    rng = np.random.default_rng(nx)
    dx = 20.94 / nx
    kx = np.fft.fftfreq(nx, d=dx) * 2 * np.pi
    ook = np.zeros(nx)
    ook[1:] = 1.0 / kx[1:]
    rho = 1e-2 * rng.standard_normal(nx)
    e = host(ops.poisson(dev(rho), dev(ook)))
    assert rel_l2(e, O.poisson(rho, ook)) <= RTOL
    rho_i = 1.0 + 1e-2 * rng.standard_normal(nx)
    for lam in [None, 0.0, 0.7]:
        e = host(ops.poisson(dev(rho_i), dev(kx), mode=1, Te=2.0, lambda_De=-1.0 if lam is None else lam))
        assert rel_l2(e, O.boltzmann_poisson(rho_i, kx, 2.0, lam)) <= RTOL


def test_poisson_batched(ops):
    nx, B = 64, 5
    rng = np.random.default_rng(0)
    rho = rng.standard_normal((B, nx))
    ooks = []
    for b in range(B):
        kx = np.fft.fftfreq(nx, d=(20.0 + b) / nx) * 2 * np.pi
        ook = np.zeros(nx)
        ook[1:] = 1.0 / kx[1:]
        ooks.append(ook)
    ooks = np.stack(ooks)
    e = host(ops.poisson(dev(rho), dev(ooks)))
    ref = np.stack([O.poisson(rho[b], ooks[b]) for b in range(B)])
    assert rel_l2(e, ref) <= RTOL
    e = host(ops.poisson(dev(rho), dev(ooks[0])))
    ref = np.stack([O.poisson(rho[b], ooks[0]) for b in range(B)])
    assert rel_l2(e, ref) <= RTOL


def test_ponderomotive_axpy_wave(ops):
    nx = 300
    rng = np.random.default_rng(1)
    a, aold, djy = (1e-2 * rng.standard_normal(nx + 2) for _ in range(3))
    dx, dt, c = 0.3, 0.02, 11.3
    np.testing.assert_allclose(host(ops.ponderomotive(dev(a), dx)), O.ponderomotive(a, dx), rtol=1e-14, atol=1e-300)
    n0, n1 = -1.0 + 0.01 * rng.standard_normal(nx), -1.0 + 0.01 * rng.standard_normal(nx)
    ref = O.wave_solver(a, aold, djy, -0.5 * (n0 + n1), c, dx, dt)["a"]
    out = host(ops.wave_step(dev(a), dev(aold), dev(djy), dev(n0), dev(n1), c, dx, dt))
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-16)
    ref = O.wave_solver(a, aold, djy, 0.0, c, dx, dt)["a"]
    out = host(ops.wave_step(dev(a), dev(aold), dev(djy), None, None, c, dx, dt))
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-16)
    e, j = rng.standard_normal(nx), rng.standard_normal(nx)
    np.testing.assert_array_equal(host(ops.axpy(dev(e), dev(j), -dt)), e - dt * j)


# ---------------------------------------------------------------------------------------------------- collisions
def _fp_cfg(nv, vmax, fp_type, krook=False, T0=1.0, m=2.0, sc_steps=0):
    dv = 2.0 * vmax / nv
    v = np.linspace(-vmax + dv / 2.0, vmax - dv / 2.0, nv)
    return {
        "grid": {
            "species_grids": {"electron": {"v": v, "dv": dv, "nv": nv, "vmax": vmax}},
            "species_params": {"electron": {"charge": -1.0, "mass": 1.0, "charge_to_mass": -1.0, "T0": T0}},
        },
        "terms": {"fokker_planck": {"is_on": True, "type": fp_type, "m": m,
                                    "self_consistent_beta": {"enabled": sc_steps > 0, "max_steps": sc_steps}},
                  "krook": {"is_on": krook}},
    }


MODEL = {"lb": 0, "dougherty": 1, "sg": 2}
SCHEME = {"central": 0, "cc": 1}


def _gpu_collide(ops, coll, f, nu_fp, nu_K, dt, n_out=None):
    from scipy.special import gammaln

    m = coll.m
    ratio = float(np.exp(gammaln(3.0 / m) - gammaln(1.0 / m)))
    return ops.collide(
        dev(f), dev(coll.v), coll.dv, dt,
        nu_fp=None if nu_fp is None else dev(nu_fp), nu_K=None if nu_K is None else dev(nu_K),
        f_mx=dev(coll.f_mx[0]), model=MODEL[coll.model], scheme=SCHEME[coll.scheme], nodrag=coll.nodrag,
        sg_m=m, sg_ratio=ratio, n_out=n_out, sc_steps=coll.sc_max_steps, sc_rtol=coll.sc_rtol, sc_atol=coll.sc_atol,
    )


@pytest.mark.parametrize("fp_type", ["lenard_bernstein", "chang_cooper", "dougherty", "chang_cooper_dougherty",
                                     "dougherty_nodrag", "super_gaussian"])
@pytest.mark.parametrize("nx,nv", [(16, 512), (8, 1024), (32, 256), (3, 64), (4, 4096), (2, 6144), (5, 96)])
def test_collide_matches_oracle(ops, fp_type, nx, nv):
    cfg = _fp_cfg(nv, 6.4, fp_type, m=3.0 if fp_type == "super_gaussian" else 2.0)
    coll = O.Collisions(cfg)
    f, x, v, dx, dv = make_f(nx, nv, seed=nv, noise=0.0)
    f = f * (1 + 0.05 * np.sin(7 * v))[None, :]
    nu = np.linspace(0.2, 1.0, nx)
    dt = 0.1
    ref = coll(nu, None, f, dt)
    out = host(_gpu_collide(ops, coll, f, nu, None, dt))
    # dougherty_nodrag subtracts dt*nu*lap(D f_M) explicitly (fokker_planck.py:414-427): a 1-ulp difference in
    # exp() is amplified by dt*nu*D/dv^2 (1e4 at nv=4096), in the reference as much as here.
    # More generally cond(I - dt nu L) ~ 1 + 4 dt nu D / dv^2 (2e4 at nv=6144, nu=1): LAPACK gtsv and the parallel
    # solve both carry cond * eps of rounding error, so the comparison tolerance scales with it.
    amp = dt * nu.max() / dv**2
    assert rel_l2(out, ref) <= max(RTOL, 2e-16 * amp)
    # weakly collisional production regime
    nu2 = 1e-5 * np.ones(nx)
    ref = coll(nu2, None, f, dt)
    out = host(_gpu_collide(ops, coll, f, nu2, None, dt))
    assert rel_l2(out, ref) <= RTOL
    # in between: the cyclic reduction terminates early, after a different number of steps per CTA
    nu3 = np.geomspace(1e-4, 3e-2, nx)
    ref = coll(nu3, None, f, dt)
    out = host(_gpu_collide(ops, coll, f, nu3, None, dt))
    assert rel_l2(out, ref) <= max(RTOL, 2e-16 * dt * nu3.max() / dv**2)


@pytest.mark.parametrize("fp_type", ["lenard_bernstein", "chang_cooper", "dougherty", "chang_cooper_dougherty",
                                     "dougherty_nodrag", "super_gaussian"])
@pytest.mark.parametrize("nx,nv,sc_steps", [(16, 128, 3), (5, 64, 1), (8, 512, 3), (3, 96, 2), (2, 4096, 3)])
def test_collide_self_consistent_beta_matches_oracle(ops, fp_type, nx, nv, sc_steps):
    """terms.fokker_planck.self_consistent_beta (fokker_planck.py:295-301, 391-410): beta refined by Newton on the
    discrete temperature (driftdiffusion.py:161-233) or, super-Gaussian, on the discrete energy flux
    (fokker_planck.py:139-210), then the same operator."""
    cfg = _fp_cfg(nv, 6.0, fp_type, m=3.0 if fp_type == "super_gaussian" else 2.0, sc_steps=sc_steps)
    coll = O.Collisions(cfg)
    assert coll.sc_max_steps == sc_steps
    f, x, v, dx, dv = make_f(nx, nv, seed=nv + sc_steps, noise=0.0, vmax=6.0)
    # rows of different widths and drifts: the discrete and continuum temperatures differ most on narrow rows
    width = np.linspace(0.05 if nv <= 128 else 0.4, 1.5, nx)[:, None]
    drift = np.linspace(-0.7, 0.9, nx)[:, None]
    f = np.exp(-np.abs(v[None, :] - drift) ** coll.m / (2 * width)) * (1 + 0.05 * np.sin(7 * v))[None, :]
    for nu in (np.linspace(0.2, 1.0, nx), 1e-5 * np.ones(nx)):
        ref = coll(nu, None, f, 0.1)
        out = host(_gpu_collide(ops, coll, f, nu, None, 0.1))
        # conditioning as in test_collide_matches_oracle; the rows here reach D = T ~ width.max() = 1.5
        assert rel_l2(out, ref) <= max(RTOL, 4e-16 * 0.1 * nu.max() * width.max() / dv**2)
    # the refinement is not a no-op on this input (otherwise the test would not see it)
    coll0 = O.Collisions(_fp_cfg(nv, 6.0, fp_type, m=coll.m, sc_steps=0))
    nu = np.linspace(0.2, 1.0, nx)
    assert rel_l2(coll0(nu, None, f, 0.1), coll(nu, None, f, 0.1)) > 1e-9


def test_collide_sc_beta_supergaussian_fixed_point_on_gpu(ops):
    """The reference's own known answer (tests/test_vlasov1d/test_super_gaussian_fp.py:127-141) through the C ABI:
    100 settling steps, then 1000 more move f by < 1e-8."""
    from scipy.special import gamma

    m, nv = 3.0, 128
    cfg = _fp_cfg(nv, 6.0, "super_gaussian", m=m, sc_steps=3)
    coll = O.Collisions(cfg)
    v, dv = coll.v, coll.dv
    vm = np.sqrt(gamma(1.0 / m) / gamma(3.0 / m))
    f = np.exp(-(np.abs(v / vm) ** m))
    f = dev((f / np.sum(f * dv))[None, :])
    nu = dev(np.ones(1))
    args = dict(nu_fp=nu, f_mx=None, model=2, scheme=1, sg_m=m, sg_ratio=float(gamma(3.0 / m) / gamma(1.0 / m)),
                sc_steps=3)
    vd = dev(v)
    for _ in range(100):
        f = ops.collide(f, vd, dv, 0.1, **args)
    f_eq = host(f)
    for _ in range(1000):
        f = ops.collide(f, vd, dv, 0.1, **args)
    f_end = host(f)
    assert np.linalg.norm(f_end - f_eq) / np.linalg.norm(f_eq) < 1e-8
    T = lambda g: np.sum(g[0] * v**2 * dv) / np.sum(g[0] * dv)
    assert abs(T(f_end) / T(f_eq) - 1.0) < 1e-8


def test_collide_strongly_collisional(ops):
    nx, nv = 8, 512
    coll = O.Collisions(_fp_cfg(nv, 6.4, "dougherty"))
    f, *_ = make_f(nx, nv, seed=3)
    nu = 200.0 * np.ones(nx)
    ref = coll(nu, None, f, 0.5)
    out = host(_gpu_collide(ops, coll, f, nu, None, 0.5))
    assert rel_l2(out, ref) <= 1e-11  # ill-conditioned limit (nu dt D/dv^2 ~ 1.6e5): allow 10x


@pytest.mark.parametrize("fp_type", ["lenard_bernstein", "dougherty", "chang_cooper", "chang_cooper_dougherty"])
@pytest.mark.parametrize("nx,nv", [(4, 512), (6, 1024), (6, 2048), (2, 4096), (2, 8192)])
def test_fused_vpush_collide_matches_oracle(ops, fp_type, nx, nv):
    """VelocityExponential followed by Collisions (vector_field.py:236-238) in one kernel."""
    coll = O.Collisions(_fp_cfg(nv, 6.4, fp_type))
    f, x, v, dx, dv = make_f(nx, nv, seed=nv + 1, noise=0.0)
    f = f * (1 + 0.05 * np.sin(7 * v))[None, :]
    rng = np.random.default_rng(nv)
    e, dex, pond = 0.3 * rng.standard_normal(nx), 0.01 * rng.standard_normal(nx), 0.02 * rng.standard_normal(nx)
    kvr = np.fft.rfftfreq(nv, d=dv) * 2 * np.pi
    q, m, dt = -1.0, 1.0, 0.1
    # strong, weak (no cyclic-reduction step needed) and in-between (early termination after a few steps)
    for nu in (np.linspace(0.2, 1.0, nx), 1e-5 * np.ones(nx), np.geomspace(1e-4, 3e-2, nx)):
        ref = coll(nu, None, O.velocity_exponential(f, kvr, e + dex, pond, dt, q, m), dt)
        out = host(ops.vpush_collide(dev(f), dev(e), dev(pond), q, m, dt, kvr[1], dev(v), dv, dev(nu),
                                     model=MODEL[coll.model], dex=dev(dex), scheme=SCHEME[coll.scheme]))
        amp = dt * nu.max() / dv**2
        assert rel_l2(out, ref) <= max(RTOL, 2e-16 * amp)


@pytest.mark.parametrize("fp_on", [True, False])
def test_collide_krook_and_density_output(ops, fp_on):
    nx, nv = 12, 256
    cfg = _fp_cfg(nv, 6.4, "dougherty", krook=True, T0=1.3)
    cfg["terms"]["fokker_planck"]["is_on"] = fp_on
    coll = O.Collisions(cfg)
    f, *_ = make_f(nx, nv, seed=9, noise=0.0)
    nu_fp = np.linspace(0.1, 0.5, nx)
    nu_K = np.linspace(0.0, 2.0, nx)
    ref = coll(nu_fp if fp_on else None, nu_K, f, 0.1)
    n_out = torch.empty(nx, dtype=torch.float64, device="cuda")
    out = host(_gpu_collide(ops, coll, f, nu_fp if fp_on else None, nu_K, 0.1, n_out=n_out))
    assert rel_l2(out, ref) <= RTOL
    np.testing.assert_allclose(host(n_out), np.sum(ref, 1) * coll.dv, rtol=1e-13)


def test_collide_conservation_50_steps(ops):
    """reference test_fp_momentum_conservation.py:44-83 through the CUDA kernel."""
    for fp_type, energy_rtol in [("dougherty", 5e-3), ("chang_cooper_dougherty", 1e-5)]:
        nx, nv, vmax = 16, 512, 6.4
        coll = O.Collisions(_fp_cfg(nv, vmax, fp_type))
        v, dv = coll.v, coll.dv
        xx = np.linspace(0, 2 * np.pi, nx, endpoint=False)
        f = (1.0 + 0.5 * np.sin(xx))[:, None] * np.exp(-((v[None, :] - 0.5) ** 2) / 2.0)
        f = f / (np.sum(np.exp(-((v - 0.5) ** 2) / 2.0)) * dv)
        fd, vd, nu = dev(f), dev(v), dev(np.ones(nx))
        for _ in range(50):
            fd = ops.collide(fd, vd, dv, 0.1, nu_fp=nu, model=1, scheme=SCHEME[coll.scheme])
        out = host(fd)
        np.testing.assert_allclose(np.sum(out, 1) * dv, np.sum(f, 1) * dv, rtol=1e-10)
        np.testing.assert_allclose(np.sum(out * v, 1) * dv, np.sum(f * v, 1) * dv, rtol=1e-6)
        np.testing.assert_allclose(np.sum(out * v**2, 1) * dv, np.sum(f * v**2, 1) * dv, rtol=energy_rtol)


# ---------------------------------------------------------------------------------------------------- in-loop saves
@pytest.mark.parametrize("nx,nv,interp", [(16, 64, False), (32, 250, True), (4096, 512, True)])
def test_save_moments_fused_pass(ops, nx, nv, interp):
    """storage.py:119-162 / 286-327: six velocity moments of the (interpolated) state in one pass over f."""
    f0, x, v, dx, dv = make_f(nx, nv, seed=1, noise=0.0)
    f1 = f0 * (1 + 0.05 * np.cos(3 * v))[None, :]
    w = 0.37
    f = f0 + w * (f1 - f0) if interp else f0
    got = host(ops.save_moments(dev(f0), dev(v), dv, dev(f1) if interp else None, w)).reshape(6, nx)
    ref = [np.sum(g, axis=1) * dv for g in (f, f * v, f * v**2, f * v**3, -np.log(np.abs(f)) * np.abs(f), f * f)]
    for k in range(6):
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-12, atol=1e-13 * np.max(np.abs(ref[k])))


@pytest.mark.parametrize("rows,n", [(6, 4096), (1, 1), (3, 5000), (6, 17280)])
def test_row_means_device_and_pinned_out(ops, rows, n):
    """Space average of the saved moments (storage.py:306-323): one launch, into device memory or a pinned host ring."""
    rng = np.random.default_rng(7)
    a = rng.standard_normal((rows, n)) + 1.0
    ref = a.mean(axis=1)
    got = host(ops.row_means(dev(a)))
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-15)
    ring = torch.zeros((4, rows), dtype=torch.float64).pin_memory()
    ops.row_means(dev(a), out=ring[2])
    torch.cuda.synchronize()
    np.testing.assert_allclose(ring[2].numpy(), ref, rtol=1e-13, atol=1e-15)
    assert float(ring[0].abs().sum() + ring[1].abs().sum() + ring[3].abs().sum()) == 0.0


def test_save_moments_log_special_values(ops):
    """The entropy moment uses a table-based logarithm on the normal range and the library on the rest: negative f (the
    reference takes |f|), subnormals, huge values, and f = 0, where -|f| log|f| is NaN in the reference too."""
    nx, nv = 8, 4096
    f, x, v, dx, dv = make_f(nx, nv, seed=3, noise=0.0)
    f[1, 100:200] *= -1.0
    f[2, 7] = 5e-324
    f[2, 9] = 1e-310
    f[3, 11] = 1e300
    f[4, 13] = 1.0
    f[4, 14] = np.nextafter(1.0, 0.0)
    f[5, 2048] = 0.0
    got = host(ops.save_moments(dev(f), dev(v), dv)).reshape(6, nx)
    with np.errstate(divide="ignore", invalid="ignore"):
        ref = np.sum(-np.log(np.abs(f)) * np.abs(f), axis=1) * dv
    assert np.isnan(ref[5]) and np.isnan(got[4][5])
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(got[4][ok], ref[ok], rtol=1e-12)
    np.testing.assert_allclose(got[0], np.sum(f, axis=1) * dv, rtol=1e-12)


def test_field_moments_match_oracle():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200.module import default_scalars, field_moments

    nx, nv = 64, 256
    f, x, v, dx, dv = make_f(nx, nv, seed=5)
    cfg = {"grid": {"species_grids": {"electron": {"v": v, "dv": dv}}, "species_params": {"electron": {"mass": 1.0}},
                    "dx": dx}}
    y = {"electron": f, "e": np.sin(x), "de": np.cos(x), "a": np.linspace(0, 1, nx + 2), "prev_a": np.zeros(nx + 2)}
    yd = {k: dev(val) for k, val in y.items()}
    ref, got = O.field_moments(cfg, y), field_moments(cfg, yd)
    for k in ("n", "j", "v", "-flogf", "f^2"):
        np.testing.assert_allclose(host(got["electron"][k]), ref["electron"][k], rtol=1e-12, atol=1e-14)
    for k in ("p", "q"):  # central moments via raw moments: error relative to the raw-moment scale
        scale = np.max(np.abs(ref["electron"]["n"])) * (np.max(np.abs(ref["electron"]["v"])) + 1.0) ** 3
        np.testing.assert_allclose(host(got["electron"][k]), ref["electron"][k], rtol=0, atol=2e-13 * scale)
    np.testing.assert_allclose(host(got["pond"]), ref["pond"], rtol=1e-13, atol=1e-16)
    sref, sgot = O.default_scalars(cfg, y), default_scalars(cfg, yd)
    for k in sref:
        np.testing.assert_allclose(float(sgot[k]), sref[k], rtol=1e-12, atol=1e-15, err_msg=k)


# ---------------------------------------------------------------------------------------------------- error behaviour
def test_errors_are_loud(ops):
    from adept_b200._lib import AdeptB200Error

    f = torch.zeros(12, 16, dtype=torch.float64, device="cuda")
    v = torch.zeros(16, dtype=torch.float64, device="cuda")
    ops.vdfdx(f, v, 0.1, 1.0)  # nx = 12: even, runs through the chirp-z path
    for nx_bad in (13, 4098):  # odd lengths and even non-powers of two above 4096 have no kernel (and no fallback)
        with pytest.raises(AdeptB200Error, match="power of two"):
            ops.vdfdx(torch.zeros(nx_bad, 16, dtype=torch.float64, device="cuda"), v, 0.1, 1.0)
    with pytest.raises(AdeptB200Error, match="no CPU path"):
        ops.vdfdx(f.cpu(), v, 0.1, 1.0)
    with pytest.raises(AdeptB200Error, match="float64"):
        ops.vdfdx(f.float(), v, 0.1, 1.0)
    g = torch.zeros(16, 16, dtype=torch.float64, device="cuda")
    with pytest.raises(AdeptB200Error, match="in-place"):
        ops.edfdv_spline(g, v, None, 1.0, 1.0, 0.1, 0.1, out=g)


@pytest.mark.parametrize("nx,nv,nxq,nvq", [(32, 256, 16, 64), (64, 512, 32, 512), (17, 33, 41, 29)])
def test_dist_save_interp2d_matches_oracle(ops, nx, nv, nxq, nvq):
    """{t, x, v} distribution save (storage.py:173-181): bilinear interpolation with NaN outside the grid, also on the
    state interpolated between two steps."""
    rng = np.random.default_rng(nx + nv)
    x = np.linspace(0.1, 20.8, nx)
    v = np.linspace(-6.3, 6.3, nv)
    f0, f1 = rng.standard_normal((nx, nv)), rng.standard_normal((nx, nv))
    xq = np.linspace(0.0, 20.94, nxq)       # reaches past both ends of x: NaN there, like the reference
    vq = np.linspace(-6.4, 6.4, nvq)
    for w, b in ((0.0, None), (0.37, f1)):
        ref = O.dist_save_xv(f0 if b is None else f0 + w * (f1 - f0), x, v, xq, vq)
        out = host(ops.interp2d(dev(f0), dev(x), dev(v), dev(xq), dev(vq), None if b is None else dev(f1), w))
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert np.max(np.abs(out[ok] - ref[ok])) <= 1e-13 * max(1.0, np.max(np.abs(ref[ok])))
    # a grid-node query hits the node exactly (searchsorted side="right")
    out = host(ops.interp2d(dev(f0), dev(x), dev(v), dev(x), dev(v)))
    assert np.max(np.abs(out - f0)) <= 1e-14


@pytest.mark.parametrize("nx,nv", [(6, 8), (12, 20), (96, 24), (1028, 16), (1728, 8), (3456, 4), (4094, 4)])
def test_vdfdx_any_even_length(ops, nx, nv):
    """Transform lengths that are not powers of two (stock decks: nx = 1028, 1728, 3456) take the chirp-z path."""
    f, x, v, dx, dv = make_f(nx, nv, seed=nx, noise=1e-3)
    kxr = np.fft.rfftfreq(nx, d=dx) * 2 * np.pi
    ref = O.space_exponential(f, kxr, v, 0.37)
    out = host(ops.vdfdx(dev(f), dev(v), 0.37, kxr[1]))
    assert rel_l2(out, ref) <= RTOL
    # a batch of two members with their own box lengths
    fb = np.stack([f, 0.5 * f[::-1].copy()])
    k1s = [kxr[1], 1.3 * kxr[1]]
    refb = np.stack([O.space_exponential(fb[i], k1s[i] / kxr[1] * kxr, v, -0.21) for i in range(2)])
    outb = host(ops.vdfdx(dev(fb), dev(v), -0.21, 0.0, k1x_batch=dev(np.array(k1s))))
    assert rel_l2(outb, refb) <= RTOL


@pytest.mark.parametrize("nx,nv", [(4, 6), (6, 96), (2, 384), (4, 1028), (2, 3000)])
def test_edfdv_exp_any_even_length(ops, nx, nv):
    f, x, v, dx, dv = make_f(nx, nv, seed=nv, noise=1e-3)
    rng = np.random.default_rng(nv)
    e, dex, pond = 0.3 * rng.standard_normal(nx), 0.01 * rng.standard_normal(nx), 0.02 * rng.standard_normal(nx)
    kvr = np.fft.rfftfreq(nv, d=dv) * 2 * np.pi
    ref = O.velocity_exponential(f, kvr, e + dex, pond, 0.1, -1.0, 1.0)
    out = host(ops.edfdv_exp(dev(f), dev(e), dev(pond), -1.0, 1.0, 0.1, kvr[1], dex=dev(dex)))
    assert rel_l2(out, ref) <= RTOL


@pytest.mark.parametrize("nx", [6, 100, 1028, 1728, 3456])
def test_poisson_any_even_length(ops, nx):
    rng = np.random.default_rng(nx)
    dx = 20.94 / nx
    rho = rng.standard_normal((3, nx))
    kx = 2 * np.pi * np.fft.fftfreq(nx, d=dx)
    ook = np.zeros(nx)
    ook[1:] = 1.0 / kx[1:]
    ref = np.stack([O.poisson(r, ook) for r in rho])
    out = host(ops.poisson(dev(rho), dev(ook)))
    assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref))
    rho_i = 1.0 + 0.1 * rng.standard_normal(nx)
    e = host(ops.poisson(dev(rho_i), dev(kx), mode=1, Te=2.0, lambda_De=0.7))
    assert rel_l2(e, O.boltzmann_poisson(rho_i, kx, 2.0, 0.7)) <= RTOL


@pytest.mark.parametrize("batch,nx", [(1, 32), (1, 4096), (3, 100), (5, 7)])
def test_field_energy_matches_numpy(ops, batch, nx):
    """mean_e2 / mean_de2 of the default save (storage.py:316-317), plain and on the interpolated state."""
    rng = np.random.default_rng(nx)
    e0, de0, e1, de1 = (rng.standard_normal((batch, nx)) for _ in range(4))
    out = host(ops.field_energy(dev(e0), dev(de0)))
    np.testing.assert_allclose(out[:, 0], np.mean(e0**2.0, axis=1), rtol=1e-13)
    np.testing.assert_allclose(out[:, 1], np.mean(de0**2.0, axis=1), rtol=1e-13)
    w = 0.37
    out = host(ops.field_energy(dev(e0), dev(de0), dev(e1), dev(de1), w))
    np.testing.assert_allclose(out[:, 0], np.mean((e0 + w * (e1 - e0)) ** 2.0, axis=1), rtol=1e-13)
    np.testing.assert_allclose(out[:, 1], np.mean((de0 + w * (de1 - de0)) ** 2.0, axis=1), rtol=1e-13)
    from adept_b200._lib import AdeptB200Error

    with pytest.raises(AdeptB200Error, match="both e1 and de1"):
        ops.field_energy(dev(e0), dev(de0), dev(e1), None, w)


@pytest.mark.parametrize("batch,nx", [(1, 512), (1, 1024), (1, 4096), (3, 512), (2, 8192)])
def test_poisson_green_matches_oracle(ops, batch, nx):
    """field.py:221-224 as a circular convolution with green = Re ifft(-i / kx): same operator as the FFT solve."""
    rng = np.random.default_rng(nx)
    dx = 20.94 / nx
    kx = np.fft.fftfreq(nx, d=dx) * 2 * np.pi
    ook = np.zeros(nx)
    ook[1:] = 1.0 / kx[1:]
    green = np.real(np.fft.ifft(-1j * ook))
    rho = 0.01 * rng.standard_normal((batch, nx)) + 1e-2 * np.cos(0.3 * np.arange(nx) * dx)[None, :]
    e = host(ops.poisson_green(dev(rho), dev(green)))
    for b in range(batch):
        ref = O.poisson(rho[b], ook)
        assert np.max(np.abs(e[b] - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref)))
        assert rel_l2(e[b], ref) <= RTOL
    from adept_b200._lib import AdeptB200Error

    with pytest.raises(AdeptB200Error, match="power of two"):
        ops.poisson_green(dev(rho[:, :100].copy()), dev(green[:100].copy()))


def test_field_energy_writes_into_pinned_host_memory(ops):
    rng = np.random.default_rng(5)
    e, de = rng.standard_normal(4096), rng.standard_normal(4096)
    ring = torch.zeros((3, 2), dtype=torch.float64).pin_memory()
    ops.field_energy(dev(e), dev(de), out=ring[1])
    torch.cuda.synchronize()
    np.testing.assert_allclose(ring[1].numpy(), [np.mean(e**2), np.mean(de**2)], rtol=1e-13)
    assert float(ring[0].abs().sum() + ring[2].abs().sum()) == 0.0


@pytest.mark.parametrize("fp_type", ["chang_cooper", "chang_cooper_dougherty"])
def test_reference_relaxation_sweep_on_gpu(ops, fp_type):
    """The reference's own relaxation assertions (tests/test_vlasov1d/test_fp_relaxation.py:84-118: Chang-Cooper,
    self-consistent beta with max_steps = 2, dt = tau, 10 collision times, five initial conditions) through the C ABI,
    and the final distribution against the oracle's."""
    from test_oracle_operators import _relax_metrics, _relax_problems, _relax_run, _sg_cfg

    cfg, v, dv = _sg_cfg(128, fp_type, sc_steps=2)
    coll = O.Collisions(cfg)
    nu, vd = dev(np.ones(1)), dev(v)
    model, scheme = MODEL[coll.model], SCHEME[coll.scheme]

    def gpu_collide(f):
        return host(ops.collide(dev(f), vd, dv, 1.0, nu_fp=nu, model=model, scheme=scheme, sc_steps=2))

    for name, f0 in _relax_problems(v, dv).items():
        hist = _relax_run(gpu_collide, fp_type, f0)
        ref = _relax_run(lambda f: coll(np.ones(1), np.zeros(1), f, 1.0), fp_type, f0)
        assert rel_l2(hist[-1], ref[-1]) <= 1e-11, name  # ten strongly collisional steps (dt nu D / dv^2 ~ 114)
        m = _relax_metrics(hist, v, dv)
        assert abs(m["rel_density"]) < 2e-13, (name, m)
        assert abs(m["T_ratio"] - 1.0) < 5e-3, (name, m)
        assert m["rmse_instant"] < 1e-4, (name, m)
        assert m["positivity"] < 1e-20, (name, m)
        if "dougherty" in fp_type:
            assert m["rmse_expected"] < 1e-2, (name, m)
            assert abs(m["momentum_drift"]) < 5e-5, (name, m)


def test_supergaussian_known_answers_on_gpu(ops):
    """tests/test_vlasov1d/test_super_gaussian_fp.py:144-216 through the C ABI (control, Maxwellian -> super-Gaussian
    with an O(nu dt) energy error, momentum of a drifting initial condition, m = 2 == Chang-Cooper Dougherty)."""
    from scipy.special import gammaln
    from test_oracle_operators import _sg_cfg, check_supergaussian_known_answers

    def make(fp_type, m, sc_steps):
        cfg, v, dv = _sg_cfg(128, fp_type, m=m, sc_steps=sc_steps)
        coll = O.Collisions(cfg)
        mm = coll.m
        ratio = float(np.exp(gammaln(3.0 / mm) - gammaln(1.0 / mm)))
        nu, vd = dev(np.ones(1)), dev(v)

        def collide(f, dt):
            return host(ops.collide(dev(f), vd, dv, dt, nu_fp=nu, model=MODEL[coll.model], scheme=SCHEME[coll.scheme],
                                    sg_m=mm, sg_ratio=ratio, sc_steps=sc_steps))

        return collide, v, dv

    check_supergaussian_known_answers(make)


def test_boltzmann_field_solver_matches_screened_poisson_on_gpu(ops):
    """test_boltzmann_electrons.py:46-75 through the C ABI (velocity sum + Boltzmann-Poisson solve)."""
    nx, nv = 64, 256
    length = 2 * np.pi / 0.1
    dx = length / nx
    x = np.linspace(dx / 2, length - dx / 2, nx)
    kx = np.fft.fftfreq(nx, d=dx) * 2 * np.pi
    vmax, k, eps, Te = 0.64, 0.1, 1e-3, 1.0
    dv = 2 * vmax / nv
    v = np.linspace(-vmax + dv / 2, vmax - dv / 2, nv)
    f = (1 + eps * np.cos(k * x))[:, None] * (np.exp(-(v**2) / 0.02) / np.sqrt(2 * np.pi * 0.01))[None, :]
    rho = torch.empty(nx, dtype=torch.float64, device="cuda")
    ops.moments(dev(f), None, dv, (rho, None, None), scale_b=(1.0, 1.0, 1.0))
    for lam, screening in [(1.0, 1 + k**2), (0.0, 1.0), (None, 1 + k**2 * Te)]:
        e = host(ops.poisson(rho, dev(kx), mode=1, Te=Te, lambda_De=-1.0 if lam is None else lam))
        np.testing.assert_allclose(e, Te * eps * k / screening * np.sin(k * x), atol=1e-8 * eps * k)


def test_collisions_conserve_density_on_asymmetric_grid_gpu():
    """tests/test_vlasov1d/test_asymmetric_velocity_grid.py:109-137 through the host Collisions object and the kernel:
    Fokker-Planck + Krook of resonance.yaml on vmin = -5, vmax = 8 conserve density to 1e-6 and match the oracle."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from copy import deepcopy
    from pathlib import Path

    import yaml

    from adept_b200 import pushers
    from adept_b200.config import build_cfg

    with open(Path(__file__).parent / "golden" / "resonance.yaml") as fh:
        d = yaml.safe_load(fh)
    d["grid"].update(vmin=-5.0, vmax=8.0)
    cfg, grid = build_cfg(deepcopy(d))
    g = cfg["grid"]
    dv = g["species_grids"]["electron"]["dv"]
    f0 = np.asarray(g["species_distributions"]["electron"][1])
    nu = np.ones(f0.shape[0])
    f1 = host(pushers.Collisions(cfg)(dev(nu), dev(nu), dev(f0), grid.dt))
    assert np.all(np.isfinite(f1))
    np.testing.assert_allclose(f1.sum(axis=1) * dv, f0.sum(axis=1) * dv, rtol=1e-6)
    ref = O.Collisions(O.build_cfg(deepcopy(d)))(nu, nu, f0, grid.dt)
    assert rel_l2(f1, ref) <= 1e-11


@pytest.mark.parametrize("nx,nv", [(32, 64), (256, 48), (4096, 16)])
def test_abs_rfft_x_matches_numpy(ops, nx, nv):
    """|rfft_x f| (the spectrum of the {t, kx, v} distribution save, storage.py:189) against numpy's rfft."""
    rng = np.random.default_rng(nx)
    f = rng.standard_normal((nx, nv))
    out = host(ops.abs_rfft_x(dev(f)))
    ref = np.abs(np.fft.rfft(f, axis=0))
    assert out.shape == ref.shape == (nx // 2 + 1, nv)
    assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(ref)


def test_dist_save_kx_block_matches_oracle():
    """{t, kx, v} distribution save through the module's save function against the oracle's restatement (one-sided kx
    axis), also on the state interpolated between two steps."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from adept_b200.module import dist_save

    nx, nv = 64, 128
    rng = np.random.default_rng(5)
    dx = 20.94 / nx
    kxr = 2 * np.pi * np.fft.rfftfreq(nx, d=dx)
    v = np.linspace(-6.35, 6.35, nv)
    f0, f1 = rng.standard_normal((nx, nv)), rng.standard_normal((nx, nv))
    kq = np.linspace(0.0, 3.0, 17)
    vq = np.linspace(-6.4, 6.4, 33)
    cfg = {"grid": {"kxr": kxr, "x": None, "species_grids": {"electron": {"v": v}}}}
    fn = dist_save("electron", kxax=kq, vax=vq)
    for w, y1 in ((0.0, None), (0.4, {"electron": dev(f1)})):
        out = host(fn(cfg, {"electron": dev(f0)}, y1, w))
        ref = O.dist_save_kxv(f0 if y1 is None else f0 + w * (f1 - f0), kxr, v, kq, vq)
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert np.max(np.abs(out[ok] - ref[ok])) <= 1e-12 * np.max(np.abs(ref[ok]))


def test_run_with_the_decks_own_save_block():
    """Vlasov1D.run(save="deck") wires the YAML save: block like get_save_quantities (storage.py:222-283): fields,
    "<species>.<label>" distribution saves, the dfdt diagnostics and the always-on default scalars."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import yaml
    from pathlib import Path

    from adept_b200.module import Vlasov1D

    with open(Path(__file__).parent / "golden" / "epw.yaml") as fh:
        deck = yaml.safe_load(fh)
    deck["grid"]["tmax"] = 2.0
    for k in ("fields", "diag-vlasov-dfdt", "diag-fp-dfdt"):
        deck["save"][k]["t"].update(tmax=2.0, nt=5)
    deck["save"]["electron"]["main"]["t"].update(tmax=2.0, nt=3)
    deck["save"]["electron"]["spec"] = {"t": {"tmin": 0.0, "tmax": 2.0, "nt": 2},
                                        "kx": {"kxmin": 0.0, "kxmax": 2.0, "nkx": 9},
                                        "v": {"vmin": -6.0, "vmax": 6.0, "nv": 25}}
    sim = Vlasov1D(deck)
    _, saved = sim.run(save="deck")
    assert set(saved) == {"fields", "electron.main", "electron.spec", "diag-vlasov-dfdt", "diag-fp-dfdt", "default"}
    assert len(saved["default"]) == sim.grid.nt - 1 or len(saved["default"]) == sim.grid.nt
    assert len(saved["electron.main"]) >= 2 and saved["electron.main"][0].shape == (32, 256)
    assert saved["electron.spec"][-1].shape == (9, 25) and bool(torch.isfinite(saved["electron.spec"][-1]).all())
    assert saved["diag-fp-dfdt"][-1].shape == (32, 256)
    assert "n" in saved["fields"][-1]["electron"] and float(saved["default"][-1]["mean_n_electron"]) > 0.99


# ------------------------------------------------------------------------------ long pencils of mixed length (bigx.cu)
@pytest.mark.parametrize("nx,nv", [(17280, 64), (8640, 128), (6912, 64), (5760, 64), (34560, 64)])
def test_vdfdx_long_mixed_length_matches_oracle(ops, nx, nv):
    """nx = 2^a m beyond one SM's shared memory (configs/vlasov-1d/iaw-turbulence-big*.yaml: nx = 17280 = 128 x 135):
    128-point FFTs, m-point DFTs + phase, inverse -- three launches through a scratch array."""
    f, x, v, dx, dv = make_f(nx, nv, seed=nx, noise=0.05, xmax=966.0)
    kxr = np.fft.rfftfreq(nx, d=dx) * 2 * np.pi
    for dt in (0.25, -0.1):
        ref = O.space_exponential(f, kxr, v, dt)
        fd = dev(f)
        out = host(ops.vdfdx(fd, dev(v), dt, kxr[1]))
        assert rel_l2(out, ref) <= RTOL
        ops.vdfdx(fd, dev(v), dt, kxr[1], out=fd)  # in place
        assert rel_l2(host(fd), ref) <= RTOL


@pytest.mark.parametrize("nx", [17280, 6912])
def test_poisson_long_mixed_length_matches_oracle(ops, nx):
    rng = np.random.default_rng(nx)
    dx = 966.0 / nx
    x = (np.arange(nx) + 0.5) * dx
    kx = np.fft.fftfreq(nx, d=dx) * 2 * np.pi
    ook = np.zeros(nx)
    ook[1:] = 1.0 / kx[1:]
    rho = 0.01 * np.sin(2 * np.pi * x / 966.0) + 1e-3 * rng.standard_normal(nx)
    ref = O.poisson(rho, ook)
    out = host(ops.poisson(dev(rho), dev(ook)))
    assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref))
    rho_b = 1.0 + rho
    for lam in (None, 0.3):
        ref = O.boltzmann_poisson(rho_b, kx, 0.05, lam)
        out = host(ops.poisson(dev(rho_b), dev(kx), mode=1, Te=0.05, lambda_De=-1.0 if lam is None else lam))
        assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref))
