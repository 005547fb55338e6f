"""Reverse-mode differentiation of the vlasov-1d step through the CUDA operators.

In the reference the step is differentiated by JAX (``jax.vjp`` of the vector field per step under diffrax's
RecursiveCheckpointAdjoint); a drop-in puts a ``jax.custom_vjp`` around each FFI operator whose backward
rule calls the adjoint entry points of ``include/adept_b200.h``.  jax is not available in this image, so
the same forward/backward pairs are registered here as ``torch.autograd.Function`` -- plumbing only: every forward and
every backward below is one call into libadept_b200.so.

  operator (forward)                       backward
  SpaceExponential        vdfdx            vdfdx with -dt                       (real circulant, unit-modulus symbol)
  VelocityExponential     edfdv_exp        w.r.t. f: edfdv_exp with -dt; w.r.t. e: edfdv_exp_bwd_accel * q/m
  compute_charge_density  moments          moments_bwd (broadcast along v)
  SpectralPoissonSolver   poisson          -poisson (antisymmetric operator)
  VelocityCubicSpline     edfdv_spline     edfdv_spline_bwd (taps scattered back; chain through the shift)
  Collisions (FP, central or Chang-Cooper, LB/Dougherty)   collide_bwd (transposed solve + moment-chain terms)
  Krook                                    krook_bwd

``leapfrog_step`` / ``sixth_step`` compose them like LeapfrogIntegrator / SixthOrderHamIntegrator + Collisions
(vector_field.py:87-95, 118-186, 238); gradients are checked against finite differences in tests/test_gpu_autodiff.py.
"""

from __future__ import annotations

import torch

from . import ops


class _Vdfdx(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, v, dt, k1x):
        ctx.v, ctx.dt, ctx.k1x = v, dt, k1x
        return ops.vdfdx(f.contiguous(), v, dt, k1x)

    @staticmethod
    def backward(ctx, g):
        return ops.vdfdx(g.contiguous(), ctx.v, -ctx.dt, ctx.k1x), None, None, None


class _EdfdvExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, e, q, m, dt, k1v):
        f, e = f.contiguous(), e.contiguous()
        ctx.save_for_backward(f, e)
        ctx.c = (q, m, dt, k1v)
        return ops.edfdv_exp(f, e, None, q, m, dt, k1v)

    @staticmethod
    def backward(ctx, g):
        f, e = ctx.saved_tensors
        q, m, dt, k1v = ctx.c
        g = g.contiguous()
        fbar = ops.edfdv_exp(g, e, None, q, m, -dt, k1v) if ctx.needs_input_grad[0] else None
        ebar = None
        if ctx.needs_input_grad[1]:
            ebar = ops.edfdv_exp_bwd_accel(f, g, e, None, q, m, dt, k1v) * (q / m)  # accel = (q e + ...)/m
        return fbar, ebar, None, None, None, None


class _ChargeDensity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, dv, q, base):
        f = f.contiguous()
        ctx.shape, ctx.coef = f.shape, dv * q
        out = torch.empty(f.shape[:-1], dtype=torch.float64, device=f.device)
        ops.moments(f, None, dv, (out, None, None), bases=(base, None, None), scale_b=(q, 1.0, 1.0))
        return out

    @staticmethod
    def backward(ctx, g):
        fbar = ops.moments_bwd((g.contiguous(), None, None), (ctx.coef, 0.0, 0.0), None, ctx.shape)
        return fbar, None, None, (g if ctx.needs_input_grad[3] else None)


class _Poisson(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rho, one_over_kx):
        ctx.k = one_over_kx
        return ops.poisson(rho.contiguous(), one_over_kx)

    @staticmethod
    def backward(ctx, g):
        return -ops.poisson(g.contiguous(), ctx.k), None


class _CollideFP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, nu_fp, v, dv, dt, model, scheme):
        f, nu_fp = f.contiguous(), nu_fp.contiguous()
        out = ops.collide(f, v, dv, dt, nu_fp=nu_fp, model=model, scheme=scheme)
        ctx.save_for_backward(f, out, nu_fp)
        ctx.c = (v, dv, dt, model, scheme)
        return out

    @staticmethod
    def backward(ctx, g):
        f, out, nu_fp = ctx.saved_tensors
        v, dv, dt, model, scheme = ctx.c
        fbar, nubar = ops.collide_bwd(f, out, g.contiguous(), v, dv, dt, nu_fp, model=model, scheme=scheme,
                                      want_nu_bar=ctx.needs_input_grad[1])
        return fbar, nubar, None, None, None, None, None


class _EdfdvSpline(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, e, q, m, dt, dv):
        f, e = f.contiguous(), e.contiguous()
        ctx.save_for_backward(f, e)
        ctx.c = (q, m, dt, dv)
        return ops.edfdv_spline(f, e, None, q, m, dt, dv)

    @staticmethod
    def backward(ctx, g):
        f, e = ctx.saved_tensors
        q, m, dt, dv = ctx.c
        fbar, abar = ops.edfdv_spline_bwd(f, g.contiguous(), e, None, q, m, dt, dv, want_f=ctx.needs_input_grad[0],
                                          want_accel=ctx.needs_input_grad[1])
        return fbar, (abar * (q / m) if abar is not None else None), None, None, None, None


class _Krook(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, nu_K, v, dv, dt, f_mx):
        f, nu_K = f.contiguous(), nu_K.contiguous()
        ctx.save_for_backward(f, nu_K)
        ctx.c = (dv, dt, f_mx)
        return ops.collide(f, v, dv, dt, nu_K=nu_K, f_mx=f_mx)

    @staticmethod
    def backward(ctx, g):
        f, nu_K = ctx.saved_tensors
        dv, dt, f_mx = ctx.c
        fbar, nubar = ops.krook_bwd(f, g.contiguous(), dv, dt, nu_K, f_mx, want_nu_bar=ctx.needs_input_grad[1])
        return fbar, nubar, None, None, None, None


def vdfdx(f, v, dt, k1x):
    return _Vdfdx.apply(f, v, float(dt), float(k1x))


def edfdv_exp(f, e, q, m, dt, k1v):
    return _EdfdvExp.apply(f, e, float(q), float(m), float(dt), float(k1v))


def charge_density(f, dv, q, base=None):
    return _ChargeDensity.apply(f, float(dv), float(q), base)


def poisson(rho, one_over_kx):
    return _Poisson.apply(rho, one_over_kx)


def collide_fp(f, nu_fp, v, dv, dt, model=1, scheme=0):
    return _CollideFP.apply(f, nu_fp, v, float(dv), float(dt), int(model), int(scheme))


def edfdv_spline(f, e, q, m, dt, dv):
    return _EdfdvSpline.apply(f, e, float(q), float(m), float(dt), float(dv))


def krook(f, nu_K, v, dv, dt, f_mx):
    return _Krook.apply(f, nu_K, v, float(dv), float(dt), f_mx)


def leapfrog_step(f, dex, nu_fp, p: dict):
    """One differentiable leapfrog Vlasov-Poisson(-Fokker-Planck) step of a single species (vector_field.py:87-95,
    238): returns (f_new, e).  ``p``: v, dv, dt, k1x, k1v, q, m, one_over_kx, ion (nullable), fp_model (None = off)."""
    fs = vdfdx(f, p["v"], p["dt"], p["k1x"])
    rho = charge_density(fs, p["dv"], p["q"], p.get("ion"))
    e = poisson(rho, p["one_over_kx"])
    f2 = _push_v(fs, e + dex, p["dt"], p)
    return _collisions(f2, nu_fp, p), e


def _push_v(f, e_total, dt, p):
    if p.get("edfdv", "exponential") == "cubic-spline":
        return edfdv_spline(f, e_total, p["q"], p["m"], dt, p["dv"])
    return edfdv_exp(f, e_total, p["q"], p["m"], dt, p["k1v"])


def _collisions(f, nu_fp, p, nu_K=None):
    if p.get("fp_model") is not None and nu_fp is not None:
        f = collide_fp(f, nu_fp, p["v"], p["dv"], p["dt"], p["fp_model"], p.get("fp_scheme", 0))
    if nu_K is not None:
        f = krook(f, nu_K, p["v"], p["dv"], p["dt"], p["f_mx"])
    return f


SIXTH_A = (0.168735950563437422448196, 0.377851589220928303880766, -0.093175079568731452657924)
_B = (0.049086460976116245491441, 0.264177609888976700200146, 0.186735929134907054308413)
_C = (-0.000069728715055305084099, -0.000625704827430047189169, -0.002213085124045325561636)
_D = (0.0, -2.916600457689847816445691e-6, 3.048480261700038788680723e-5)
_E3 = 4.985549387875068121593988e-7


def sixth_substep_times(dt):
    """Offsets t + dt_array[s] at which the driver field of substep s is evaluated (vector_field.py:148-157)."""
    a1, a2, a3 = SIXTH_A
    return [0.0, a1 * dt, (a1 + a2) * dt, (a1 + a2 + a3) * dt, (a1 + a2 + a3 + a2) * dt, (a1 + a2 + a3 + a2 + a1) * dt]


def sixth_step(f, dex, nu_fp, p: dict, nu_K=None):
    """One differentiable sixth-order Hamiltonian-splitting step (SixthOrderHamIntegrator, vector_field.py:118-186) +
    collisions of a single species: ``dex`` is the list of the six driver fields at ``sixth_substep_times``; returns
    (f_new, e) with e the field of the last substep."""
    dt = p["dt"]
    a1, a2, a3 = SIXTH_A
    D1 = _B[0] + 2.0 * _C[0] * dt**2.0
    D2 = _B[1] + 2.0 * _C[1] * dt**2.0 + 4.0 * _D[1] * dt**4.0
    D3 = _B[2] + 2.0 * _C[2] * dt**2.0 + 4.0 * _D[2] * dt**4.0 - 8.0 * _E3 * dt**6.0
    kicks, drifts = (D1, D2, D3, D3, D2, D1), (a1, a2, a3, a2, a1)
    e = None
    for i in range(6):
        rho = charge_density(f, p["dv"], p["q"], p.get("ion"))
        e = poisson(rho, p["one_over_kx"])
        f = _push_v(f, e + dex[i], kicks[i] * dt, p)
        if i < 5:
            f = vdfdx(f, p["v"], drifts[i] * dt, p["k1x"])
    return _collisions(f, nu_fp, p, nu_K), e
