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
  Collisions (FP, central, LB/Dougherty)   collide_bwd (transposed solve + moment-chain terms)

``leapfrog_step`` composes them like LeapfrogIntegrator + Collisions (vector_field.py:87-95, 238); gradients are checked
against finite differences in tests/test_gpu_autodiff.py.
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
    def forward(ctx, f, nu_fp, v, dv, dt, model):
        f, nu_fp = f.contiguous(), nu_fp.contiguous()
        out = ops.collide(f, v, dv, dt, nu_fp=nu_fp, model=model, scheme=0)
        ctx.save_for_backward(f, out, nu_fp)
        ctx.c = (v, dv, dt, model)
        return out

    @staticmethod
    def backward(ctx, g):
        f, out, nu_fp = ctx.saved_tensors
        v, dv, dt, model = ctx.c
        fbar, nubar = ops.collide_bwd(f, out, g.contiguous(), v, dv, dt, nu_fp, model=model, scheme=0,
                                      want_nu_bar=ctx.needs_input_grad[1])
        return fbar, nubar, None, None, None, None


def vdfdx(f, v, dt, k1x):
    return _Vdfdx.apply(f, v, float(dt), float(k1x))


def edfdv_exp(f, e, q, m, dt, k1v):
    return _EdfdvExp.apply(f, e, float(q), float(m), float(dt), float(k1v))


def charge_density(f, dv, q, base=None):
    return _ChargeDensity.apply(f, float(dv), float(q), base)


def poisson(rho, one_over_kx):
    return _Poisson.apply(rho, one_over_kx)


def collide_fp(f, nu_fp, v, dv, dt, model=1):
    return _CollideFP.apply(f, nu_fp, v, float(dv), float(dt), int(model))


def leapfrog_step(f, dex, nu_fp, p: dict):
    """One differentiable leapfrog Vlasov-Poisson(-Fokker-Planck) step of a single species (vector_field.py:87-95,
    238): returns (f_new, e).  ``p``: v, dv, dt, k1x, k1v, q, m, one_over_kx, ion (nullable), fp_model (None = off)."""
    fs = vdfdx(f, p["v"], p["dt"], p["k1x"])
    rho = charge_density(fs, p["dv"], p["q"], p.get("ion"))
    e = poisson(rho, p["one_over_kx"])
    f2 = edfdv_exp(fs, e + dex, p["q"], p["m"], p["dt"], p["k1v"])
    if p.get("fp_model") is not None and nu_fp is not None:
        f2 = collide_fp(f2, nu_fp, p["v"], p["dv"], p["dt"], p["fp_model"])
    return f2, e
