"""Thin torch-tensor front end of the C ABI: one function per entry point of include/adept_b200.h.

torch is used for device memory and streams only; every function enqueues exactly one hand-written CUDA kernel
on the current torch stream.  Inputs must be float64 CUDA tensors (no CPU path exists; a CPU tensor raises).
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import AdeptB200Error

LAUNCHES = 0  # number of adept_b200 kernels enqueued by this process (bench.py reports it)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, name, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise AdeptB200Error(f"{name}: tensor required")
    if not isinstance(t, torch.Tensor):
        raise AdeptB200Error(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise AdeptB200Error(f"{name}: adept_b200 has no CPU path; tensor must live on a CUDA device")
    if t.dtype != torch.float64:
        raise AdeptB200Error(f"{name}: expected float64, got {t.dtype}")
    if not t.is_contiguous():
        raise AdeptB200Error(f"{name}: tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def _ptr32(t, name, allow_none=False):
    """float32 twin of :func:`_ptr` for the ``_f32`` entry points."""
    if t is None and allow_none:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
        raise AdeptB200Error(f"{name}: expected a contiguous float32 CUDA tensor")
    return C.c_void_p(t.data_ptr())


def _shape3(f):
    if f.dim() == 2:
        return 1, f.shape[0], f.shape[1]
    if f.dim() == 3:
        return f.shape[0], f.shape[1], f.shape[2]
    raise AdeptB200Error(f"f must be [nx, nv] or [batch, nx, nv], got {tuple(f.shape)}")


def prepare(n: int) -> None:
    """Build the twiddle tables for transform length n on the current device (needed before graph capture)."""
    _lib.check(_lib.load().adept_b200_prepare(int(n)), "prepare")


def is_long_mixed_nx(nx: int, nv: int) -> bool:
    """True for the pencils adept_b200_vdfdx_scratch_f64 serves: nx > 4096, not a power of two, nx = 2^a m with a >= 6
    and a small odd m, nv % 64 == 0 (e.g. nx = 17280 = 128 x 135)."""
    if nx <= 4096 or nx & (nx - 1) == 0 or nv % 64:
        return False
    a, m = 0, nx
    while m % 2 == 0:
        m //= 2
        a += 1
    return a >= 6 and m * (1 << max(a - 8, 0)) <= 150


_SCRATCH = {}


def vdfdx(f, v, dt, k1x, out=None, k1x_batch=None):
    """x-advection (SpaceExponential, vlasov.py:234-251)."""
    b, nx, nv = _shape3(f)
    out = torch.empty_like(f) if out is None else out
    if is_long_mixed_nx(nx, nv):
        key = (str(f.device), tuple(f.shape))
        if key not in _SCRATCH:
            _SCRATCH.clear()  # one work array at a time
            _SCRATCH[key] = torch.empty_like(f)
        rc = _lib.load().adept_b200_vdfdx_scratch_f64(
            _ptr(f, "f"), _ptr(out, "out"), _ptr(_SCRATCH[key], "scratch"), b, nx, nv, _ptr(v, "v"), float(dt),
            float(k1x), _ptr(k1x_batch, "k1x_batch", True), _stream())
        _lib.check(rc, "vdfdx (long pencils)")
        _count(3)
        return out
    rc = _lib.load().adept_b200_vdfdx_f64(
        _ptr(f, "f"), _ptr(out, "out"), b, nx, nv, _ptr(v, "v"), float(dt), float(k1x),
        _ptr(k1x_batch, "k1x_batch", True), _stream(),
    )
    _lib.check(rc, "vdfdx")
    _count()
    return out


def vpush_collide(f, e, pond, q, m, dt, k1v, v, dv, nu_fp, model=1, dex=None, out=None, scheme=0):
    """Fused spectral v-advection + Fokker-Planck step (vector_field.py:236-238), one read and one write of f."""
    b, nx, nv = _shape3(f)
    out = torch.empty_like(f) if out is None else out
    rc = _lib.load().adept_b200_vpush_collide_f64(
        _ptr(f, "f"), _ptr(out, "out"), b, nx, nv, _ptr(e, "e"), _ptr(dex, "dex", True), _ptr(pond, "pond", True),
        float(q), float(m), float(dt), float(k1v), _ptr(v, "v"), float(dv), _ptr(nu_fp, "nu_fp"), int(model),
        int(scheme), _stream())
    _lib.check(rc, "vpush_collide")
    _count()
    return out


def _peer_array(ptrs):
    import ctypes as C

    return (C.c_void_p * len(ptrs))(*[int(q) for q in ptrs])


def vpush_collide_p2p(in_ptrs, out_ptrs, row0_global, nx_local, nv, e, pond, q, m, dt, k1v, v, dv, nu_fp, model=1,
                      dex=None, scheme=0, stage=None, round_counters=None, n_movers=0, nx_global=0):
    """Fused v-advection + Fokker-Planck step of this rank's ``nx_local`` rows of a grid whose buffers are all
    v-sharded ``[nx, nv / P]``: cells are read from and written to the owning ranks' buffers over peer memory
    (``in_ptrs`` / ``out_ptrs``: one device pointer per rank).  With ``stage`` ([nx_local, nv] scratch) and
    ``round_counters`` (int32 [nx_local / n_movers]) the input rows are gathered by ``n_movers`` mover CTAs of the same
    launch (``adept_b200_vpush_collide_p2p_staged_f64``)."""
    if stage is not None:
        rc = _lib.load().adept_b200_vpush_collide_p2p_staged_f64(
            _peer_array(in_ptrs), _peer_array(out_ptrs), len(out_ptrs), int(row0_global), int(nx_local), int(nv),
            _ptr(e, "e"), _ptr(dex, "dex", True), _ptr(pond, "pond", True), float(q), float(m), float(dt), float(k1v),
            _ptr(v, "v"), float(dv), _ptr(nu_fp, "nu_fp"), int(model), int(scheme), _ptr(stage, "stage"),
            None if round_counters is None else C.c_void_p(round_counters.data_ptr()), int(n_movers), int(nx_global),
            _stream())
        _lib.check(rc, "vpush_collide_p2p_staged")
        _count()
        return
    rc = _lib.load().adept_b200_vpush_collide_p2p_f64(
        _peer_array(in_ptrs), _peer_array(out_ptrs), len(out_ptrs), int(row0_global), int(nx_local), int(nv),
        _ptr(e, "e"), _ptr(dex, "dex", True), _ptr(pond, "pond", True), float(q), float(m), float(dt), float(k1v),
        _ptr(v, "v"), float(dv), _ptr(nu_fp, "nu_fp"), int(model), int(scheme), _stream())
    _lib.check(rc, "vpush_collide_p2p")
    _count()


# ---------------------------------------------------------------------------------------------------- adjoints
def edfdv_exp_bwd_accel(f_in, g, e, pond, q, m, dt, k1v, dex=None, out=None):
    """accel_bar[.., i] = sum_j g_ij d f'_ij / d accel_i of :func:`edfdv_exp` (f_in = the forward input)."""
    b, nx, nv = _shape3(f_in)
    out = torch.empty(f_in.shape[:-1], dtype=torch.float64, device=f_in.device) if out is None else out
    rc = _lib.load().adept_b200_edfdv_exp_bwd_accel_f64(
        _ptr(f_in, "f_in"), _ptr(g, "g"), b, nx, nv, _ptr(e, "e"), _ptr(dex, "dex", True), _ptr(pond, "pond", True),
        float(q), float(m), float(dt), float(k1v), _ptr(out, "out"), _stream())
    _lib.check(rc, "edfdv_exp_bwd_accel")
    _count()
    return out


def moments_bwd(out_bars, coefs, v, shape, accumulate_into=None):
    """f_bar = sum_k coefs[k] * out_bars[k][:, None] * v^k (adjoint of :func:`moments`)."""
    dev = next(o for o in out_bars if o is not None).device
    fbar = torch.empty(shape, dtype=torch.float64, device=dev) if accumulate_into is None else accumulate_into
    b, nx, nv = _shape3(fbar)
    PA = C.c_void_p * 3
    arr = PA(*[_ptr(o, f"out_bar{k}", True) for k, o in enumerate(out_bars)])
    cf = (C.c_double * 3)(*[float(c) for c in coefs])
    rc = _lib.load().adept_b200_moments_bwd_f64(arr, cf, b, nx, nv, _ptr(v, "v", True),
                                                int(accumulate_into is not None), _ptr(fbar, "fbar"), _stream())
    _lib.check(rc, "moments_bwd")
    _count()
    return fbar


def collide_bwd(f_in, f_new, g, v, dv, dt, nu_fp, model=1, scheme=0, want_nu_bar=False):
    """(f_bar, nu_bar) of the Fokker-Planck step of :func:`collide` (central differencing, LB / Dougherty)."""
    b, nx, nv = _shape3(f_in)
    fbar = torch.empty_like(f_in)
    nubar = torch.empty(f_in.shape[:-1], dtype=torch.float64, device=f_in.device) if want_nu_bar else None
    rc = _lib.load().adept_b200_collide_bwd_f64(
        _ptr(f_in, "f_in"), _ptr(f_new, "f_new"), _ptr(g, "g"), _ptr(fbar, "fbar"), _ptr(nubar, "nubar", True), b, nx,
        nv, _ptr(v, "v"), float(dv), float(dt), _ptr(nu_fp, "nu_fp"), int(model), int(scheme), _stream())
    _lib.check(rc, "collide_bwd")
    _count()
    return fbar, nubar


def edfdv_spline_bwd(f_in, g, e, pond, q, m, dt, dv, dex=None, want_f=True, want_accel=True):
    """(f_bar, accel_bar) of :func:`edfdv_spline` (f_in = the forward input)."""
    b, nx, nv = _shape3(f_in)
    fbar = torch.empty_like(f_in) if want_f else None
    abar = torch.empty(f_in.shape[:-1], dtype=torch.float64, device=f_in.device) if want_accel else None
    rc = _lib.load().adept_b200_edfdv_spline_bwd_f64(
        _ptr(f_in, "f_in"), _ptr(g, "g"), b, nx, nv, _ptr(e, "e"), _ptr(dex, "dex", True), _ptr(pond, "pond", True),
        float(q), float(m), float(dt), float(dv), _ptr(fbar, "fbar", True), _ptr(abar, "abar", True), _stream())
    _lib.check(rc, "edfdv_spline_bwd")
    _count()
    return fbar, abar


def krook_bwd(f_in, g, dv, dt, nu_K, f_mx, want_nu_bar=False):
    """(f_bar, nu_bar) of the Krook step of :func:`collide`."""
    b, nx, nv = _shape3(f_in)
    fbar = torch.empty_like(f_in)
    nubar = torch.empty(f_in.shape[:-1], dtype=torch.float64, device=f_in.device) if want_nu_bar else None
    rc = _lib.load().adept_b200_krook_bwd_f64(_ptr(f_in, "f_in"), _ptr(g, "g"), b, nx, nv, float(dv), float(dt),
                                              _ptr(nu_K, "nu_K"), _ptr(f_mx, "f_mx"), _ptr(fbar, "fbar"),
                                              _ptr(nubar, "nubar", True), _stream())
    _lib.check(rc, "krook_bwd")
    _count()
    return fbar, nubar


# ----------------------------------------------------------------------------------------------- vlasov-1d2v
def marginal(f, wperp, out=None):
    """F[..., v] = sum_p f[..., v, p] wperp[p] (vector_field.py:36-38 of adept/_vlasov1d2v)."""
    np_ = f.shape[-1]
    rows = f.numel() // np_
    out = torch.empty(f.shape[:-1], dtype=torch.float64, device=f.device) if out is None else out
    rc = _lib.load().adept_b200_marginal_f64(_ptr(f, "f"), _ptr(wperp, "wperp"), rows, np_, _ptr(out, "out"), _stream())
    _lib.check(rc, "marginal")
    _count()
    return out


def transpose_last2(f, out=None):
    """[..., n0, n1] -> [..., n1, n0], out of place."""
    n0, n1 = f.shape[-2], f.shape[-1]
    b = f.numel() // (n0 * n1)
    out = torch.empty(tuple(f.shape[:-2]) + (n1, n0), dtype=torch.float64, device=f.device) if out is None else out
    rc = _lib.load().adept_b200_transpose_f64(_ptr(f, "f"), _ptr(out, "out"), b, n0, n1, _stream())
    _lib.check(rc, "transpose")
    _count()
    return out


def collide_coef(f, v, dv, dt, nu_fp, model=1, scheme=0, nodrag=False, sc_steps=0, sc_rtol=1e-8, sc_atol=1e-12,
                 coef_in=None, coef_out=None, coef_div=1, out=None):
    """Fokker-Planck step whose (vbar, beta) are recorded (coef_out[rows, 2]) or imposed per group of ``coef_div`` rows
    (coef_in[rows / coef_div, 2]); see adept_b200_collide_coef_f64."""
    b, nx, nv = _shape3(f)
    out = torch.empty_like(f) if out is None else out
    rc = _lib.load().adept_b200_collide_coef_f64(
        _ptr(f, "f"), _ptr(out, "out"), b, nx, nv, _ptr(v, "v"), float(dv), float(dt), _ptr(nu_fp, "nu_fp"), int(model),
        int(scheme), int(bool(nodrag)), int(sc_steps), float(sc_rtol), float(sc_atol), _ptr(coef_in, "coef_in", True),
        _ptr(coef_out, "coef_out", True), int(coef_div), _stream())
    _lib.check(rc, "collide_coef")
    _count()
    return out


# ------------------------------------------------------------------------------------- single precision (extra)
def vdfdx_f32(f, v, dt, k1x, out=None, k1x_batch=None):
    """x-advection of a float32 distribution (v, k1x in double: phases are formed in fp64)."""
    b, nx, nv = _shape3(f)
    out = torch.empty_like(f) if out is None else out
    rc = _lib.load().adept_b200_vdfdx_f32(_ptr32(f, "f"), _ptr32(out, "out"), b, nx, nv, _ptr(v, "v"), float(dt),
                                          float(k1x), _ptr(k1x_batch, "k1x_batch", True), _stream())
    _lib.check(rc, "vdfdx_f32")
    _count()
    return out


def edfdv_exp_f32(f, e, pond, q, m, dt, k1v, out=None, dex=None):
    """Spectral v-advection of a float32 distribution (fields in double)."""
    b, nx, nv = _shape3(f)
    out = torch.empty_like(f) if out is None else out
    rc = _lib.load().adept_b200_edfdv_exp_f32(_ptr32(f, "f"), _ptr32(out, "out"), b, nx, nv, _ptr(e, "e"),
                                              _ptr(dex, "dex", True), _ptr(pond, "pond", True), float(q), float(m),
                                              float(dt), float(k1v), _stream())
    _lib.check(rc, "edfdv_exp_f32")
    _count()
    return out


def collide_f32(f, v, dv, dt, nu_fp=None, nu_K=None, f_mx=None, model=1, scheme=0, n_out=None, out=None):
    """Fokker-Planck (LB / Dougherty, central or Chang-Cooper) + Krook on a float32 distribution."""
    b, nx, nv = _shape3(f)
    out = torch.empty_like(f) if out is None else out
    rc = _lib.load().adept_b200_collide_f32(_ptr32(f, "f"), _ptr32(out, "out"), b, nx, nv, _ptr(v, "v"), float(dv),
                                            float(dt), _ptr(nu_fp, "nu_fp", True), _ptr(nu_K, "nu_K", True),
                                            _ptr(f_mx, "f_mx", True), int(model), int(scheme),
                                            _ptr32(n_out, "n_out", True), _stream())
    _lib.check(rc, "collide_f32")
    _count()
    return out


def abs_rfft_x(f, out=None):
    """|rfft(f, axis=x)|: [.., nx/2 + 1, nv] (the spectrum of the {t, kx, v} distribution save, storage.py:189)."""
    b, nx, nv = _shape3(f)
    shape = tuple(f.shape[:-2]) + (nx // 2 + 1, nv)
    out = torch.empty(shape, dtype=torch.float64, device=f.device) if out is None else out
    rc = _lib.load().adept_b200_abs_rfft_x_f64(_ptr(f, "f"), _ptr(out, "out"), b, nx, nv, _stream())
    _lib.check(rc, "abs_rfft_x")
    _count()
    return out


def save_moments(f0, v, dv, f1=None, w=0.0, out=None):
    """[6, batch*nx] = dv * sum_v {f, f v, f v^2, f v^3, -|f| log|f|, f^2} of f = f0 + w (f1 - f0) (storage.py:119-162,
    286-327), one pass over f, the interpolated distribution is never materialised."""
    b, nx, nv = _shape3(f0)
    out = torch.empty((6, b * nx), dtype=torch.float64, device=f0.device) if out is None else out
    rc = _lib.load().adept_b200_save_moments_f64(_ptr(f0, "f0"), _ptr(f1, "f1", True), float(w), b, nx, nv,
                                                 _ptr(v, "v"), float(dv), _ptr(out, "out"), _stream())
    _lib.check(rc, "save_moments")
    _count()
    return out


def interp2d(f0, x, v, xq, vq, f1=None, w=0.0, out=None):
    """Bilinear interpolation of f[nx, nv] (or of f0 + w (f1 - f0)) on the mesh xq x vq; NaN outside the grid."""
    if f0.dim() != 2:
        raise _lib.AdeptB200Error("interp2d: f must be [nx, nv]")
    nx, nv = f0.shape
    out = torch.empty((xq.numel(), vq.numel()), dtype=torch.float64, device=f0.device) if out is None else out
    rc = _lib.load().adept_b200_interp2d_f64(
        _ptr(f0, "f0"), _ptr(f1, "f1", True), float(w), nx, nv, _ptr(x, "x"), _ptr(v, "v"), _ptr(xq, "xq"),
        _ptr(vq, "vq"), int(xq.numel()), int(vq.numel()), _ptr(out, "out"), _stream())
    _lib.check(rc, "interp2d")
    _count()
    return out


def filter_x(f, filt, zeros_v, out=None):
    """Real per-mode multiplier along x: irfft(filt[:, None] * rfft(f, axis=x)) (HouLiFilter, vlasov.py:215-220)."""
    b, nx, nv = _shape3(f)
    out = torch.empty_like(f) if out is None else out
    rc = _lib.load().adept_b200_filter_x_f64(_ptr(f, "f"), _ptr(out, "out"), b, nx, nv, _ptr(filt, "filt"),
                                             _ptr(zeros_v, "zeros_v"), _stream())
    _lib.check(rc, "filter_x")
    _count()
    return out


def vdfdx_rho_parts(f) -> int:
    """Rows of the partial-sum scratch that :func:`vdfdx_rho` needs for a distribution of this shape."""
    b, nx, nv = _shape3(f)
    return int(_lib.load().adept_b200_vdfdx_rho_parts(b, nx, nv))


def vdfdx_rho(f, v, dt, k1x, parts, out=None, k1x_batch=None):
    """x-advection fused with the first stage of the velocity sum of its result (vlasov.py:234-251 + field.py:197-208).

    ``parts`` is a [nparts, batch*nx] scratch tensor (nparts >= vdfdx_rho_parts(f)); finish with :func:`reduce_parts`."""
    b, nx, nv = _shape3(f)
    out = torch.empty_like(f) if out is None else out
    rc = _lib.load().adept_b200_vdfdx_rho_f64(
        _ptr(f, "f"), _ptr(out, "out"), b, nx, nv, _ptr(v, "v"), float(dt), float(k1x),
        _ptr(k1x_batch, "k1x_batch", True), _ptr(parts, "parts"), int(parts.shape[0]), _stream(),
    )
    _lib.check(rc, "vdfdx_rho")
    _count(2)
    return out


def copy2d(dst_ptr, dst_pitch, src_ptr, src_pitch, width, height):
    """``height`` rows of ``width`` doubles from ``src_ptr`` to ``dst_ptr`` (raw device addresses, pitches in doubles) as
    one copy-engine transfer on the current stream; ``src_ptr`` may be a peer-mapped buffer."""
    rc = _lib.load().adept_b200_copy2d_f64(C.c_void_p(int(dst_ptr)), int(dst_pitch), C.c_void_p(int(src_ptr)),
                                           int(src_pitch), int(width), int(height), _stream())
    _lib.check(rc, "copy2d")


def sum_peers(ptrs, n, out):
    """out[i] = sum_r peer_r[i] in rank order over peer-mapped buffers (``ptrs``: one device pointer per rank)."""
    rc = _lib.load().adept_b200_sum_peers_f64(_peer_array(ptrs), len(ptrs), int(n), _ptr(out, "out"), _stream())
    _lib.check(rc, "sum_peers")
    _count()
    return out


def vdfdx_field_peers(f, v_loc, dt, k1x, parts, out, peers):
    """First launch of a sharded-grid step: x-advection of this rank's columns with the charge-density exchange over
    peer memory and the field solve of the whole grid in the tail (``adept_b200_vdfdx_field_peers_f64``).  ``peers``:
    dict with the inbox / flag pointers of every rank, this rank's scratch tensors and the driver's host factors (see
    ``ShardedVlasov1D._setup_p2p``); its ``rho``, ``e`` and ``dex`` tensors are overwritten."""
    nx, nvl = f.shape
    fp = _lib.FieldPeers()
    P = len(peers["share_ptrs"])
    fp.n_peers, fp.my_rank, fp.epoch = P, int(peers["rank"]), int(peers["epoch"])
    for r in range(P):
        fp.share_in[r], fp.flag_in[r] = int(peers["share_ptrs"][r]), int(peers["flag_ptrs"][r])
    fp.sync_counter = peers["counter"].data_ptr()
    fp.ion_share = None if peers["ion_share"] is None else _ptr(peers["ion_share"], "ion_share").value
    fp.dv, fp.charge, fp.dx = float(peers["dv"]), float(peers["charge"]), float(peers["dx"])
    fp.green, fp.rho, fp.e = (_ptr(peers[k], k).value for k in ("green", "rho", "e"))
    fp.dex, fp.pond, fp.a_zero = (_ptr(peers[k], k).value for k in ("dex", "pond", "a_zero"))
    n_ex = len(peers["ex_w"])
    fp.n_ex = n_ex
    if n_ex:
        fp.ex_space, fp.ex_kx = _ptr(peers["ex_space"], "ex_space").value, _ptr(peers["ex_kx"], "ex_kx").value
        for d in range(n_ex):
            fp.ex_w[d], fp.ex_a0[d] = float(peers["ex_w"][d]), float(peers["ex_a0"][d])
            fp.ex_tenv[d], fp.ex_wt[d] = float(peers["ex_tenv"][d]), float(peers["ex_wt"][d])
    rc = _lib.load().adept_b200_vdfdx_field_peers_f64(_ptr(f, "f"), _ptr(out, "out"), nx, nvl, _ptr(v_loc, "v"),
                                                      float(dt), float(k1x), _ptr(parts, "parts"), int(parts.shape[0]),
                                                      C.byref(fp), _stream())
    _lib.check(rc, "vdfdx_field_peers")
    _count()
    return out


def ex_driver(ex_space, ex_kx, w, a0, tenv, wt, out=None):
    """dex[i] = sum_d ((tenv[d] space[d, i]) w[d]) a0[d] sin(kx[d, i] - wt[d]) (field.py:21-33), one launch; ex_space /
    ex_kx: [n_ex, n] device tensors (or None when there is no driver), the rest host sequences."""
    n_ex = len(w)
    if n_ex == 0:
        return out.zero_() if out is not None else None
    n = ex_space.shape[-1]
    out = torch.empty(n, dtype=torch.float64, device=ex_space.device) if out is None else out
    arr = lambda xs: (C.c_double * n_ex)(*[float(x) for x in xs])  # noqa: E731
    rc = _lib.load().adept_b200_ex_driver_f64(_ptr(ex_space, "ex_space"), _ptr(ex_kx, "ex_kx"), n_ex, arr(w), arr(a0),
                                              arr(tenv), arr(wt), n, _ptr(out, "out"), _stream())
    _lib.check(rc, "ex_driver")
    _count()
    return out


def reduce_parts(parts, scale_a, scale_b=1.0, base=None, out=None):
    """out = base + scale_b * ((sum_p parts[p]) * scale_a), fixed summation order."""
    n = parts.shape[1]
    out = torch.empty(n, dtype=torch.float64, device=parts.device) if out is None else out
    rc = _lib.load().adept_b200_reduce_parts_f64(
        _ptr(parts, "parts"), int(parts.shape[0]), n, float(scale_a), float(scale_b), _ptr(base, "base", True),
        _ptr(out, "out"), _stream(),
    )
    _lib.check(rc, "reduce_parts")
    _count()
    return out


def edfdv_exp(f, e, pond, q, m, dt, k1v, dex=None, out=None):
    """Spectral v-advection (VelocityExponential, vlasov.py:74-91); e := e + dex inside the kernel."""
    b, nx, nv = _shape3(f)
    out = torch.empty_like(f) if out is None else out
    rc = _lib.load().adept_b200_edfdv_exp_f64(
        _ptr(f, "f"), _ptr(out, "out"), b, nx, nv, _ptr(e, "e"), _ptr(dex, "dex", True), _ptr(pond, "pond", True),
        float(q), float(m), float(dt), float(k1v), _stream(),
    )
    _lib.check(rc, "edfdv_exp")
    _count()
    return out


def edfdv_spline(f, e, pond, q, m, dt, dv, dex=None, out=None):
    """Cubic-spline v-advection (VelocityCubicSpline, vlasov.py:106-172); out-of-place."""
    b, nx, nv = _shape3(f)
    out = torch.empty_like(f) if out is None else out
    rc = _lib.load().adept_b200_edfdv_spline_f64(
        _ptr(f, "f"), _ptr(out, "out"), b, nx, nv, _ptr(e, "e"), _ptr(dex, "dex", True), _ptr(pond, "pond", True),
        float(q), float(m), float(dt), float(dv), _stream(),
    )
    _lib.check(rc, "edfdv_spline")
    _count()
    return out


def moments(f, v, scale_a, outs, bases=(None, None, None), scale_b=(1.0, 1.0, 1.0)):
    """out_k = base_k + scale_b[k] * (sum_j f v^k * scale_a) for each non-None out_k (k = 0, 1, 2)."""
    b, nx, nv = _shape3(f)
    PA = C.c_void_p * 3
    out_arr = PA(*[(_ptr(o, f"out{k}", True)) for k, o in enumerate(outs)])
    base_arr = PA(*[(_ptr(o, f"base{k}", True)) for k, o in enumerate(bases)])
    sb = (C.c_double * 3)(*[float(s) for s in scale_b])
    rc = _lib.load().adept_b200_moments_f64(
        _ptr(f, "f"), b, nx, nv, _ptr(v, "v", True), float(scale_a), base_arr, out_arr, sb, _stream()
    )
    _lib.check(rc, "moments")
    _count()
    return outs


def poisson(rho, kmul, mode=0, Te=1.0, lambda_De=-1.0, out=None):
    """Spectral field solve (field.py:210-224 / 282-298).  rho [nx] or [batch, nx]; kmul [nx] or [batch, nx]."""
    nx = rho.shape[-1]
    batch = rho.numel() // nx
    stride = nx if (kmul.dim() == 2 and kmul.shape[0] == batch and batch > 1) else 0
    out = torch.empty_like(rho) if out is None else out
    rc = _lib.load().adept_b200_poisson_f64(
        _ptr(rho, "rho"), _ptr(kmul, "kmul"), stride, _ptr(out, "out"), batch, nx, int(mode), float(Te),
        float(lambda_De), _stream(),
    )
    _lib.check(rc, "poisson")
    _count()
    return out


def poisson_green(rho, green, out=None):
    """Plain Poisson solve as a circular convolution with ``green = Re ifft(-i / kx)`` (field.py:221-224 is linear in
    rho), spread over the whole GPU.  rho [nx] or [batch, nx]; green [nx] (shared) or [batch, nx]."""
    nx = rho.shape[-1]
    batch = rho.numel() // nx
    stride = nx if (green.dim() == 2 and green.shape[0] == batch and batch > 1) else 0
    out = torch.empty_like(rho) if out is None else out
    rc = _lib.load().adept_b200_poisson_green_f64(_ptr(rho, "rho"), _ptr(green, "green"), stride, _ptr(out, "out"),
                                                  batch, nx, _stream())
    _lib.check(rc, "poisson_green")
    _count()
    return out


def row_means(a, out=None):
    """out[r] = mean(a[r, :]) for a 2-D tensor, one launch (the space average of the saved moments, storage.py:306-323).
    ``out`` may be a PINNED host tensor like :func:`field_energy`'s: valid after the stream has been synchronised."""
    if a.dim() != 2:
        raise AdeptB200Error("row_means: expected a 2-D tensor")
    rows, n = a.shape
    if out is None:
        out = torch.empty(rows, dtype=torch.float64, device=a.device)
    if not out.is_cuda and out.is_pinned() and out.dtype == torch.float64 and out.is_contiguous():
        out_ptr = C.c_void_p(out.data_ptr())
    else:
        out_ptr = _ptr(out, "out")
    rc = _lib.load().adept_b200_row_means_f64(_ptr(a, "a"), int(rows), int(n), out_ptr, _stream())
    _lib.check(rc, "row_means")
    _count()
    return out


def field_energy(e, de, e1=None, de1=None, w=0.0, out=None):
    """{mean(e^2), mean(de^2)} per member in one launch (storage.py:316-317), optionally of the state interpolated
    towards (e1, de1) with weight w.  Returns a [batch, 2] (or [2]) tensor.  ``out`` may be a PINNED host tensor
    (page-locked memory is mapped into the device's address space): the kernel then writes the scalars straight into
    host memory, which is the device-to-host transfer of a per-step diagnostic without a separate copy; the values are
    valid after the stream has been synchronised."""
    nx = e.shape[-1]
    batch = e.numel() // nx
    if out is None:
        out = torch.empty(e.shape[:-1] + (2,), dtype=torch.float64, device=e.device)
    if not out.is_cuda and out.is_pinned() and out.dtype == torch.float64 and out.is_contiguous():
        out_ptr = C.c_void_p(out.data_ptr())
    else:
        out_ptr = _ptr(out, "out")
    rc = _lib.load().adept_b200_field_energy_f64(_ptr(e, "e"), _ptr(de, "de"), _ptr(e1, "e1", True),
                                                 _ptr(de1, "de1", True), float(w), batch, nx, out_ptr, _stream())
    _lib.check(rc, "field_energy")
    _count()
    return out


def axpy(a, b, s, out=None):
    """out = a + s*b (Ampere update, field.py:354)."""
    out = torch.empty_like(a) if out is None else out
    rc = _lib.load().adept_b200_axpy_f64(_ptr(a, "a"), _ptr(b, "b"), float(s), _ptr(out, "out"), a.numel(), _stream())
    _lib.check(rc, "axpy")
    _count()
    return out


def ponderomotive(a, dx, out=None):
    """pond = -0.5 gradient(a^2, dx)[1:-1] (field.py:495); a is [nx+2] or [batch, nx+2]."""
    nx = a.shape[-1] - 2
    batch = a.numel() // (nx + 2)
    out = a.new_empty(a.shape[:-1] + (nx,)) if out is None else out
    rc = _lib.load().adept_b200_ponderomotive_f64(_ptr(a, "a"), _ptr(out, "out"), batch, nx, float(dx), _stream())
    _lib.check(rc, "ponderomotive")
    _count()
    return out


def wave_step(a, aold, djy, ne_n, ne_np1, c, dx, dt, out=None):
    """One wave-equation step (WaveSolver, field.py:109-157); returns the new a."""
    nx = a.shape[-1] - 2
    batch = a.numel() // (nx + 2)
    out = torch.empty_like(a) if out is None else out
    rc = _lib.load().adept_b200_wave_step_f64(
        _ptr(a, "a"), _ptr(aold, "aold"), _ptr(djy, "djy"), _ptr(ne_n, "ne_n", True), _ptr(ne_np1, "ne_np1", True),
        _ptr(out, "out"), batch, nx, float(c), float(dx), float(dt), _stream(),
    )
    _lib.check(rc, "wave_step")
    _count()
    return out


def collide(f, v, dv, dt, nu_fp=None, nu_K=None, f_mx=None, model=1, scheme=0, nodrag=False, sg_m=2.0, sg_ratio=0.5,
            n_out=None, out=None, sc_steps=0, sc_rtol=1e-8, sc_atol=1e-12):
    """Fokker-Planck (delta-form implicit) + Krook (fokker_planck.py:368-484); ``sc_steps > 0`` refines beta by the
    self-consistent Newton solve first (driftdiffusion.py:161-283, fokker_planck.py:139-210)."""
    b, nx, nv = _shape3(f)
    out = torch.empty_like(f) if out is None else out
    rc = _lib.load().adept_b200_collide_sc_f64(
        _ptr(f, "f"), _ptr(out, "out"), b, nx, nv, _ptr(v, "v"), float(dv), float(dt), _ptr(nu_fp, "nu_fp", True),
        _ptr(nu_K, "nu_K", True), _ptr(f_mx, "f_mx", True), int(model), int(scheme), int(bool(nodrag)), float(sg_m),
        float(sg_ratio), _ptr(n_out, "n_out", True), int(sc_steps), float(sc_rtol), float(sc_atol), _stream(),
    )
    _lib.check(rc, "collide")
    _count()
    return out
