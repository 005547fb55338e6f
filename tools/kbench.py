"""Per-kernel micro-benchmark at the C3 size (development aid, not the bench contract).

    python tools/kbench.py [nx nv [iters]]
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import numpy as np
import torch

from adept_b200 import ops

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
PEAK = 6544.0

vmax = 6.4
dv = 2 * vmax / nv
v = np.linspace(-vmax + dv / 2, vmax - dv / 2, nv)
xmax = 20.94
dx = xmax / nx
x = np.linspace(dx / 2, xmax - dx / 2, nx)
f = (1 + 0.01 * np.cos(0.3 * x))[:, None] * np.exp(-v**2 / 2)[None, :] / (np.sum(np.exp(-v**2 / 2)) * dv)
fd = torch.as_tensor(f, device="cuda")
gd = torch.empty_like(fd)
vd = torch.as_tensor(v, device="cuda")
e = torch.as_tensor(1e-2 * np.sin(0.3 * x), device="cuda")
nu = torch.full((nx,), 1e-5, dtype=torch.float64, device="cuda")
rho = torch.empty(nx, dtype=torch.float64, device="cuda")
k1x, k1v = 2 * np.pi / xmax, 2 * np.pi / (nv * dv)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")


def timeit(name, fn, bytes_per_cell=16.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        t.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(t) * 1e-3)
    ts = np.array(ts)
    gbs = nx * nv * bytes_per_cell / ts / 1e9
    print(f"{name:14s} median {np.median(ts)*1e6:8.1f} us  best {ts.min()*1e6:8.1f} us   "
          f"{np.median(gbs):7.0f} GB/s  = {np.median(gbs)/PEAK*100:5.1f}% of measured HBM peak")


timeit("copy(torch)", lambda: gd.copy_(fd))
timeit("vdfdx", lambda: ops.vdfdx(fd, vd, 0.1, k1x, out=gd))
timeit("edfdv_exp", lambda: ops.edfdv_exp(fd, e, None, -1.0, 1.0, 0.1, k1v, out=gd))
timeit("edfdv_spline", lambda: ops.edfdv_spline(fd, e, None, -1.0, 1.0, 0.1, dv, out=gd))
timeit("moments(n)", lambda: ops.moments(fd, vd, dv, (rho, None, None)), 8.0)
timeit("collide_dough", lambda: ops.collide(fd, vd, dv, 0.1, nu_fp=nu, model=1, scheme=0, out=gd))
timeit("collide_cc", lambda: ops.collide(fd, vd, dv, 0.1, nu_fp=nu, model=1, scheme=1, out=gd))
timeit("poisson", lambda: ops.poisson(rho, rho), 0.0)
parts = torch.zeros((ops.vdfdx_rho_parts(fd), nx), dtype=torch.float64, device="cuda")
timeit("vdfdx_rho(tma)", lambda: ops.vdfdx_rho(fd, vd, 0.1, k1x, parts, out=gd))
timeit("vpush_collide", lambda: ops.vpush_collide(fd, e, None, -1.0, 1.0, 0.1, k1v, vd, dv, nu, model=1, out=gd), 32.0)
nu_cc = nu
timeit("vpush_collide_cc", lambda: ops.vpush_collide(fd, e, None, -1.0, 1.0, 0.1, k1v, vd, dv, nu_cc, model=1, out=gd, scheme=1), 32.0)
mom6 = torch.empty((6, nx), dtype=torch.float64, device="cuda")
timeit("save_moments", lambda: ops.save_moments(fd, vd, dv, out=mom6), 8.0)
timeit("save_mom_interp", lambda: ops.save_moments(fd, vd, dv, gd, 0.3, out=mom6), 16.0)
# single precision (8 B/cell per operator application)
f32, g32 = fd.float(), torch.empty_like(fd, dtype=torch.float32)
timeit("vdfdx_f32", lambda: ops.vdfdx_f32(f32, vd, 0.1, k1x, out=g32), 8.0)
timeit("edfdv_exp_f32", lambda: ops.edfdv_exp_f32(f32, e, None, -1.0, 1.0, 0.1, k1v, out=g32), 8.0)
timeit("collide_f32", lambda: ops.collide_f32(f32, vd, dv, 0.1, nu_fp=nu, model=1, scheme=0, out=g32), 8.0)
# non-power-of-two lengths (chirp-z path), same cell count as a 2048 x 2048 grid for orientation
nb = 3456
fb = torch.as_tensor(np.ascontiguousarray(f[:nb, :1024]), device="cuda")
gb = torch.empty_like(fb)
vb = torch.as_tensor(v[:1024].copy(), device="cuda")
for _name, _fn, _cells in (("vdfdx 3456x1024 (bluestein)", lambda: ops.vdfdx(fb, vb, 0.1, k1x, out=gb), nb * 1024),):
    for _ in range(3):
        _fn()
    torch.cuda.synchronize()
    s_, t_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_.record()
    for _ in range(5):
        _fn()
    t_.record()
    torch.cuda.synchronize()
    us = s_.elapsed_time(t_) * 1e3 / 5
    print(f"{_name:28s} {us:8.1f} us   {_cells * 16 / us / 1e3:7.0f} GB/s")
