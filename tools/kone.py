"""Run one operator a few times at the C3 size (for ncu captures).   python tools/kone.py <op> [nx nv iters]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import numpy as np
import torch

from adept_b200 import ops

op = sys.argv[1]
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
nv = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
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
fns = {
    "vdfdx": lambda: ops.vdfdx(fd, vd, 0.1, k1x, out=gd),
    "edfdv_exp": lambda: ops.edfdv_exp(fd, e, None, -1.0, 1.0, 0.1, k1v, out=gd),
    "edfdv_spline": lambda: ops.edfdv_spline(fd, e, None, -1.0, 1.0, 0.1, dv, out=gd),
    "collide": lambda: ops.collide(fd, vd, dv, 0.1, nu_fp=nu, model=1, scheme=0, out=gd),
    "collide_cc": lambda: ops.collide(fd, vd, dv, 0.1, nu_fp=nu, model=1, scheme=1, out=gd),
    "moments": lambda: ops.moments(fd, vd, dv, (rho, None, None)),
}
for _ in range(iters):
    fns[op]()
torch.cuda.synchronize()
