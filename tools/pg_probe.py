import sys; sys.path.insert(0, '.')
import numpy as np, torch
from adept_b200 import ops
nx = 4096
rho = torch.randn(nx, dtype=torch.float64, device='cuda'); green = torch.randn(nx, dtype=torch.float64, device='cuda')
ook = torch.randn(nx, dtype=torch.float64, device='cuda')
out = torch.empty_like(rho)
for name, fn in (("poisson_green", lambda: ops.poisson_green(rho, green, out=out)), ("poisson_fft", lambda: ops.poisson(rho, ook, out=out))):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 200 * 1e3, "us per call (back to back)")
