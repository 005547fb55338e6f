import sys; sys.path.insert(0, "tests")
import numpy as np, torch
from test_gpu_step import VARIANTS, rel_l2
from oracle import vlasov1d as O
from adept_b200.module import Vlasov1D
for name, deck in VARIANTS.items():
    sim = Vlasov1D(deck); cfg = O.build_cfg(deck); vf = O.VlasovMaxwell(cfg); y = O.init_state(cfg)
    dt = cfg["grid"]["dt"]; t0 = 30.0 if cfg["drivers"].get("ex") else 0.0
    worst = {}; absw = {}
    for n in range(6):
        t = t0 + n * dt
        st = {k: torch.as_tensor(v, device="cuda").contiguous() for k, v in y.items()}
        yg = sim.vector_field(t, st, None); y = vf(t, y, None)
        for k in y:
            g = yg[k].cpu().numpy()
            worst[k] = max(worst.get(k, 0), rel_l2(g, y[k])); absw[k] = max(absw.get(k, 0), np.abs(g - y[k]).max())
    print(name, {k: f"{worst[k]:.1e}|abs {absw[k]:.1e}|max {np.abs(y[k]).max():.1e}" for k in worst})
