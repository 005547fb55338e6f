import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
from adept_b200.ensemble import EnsembleVlasov1D
from bench import c3_deck
for B in (16, 128, 256):
    decks = []
    for k0 in np.linspace(0.2, 0.4, B):
        d = c3_deck(64, 512)
        d["grid"]["xmax"] = 2 * np.pi / k0
        d["density"]["species-background"]["wavenumber"] = float(k0)
        d["drivers"]["ex"]["0"]["params"].update(k0=float(k0), a0=1e-3, w0=float(np.sqrt(1 + 3 * k0**2)))
        decks.append(d)
    ens = EnsembleVlasov1D(decks)
    ens.t, ens.step_index = 30.0, 300
    for _ in range(10): ens.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): ens.step()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(json.dumps({"B": B, "host_issue_us_per_step": t_host / 200 * 1e6, "wall_us_per_step": t_all / 200 * 1e6}))
