"""Time the reference's one published benchmark deck (configs/vlasov-1d/iaw-turbulence-big-bench.yaml: nx = 17280,
nv = 2048, sixth-order, cubic-spline, Boltzmann electrons, stochastic forcing; 45 ms/step on 4 x A100-40GB per
configs/vlasov-1d/run-iaw-big.sbatch) on one B200."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, yaml
from adept_b200 import _lib
from adept_b200.module import Vlasov1D
import ctypes

def deck(time="sixth", tmax=50.0):
    return yaml.safe_load(f"""
units: {{normalizing_temperature: 1000eV, normalizing_density: 1.0e19/cc, reference: ion, A: 1.0, Z: 1.0}}
density:
  quasineutrality: true
  species-ion-background: {{noise_seed: 416, noise_type: gaussian, noise_val: 0.0, v0: 0.0, T0: 1.0, m: 2.0, basis: uniform, baseline: 1.0}}
grid: {{dt: 0.25, nx: 17280, tmin: 0.0, tmax: {tmax}, xmax: 966.0, xmin: 0.0}}
save: {{}}
solver: vlasov-1d
drivers:
  ex: {{}}
  ey: {{}}
  ex_stochastic: {{modes: [1], amplitude: 0.0010352, tau: 966.0, seed: 42}}
diagnostics: {{diag-vlasov-dfdt: false, diag-fp-dfdt: false}}
terms:
  field: poisson-boltzmann
  boltzmann_electrons: {{Te: 0.05}}
  edfdv: cubic-spline
  time: {time}
  species:
  - {{name: ion, charge: 1.0, mass: 1.0, vmax: 6.4, nv: 2048, density_components: [species-ion-background]}}
  fokker_planck: {{is_on: false, type: Dougherty, time: {{baseline: 1.0, bump_or_trough: bump, center: 0.0, rise: 25.0, slope: 0.0, bump_height: 0.0, width: 100000.0}}, space: {{baseline: 1.0, bump_or_trough: bump, center: 0.0, rise: 25.0, slope: 0.0, bump_height: 0.0, width: 100000.0}}}}
  krook: {{is_on: false, time: {{baseline: 1.0, bump_or_trough: bump, center: 0.0, rise: 25.0, slope: 0.0, bump_height: 0.0, width: 100000.0}}, space: {{baseline: 1.0, bump_or_trough: bump, center: 0.0, rise: 25.0, slope: 0.0, bump_height: 0.0, width: 100000.0}}}}
""")

def run(time, nsteps=10):
    sim = Vlasov1D(deck(time))
    for _ in range(3): sim.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nsteps): sim.step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nsteps
    lib = _lib.load(); lib.adept_b200_profile(1)
    for _ in range(3): sim.step()
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16); lib.adept_b200_profile_report(buf, len(buf)); lib.adept_b200_profile(0)
    kern = {}
    for ln in buf.value.decode().splitlines():
        nm, cnt, tot = ln.split(); kern[nm] = {"n_per_step": int(cnt) / 3, "avg_us": float(tot) / int(cnt) * 1e3}
    finite = bool(torch.isfinite(sim.state["ion"]).all())
    return {"time": time, "ms_per_step": ms, "cell_updates_per_s": 17280 * 2048 / (ms * 1e-3), "finite": finite, "kernels": kern}

if __name__ == "__main__":
    for t in ("sixth", "leapfrog"):
        print(json.dumps(run(t)), flush=True)
