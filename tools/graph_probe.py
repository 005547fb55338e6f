import sys, time
sys.path.insert(0, '.')
import torch, yaml
from pathlib import Path
from bench import c3_deck
from adept_b200.module import Vlasov1D
with open('tests/golden/epw.yaml') as fh: c1 = yaml.safe_load(fh)
def timeit(step, n, warm=10):
    for _ in range(warm): step()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)*1e3/n
for key, deck in (("C1", c1), ("C2", c3_deck(64,512)), ("1024x1024", c3_deck(1024,1024))):
    sim = Vlasov1D(deck); sim.t, sim.step_index = 30.0, 300
    us = timeit(sim.step, 300)
    g = 1e6/ sim.graph_steps_per_second(300)
    print(key, "eager us/step", round(us,1), "graph us/step", round(g,1), flush=True)
