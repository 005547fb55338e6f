"""Throughput of an ensemble of independent 64 x 512 runs on one GPU (BASELINE.json configs[3]: 1024 members over 8
GPUs = 128 members per GPU).  Development aid; bench.py is the contract.

    python tools/bench_ensemble.py [members nx nv steps]
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ctypes  # noqa: E402

from adept_b200 import _lib  # noqa: E402
from adept_b200.ensemble import EnsembleVlasov1D  # noqa: E402
from bench import c3_deck  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nv = int(sys.argv[3]) if len(sys.argv) > 3 else 512
K = int(sys.argv[4]) if len(sys.argv) > 4 else 100
decks = []
for k0 in np.linspace(0.2, 0.4, int(np.sqrt(B))):
    for a0 in np.logspace(-4, -1, B // int(np.sqrt(B))):
        d = c3_deck(nx, nv)
        d["grid"]["xmax"] = 2 * np.pi / k0
        d["density"]["species-background"]["wavenumber"] = float(k0)
        d["drivers"]["ex"]["0"]["params"].update(k0=float(k0), a0=float(a0), w0=float(np.sqrt(1 + 3 * k0**2)))
        decks.append(d)
# under torchrun (one rank per GPU) the members are split across the ranks: no data-path collective, the only
# communication is the barrier around the timed region and the max over ranks of the device time
import os  # noqa: E402

import torch.distributed as dist  # noqa: E402

from adept_b200.ensemble import member_slice  # noqa: E402

world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
if world > 1:
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
n_total = len(decks)
decks = decks[member_slice(n_total, rank, world)]
ens = EnsembleVlasov1D(decks)
ens.t, ens.step_index = 30.0, 300
for _ in range(5):
    ens.step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    ens.step()
e1.record()
torch.cuda.synchronize()
el = e0.elapsed_time(e1) * 1e-3
if world > 1:
    t = torch.tensor([el], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    el = float(t.item())
lib = _lib.load()
lib.adept_b200_profile(1)
for _ in range(K):
    ens.step()
torch.cuda.synchronize()
buf = ctypes.create_string_buffer(1 << 16)
lib.adept_b200_profile_report(buf, len(buf))
kern = {l.split()[0]: round(float(l.split()[2]) / int(l.split()[1]) * 1e3, 1) for l in buf.value.decode().splitlines()}
cells = n_total * nx * nv
if rank == 0:
    print(json.dumps({"members": n_total, "n_gpus": world, "nx": nx, "nv": nv, "us_per_step": el / K * 1e6,
                      "cell_updates_per_s": cells * K / el,
                      "frac_of_48B_roofline": 48 * cells * K / el / (6548.5e9 * world), "kernel_us": kern}))
if world > 1:
    dist.destroy_process_group()
