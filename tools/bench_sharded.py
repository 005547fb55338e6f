"""Strong-scaling timing of ONE 4096 x 4096 grid sharded over the ranks (development aid; bench.py is the contract).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_sharded.py [nx nv steps [nccl|p2p|auto]]
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import c3_deck  # noqa: E402

from adept_b200.sharded import ShardedVlasov1D  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = int(sys.argv[3]) if len(sys.argv) > 3 else 30
transpose = sys.argv[4] if len(sys.argv) > 4 else "auto"  # nccl | p2p | auto
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sim = ShardedVlasov1D(c3_deck(nx, nv), transpose=transpose)
sim.t, sim.step_index = 30.0, 300
for _ in range(30):  # NCCL and the symmetric-memory rendezvous keep initialising lazily for tens of steps at 4+ ranks
    sim.step()
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    sim.step()
e1.record()
dist.barrier()
torch.cuda.synchronize()
# second pass with the library's per-kernel event timing (rank 0 reports its own kernels)
import ctypes  # noqa: E402

from adept_b200 import _lib  # noqa: E402

lib = _lib.load()
lib.adept_b200_profile(1)
for _ in range(K):
    sim.step()
torch.cuda.synchronize()
buf = ctypes.create_string_buffer(1 << 16)
lib.adept_b200_profile_report(buf, len(buf))
kernels = {}
for line in buf.value.decode().splitlines():
    name, count, ms = line.split()
    kernels[name] = round(float(ms) / int(count) * 1e3, 1)
lib.adept_b200_profile(0)
tm = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
dist.all_reduce(tm, op=dist.ReduceOp.MAX)
if rank == 0:
    el = float(tm.item())
    mode = ("single grid, v-sharded, transposes fused into the kernels' stores over peer memory + all-reduce"
            if sim.p2p is not None else "single grid, v-sharded (all-to-all x2 + all-reduce per step)")
    print(json.dumps({"mode": mode, "n_gpus": world, "nx": nx,
                      "nv": nv, "steps": K, "ms_per_step": el / K * 1e3, "cell_updates_per_s": nx * nv * K / el, "rank0_kernel_us": kernels}))
dist.destroy_process_group()
