"""Strong-scaling timing of ONE 4096 x 4096 grid sharded over the ranks (development aid; bench.py is the contract).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_sharded.py [nx nv steps]
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
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sim = ShardedVlasov1D(c3_deck(nx, nv))
sim.t, sim.step_index = 30.0, 300
for _ in range(5):
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
tm = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
dist.all_reduce(tm, op=dist.ReduceOp.MAX)
if rank == 0:
    el = float(tm.item())
    print(json.dumps({"mode": "single grid, v-sharded (all-to-all x2 + all-reduce per step)", "n_gpus": world, "nx": nx,
                      "nv": nv, "steps": K, "ms_per_step": el / K * 1e3, "cell_updates_per_s": nx * nv * K / el}))
dist.destroy_process_group()
