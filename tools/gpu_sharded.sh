#!/bin/bash
# Sharded-grid session on N GPUs: parity tests, then bench.py under torchrun (its extras carry the sharded grid's parity
# against the oracle and its strong-scaling time).
R=${1:-rXX}; N=${2:-2}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 600 > $O/${R}_pytest_sharded.log 2>&1; echo "pytest exit $?"; tail -3 $O/${R}_pytest_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/${R}_bench_${N}gpu.json 2> $O/${R}_bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.loads(open("$O/${R}_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"])
    print(json.dumps(d["extra"].get("sharded"))[:1200])
    print(json.dumps(d["extra"].get("ensemble_c4"))[:500])
except Exception as e:
    print("bench failed", e); print(open("$O/${R}_bench_${N}gpu.err").read()[-3000:])
PY
