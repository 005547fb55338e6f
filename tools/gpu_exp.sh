#!/bin/bash
# One-off experiment session (1 GPU): parity of the touched kernels, kernel timings, ncu capture of save_moments.
R=${1:-r02t}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q -x --timeout 600 -k "vpush or collide or save_moments or step or edfdv" 2>&1 | tail -2
python tools/kbench.py 4096 4096 10 2>&1 | grep -E "edfdv_exp |save_mom|vpush_collide|collide_"
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras > $O/${R}_bench.json 2> $O/${R}_bench.err
python - <<PY
import json
d = json.loads(open("$O/${R}_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"] * 1e3, 1), {k: round(x["avg_us"], 1) for k, x in d["kernels"].items()}, "e2e", round(d["e2e"]["ms_per_step"] * 1e3, 1), round(d["e2e"]["without_default_save"]["ms_per_step"] * 1e3, 1))
PY
ncu --set full --clock-control none --import-source on -k regex:'save_moments' -c 2 -f -o $O/${R}_savemom python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > $O/${R}_ncu_savemom.log 2>&1
ls -la $O | grep ${R}
