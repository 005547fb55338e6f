#!/bin/bash
# Quick A/B of the x-push variants: parity tests touching the x-push, per-kernel bench, short bench in both modes.
R=${1:-r02a}
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q -x --timeout 600 -k "vdfdx or step or field" > $O/${R}_pytest_x.log 2>&1; echo "pytest exit $?"; tail -5 $O/${R}_pytest_x.log
for mode in tma dual; do
  ADEPT_B200_XPUSH=$mode python tools/kbench.py 4096 4096 10 > $O/${R}_kbench_$mode.txt 2>&1; grep -E "vdfdx|error|Error" $O/${R}_kbench_$mode.txt
  ADEPT_B200_XPUSH=$mode python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > $O/${R}_bench_$mode.json 2> $O/${R}_bench_$mode.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${R}_bench_$mode.json").read().strip().splitlines()[-1])
    print("$mode", d["ms_per_step"], d["kernels"], d["e2e"]["ms_per_step"])
except Exception as e:
    print("$mode bench failed", e); print(open("$O/${R}_bench_$mode.err").read()[-2000:])
PY
done
