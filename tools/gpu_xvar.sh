#!/bin/bash
# A/B of the nx = 4096 x-push variants (ADEPT_B200_XVAR): parity tests on the x-push paths, then per-variant step time.
R=${1:-r02h}
VARS=${2:-"1 4"}
O=gpurun_out
mkdir -p $O
for v in $VARS; do
  ADEPT_B200_XVAR=$v timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q -x --timeout 600 -k "vdfdx or step or field" > $O/${R}_pytest_x$v.log 2>&1; echo "var $v pytest exit $?"; tail -3 $O/${R}_pytest_x$v.log
  ADEPT_B200_XVAR=$v timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras > $O/${R}_bench_x$v.json 2> $O/${R}_bench_x$v.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${R}_bench_x$v.json").read().strip().splitlines()[-1])
    print("var $v", d["ms_per_step"], {k: round(x["avg_us"], 1) for k, x in d["kernels"].items()}, d["e2e"]["ms_per_step"])
except Exception as e:
    print("var $v bench failed", e); print(open("$O/${R}_bench_x$v.err").read()[-2000:])
PY
done
