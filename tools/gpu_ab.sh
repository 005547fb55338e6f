#!/bin/bash
# A/B of kernel variants selected by environment variables: for each setting (e.g. "ADEPT_B200_XVAR=5"), the parity
# tests that touch the two step kernels, then the step time.   bash tools/gpu_ab.sh r02i "A=1" "B=2 C=3"
R=$1; shift
O=gpurun_out
mkdir -p $O
i=0
for setting in "$@"; do
  i=$((i+1))
  env $setting timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_step.py -m gpu -q -x --timeout 600 -k "vdfdx or step or field or vpush or collide or save_moments" > $O/${R}_pytest_$i.log 2>&1; echo "[$setting] pytest exit $?"; tail -2 $O/${R}_pytest_$i.log
  env $setting timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras > $O/${R}_bench_$i.json 2> $O/${R}_bench_$i.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${R}_bench_$i.json").read().strip().splitlines()[-1])
    print("[$setting]", round(d["ms_per_step"] * 1e3, 1), {k: round(x["avg_us"], 1) for k, x in d["kernels"].items()}, "e2e", round(d["e2e"]["ms_per_step"] * 1e3, 1), round(d["e2e"]["without_default_save"]["ms_per_step"] * 1e3, 1))
except Exception as e:
    print("[$setting] bench failed", e); print(open("$O/${R}_bench_$i.err").read()[-2000:])
PY
done
python tools/kbench.py 4096 4096 10 2>&1 | grep -E "save_mom|vdfdx_rho|vpush_collide|edfdv_exp " 
