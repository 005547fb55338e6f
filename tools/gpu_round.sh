#!/bin/bash
# One GPU session of a round: tests, bench (both arms), per-kernel bench, ncu launch list + full capture.  Every step
# runs under its own timeout: a hung kernel must not hold the box until gpurun's limit.
#   gpurun --timeout 900 -- 'bash tools/gpu_round.sh r02v'
R=${1:-rXX}
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q --timeout 120 > $O/${R}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/${R}_pytest.log
timeout 200 python bench.py --steps 100 --warmup 5 > $O/${R}_bench.json 2> $O/${R}_bench.err; cat $O/${R}_bench.json | cut -c1-400
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2>> $O/${R}_bench.err
timeout 100 python tools/kbench.py > $O/${R}_kbench.txt 2>&1; cat $O/${R}_kbench.txt
timeout 100 python tools/bigbench_probe.py > $O/${R}_bigbench.txt 2>&1; tail -2 $O/${R}_bigbench.txt | cut -c1-200
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > $O/${R}_ncu_bench.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'vdfdx_tma|field_fused|vpush_collide|save_moments' -s 8 -c 4 -f \
    -o $O/${R}_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > $O/${R}_ncu_full.log 2>&1
ls -la $O | grep ${R}
