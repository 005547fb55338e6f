#!/bin/bash
# Sharded-grid session on N GPUs: parity tests, then the strong-scaling timing of the 4096^2 grid with the peer
# transfers as bulk copies / TMA stores (default) and as per-thread loads and stores (ADEPT_B200_PEER_TMA=0).
R=${1:-r02j}; N=${2:-2}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x --timeout 240 > $O/${R}_pytest_sharded.log 2>&1; echo "pytest exit $?"; tail -3 $O/${R}_pytest_sharded.log
for mode in ${3:-4 8 0}; do
  ADEPT_B200_SHARDED_CE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/bench_sharded.py 4096 4096 100 p2p > $O/${R}_sharded_${N}gpu_ce$mode.txt 2>&1; echo "CE=$mode:"; tail -2 $O/${R}_sharded_${N}gpu_ce$mode.txt | cut -c1-600
done
