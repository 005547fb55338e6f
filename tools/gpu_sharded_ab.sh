#!/bin/bash
# Strong-scaling timing of the sharded 4096^2 grid on N GPUs for a list of environment settings.
#   bash tools/gpu_sharded_ab.sh r02q 4 "ADEPT_B200_PEER_TMA=1" "ADEPT_B200_PEER_TMA=0"
R=$1; N=$2; shift; shift
O=gpurun_out
mkdir -p $O
i=0
for setting in "$@"; do
  i=$((i+1))
  env $setting timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$i \
    tools/bench_sharded.py 4096 4096 100 p2p > $O/${R}_sharded_${N}gpu_$i.txt 2>&1; echo "[$setting]"; tail -1 $O/${R}_sharded_${N}gpu_$i.txt | cut -c150-600
done
