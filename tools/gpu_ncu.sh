#!/bin/bash
# ncu --set full capture of the fused kernel for the given workloads (after a plain run of the same command exited 0)
set -u
mkdir -p gpurun_out
T=${TAG:-ncu}
for w in ${WORKLOADS:-gabor}; do
  python bench.py --workload $w --utts 1024 --steps 2 --warmup 3 --kernel-only > gpurun_out/${T}_plain_$w.log 2>&1 || { echo "plain run failed for $w"; continue; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_features -s 4 -c 1 -o gpurun_out/${T}_ncu_$w -f \
     python bench.py --workload $w --utts 1024 --steps 2 --warmup 3 --kernel-only > gpurun_out/${T}_ncu_$w.log 2>&1
done
ls -la gpurun_out | grep ${T}
