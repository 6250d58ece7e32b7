#!/bin/bash
# A/B timing of library builds on ONE box: LIBS="a.so b.so ..." (paths under the repo), interleaved, REPS rounds
set -u
mkdir -p gpurun_out
for rep in $(seq 1 ${REPS:-2}); do
  for lib in $LIBS; do
    for w in ${WORKLOADS:-mel gabor}; do
      AUD_B200_LIB=$PWD/$lib python bench.py --workload $w --utts 1024 --steps 100 --warmup 5 --kernel-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$lib $w rep$rep', round(d['ms_per_step']*1000,1), 'us', d['clocks']['sm_mhz'])"
    done
  done
done
