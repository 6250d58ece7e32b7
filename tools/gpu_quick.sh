#!/bin/bash
# quick GPU check: a parity subset + the three bench workloads (kernel-only numbers)
set -u
mkdir -p gpurun_out
T=${TAG:-quick}
python -m pytest tests/test_gpu_parity.py tests/test_gpu_levels.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${T}_tests.log
for w in ${WORKLOADS:-mel gabor mfcc}; do
  python bench.py --workload $w --utts 1024 --steps 50 --warmup 5 --kernel-only > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
done
tail -3 gpurun_out/${T}_tests.log
for w in ${WORKLOADS:-mel gabor mfcc}; do python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_$w.json'));print('$w',round(d['ms_per_step']*1000,1),'us',d['value'])"; done
if [ -n "${EXTRA_OPTS:-}" ]; then
  for w in ${WORKLOADS:-mel gabor mfcc}; do
    python bench.py --workload $w --utts 1024 --steps 50 --warmup 5 --kernel-only $EXTRA_OPTS > gpurun_out/${T}_bench_${w}_opt.json 2> gpurun_out/${T}_bench_${w}_opt.err
    python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_${w}_opt.json'));print('$w [$EXTRA_OPTS]',round(d['ms_per_step']*1000,1),'us',d['value'])"
  done
fi
