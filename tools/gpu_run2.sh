#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r02b}
python tools/measure_fp32.py > gpurun_out/${T}_fp32_peaks.json 2> gpurun_out/${T}_fp32_peaks.err
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${T}_tests.log
cp gpurun_out/parity_errors.json gpurun_out/${T}_parity_errors.json 2>/dev/null
for w in mel gabor mfcc; do
  python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
done
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_features -s 4 -c 1 -o gpurun_out/${T}_ncu_mel -f \
     python bench.py --workload mel --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${T}_ncu_mel.log 2>&1
fi
tail -5 gpurun_out/${T}_tests.log
cat gpurun_out/${T}_fp32_peaks.json
for w in mel gabor mfcc; do python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_$w.json'));print('$w',round(d['ms_per_step']*1000,1),'us',d['value'])"; done
