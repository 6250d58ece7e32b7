#!/bin/bash
# Round-2 measurement set on one box: tests, bench lines, launch list, ncu --set full of the headline kernel.
set -u
mkdir -p gpurun_out
T=${TAG:-r02}
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}_tests.log
cp gpurun_out/parity_errors.json gpurun_out/${T}_parity.json 2>/dev/null
python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "default exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
python bench.py --workload mel > gpurun_out/${T}_bench_mel.json 2> gpurun_out/${T}_bench_mel.err
python bench.py --workload mfcc > gpurun_out/${T}_bench_mfcc.json 2> gpurun_out/${T}_bench_mfcc.err
# launch list of the default command (after it exited 0 without ncu)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/${T}_ncu_launches.log 2>&1
# ncu --set full of one at-size launch of the headline kernel (mel + gabor, ~7,100 utterances)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_features -s 12 -c 1 -o gpurun_out/${T}_ncu_gabor_atsize -f \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/${T}_ncu_gabor_atsize.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_features -s 4 -c 1 -o gpurun_out/${T}_ncu_mel -f \
   python bench.py --workload mel --utts 1024 --steps 2 --warmup 3 --kernel-only > gpurun_out/${T}_ncu_mel.log 2>&1
cat gpurun_out/${T}_tests.log
ls -la gpurun_out | grep ${T}_
