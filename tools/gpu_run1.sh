#!/bin/bash
# GPU session 1 of round 2: regression + new parity cases on the round-1 kernel, compute-sanitizer, baselines, ncu.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_levels.py 2>&1 | tail -15 > gpurun_out/r02a_tests_old.log
python -m pytest tests/test_gpu_levels.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r02a_tests_levels.log
cp gpurun_out/parity_errors.json gpurun_out/r02a_parity_errors.json 2>/dev/null
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_cases.py > gpurun_out/r02a_sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r02a_sanitizer_$tool.log
done
for w in mel gabor mfcc; do
  python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > gpurun_out/r02a_bench_$w.json 2> gpurun_out/r02a_bench_$w.err
done
for w in gabor mfcc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_features -s 4 -c 1 -o gpurun_out/r02a_ncu_$w -f \
     python bench.py --workload $w --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r02a_ncu_$w.log 2>&1
done
ls -la gpurun_out
