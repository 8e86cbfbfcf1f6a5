#!/bin/bash
# One GPU session for the section-8f entry points: parity tests, then timings at BASELINE config 3's shape.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 330 python -m pytest tests/test_gpu_auto.py -q -m gpu --maxfail=40 --tb=short -p no:cacheprovider > gpurun_out/auto_tests.log 2>&1
echo "pytest exit $?"; tail -n 60 gpurun_out/auto_tests.log | cut -c1-220
timeout 240 python tools/auto_bench.py --nbf 120 --nocc 20 --out gpurun_out/auto_bench_c3.json > gpurun_out/auto_bench.log 2>&1
echo "auto_bench exit $?"; tail -n 5 gpurun_out/auto_bench.log | cut -c1-3000
