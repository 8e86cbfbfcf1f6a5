#!/bin/bash
# Full GPU session: every -m gpu parity test, the bench line, the section-8f timings (with the
# fine-grained trace timers on for the (T) breakdown).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests -q -m gpu --maxfail=40 --tb=short -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "pytest exit $?"; tail -n 40 gpurun_out/gpu_tests.log | cut -c1-220
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; cut -c1-2500 gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err | cut -c1-300
JUES_B200_TRACE=1 timeout 240 python tools/auto_bench.py --nbf 120 --nocc 20 --out gpurun_out/auto_bench_c3_trace.json > gpurun_out/auto_bench.log 2>&1
echo "auto_bench exit $?"; tail -n 3 gpurun_out/auto_bench.log | cut -c1-3500
