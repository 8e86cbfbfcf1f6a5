echo "== fused (T) experimental test"; JUES_B200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_auto.py -q -k fused 2>&1 | tail -3
echo "== (T) A/B at C3: two-kernel"; JUES_B200_TRACE=1 timeout 600 python tools/auto_bench.py --nbf 120 --nocc 20 --out gpurun_out/auto_plain.json 2>&1 | tail -2
echo "== (T) A/B at C3: fused"; JUES_B200_PT_FUSED=1 JUES_B200_TRACE=1 timeout 600 python tools/auto_bench.py --nbf 120 --nocc 20 --out gpurun_out/auto_fused.json 2>&1 | tail -2
echo "== bench N=1"; timeout 900 python bench.py 2>gpurun_out/bench1.err > gpurun_out/bench1.json; echo rc=$?; cut -c1-400 gpurun_out/bench1.json
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench2.err > gpurun_out/bench2.json; echo rc=$?; cut -c1-400 gpurun_out/bench2.json
