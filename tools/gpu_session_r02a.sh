echo "== pytest gpu (1 GPU)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench N=1 quick"; timeout 600 python bench.py --steps 10 --warmup 3 --no-next-rows 2>gpurun_out/bench1.err | tee gpurun_out/bench1.json | cut -c1-1500
