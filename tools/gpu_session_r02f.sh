echo "== pytest gpu (1 GPU)"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== gemm shapes"; timeout 300 python tools/gemm_shapes_bench.py
echo "== sweep gemm list C3"; timeout 300 python tools/sweep_gemm_list.py 120 20 2>&1 | tail -62
echo "== bench N=1 (no large, no cpu)"; timeout 900 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench1c.err > gpurun_out/bench1c.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1c.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
print('e2e',d['e2e']['s_per_do_rccsd'],d['e2e']['phases_ms_rank0'])
print('steps',d['ms_each_step_rank0'])
print('host',d['host_issue_ms_rank0'])
print('roof',d['roofline']['shape'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline']['sweep_frac_of_peak'])
PY
