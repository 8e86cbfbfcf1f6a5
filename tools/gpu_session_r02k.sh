echo "== pytest gpu full"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== sweep gemm list C3"; timeout 300 python tools/sweep_gemm_list.py 120 20 > gpurun_out/gemm_list_c3_r02k.txt 2>&1; head -10 gpurun_out/gemm_list_c3_r02k.txt; tail -3 gpurun_out/gemm_list_c3_r02k.txt
echo "== bench N=1 full"; timeout 1200 python bench.py 2>gpurun_out/bench1h.err > gpurun_out/bench1h.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1h.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
print('e2e',d['e2e']['s_per_do_rccsd'],d['e2e']['value'],d['e2e']['phases_ms_rank0'])
print('roof',json.dumps(d['roofline'])[:700])
print('tei',json.dumps(d['tei_transform'])[:600])
print('large',json.dumps(d['large'])[:900])
print('cpu',json.dumps(d['cpu_baseline'])[:500])
print('next',json.dumps(d['next_rows'])[:600])
print('clocks',d['clocks'])
PY
