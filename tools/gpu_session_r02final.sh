echo "== pytest gpu full"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench N=1 (no cpu leg)"; timeout 1200 python bench.py --no-cpu-baseline 2>gpurun_out/bench1final.err > gpurun_out/bench1final.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1final.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'],'e2e',d['e2e']['s_per_do_rccsd'])
print('large',{k:{q:v.get(q) for q in ('s_per_iteration','transform_s','frac_of_fp64_peak_per_gpu','transform_frac_of_fp64_peak_per_gpu','s_per_call','error')} for k,v in d['large'].items() if isinstance(v,dict)})
print('next_rows',{k:d['next_rows'].get(k) for k in ('s_per_call','ms_per_sweep_median','triples_ms','error')})
PY
