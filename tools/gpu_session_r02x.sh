echo "== pytest gpu full"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== full default bench"; timeout 2400 python bench.py 2>gpurun_out/bench1x.err > gpurun_out/bench1x.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1x.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'],d['parity'].get('large_invariance'))
print('e2e',d['e2e']['s_per_do_rccsd'], 'launches', d['gpu_launches'])
print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','ms_per_launch','share_of_sweep','sweep_frac_of_peak','isolated_tflops')})
print('cpu',{k:d['cpu_baseline'][k] for k in ('value','cores','s_per_iteration','s_transforms')})
nr=d['next_rows']; print('next_rows',{k:nr.get(k) for k in ('s_per_call','iterations','ms_per_sweep_median','triples_ms','error')})
print('tei',d['tei_transform'])
print('large',json.dumps(d['large'])[:2600])
print('clocks',d['clocks'])
PY
tail -3 gpurun_out/bench1x.err
