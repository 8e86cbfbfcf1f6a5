echo "== pytest gpu full (narrow-arithmetic generator)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench N=1 with large"; timeout 900 python bench.py --no-cpu-baseline --no-next-rows 2>gpurun_out/bench_gen.err > gpurun_out/bench_gen.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_gen.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'parity',d['parity']['status'],d['parity']['max_abs_dE'],d['parity'].get('large_invariance',{}).get('max_abs_dE'))
for k,v in d['large'].items():
    if isinstance(v,dict): print(k,{q:v.get(q) for q in ('s_per_iteration','s_per_call','transform_s','ao_generation_ms','transform_frac_of_fp64_peak_per_gpu','error')})
PY
