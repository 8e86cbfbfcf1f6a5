echo "== pytest gpu full"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== auto bench C3 (get_fock / AutoRCCSD / (T) / mRCCD)"; timeout 900 python tools/auto_bench.py --nbf 120 --nocc 20 > gpurun_out/auto_bench_c3_r02r.json 2>gpurun_out/auto_bench_r02r.err; tail -c 1500 gpurun_out/auto_bench_c3_r02r.json
echo "== full default bench"; timeout 2400 python bench.py 2>gpurun_out/bench1r.err > gpurun_out/bench1r.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1r.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
print('e2e',d['e2e'])
print('cpu',d['cpu_baseline'])
print('next_rows',d['next_rows'])
print('tei',d['tei_transform'])
print('clocks',d['clocks'])
PY
echo "== reference arm"; timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-1200
