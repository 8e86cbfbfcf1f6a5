echo "== pytest gpu full (cc_energy as a dot product)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench N=1 quick"; timeout 600 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench_energy.err > gpurun_out/bench_energy.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_energy.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'],'e2e',d['e2e']['s_per_do_rccsd'])
PY
echo "== in-situ cc_energy"; timeout 300 python tools/sweep_gemm_list.py 120 20 2>&1 | grep "cc_energy\|^sweep" | tail -2; timeout 600 python tools/sweep_gemm_list.py 300 60 2>&1 | grep "cc_energy\|^sweep" | tail -2
