echo "== skinny gemm tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "skinny or test_gemm" 2>&1 | tail -5
echo "== pytest gpu full"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== sweep gemm list C3"; timeout 300 python tools/sweep_gemm_list.py 120 20 > gpurun_out/gemm_list_c3_r02i.txt 2>&1; tail -62 gpurun_out/gemm_list_c3_r02i.txt
echo "== bench N=1 quick"; timeout 900 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench1e.err > gpurun_out/bench1e.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1e.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
print('e2e',d['e2e']['s_per_do_rccsd'],d['e2e']['phases_ms_rank0'])
print('roof',d['roofline']['shape'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline']['sweep_frac_of_peak'],d['roofline']['all_gemms_share_of_sweep'])
PY
