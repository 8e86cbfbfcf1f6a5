echo "== pytest gpu full"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== sweep list C3"; timeout 300 python tools/sweep_gemm_list.py 120 20 > gpurun_out/gemm_list_c3_r02p.txt 2>&1; tail -1 gpurun_out/gemm_list_c3_r02p.txt; grep "aux\|permute" gpurun_out/gemm_list_c3_r02p.txt
echo "== sweep list nbf=300 nocc=60"; timeout 600 python tools/sweep_gemm_list.py 300 60 > gpurun_out/gemm_list_n300_r02p.txt 2>&1; tail -1 gpurun_out/gemm_list_n300_r02p.txt; grep "aux\|permute" gpurun_out/gemm_list_n300_r02p.txt
echo "== bench N=1 quick"
timeout 900 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench1p.err > gpurun_out/bench1p.json; echo rc=$?; python - <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench1p.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
PY
