echo "== pytest gpu full"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== sweep list C3 (packed permutes)"; timeout 300 python tools/sweep_gemm_list.py 120 20 > gpurun_out/gemm_list_c3_r02l.txt 2>&1; grep "permute\|^sweep" gpurun_out/gemm_list_c3_r02l.txt
echo "== bench N=1 quick"; timeout 900 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench1i.err > gpurun_out/bench1i.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1i.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
print('e2e',d['e2e']['s_per_do_rccsd'],d['e2e']['phases_ms_rank0'])
PY
echo "== bench N=1 quick (old permutes)"; JUES_B200_NO_PACKED_PERMUTE=1 timeout 900 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench1j.err > gpurun_out/bench1j.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1j.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
PY
