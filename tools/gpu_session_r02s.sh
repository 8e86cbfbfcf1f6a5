echo "== pytest gpu full (amp extras)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== sweep list C3"; timeout 300 python tools/sweep_gemm_list.py 120 20 > gpurun_out/gemm_list_c3_r02s.txt 2>&1; tail -1 gpurun_out/gemm_list_c3_r02s.txt; grep "aux\|permute" gpurun_out/gemm_list_c3_r02s.txt
for mode in extras noextras; do
echo "== bench N=1 quick ($mode)"
if [ $mode = noextras ]; then export JUES_B200_NO_AMP_EXTRAS=1; else unset JUES_B200_NO_AMP_EXTRAS; fi
timeout 900 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench1s_$mode.err > gpurun_out/bench1s_$mode.json; echo rc=$?; python - $mode <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/bench1s_{sys.argv[1]}.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
PY
done
unset JUES_B200_NO_AMP_EXTRAS
echo "== diag next_rows (fresh)"; timeout 600 python tools/diag_next_rows.py 2>&1 | tail -3
echo "== diag next_rows (after rccsd runs)"; timeout 600 python tools/diag_next_rows.py rccsd_first 2>&1 | tail -3
