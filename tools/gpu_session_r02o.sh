echo "== pytest gpu full"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== sweep list C3 (re-laid sweep)"; timeout 300 python tools/sweep_gemm_list.py 120 20 > gpurun_out/gemm_list_c3_r02o.txt 2>&1; tail -1 gpurun_out/gemm_list_c3_r02o.txt
echo "== bench N=1 quick"
timeout 900 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench1o.err > gpurun_out/bench1o.json; echo rc=$?; python - <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench1o.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','shape','ms_per_launch','share_of_sweep','sweep_frac_of_peak')})
print('e2e',d['e2e']['s_per_do_rccsd'],d['e2e']['phases_ms_rank0'])
PY
echo "== ncu full: element-wise kernels of a sweep"
JUES_B200_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:"amp_combos|residual_finish|cc_energy|ring_combine" -s 8 -c 8 -o gpurun_out/prof_aux_r02o python tools/sweep_for_ncu.py 120 20 4 > gpurun_out/ncu_aux_o.log 2>&1; tail -2 gpurun_out/ncu_aux_o.log
ncu -i gpurun_out/prof_aux_r02o.ncu-rep --page raw --csv > gpurun_out/ncu_aux_r02o_raw.csv 2>/dev/null; python tools/ncu_csv_to_json.py gpurun_out/ncu_aux_r02o_raw.csv gpurun_out/ncu_aux_kernels_r02o.json; rm -f gpurun_out/prof_aux_r02o.ncu-rep
