echo "== pytest gpu full"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== sweep list C3 (re-laid sweep)"; timeout 300 python tools/sweep_gemm_list.py 120 20 > gpurun_out/gemm_list_c3_r02n.txt 2>&1; tail -1 gpurun_out/gemm_list_c3_r02n.txt
echo "== sweep list C3 (plain sweep)"; JUES_B200_PLAIN_SWEEP=1 timeout 300 python tools/sweep_gemm_list.py 120 20 > gpurun_out/gemm_list_c3_r02n_plain.txt 2>&1; tail -1 gpurun_out/gemm_list_c3_r02n_plain.txt
for mode in relaid plain; do
echo "== bench N=1 quick ($mode)"
if [ $mode = plain ]; then export JUES_B200_PLAIN_SWEEP=1; else unset JUES_B200_PLAIN_SWEEP; fi
timeout 900 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench1n_$mode.err > gpurun_out/bench1n_$mode.json; echo rc=$?; python - $mode <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/bench1n_{sys.argv[1]}.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','shape','ms_per_launch','share_of_sweep','sweep_frac_of_peak')})
PY
done
unset JUES_B200_PLAIN_SWEEP
echo "== gemm autotune (skinny shapes, 10 configurations)"; timeout 600 python tools/gemm_autotune.py skinny > gpurun_out/gemm_autotune_r02n.log 2>&1; cut -c1-220 gpurun_out/gemm_autotune_r02n.log
