echo "== skinny gemm tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "skinny or test_gemm or rccsd_every or rccd_every" 2>&1 | tail -3
echo "== sweep gemm list C3 (skip on)"; timeout 300 python tools/sweep_gemm_list.py 120 20 > gpurun_out/gemm_list_c3_r02j.txt 2>&1; grep -v "^ .*permute" gpurun_out/gemm_list_c3_r02j.txt | tail -52
echo "== sweep gemm list C3 (skip off)"; JUES_B200_LIB=$PWD/tools/_libjues_old.so timeout 300 python tools/sweep_gemm_list.py 120 20 > gpurun_out/gemm_list_c3_r02j_noskip.txt 2>&1; grep -v "^ .*permute" gpurun_out/gemm_list_c3_r02j_noskip.txt | tail -52
echo "== bench N=1 quick (skip on)"; timeout 900 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench1f.err > gpurun_out/bench1f.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1f.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
print('e2e',d['e2e']['s_per_do_rccsd'],d['e2e']['phases_ms_rank0'])
PY
echo "== bench N=1 quick (skip off)"; JUES_B200_LIB=$PWD/tools/_libjues_old.so timeout 900 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench1g.err > gpurun_out/bench1g.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1g.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
PY
