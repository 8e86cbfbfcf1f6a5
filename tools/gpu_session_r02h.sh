echo "== test_multigpu (2 GPUs)"; timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -5
echo "== bench N=1 quick"; timeout 900 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench1d.err > gpurun_out/bench1d.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1d.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
print('e2e',d['e2e']['s_per_do_rccsd'],d['e2e']['phases_ms_rank0'])
print('roof',d['roofline']['shape'],d['roofline']['ms_per_launch'],d['roofline']['frac'],d['roofline']['sweep_frac_of_peak'],d['roofline']['all_gemms_share_of_sweep'])
PY
echo "== sweep gemm list C3"; timeout 300 python tools/sweep_gemm_list.py 120 20 2>&1 | tail -62 | awk '{print}' > gpurun_out/gemm_list_c3_r02h.txt; tail -3 gpurun_out/gemm_list_c3_r02h.txt
echo "== ncu full: non-GEMM kernels of a sweep"
JUES_B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none -k regex:"amp_combos|residual_finish|cc_energy|permute|pack_tau|unpack_ladder|splitk" -s 40 -c 30 -o gpurun_out/prof_aux_r02 python tools/sweep_for_ncu.py 120 20 3 > gpurun_out/ncu_aux.log 2>&1; tail -2 gpurun_out/ncu_aux.log
ncu -i gpurun_out/prof_aux_r02.ncu-rep --page raw --csv > gpurun_out/ncu_aux_r02_raw.csv 2>/dev/null; ls -la gpurun_out/prof_aux_r02.ncu-rep; rm -f gpurun_out/prof_aux_r02.ncu-rep
echo "== ncu full: GEMMs of a sweep"
JUES_B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_tma_dmma -s 92 -c 46 -o gpurun_out/prof_gemm_r02 python tools/sweep_for_ncu.py 120 20 3 > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
ncu -i gpurun_out/prof_gemm_r02.ncu-rep --page raw --csv > gpurun_out/ncu_gemm_r02_raw.csv 2>/dev/null; ls -la gpurun_out/prof_gemm_r02.ncu-rep; rm -f gpurun_out/prof_gemm_r02.ncu-rep
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --no-large 2>gpurun_out/bench2d.err > gpurun_out/bench2d.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench2d.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
print('e2e',d['e2e']['s_per_do_rccsd'],d['e2e']['phases_ms_rank0'])
PY
