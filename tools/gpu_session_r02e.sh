echo "== pytest gpu (1 GPU)"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== sweep gemm list C3"; timeout 300 python tools/sweep_gemm_list.py 120 20 2>&1 | tail -75
echo "== bench N=1 (no large, no cpu)"; timeout 900 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench1b.err > gpurun_out/bench1b.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1b.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
print('e2e',d['e2e']['s_per_do_rccsd'],d['e2e']['phases_ms_rank0'])
print('steps',d['ms_each_step_rank0'])
print('host',d['host_issue_ms_rank0'])
PY
echo "== dist_check 2 ranks"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py 2>&1 | tail -3
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --no-large 2>gpurun_out/bench2b.err > gpurun_out/bench2b.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench2b.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'])
print('e2e',d['e2e']['s_per_do_rccsd'],d['e2e']['phases_ms_rank0'])
print('steps',d['ms_each_step_rank0'])
print('roof',d['roofline']['comm_ms_per_traced_sweep'])
PY
