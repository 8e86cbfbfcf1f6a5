echo "== pytest gpu full (2-GPU box)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== dist_check 2 ranks"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py 2>&1 | tail -8
echo "== bench N=2"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 2>gpurun_out/bench2q.err > gpurun_out/bench2q.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench2q.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity'])
print('e2e',d['e2e']['s_per_do_rccsd'])
print('large',json.dumps(d['large'])[:2500])
PY
tail -5 gpurun_out/bench2q.err
