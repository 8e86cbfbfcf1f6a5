echo "== dist_check 8 ranks"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py 2>&1 | grep -E "DIST_PARITY|errs|Error|error" | tail -8
echo "== bench N=8"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 2>gpurun_out/bench8v.err > gpurun_out/bench8v.json; echo rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench8v.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity'])
print('e2e',d['e2e']['s_per_do_rccsd'], 'comm', d['roofline'].get('comm_ms_per_traced_sweep'))
print('large',json.dumps(d['large'])[:4000])
PY
tail -3 gpurun_out/bench8v.err
