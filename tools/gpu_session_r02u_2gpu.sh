echo "== pytest multigpu"; timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3
echo "== dist_check 2 ranks"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py 2>&1 | grep -E "DIST_PARITY|errs|Error|error" | tail -8
for mode in overlap; do
if [ $mode = nooverlap ]; then export JUES_B200_NO_OVERLAP=1; else unset JUES_B200_NO_OVERLAP; fi
echo "== bench N=2 ($mode)"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 2>gpurun_out/bench2u_$mode.err > gpurun_out/bench2u_$mode.json; echo rc=$?; python - $mode <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/bench2u_{sys.argv[1]}.json').read())
print('large',json.dumps(d.get('large'))[:1500]); print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'],'e2e',d['e2e']['s_per_do_rccsd'], 'comm', d['roofline'].get('comm_ms_per_traced_sweep'))
PY
done
