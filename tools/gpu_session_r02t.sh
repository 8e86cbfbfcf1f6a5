echo "== pytest gpu full (side branch)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== determinism stress"; timeout 900 python tools/diag_overlap_stress.py 12 2>&1 | tail -12
for mode in overlap nooverlap; do
echo "== bench N=1 quick ($mode)"
if [ $mode = nooverlap ]; then export JUES_B200_NO_OVERLAP=1; else unset JUES_B200_NO_OVERLAP; fi
timeout 900 python bench.py --no-large --no-cpu-baseline --no-next-rows 2>gpurun_out/bench1t_$mode.err > gpurun_out/bench1t_$mode.json; echo rc=$?; python - $mode <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/bench1t_{sys.argv[1]}.json').read())
print('value',d['value'],'ms',d['ms_per_step'],'graph',d['graph_replayed_sweeps'],'parity',d['parity']['status'],d['parity']['max_abs_dE'], 'e2e', d['e2e']['s_per_do_rccsd'])
PY
done
unset JUES_B200_NO_OVERLAP
echo "== auto bench C3"; timeout 900 python tools/auto_bench.py --nbf 120 --nocc 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k in ('auto_rccsd_canonical','auto_rccsd_noncanonical','mrccd_diis'):
    print(k, {q:d[k].get(q) for q in ('wall_s','iterations','ms_per_sweep_median','ecc')})
"
