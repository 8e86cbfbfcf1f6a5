#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_auto.py tests/test_golden.py -q -m gpu -k "rccsd or rccd or auto or pT or triples or golden or mrccd" --tb=short -p no:cacheprovider 2>&1 | tail -n 15 | cut -c1-200
for n in 1 2 3; do
for cfg in "JUES_B200_NO_GRAPH=1 --steps 60" "X=1 --steps 60"; do
  envs=${cfg%% *}; fl=${cfg#* }
  env $envs BENCH_SAMPLER_PERIOD=0.1 timeout 200 python bench.py --no-cpu-baseline $fl > gpurun_out/b.json 2> gpurun_out/b.err
  python - "$cfg" <<'PY'
import json,sys
d=json.load(open('gpurun_out/b.json')); s=d['ms_each_step_rank0']; h=d['host_issue_ms_rank0']
print(sys.argv[1], 'ms_per_step', round(d['ms_per_step'],3), 'max', max(s), 'n>10ms', sum(1 for x in s if x>10), 'graph', d['graph_replayed_sweeps'], 'host_issue_last_ms', h[-1], 'gpu_total_ms', round(sum(s),1), 'e2e', round(d['e2e']['s_per_do_rccsd'],4))
PY
done; done
cp gpurun_out/b.json gpurun_out/bench_graph.json
JUES_B200_TRACE=1 timeout 240 python tools/auto_bench.py --nbf 120 --nocc 20 --out gpurun_out/auto_bench_c3_trace.json > gpurun_out/auto_bench.log 2>&1
timeout 240 python tools/auto_bench.py --nbf 120 --nocc 20 --out gpurun_out/auto_bench_c3.json > gpurun_out/auto_bench.log 2>&1
python - <<'PY'
import json
for f in ('auto_bench_c3_trace','auto_bench_c3'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    a=d['auto_rccsd_canonical']; print(f,{k:a[k] for k in ('iterations','ept','cc.triples_ms','pt_tflops','pt_frac_of_fp64_peak','ms_per_sweep_median','wall_s')}, {k:v for k,v in a.get('trace_ms_sum',{}).items() if k.startswith('pt.')}, d['auto_rccsd_noncanonical']['ms_per_sweep_median'], d['auto_rccsd_noncanonical']['wall_s'], d['mrccd_diis']['wall_s'])
PY
