#!/bin/bash
# One multi-GPU measurement session: tools/scale_session.sh N [big_nbf big_nocc]
N=$1; BIG_NBF=${2:-300}; BIG_NOCC=${3:-60}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29541 tools/dist_check.py 2>&1 | grep -E "DIST_PARITY|rror" | tee gpurun_out/dist_check_${N}gpu.txt
timeout 300 $TR --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print({k:d[k] for k in ('n_gpus','value','ms_per_step')}, d['config']['workload'][:40], d['e2e']['s_per_do_rccsd'], d['roofline']['sweep_frac_of_peak'], d['tei_transform']['cc_classes'])" || tail -5 gpurun_out/bench_n$N.err
timeout 400 $TR --master-port 29543 tools/big_run.py --nbf $BIG_NBF --nocc $BIG_NOCC --sweeps 2 --what rccsd --out gpurun_out/big_rccsd_${BIG_NBF}_${BIG_NOCC}_${N}gpu.json 2>&1 | tail -1 | cut -c1-900
if [ "$N" = "8" ]; then
  timeout 300 $TR --master-port 29544 tools/big_run.py --nbf 500 --nocc 60 --what rmp2 --out gpurun_out/big_rmp2_500_60_${N}gpu.json 2>&1 | tail -1 | cut -c1-700
fi
