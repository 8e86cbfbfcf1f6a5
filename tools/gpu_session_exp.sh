#!/bin/bash
mkdir -p gpurun_out
timeout 120 ncu --set full --clock-control none --import-source on -k regex:dgemm_tma_dmma -s 2 -c 1 -f -o gpurun_out/ncu_gemm_ring python tools/gemm_one.py T N 2000 2000 2000 > gpurun_out/ncu_gemm_ring.log 2>&1
echo "ring exit $?"; tail -n 2 gpurun_out/ncu_gemm_ring.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_launches_sweep.csv python tools/sweep_for_ncu.py 120 20 3 > gpurun_out/ncu_sweep.log 2>&1
echo "list exit $?"; wc -l gpurun_out/ncu_launches_sweep.csv
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:dgemm_tma_dmma<0, 1, (80|96|112|128), 128' -s 1 -c 1 -f -o gpurun_out/ncu_gemm_saladder python tools/sweep_for_ncu.py 120 20 2 > gpurun_out/ncu_gemm_saladder.log 2>&1
echo "saladder exit $?"; tail -n 3 gpurun_out/ncu_gemm_saladder.log | cut -c1-200
for k in pack_tau_sa_kernel unpack_ladder_sa_kernel; do
timeout 120 ncu --set full --clock-control none -k regex:$k -s 1 -c 1 -f -o gpurun_out/ncu_$k python tools/sweep_for_ncu.py 120 20 2 > gpurun_out/ncu_$k.log 2>&1
echo "$k exit $?"
done
ls -la gpurun_out/*.ncu-rep
