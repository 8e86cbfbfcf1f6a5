echo "== pytest gpu full (final binary)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== sweep list nbf=300 nocc=60 (final)"; timeout 600 python tools/sweep_gemm_list.py 300 60 > gpurun_out/gemm_list_n300_r02last.txt 2>&1; tail -1 gpurun_out/gemm_list_n300_r02last.txt; grep "aux\|permute" gpurun_out/gemm_list_n300_r02last.txt
echo "== sweep list C3 (final)"; timeout 300 python tools/sweep_gemm_list.py 120 20 > gpurun_out/gemm_list_c3_r02last.txt 2>&1; tail -1 gpurun_out/gemm_list_c3_r02last.txt
