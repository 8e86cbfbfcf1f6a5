echo "== cold transform stress (new lib)"; DIAG_REPS=3 timeout 200 python tools/diag_transform_stress.py 2>&1 | cut -c1-300
echo "== gemm shapes NEW"; timeout 300 python tools/gemm_shapes_bench.py
echo "== gemm shapes OLD"; JUES_B200_LIB=$PWD/tools/_libjues_old.so timeout 300 python tools/gemm_shapes_bench.py
echo "== gemm shapes NEW again"; timeout 300 python tools/gemm_shapes_bench.py
echo "== slab diag"; timeout 300 python tools/diag_slab_transform.py
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
