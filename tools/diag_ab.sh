for mode in P NP P NP P NP; do
  if [ $mode = NP ]; then export JUES_B200_GEMM_NONPERSISTENT=1; else unset JUES_B200_GEMM_NONPERSISTENT; fi
  echo "== mode $mode"
  DIAG_REPS=2 DIAG_SLABS=1 timeout 200 python tools/diag_transform_stress.py 2>&1 | cut -c1-400
done
