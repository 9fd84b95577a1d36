# thread-per-matrix (default) against the warp kernels (MODCHOL_TPM_MAX=0 / MODCHOL_TPS_MAX=0) at small batches
for b in 512 2048 8192; do
for t in 8 0; do
  echo "=== batch $b MODCHOL_TPM_MAX=$t"
  MODCHOL_TPM_MAX=$t python tools/quick_bench.py --cases se99:4:$b,se99:6:$b,se99:8:$b --modes strict,fast --families uniform --reps 9 | awk '{print $1,$2,$3,$5,$6,$7,$8,$9,$10}'
done; done
