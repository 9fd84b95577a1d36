python -m pytest tests -m gpu -x -q -k "thread_per_matrix or small_kernel_families or strict_bit_exact_vs_oracle or specials or ties" 2>&1 | tail -5
for t in 0 8; do
  echo "=== MODCHOL_TPM_MAX=$t"
  MODCHOL_TPM_MAX=$t python tools/quick_bench.py --cases se99:2:4194304,se99:3:4194304,se99:4:4194304,se99:5:2097152,se99:6:2097152,se99:8:1048576,gmw81:4:4194304,gmw81:8:1048576,se90:4:4194304 --modes strict,fast --families uniform,wishart --reps 5
done
