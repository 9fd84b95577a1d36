python -m pytest tests -m gpu -x -q -k "thread_per_matrix or small_kernel_families or eight_byte or fused" 2>&1 | tail -3
for v in "" d4; do
  if [ -z "$v" ]; then unset MODCHOL_LIB; else export MODCHOL_LIB=$PWD/modcholesky_b200/csrc/libmodchol_b200.$v.so; fi
  echo "=== variant ${v:-base}"
  python tools/quick_bench.py --cases se99:4:4194304,gmw81:4:4194304,se90:4:4194304,se99:8:1048576,gmw81:8:1048576 --modes fast --families uniform,wishart --reps 7
done
MODCHOL_LIB=$PWD/modcholesky_b200/csrc/libmodchol_b200.d4.so python -m pytest tests -m gpu -x -q -k "thread_per_matrix or eight_byte" 2>&1 | tail -2
