python -m pytest tests -m gpu -x -q -k "thread_per_matrix or small_kernel_families or eight_byte" 2>&1 | tail -4
echo "=== direct kernels with 256-bit loads"
MODCHOL_TPS_MAX=16 python tools/quick_bench.py --cases se99:8:1048576,gmw81:8:1048576,se99:12:1048576,se99:16:1048576 --modes fast --families uniform,wishart --reps 5
