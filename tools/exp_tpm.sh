python -m pytest tests -m gpu -x -q -k "thread_per_matrix or small_kernel_families or eight_byte" 2>&1 | tail -4
echo "=== thread-per-matrix kernels, 8 < n <= 16 (MODCHOL_TPS_MAX=16)"
MODCHOL_TPS_MAX=16 python tools/quick_bench.py --cases se99:9:1048576,se99:10:1048576,se99:11:1048576,se99:12:1048576,se99:13:1048576,se99:14:1048576,se99:16:1048576,gmw81:9:1048576,gmw81:10:1048576,gmw81:12:1048576,gmw81:16:1048576,se90:12:1048576,se90:16:1048576 --modes fast --families uniform,wishart --reps 5
