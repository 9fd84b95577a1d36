python -m pytest tests -m gpu -x -q -k "thread_per_matrix" 2>&1 | tail -6
echo "=== GMW81 thread kernels (MODCHOL_TPS_MAX=16)"
MODCHOL_TPS_MAX=16 python tools/quick_bench.py --cases gmw81:9:1048576,gmw81:10:1048576,gmw81:12:1048576,gmw81:14:1048576,gmw81:16:1048576 --modes fast --families uniform --reps 5
echo "=== GMW81 warp kernels (MODCHOL_TPS_MAX=0)"
MODCHOL_TPS_MAX=0 python tools/quick_bench.py --cases gmw81:9:1048576,gmw81:10:1048576,gmw81:12:1048576,gmw81:14:1048576,gmw81:16:1048576 --modes fast,strict --families uniform --reps 5
