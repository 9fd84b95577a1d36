python -m pytest tests -m gpu -x -q -k "thread_per_matrix or small_kernel_families" 2>&1 | tail -3
echo "=== MODCHOL_TPM_MAX=8 (tpr)"
MODCHOL_TPM_MAX=8 python tools/quick_bench.py --cases se99:7:2097152,se99:8:1048576,gmw81:8:1048576,se90:8:1048576 --modes fast --families uniform,wishart,pd --reps 5
echo "=== MODCHOL_TPM_MAX=0 (warp kernels)"
MODCHOL_TPM_MAX=0 python tools/quick_bench.py --cases se99:7:2097152,se99:8:1048576,gmw81:8:1048576,se90:8:1048576 --modes fast,strict --families uniform,wishart,pd --reps 5
