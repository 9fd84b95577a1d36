python -m pytest tests -m gpu -x -q -k "thread_per_matrix or small_kernel_families" 2>&1 | tail -4
echo "=== tpr kernels"
python tools/quick_bench.py --cases se99:2:4194304,se99:3:4194304,se99:4:4194304,se99:5:2097152,se99:6:2097152,se99:7:2097152,se99:8:1048576,se90:4:4194304,se90:6:2097152,se90:8:1048576 --modes fast --families uniform,wishart --reps 5
