# Final round-2 evidence refresh: thread-per-matrix captures, the size sweep and the bench line of
# the final build (every ncu run follows a plain run of the same command).
set -x
bash tools/prof_tpm.sh
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_line.json 2> gpurun_out/bench_final.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu7.log 2>&1
python tools/quick_bench.py --cases se99:2:4194304,se99:3:4194304,se99:4:4194304,se99:6:2097152,se99:8:1048576,se99:10:1048576,se99:12:1048576,se99:16:1048576,se99:24:1048576,se99:32:1048576,se99:48:200000,se90:64:200000,se99:96:40000,se99:128:20000 --reps 3 > gpurun_out/r02_sweep_a.txt 2>&1
python tools/quick_bench.py --cases gmw81:4:4194304,gmw81:8:1048576,gmw81:16:1048576,gmw81:32:1048576,gmw81:64:200000,se90:4:4194304,se90:8:1048576,se90:16:1048576,se90:32:1048576,gmw81:128:20000,gmw81:200:4096,se99:200:4096,gmw81:256:4096,se99:256:4096,se90:256:4096 --reps 3 >> gpurun_out/r02_sweep_a.txt 2>&1
python tools/quick_bench.py --cases gmw81:384:2048,se99:384:2048,gmw81:512:1184,se99:512:1184,gmw81:1024:592,se99:1024:592 --reps 2 --families uniform >> gpurun_out/r02_sweep_a.txt 2>&1
ls -la gpurun_out | tail -5
