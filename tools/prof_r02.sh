set -x
N="ncu --set full --clock-control none --import-source on -c 1"
python tools/prof_case.py se99 32 131072 uniform fast > gpurun_out/plain1.log 2>&1 && $N -k regex:wip_kernel -s 1 -o gpurun_out/r02_se99_n32_fast python tools/prof_case.py se99 32 131072 uniform fast > gpurun_out/ncu1.log 2>&1
python tools/prof_case.py se99 32 131072 uniform strict > gpurun_out/plain2.log 2>&1 && $N -k regex:wsm_se -s 1 -o gpurun_out/r02_se99_n32_strict python tools/prof_case.py se99 32 131072 uniform strict > gpurun_out/ncu2.log 2>&1
python tools/prof_case.py gmw81 256 1184 uniform fast > gpurun_out/plain3.log 2>&1 && $N -k regex:bsm_gmw -s 1 -o gpurun_out/r02_gmw81_n256_fast python tools/prof_case.py gmw81 256 1184 uniform fast > gpurun_out/ncu3.log 2>&1
python tools/prof_case.py se99 256 1184 uniform fast > gpurun_out/plain4.log 2>&1 && $N -k regex:bsm_se -s 1 -o gpurun_out/r02_se99_n256_fast python tools/prof_case.py se99 256 1184 uniform fast > gpurun_out/ncu4.log 2>&1
python tools/prof_case.py se90 64 65536 uniform fast > gpurun_out/plain5.log 2>&1 && $N -k regex:wsm_se -s 1 -o gpurun_out/r02_se90_n64_fast python tools/prof_case.py se90 64 65536 uniform fast > gpurun_out/ncu5.log 2>&1
python tools/prof_case.py se99 16 262144 uniform fast > gpurun_out/plain6.log 2>&1 && $N -k regex:wip_kernel -s 1 -o gpurun_out/r02_se99_n16_fast python tools/prof_case.py se99 16 262144 uniform fast > gpurun_out/ncu6.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu7.log 2>&1
python tools/quick_bench.py --cases se99:4:1048576,se99:8:1048576,se99:16:1048576,se99:32:1048576,se90:64:200000,se99:96:40000,se99:128:20000 --reps 3 > gpurun_out/r02_sweep_a.txt 2>&1
python tools/quick_bench.py --cases gmw81:32:1048576,gmw81:64:200000,se90:32:1048576,gmw81:128:20000,gmw81:200:4096,se99:200:4096,gmw81:256:4096,se99:256:4096,se90:256:4096 --reps 3 >> gpurun_out/r02_sweep_a.txt 2>&1
python tools/quick_bench.py --cases gmw81:384:2048,se99:384:2048,gmw81:512:1184,se99:512:1184,gmw81:1024:592,se99:1024:592 --reps 2 --families uniform >> gpurun_out/r02_sweep_a.txt 2>&1
ls -la gpurun_out | tail -5
