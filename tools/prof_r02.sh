# Round-2 ncu captures (one GPU; every ncu run follows a plain run of the same command) and the
# size sweep.  Summaries: tools/make_profiles.sh.
set -x
N="ncu --set full --clock-control none --import-source on -c 1"
cap() {  # name kernel-regex algo n batch family mode
  python tools/prof_case.py $3 $4 $5 $6 $7 > gpurun_out/plain_$1.log 2>&1 && $N -k regex:$2 -s 1 -o gpurun_out/$1 python tools/prof_case.py $3 $4 $5 $6 $7 > gpurun_out/ncu_$1.log 2>&1
}
cap r02_se99_n32_fast wip_kernel se99 32 131072 uniform fast
cap r02_se99_n32_fast_wishart wip_kernel se99 32 131072 wishart fast
cap r02_se99_n32_strict wsm_se se99 32 131072 uniform strict
cap r02_gmw81_n256_fast bsm_gmw gmw81 256 1184 uniform fast
cap r02_se99_n256_fast bsm_se se99 256 1184 uniform fast
cap r02_se90_n64_fast wip2_kernel se90 64 65536 uniform fast
cap r02_se99_n16_fast wipp_kernel se99 16 262144 uniform fast
cap r02_verify_n32 verify_warp se99 32 131072 uniform fast
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu7.log 2>&1
python tools/quick_bench.py --cases se99:4:1048576,se99:8:1048576,se99:12:1048576,se99:16:1048576,se99:24:1048576,se99:32:1048576,se99:48:200000,se90:64:200000,se99:96:40000,se99:128:20000 --reps 3 > gpurun_out/r02_sweep_a.txt 2>&1
python tools/quick_bench.py --cases gmw81:16:1048576,gmw81:32:1048576,gmw81:64:200000,se90:16:1048576,se90:32:1048576,gmw81:128:20000,gmw81:200:4096,se99:200:4096,gmw81:256:4096,se99:256:4096,se90:256:4096 --reps 3 >> gpurun_out/r02_sweep_a.txt 2>&1
python tools/quick_bench.py --cases gmw81:384:2048,se99:384:2048,gmw81:512:1184,se99:512:1184,gmw81:1024:592,se99:1024:592 --reps 2 --families uniform >> gpurun_out/r02_sweep_a.txt 2>&1
MODCHOL_LWM_MIN=65 python tools/quick_bench.py --cases gmw81:128:20000,se99:128:20000,gmw81:256:4096,se99:256:4096 --modes fast --families uniform --reps 2 > gpurun_out/r02_sweep_warp_panel.txt 2>&1
ls -la gpurun_out | tail -5
