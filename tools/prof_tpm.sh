# ncu captures of the thread-per-matrix kernels (each after a plain run of the same command)
N="ncu --set full --clock-control none --import-source on -c 1"
cap() {  # name kernel-regex algo n batch family mode
  python tools/prof_case.py $3 $4 $5 $6 $7 > gpurun_out/plain_$1.log 2>&1 && $N -k regex:$2 -s 1 -o gpurun_out/$1 python tools/prof_case.py $3 $4 $5 $6 $7 > gpurun_out/ncu_$1.log 2>&1
}
cap r02_tpr_se99_n4 tpr_kernel se99 4 4194304 uniform fast
cap r02_tpr_se99_n8 tpr_kernel se99 8 1048576 uniform fast
cap r02_tps_se99_n12 tps_kernel se99 12 1048576 uniform strict
