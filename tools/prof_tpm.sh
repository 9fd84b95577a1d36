N="ncu --set full --clock-control none --import-source on -c 1"
python tools/prof_case.py se99 4 1048576 uniform fast > gpurun_out/plain_tpm4.log 2>&1 && $N -k regex:tpm_kernel -s 1 -o gpurun_out/r02_tpm_se99_n4 python tools/prof_case.py se99 4 1048576 uniform fast > gpurun_out/ncu_tpm4.log 2>&1
tail -2 gpurun_out/plain_tpm4.log gpurun_out/ncu_tpm4.log
