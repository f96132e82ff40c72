set -x
PF_EVERY=8163 python scripts/dev_probe.py c5 1048576 2>&1 | grep -E "^K2|SOLVE" | tail -3 && PF_EVERY=8163 ncu --set full --clock-control none --import-source on -k regex:fo1_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_c5_fo1_stateonly python scripts/dev_probe.py c5 1048576 > gpurun_out/ncu_c5.log 2>&1
PF_EVERY=64 python scripts/dev_probe.py c3 1048576 2>&1 | grep -E "^K2|SOLVE" | tail -3 && PF_EVERY=64 ncu --set full --clock-control none --import-source on -k regex:fo_rk4_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2_c3_nf2_strided python scripts/dev_probe.py c3 1048576 > gpurun_out/ncu_c3.log 2>&1
PF_EVERY=512 python scripts/dev_probe.py c4 65536 2>&1 | grep -E "^K2|SOLVE" | tail -3 && PF_EVERY=512 ncu --set full --clock-control none --import-source on -k regex:fo_rk4_fact -s 1 -c 1 -f -o gpurun_out/prof_r2_c4_nf8_fact python scripts/dev_probe.py c4 65536 > gpurun_out/ncu_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep
