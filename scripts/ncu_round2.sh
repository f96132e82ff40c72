# Round-2 ncu evidence (run under gpurun, one GPU), each capture only after the same command
# exited 0 without ncu:
#   * launch list of the bench command (headline loop only: the extra phases have their own lines)
#   * --set full captures of the headline fused kernel and of the compute-bound kernels
set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fo1_fused -s 2 -c 1 -f -o gpurun_out/prof_r2_fused_headline \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_fused.log 2>&1
PF_EVERY=8163 python scripts/dev_probe.py c5 1048576 2>&1 | grep -E "^K2" | tail -1 && \
PF_EVERY=8163 ncu --set full --clock-control none --import-source on -k regex:fo1_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2b_c5_fo1_stateonly \
    python scripts/dev_probe.py c5 1048576 > gpurun_out/ncu_c5.log 2>&1
PF_EVERY=64 python scripts/dev_probe.py c3 1048576 2>&1 | grep -E "^K2" | tail -1 && \
PF_EVERY=64 ncu --set full --clock-control none --import-source on -k regex:fo_rk4_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2b_c3_nf2_strided \
    python scripts/dev_probe.py c3 1048576 > gpurun_out/ncu_c3.log 2>&1
PF_EVERY=512 python scripts/dev_probe.py c4 65536 2>&1 | grep -E "^K2" | tail -1 && \
PF_EVERY=512 ncu --set full --clock-control none --import-source on -k regex:fo_rk4_fact -s 1 -c 1 -f -o gpurun_out/prof_r2b_c4_nf8_fact \
    python scripts/dev_probe.py c4 65536 > gpurun_out/ncu_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches.csv
