# Round-2 ncu evidence (run under gpurun, one GPU): launch list of the bench command and full
# captures of the nf=8 full-trajectory kernel and the headline fused kernel.
set -x
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_list.log 2>&1
python scripts/dev_probe.py c4 8192 2>&1 | grep -E "^K2" | tail -2 && \
ncu --set full --clock-control none --import-source on -k regex:fo_rk4_fact -s 1 -c 1 -f -o gpurun_out/prof_r2_c4_nf8_full \
    python scripts/dev_probe.py c4 8192 > gpurun_out/ncu_c4full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fo1_fused -s 2 -c 1 -f -o gpurun_out/prof_r2_fused_headline \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches.csv
