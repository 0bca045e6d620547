set -x
mkdir -p gpurun_out
LMG_BENCH_CPU_S=0 LMG_C3_SIM_WORLD=8 LMG_DEBUG_TIMING=1 LMG_LANES=1 timeout 900 python bench.py --config c3 --steps 1 --warmup 1 > gpurun_out/bench_c3_rank0_l1.json 2> gpurun_out/bench_c3_rank0_l1.err
grep -E "lmg host" gpurun_out/bench_c3_rank0_l1.err | tail -45
LMG_BENCH_CPU_S=0 LMG_C3_SIM_WORLD=8 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 70 -c 900 --csv --log-file gpurun_out/launches_c3_rank0.csv python bench.py --config c3 --steps 1 --warmup 1 > gpurun_out/ncu_c3.log 2>&1
LMG_BENCH_CPU_S=0 LMG_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_probe_find2' --launch-skip 1 -c 2 -f -o gpurun_out/prof_r2e python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
