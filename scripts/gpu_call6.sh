set -x
mkdir -p gpurun_out
LMG_BENCH_CPU_S=0 LMG_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_wfa_reg' --launch-skip 2 -c 1 -f -o gpurun_out/prof_r2d python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
