set -x
mkdir -p gpurun_out
LMG_BENCH_CPU_S=0 LMG_C3_SIM_WORLD=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pa_anchors3' --launch-skip 2 -c 1 -f -o gpurun_out/prof_r2f python bench.py --config c3 --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -4
