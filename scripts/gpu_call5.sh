set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -30 gpurun_out/pytest_gpu.txt
LMG_BENCH_CPU_S=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
LMG_C5_PER_MASK=150000 timeout 300 python bench.py --config c5 --steps 5 --warmup 2 > gpurun_out/bench_c5_1gpu.json 2> gpurun_out/bench_c5_1gpu.err; tail -3 gpurun_out/bench_c5_1gpu.err
LMG_BENCH_CPU_S=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_wfa_reg|k_pa_chain|k_extend_run' --launch-skip 6 -c 6 -f -o gpurun_out/prof_r2c python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
LMG_BENCH_CPU_S=0 LMG_DEBUG_TIMING=1 LMG_LANES=1 timeout 300 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_c2_l1.json 2> gpurun_out/bench_c2_l1.err
ls -la gpurun_out
