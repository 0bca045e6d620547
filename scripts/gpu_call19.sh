set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
B="timeout 300 python bench.py --config c4 --steps 5 --warmup 3"
LMG_BENCH_CPU_S=0 $B > gpurun_out/bench_c4_reg8.json 2> gpurun_out/bench_c4_reg8.err
LMG_NO_WFA_REG8=1 LMG_BENCH_CPU_S=0 $B > gpurun_out/bench_c4_noreg8.json 2> gpurun_out/bench_c4_noreg8.err
LMG_LANES=4 LMG_BENCH_CPU_S=0 $B > gpurun_out/bench_c4_reg8_l4.json 2> gpurun_out/bench_c4_reg8_l4.err
LMG_LANES=1 LMG_DEBUG_TIMING=1 LMG_BENCH_CPU_S=0 timeout 300 python bench.py --config c4 --steps 1 --warmup 1 > gpurun_out/bench_c4_dbg.json 2> gpurun_out/bench_c4_dbg.err
LMG_BENCH_CPU_S=0 timeout 300 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_c2_after_reg8.json 2> gpurun_out/bench_c2_after_reg8.err
ls -la gpurun_out | tail -8
