set -x
mkdir -p gpurun_out
LMG_BENCH_CPU_S=0 LMG_DEBUG_TIMING=1 LMG_LANES=1 timeout 400 python bench.py --config c4 --steps 1 --warmup 1 > gpurun_out/bench_c4_l1.json 2> gpurun_out/bench_c4_l1.err
grep "register WFA" gpurun_out/bench_c4_l1.err | tail -3
LMG_BENCH_CPU_S=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
