set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -12 gpurun_out/pytest_gpu.txt
LMG_BENCH_CPU_S=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
LMG_BENCH_CPU_S=0 LMG_C3_SIM_WORLD=8 timeout 900 python bench.py --config c3 --steps 2 --warmup 1 > gpurun_out/bench_c3_rank0.json 2> gpurun_out/bench_c3_rank0.err; tail -3 gpurun_out/bench_c3_rank0.err
ls -la gpurun_out | tail -8
