set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
LMG_BENCH_CPU_S=0 timeout 300 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/bench_c4_fastgen.json 2> gpurun_out/bench_c4_fastgen.err
LMG_LANES=1 LMG_BENCH_CPU_S=0 timeout 300 python bench.py --config c4 --steps 3 --warmup 2 > gpurun_out/bench_c4_fastgen_l1.json 2> gpurun_out/bench_c4_fastgen_l1.err
LMG_LANES=3 LMG_BENCH_CPU_S=0 timeout 300 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/bench_c4_fastgen_l3.json 2> gpurun_out/bench_c4_fastgen_l3.err
ls -la gpurun_out | tail -5
