set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
LMG_BENCH_CPU_S=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2_prio.json 2> gpurun_out/bench_c2_prio.err
LMG_NO_PRIO_LOOKUP=1 LMG_BENCH_CPU_S=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2_noprio.json 2> gpurun_out/bench_c2_noprio.err
LMG_BENCH_CPU_S=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2_prio2.json 2> gpurun_out/bench_c2_prio2.err
ls -la gpurun_out | tail -5
