set -x
nproc; free -g | head -2; df -h /tmp | tail -1; nvidia-smi -L; cat /sys/fs/cgroup/cpu.max 2>/dev/null; python -c "import os; print(len(os.sched_getaffinity(0)))"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_gpu.txt
tail -40 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -5 gpurun_out/bench_c2.err
LMG_C5_PER_MASK=150000 timeout 300 python bench.py --config c5 --steps 5 --warmup 2 > gpurun_out/bench_c5_1gpu.json 2> gpurun_out/bench_c5_1gpu.err; tail -3 gpurun_out/bench_c5_1gpu.err
timeout 400 python bench.py --config c4 --steps 3 --warmup 1 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -3 gpurun_out/bench_c4.err
LMG_BENCH_CPU_S=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out
