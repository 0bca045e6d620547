set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -30 gpurun_out/pytest_gpu.txt
LMG_BENCH_CPU_S=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
LMG_BENCH_CPU_S=0 timeout 400 python bench.py --config c4 --steps 3 --warmup 1 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -3 gpurun_out/bench_c4.err
LMG_BENCH_CPU_S=0 LMG_C3_GENOMES=2000 LMG_C3_QUERIES=1000 timeout 600 python bench.py --config c3 --steps 2 --warmup 1 > gpurun_out/bench_c3_mini.json 2> gpurun_out/bench_c3_mini.err; tail -5 gpurun_out/bench_c3_mini.err
LMG_C5_PER_MASK=150000 timeout 300 python bench.py --config c5 --steps 5 --warmup 2 > gpurun_out/bench_c5_1gpu.json 2> gpurun_out/bench_c5_1gpu.err; tail -3 gpurun_out/bench_c5_1gpu.err
LMG_BENCH_CPU_S=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 1200 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_b.log 2>&1
LMG_BENCH_CPU_S=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_probe_find2|k_wfa_reg|k_capture2|k_pa_anchors3' --launch-skip 8 -c 8 -f -o gpurun_out/prof_r2b python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
