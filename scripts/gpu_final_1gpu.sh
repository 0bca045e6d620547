set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu_final.txt
tail -5 gpurun_out/pytest_gpu_final.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_c2_reference.json 2> gpurun_out/bench_final_c2_reference.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final_c2.json 2> gpurun_out/bench_final_c2.err; tail -3 gpurun_out/bench_final_c2.err
timeout 600 python bench.py --config c4 --steps 3 --warmup 1 > gpurun_out/bench_final_c4.json 2> gpurun_out/bench_final_c4.err
LMG_C5_PER_MASK=150000 timeout 300 python bench.py --config c5 --steps 5 --warmup 2 > gpurun_out/bench_final_c5_1gpu.json 2> gpurun_out/bench_final_c5_1gpu.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
LMG_BENCH_CPU_S=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 1200 --csv --log-file gpurun_out/launches_final_c2.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out | tail -12
