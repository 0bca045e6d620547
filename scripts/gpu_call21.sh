set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_final2.json 2> gpurun_out/bench_c2_final2.err
timeout 300 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/bench_c4_final.json 2> gpurun_out/bench_c4_final.err
LMG_BENCH_CPU_S=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 1500 --csv --log-file gpurun_out/launches_c2_final.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out | tail -5
