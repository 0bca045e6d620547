set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
LMG_BENCH_CPU_S=0 LMG_LANES=1 timeout 400 ncu --metrics $M --clock-control none -k regex:'k_probe_find2|k_surv_|k_bucket_scan' --launch-skip 3 -c 6 --csv --log-file gpurun_out/ncu_find2_regroup2.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_a.log 2>&1
LMG_C5_PER_MASK=150000 timeout 300 python bench.py --config c5 --steps 5 --warmup 2 > gpurun_out/bench_c5_150k_regroup2.json 2> gpurun_out/bench_c5_150k_regroup2.err
timeout 300 python bench.py --config c5 --steps 5 --warmup 2 > gpurun_out/bench_c5_500k_regroup2.json 2> gpurun_out/bench_c5_500k_regroup2.err
ls -la gpurun_out | tail -8
