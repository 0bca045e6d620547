set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
B="timeout 600 python bench.py --steps 8 --warmup 3"
LMG_BENCH_CPU_S=0 $B > gpurun_out/bench_c2_regroup.json 2> gpurun_out/bench_c2_regroup.err
LMG_NO_REGROUP=1 LMG_BENCH_CPU_S=0 $B > gpurun_out/bench_c2_noregroup.json 2> gpurun_out/bench_c2_noregroup.err
LMG_L2_HINTS=1 LMG_BENCH_CPU_S=0 $B > gpurun_out/bench_c2_regroup_hints.json 2> gpurun_out/bench_c2_regroup_hints.err
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum
LMG_BENCH_CPU_S=0 LMG_LANES=1 timeout 400 ncu --metrics $M --clock-control none -k regex:'k_probe_find2|k_surv_|k_bucket_scan' --launch-skip 4 -c 8 --csv --log-file gpurun_out/ncu_find2_regroup.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_a.log 2>&1
LMG_C5_PER_MASK=150000 timeout 300 python bench.py --config c5 --steps 5 --warmup 2 > gpurun_out/bench_c5_150k_regroup.json 2> gpurun_out/bench_c5_150k_regroup.err
LMG_NO_REGROUP=1 LMG_C5_PER_MASK=150000 timeout 300 python bench.py --config c5 --steps 5 --warmup 2 > gpurun_out/bench_c5_150k_noregroup.json 2> gpurun_out/bench_c5_150k_noregroup.err
timeout 300 python bench.py --config c5 --steps 5 --warmup 2 > gpurun_out/bench_c5_500k_regroup.json 2> gpurun_out/bench_c5_500k_regroup.err
ls -la gpurun_out | tail -12
