set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
B="timeout 600 python bench.py --steps 8 --warmup 3"
LMG_BENCH_CPU_S=0 $B > gpurun_out/bench_c2_hints.json 2> gpurun_out/bench_c2_hints.err
LMG_NO_L2_HINTS=1 LMG_BENCH_CPU_S=0 $B > gpurun_out/bench_c2_nohints.json 2> gpurun_out/bench_c2_nohints.err
LMG_CSTART32=1 LMG_BENCH_CPU_S=0 $B > gpurun_out/bench_c2_cstart32.json 2> gpurun_out/bench_c2_cstart32.err
LMG_L2_FETCH=32 LMG_BENCH_CPU_S=0 $B > gpurun_out/bench_c2_fetch32.json 2> gpurun_out/bench_c2_fetch32.err
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum
LMG_BENCH_CPU_S=0 LMG_LANES=1 timeout 400 ncu --metrics $M --clock-control none -k regex:'k_probe_find2' --launch-skip 1 -c 2 --csv --log-file gpurun_out/ncu_find2_hints.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_a.log 2>&1
LMG_NO_L2_HINTS=1 LMG_BENCH_CPU_S=0 LMG_LANES=1 timeout 400 ncu --metrics $M --clock-control none -k regex:'k_probe_find2' --launch-skip 1 -c 2 --csv --log-file gpurun_out/ncu_find2_nohints.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_b.log 2>&1
LMG_L2_FETCH=32 LMG_BENCH_CPU_S=0 LMG_LANES=1 timeout 400 ncu --metrics $M --clock-control none -k regex:'k_probe_find2' --launch-skip 1 -c 2 --csv --log-file gpurun_out/ncu_find2_fetch32.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_c.log 2>&1
timeout 300 python bench.py --config c5 --steps 5 --warmup 2 > gpurun_out/bench_c5_hints.json 2> gpurun_out/bench_c5_hints.err
LMG_NO_L2_HINTS=1 timeout 300 python bench.py --config c5 --steps 5 --warmup 2 > gpurun_out/bench_c5_nohints.json 2> gpurun_out/bench_c5_nohints.err
ls -la gpurun_out | tail -12
