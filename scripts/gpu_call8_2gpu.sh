set -x
mkdir -p gpurun_out
df -h /dev/shm /tmp | tail -2; nvidia-smi -L | head -3; nproc; cat /sys/fs/cgroup/cpu.max
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
LMG_BENCH_CPU_S=6 LMG_C3_GENOMES=4000 LMG_C3_QUERIES=2000 LMG_C3_WAVES=2 timeout 900 $TR bench.py --gpus 2 --config c3 --steps 2 --warmup 1 > gpurun_out/bench_c3_2gpu_mini.json 2> gpurun_out/bench_c3_2gpu_mini.err; tail -6 gpurun_out/bench_c3_2gpu_mini.err
timeout 600 $TR bench.py --gpus 2 --config c5 --steps 5 --warmup 2 > gpurun_out/bench_c5_2gpu.json 2> gpurun_out/bench_c5_2gpu.err; tail -4 gpurun_out/bench_c5_2gpu.err
LMG_BENCH_CPU_S=6 timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_c2_2gpu.json 2> gpurun_out/bench_c2_2gpu.err; tail -4 gpurun_out/bench_c2_2gpu.err
ls -la gpurun_out | head -30
