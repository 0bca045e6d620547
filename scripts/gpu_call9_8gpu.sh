set -x
mkdir -p gpurun_out
df -h /dev/shm /tmp 2>&1 | tail -2 > gpurun_out/box_8gpu.txt; nproc >> gpurun_out/box_8gpu.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/box_8gpu.txt; free -g | head -2 >> gpurun_out/box_8gpu.txt; nvidia-smi -L >> gpurun_out/box_8gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
LMG_BENCH_CPU_S=10 LMG_C3_WAVES=2 timeout 1100 $TR bench.py --gpus 8 --config c3 --steps 3 --warmup 1 > gpurun_out/bench_c3_8gpu.json 2> gpurun_out/bench_c3_8gpu.err; tail -12 gpurun_out/bench_c3_8gpu.err | cut -c1-300
rm -rf /tmp/lmg_bench/c3d_*
timeout 400 $TR bench.py --gpus 8 --config c5 --steps 5 --warmup 2 > gpurun_out/bench_c5_8gpu.json 2> gpurun_out/bench_c5_8gpu.err; tail -3 gpurun_out/bench_c5_8gpu.err | cut -c1-300
LMG_BENCH_CPU_S=0 timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_c2_8gpu.json 2> gpurun_out/bench_c2_8gpu.err; tail -3 gpurun_out/bench_c2_8gpu.err | cut -c1-300
ls -la gpurun_out | head -30
