set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_boundary.py -m gpu -q -k "switches" --timeout 180 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_switches.txt
tail -5 gpurun_out/pytest_switches.txt
