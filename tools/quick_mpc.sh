set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mpc_gpu.py -m gpu -x -q 2>&1 | tail -30
timeout 600 python bench.py --no-cpu --steps 20 --warmup 5 2>&1 | tail -3
