set -u
timeout 900 python -m pytest tests/test_ik_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --no-cpu --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['kernels_ms'], d['stats'], d['e2e'])"
