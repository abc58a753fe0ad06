#!/bin/bash
# horizon sweep (BASELINE configs[4] + the shipped N = 200): full tick at B = 65536 on one GPU, the kernel family wc_mpc_create picks
for n in "$@"; do
  timeout 600 python bench.py --no-cpu --no-extras --steps 8 --warmup 4 --horizon $n 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); k=d['kernels_ms']
print('| $n | %s | %.3e | %.2f | %.2f | %.3e | %.3e |' % (d['mpc_path'], d['value'], k['mpc_admm'] + k['mpc_prepare'], k['fk_qpik'], d['e2e']['value'], d['closed_loop']['value']))"
done
