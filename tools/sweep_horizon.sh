#!/bin/bash
# horizon sweep (BASELINE configs[4] + the shipped N = 200): full tick on one GPU, chain path vs cooperative path
for n in "$@"; do
  b=65536; if [ $n -ge 100 ]; then b=16384; fi
  for path in auto coop; do
    if [ $path = coop ]; then export WC_MPC_PATH=coop; else unset WC_MPC_PATH; fi
    timeout 400 python bench.py --no-cpu --steps 12 --warmup 4 --horizon $n --batch $b 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=$n B=$b', d['config']['mpc_path'], 'ticks/s %.3e' % d['value'], 'mpc_admm ms %.3f' % d['kernels_ms']['mpc_admm'], 'e2e %.3e' % d['e2e']['value'])"
  done
done
