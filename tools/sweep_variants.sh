#!/bin/bash
# bench line of every library variant under build_variants/; run on the GPU box
for f in build_variants/lib_*.so; do
  WC_LIB_PATH=$PWD/$f timeout 300 python bench.py --no-cpu --no-extras --steps 40 --warmup 10 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); k=d['kernels_ms']; print('$f', round(d['value']/1e6,2), {a: round(b,4) for a,b in k.items()})"
done
