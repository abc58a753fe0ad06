#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list + full capture of the two hot kernels.
# Usage (here): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh rNN'
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 6 --warmup 3 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?"
# BASELINE configs[3] size on one GPU (maximum batch): 524 288 instances
timeout 400 python bench.py --batch 524288 --steps 10 --warmup 5 --no-cpu --no-extras > $OUT/bench_524k_$TAG.json 2> $OUT/bench_524k_$TAG.err; echo "bench 524k rc=$?"
CMD="python bench.py --steps 3 --warmup 12 --no-cpu --no-extras"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu1_$TAG.log 2>&1
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"mpc_chain|mpc_prepare|ik_fast|ik_kernel" -s 86 -c 6 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu2_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ik_osqp" -s 4 -c 1 -f -o $OUT/prof_osqp_$TAG python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu3_$TAG.log 2>&1
echo done
tail -c 1500 $OUT/pytest_gpu_$TAG.log
cat $OUT/bench_$TAG.json | head -c 3000
