#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + full captures of the hot kernels.
# Usage (here): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh rNN'            (everything but the osqp capture)
#               gpurun --timeout 900 -- 'PARTS="osqp" bash tools/gpu_round.sh rNN'  (selected parts)
# What a call writes under gpurun_out/ must stay below 64 MiB, or nothing is copied back: the two full-set captures
# (~55 MB and ~15 MB with sources) go in separate calls.
set -u
TAG=${1:-r1}
PARTS=${PARTS:-"tests bench ref big launches full"}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $PARTS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.csv 2>&1
if has tests; then timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log; tail -c 1500 $OUT/pytest_gpu_$TAG.log; fi
if has bench; then timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; head -c 4000 $OUT/bench_$TAG.json; fi
if has ref; then timeout 300 python bench.py --impl reference --steps 6 --warmup 3 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?"; fi
# BASELINE configs[3] size on one GPU (maximum batch): 524 288 instances
if has big; then timeout 400 python bench.py --batch 524288 --steps 10 --warmup 5 --no-cpu --no-extras > $OUT/bench_524k_$TAG.json 2> $OUT/bench_524k_$TAG.err; echo "bench 524k rc=$?"; head -c 1200 $OUT/bench_524k_$TAG.json; fi
CMD="python bench.py --steps 3 --warmup 12 --no-cpu --no-extras"
if has launches; then
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu1_$TAG.log 2>&1
fi
if has full; then
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"mpc_split|mpc_prepare|mpc_chain_setup|ik_tpi|ik_kernel" -s 60 -c 6 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu2_$TAG.log 2>&1
fi
if has osqp; then
CMD3="python bench.py --steps 3 --warmup 3 --no-cpu"
$CMD3 > $OUT/plain3_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ik_osqp" -s 4 -c 1 -f -o $OUT/prof_osqp_$TAG $CMD3 > $OUT/ncu3_$TAG.log 2>&1
tail -c 1500 $OUT/plain3_$TAG.log
fi
du -sh $OUT
echo done
