#!/bin/bash
# One gpurun call: full bench line, ncu launch list of the same command, one full ncu capture of the fused kernel.
# Usage (on the GPU box): bash tools/profile_round.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --timed-only --no-graph"
timeout 300 $CMD > $OUT/plain_${TAG}.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 3 --timed-only --no-graph"
timeout 300 $CMD2 > $OUT/plain2_${TAG}.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_layer -s 13 -c 1 -o $OUT/prof_${TAG} $CMD2 > $OUT/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
