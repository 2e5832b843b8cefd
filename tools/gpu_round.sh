#!/bin/bash
# One gpurun call of a development round: GPU tests, smoke, bench (both arms).  Usage: bash tools/gpu_round.sh <tag> [pytest args]
set -u
TAG=${1:-r2}; shift || true
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $OUT/gpu_${TAG}.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q "$@" > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_${TAG}.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; grep smoke $OUT/smoke_${TAG}.log
timeout 900 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"; tail -3 $OUT/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/bench_ref_${TAG}.json 2> $OUT/bench_ref_${TAG}.err; echo "ref rc=$?"
head -c 600 $OUT/bench_${TAG}.json; echo; head -c 400 $OUT/bench_ref_${TAG}.json
