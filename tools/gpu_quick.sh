#!/bin/bash
# Quick kernel iteration on the GPU box: the tensor-core parity tests, then the device-timed cfg3 number (and small batches).
# Usage: bash tools/gpu_quick.sh <tag> [extra bench batches...]
set -u
TAG=${1:-q}; shift || true
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_pinned_configs.py -m gpu -q -x > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_${TAG}.log
timeout 300 python bench.py --steps 20 --warmup 5 --timed-only > $OUT/timed_${TAG}.json 2> $OUT/timed_${TAG}.err; echo "timed rc=$?"; cat $OUT/timed_${TAG}.json
for B in "$@"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --timed-only --batch $B 2>/dev/null | sed "s/^/batch $B: /"
done
