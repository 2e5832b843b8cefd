#!/bin/bash
# One gpurun call while developing the tf32 x 3 path: its parity tests, then (if green) the whole GPU suite and the cfg3 timing of
# precision fp32 / fp32_simt.   Usage: bash tools/gpu_t3.sh <tag>
set -u
TAG=${1:-t3}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_t3.py -m gpu -q -rA > $OUT/pytest_t3_${TAG}.log 2>&1; RC=$?; echo "t3 pytest rc=$RC"
grep -E "max\|d\||passed|failed|Error|error" $OUT/pytest_t3_${TAG}.log | sort | uniq | head -150
if [ $RC -eq 0 ]; then
  timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/pytest_all_${TAG}.log 2>&1; echo "all pytest rc=$?"; tail -5 $OUT/pytest_all_${TAG}.log
  for P in fp32 fp32_simt; do
    timeout 300 python bench.py --steps 5 --warmup 3 --timed-only --precision $P > $OUT/timed_${P}_${TAG}.json 2> $OUT/timed_${P}_${TAG}.err; echo "$P rc=$?"; head -c 400 $OUT/timed_${P}_${TAG}.json; echo
  done
fi
