#!/bin/bash
# Multi-GPU round on one box: NCCL bitwise test + bench at N ranks.  Usage: bash tools/gpu_multi.sh <tag> <N> [N2 ...]
set -u
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_${TAG}.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $OUT/pytest_multi_${TAG}.log 2>&1; echo "pytest multi rc=$?"; tail -3 $OUT/pytest_multi_${TAG}.log
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --no-cpu-baseline --no-configs > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
  fi
  echo "bench N=$N rc=$?"; python - <<P
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_n$N.json"))
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","sharded_equals_single_gpu_bitwise")}, "e2e", d["e2e"]["value"], "weak", d.get("weak"), "cfg5", d.get("cfg5"))
except Exception as e:
    print("no line:", e); print(open("$OUT/bench_${TAG}_n$N.err").read()[-1500:])
P
done
