#!/bin/bash
# full ncu captures (source-level samples) of one t3 GEMM launch and one t3 cross-attention launch.  Usage: bash tools/gpu_t3ncu.sh <tag>
set -u
TAG=${1:-t3}
OUT=gpurun_out; mkdir -p $OUT
CMD="python tools/prof_classes.py fp32 256 400 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:t3_gemm -s 55 -c 1 -o $OUT/prof_t3gemm_${TAG} $CMD > $OUT/ncu_t3gemm_${TAG}.log 2>&1; echo "gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:t3_attn -s 9 -c 1 -o $OUT/prof_t3attn_${TAG} $CMD > $OUT/ncu_t3attn_${TAG}.log 2>&1; echo "attn rc=$?"
