#!/bin/bash
# full ncu captures (source-level samples) of t3 GEMM launches (skip counts given) and a metric pass.  Usage: bash tools/gpu_t3ncu.sh <tag> <skip>...
set -u
TAG=${1:-t3}; shift
OUT=gpurun_out; mkdir -p $OUT
CMD="python tools/prof_classes.py fp32 256 400 1"
for S in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:t3_gemm -s $S -c 1 -o $OUT/prof_t3gemm_${TAG}_$S $CMD > $OUT/ncu_t3gemm_${TAG}_$S.log 2>&1; echo "gemm $S rc=$?"
done
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__registers_per_thread
timeout 600 ncu --metrics $M --clock-control none -k regex:t3_ -s 40 -c 12 --csv --log-file $OUT/ncu_t3_${TAG}.csv $CMD > $OUT/ncu_t3_${TAG}.log 2>&1; echo "ncu rc=$?"
