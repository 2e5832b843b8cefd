#!/bin/bash
# per-class times of the tf32 x 3 path at cfg3 + a short ncu metric pass over its kernels.  Usage: bash tools/gpu_t3prof.sh <tag>
set -u
TAG=${1:-t3}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/prof_classes.py fp32 > $OUT/classes_fp32_${TAG}.txt 2>&1; cat $OUT/classes_fp32_${TAG}.txt | tail -12
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers
timeout 600 ncu --metrics $M --clock-control none -k regex:t3_ -s 40 -c 36 --csv --log-file $OUT/ncu_t3_${TAG}.csv python tools/prof_classes.py fp32 256 400 1 > $OUT/ncu_t3_${TAG}.log 2>&1; echo "ncu rc=$?"
