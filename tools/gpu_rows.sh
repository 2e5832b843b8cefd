#!/bin/bash
# Kernel-class rows: CUDA-event table (L2 flushed between repetitions) + one ncu pass with DRAM bytes and pipe utilisation.
set -u
TAG=${1:-r2}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python tools/kernel_rows.py > $OUT/rows_${TAG}.txt 2> $OUT/rows_${TAG}.err; echo "rows rc=$?"; grep -v "^JSON\|^\[build" $OUT/rows_${TAG}.txt
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $OUT/rows_ncu_${TAG}.csv python tools/kernel_rows.py --once > $OUT/rows_ncu_${TAG}.log 2>&1; echo "ncu rc=$?"
python tools/kernel_rows.py --summarise $OUT/rows_ncu_${TAG}.csv > $OUT/rows_ncu_${TAG}.txt 2>&1; tail -60 $OUT/rows_ncu_${TAG}.txt
