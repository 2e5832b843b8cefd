#!/bin/bash
# One gpurun call for a committed state: GPU tests, smoke, bench line, ncu launch list and one full capture of the fused kernel.
# Usage (on the GPU box): bash tools/gpu_final.sh <tag>
set -u
TAG=${1:-r2e}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_${TAG}.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; grep -i smoke $OUT/smoke_${TAG}.log | tail -3
timeout 900 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"; tail -3 $OUT/bench_${TAG}.err
CMD="python bench.py --steps 2 --warmup 3 --timed-only --no-graph"
timeout 300 $CMD > $OUT/plain_${TAG}.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 3 --timed-only --no-graph"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_layer -s 13 -c 1 -o $OUT/prof_${TAG} $CMD2 > $OUT/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
head -c 700 $OUT/bench_${TAG}.json; echo
# the fp32-grade tensor-core path (precision="fp32"): per-class times and one ncu metric pass over the kernels of a decoder step
timeout 300 python tools/prof_classes.py fp32 > $OUT/classes_fp32_${TAG}.txt 2>&1; tail -9 $OUT/classes_fp32_${TAG}.txt
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__registers_per_thread
timeout 600 ncu --metrics $M --clock-control none -k regex:t3_ -s 66 -c 20 --csv --log-file $OUT/ncu_t3_${TAG}.csv python tools/prof_classes.py fp32 256 400 1 > $OUT/ncu_t3_${TAG}.log 2>&1; echo "t3 ncu rc=$?"
