#!/bin/bash
# Development helper (GPU box): per-phase clocks (clk variant) and the event trace (trace variant) of the fused kernel at cfg3.
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-d}
EDTTS_LAYER_CLOCKS=1 EDTTS_LIB=$PWD/edge_diffusion_tts_b200/lib/libedtts_clk.so timeout 300 python bench.py --steps 1 --warmup 1 --timed-only --no-graph > $OUT/clk_${TAG}.json 2> $OUT/clk_${TAG}.err; echo "clk rc=$?"
tail -12 $OUT/clk_${TAG}.err
EDTTS_LIB=$PWD/edge_diffusion_tts_b200/lib/libedtts_trace.so timeout 300 python bench.py --steps 1 --warmup 1 --timed-only --no-graph > $OUT/trace_${TAG}.json 2> $OUT/trace_${TAG}.err; echo "trace rc=$?"
grep -c TRACE $OUT/trace_${TAG}.err
