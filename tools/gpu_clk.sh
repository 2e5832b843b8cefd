#!/bin/bash
# Development helper (GPU box): per-phase clocks of the fused kernel at cfg3 for each named library variant (built with -DEDTTS_DEBUG_CLOCKS).
set -u
OUT=gpurun_out; mkdir -p $OUT
for v in "$@"; do
  EDTTS_LAYER_CLOCKS=1 EDTTS_LIB=$PWD/edge_diffusion_tts_b200/lib/libedtts_$v.so timeout 300 python bench.py --steps 1 --warmup 1 --timed-only --no-graph > $OUT/clk_$v.json 2> $OUT/clk_$v.err; echo "$v rc=$?"
  tail -4 $OUT/clk_$v.err
done
