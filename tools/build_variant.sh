#!/bin/bash
# Development helper: build libedtts_<name>.so with extra nvcc flags on tc_layer.cu (A/B of compile-time kernel options).
# Usage: tools/build_variant.sh <name> "<extra flags>"; run with EDTTS_LIB=edge_diffusion_tts_b200/lib/libedtts_<name>.so
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
python -c "import sys; sys.path.insert(0, '$ROOT'); import __graft_entry__ as g; g.build()" > /dev/null
mkdir -p $ROOT/build/var_$NAME
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets $@ \
  -c -o $ROOT/build/var_$NAME/tc_layer.o $ROOT/edge_diffusion_tts_b200/csrc/tc_layer.cu
OBJS=$(ls $ROOT/build/*.o | grep -v "/tc_layer.o")
nvcc -shared -o $ROOT/edge_diffusion_tts_b200/lib/libedtts_$NAME.so $OBJS $ROOT/build/var_$NAME/tc_layer.o
echo built libedtts_$NAME.so
