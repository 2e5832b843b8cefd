#!/bin/bash
# Development helper (GPU box): time the cfg3 generate with each library variant given.  Usage: [AB_ARGS="--batch 148"] tools/ab.sh name1 name2 ...
for v in "$@"; do
  L=edge_diffusion_tts_b200/lib/libedtts_$v.so
  [ "$v" = "base" ] && L=edge_diffusion_tts_b200/lib/libedtts.so
  r=$(EDTTS_LIB=$PWD/$L timeout 200 python bench.py --steps 10 --warmup 3 --timed-only $AB_ARGS 2>/dev/null | python -c "import sys,json; print(json.loads(sys.stdin.read())['ms_per_step'])")
  echo "$v ms_per_step $r"
done
