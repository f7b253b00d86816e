#!/bin/bash
# times the QUADPACK node kernel of every library under _variants/ (tools/bench_quadpack.py); run on a GPU box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ref=""
for so in _variants/*.so; do
  name=$(basename $so .so)
  AP_LIB_PATH=$PWD/$so python tools/bench_quadpack.py --n 2000000 --out gpurun_out/q_$name.npy ${ref:+--ref $ref} 2>/dev/null | tail -1
  [ -z "$ref" ] && ref=gpurun_out/q_$name.npy
done
rm -f gpurun_out/q_*.npy
