#!/bin/bash
# Build an alternative libpixelflow_gpu.so for kernel tuning experiments:
#   tools/build_variant.sh NAME FILE -DMACRO=... ...   ->  exp/libpf_NAME.so  (run with PIXELFLOW_GPU_LIB=exp/libpf_NAME.so)
# FILE = the one source (without .cu) that is rebuilt with the extra flags, e.g. pf_sor_tma2; every other object is
# taken from the product build in pixelflow_b200/build (run `python -m pixelflow_b200.build` first).
set -e
cd "$(dirname "$0")/.."
name=$1; file=$2; shift 2
mkdir -p exp/obj_$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-O2,-fvisibility=hidden \
  -Xptxas -v "$@" -I include -c pixelflow_b200/csrc/$file.cu -o exp/obj_$name/$file.o 2> exp/obj_$name/$file.log
objs=""
for o in pixelflow_b200/build/*.cu.o; do
  if [ "$(basename $o)" != "$file.cu.o" ]; then objs="$objs $o"; fi
done
nvcc -shared -o exp/libpf_$name.so $objs exp/obj_$name/$file.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -ldl
grep "registers\|spill" exp/obj_$name/$file.log | tail -2
ls -la exp/libpf_$name.so
