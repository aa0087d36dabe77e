#!/bin/bash
# Build an alternative libpixelflow_gpu.so for kernel tuning experiments:
#   tools/build_variant.sh NAME -DPF_TMA_TR=8 -DPF_TMA_TW=32 ...   ->  exp/libpf_NAME.so  (run with PIXELFLOW_GPU_LIB=exp/libpf_NAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p exp/obj_$name
objs=""
for f in pf_api pf_kernels pf_sor pf_sor_fused pf_sor_tma pf_comm; do
  extra=""
  if [ $f = pf_sor_tma ]; then extra="$*"; fi
  if [ $f = pf_sor_tma ] || [ ! -f exp/obj_common/$f.o ]; then
    mkdir -p exp/obj_common
    out=exp/obj_common/$f.o
    if [ $f = pf_sor_tma ]; then out=exp/obj_$name/$f.o; fi
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-O2,-fvisibility=hidden \
      -Xptxas -v $extra -I include -c pixelflow_b200/csrc/$f.cu -o $out 2> exp/obj_$name/$f.log &
  fi
done
wait
for f in pf_api pf_kernels pf_sor pf_sor_fused pf_comm; do objs="$objs exp/obj_common/$f.o"; done
nvcc -shared -o exp/libpf_$name.so $objs exp/obj_$name/pf_sor_tma.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -ldl
grep -A2 "sor_tma_kernel" exp/obj_$name/pf_sor_tma.log | grep "registers\|spill" | head -3
ls -la exp/libpf_$name.so
