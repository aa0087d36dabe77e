#!/usr/bin/env bash
# ncu --set full of one SOR kernel on the 256^3 (and optionally S1) porous channel
# usage: tools/r02_ncu.sh TAG VARIANT KERNEL_REGEX [LIB] [GRID...]
set -u
tag=$1; variant=$2; regex=$3; lib=${4:-}; shift 4 || true
grid=${*:-256 256 256}
out=gpurun_out
mkdir -p $out
libarg=""
if [ -n "$lib" ] && [ "$lib" != "-" ]; then export PIXELFLOW_GPU_LIB=$PWD/$lib; fi
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s 30 -c 2 -f -o $out/${tag} \
  python tools/sor_lab.py --variant $variant --grid $grid --steps 1 --warmup 1 --iters 40 --graph 0 > $out/${tag}_run.log 2>&1
echo "ncu rc=$?"; tail -3 $out/${tag}_run.log
ls -la $out/${tag}.ncu-rep
