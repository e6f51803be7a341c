#!/bin/bash
# Build tuning variants of libpicgolf.so into particleincellcodegolf.jl_b200/lib/variants/ (git-ignored, travel with gpurun).
#   tools/build_variants.sh name1 "-DPG_CP_THREADS=128 -DPG_CP_MINBLOCKS=3" name2 "..."
# Select one at run time with PICGOLF_LIB=particleincellcodegolf.jl_b200/lib/variants/libpicgolf_<name>.so
set -e
cd "$(dirname "$0")/.."
out=particleincellcodegolf.jl_b200/lib/variants
mkdir -p $out
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -shared \
    $flags -Xptxas -v -o $out/libpicgolf_$name.so particleincellcodegolf.jl_b200/csrc/picgolf.cu -ldl 2>&1 | grep -A2 "fp_pass_polyILb0" | grep -v "^--" | sed "s/^/[$name] /" &
done
wait
