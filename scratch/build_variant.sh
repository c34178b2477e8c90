#!/bin/bash
# usage: scratch/build_variant.sh NAME "-DYC_EM_UNROLL1=2 ..."   -> scratch/variants/NAME.so (git-ignored, travels to the GPU box)
set -e
name=$1; shift
cd "$(dirname "$0")/../yanocomp_b200/csrc"
out=../../scratch/variants/$name
mkdir -p $out
for f in yc_api yc_ks yc_ks_warp yc_fit yc_fit_warp yc_assemble; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC "$@" -c $f.cu -o $out/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/variants/$name.so $out/*.o
rm -rf $out
echo built scratch/variants/$name.so
