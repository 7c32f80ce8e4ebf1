#!/bin/bash
# A/B build of one translation unit with extra -D flags: tools/ab_build.sh NAME FILE.cu "-DX=1 ..." -> build/NAME/librife_b200.so
# (select with RIFE_B200_LIB=build/NAME/librife_b200.so)
set -e
cd "$(dirname "$0")/../eccv2022-rife_b200/csrc"
name=$1; src=$2; flags=$3
out=../../build/$name; mkdir -p $out
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr -Xptxas -v $flags -c $src -o $out/${src%.cu}.o 2> $out/ptxas.log
objs=""; for f in api pixel_kernels conv_simt conv_umma conv_halo conv_strip conv_gather conv_head generic_ops scene_ssim; do if [ -f $out/$f.o ]; then objs="$objs $out/$f.o"; else objs="$objs $f.o"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/librife_b200.so $objs -cudart static
grep -A3 "block_input_h_kernelILi2" $out/ptxas.log | grep -i "used\|spill" | head -4
