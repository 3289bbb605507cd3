#!/bin/bash
# Builds an A/B variant of the library under build_ab/: tools/build_ab.sh <name> <-D flags...>
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -o build_ab/$name.so astro_b200/csrc/astro_b200.cu && echo built build_ab/$name.so
