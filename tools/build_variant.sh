#!/bin/bash
# usage: tools/build_variant.sh <name> <extra nvcc flags...>  -> build/libplantos_<name>.so (experiments only)
NAME=$1; shift
mkdir -p build
cd rl_env_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" -o ../../build/libplantos_$NAME.so plantos_abi.cu
