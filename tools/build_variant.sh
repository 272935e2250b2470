#!/bin/bash
# usage: build_variant.sh <name> [-Dflags...]   -> scratch/variants/<name>.so (+ .log with ptxas -v)
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -shared -DQMPC_NO_XCHECK \
  -Xptxas -v "$@" -o scratch/variants/$name.so quaternion_mpc_b200/csrc/*.cu > scratch/variants/$name.log 2>&1
grep -A2 "qmpc_coop_kernelINS_9QuatModelILi4" scratch/variants/$name.log | grep -E "registers|spill" | head -4
