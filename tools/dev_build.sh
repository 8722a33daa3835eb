#!/bin/bash
# Development builds of the engine with only the BASELINE instantiation (GoalContinuous3P-v0): seconds, not minutes.
#   tools/dev_build.sh NAME [extra nvcc flags...]   -> spacegym_b200/libsg_dev_NAME.so (+ /tmp/sg_dev_NAME.ptxas)
# Compare candidates on one GPU box with tools/dev_eval.sh.
set -e
name="$1"; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v \
  -DSG_DEV_ONLY_GOAL3P "$@" -shared spacegym_b200/csrc/sg_engine.cu -o spacegym_b200/libsg_dev_${name}.so \
  2> /tmp/sg_dev_${name}.ptxas
grep -A1 "hot_kernel\|cont_kernel\|tail2_kernel" /tmp/sg_dev_${name}.ptxas | grep -B1 "Used" | grep "Used\|Compiling" | paste - - | sed 's/ptxas info    : //g' | c++filt | cut -c1-220 | grep -v "float" | head
