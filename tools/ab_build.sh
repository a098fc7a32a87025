#!/bin/bash
# Build a variant of the library for A/B timing:  tools/ab_build.sh NAME -DSQ_SOMETHING=1 ...
# -> odam_b200/lib/ab/libodam_sq_NAME.so ; run with ODAM_SQ_LIB=$PWD/odam_b200/lib/ab/libodam_sq_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p odam_b200/lib/ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xptxas -regUsageLevel=10 -Xcompiler -fPIC -shared "$@" \
     -o odam_b200/lib/ab/libodam_sq_$name.so odam_b200/csrc/sq_kernels.cu
echo odam_b200/lib/ab/libodam_sq_$name.so
