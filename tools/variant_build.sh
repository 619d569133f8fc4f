#!/bin/bash
# Build an experimental variant of ONE kernel source with extra -D flags and link it with the objects of the
# regular build:  tools/variant_build.sh <name> <source.cu> -DFOO -DBAR   ->  build/variants/<name>.so
# Run it with AASIST_B200_LIB=build/variants/<name>.so (timing experiments only; never the product path).
set -e
NAME=$1; SRC=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/aasist_b200/csrc
mkdir -p $ROOT/build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --expt-relaxed-constexpr -Xptxas -v "$@" -c $CS/$SRC -o $ROOT/build/variants/$NAME.o 2>&1 | grep -A2 "$(basename $SRC .cu)_kernel\|block0_tc_kernel\|conv_tc_kernel\|graph_kernel" | grep "spill\|Used" | head -4
OBJS=$(ls $CS/*.o | grep -v "/$(basename $SRC .cu).o")
nvcc -shared -o $ROOT/build/variants/$NAME.so $OBJS $ROOT/build/variants/$NAME.o -cudart static -Xlinker --exclude-libs,ALL
echo built build/variants/$NAME.so
