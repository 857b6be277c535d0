#!/bin/bash
# A/B build of the library with extra -D flags for sweep3d.cu: tools/build_ab.sh <name> "<flags>" -> umt_b200/ab/libumtsweep_<name>.so (select with UMT_LIB)
set -e
NAME=$1; FLAGS=$2
cd "$(dirname "$0")/../umt_b200/csrc"
make -s >/dev/null
mkdir -p ../ab build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 -Xptxas -v $FLAGS -c sweep3d.cu -o build/ab/sweep3d_$NAME.o 2> build/ab/sweep3d_$NAME.log
grep -A2 "plan_kernel" build/ab/sweep3d_$NAME.log | grep "registers\|spill" 
OBJS=$(ls build/*.o | grep -v sweep3d.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../ab/libumtsweep_$NAME.so $OBJS build/ab/sweep3d_$NAME.o -lcudart -ldl
echo built umt_b200/ab/libumtsweep_$NAME.so
