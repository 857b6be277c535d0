#!/bin/bash
# A/B build of the library with extra -D flags for one source file (default sweep3d.cu):
#   tools/build_ab.sh <name> "<flags>" [file.cu] -> umt_b200/ab/libumtsweep_<name>.so (select with UMT_LIB)
set -e
NAME=$1; FLAGS=$2; SRC=${3:-sweep3d.cu}; BASE=${SRC%.cu}
cd "$(dirname "$0")/../umt_b200/csrc"
make -s >/dev/null
mkdir -p ../ab build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 -Xptxas -v $FLAGS -c $SRC -o build/ab/${BASE}_$NAME.o 2> build/ab/${BASE}_$NAME.log
grep "registers\|spill" build/ab/${BASE}_$NAME.log | head -12
OBJS=$(ls build/*.o | grep -v "build/$BASE.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../ab/libumtsweep_$NAME.so $OBJS build/ab/${BASE}_$NAME.o -lcudart -ldl
echo built umt_b200/ab/libumtsweep_$NAME.so
