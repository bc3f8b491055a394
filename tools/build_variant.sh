#!/bin/bash
# build_variant.sh NAME DT_SOURCE [extra nvcc flags]: links a copy of the library whose dt.o is compiled from DT_SOURCE with extra
# flags -> build/variants/libpbd_b200_NAME.so (A/B runs on the GPU box: PBD_B200_LIB=build/variants/libpbd_b200_NAME.so)
set -e
NAME=$1; SRC=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/partsbaseddetector_b200/csrc
mkdir -p $ROOT/build/variants
make -C $C -j8 > /dev/null
cp $SRC $C/_variant_dt.cu
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --fmad=false "$@" -Xptxas -v -c $C/_variant_dt.cu -o $ROOT/build/variants/dt_$NAME.o 2>&1 | grep -E "Used" | head -1
rm -f $C/_variant_dt.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/build/variants/libpbd_b200_$NAME.so $C/pyramid.o $C/hog.o $C/response.o $C/response_tc.o \
  $ROOT/build/variants/dt_$NAME.o $C/backtrack.o $C/nms.o $C/engine.o $C/model.o $C/matfile.o $C/ingest.o $C/abi.o -lz -Xlinker --no-undefined
echo built $NAME
