#!/bin/bash
# Build a tuning variant of the library: recompile only the listed translation units with extra flags and link them with
# the objects of the main build.   tools/build_variant.sh NAME "-DENF_GRAD_VECS=8" enf_chain_inst_f32_vec [more TUs]
set -e
NAME=$1; EXTRA=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/euclidiannormalizingflows.jl_b200/csrc
OBJ=$ROOT/build/obj_$NAME
mkdir -p $OBJ $ROOT/build/variants
cp $ROOT/build/obj/*.o $OBJ/
for tu in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function \
     -I$ROOT/include -I$CSRC --expt-relaxed-constexpr --compress-mode=size $EXTRA -c $CSRC/$tu.cu -o $OBJ/$tu.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/build/variants/libenf_$NAME.so $OBJ/*.o -lcudart -ldl
echo built $ROOT/build/variants/libenf_$NAME.so
