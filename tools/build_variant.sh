#!/bin/bash
# Build a tuning variant of libcph_b200.so next to the product library:
#   tools/build_variant.sh TAG "-DCPH_EVAL_MAXNREG=72 -DCPH_REFINE=3"
# -> constant_ph_b200/csrc/variants/libcph_b200_TAG.so; select it with CPH_B200_LIB=<path> (capi.py).
set -e
TAG=$1; shift
DIR=$(cd "$(dirname "$0")/../constant_ph_b200/csrc" && pwd)
mkdir -p "$DIR/variants"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a $* -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-fopenmp --expt-relaxed-constexpr -shared -o "$DIR/variants/libcph_b200_$TAG.so" \
  "$DIR"/cph_api.cu "$DIR"/neigh.cu "$DIR"/pair.cu "$DIR"/sites.cu "$DIR"/comm.cu "$DIR"/bonded.cu "$DIR"/ljstates.cu "$DIR"/microbench.cu \
  -lcudart -ldl -lgomp
echo "$DIR/variants/libcph_b200_$TAG.so"
