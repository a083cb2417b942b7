#!/bin/bash
# development aid: build an A/B variant of the library with extra -D flags for list_decode.cu only
#   tools/variant.sh NAME [-DFLAG ...]   ->  build/variants/libpolargpu_NAME.so   (use with POLARGPU_LIB=...)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/build/variants
cd $root/polardecoding_b200/csrc
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DPOLAR_DEV_CASES "$@" -c -o $root/build/variants/list_$name.o list_decode.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/build/variants/libpolargpu_$name.so api.o channel.o $root/build/variants/list_$name.o bp_decode.o bp_decode_h2.o hostlogic.o -ldl
echo built $root/build/variants/libpolargpu_$name.so
