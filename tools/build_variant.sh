#!/bin/bash
# Build a tuning variant of liblqgk.so: recompiles ONE dimension tuple with extra -D flags and links it with the
# other (already built) objects.  Usage: tools/build_variant.sh <tag> "<x b u y d>" [-DFLAG ...]
# -> lqg_b200/csrc/variants/liblqgk_<tag>.so  (use with LQGK_LIB_PATH=...; experiments only)
set -e
tag=$1; read x b u y d <<< "$2"; shift 2
cd "$(dirname "$0")/../lqg_b200/csrc"
mkdir -p variants
big=""; if [ $((x + b)) -gt 12 ]; then big="-DLQGK_BIG"; fi
obj=variants/inst_${tag}.o
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" $big \
  -DLQGK_INST_X=$x -DLQGK_INST_B=$b -DLQGK_INST_U=$u -DLQGK_INST_Y=$y -DLQGK_INST_D=$d -c lqgk_inst.cu -o $obj
others=$(ls lqgk_inst_*.o | grep -v "lqgk_inst_${x}_${b}_${u}_${y}_${d}.o")
/usr/local/cuda/bin/nvcc -shared -o variants/liblqgk_${tag}.so lqgk_api.o $obj $others -gencode arch=compute_100a,code=sm_100a
echo built variants/liblqgk_${tag}.so
