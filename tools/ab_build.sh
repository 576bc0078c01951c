#!/bin/bash
# usage: tools/ab_build.sh <name> [-DMACRO ...]   -> tools/ab/lib<name>.so : the library with csrc/attention.cu rebuilt with extra
# macros (A/B experiments; compared in one process by tools/attn_ab.py)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/ab tools/ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include "$@" \
  -c fastdm_b200/csrc/attention.cu -o build/ab/attention_$name.o
objs=$(ls build/obj/*.o | grep -v "/attention.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o tools/ab/lib$name.so $objs build/ab/attention_$name.o
echo "built tools/ab/lib$name.so"
