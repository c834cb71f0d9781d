#!/bin/bash
# A/B library with different compile-time knobs in ONE source file, linked against the objects of the normal build:
#   bash tools/build_variant.sh <tag> <file.cu> "-DKNOB=1 ..."   ->  mimic3_b200/libm3b200_<tag>.so  (select with M3B200_LIBRARY)
set -e
TAG=$1; SRC=$2; DEFS=$3
python -m mimic3_b200.build > /dev/null
mkdir -p build/obj_var
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I include -I mimic3_b200/csrc $DEFS \
  -c mimic3_b200/csrc/$SRC -o build/obj_var/${TAG}.o
OBJS=$(ls build/obj/*.o | grep -v "/$SRC.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o mimic3_b200/libm3b200_${TAG}.so $OBJS build/obj_var/${TAG}.o -Xlinker --no-undefined
echo mimic3_b200/libm3b200_${TAG}.so
