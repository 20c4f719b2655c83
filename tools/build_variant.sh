#!/bin/bash
# A/B build of libspeecht5_b200.so for tuning runs: tools/build_variant.sh NAME FILE.cu -DMACRO ...
# recompiles FILE.cu with the extra flags, links it with the other objects of the regular build into
# speecht5_b200/lib/variant_NAME.so (select it with ST5_LIB=... at run time). The regular library must be built first.
set -eu
NAME=$1; SRC=$2; shift 2
cd "$(dirname "$0")/.."
OBJ=speecht5_b200/lib/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -cudart static"
/usr/local/cuda/bin/nvcc $FLAGS "$@" -c speecht5_b200/csrc/$SRC -o $OBJ/variant_${NAME}.o
OTHERS=$(ls $OBJ/*.o | grep -v "variant_" | grep -v "/${SRC%.cu}.o")
/usr/local/cuda/bin/nvcc $FLAGS -shared $OTHERS $OBJ/variant_${NAME}.o -o speecht5_b200/lib/variant_${NAME}.so
echo speecht5_b200/lib/variant_${NAME}.so
