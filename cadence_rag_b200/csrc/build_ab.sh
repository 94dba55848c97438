#!/bin/sh
# A/B build: the same library with extra compiler flags, as build/ab/libcadence_dense_<name>.so
# (load it with CADENCE_DENSE_LIB=...).  Usage: sh build_ab.sh <name> "<extra nvcc flags>" [files to recompile...]
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
NAME=$1; EXTRA=$2; shift 2
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OBJ="$HERE/../../build/obj"; AB="$HERE/../../build/ab"; mkdir -p "$AB/obj_$NAME"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++"
objs=""
for f in abi store filter exact_scan rrf topk_merge gemm_topk tech_lane hybrid peer; do
    use="$OBJ/$f.o"
    for g in "$@"; do
        if [ "$g" = "$f" ]; then
            extra=""; [ "$f" = "rrf" ] && extra="-fmad=false"
            $NVCC $FLAGS $extra $EXTRA -c "$HERE/$f.cu" -o "$AB/obj_$NAME/$f.o"
            use="$AB/obj_$NAME/$f.o"
        fi
    done
    objs="$objs $use"
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$AB/libcadence_dense_$NAME.so" $objs -ccbin /usr/bin/g++
echo "built $AB/libcadence_dense_$NAME.so"
