#!/bin/sh
# Builds cadence_rag_b200/libcadence_dense.so for sm_100a (B200) only.
# Usage: sh cadence_rag_b200/csrc/build.sh   (called by __graft_entry__.build())
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/../libcadence_dense.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OBJ="$HERE/../../build/obj"
mkdir -p "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++"
pids=""
for f in abi store filter exact_scan rrf topk_merge gemm_topk tech_lane hybrid peer; do
    extra=""
    # K5 must not contract fp64 add/div chains (bit-exact RRF)
    [ "$f" = "rrf" ] && extra="-fmad=false"
    if [ ! -f "$OBJ/$f.o" ] || [ "$HERE/$f.cu" -nt "$OBJ/$f.o" ] || [ "$HERE/common.cuh" -nt "$OBJ/$f.o" ] \
       || [ "$HERE/../../include/cadence_dense.h" -nt "$OBJ/$f.o" ]; then
        $NVCC $FLAGS $extra -c "$HERE/$f.cu" -o "$OBJ/$f.o" &
        pids="$pids $!"
    fi
done
for p in $pids; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$OBJ"/abi.o "$OBJ"/store.o "$OBJ"/filter.o "$OBJ"/exact_scan.o "$OBJ"/rrf.o \
    "$OBJ"/topk_merge.o "$OBJ"/gemm_topk.o "$OBJ"/tech_lane.o "$OBJ"/hybrid.o "$OBJ"/peer.o -ccbin /usr/bin/g++
echo "built $OUT"
