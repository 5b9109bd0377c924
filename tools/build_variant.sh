#!/bin/bash
# Builds a kernel-variant copy of libldpc_b200.so for A/B measurements: tools/build_variant.sh NAME "-DRES_DUP=0 ..."
# -> build/libldpc_b200_NAME.so (select it with LDPC_B200_SO=build/libldpc_b200_NAME.so; travels to the GPU box).
# Only the translation units of the RESIDENT kernels see the RES_* switches; the other objects are taken from the
# regular build in csrc/ (run make there first).  ALL=1 recompiles everything with the flags.
set -e
NAME=$1; FLAGS=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/ecc-ldpc_b200/csrc; OBJ=$ROOT/build/obj_$NAME
mkdir -p $OBJ
NV="/usr/local/cuda/bin/nvcc $FLAGS -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall -Xptxas -v --fmad=true -ccbin g++"
RES="resident_engine resident_trace resident_latency"
REST="api stream_engine onchip_engine channel host_ecc multi_gpu mailbox"
if [ -n "$ALL" ]; then RES="$RES $REST"; REST=""; fi
pids=()
for f in $RES; do
  ( cd $SRC && $NV -c $f.cu -o $OBJ/$f.o 2> $OBJ/$f.log ) & pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
for f in $REST; do cp $SRC/$f.o $OBJ/$f.o; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/build/libldpc_b200_$NAME.so $OBJ/*.o -cudart static -ldl
echo built build/libldpc_b200_$NAME.so
