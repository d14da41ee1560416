#!/bin/bash
# Builds libpn2b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
OUT=../libpn2b200.so
SRCS="api.cu fps.cu ballquery.cu ballgrid.cu group.cu threenn.cu linear_simt.cu bn.cu linear_tc.cu bwd_fused.cu sa_fused.cu head.cu vote.cu slicer.cu optim.cu"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include ${PN2_NVCC_EXTRA}"
mkdir -p build
objs=""
pids=""
for s in $SRCS; do
  o=build/${s%.cu}.o
  objs="$objs $o"
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ tc_common.cuh -nt "$o" ] || [ ../../include/pn2b200.h -nt "$o" ]; then
    $NVCC $FLAGS -c "$s" -o "$o" &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $objs -lcudart
echo "built $(readlink -f $OUT)"
