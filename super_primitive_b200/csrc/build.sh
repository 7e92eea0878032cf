#!/bin/bash
# Builds libspb200.so (sm_100a) next to the sources.  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -ftz=true -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
$NVCC $FLAGS -shared -o libspb200.so spb_align.cu spb_geom.cu spb_solve.cu spb_reinit.cu spb_window.cu spb_ingest.cu "$@"
