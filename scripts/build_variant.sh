#!/bin/bash
# Builds a tuning variant of the library next to the default one: libspb200_<name>.so with extra -D flags.
# Usage: scripts/build_variant.sh <name> [-DSPB_X=1 ...]     (variants are selected at run time with SPB200_LIB=...)
set -e
NAME=$1; shift
cd "$(dirname "$0")/../super_primitive_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -ftz=true -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -shared \
    -o libspb200_$NAME.so spb_align.cu spb_geom.cu spb_solve.cu spb_reinit.cu spb_window.cu spb_ingest.cu spb_fill.cu
echo "built libspb200_$NAME.so ($*)"
