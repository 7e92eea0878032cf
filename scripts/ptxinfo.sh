#!/bin/bash
# Registers / spills of the batched fused kernels for a set of -D flags (offline, no GPU):
#   scripts/ptxinfo.sh -DSPB_GROUP=4 -DSPB_TOUCH=1
cd "$(dirname "$0")/../super_primitive_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -ftz=true -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" -c spb_align.cu -o /tmp/a_$$.o 2>&1 | grep -A2 "k_align_globalILi[01]ELi6ELb0" | grep -v "^--" | paste - - - | sed -e 's/ptxas info    ://g' -e 's/.*k_align_globalILi\([01]\).*Function properties for [^ ]*/mode\1:/' -e 's/, used 1 barriers.*//' | tr -s ' \t' ' '
rm -f /tmp/a_$$.o
