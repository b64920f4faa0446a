#!/bin/bash
# Builds variants/libnls_<name>.so with extra nvcc flags for kernels_lines3d.cu only (the other objects are compiled once).
# Usage: tools/build_variant.sh name [-DL3_PAIR_UNROLL=2 ...]; run with NLS_B200_LIB=variants/libnls_<name>.so
set -e
cd "$(dirname "$0")/../nonlinearsolids.jl_b200/csrc"
name=$1; shift
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
O=../../variants/obj; mkdir -p $O
for s in nls_api.cu kernels_assembly.cu kernels_gather.cu kernels_twophase.cu kernels_mg.cu kernels_direct.cu kernels_j2.cu kernels_krylov.cu kernels_dynamics.cu kernels_p2p.cu host_tables.cpp; do
  [ $O/$s.o -nt $s ] || nvcc $F -c $s -o $O/$s.o &
done
nvcc $F "$@" -Xptxas -v -c kernels_lines3d.cu -o $O/lines_$name.o 2>&1 | grep -A2 "k_asm_q1_3d_linesILb1ELb1ELb0ELb0" | grep "spill\|Used"
wait
nvcc -shared -o ../../variants/libnls_$name.so $O/lines_$name.o $O/*.cu.o $O/host_tables.cpp.o -ldl
