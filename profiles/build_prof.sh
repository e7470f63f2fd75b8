#!/bin/bash
# Builds lib/libvffem_b200_prof.so: the product library with the cycle probes of the assembly
# kernels compiled in (-DVF_PIPE_PROF).  Run after __graft_entry__.build().
set -e
cd "$(dirname "$0")/../vf-fem_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
  -DVF_PIPE_PROF -c -o build/assembly_prof.o csrc/assembly.cu
nvcc --shared -o lib/libvffem_b200_prof.so build/assembly_prof.o build/krylov.o build/member.o build/vffem_b200.o
