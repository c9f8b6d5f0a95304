#!/bin/bash
# compute ceiling of the packed p-c pair evaluation (build first:
#   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -Iinclude \
#        -o tools/bin/pc_pair_ceiling tools/pc_pair_ceiling.cu)
tools/bin/pc_pair_ceiling
