#!/bin/sh
# ptxas_info.sh [extra nvcc flags] : registers / spills of every kernel of libdla_b200 (compile only, no GPU needed)
cd "$(dirname "$0")/../gpy_dla_detection_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 "$@" -Xptxas -v -c -o /tmp/dla_ptxas_check.o dla_b200.cu 2>&1 | grep -i " error\|Compiling entry\|Used\|spill" | sed 's/ptxas info    : //; s/Compiling entry function //; s/ for .sm_100a.//'
