#!/bin/bash
set -e
cd "$(dirname "$0")"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -ccbin /usr/bin/g++ -o microbench microbench.cu
