#!/bin/bash
# builds the microbenchmarks next to their sources (binaries are git-ignored; they travel with gpurun)
set -e
cd "$(dirname "$0")/../.."
python -c "import websplat_b200 as ws; ws.build_library(force=False)"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I web-splat_b200/csrc -I include \
     profiles/microbench/sort_vs_cub.cu web-splat_b200/build/radix_sort.o -o profiles/microbench/sort_vs_cub
