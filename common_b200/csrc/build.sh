#!/bin/bash
# Builds libmscope_b200.so in-tree for sm_100a (the only target).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v \
  -shared -o libmscope_b200.so mscope_b200.cu 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "Xptxas" | head -20 || true
echo "built $(pwd)/libmscope_b200.so"
