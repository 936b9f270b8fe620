#!/bin/bash
# Instrumented build of the library (-DMDBN_SKINNY_DEBUG: the MDBN_SKINNY_DEBUG=<bits> skip flags of the skinny kernel,
# timing experiments only — results are wrong).  Use with MDBN_B200_LIB=mdbn_b200/csrc/libmdbn_b200_dbg.so
set -e
cd "$(dirname "$0")/../mdbn_b200/csrc"
NCCL_INC=$(python -c "import os,sys;print(next((os.path.join(b,'nvidia','nccl','include') for b in sys.path if os.path.exists(os.path.join(b,'nvidia','nccl','include','nccl.h'))),'.'))")
mkdir -p build_dbg
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DMDBN_SKINNY_DEBUG -I "$NCCL_INC" -c $f -o build_dbg/${f%.cu}.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a build_dbg/*.o -o libmdbn_b200_dbg.so -ldl
echo built libmdbn_b200_dbg.so
