#!/bin/bash
# Build a tuning variant of libdust_b200.so for tools/ab.sh:  tools/mkvar.sh <name> "<extra nvcc flags>"
set -e
cd "$(dirname "$0")/../dust_b200/csrc"
mkdir -p ../lib/var
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall -shared $2 -o ../lib/var/$1.so dust_abi.cu
echo "built lib/var/$1.so ($2)"
