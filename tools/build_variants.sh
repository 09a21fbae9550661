#!/bin/bash
# Experiment builds of libwassgpu.so with other rows-per-band / ring-depth settings of the fused sweeps (K == 1):
#   tools/build_variants.sh "15 6" "11 8" "7 5"    -> wass_b200/variants/libwassgpu_r15n6.so ...
# Select one with WSG_LIB=<path> (wass_b200/capi.py).  The default build is not touched.
set -e
cd "$(dirname "$0")/.."
python -m wass_b200.build > /dev/null
OBJ=wass_b200/build
mkdir -p wass_b200/variants
for v in "$@"; do
  set -- $v; R=$1; NS=$2; shift 2; EXTRA="$*"
  name=r${R}n${NS}$(echo "$EXTRA" | tr -d ' =-' | tr 'A-Z' 'a-z' | sed 's/dwsg_//g')
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ \
      -DWSG_SW_ROWS1=$R -DWSG_SW_NS1=$NS $EXTRA -c wass_b200/csrc/sweep_kernels.cu -o $OBJ/sweep_kernels_$name.o &&
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o wass_b200/variants/libwassgpu_$name.so \
      $(ls $OBJ/*.o | grep -v "sweep_kernels") $OBJ/sweep_kernels_$name.o -lcudart_static -lpthread -ldl -lrt &&
    echo built wass_b200/variants/libwassgpu_$name.so ) &
done
wait
