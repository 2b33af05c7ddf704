#!/bin/bash
# build_variant.sh NAME [extra nvcc flags]: compile only the classic rule-set + misc TUs of the working tree
# into gpurun_out/../variants/libsnk_NAME.so for A/B timing through SNK_LIB (classic rules only).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants/_o_$name
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $*"
C=snakes_b200/csrc
nvcc $F -c -o variants/_o_$name/api.o $C/snk_api.cu &
nvcc $F -DSNK_TU=0 -DLANE_COMBOS_OVERRIDE -c -o variants/_o_$name/k0.o $C/snk_kernels.cu &
nvcc $F -DSNK_TU=1 -DLANE_COMBOS_OVERRIDE -c -o variants/_o_$name/k1.o $C/snk_kernels.cu &
nvcc $F -DSNK_TU=2 -DLANE_COMBOS_OVERRIDE -c -o variants/_o_$name/k2.o $C/snk_kernels.cu &
nvcc $F -DSNK_TU=3 -DLANE_COMBOS_OVERRIDE -c -o variants/_o_$name/k3.o $C/snk_kernels.cu &
wait
for o in api k0 k1 k2 k3; do [ -f variants/_o_$name/$o.o ] || { echo "FAILED: $o did not compile"; rm -rf variants/_o_$name; exit 1; }; done
nvcc --shared -gencode arch=compute_100a,code=sm_100a -o variants/libsnk_$name.so variants/_o_$name/*.o
rm -rf variants/_o_$name
echo built variants/libsnk_$name.so
