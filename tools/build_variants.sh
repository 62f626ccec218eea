#!/bin/bash
# Dev tool: builds librerevst_b200_<name>.so variants of the tensor-core convolution with different -D switches
# (kernel experiments; select one at run time with RRV_LIB_PATH).  usage: tools/build_variants.sh name "-DX=1 -DY=0" ...
set -e
cd "$(dirname "$0")/../rerevst-code_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
pids=()
args=("$@")
for ((i = 0; i < ${#args[@]}; i += 2)); do
  name=${args[i]}; defs=${args[i+1]}
  ( mkdir -p build/var_$name
    $NVCC $FLAGS $defs -Xptxas -v -c conv_tc.cu -o build/var_$name/conv_tc.o 2> build/var_$name/ptxas.log
    objs=$(ls build/*.o | grep -v conv_tc.o)
    $NVCC -shared -o librerevst_b200_$name.so build/var_$name/conv_tc.o $objs -lcudart ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
ls -la librerevst_b200_*.so
