#!/bin/bash
# Builds librerevst_b200.so for sm_100a in-tree (no torch dependency: plain CUDA runtime).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
# (-diag-suppress 128: "loop is not reachable" in the staged-only instantiations of epilogue_chunk, by construction)
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress 128 -Wno-deprecated-gpu-targets ${RRV_DEFS}"
mkdir -p build
pids=()
for f in capi conv_ffma conv_tc first_layer pointwise stats warp; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ rrv_common.cuh -nt build/$f.o ] || [ ../../include/rerevst_b200.h -nt build/$f.o ] || { [ $f = conv_tc ] && [ -f tc_ptx.cuh ] && [ tc_ptx.cuh -nt build/$f.o ]; }; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -Wno-deprecated-gpu-targets -shared -o librerevst_b200.so build/*.o -lcudart
echo "built $(pwd)/librerevst_b200.so"
# the optional video-output side library (include/rerevst_b200_io.h): host code + nvJPEG; nothing on the stylization path loads it
if [ ! -f librerevst_b200_io.so ] || [ mjpg_io.cpp -nt librerevst_b200_io.so ] || [ ../../include/rerevst_b200_io.h -nt librerevst_b200_io.so ]; then
  # (not fatal: a toolkit without nvJPEG still builds the core library; video_io raises when its library is missing)
  $NVCC -Wno-deprecated-gpu-targets -O2 -std=c++17 -Xcompiler -fPIC -shared mjpg_io.cpp -o librerevst_b200_io.so -lnvjpeg -lcudart \
    || echo "warning: librerevst_b200_io.so not built (nvJPEG missing?)"
fi
[ -f librerevst_b200_io.so ] && echo "built $(pwd)/librerevst_b200_io.so"
true
