#!/bin/sh
# Build libradiofm_b200.so (in-tree) for sm_100a.  -fmad=false: see rfm_math.cuh.
set -e
cd "$(dirname "$0")"
OUT=../libradiofm_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVFLAGS="$ARCH -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off"
mkdir -p build
$NVCC $NVFLAGS -c rfm_kernels.cu -o build/rfm_kernels.o
$NVCC $NVFLAGS -c rfm_api.cu -o build/rfm_api.o
$NVCC $NVFLAGS -c rfm_probe.cu -o build/rfm_probe.o
$NVCC $NVFLAGS -c rfm_freqshift.cu -o build/rfm_freqshift.o
$NVCC $NVFLAGS -c rfm_downconvert.cu -o build/rfm_downconvert.o
$NVCC $NVFLAGS -c rfm_primitives.cu -o build/rfm_primitives.o
g++ -O2 -std=c++17 -fPIC -fvisibility=hidden -ffp-contract=off -c rfm_plan.cpp -o build/rfm_plan.o
g++ -O2 -std=c++17 -fPIC -fvisibility=hidden -ffp-contract=off -c rfm_rdssync.cpp -o build/rfm_rdssync.o
g++ -O2 -std=c++17 -fPIC -fvisibility=hidden -ffp-contract=off -c rfm_rdsgroup.cpp -o build/rfm_rdsgroup.o
g++ -O2 -std=c++17 -fPIC -fvisibility=hidden -ffp-contract=off -I/usr/local/cuda/include -c rfm_demux.cpp -o build/rfm_demux.o
$NVCC $ARCH -shared -o $OUT build/rfm_kernels.o build/rfm_api.o build/rfm_probe.o build/rfm_freqshift.o build/rfm_downconvert.o build/rfm_primitives.o build/rfm_plan.o build/rfm_rdssync.o build/rfm_rdsgroup.o build/rfm_demux.o -lcudart_static -lpthread -ldl -lrt
echo "built $(readlink -f $OUT)"
