#!/bin/bash
# Build a variant of the library with extra -D flags on the sweep-kernel TU (kernel experiments):
#   tools/build_variant.sh narrow "-DIFX_PPE_NC2_WIDE=1"   ->  tools/_bin/lib_narrow.so
# Run it with IFX_LIBRARY=tools/_bin/lib_NAME.so python bench.py ...
set -e
cd "$(dirname "$0")/.."
python immerseflow_b200/build.py > /dev/null
mkdir -p tools/_bin
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -fmad=false $2 -x cu -c immerseflow_b200/csrc/kernels_v4.cu -o /tmp/v4_$1.o
OBJS=$(ls immerseflow_b200/_build/*.o | grep -v kernels_v4.o)
nvcc $ARCH -shared -o tools/_bin/lib_$1.so $OBJS /tmp/v4_$1.o -Xcompiler -fPIC
echo tools/_bin/lib_$1.so
