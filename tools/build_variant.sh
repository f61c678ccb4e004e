#!/usr/bin/env bash
# tools/build_variant.sh NAME SRC.cu [extra nvcc flags]: link tools/variants/libNAME.so = in-tree objects with SRC.cu's object replaced
set -euo pipefail
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
base=$(basename "${src%.cu}")
mkdir -p tools/variants build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --expt-relaxed-constexpr -cudart static -Isnuffy_b200/csrc "$@" -c "$src" -o build/variants/${base}_$name.o
objs=$(ls build/*.o | grep -v "/$base.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xlinker --exclude-libs,ALL -o tools/variants/lib$name.so $objs build/variants/${base}_$name.o
echo "built tools/variants/lib$name.so"
