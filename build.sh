#!/usr/bin/env bash
# Builds snuffy_b200/libsnuffy_b200.so (sm_100a only) and the oracle's C pieces.  Usage: ./build.sh [-v]
set -euo pipefail
cd "$(dirname "$0")"
SRC=snuffy_b200/csrc
OUT=snuffy_b200/libsnuffy_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden
       --expt-relaxed-constexpr -cudart static)
[[ "${1:-}" == "-v" ]] && FLAGS+=(-Xptxas -v)
mkdir -p build
pids=()
for f in "$SRC"/*.cu; do
  o=build/$(basename "${f%.cu}").o
  if [[ ! -f "$o" || "$f" -nt "$o" || "$SRC/common.cuh" -nt "$o" || "$SRC/tc_ptx.cuh" -nt "$o" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" &
    pids+=($!)
  fi
done
fail=0
for p in "${pids[@]:-}"; do
  if [[ -n "$p" ]] && ! wait "$p"; then fail=1; fi
done
if [[ $fail -ne 0 ]]; then echo "build.sh: a compile step failed" >&2; exit 1; fi
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xlinker --exclude-libs,ALL -o "$OUT" build/*.o
echo "built $OUT"
