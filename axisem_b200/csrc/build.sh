#!/bin/bash
# Builds the product library in-tree for sm_100a (and the -fmad=false twin used by the
# bit-exact parity tests).  Called by __graft_entry__.build().
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/.."
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
COMMON=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --ftz=true
        -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xlinker -Bsymbolic -Xlinker "--version-script=$HERE/axb.map" -shared -Xptxas -v)
"$NVCC" "${COMMON[@]}" -o "$OUT/libaxisem_b200.so" "$HERE/axb_api.cu" 2> "$HERE/ptxas_fast.log" || { cat "$HERE/ptxas_fast.log"; exit 1; }
"$NVCC" "${COMMON[@]}" -fmad=false -DAXB_STRICT=1 -o "$OUT/libaxisem_b200_strict.so" "$HERE/axb_api.cu" 2> "$HERE/ptxas_strict.log" || { cat "$HERE/ptxas_strict.log"; exit 1; }
# keep the tracked -Xptxas -v logs deterministic (registers / spills / shared memory per kernel)
sed -i '/Compile time = /d' "$HERE/ptxas_fast.log" "$HERE/ptxas_strict.log"
echo "built $OUT/libaxisem_b200.so and libaxisem_b200_strict.so"
