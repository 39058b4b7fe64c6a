#!/usr/bin/env bash
# Builds libsimplex_b200.so IN-TREE for sm_100a (B200). nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libsimplex_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC,-O3,-Wall,-Wno-unused-function -Xptxas -v \
    --shared -o "${OUT}" "${HERE}/sb200_abi.cu" ${SB200_EXTRA_NVCC_FLAGS:-} 2> "${HERE}/../build_ptxas.log" || {
    cat "${HERE}/../build_ptxas.log" >&2; exit 1; }
echo "built ${OUT}"
