#!/usr/bin/env bash
# Build libgcnb200.so (sm_100a only) next to the Python package.  No GPU needed: nvcc cross-compiles.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libgcnb200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
srcs=("${here}"/*.cu)
"${NVCC}" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC,-O3,-fvisibility=hidden -shared -cudart static \
  ${GCNB_NVCC_EXTRA:-} -o "${out}" "${srcs[@]}"
echo "built ${out}"
