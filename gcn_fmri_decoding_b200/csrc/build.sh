#!/usr/bin/env bash
# Build libgcnb200.so (sm_100a only) next to the Python package.  No GPU needed: nvcc cross-compiles.
# Sources compile in parallel into csrc/build/*.o (only when newer than the object), then link.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libgcnb200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
obj="${here}/build"
mkdir -p "${obj}"
flags=(-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3,-fvisibility=hidden ${GCNB_NVCC_EXTRA:-})
newest_hdr=$(ls -t "${here}"/*.cuh "${here}"/../../include/*.h | head -1)
pids=()
objs=()
for src in "${here}"/*.cu; do
  o="${obj}/$(basename "${src%.cu}").o"
  objs+=("${o}")
  if [[ ! -f "${o}" || "${src}" -nt "${o}" || "${newest_hdr}" -nt "${o}" ]]; then
    "${NVCC}" "${flags[@]}" -c "${src}" -o "${o}" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "${p}" ]] && wait "${p}"; done
"${NVCC}" -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o "${out}" "${objs[@]}"
echo "built ${out}"
