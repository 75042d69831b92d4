#!/usr/bin/env bash
# Debug build with the clock64 trace points of k_cheb_fwd_umma compiled in (-DGCNB_TRACE), next to the release library:
#   tools/build_trace.sh && GCNB_LIB_PATH=gcn_fmri_decoding_b200/csrc/build_trace/libgcnb200_trace.so python tools/umma_trace.py f1
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/../gcn_fmri_decoding_b200/csrc" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
obj="${here}/build_trace"
mkdir -p "${obj}"
flags=(-std=c++17 -O3 -lineinfo -DGCNB_TRACE -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3,-fvisibility=hidden)
pids=()
objs=()
for src in "${here}"/*.cu; do
  o="${obj}/$(basename "${src%.cu}").o"
  objs+=("${o}")
  if [[ ! -f "${o}" || "${src}" -nt "${o}" || "$(ls -t "${here}"/*.cuh | head -1)" -nt "${o}" ]]; then
    "${NVCC}" "${flags[@]}" -c "${src}" -o "${o}" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "${p}" ]] && wait "${p}"; done
"${NVCC}" -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o "${obj}/libgcnb200_trace.so" "${objs[@]}"
echo "built ${obj}/libgcnb200_trace.so"
