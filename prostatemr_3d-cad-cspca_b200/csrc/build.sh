#!/usr/bin/env bash
# Builds libm1b200.so (sm_100a only) in-tree: prostatemr_3d-cad-cspca_b200/lib/libm1b200.so
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../lib"
mkdir -p "$out" "$here/_obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr
       -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function -I"$here/../../include")
if [[ "${M1_PTXAS_V:-0}" == "1" ]]; then FLAGS+=(-Xptxas -v); fi
objs=()
pids=()
for src in "$here"/*.cu; do
  obj="$here/_obj/$(basename "${src%.cu}").o"
  objs+=("$obj")
  if [[ ! -f "$obj" || "$src" -nt "$obj" || "$here/common.cuh" -nt "$obj" || "$here/tc_common.cuh" -nt "$obj" || "$here/../../include/m1b200.h" -nt "$obj" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" &
    pids+=($!)
  fi
done
for pid in "${pids[@]:-}"; do [[ -n "$pid" ]] && wait "$pid"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libm1b200.so" "${objs[@]}" -lcudart
echo "built $out/libm1b200.so"
