#!/usr/bin/env bash
# Builds libtimegan_b200.so for sm_100a next to the Python host package (in-tree, travels with gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libtimegan_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall -Xptxas -v
       --expt-relaxed-constexpr -I"$HERE" -I"$HERE/../../include")
SRCS=(api proj_bf16 gru_fwd gru_bwd gru_jvp gru_bigh gru_cluster gemm_ffma proj_tcgen05 wgrad_tcgen05 wgrad_gru_tcgen05 losses optim rng peer_allreduce eval_stats head)
mkdir -p "$HERE/build"
pids=()
for s in "${SRCS[@]}"; do
  if [[ ! -f "$HERE/build/$s.o" || "$HERE/$s.cu" -nt "$HERE/build/$s.o" || -n "$(find "$HERE" "$HERE/../../include" -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$HERE/build/$s.o" 2>/dev/null)" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$HERE/$s.cu" -o "$HERE/build/$s.o" > "$HERE/build/$s.log" 2>&1 || { cat "$HERE/build/$s.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
# the shared CUDA runtime (the one torch has already loaded): the library then carries no copy of libcudart
"$NVCC" -shared -o "$OUT" $(printf "$HERE/build/%s.o " "${SRCS[@]}") --cudart shared -ldl -lrt -lpthread
echo "built $OUT"
