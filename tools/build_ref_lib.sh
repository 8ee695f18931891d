#!/bin/bash
# Build libflair_b200.so of another commit for A/B timing on the same GPU box:
#   tools/build_ref_lib.sh <git-ref>   ->  flair_b200/_ab/lib_<ref>.so   (git-ignored, travels with gpurun)
#   FLAIR_B200_LIB=flair_b200/_ab/lib_<ref>.so python tests/gpu_probes/conv_graph.py ...
set -euo pipefail
ref=$1
root=$(git rev-parse --show-toplevel)
tmp=$(mktemp -d)
git -C "$root" archive "$ref" flair_b200/csrc include | tar -x -C "$tmp"
mkdir -p "$root/flair_b200/_ab"
objs=()
for f in "$tmp"/flair_b200/csrc/*.cu; do
  o="$tmp/$(basename "$f" .cu).o"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -c "$f" -o "$o" &
  objs+=("$o")
done
wait
/usr/local/cuda/bin/nvcc -shared -o "$root/flair_b200/_ab/lib_${ref}.so" "${objs[@]}" -lcudart
rm -rf "$tmp"
echo "$root/flair_b200/_ab/lib_${ref}.so"
