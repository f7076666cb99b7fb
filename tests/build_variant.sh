#!/bin/bash
# Developer tool: build libpdabm.so of a git revision (or of the working tree, rev = WORK) into build_ab/<name>.so
# for A/B timing in one GPU session (PDABM_LIB=build_ab/<name>.so python bench.py ...).
set -e
rev=$1; name=$2
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/build_ab"
tmp=$(mktemp -d)
mkdir -p "$tmp/pkg/csrc" "$tmp/include"
pkg=flamegpu2-prisoners-dilemma-abm_b200
if [ "$rev" = WORK ]; then
  cp "$root/$pkg/csrc/"* "$tmp/pkg/csrc/"; cp "$root/include/pdabm.h" "$tmp/include/"
else
  for f in $(git -C "$root" ls-tree --name-only "$rev" "$pkg/csrc/"); do git -C "$root" show "$rev:$f" > "$tmp/pkg/csrc/$(basename $f)"; done
  git -C "$root" show "$rev:include/pdabm.h" > "$tmp/include/pdabm.h"
fi
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -shared $EXTRA "$tmp/pkg/csrc/pdabm.cu" -o "$root/build_ab/$name.so"
rm -rf "$tmp"
echo "built build_ab/$name.so from $rev"
