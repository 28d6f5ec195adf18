#!/bin/bash
# usage: tools/variant_bench.sh lib1.so lib2.so ...  -- swaps the product library and times the view-batch kernels
cp rasterizer_b200/librasterizer_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp "$v" rasterizer_b200/librasterizer_b200.so
  echo "== $v"; timeout 300 python tools/experiments.py 2>&1 | grep -E "t2_gw[48]_(full|noqueries)"
done
cp /tmp/lib_orig.so rasterizer_b200/librasterizer_b200.so
echo "== current"; timeout 300 python tools/experiments.py 2>&1 | grep -E "t2_gw[48]_(full|noqueries)"
