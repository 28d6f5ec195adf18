#!/bin/bash
# usage: tools/variants_time.sh out.jsonl case... -- times every library under rasterizer_b200/variants/ and the product's own
out=$1; shift
: > $out
for lib in rasterizer_b200/librasterizer_b200.so rasterizer_b200/variants/lib_*.so; do
  ORZ_LIB=$PWD/$lib timeout 300 python tools/step_time.py "$@" >> $out 2>> ${out%.jsonl}.err
done
python - "$out" <<'P'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print(f"{d['lib']:34s}", "  ".join(f"{k}={v['ms_median']:.3f}ms[{v['vis_checksum'] % 100000},{v.get('launches_per_step',0):.0f}]" for k, v in d.items() if k != 'lib'))
P
