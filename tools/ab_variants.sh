#!/bin/bash
# A/B timing of prebuilt library variants (psi-release_b200/lib/variants/*.so): value, graph iteration time and the
# eager per-kernel times of bench.py for each; restores base.so at the end.
cd "$(dirname "$0")/.."
for f in psi-release_b200/lib/variants/*.so; do
  n=$(basename $f .so)
  cp $f psi-release_b200/lib/libpsi_b200.so
  timeout 120 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); r=d['roofline']
print('$n', round(d['value'],1), round(r['iteration_ms_graph']*1e3,1), round(r['launch_ms']*1e3,1), {k: round(v['launch_ms']*1e3,1) for k,v in r['others'].items()})"
done
cp psi-release_b200/lib/variants/base.so psi-release_b200/lib/libpsi_b200.so
