#!/bin/bash
# GPU side: A/B of two builds of the library on the same box:  tools/ab_bench.sh "<bench args>" libA.so libB.so [libA.so libB.so ...]
# (SE_LIB_OUT=... python -m sketchedit_b200.build --force builds a variant; SE_B200_LIB selects it at load time)
args=$1; shift
for lib in "$@"; do
  if [ "$lib" = default ]; then unset SE_B200_LIB; else export SE_B200_LIB=$PWD/sketchedit_b200/$lib; fi
  echo "=== $lib"
  timeout 600 python bench.py --no-latency $args 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
r=d['roofline']
print('value %.0f  ms %.3f  e2e %.0f  instrumented_ms %.3f  frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['instrumented_ms_per_step'], r['frac']))
for row in r['per_class'][:${ROWS:-12}]: print('   %8.1f us  %s' % (row[2], row[0]))
"
done
