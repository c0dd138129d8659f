#!/bin/bash
# GPU side: A/B of an environment switch on the same box:  tools/ab_env.sh "<bench args>" VAR=a VAR=b [VAR=a VAR=b ...]   ("-" = unset)
args=$1; shift
for kv in "$@"; do
  echo "=== $kv"
  ( [ "$kv" != "-" ] && export "$kv"; timeout 600 python bench.py --no-latency $args 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
r=d['roofline']
print('value %.0f  ms %.3f  e2e %.0f  instrumented_ms %.3f  frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['instrumented_ms_per_step'], r['frac']))
for row in r['per_class'][:${ROWS:-12}]: print('   %8.1f us  %s' % (row[2], row[0]))
" )
done
