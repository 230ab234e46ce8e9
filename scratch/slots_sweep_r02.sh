#!/bin/bash
# Track-slot count sweep on the final build (two streams share the slots)
B="python bench.py --no-extra --no-cpu-baseline --steps 3 --warmup 3"
line() {
  python - "$1" <<'PY'
import json,sys
d=json.load(open('/tmp/line.json'))
print(sys.argv[1], '%.4g track-steps/s' % d['value'], '%.2f ms' % d['ms_per_step'], 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
}
for w in testem3 cms-scale; do
  for n in 524288 786432 1048576 1310720 1572864; do
    $B --workload $w --slots $n 2>/dev/null | tail -1 > /tmp/line.json; line "$w slots=$n"
  done
done
