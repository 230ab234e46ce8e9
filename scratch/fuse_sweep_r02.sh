#!/bin/bash
# Fused-launch threshold sweep on the final build (no rebuild): B200_FUSE_THRESHOLD
B="python bench.py --no-extra --no-cpu-baseline --steps 3 --warmup 3"
line() {
  python - "$1" <<'PY'
import json,sys
d=json.load(open('/tmp/line.json'))
print(sys.argv[1], '%.4g track-steps/s' % d['value'], '%.2f ms' % d['ms_per_step'], 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
}
for w in testem3 cms-scale; do
  for t in ${SWEEP:-32768 65536 98304 131072 196608}; do
    B200_FUSE_THRESHOLD=$t $B --workload $w 2>/dev/null | tail -1 > /tmp/line.json; line "$w fuse<=$t"
  done
done
