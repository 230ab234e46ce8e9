#!/bin/bash
# Charged along-step register budget on the final code: 8 / 6 / 5 resident blocks of 128 threads
B="python bench.py --no-extra --no-cpu-baseline --steps 3 --warmup 3"
line() {
  python - "$1" <<'PY'
import json,sys
d=json.load(open('/tmp/line.json'))
print(sys.argv[1], '%.4g track-steps/s' % d['value'], '%.2f ms' % d['ms_per_step'], 'along-step %.2f ms' % (1e3 * d['roofline']['per_action_seconds']['along-step-general-linear']), 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
}
for lib in default celeritas_b200/lib_a6.so celeritas_b200/lib_a5.so; do
  if [ $lib = default ]; then unset CELERITAS_B200_LIB; else export CELERITAS_B200_LIB=$PWD/$lib; fi
  for rep in 1 2; do $B 2>/dev/null | tail -1 > /tmp/line.json; line "$lib testem3"; done
done
