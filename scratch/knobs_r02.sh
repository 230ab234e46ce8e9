#!/bin/bash
# Runtime-knob sweep on one B200 (no rebuild): streams per GPU and the device-resident loop's
# threshold, TestEm3 and CMS-scale. Usage: gpurun -- bash scratch/knobs_r02.sh
B="python bench.py --no-extra --no-cpu-baseline --steps 3 --warmup 3"
run() { # label, env..., args
  label=$1; shift
  env "$@" 2>/dev/null | tail -1 > /tmp/knob_line.json
  python - "$label" <<'PY'
import json,sys
d=json.load(open('/tmp/knob_line.json'))
print(sys.argv[1], '%.4g track-steps/s' % d['value'], '%.1f ms' % d['ms_per_step'], 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
}
run "cms-scale default        " A=1 $B --workload cms-scale
run "cms-scale streams=3      " A=1 $B --workload cms-scale --streams 3
run "cms-scale streams=4      " A=1 $B --workload cms-scale --streams 4
run "cms-scale tail=1024      " B200_TAIL_THRESHOLD=1024 $B --workload cms-scale
run "cms-scale tail=4096      " B200_TAIL_THRESHOLD=4096 $B --workload cms-scale
run "cms-scale tail off       " B200_TAIL_THRESHOLD=4294967295 $B --workload cms-scale
run "testem3 default          " A=1 $B
run "testem3 streams=3        " A=1 $B --streams 3
run "testem3 streams=4        " A=1 $B --streams 4
run "testem3 tail=1024        " B200_TAIL_THRESHOLD=1024 $B
run "simple-cms default       " A=1 $B --workload simple-cms
