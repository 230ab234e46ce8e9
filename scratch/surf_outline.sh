#!/bin/bash
# A/B of the surface-dispatch layouts (csrc/orange.cuh: B2_SURF_OUTLINE) on one B200.
# default library = level 1; celeritas_b200/lib_s2.so = level 2 (planes only inline).
B="python bench.py --no-extra --no-cpu-baseline --steps 3 --warmup 3"
line() {
  python - "$1" <<'PY'
import json,sys
d=json.load(open('/tmp/line.json'))
print(sys.argv[1], '%.4g track-steps/s' % d['value'], '%.2f ms' % d['ms_per_step'], 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
}
for lib in default celeritas_b200/lib_s2.so; do
  if [ $lib = default ]; then unset CELERITAS_B200_LIB; else export CELERITAS_B200_LIB=$PWD/$lib; fi
  echo "== $lib: parity"
  timeout 600 python -m pytest tests/test_gpu_orange.py tests/test_gpu_many_faces.py tests/test_gpu_testem3.py tests/test_gpu_field.py tests/test_gpu_cms_scale.py tests/test_gpu_golden.py -x -q --timeout 250 2>&1 | tail -2
  for w in testem3 cms-scale simple-cms; do
    for rep in 1 2; do
      $B --workload $w 2>/dev/null | tail -1 > /tmp/line.json; line "$lib $w"
    done
  done
done
