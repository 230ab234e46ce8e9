#!/bin/bash
# usage: scratch/variants2.sh lib_a.so ...  : default bench config (2 streams) + single-stream action breakdown
for lib in "$@"; do
  CELERITAS_B200_LIB=$PWD/celeritas_b200/$lib python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
t=d['roofline']['per_action_seconds']
print('$lib', '%.4g'%d['value'], '%.1f ms'%d['ms_per_step'], ' '.join('%s=%.1f'%(k[:10],v*1e3) for k,v in t.items()))"
done
