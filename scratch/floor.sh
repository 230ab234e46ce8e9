#!/bin/bash
# per-iteration floor: one 1 GeV primary per pass
for lib in "$@"; do
CELERITAS_B200_LIB=$PWD/celeritas_b200/$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline --events 1 --primaries-per-event 1 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
t=d['roofline']['per_action_seconds']
n=d['num_step_iterations']/d['steps']
print('$lib iters/pass', n, 'us/iter %.1f' % (d['ms_per_step']*1e3/n), {k[:10]: round(v*1e6/n,1) for k,v in t.items()})"
done
