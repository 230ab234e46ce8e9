#!/bin/bash
# usage: scratch/streams.sh LIB S1 S2 ...
lib=$1; shift
for S in "$@"; do
CELERITAS_B200_LIB=$PWD/celeritas_b200/$lib python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams $S 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$lib streams=$S', '%.4g'%d['value'], '%.1f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'iters', d['num_step_iterations'])"
done
