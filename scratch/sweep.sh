#!/bin/bash
# usage: scratch/sweep.sh "slots list" "streams list" [fuse]
for SL in $1; do for S in $2; do
B200_FUSE_THRESHOLD=${3:-131072} python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams $S --slots $SL 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('slots=$SL streams=$S', '%.4g'%d['value'], '%.1f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'iters', d['num_step_iterations']//3)"
done; done
