#!/bin/bash
# usage: scratch/fuse.sh T1 T2 ...  : bench at fuse thresholds (4294967295 = never)
for T in "$@"; do
B200_FUSE_THRESHOLD=$T python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams ${STREAMS:-1} 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('fuse<=$T streams=${STREAMS:-1}', '%.4g'%d['value'], '%.1f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'iters', d['num_step_iterations'], 'launches', d['gpu_launches'])"
done
