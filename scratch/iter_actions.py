"""Per-action device time by iteration-size class (single stream, per-action kernels)."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import celeritas_b200 as cb
import bench
params = cb.Params(bench.IMAGE)
st = cb.Stepper(params, 1 << 20, fuse_threshold=0xffffffff)
prim, offsets = bench.make_events(100, 100, 0, params.find_particle(11), cb.PRIMARY_DTYPE)
for rep in range(2):
    st.reseed(0)
    c = st.step(prim)
    while c['alive'] or c['queued']:
        c = st.step()
st.set_action_times(True)
st.reseed(0)
rows = []
prev = st.action_times
c = st.step(prim)
while True:
    cur = st.action_times
    rows.append((c['active'], {k: cur[k] - prev[k] for k in cur}))
    prev = cur
    if not (c['alive'] or c['queued']):
        break
    c = st.step()
edges = [0, 16384, 131072, 524288, 1 << 21]
for lo, hi in zip(edges[:-1], edges[1:]):
    sel = [r for r in rows if lo <= r[0] < hi]
    if not sel:
        continue
    tot = {k: sum(r[1][k] for r in sel) for k in sel[0][1]}
    steps = sum(r[0] for r in sel)
    print('active [%d,%d): iters %d steps %.3g total %.1f ms' % (lo, hi, len(sel), steps, sum(tot.values()) * 1e3))
    print('   ', ' '.join('%s=%.1fms(%.2fns)' % (k[:10], v * 1e3, v * 1e9 / steps) for k, v in tot.items()))
