"""Per-iteration wall time and track counts of one bench pass (single stream)."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import celeritas_b200 as cb
import bench
workload = sys.argv[1] if len(sys.argv) > 1 else 'testem3'
wl = bench.WORKLOADS[workload]
params = cb.Params(wl['image'])
import os
fuse = os.environ.get('FUSE')
st = cb.Stepper(params, 1 << 20, **({'fuse_threshold': int(fuse)} if fuse else {}))
prim, offsets = bench.make_workload_events(workload, params, wl['events'], wl['per_event'], 0,
                                           cb.PRIMARY_DTYPE)
for rep in range(3):
    st.reseed(0)
    rows = []
    t0 = time.perf_counter()
    c = st.step(prim)
    t1 = time.perf_counter(); rows.append((c['active'], c['alive'], c['queued'], t1 - t0))
    while c['alive'] or c['queued']:
        t0 = time.perf_counter()
        c = st.step()
        t1 = time.perf_counter(); rows.append((c['active'], c['alive'], c['queued'], t1 - t0))
a = np.array(rows)
print('iterations', len(a), 'total ms %.1f' % (a[:, 3].sum() * 1e3), 'track-steps %.3g' % a[:, 0].sum())
edges = [0, 16, 128, 1024, 4096, 16384, 65536, 131072, 262144, 524288, 1 << 21]
for lo, hi in zip(edges[:-1], edges[1:]):
    m = (a[:, 0] >= lo) & (a[:, 0] < hi)
    if m.any():
        print('active [%7d,%7d): iters %3d  time %6.1f ms  steps %.3g  us/iter %.0f  ns/track-step %.2f' % (
            lo, hi, m.sum(), a[m, 3].sum() * 1e3, a[m, 0].sum(), a[m, 3].mean() * 1e6,
            a[m, 3].sum() * 1e9 / max(a[m, 0].sum(), 1)))
np.save('gpurun_out/iter_profile.npy', a)
