"""One bench pass on one stream with the device-resident loop (for ncu)."""
import sys
sys.path.insert(0, '.')
import celeritas_b200 as cb
import bench
workload = sys.argv[1] if len(sys.argv) > 1 else 'cms-scale'
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 64
wl = bench.WORKLOADS[workload]
params = cb.Params(wl['image'])
st = cb.Stepper(params, 1 << 19, tail_threshold=thr)
prim, offsets = bench.make_workload_events(workload, params, wl['events'], wl['per_event'], 0,
                                           cb.PRIMARY_DTYPE)
r = st.run_events(prim, offsets, merge_events=True)
print(r, st.tail_iterations)
