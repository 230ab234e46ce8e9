"""One bench pass on one stream (for ncu)."""
import sys
sys.path.insert(0, '.')
import celeritas_b200 as cb
import bench
workload = sys.argv[1] if len(sys.argv) > 1 else 'testem3'
wl = bench.WORKLOADS[workload]
params = cb.Params(wl['image'])
st = cb.Stepper(params, 1 << 20)
prim, offsets = bench.make_workload_events(workload, params, wl['events'], wl['per_event'], 0,
                                           cb.PRIMARY_DTYPE)
r = st.run_events(prim, offsets, merge_events=True)
print(r)
