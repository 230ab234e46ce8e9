"""One bench pass on one stream (for ncu)."""
import sys
sys.path.insert(0, '.')
import celeritas_b200 as cb
import bench
params = cb.Params(bench.IMAGE)
st = cb.Stepper(params, 1 << 20)
prim, offsets = bench.make_events(100, 100, 0, params.find_particle(11), cb.PRIMARY_DTYPE)
r = st.run_events(prim, offsets, merge_events=True)
print(r)
