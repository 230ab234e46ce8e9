import sys
sys.path.insert(0, '.')
import celeritas_b200 as cb
params = cb.Params('data/images/simple-cms-em-field.b2img')
st = cb.Stepper(params, 1 << 20)
opts = {'seed': 7, 'pdg': [11, 22], 'num_events': 100, 'primaries_per_event': 10,
        'energy': 10000.0, 'position': [0, 0, 0], 'direction': {'distribution': 'isotropic'}}
prim, offsets = params.generate_primaries(opts)
for rep in range(2):
    r = st.run_events(prim, offsets, merge_events=True)
print(r)
st.set_action_times(True)
before = st.action_times
r = st.run_events(prim, offsets, merge_events=True)
after = st.action_times
print('per-action ms:', {k[:14]: round((after[k] - before[k]) * 1e3, 1) for k in after})
import numpy as np
c = st.step(prim); sizes = [c['active']]
while c['alive'] or c['queued']:
    c = st.step(); sizes.append(c['active'])
a = np.array(sizes)
print('iterations', len(a), 'max active', a.max(), 'iters < 65536:', (a < 65536).sum(), 'steps in them %.3g of %.3g' % (a[a < 65536].sum(), a.sum()))
