"""Who lives longest in a CMS-scale pass? The events the bench gives to stream k of rank r
(two streams per rank), stepped by hand: for the last tracks alive their particle, energy,
volume, step count, position and step length. Usage: python scratch/cms_tail_who.py <rank> <k>"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import celeritas_b200 as cb
import bench
rank, k = int(sys.argv[1]), int(sys.argv[2])
nstreams = 2
wl = bench.WORKLOADS['cms-scale']
params = cb.Params(wl['image'])
st = cb.Stepper(params, (1 << 20) // nstreams, stream_id=rank * nstreams + k)
prim, offsets = bench.make_workload_events('cms-scale', params, wl['events'], wl['per_event'],
                                           rank * wl['events'], cb.PRIMARY_DTYPE)
mine = np.concatenate([prim[offsets[e]:offsets[e + 1]] for e in range(k, wl['events'], nstreams)])
labels = params.volume_labels
c = st.step(mine)
it = 0
while c['alive'] or c['queued']:
    c = st.step()
    it += 1
    if c['alive'] and c['alive'] <= 2 and it % 150 == 0:
        alive = np.nonzero(st.get('status') != 0)[0]
        for s in alive[:2]:
            print(it, 'particle', int(st.get('particle_id')[s]), 'E %.4g MeV' % st.get('energy')[s],
                  labels[int(st.get('volume_id')[s])], 'steps', int(st.get('num_steps')[s]),
                  'pos', np.round(st.get('pos')[s], 2), 'dir', np.round(st.get('dir')[s], 3),
                  'step %.4g cm' % st.get('step_length')[s], flush=True)
print('rank', rank, 'stream', k, 'iterations', it + 1)
