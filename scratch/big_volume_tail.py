"""Device-resident loop on the many-faces problem (every distance search is in the 113-face
volume): B200_TAIL_COOP=1 (one warp per track, lanes share the big-volume search) against
B200_TAIL_COOP=0 (one lane per track). Step-phase time per iteration from the loop's own
device timers (B200_TAIL_DUMP), for iterations with at most 16 tracks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import celeritas_b200 as cb

dump = 'gpurun_out/tail_dump_coop%s.txt' % os.environ.get('B200_TAIL_COOP', '1')
if os.path.exists(dump):
    os.remove(dump)
os.environ['B200_TAIL_DUMP'] = dump
params = cb.Params('data/images/many-faces.b2img')
e = params.find_particle(11)
st = cb.Stepper(params, 1024, tail_threshold=1024)
for rep in range(3):
    st.reseed(rep)
    prim = cb.make_primaries(6, particle_id=e, energy=100.0, pos=(0.3, 0.2, 5.0), direction=(0, 0, 1))
    c = st.step(prim)
    while c['alive'] or c['queued']:
        c = st.advance(1024)[-1]
a = np.loadtxt(dump)  # active, A, B (step), C1, C2 [ns], total
small = a[a[:, 0] <= 16]
big = a[a[:, 0] > 16]
print('TAIL_COOP=%s: %d iterations; <=16 tracks: %d iterations, step phase %.1f us mean (%.1f median), '
      'whole iteration %.1f us; >16 tracks: %d iterations, step phase %.1f us'
      % (os.environ.get('B200_TAIL_COOP', '1'), len(a), len(small), small[:, 2].mean() / 1e3,
         np.median(small[:, 2]) / 1e3, small[small[:, 5] > 0][:, 5].mean() / 1e3, len(big),
         big[:, 2].mean() / 1e3 if len(big) else 0))
