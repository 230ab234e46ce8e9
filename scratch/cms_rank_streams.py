"""The CMS-scale pass of one rank of the bench, per stream: iterations and seconds.
Usage: python scratch/cms_rank_streams.py <rank> <nstreams>"""
import sys
sys.path.insert(0, '.')
import numpy as np
import celeritas_b200 as cb
import bench
rank, nstreams = int(sys.argv[1]), int(sys.argv[2])
wl = bench.WORKLOADS['cms-scale']
params = cb.Params(wl['image'])
steppers = [cb.Stepper(params, (1 << 20) // nstreams, stream_id=rank * nstreams + k)
            for k in range(nstreams)]
prim, offsets = bench.make_workload_events('cms-scale', params, wl['events'], wl['per_event'],
                                           rank * wl['events'], cb.PRIMARY_DTYPE)
for rep in range(3):
    for s in steppers:
        s.reseed(0) if hasattr(s, 'reseed') else None
    if nstreams == 1:
        r = steppers[0].run_events(prim, offsets, merge_events=True)
        print(rep, r)
    else:
        per, sec = cb.run_events_streams(steppers, prim, offsets, merge_events=True)
        print(rep, 'pass %.1f ms' % (sec * 1e3),
              [(p['num_step_iterations'], round(p['seconds'] * 1e3, 1), p['num_steps']) for p in per])
