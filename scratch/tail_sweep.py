"""Time whole passes with the device-resident loop at different thresholds / grid sizes."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import celeritas_b200 as cb
import bench

workload = sys.argv[1] if len(sys.argv) > 1 else 'testem3'
wl = bench.WORKLOADS[workload]
params = cb.Params(wl['image'])
prim, offsets = bench.make_workload_events(workload, params, wl['events'], wl['per_event'], 0, cb.PRIMARY_DTYPE)
NEVER = 0xffffffff
for nstreams in (1, 2):
    for thr, blocks in ((NEVER, 0), (64, 16), (64, 148), (1024, 148), (4096, 148), (16384, 148), (1024, 32), (1024, 296)):
        if blocks:
            os.environ['B200_TAIL_BLOCKS'] = str(blocks)
        steppers = [cb.Stepper(params, (1 << 20) // nstreams, stream_id=k, tail_threshold=thr) for k in range(nstreams)]
        def one():
            for st in steppers:
                st.calo_clear()
            if nstreams == 1:
                r = steppers[0].run_events(prim, offsets, merge_events=True)
                return r['seconds'], r['num_step_iterations'], r['num_steps']
            per, secs = cb.run_events_streams(steppers, prim, offsets, merge_events=True)
            return secs, max(x['num_step_iterations'] for x in per), sum(x['num_steps'] for x in per)
        for _ in range(2):
            one()
        ts = [one() for _ in range(3)]
        tail_it = sum(st.tail_iterations for st in steppers)
        print('streams %d thr %10d blocks %3d: %.1f ms  iters %d steps %d tail_iters %d' % (
            nstreams, thr, blocks, 1e3 * min(t[0] for t in ts), ts[0][1], ts[0][2], tail_it), flush=True)
        del steppers
