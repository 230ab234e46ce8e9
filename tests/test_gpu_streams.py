"""Several steppers (CUDA streams) transporting events concurrently from host threads
(b200_run_events_streams) against the reference's per-stream host Steppers: every stream
must end in the reference's state for that stream id, bit for bit in the integers and the
RNG words, and the summed calorimeter tallies must agree."""
import json

import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('merge', [False, True])
@pytest.mark.parametrize('nstreams', [2, 3])
def test_streams_match_reference_streams(nstreams, merge):
    import celeritas_b200 as cb
    import celerref
    from parity import compare_states
    slots, nevents, per_event = 2048, 7, 2
    cfg = json.load(open(data_path('images', 'testem3-small.json')))
    cfg['max_streams'] = nstreams
    refp = celerref.Problem(cfg)
    params = cb.Params(data_path('images', 'testem3-small.b2img'))
    steppers = [cb.Stepper(params, slots, stream_id=k) for k in range(nstreams)]
    prim = cb.make_primaries(nevents * per_event, particle_id=params.find_particle(11),
                             energy=300.0, pos=(-22, 0, 0), direction=(1, 0, 0),
                             event_of=lambda i: i // per_event)
    offsets = np.arange(0, len(prim) + 1, per_event, dtype=np.uint32)
    results, seconds = cb.run_events_streams(steppers, prim, offsets, merge_events=merge)
    assert seconds > 0 and len(results) == nstreams

    for k in range(nstreams):
        ref = refp.stepper(slots, stream_id=k)
        events = [e for e in range(nevents) if e % nstreams == k]
        batches = ([np.concatenate([prim[offsets[e]:offsets[e + 1]] for e in events])]
                   if merge else [prim[offsets[e]:offsets[e + 1]] for e in events])
        steps = iters = 0
        for batch in batches:
            ref.reseed(int(batch[0]['event_id']))
            c = ref.step(batch)
            while True:
                steps += c['active']
                iters += 1
                if not (c['alive'] or c['queued']):
                    break
                c = ref.step()
        assert results[k]['num_steps'] == steps
        assert results[k]['num_step_iterations'] == iters
        assert results[k]['num_primaries'] == sum(len(b) for b in batches)
        compare_states(ref, steppers[k], step=-1)
    # SimpleCalo sums over streams
    ndet = params.num_detectors
    assert np.allclose(refp.calo(ndet), sum(s.calo() for s in steppers), rtol=1e-9, atol=1e-9)
