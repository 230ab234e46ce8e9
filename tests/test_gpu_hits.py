"""Step/hit output: the reference's StepCollector gathers the pre- and
post-step points of every track that steps in a sensitive volume and DetectorSteps compacts
them for the user callback (src/celeritas/user/StepCollector.cc, detail/StepGatherExecutor.hh,
DetectorSteps.cu:150-200: thrust::copy_if + a gather kernel + one device-to-host copy per
field). Here: the pre-step point is kept per slot by the pre-step launch, and after the
tallies one stable partition (the counting sort of csrc/kernels_sort.cu) plus one gather
kernel write the compact records.

Checked after EVERY step iteration against the reference's own DetectorStepOutput on the same
problem: the same hits in the same (slot) order; ids exactly, reals to 1e-7."""
import json

import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu

INT_FIELDS = ['detector', 'track_id', 'event_id', 'parent_id', 'track_step_count', 'particle']


@pytest.mark.parametrize('nonzero', [False, True], ids=['all-steps', 'nonzero-edep'])
def test_hits_match_reference_every_iteration(nonzero, tmp_path):
    import celeritas_b200 as cb
    import celerref
    from celerref import HIT_FIELDS
    cfg = json.load(open(data_path('images', 'testem3-small.json')))
    del cfg['simple_calo']
    cfg['hit_volumes'] = ['gap_%d' % i for i in range(0, 50, 2)] + ['absorber_3', 'absorber_4']
    cfg['hits_nonzero_edep'] = nonzero
    problem = celerref.Problem(cfg)
    image = str(tmp_path / 'hits.b2img')
    problem.export_image(image)
    params = cb.Params(image)
    slots = 4096
    ref, gpu = problem.stepper(slots), cb.Stepper(params, slots)
    prim = cb.make_primaries(3, particle_id=params.find_particle(11), energy=500.0,
                             pos=(-22, 0, 0), direction=(1, 0, 0))
    cr, cg = ref.step(prim), gpu.step(prim)
    total, it = 0, 0
    while True:
        assert cr == cg
        want, got = problem.hits(), gpu.hits()
        n = len(want['detector'])
        assert len(got['detector']) == n, 'iteration %d: %d hits, reference %d' % (
            it, len(got['detector']), n)
        for f in HIT_FIELDS:
            if f in INT_FIELDS:
                assert np.array_equal(want[f], got[f]), (it, f)
            else:
                assert np.allclose(want[f], got[f], rtol=1e-7, atol=1e-7), (it, f)
        if nonzero and n:
            assert (got['energy_deposition'] != 0).all()
        total += n
        if not (cr['alive'] or cr['queued']):
            break
        cr, cg = ref.step(), gpu.step()
        it += 1
    assert it > 100 and total > (500 if nonzero else 2000)
    # hits were taken in several detectors and for all three particle types
    assert gpu.launch_count > 0
