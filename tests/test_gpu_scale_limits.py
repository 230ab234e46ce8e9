"""Branches that only large problems reach (VERDICT r01, weak point 10):

* the end-of-step scan of block totals beyond 2^21 slots per stream (k_end_pass2, `per > 16`:
  a thread's run of block totals no longer fits its registers);
* tallies with more bins than the per-block shared-memory copy holds (1024): SimpleCalo over
  1500 detector volumes (O(10^4)-cell calorimeters are what SURVEY 2.2 sizes CMS for) and a
  step diagnostic with 3 x 502 bins fall back to global atomics.

All against the reference's host Stepper on the same inputs."""
import json

import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu

NEVER = 0xffffffff


def test_end_of_step_scan_beyond_two_million_slots():
    import celeritas_b200 as cb
    import celerref
    from parity import lockstep
    cfg = {'problem': 'simple-compton', 'geometry_file': 'data/geometry/two-boxes.org.json',
           'seed': 20220511, 'initializer_capacity': 1 << 16}
    slots = 1 << 22
    problem = celerref.Problem(cfg)
    ref = problem.stepper(slots)
    import tempfile, os
    with tempfile.TemporaryDirectory() as tmp:
        image = os.path.join(tmp, 'compton.b2img')
        problem.export_image(image)
        params = cb.Params(image)
    gpu = cb.Stepper(params, slots, fuse_threshold=NEVER, tail_threshold=NEVER)
    prim = cb.make_primaries(3000, particle_id=0, energy=100.0, pos=(-22, 0, 0),
                             direction=(1, 0, 0))
    hist = lockstep(ref, gpu, prim, max_iters=5, compare_every=5)
    assert hist[0]['active'] == 3000 and len(hist) == 6 and hist[-1]['alive'] > 100


@pytest.mark.parametrize('fuse', [0, NEVER], ids=['fused', 'per-action'])
def test_calo_with_more_detectors_than_shared_bins(fuse, tmp_path):
    import celeritas_b200 as cb
    import celerref
    from parity import lockstep
    from test_gpu_field import isotropic_mix
    cfg = json.load(open(data_path('images', 'cms-scale-small.json')))
    labels = cb.Params(data_path('images', 'cms-scale-small.b2img')).volume_labels
    seen, detectors = set(), []
    for name in labels:
        if name and not name.startswith('[') and name not in seen and labels.count(name) == 1:
            seen.add(name)
            detectors.append(name)
    detectors = detectors[:1500]
    assert len(detectors) == 1500
    cfg['simple_calo'] = detectors
    problem = celerref.Problem(cfg)
    image = str(tmp_path / 'cms-1500-detectors.b2img')
    problem.export_image(image)
    params = cb.Params(image)
    assert params.num_detectors == 1500
    slots = 16384
    ref = problem.stepper(slots)
    gpu = cb.Stepper(params, slots, fuse_threshold=fuse)
    hist = lockstep(ref, gpu, isotropic_mix(6, 1000.0, params, seed=21), max_iters=50000,
                    compare_every=25, rtol=1e-5, atol=1e-5)
    assert not (hist[-1]['alive'] or hist[-1]['queued'])
    want, got = problem.calo(1500), gpu.calo()
    assert np.allclose(want, got, rtol=1e-9, atol=1e-9)
    assert (got > 0).sum() > 40 and got.sum() > 50.0


@pytest.mark.parametrize('fuse', [0, NEVER], ids=['fused', 'per-action'])
def test_step_diagnostic_beyond_shared_bins(fuse):
    import celeritas_b200 as cb
    import celerref
    bins = 500  # 3 particles x 502 bins > 1024
    cfg = json.load(open(data_path('images', 'testem3-small.json')))
    cfg.update(action_diagnostic=True, step_diagnostic_bins=bins)
    refp = celerref.Problem(cfg)
    params = cb.Params(data_path('images', 'testem3-small.b2img'))
    slots = 4096
    ref = refp.stepper(slots)
    gpu = cb.Stepper(params, slots, action_diagnostic=True, step_diagnostic_bins=bins,
                     fuse_threshold=fuse)
    prim = cb.make_primaries(4, particle_id=params.find_particle(11), energy=500.0,
                             pos=(-22, 0, 0), direction=(1, 0, 0))
    cr, cg = ref.step(prim), gpu.step(prim)
    while cr['alive'] or cr['queued']:
        assert cr == cg
        cr, cg = ref.step(), gpu.step()
    want = refp.diagnostic(steps=True)
    assert want.shape == (3, bins + 2) and want.sum() > 1000
    assert np.array_equal(want, gpu.step_diagnostic())
