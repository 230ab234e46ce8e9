"""GPU parity: two-box geometry with hand-made Compton tables (no Geant4 data).

This is the reference's own SimpleComptonTest (test/celeritas/global/Stepper.test.cc:194-225):
919 step iterations, 53.8125 steps per primary, initializer-queue high-water mark 6 at
iteration 1 -- for both the host and device builds of the reference.
"""
import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu

CONFIG = {'problem': 'simple-compton', 'geometry_file': 'data/geometry/two-boxes.org.json',
          'seed': 20220511}


def primaries(n):
    import celeritas_b200 as cb
    return cb.make_primaries(n, particle_id=0, energy=100.0, pos=(-22, 0, 0), direction=(1, 0, 0))


def test_reference_golden_counts():
    import celeritas_b200 as cb
    params = cb.Params(data_path('images', 'simple-compton.b2img'))
    step = cb.Stepper(params, 64)
    c = step.step(primaries(32))
    active, queued = [c['active']], [c['queued']]
    while c['queued'] > 0 or c['alive'] > 0:
        c = step.step()
        active.append(c['active'])
        queued.append(c['queued'])
    assert len(active) == 919
    assert sum(active) / 32 == 53.8125
    assert (queued.index(max(queued)), max(queued)) == (1, 6)
    assert step.launch_count > 0


def test_lockstep_with_reference():
    import celeritas_b200 as cb
    import celerref
    from parity import lockstep
    ref = celerref.Problem(CONFIG).stepper(64)
    params = cb.Params(data_path('images', 'simple-compton.b2img'))
    gpu = cb.Stepper(params, 64)
    hist = lockstep(ref, gpu, primaries(32))
    assert len(hist) == 919


def test_lockstep_many_slots():
    import celeritas_b200 as cb
    import celerref
    from parity import lockstep
    ref = celerref.Problem(CONFIG).stepper(1024)
    params = cb.Params(data_path('images', 'simple-compton.b2img'))
    gpu = cb.Stepper(params, 1024)
    lockstep(ref, gpu, primaries(1000), max_iters=60, compare_every=5)


def test_reseed_reproducible():
    import celeritas_b200 as cb
    params = cb.Params(data_path('images', 'simple-compton.b2img'))
    step = cb.Stepper(params, 16)
    step.reseed(123)
    a = step.get('rng').copy()
    step.reseed(3456)
    b = step.get('rng').copy()
    step.reseed(123)
    assert np.array_equal(a, step.get('rng'))
    assert not np.array_equal(a, b)


def test_reseed_matches_reference():
    import celeritas_b200 as cb
    import celerref
    ref = celerref.Problem(CONFIG).stepper(16)
    params = cb.Params(data_path('images', 'simple-compton.b2img'))
    gpu = cb.Stepper(params, 16)
    assert np.array_equal(ref.get('rng'), gpu.get('rng'))  # mt19937 initial fill
    for ev in (0, 1, 123, 2 ** 33 + 5):
        ref.reseed(ev)
        gpu.reseed(ev)
        assert np.array_equal(ref.get('rng'), gpu.get('rng'))
