"""GPU parity with a 1 T uniform magnetic field in the simple-CMS geometry (cylinders +
planes): Dormand-Prince field propagation, boundary crossings along chords, Urban MSC and
energy-loss fluctuations; lock-step with the reference's AlongStepUniformMscAction."""
import json

import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu


def isotropic_mix(n, energy, params, seed=1):
    import celeritas_b200 as cb
    rng = np.random.default_rng(seed)
    prim = cb.make_primaries(n, energy=energy, pos=(0, 0, 0))
    d = rng.normal(size=(n, 3))
    prim['dir'] = d / np.linalg.norm(d, axis=1)[:, None]
    ids = [params.find_particle(11), params.find_particle(22)]
    prim['particle_id'] = [ids[i % 2] for i in range(n)]
    return prim


@pytest.mark.parametrize('fuse', [0, 0xffffffff], ids=['fused', 'per-action'])
@pytest.mark.parametrize('energy,nprim,slots', [(10.0, 16, 1024), (1000.0, 4, 65536)])
def test_lockstep_field(energy, nprim, slots, fuse):
    import celeritas_b200 as cb
    import celerref
    from parity import lockstep
    cfg = json.load(open(data_path('images', 'simple-cms-em-field.json')))
    refp = celerref.Problem(cfg)
    ref = refp.stepper(slots)
    params = cb.Params(data_path('images', 'simple-cms-em-field.b2img'))
    gpu = cb.Stepper(params, slots, fuse_threshold=fuse)
    hist = lockstep(ref, gpu, isotropic_mix(nprim, energy, params), max_iters=20000)
    assert not (hist[-1]['alive'] or hist[-1]['queued'])
    assert np.allclose(refp.calo(5), gpu.calo(), rtol=1e-9, atol=1e-9)
