"""GPU parity on the CMS-scale stand-in geometry (tools/make_cms_scale.py; BASELINE configs 3
and 4): four universe levels, BIH trees over 276 and 2304 volumes, general planes (phi
sectors), two rect arrays clipped by cylindrical parents, a daughter placed twice, in a 1 T
uniform field. Lock-step with the reference's host Stepper on isotropic e-/gamma primaries from
the origin: every integer state field, the volume/surface ids at every level the reference
exposes and the RNG words identical after every step iteration (tests/parity.py)."""
import json

import numpy as np
import pytest

from conftest import data_path
from test_gpu_field import isotropic_mix

pytestmark = pytest.mark.gpu

# Real-valued fields: 1e-5 here instead of the 1e-7 of the other problems. Integers and RNG
# words are still compared exactly. Low-energy electrons spiral through metres of vacuum
# (tracker gas, muon gaps) in the field; their gyration phase s/R turns the 1e-10 relative
# energy difference that CUDA libm vs glibc log/exp/sin/cos leaves after a few dozen
# multiple-scattering steps into 1e-6 in direction and position (measured drift, step by
# step: profiles/parity_r01.md, section "CMS-scale").
REAL_TOL = 1e-5


@pytest.mark.parametrize('fuse', [0, 0xffffffff], ids=['fused', 'per-action'])
@pytest.mark.parametrize('energy,nprim,slots,seed', [(100.0, 32, 4096, 3), (1000.0, 8, 65536, 5),
                                                     (10000.0, 2, 262144, 7)])
def test_lockstep_cms_scale(energy, nprim, slots, seed, fuse):
    import celeritas_b200 as cb
    import celerref
    from parity import lockstep
    cfg = json.load(open(data_path('images', 'cms-scale-small.json')))
    refp = celerref.Problem(cfg)
    ref = refp.stepper(slots)
    params = cb.Params(data_path('images', 'cms-scale-small.b2img'))
    gpu = cb.Stepper(params, slots, fuse_threshold=fuse)
    hist = lockstep(ref, gpu, isotropic_mix(nprim, energy, params, seed=seed), max_iters=50000,
                    rtol=REAL_TOL, atol=REAL_TOL)
    assert not (hist[-1]['alive'] or hist[-1]['queued'])
    ndet = len(cfg['simple_calo'])
    assert np.allclose(refp.calo(ndet), gpu.calo(), rtol=1e-9, atol=1e-9)
    # the showers reach the rect-array calorimeters
    assert gpu.calo().sum() > 0.5 * nprim * energy


@pytest.mark.parametrize('fuse', [0, 0xffffffff], ids=['fused', 'per-action'])
def test_lockstep_cms_scale_init_charge(fuse):
    """The bench image (track_order init_charge, the reference's GPU default) in lock-step:
    slot assignment by charge through four universe levels, incl. starts from the queue."""
    import celeritas_b200 as cb
    import celerref
    from parity import lockstep
    cfg = json.load(open(data_path('images', 'cms-scale.json')))
    # host-side capacity of the reference only (no overflow at this size; results do not
    # depend on it)
    cfg['initializer_capacity'] = 1 << 18
    refp = celerref.Problem(cfg)
    slots = 2048  # small on purpose: initializers queue up and start in later iterations
    ref = refp.stepper(slots)
    params = cb.Params(data_path('images', 'cms-scale.b2img'))
    gpu = cb.Stepper(params, slots, fuse_threshold=fuse)
    hist = lockstep(ref, gpu, isotropic_mix(6, 1000.0, params, seed=13), max_iters=50000,
                    rtol=REAL_TOL, atol=REAL_TOL)
    assert not (hist[-1]['alive'] or hist[-1]['queued'])
    assert max(h['queued'] for h in hist) > 0
    assert np.allclose(refp.calo(len(cfg['simple_calo'])), gpu.calo(), rtol=1e-9, atol=1e-9)


def test_cms_scale_init_charge_many_events():
    """The bench configuration (track_order init_charge, events merged) at a reduced size:
    most of the energy is deposited in the tallied calorimeter cells, and two streams agree
    with a single-stream run of the same slot count on the number of track-steps (events are
    independent: the RNG is reseeded from the event id and the slot)."""
    import celeritas_b200 as cb
    params = cb.Params(data_path('images', 'cms-scale.b2img'))
    opts = {'seed': 11, 'pdg': [11, 22], 'num_events': 8, 'primaries_per_event': 4,
            'energy': 10000.0, 'position': [0, 0, 0], 'direction': {'distribution': 'isotropic'}}
    prim, offsets = params.generate_primaries(opts)
    one = cb.Stepper(params, 1 << 18, stream_id=0)
    res1, _ = cb.run_events_streams([one], prim, offsets, merge_events=False)
    two = [cb.Stepper(params, 1 << 18, stream_id=k) for k in range(2)]
    res2, _ = cb.run_events_streams(two, prim, offsets, merge_events=False)
    assert sum(r['num_steps'] for r in res1) == sum(r['num_steps'] for r in res2)
    calo1 = one.calo()
    calo2 = sum(st.calo() for st in two)
    assert np.allclose(calo1, calo2, rtol=1e-9)
    assert 0.5 < calo1.sum() / (len(prim) * 10000.0) <= 1.0
