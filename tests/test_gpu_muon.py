"""Muon processes: MuHadIonizationInteractor with the Bethe-Bloch and muon
Bethe-Bloch (radiative-correction) delta-ray distributions and MuBremsstrahlungInteractor
(/root/reference/src/celeritas/em/interactor/MuHadIonizationInteractor.hh:103-146,
em/distribution/{BetheBloch,MuBB,BraggICRU73QO}EnergyDistribution.hh,
em/interactor/MuBremsstrahlungInteractor.hh:104-181, em/xs/MuBremsDiffXsCalculator.hh) on the
reference's four-steel-slabs export with its mu-/mu+ tables kept, in lock-step with the
reference's host Stepper: integers and RNG words identical, reals at 1e-7. Muons take the
charged along-step without multiple scattering (the export has none for them) and with
energy-loss fluctuations for a heavy particle.

The Bragg / ICRU73QO models (below 200 keV) are loaded and dispatched but cannot fire in this
export: their delta-ray threshold is only reachable in the vacuum material.
"""
import collections
import json

import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu

NEVER_FUSE = 0xffffffff
NAME = 'four-steel-slabs-muon'


@pytest.mark.parametrize('fuse', [0, NEVER_FUSE], ids=['default', 'per-action'])
def test_lockstep_muons(fuse):
    import celeritas_b200 as cb
    import celerref
    from parity import compare_states
    cfg = json.load(open(data_path('images', NAME + '.json')))
    refp = celerref.Problem(cfg)
    slots = 65536
    ref = refp.stepper(slots)
    params = cb.Params(data_path('images', NAME + '.b2img'))
    gpu = cb.Stepper(params, slots, fuse_threshold=fuse)
    labels = params.action_labels
    for name in ('ioni-icru73qo', 'ioni-bragg', 'ioni-bethe-bloch', 'ioni-mu-bethe-bloch',
                 'brems-muon'):
        assert name in labels
    mu_minus, mu_plus = params.find_particle(13), params.find_particle(-13)
    rng = np.random.default_rng(3)

    def group(n, lo, hi, pos):
        p = np.concatenate([
            cb.make_primaries(n, particle_id=mu_minus, energy=1.0, pos=pos, direction=(0, 0, 1)),
            cb.make_primaries(n, particle_id=mu_plus, energy=1.0, pos=pos, direction=(0, 0, 1))])
        p['energy'] = np.exp(rng.uniform(np.log(lo), np.log(hi), 2 * n))
        return p

    prim = np.concatenate([group(256, 0.05, 50000.0, (0, 0, -10)),    # whole energy range
                           group(64, 0.06, 0.3, (0, 0, 0)),           # stopping in steel
                           group(1024, 5000.0, 100000.0, (1, 1, -10))])  # bremsstrahlung
    cr, cg = ref.step(prim), gpu.step(prim)
    count = collections.Counter()
    it = 0
    while True:
        assert cr == cg, (it, cr, cg)
        if it % 3 == 0:
            compare_states(ref, gpu, it)
        active = ref.get('status') != 0
        mine = gpu.get('post_step_action')[active]
        assert np.array_equal(ref.get('post_step_action')[active], mine), it
        muon = gpu.get('particle_id')[active] >= min(mu_minus, mu_plus)
        count.update(labels[a] for a in mine[muon] if a < len(labels))
        if not (cr['alive'] or cr['queued']):
            break
        cr, cg = ref.step(), gpu.step()
        it += 1
    compare_states(ref, gpu, it)
    assert count['ioni-mu-bethe-bloch'] > 5000, count
    assert count['ioni-bethe-bloch'] > 100, count
    assert count['brems-muon'] >= 5, count
    assert np.allclose(refp.calo(4), gpu.calo(), rtol=1e-9, atol=1e-9)
