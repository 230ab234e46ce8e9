"""Single Coulomb scattering: CoulombScatteringInteractor with the Wentzel
OK&VI helper, Mott correction, nuclear form factor and isotope selection
(/root/reference/src/celeritas/em/interactor/CoulombScatteringInteractor.hh:105-176,
em/xs/WentzelHelper.hh, em/distribution/WentzelDistribution.hh, em/xs/MottRatioCalculator.hh,
em/xs/NuclearFormFactors.hh, mat/IsotopeSelector.hh) on the reference's four-steel-slabs
export with its eCoulombScattering tables (e-/e+ above 100 MeV) kept, in lock-step with the
reference's host Stepper: integers and RNG words identical, reals at 1e-7.
"""
import json

import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu

NEVER_FUSE = 0xffffffff
# Integers and RNG words are compared exactly. Reals at 1e-6 instead of the usual 1e-7: the
# scattering angles of 10 GeV electrons are micro-radians and below, where cos(theta) sits at
# the resolution of a double next to 1; a last-bit difference between CUDA's and glibc's
# exp/log in the screening coefficient then moves sin(theta) by ~1e-8 per scattering, and a
# 128-shower run accumulates up to 1.7e-7 in a direction component (measured).
REAL_TOL = 1e-6
NAME = 'four-steel-slabs-coulomb'


@pytest.mark.parametrize('slots,fuse', [(65536, 0), (1 << 18, NEVER_FUSE)],
                         ids=['65536-slots', '262144-slots'])
def test_lockstep_coulomb(slots, fuse):
    """The exported mean free path is ~1.7 m in steel, about 0.4 interactions per 10 GeV
    shower: 128 showers (3.5 million track-steps) give a few dozen. Problems with this model
    run one launch per action (the fused step carries the core interactors only): with
    65 536 slots primaries queue up, with 2^18 slots all start at once; the interactions run
    over the per-model lists."""
    import celeritas_b200 as cb
    import celerref
    from parity import compare_states
    cfg = json.load(open(data_path('images', NAME + '.json')))
    refp = celerref.Problem(cfg)
    ref = refp.stepper(slots)
    params = cb.Params(data_path('images', NAME + '.b2img'))
    gpu = cb.Stepper(params, slots, fuse_threshold=fuse)
    assert 'coulomb-wentzel' in params.action_labels
    coulomb = params.action_labels.index('coulomb-wentzel')
    prim = np.concatenate([
        cb.make_primaries(64, particle_id=params.find_particle(11), energy=10000.0,
                          pos=(0, 0, -10), direction=(0, 0, 1)),
        cb.make_primaries(64, particle_id=params.find_particle(-11), energy=10000.0,
                          pos=(1, 1, -10), direction=(0, 0, 1))])
    cr, cg = ref.step(prim), gpu.step(prim)
    count, steps, it = 0, 0, 0
    while True:
        assert cr == cg, (it, cr, cg)
        if it % 4 == 0:
            compare_states(ref, gpu, it, rtol=REAL_TOL, atol=REAL_TOL)
        active = ref.get('status') != 0
        mine = gpu.get('post_step_action')[active]
        assert np.array_equal(ref.get('post_step_action')[active], mine), it
        count += int(np.count_nonzero(mine == coulomb))
        steps += cr['active']
        if not (cr['alive'] or cr['queued']):
            break
        cr, cg = ref.step(), gpu.step()
        it += 1
    compare_states(ref, gpu, it, rtol=REAL_TOL, atol=REAL_TOL)
    assert count > 10, count
    assert steps > 2000000
    assert np.allclose(refp.calo(4), gpu.calo(), rtol=1e-9, atol=1e-9)
