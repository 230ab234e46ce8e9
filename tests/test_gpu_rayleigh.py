"""Rayleigh scattering: RayleighInteractor
(/root/reference/src/celeritas/em/interactor/RayleighInteractor.hh:107-199) on the reference's
bundled four-steel-slabs export with its LivermoreRayleigh cross sections kept
(tools/make_physics.py), in lock-step with the reference's host Stepper: integers and the
six RNG words of every slot identical (the rejection loop draws 3 + 2 numbers per trial, so
the words pin the accepted trial count), reals at 1e-7. Problems with this model run one
launch per action: the fused step and the device-resident loop are built with the core
interactors only (csrc/interact.cuh: run_interaction<EXTRA>).
"""
import json

import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu

NEVER_FUSE = 0xffffffff
NAME = 'four-steel-slabs-rayleigh'


def setup(slots, fuse):
    import celeritas_b200 as cb
    import celerref
    cfg = json.load(open(data_path('images', NAME + '.json')))
    refp = celerref.Problem(cfg)
    params = cb.Params(data_path('images', NAME + '.b2img'))
    return refp, refp.stepper(slots), params, cb.Stepper(params, slots, fuse_threshold=fuse)


def gammas(params, energies):
    import celeritas_b200 as cb
    p = cb.make_primaries(len(energies), particle_id=params.find_particle(22), energy=1.0,
                          pos=(0, 0, -10), direction=(0, 0, 1))
    p['energy'] = energies
    return p


@pytest.mark.parametrize('fuse', [0, NEVER_FUSE], ids=['default', 'per-action'])
def test_lockstep_rayleigh(fuse):
    """Low-energy photons (20 keV .. 1 MeV; Rayleigh is ~10 % of the attenuation in steel at
    50 keV) through the four slabs; every Rayleigh interaction is counted from the
    post-step action of the reference's state."""
    from parity import lockstep, compare_states
    refp, ref, params, gpu = setup(1024, fuse)
    assert 'scat-rayleigh' in params.action_labels
    rayleigh = params.action_labels.index('scat-rayleigh')
    rng = np.random.default_rng(7)
    prim = gammas(params, np.exp(rng.uniform(np.log(0.02), np.log(1.0), 512)))
    cr, cg = ref.step(prim), gpu.step(prim)
    count = 0
    it = 0
    while True:
        assert cr == cg, (it, cr, cg)
        compare_states(ref, gpu, it)
        active = ref.get('status') != 0
        count += int(np.count_nonzero(ref.get('post_step_action')[active] == rayleigh))
        if not (cr['alive'] or cr['queued']):
            break
        cr, cg = ref.step(), gpu.step()
        it += 1
    assert count > 100, count
    assert np.allclose(refp.calo(4), gpu.calo(), rtol=1e-9, atol=1e-9)


def test_rayleigh_with_showers():
    """The full list (MSC, fluctuations, bremsstrahlung, pair production, photoelectric,
    Compton, annihilation) with Rayleigh on top: 100 MeV electrons, whole showers."""
    from parity import lockstep
    import celeritas_b200 as cb
    refp, ref, params, gpu = setup(4096, 0)
    prim = cb.make_primaries(8, particle_id=params.find_particle(11), energy=100.0,
                             pos=(0, 0, -10), direction=(0, 0, 1))
    hist = lockstep(ref, gpu, prim, compare_every=1)
    assert sum(h['active'] for h in hist) > 3000
    assert np.allclose(refp.calo(4), gpu.calo(), rtol=1e-9, atol=1e-9)


def test_fused_step_refuses_problems_with_extra_models():
    """The fused step (and the device-resident loop) are built with the core interactors
    only: the launcher refuses a problem that has Rayleigh / Coulomb / muon models instead of
    stepping it with an interactor missing, and the Stepper runs such problems one launch
    per action whatever the fuse threshold (both configurations above reproduce the
    reference)."""
    import ctypes as C
    import celeritas_b200 as cb
    L = cb.load_library()
    params = cb.Params(data_path('images', NAME + '.b2img'))
    state = C.c_void_p()
    assert L.b200_state_create(params.h, 0, 256, C.byref(state)) == 0
    try:
        L.b200_step_fused.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        rc = L.b200_step_fused(L.b200_params_view(params.h), L.b200_state_view(state), None)
        assert rc == 10001  # B200_ERR_INVALID_ARGUMENT
    finally:
        L.b200_state_destroy(state)
