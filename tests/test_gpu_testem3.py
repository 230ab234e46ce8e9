"""GPU parity on the TestEm3 stand-in (50 steel/lAr layers, see tools/make_physics.py):
lock-step comparison with the reference's host Stepper, slot by slot, including the
RNG stream state of every slot (which pins the number and order of random draws).
"""
import json

import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu


NEVER_FUSE = 0xffffffff


def setup(name, slots, fuse_threshold=0):
    """fuse_threshold: 0 = library default (these small problems then run every iteration
    as ONE fused launch); NEVER_FUSE = one launch per action (the path large iterations
    take). Both must reproduce the reference."""
    import celeritas_b200 as cb
    import celerref
    cfg = json.load(open(data_path('images', name + '.json')))
    ref_problem = celerref.Problem(cfg)
    ref = ref_problem.stepper(slots)
    params = cb.Params(data_path('images', name + '.b2img'))
    gpu = cb.Stepper(params, slots, fuse_threshold=fuse_threshold)
    return ref_problem, ref, params, gpu


def electrons(n, energy, params):
    import celeritas_b200 as cb
    return cb.make_primaries(n, particle_id=params.find_particle(11), energy=energy,
                             pos=(-22, 0, 0), direction=(1, 0, 0))


@pytest.mark.parametrize('energy,nprim,slots,iters', [(10.0, 8, 256, 200), (1000.0, 2, 4096, 400)])
def test_lockstep_nomsc(energy, nprim, slots, iters):
    from parity import lockstep
    _, ref, params, gpu = setup('testem3-nomsc', slots)
    lockstep(ref, gpu, electrons(nprim, energy, params), max_iters=iters)


@pytest.mark.parametrize('fuse', [0, NEVER_FUSE], ids=['fused', 'per-action'])
@pytest.mark.parametrize('energy,nprim,slots,iters', [(10.0, 8, 256, 300), (1000.0, 2, 4096, 100000)])
def test_lockstep_full_em(energy, nprim, slots, iters, fuse):
    from parity import lockstep
    _, ref, params, gpu = setup('testem3-small', slots, fuse)
    lockstep(ref, gpu, electrons(nprim, energy, params), max_iters=iters)


def test_mixed_fused_and_per_action_iterations():
    """A threshold inside the shower's size range: iterations switch between the fused
    launch and the per-action launches (with their per-model interaction lists)."""
    from parity import lockstep
    _, ref, params, gpu = setup('testem3-small', 4096, fuse_threshold=150)
    hist = lockstep(ref, gpu, electrons(3, 1000.0, params), compare_every=7)
    sizes = [h['active'] for h in hist]
    assert min(sizes) < 150 < max(sizes)


def test_calo_tally_matches_reference():
    """Per-layer energy deposition of a whole shower vs the reference (same RNG streams)."""
    from parity import lockstep
    refp, ref, params, gpu = setup('testem3-small', 8192, NEVER_FUSE)
    lockstep(ref, gpu, electrons(4, 1000.0, params), compare_every=50)
    a = refp.calo(100)
    b = gpu.calo()
    assert a.sum() > 3000  # most of 4 GeV is deposited in the calorimeter
    # Same tracks, same order of magnitude of roundoff: per-bin relative tolerance
    assert np.allclose(a, b, rtol=1e-9, atol=1e-9)


def test_lockstep_nested_geometry():
    """Two-level geometry: 50 translated daughters of one gap+absorber universe
    (test/orange/data/testem3.org.json). Whole 1 GeV showers in lock-step."""
    from parity import lockstep
    refp, ref, params, gpu = setup('testem3-nested', 4096)
    hist = lockstep(ref, gpu, electrons(2, 1000.0, params), compare_every=1)
    assert not (hist[-1]['alive'] or hist[-1]['queued'])
    assert np.allclose(refp.calo(2), gpu.calo(), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize('fuse', [0, NEVER_FUSE], ids=['fused', 'per-action'])
@pytest.mark.parametrize('energy,nprim,slots', [(10.0, 8, 256), (1000.0, 3, 4096), (1000.0, 4, 512)])
def test_lockstep_init_charge(energy, nprim, slots, fuse):
    """TrackOrder::init_charge (the reference's GPU default, RunnerInputIO.json.cc:113-120):
    starting tracks are partitioned by charge, neutral ones take the lowest vacancies and
    charged ones the highest, and secondaries never reuse their parent's slot. Slot
    assignment is part of the compared state (every field is compared slot by slot); the
    512-slot case keeps initializers queued so that partial starts are exercised."""
    from parity import lockstep
    import celeritas_b200 as cb
    _, ref, params, gpu = setup('testem3-small-initcharge', slots, fuse)
    prim = electrons(nprim, energy, params)
    # mixed charges among the primaries too
    prim['particle_id'][1::2] = params.find_particle(22)
    hist = lockstep(ref, gpu, prim, max_iters=100000)
    assert not (hist[-1]['alive'] or hist[-1]['queued'])
    if slots == 512:
        assert max(h['queued'] for h in hist) > 0


def test_action_sorted_lists_match_reference_actions():
    """The per-model interaction lists (the action-sorted order the interaction kernel
    runs in) hold exactly the slots whose post-step action in the REFERENCE state is that
    model, every step; the dense charged/neutral lists of the next step hold exactly the
    reference's active slots of each charge, in slot order."""
    import celeritas_b200 as cb
    _, ref, params, gpu = setup('testem3-small-initcharge', 2048, NEVER_FUSE)
    prim = electrons(4, 1000.0, params)
    cr, cg = ref.step(prim), gpu.step(prim)
    charge_of = {params.find_particle(11): 1, params.find_particle(-11): 1,
                 params.find_particle(22): 0}
    checked = interactions = 0
    for it in range(120):
        assert cr == cg
        lists = gpu.interaction_lists()
        action = ref.get('post_step_action')
        status = ref.get('status')
        for act, slots in lists.items():
            # in init_charge mode a slot keeps its post-step action until the next step
            want = np.nonzero((action == act))[0]
            assert np.array_equal(slots, want.astype(np.uint32)), (it, act)
            interactions += len(slots)
        if it % 10 == 0:
            pid = ref.get('particle_id')
            alive = np.nonzero(status == 2)[0]
            charged = np.array([s for s in alive if charge_of[int(pid[s])]], dtype=np.uint32)
            neutral = np.array([s for s in alive if not charge_of[int(pid[s])]], dtype=np.uint32)
            gc, gn = gpu.dense_lists((len(charged), len(neutral)))
            assert np.array_equal(gc, charged) and np.array_equal(gn, neutral), it
            checked += 1
        if not (cr['alive'] or cr['queued']):
            break
        cr, cg = ref.step(), gpu.step()
    assert interactions > 500 and checked >= 5


@pytest.mark.parametrize('fuse', [0, NEVER_FUSE], ids=['fused', 'per-action'])
def test_lockstep_combined_brem(fuse):
    """celer-sim `brem_combined`: CombinedBremInteractor (Seltzer-Berger below 1 GeV,
    relativistic with LPM above; em/interactor/CombinedBremInteractor.hh:132-170) behind one
    action. 2 GeV electrons in liquid argon exercise both samplers; whole showers in
    lock-step, integers and RNG words identical."""
    import celeritas_b200 as cb
    from parity import lockstep
    refp, ref, params, gpu = setup('lar-sphere-combined', 4096, fuse)
    assert 'brems-combined' in params.action_labels
    prim = cb.make_primaries(4, particle_id=params.find_particle(11), energy=2000.0,
                             pos=(0, 0, 0), direction=(1, 0, 0))
    hist = lockstep(ref, gpu, prim, compare_every=1)
    assert sum(h['active'] for h in hist) > 10000
    assert np.allclose(refp.calo(1), gpu.calo(), rtol=1e-9, atol=1e-9)
