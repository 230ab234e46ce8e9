"""TrackOrder::reindex_* (SortTracksAction, src/celeritas/track/SortTracksAction.cc:46-131,
track/detail/TrackSortUtils.{cc,cu}): the reference keeps a permutation of all track slots
sorted by a key (active/inactive, particle type, along-step action, step-limit action, or
the two action sorts one after the other). Its sorts are not stable (std::sort /
thrust::sort_by_key), so SURVEY 8(a) a20 asks for: the same key sequence, and the same SET of
slots per key. The problem image is exported here, at test time, from the reference's own
CoreParams built with that track order (the sort actions are entries of its action table).

Everything else is the usual lock-step comparison: per-slot state, RNG words and counters of
every iteration are identical to the reference's, whatever the order."""
import json

import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu

ORDERS = ['reindex_status', 'reindex_particle_type', 'reindex_along_step_action',
          'reindex_step_limit_action', 'reindex_both_action']


def setup(order, tmp_path, slots):
    import celeritas_b200 as cb
    import celerref
    cfg = json.load(open(data_path('images', 'testem3-small.json')))
    cfg['track_order'] = order
    problem = celerref.Problem(cfg)
    image = str(tmp_path / ('testem3-%s.b2img' % order))
    problem.export_image(image)
    params = cb.Params(image)
    return problem, problem.stepper(slots), params, cb.Stepper(params, slots)


@pytest.mark.parametrize('order', ORDERS)
def test_sorted_permutation_matches_reference(order, tmp_path):
    import celeritas_b200 as cb
    from parity import compare_states
    slots = 2048
    problem, ref, params, gpu = setup(order, tmp_path, slots)
    labels = params.action_labels
    assert any(l.startswith('sort-tracks-') for l in labels)
    assert any(l.startswith('sort-tracks-') for l in gpu.step_action_labels)
    nkeys = {'reindex_status': 1, 'reindex_particle_type': params.num_particles}.get(
        order, len(labels))
    prim = cb.make_primaries(2, particle_id=params.find_particle(11), energy=1000.0,
                             pos=(-22, 0, 0), direction=(1, 0, 0))
    cr, cg = ref.step(prim), gpu.step(prim)
    it, sizes = 0, []
    while True:
        assert cr == cg, 'iteration %d: ref=%s gpu=%s' % (it, cr, cg)
        if it % 5 == 0:
            compare_states(ref, gpu, it)
            want, got = ref.get('track_slots'), gpu.get('sort_slots')
            off = gpu.sort_offsets()[:nkeys + 2]
            # a permutation of all slots, with ranges that cover it
            assert np.array_equal(np.sort(got), np.arange(slots))
            assert np.array_equal(np.sort(want), np.arange(slots))
            assert off[0] == 0 and off[-1] == slots and np.all(np.diff(off.astype(np.int64)) >= 0)
            # same set of slots per key: the reference's permutation, cut at OUR offsets
            for k in range(nkeys + 1):
                a, b = int(off[k]), int(off[k + 1])
                assert np.array_equal(np.sort(want[a:b]), np.sort(got[a:b])), \
                    'iteration %d key %d' % (it, k)
            if order == 'reindex_status':
                assert off[1] == cr['active']
            if order in ('reindex_step_limit_action', 'reindex_both_action'):
                # the keys of tracks that survived the step are still in place: their
                # sequence along the permutation is sorted, and the same as the reference's
                # (a slot taken over by a secondary at the end of the step has a new key)
                status = gpu.get('status')
                key = np.minimum(gpu.get('post_step_action'), nkeys)
                same = gpu.get('num_steps') > 0
                keep = lambda perm: key[perm][(status[perm] != 0) & same[perm]]
                assert np.all(np.diff(keep(got).astype(np.int64)) >= 0)
                assert np.array_equal(keep(got), keep(want))
                sizes.append(int((np.diff(off.astype(np.int64)) > 0).sum()))
        if not (cr['alive'] or cr['queued']):
            break
        cr, cg = ref.step(), gpu.step()
        it += 1
    assert it > 100
    if sizes:
        assert max(sizes) >= 5  # several distinct step-limit actions were in play at once


def test_step_limit_ranges_are_the_interaction_lists(tmp_path):
    """The per-model interaction lists the discrete-select launch builds (what the
    interaction kernels run over) are exactly the sorted permutation's action ranges."""
    import celeritas_b200 as cb
    problem, ref, params, gpu = setup('reindex_step_limit_action', tmp_path, 4096)
    prim = cb.make_primaries(4, particle_id=params.find_particle(11), energy=1000.0,
                             pos=(-22, 0, 0), direction=(1, 0, 0))
    gpu.step(prim)
    seen = 0
    for it in range(60):
        c = gpu.step()
        perm, off = gpu.get('sort_slots'), gpu.sort_offsets()
        for action, slots in gpu.interaction_lists().items():
            a, b = int(off[action]), int(off[action + 1])
            assert np.array_equal(np.sort(perm[a:b]), np.sort(slots)), (it, action)
            seen += len(slots)
    assert seen > 500


def test_celer_sim_accepts_reindex_orders(tmp_path):
    """celer-sim's `track_order` values (app/celer-sim/RunnerInput.hh) are honoured when the
    image was exported with them, and refused with a message otherwise."""
    import celeritas_b200 as cb
    import celerref
    cfg = json.load(open(data_path('images', 'testem3-small.json')))
    cfg['track_order'] = 'reindex_step_limit_action'
    image = str(tmp_path / 'sorted.b2img')
    celerref.Problem(cfg).export_image(image)
    inp = {'_format': 'celer-sim', 'use_device': True, 'image_file': image,
           'geometry_file': cfg['geometry_file'], 'physics_file': cfg['physics_file'],
           'num_track_slots': 4096, 'initializer_capacity': cfg['initializer_capacity'],
           'secondary_stack_factor': 3, 'seed': cfg['seed'], 'simple_calo': cfg['simple_calo'],
           'track_order': 'reindex_step_limit_action', 'merge_events': True,
           'primary_options': {'seed': 0, 'pdg': [11], 'num_events': 2, 'primaries_per_event': 2,
                               'energy': {'distribution': 'delta', 'params': [100.0]},
                               'position': {'distribution': 'delta', 'params': [-22, 0, 0]},
                               'direction': {'distribution': 'delta', 'params': [1, 0, 0]}}}
    report = cb.celer_sim_run(inp)
    assert report['input']['track_order'] == 'reindex_step_limit_action'
    assert report['result']['runner']['num_steps'][0] > 100
    inp['track_order'] = 'init_charge'
    with pytest.raises(cb.B200Error) as e:
        cb.celer_sim_run(inp)
    assert 'sort' in str(e.value)
    inp['track_order'] = 'reindex_everything'
    with pytest.raises(cb.B200Error):
        cb.celer_sim_run(inp)


def test_reindex_shuffle_is_slot_identical():
    """TrackOrder::reindex_shuffle: the reference shuffles its thread -> slot map once at
    construction (global/CoreTrackData.cc:52-56, detail/TrackSlotUtils.cc:21-32: std::shuffle
    with mt19937 seeded by the slot count). Which thread works on a slot changes; what happens
    in the slot does not. The reference built with that order is compared slot by slot with
    this library running the stock image (threads walk dense lists here, there is no
    thread -> slot map to shuffle), and celer-sim accepts the value."""
    import celeritas_b200 as cb
    import celerref
    from parity import lockstep
    cfg = json.load(open(data_path('images', 'testem3-small.json')))
    ref_problem = celerref.Problem(dict(cfg, track_order='reindex_shuffle'))
    slots = 4096
    ref = ref_problem.stepper(slots)
    shuffled = ref.get('track_slots')
    assert np.array_equal(np.sort(shuffled), np.arange(slots))
    assert not np.array_equal(shuffled, np.arange(slots))
    params = cb.Params(data_path('images', 'testem3-small.b2img'))
    gpu = cb.Stepper(params, slots)
    prim = cb.make_primaries(3, particle_id=params.find_particle(11), energy=1000.0,
                             pos=(-22, 0, 0), direction=(1, 0, 0))
    hist = lockstep(ref, gpu, prim, compare_every=3)
    assert sum(h['active'] for h in hist) > 20000
    assert np.allclose(ref_problem.calo(100), gpu.calo(), rtol=1e-9, atol=1e-9)

    from conftest import REPO
    inp = {'_format': 'celer-sim', 'use_device': True, 'base_dir': REPO,
           'image_file': 'data/images/testem3-small.b2img',
           'geometry_file': cfg['geometry_file'], 'physics_file': cfg['physics_file'],
           'seed': cfg['seed'], 'num_track_slots': 1024,
           'initializer_capacity': cfg['initializer_capacity'], 'secondary_stack_factor': 3,
           'simple_calo': cfg['simple_calo'], 'track_order': 'reindex_shuffle',
           'primary_options': {'seed': 0, 'pdg': [11], 'num_events': 1, 'primaries_per_event': 2,
                               'energy': {'distribution': 'delta', 'params': [100.0]},
                               'position': {'distribution': 'delta', 'params': [-22, 0, 0]},
                               'direction': {'distribution': 'delta', 'params': [1, 0, 0]}}}
    report = cb.celer_sim_run(inp)
    assert report['input']['track_order'] == 'reindex_shuffle'
    none = cb.celer_sim_run(dict(inp, track_order='none'))
    assert report['result']['runner']['num_steps'] == none['result']['runner']['num_steps']
