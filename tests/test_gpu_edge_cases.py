"""Edge cases of the step loop, each against the reference's behaviour for the same input:
empty steps, primaries outside the geometry, more primaries than slots, kill_active,
capacity errors, bad event ids."""
import json

import numpy as np
import pytest

from conftest import REPO, data_path

pytestmark = pytest.mark.gpu


def setup(name, slots, **stepper_kw):
    import celeritas_b200 as cb
    import celerref
    cfg = json.load(open(data_path('images', name + '.json')))
    refp = celerref.Problem(cfg)
    params = cb.Params(data_path('images', name + '.b2img'))
    return refp, refp.stepper(slots), params, cb.Stepper(params, slots, **stepper_kw)


def test_empty_steps_and_warm_up():
    from parity import compare_states
    _, ref, params, gpu = setup('testem3-small', 64)
    gpu.warm_up()
    zero = dict(generated=0, queued=0, active=0, alive=0)
    assert ref.step() == zero
    assert gpu.step() == zero
    assert gpu.step(np.zeros(0, dtype=gpu.get('rng').dtype)) == zero
    compare_states(ref, gpu, 0)


@pytest.mark.parametrize('fuse', [0, 0xffffffff], ids=['fused', 'per-action'])
def test_primaries_outside_or_on_the_world_boundary(fuse):
    """Tracks that start outside the geometry are flagged errored at initialization and
    removed by the tracking cut (InitTracksExecutor.hh:128-150, TrackingCutExecutor.hh)."""
    import celeritas_b200 as cb
    from parity import lockstep
    _, ref, params, gpu = setup('testem3-small', 64, fuse_threshold=fuse)
    prim = cb.make_primaries(6, particle_id=params.find_particle(11), energy=50.0,
                             pos=(-22, 0, 0), direction=(1, 0, 0))
    prim['pos'][1] = (1e4, 0, 0)       # far outside
    prim['pos'][3] = (0, -1e4, 5)      # far outside
    prim['particle_id'][4] = params.find_particle(22)
    hist = lockstep(ref, gpu, prim, max_iters=10000)
    assert hist[0]['active'] == 6
    assert not (hist[-1]['alive'] or hist[-1]['queued'])


@pytest.mark.parametrize('image', ['testem3-small', 'testem3-small-initcharge'])
def test_more_primaries_than_slots(image):
    import celeritas_b200 as cb
    from parity import lockstep
    _, ref, params, gpu = setup(image, 32)
    prim = cb.make_primaries(100, particle_id=params.find_particle(22), energy=5.0,
                             pos=(-22, 0, 0), direction=(1, 0, 0))
    prim['particle_id'][::3] = params.find_particle(11)
    hist = lockstep(ref, gpu, prim, max_iters=100000)
    assert hist[0]['queued'] == 100 - 32 and hist[0]['active'] == 32
    assert not (hist[-1]['alive'] or hist[-1]['queued'])


def test_kill_active_matches_reference():
    """Stepper::kill_active (global/Stepper.cc:177-182, detail/KillActive.hh): every active
    track is flagged errored and deposits its energy through the tracking cut."""
    import celeritas_b200 as cb
    from parity import compare_states
    refp, ref, params, gpu = setup('testem3-small', 1024)
    prim = cb.make_primaries(4, particle_id=params.find_particle(11), energy=500.0,
                             pos=(-22, 0, 0), direction=(1, 0, 0))
    cr, cg = ref.step(prim), gpu.step(prim)
    for _ in range(12):
        cr, cg = ref.step(), gpu.step()
    assert cr == cg and cr['alive'] > 20
    ref.kill_active()
    gpu.kill_active()
    cr, cg = ref.step(), gpu.step()
    assert cr == cg
    compare_states(ref, gpu, 13)
    # queued initializers survive a kill; finish the event
    while cr['alive'] or cr['queued']:
        cr, cg = ref.step(), gpu.step()
        assert cr == cg
    compare_states(ref, gpu, -1)
    assert np.allclose(refp.calo(params.num_detectors), gpu.calo(), rtol=1e-9, atol=1e-9)


def test_capacity_errors_have_the_reference_messages():
    import celeritas_b200 as cb
    cfg = json.load(open(data_path('images', 'testem3-small.json')))
    prim_opts = {'seed': 0, 'pdg': [11], 'num_events': 1, 'primaries_per_event': 200,
                 'energy': 1000.0, 'position': [-22, 0, 0], 'direction': [1, 0, 0]}
    base = {'use_device': True, 'image_file': data_path('images', 'testem3-small.b2img'),
            'geometry_file': cfg['geometry_file'], 'primary_options': prim_opts,
            'seed': 1, 'num_track_slots': 256, 'secondary_stack_factor': 3, 'warm_up': False}
    # too small for the primaries (ExtendFromPrimariesAction.cc:107-113)
    with pytest.raises(cb.B200Error) as err:
        cb.celer_sim_run(dict(base, initializer_capacity=100))
    assert 'insufficient initializer capacity (100) with size (0) for primaries (200)' \
        in str(err.value)
    # too small for the secondaries (ExtendFromSecondariesAction.cc:89-95)
    with pytest.raises(cb.B200Error) as err:
        cb.celer_sim_run(dict(base, initializer_capacity=300))
    assert 'insufficient capacity (300) for track initializers' in str(err.value)
    # and large enough
    out = cb.celer_sim_run(dict(base, initializer_capacity=1 << 18))
    assert out['result']['runner']['num_aborted'] == [0]


def test_event_id_out_of_range_is_rejected():
    import celeritas_b200 as cb
    _, _, params, gpu = setup('testem3-small', 64)
    prim = cb.make_primaries(2, particle_id=params.find_particle(11), energy=10.0,
                             pos=(-22, 0, 0), direction=(1, 0, 0), event_of=lambda i: 64 + i)
    with pytest.raises(cb.B200Error) as err:
        gpu.step(prim)
    assert 'max_events' in str(err.value)


def test_corrupt_images_fail_at_load(tmp_path):
    """A truncated or hand-edited problem image is rejected by the loader (index columns are
    checked against the columns they point into) instead of becoming out-of-bounds device
    reads inside the kernels."""
    import celeritas_b200 as cb
    raw = open(data_path('images', 'testem3-small.b2img'), 'rb').read()

    def patched(name, value, index=0):
        key = name.encode()
        at = raw.index(len(key).to_bytes(4, 'little') + key) + 4 + len(key) + 12 + 4 * index
        return raw[:at] + int(value).to_bytes(4, 'little') + raw[at + 4:]

    cases = {
        'truncated': raw[:len(raw) // 2],
        'grid id out of range': patched('phys.pp_grid', 1000),
        'grid beyond the value pool': patched('phys.grid_value_offset', 1 << 30),
        'material id out of range': patched('geomat.volume_material', 77, index=3),
        'surface id range': patched('geo.vol_face_end', 1 << 20, index=5),
        'element id out of range': patched('mat.elcomp_element', 99),
    }
    for what, data in cases.items():
        path = tmp_path / 'bad.b2img'
        path.write_bytes(data)
        with pytest.raises(cb.B200Error) as e:
            cb.Params(str(path))
        assert 'image' in str(e.value), (what, str(e.value))
        with pytest.raises(cb.B200Error):
            cb.Params(image_bytes=data)
    # the untouched bytes load from memory as well as from the file
    assert cb.Params(image_bytes=raw).num_detectors == cb.Params(
        data_path('images', 'testem3-small.b2img')).num_detectors
