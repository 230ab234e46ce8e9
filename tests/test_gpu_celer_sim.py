"""GPU parity of the celer-sim front end and its diagnostics against the reference:

* PrimaryGenerator (phys/PrimaryGenerator.cc): identical primaries, byte for byte;
* ActionDiagnostic / StepDiagnostic (user/detail/*DiagnosticExecutor.hh): identical tallies;
* Transporter / RunnerOutput (app/celer-sim/Transporter.cc, RunnerOutput.cc): identical
  per-iteration track counts, step/track totals and calorimeter tallies for a celer-sim
  JSON input run through b200_celer_sim_run.
"""
import json

import numpy as np
import pytest

from conftest import REPO, data_path

pytestmark = pytest.mark.gpu

PRIMARY_OPTIONS = {
    '_format': 'primary-generator', 'seed': 12345, 'pdg': [11, 22, -11],
    'num_events': 5, 'primaries_per_event': 7,
    'energy': {'distribution': 'delta', 'params': [150.0]},
    'position': {'distribution': 'box', 'params': [-22, -5, -5, -21, 5, 5]},
    'direction': {'distribution': 'isotropic'},
}


def reference_problem(name, **extra):
    import celerref
    cfg = json.load(open(data_path('images', name + '.json')))
    cfg.update(extra)
    return celerref.Problem(cfg)


def test_primary_generator_matches_reference():
    import celeritas_b200 as cb
    refp = reference_problem('testem3-small')
    params = cb.Params(data_path('images', 'testem3-small.b2img'))
    for opts in (PRIMARY_OPTIONS,
                 dict(PRIMARY_OPTIONS, pdg=22, energy=10.0, position=[0, 0, 0],
                      direction=[0, 0, 1], seed=0),
                 dict(PRIMARY_OPTIONS, position=[-22, 0, 0], num_events=64,
                      primaries_per_event=33, seed=987654321)):
        expected = refp.generate_primaries(opts)
        actual, offsets = params.generate_primaries(opts)
        assert len(actual) == opts['num_events'] * opts['primaries_per_event']
        assert offsets[-1] == len(actual)
        assert actual.tobytes() == expected.tobytes()


def test_primary_generator_errors_match_reference():
    """Same validation as PrimaryGeneratorOptions.cc:22-50,61-123."""
    import celeritas_b200 as cb
    params = cb.Params(data_path('images', 'testem3-small.b2img'))
    refp = reference_problem('testem3-small')
    bad = [dict(PRIMARY_OPTIONS, energy={'distribution': 'box', 'params': [1, 2]}),
           dict(PRIMARY_OPTIONS, position={'distribution': 'box', 'params': [1, 2, 3]}),
           dict(PRIMARY_OPTIONS, direction={'distribution': 'box', 'params': [0] * 6}),
           dict(PRIMARY_OPTIONS, direction={'distribution': 'isotropic', 'params': [1.0]}),
           dict(PRIMARY_OPTIONS, _format='celer-sim')]
    for opts in bad:
        with pytest.raises(RuntimeError) as ref_err:
            refp.generate_primaries(opts)
        with pytest.raises(cb.B200Error) as err:
            params.generate_primaries(opts)
        # our message is the reference's validation message
        ref_msg = str(ref_err.value)
        msg = str(err.value).split(': ', 1)[1]
        assert msg in ref_msg, (msg, ref_msg)


def run_reference(refp, primaries, offsets, slots, max_steps=0):
    """The celer-sim loop on the reference's host Stepper: events in turn on one stream,
    no reseeding (app/celer-sim/Transporter.cc:84-179)."""
    ref = refp.stepper(slots)
    events = []
    for e in range(len(offsets) - 1):
        counts = ref.step(primaries[offsets[e]:offsets[e + 1]])
        hist = [counts]
        remaining = max_steps
        while counts['alive'] or counts['queued']:
            if max_steps:
                remaining -= 1
                if remaining == 0:
                    break
            counts = ref.step()
            hist.append(counts)
        events.append(hist)
    return ref, events


@pytest.mark.parametrize('fuse', [0, 0xffffffff], ids=['fused', 'per-action'])
def test_diagnostics_match_reference(fuse):
    import celeritas_b200 as cb
    slots, bins = 2048, 30
    refp = reference_problem('testem3-small', action_diagnostic=True, step_diagnostic_bins=bins)
    params = cb.Params(data_path('images', 'testem3-small.b2img'))
    gpu = cb.Stepper(params, slots, action_diagnostic=True, step_diagnostic_bins=bins,
                     fuse_threshold=fuse)
    assert 'action-diagnostic' in gpu.step_action_labels
    assert 'step-diagnostic' in gpu.step_action_labels
    opts = dict(PRIMARY_OPTIONS, num_events=1, primaries_per_event=6, position=[-22, 0, 0],
                direction=[1, 0, 0])
    prim, offsets = params.generate_primaries(opts)
    _, events = run_reference(refp, prim, offsets, slots)
    counts = gpu.step(prim)
    n = 1
    while counts['alive'] or counts['queued']:
        counts = gpu.step()
        n += 1
    assert n == len(events[0])
    # steps per track: identical [particle][bin] table
    ref_steps = refp.diagnostic(steps=True)
    assert ref_steps.shape == (3, bins + 2)
    assert ref_steps.sum() > 1000
    assert ref_steps[:, bins + 1].sum() > 0  # the overflow bin is exercised
    assert np.array_equal(ref_steps, gpu.step_diagnostic())
    # post-step actions: identical counts per (particle, action label)
    ref_actions = refp.diagnostic(steps=False)
    gpu_actions = gpu.action_diagnostic()
    ref_labels, gpu_labels = refp.action_labels(), gpu.all_action_labels
    assert sorted(ref_labels) == sorted(gpu_labels)
    assert ref_actions.sum() == sum(h['active'] for h in events[0])
    used = 0
    for a, label in enumerate(ref_labels):
        assert np.array_equal(ref_actions[:, a], gpu_actions[:, gpu_labels.index(label)]), label
        used += bool(ref_actions[:, a].any())
    assert used >= 8  # msc/eloss/boundary/discrete models of three particle types
    gpu.diagnostics_clear()
    assert gpu.action_diagnostic().sum() == 0 and gpu.step_diagnostic().sum() == 0


@pytest.mark.parametrize('order', ['none', 'init_charge'])
@pytest.mark.parametrize('merge', [False, True])
def test_celer_sim_run_matches_reference(merge, order):
    """`track_order` is a run option (RunnerInput.hh:123): the SAME image runs with either
    slot assignment and must match the reference built with that order."""
    import celeritas_b200 as cb
    slots = 4096
    opts = dict(PRIMARY_OPTIONS, num_events=3, primaries_per_event=4, pdg=[11, 22])
    cfg = json.load(open(data_path('images', 'testem3-small.json')))
    run_input = {
        '_format': 'celer-sim', 'use_device': True,
        'image_file': 'data/images/testem3-small.b2img', 'base_dir': REPO,
        'geometry_file': cfg['geometry_file'], 'physics_file': cfg['physics_file'],
        'primary_options': opts, 'seed': cfg['seed'], 'num_track_slots': slots,
        'initializer_capacity': cfg['initializer_capacity'], 'secondary_stack_factor': 3,
        'simple_calo': cfg['simple_calo'], 'action_diagnostic': True, 'step_diagnostic': True,
        'step_diagnostic_bins': 50, 'merge_events': merge, 'action_times': True,
        'warm_up': True, 'track_order': order,
    }
    out = cb.celer_sim_run(run_input)
    runner = out['result']['runner']
    assert out['input']['track_order'] == order

    refp = reference_problem('testem3-small', action_diagnostic=True, step_diagnostic_bins=50,
                             track_order=order)
    prim = refp.generate_primaries(opts)
    offsets = [0, len(prim)] if merge else list(range(0, len(prim) + 1, 4))
    ref = refp.stepper(slots)
    ref.step()  # warm-up iteration (Runner::warm_up)
    events = []
    for e in range(len(offsets) - 1):
        counts = ref.step(prim[offsets[e]:offsets[e + 1]])
        hist = [counts]
        while counts['alive'] or counts['queued']:
            counts = ref.step()
            hist.append(counts)
        events.append(hist)

    assert runner['num_streams'] == 1
    assert runner['num_track_slots'] == [slots] * len(events)
    assert runner['num_step_iterations'] == [len(h) for h in events]
    assert runner['num_steps'] == [sum(c['active'] for c in h) for h in events]
    assert runner['max_queued'] == [max(c['queued'] for c in h) for h in events]
    assert runner['num_aborted'] == [0] * len(events)
    for key, ref_key in (('active', 'active'), ('alive', 'alive'), ('generated', 'generated'),
                         ('initializers', 'queued')):
        assert runner[key] == [[c[ref_key] for c in h] for h in events], key
    assert len(runner['time']['steps']) == len(events)
    assert runner['time']['actions']['pre-step'] > 0
    # tracks created: primaries + all secondaries (cumulative over events, as the
    # reference's per-event counters are never reset between events)
    total_tracks = runner['num_tracks'][-1]
    assert total_tracks == refp.diagnostic(steps=True).sum()
    # diagnostics and calorimeter
    assert np.array_equal(np.array(out['result']['step-diagnostic']['steps']),
                          refp.diagnostic(steps=True))
    ref_actions = refp.diagnostic(steps=False)
    ref_labels = refp.action_labels()
    gpu_labels = out['internal']['actions']['label']
    gpu_actions = np.array(out['result']['action-diagnostic']['actions'])
    for a, label in enumerate(ref_labels):
        assert np.array_equal(ref_actions[:, a], gpu_actions[:, gpu_labels.index(label)]), label
    calo = out['result']['simple_calo']
    assert calo['volume_labels'] == cfg['simple_calo']
    assert np.allclose(calo['energy_deposition'], refp.calo(len(cfg['simple_calo'])),
                       rtol=1e-9, atol=1e-9)


def test_celer_sim_max_steps_aborts_like_reference():
    """max_steps counts step iterations; leftover tracks are reported as aborted and the
    state is reset (Transporter.cc:133-141,166-174)."""
    import celeritas_b200 as cb
    slots = 512
    opts = dict(PRIMARY_OPTIONS, num_events=2, primaries_per_event=2, pdg=[11],
                position=[-22, 0, 0], direction=[1, 0, 0])
    cfg = json.load(open(data_path('images', 'testem3-small.json')))
    run_input = {
        'use_device': True, 'image_file': data_path('images', 'testem3-small.b2img'),
        'geometry_file': cfg['geometry_file'], 'primary_options': opts, 'seed': cfg['seed'],
        'num_track_slots': slots, 'initializer_capacity': cfg['initializer_capacity'],
        'secondary_stack_factor': 3, 'max_steps': 25, 'warm_up': False,
        'track_order': 'none',
    }
    runner = cb.celer_sim_run(run_input)['result']['runner']
    refp = reference_problem('testem3-small')
    prim = refp.generate_primaries(opts)
    ref = refp.stepper(slots)
    counts = ref.step(prim[:2])
    hist = [counts]
    for _ in range(24):
        counts = ref.step()
        hist.append(counts)
    assert runner['num_step_iterations'] == [25, 25]
    assert runner['active'][0] == [c['active'] for c in hist]
    assert runner['num_aborted'][0] == hist[-1]['alive'] + hist[-1]['queued'] > 0


def test_celer_sim_input_errors():
    import celeritas_b200 as cb
    base = {'use_device': True, 'image_file': data_path('images', 'testem3-small.b2img'),
            'geometry_file': 'x', 'primary_options': PRIMARY_OPTIONS, 'num_track_slots': 64,
            'initializer_capacity': 1024, 'secondary_stack_factor': 3}
    for change, fragment in (
            ({'primary_options': None}, 'either a event filename or options'),
            ({'event_file': 'events.hepmc3'}, 'but not both'),
            ({'initializer_capacity': None}, 'initializer_capacity'),
            ({'use_device': False}, 'no host track loop'),
            ({'num_track_slots': 0}, 'nonpositive num_track_slots'),
            ({'field_options': {'minimum_step': 1}}, "'field_options' cannot be specified"),
            ({'field': [0, 0, 1]}, 'without a uniform-field'),
            ({'_format': 'other'}, 'invalid format'),
            ({'step_limiter': 0.5}, "'step_limiter'"),
            ({'brem_combined': True}, "'brem_combined'"),
            ({'simple_calo': ['world']}, 'simple_calo')):
        inp = dict(base)
        for k, v in change.items():
            if v is None:
                inp.pop(k)
            else:
                inp[k] = v
        with pytest.raises(cb.B200Error) as err:
            cb.celer_sim_run(inp)
        assert fragment in str(err.value), (change, str(err.value))


def test_celer_sim_builds_the_geometry_from_geometry_file(tmp_path):
    """`geometry_file` is read: the ORANGE JSON file is built natively and
    replaces the image's geometry columns; a file that describes another geometry than the
    one the image's materials were exported for is refused."""
    import celeritas_b200 as cb
    cfg = json.load(open(data_path('images', 'testem3-small.json')))
    run_input = {
        '_format': 'celer-sim', 'use_device': True,
        'image_file': 'data/images/testem3-small.b2img', 'base_dir': REPO,
        'geometry_file': cfg['geometry_file'], 'physics_file': cfg['physics_file'],
        'primary_options': dict(PRIMARY_OPTIONS, num_events=1, primaries_per_event=2, pdg=[11]),
        'seed': cfg['seed'], 'num_track_slots': 2048,
        'initializer_capacity': cfg['initializer_capacity'], 'secondary_stack_factor': 3,
        'simple_calo': cfg['simple_calo'],
    }
    good = cb.celer_sim_run(run_input)
    assert good['result']['runner']['num_steps'][0] > 100
    # the same run with the geometry columns of the image (file missing -> image geometry)
    moved = dict(run_input, geometry_file='data/geometry/does-not-exist.org.json')
    same = cb.celer_sim_run(moved)
    assert same['result']['runner']['num_steps'] == good['result']['runner']['num_steps']
    assert same['result']['runner']['active'] == good['result']['runner']['active']
    wrong = dict(run_input, geometry_file='data/geometry/simple-cms.org.json')
    with pytest.raises(cb.B200Error) as e:
        cb.celer_sim_run(wrong)
    assert 'geometry_file' in str(e.value)


def test_celer_sim_reads_physics_file():
    """`physics_file` is read: a reference ROOT export is decoded by the
    library's own reader and must hold the particles and elements the image's tables were
    built for. lar-sphere-combined.b2img was exported from data/physics/lar-sphere-em.json,
    derived from the decoded form of the reference's test/celeritas/data/lar-sphere.root."""
    import celeritas_b200 as cb
    cfg = json.load(open(data_path('images', 'lar-sphere-combined.json')))
    run_input = {
        '_format': 'celer-sim', 'use_device': True,
        'image_file': 'data/images/lar-sphere-combined.b2img', 'base_dir': REPO,
        'geometry_file': cfg['geometry_file'],
        'physics_file': 'tests/golden/root/lar-sphere.root',
        'primary_options': {
            '_format': 'primary-generator', 'seed': 1, 'pdg': [11, 22], 'num_events': 1,
            'primaries_per_event': 8,
            'energy': {'distribution': 'delta', 'params': [100.0]},
            'position': {'distribution': 'delta', 'params': [0, 0, 0]},
            'direction': {'distribution': 'isotropic'}},
        'seed': cfg['seed'], 'num_track_slots': 2048,
        'initializer_capacity': cfg['initializer_capacity'], 'secondary_stack_factor': 3,
        'simple_calo': cfg['simple_calo'], 'brem_combined': True, 'max_steps': 100000,
    }
    out = cb.celer_sim_run(run_input)
    assert out['result']['runner']['num_steps'][0] > 50
    assert out['result']['runner']['num_aborted'] == [0]
    assert 'particles' in out['internal']['physics_file']
    as_json = cb.celer_sim_run(dict(run_input, physics_file=cfg['physics_file']))
    assert as_json['result']['runner']['num_steps'] == out['result']['runner']['num_steps']
    # silicon / lead / iron ...: not the physics of an image with liquid-argon tables
    with pytest.raises(cb.B200Error) as e:
        cb.celer_sim_run(dict(run_input, physics_file='tests/golden/root/simple-cms.root'))
    assert 'physics_file' in str(e.value)
