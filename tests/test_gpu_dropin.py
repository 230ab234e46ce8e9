"""The COMPILED drop-in: celeritas_b200/adapter/B200Actions.cc is built against
the reference's headers and linked with the reference's own CUDA build
(oracle/_ref/libcelerref_dropin.so, `make -C oracle dropin`). Every B200 step action is a
CoreStepActionInterface adapter (src/corecel/sys/ActionInterface.hh:175-186) registered in an
ActionRegistry under the reference's ids, run by the reference's OWN ActionSequence
(src/celeritas/global/ActionSequence.cc:77-138) on the reference's OWN CoreState<device>,
whose AuxStateVec holds the SoA track state; the problem crosses in memory
(b200_params_create_from_memory). The stepper around it implements StepperInterface.

Checked against the reference's SimpleComptonTest golden values
(test/celeritas/global/Stepper.test.cc:194-209: 919 iterations, 53.8125 steps per primary,
queue high-water mark 6 at iteration 1) and, slot by slot, against the reference's host
Stepper run here.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import REPO, data_path

pytestmark = pytest.mark.gpu

DROPIN = os.path.join(REPO, 'oracle', '_ref', 'libcelerref_dropin.so')
SNAP_FIELDS = ['status', 'track_id', 'particle_id', 'volume_id', 'post_step_action', 'rng']


def run_dropin(tmp_path, cfg, slots, nprim, particle_id, energy, max_iters, every):
    assert os.path.exists(DROPIN), 'build it with `make -C oracle dropin`'
    cfg_path = tmp_path / 'config.json'
    cfg_path.write_text(json.dumps(cfg))
    out = tmp_path / 'dropin.npz'
    cmd = [sys.executable, os.path.join(REPO, 'tests', 'dropin_run.py'), str(cfg_path), str(out),
           str(slots), str(nprim), str(particle_id), str(energy), str(max_iters), str(every)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=REPO)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return np.load(out)


def reference_host(cfg, slots, prim, max_iters, every):
    """The reference's host Stepper on the same problem: history + state snapshots."""
    import celerref
    ref = celerref.Problem(cfg).stepper(slots)
    c = ref.step(prim)
    history, snaps, it = [], {}, 0
    while True:
        history.append([c['generated'], c['queued'], c['active'], c['alive']])
        if every and it % every == 0:
            for f in SNAP_FIELDS + ['energy']:
                snaps['%s_%d' % (f, it)] = ref.get(f)
        if not (c['queued'] > 0 or c['alive'] > 0) or it >= max_iters:
            break
        c = ref.step()
        it += 1
    return np.array(history), snaps


def compare(d, history, snaps):
    assert np.array_equal(d['history'], history)
    for key, want in snaps.items():
        got = d[key]
        field = key.rsplit('_', 1)[0]
        active = snaps['status_' + key.rsplit('_', 1)[1]] != 0
        if field == 'energy':
            assert np.allclose(got[active], want[active], rtol=1e-7, atol=1e-7), key
        elif field in ('track_id', 'particle_id'):
            assert np.array_equal(got[active], want[active]), key
        else:
            assert np.array_equal(got, want), key


def test_simple_compton_golden_through_reference_action_sequence(tmp_path):
    import celeritas_b200 as cb
    cfg = {'problem': 'simple-compton', 'geometry_file': 'data/geometry/two-boxes.org.json',
           'seed': 20220511}
    d = run_dropin(tmp_path, cfg, 64, 32, 0, 100.0, 100000, 50)
    h = d['history']
    active, queued = h[:, 2], h[:, 1]
    # test/celeritas/global/Stepper.test.cc:194-209
    assert len(h) == 919
    assert active.sum() / 32 == 53.8125
    assert (int(np.argmax(queued)), int(queued.max())) == (1, 6)
    # the sequence is the reference's: its ActionSequence sorted our adapters by (order, id)
    seq = list(d['sequence'])
    assert seq[0] == 'extend-from-primaries' and seq[-1] == 'extend-from-secondaries'
    assert seq.index('pre-step') < seq.index('along-step-neutral') \
        < seq.index('physics-discrete-select') < seq.index('scat-klein-nishina') \
        < seq.index('geo-boundary') < seq.index('tracking-cut')
    # ids and labels of the B200 registry are the reference registry's
    assert set(seq) <= set(d['registry'])
    # the reference CoreState's own counters were filled by the end-of-step adapter
    assert list(d['ref_counters'][[1, 3, 4]]) == [0, int(active[-1]), 0]
    assert int(d['launches'][0]) > 919 * 5
    prim = cb.make_primaries(32, particle_id=0, energy=100.0, pos=(-22, 0, 0),
                             direction=(1, 0, 0))
    history, snaps = reference_host(cfg, 64, prim, 100000, 50)
    compare(d, history, snaps)


def test_full_em_shower_through_reference_action_sequence(tmp_path):
    import celeritas_b200 as cb
    cfg = json.load(open(data_path('images', 'testem3-small.json')))
    params = cb.Params(data_path('images', 'testem3-small.b2img'))
    electron = params.find_particle(11)
    d = run_dropin(tmp_path, cfg, 4096, 2, electron, 1000.0, 100000, 10)
    prim = cb.make_primaries(2, particle_id=electron, energy=1000.0, pos=(-22, 0, 0),
                             direction=(1, 0, 0))
    history, snaps = reference_host(cfg, 4096, prim, 100000, 10)
    assert len(history) > 100
    compare(d, history, snaps)
