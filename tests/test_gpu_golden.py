"""Replay of the committed golden fixtures (tests/golden/*.json, produced from the reference
by tests/golden/make_golden.py): per-iteration counters, CRCs of every integer state field
and of the XORWOW words, and the calorimeter tallies. Needs no reference at run time."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import REPO, data_path

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(REPO, 'tests', 'golden')
sys.path.insert(0, GOLDEN)
CASES = sorted(f[:-5] for f in os.listdir(GOLDEN) if f.endswith('.json'))


@pytest.mark.parametrize('fuse', [0, 0xffffffff], ids=['fused', 'per-action'])
@pytest.mark.parametrize('name', CASES)
def test_replay_golden(name, fuse):
    import celeritas_b200 as cb
    from make_golden import primaries_for, state_crcs
    gold = json.load(open(os.path.join(GOLDEN, name + '.json')))
    case = gold['case']
    params = cb.Params(data_path('images', case['image'] + '.b2img'))
    # the fixture's particle numbering is the image's
    for pdg, pid in gold['particle_ids'].items():
        found = params.find_particle(int(pdg))
        assert found is None or found == pid
    gpu = cb.Stepper(params, case['slots'], fuse_threshold=fuse)
    prim = primaries_for(case, params.find_particle, cb.make_primaries)
    c = gpu.step(prim)
    for it, want in enumerate(gold['steps']):
        got = dict(c)
        crc = want.get('crc')
        assert got == {k: want[k] for k in got}, 'iteration %d counters' % it
        if crc is not None:
            assert state_crcs(gpu) == crc, 'iteration %d state CRCs' % it
        if it + 1 < len(gold['steps']):
            c = gpu.step()
    assert not (c['alive'] or c['queued'])
    if 'calo' in gold:
        assert np.allclose(gpu.calo(), gold['calo'], rtol=1e-9, atol=1e-9)


def test_fixtures_cover_the_configurations():
    assert {'simple-compton', 'testem3-small', 'testem3-small-initcharge',
            'simple-cms-em-field'} <= set(CASES)
    gold = json.load(open(os.path.join(GOLDEN, 'simple-compton.json')))
    # the reference's own gold numbers (test/celeritas/global/Stepper.test.cc:194-209)
    steps = gold['steps']
    assert len(steps) == 919
    assert sum(s['active'] for s in steps) / 32 == 53.8125
