"""CPU tests (no GPU): the oracle is pinned against the reference's own golden vectors.

 * oracle/restate.py (pure-Python restatement) vs KATs copied from the reference's tests;
 * oracle/_ref (the reference's own host code, compiled by oracle/Makefile) vs the golden
   values of the reference's stepping-loop tests, and vs the restatement.
"""
import math
import os

import numpy as np
import pytest

import restate
from conftest import REPO, data_path

try:
    import celerref
    HAVE_REF = celerref.available()
except Exception:  # pragma: no cover
    HAVE_REF = False
needs_ref = pytest.mark.skipif(not HAVE_REF, reason='oracle/_ref/libcelerref.so not built')

# test/celeritas/random/XorwowRngEngine.test.cc:156-171 (seed 12345, stream 0, 8 states)
XORWOW_INITIAL = [
    2421091215, 3647994171, 2504472727, 1236778574, 4083156575, 63361926, 3719645674,
    843467800, 1265623178, 295820715, 1583721852, 802677129, 3794549800, 1642707272,
    4266580851, 2668696688, 2910059606, 1707659088, 3955349927, 2857721444, 2773100230,
    3321656875, 1176613630, 909057096, 4173021154, 338389676, 2806912494, 1345761716,
    149057928, 630801564, 3118211368, 3857808320, 4193588147, 925742588, 1585365047,
    3244057179, 3428095051, 118856847, 945254054, 2395966273, 1370167352, 1607766504,
    3084411954, 2675509253, 2542521715, 327503606, 3527767224, 154218656]
# XorwowRngEngine.test.cc:183-190 (stream 1, first 8 words)
XORWOW_STREAM1 = [600837418, 1595898312, 3746176631, 2544092812, 689723186, 2087379088,
                  2231971747, 2290977355]


def test_xorwow_initial_states_golden():
    states = restate.initial_xorwow_states(12345, 0, 8)
    assert [w for s in states for w in s] == XORWOW_INITIAL


def test_xorwow_initial_states_stream1_golden():
    states = restate.initial_xorwow_states(12345, 1, 2)
    assert [w for s in states for w in s][:8] == XORWOW_STREAM1


def _jump_tables():
    """Jump polynomials travel as data in every problem image (rng.params)."""
    import struct
    raw = open(data_path('images', 'simple-compton.b2img'), 'rb').read()
    key = b'rng.params'
    i = raw.index(key) + len(key)
    dtype, count = struct.unpack_from('<IQ', raw, i)
    vals = struct.unpack_from('<%dI' % count, raw, i + 12)
    assert dtype == 1 and count == 321
    jump = [list(vals[1 + 5 * k:6 + 5 * k]) for k in range(32)]
    jump_sub = [list(vals[161 + 5 * k:166 + 5 * k]) for k in range(32)]
    return vals[0], jump, jump_sub


def test_xorwow_jump_equals_stepping():
    """XorwowRngEngine.test.cc:208-240: initialising at offset n == calling next() n times."""
    _, jump, jump_sub = _jump_tables()
    rng = restate.Xorwow.from_seed(12345, 0, 0, jump, jump_sub)
    for offset in range(0, 300):
        skip = restate.Xorwow.from_seed(12345, 0, offset, jump, jump_sub)
        assert rng() == skip()
    rng = restate.Xorwow.from_seed(12345, 0, 0, jump, jump_sub)
    skip = restate.Xorwow.from_seed(12345, 0, 0, jump, jump_sub)
    for count in (4, 21, 170, 6553):
        skip.discard(count, jump)
        for _ in range(count):
            rng()
        assert rng() == skip()


def test_klein_nishina_golden():
    """test/celeritas/em/KleinNishina.test.cc:88-138: 10 MeV photon along +z, mt19937."""
    rng = restate.Mt19937()
    canonical = lambda: restate.canonical_std(rng)
    inv_mass = 1 / 0.5109989461
    exp_e = [0.4581502636229, 1.325852509857, 9.837250571445, 0.5250297816972]
    exp_cos = [-0.0642523962721, 0.6656882878883, 0.9991545931877, 0.07782377978055]
    exp_ee = [9.541849736377, 8.674147490143, 0.1627494285554, 9.474970218303]
    exp_cose = [0.998962567429, 0.9941635460938, 0.3895748042313, 0.9986216572142]
    for i in range(4):
        e, d, ee, de, dep = restate.klein_nishina(10.0, [0, 0, 1], inv_mass, canonical)
        assert e == pytest.approx(exp_e[i], rel=1e-11)
        assert d[2] == pytest.approx(exp_cos[i], rel=1e-10)
        assert ee == pytest.approx(exp_ee[i], rel=1e-11)
        assert de[2] == pytest.approx(exp_cose[i], rel=1e-10)


def test_xs_calculator_golden():
    """test/celeritas/grid/XsCalculator.test.cc:37-139 (simple, scaled_lowest, scaled_middle)."""
    def grid(emin, emax, n, f):
        front = math.log(emin)
        delta = (math.log(emax) - front) / (n - 1)
        vals = [f(math.exp(front + delta * i)) for i in range(n)]
        return front, math.log(emax), vals, delta

    front, back, vals, delta = grid(1.0, 1e5, 6, lambda e: e)
    no_scaling = 0xFFFFFFFF
    for e, want in [(1, 1.0), (1e2, 1e2), (1e5 - 1e-6, 1e5 - 1e-6), (1e5, 1e5), (5, 5),
                    (0.0001, 1.0), (1e7, 1e5)]:
        assert restate.calc_xs(front, back, vals, no_scaling, e) == pytest.approx(want, rel=1e-12)
    # values of 1, scaled by E from index 0
    front, back, vals, delta = grid(0.1, 1e4, 6, lambda e: 1.0)
    vals = [v * math.exp(front + delta * i) for i, v in enumerate(vals)]
    for e, want in [(0.1, 1), (1e2, 1), (1e4, 1), (0.2, 1), (5, 1), (0.0001, 1000), (1e5, 0.1)]:
        assert restate.calc_xs(front, back, vals, 0, e) == pytest.approx(want, rel=1e-12)
    # values of 3, scaled by E from index 3
    front, back, vals, delta = grid(0.1, 1e4, 6, lambda e: 3.0)
    vals = [v * (math.exp(front + delta * i) if i >= 3 else 1) for i, v in enumerate(vals)]
    for e, want in [(0.1, 3), (1e2, 3), (1e4, 3), (0.2, 3), (5, 3), (0.0001, 3), (1e5, 0.3)]:
        assert restate.calc_xs(front, back, vals, 3, e) == pytest.approx(want, rel=1e-12)


def test_range_calculators_golden():
    """test/celeritas/grid/RangeCalculator.test.cc:20-62 and InverseRangeCalculator.test.cc:
    20-72: energy grid 10 .. 1e4 MeV in 4 points, range = E / 20."""
    front, back = math.log(10.0), math.log(1e4)
    delta = (back - front) / 3
    values = [0.05 * math.exp(front + delta * i) for i in range(4)]
    for e, want in [(1, 0.5 * math.sqrt(1 / 10.)), (2, 0.5 * math.sqrt(2 / 10.)), (10, 0.5),
                    (20, 1.0), (100, 5.0), (1e4, 500), (1.001e4, 500)]:
        assert restate.calc_range(front, back, values, e) == pytest.approx(want, rel=1e-12)
    values[-1] = 500.0
    for r, want in [(0.5 * math.sqrt(1 / 10.), 1.0), (0.5 * math.sqrt(2 / 10.), 2.0), (0.5, 10.0),
                    (1, 20.0), (5, 100.0), (500, 1e4)]:
        assert restate.calc_inverse_range(front, back, values, r) == pytest.approx(want, rel=1e-12)


def test_logic_evaluator():
    """Logic strings of the reference's two-boxes geometry (data/geometry/two-boxes.org.json)."""
    inner = restate.parse_logic('0 1 ~ & 2 & 3 ~ & 4 & 5 ~ &')
    # inside the box: outside(+) of the lower planes, inside(-) of the upper planes
    assert restate.eval_logic(inner, [1, 0, 1, 0, 1, 0])
    assert not restate.eval_logic(inner, [0, 0, 1, 0, 1, 0])
    exterior = restate.parse_logic('0 1 ~ & 2 & 3 ~ & 4 & 5 ~ & ~')
    assert not restate.eval_logic(exterior, [1, 0, 1, 0, 1, 0])
    assert restate.eval_logic(exterior, [1, 1, 1, 0, 1, 0])


# --------------------------------------------------------------------------- #
# The reference itself (oracle/_ref)
# --------------------------------------------------------------------------- #
SIMPLE_COMPTON = {'problem': 'simple-compton',
                  'geometry_file': 'data/geometry/two-boxes.org.json', 'seed': 20220511}


@needs_ref
def test_ref_simple_compton_golden():
    """test/celeritas/global/Stepper.test.cc:194-209: 919 iterations, 53.8125 steps/primary,
    initializer queue high-water mark 6 at iteration 1."""
    step = celerref.Problem(SIMPLE_COMPTON).stepper(64)
    prim = celerref.make_primaries(32, energy=100.0, pos=(-22, 0, 0), direction=(1, 0, 0))
    c = step.step(prim)
    active, queued = [c['active']], [c['queued']]
    while c['queued'] > 0 or c['alive'] > 0:
        c = step.step()
        active.append(c['active'])
        queued.append(c['queued'])
    assert len(active) == 919
    assert sum(active) / 32 == 53.8125
    assert (queued.index(max(queued)), max(queued)) == (1, 6)


@needs_ref
def test_ref_initial_rng_matches_restatement():
    step = celerref.Problem(SIMPLE_COMPTON).stepper(8)
    want = restate.initial_xorwow_states(20220511, 0, 8)
    assert step.get('rng').tolist() == want


@needs_ref
def test_ref_reseed_matches_restatement():
    """reseed_rng: subsequence = event * num_slots + slot (random/RngReseed.cu:38-41)."""
    seed, jump, jump_sub = _jump_tables()
    assert seed == 20220511
    step = celerref.Problem(SIMPLE_COMPTON).stepper(4)
    for event in (0, 3, 123):
        step.reseed(event)
        got = step.get('rng').tolist()
        for slot in range(4):
            r = restate.Xorwow.from_seed(seed, event * 4 + slot, 0, jump, jump_sub)
            assert got[slot] == r.state()


@needs_ref
def test_ref_first_mfp_draw_matches_restatement():
    """After one step the per-slot RNG has advanced by whole canonical draws; the first draw
    is the interaction MFP = -log(xi) (phys/detail/PreStepExecutor.hh:78-83)."""
    step = celerref.Problem(SIMPLE_COMPTON).stepper(4)
    init = step.get('rng').tolist()
    prim = celerref.make_primaries(4, energy=100.0, pos=(-22, 0, 0), direction=(1, 0, 0))
    step.step(prim)
    # primaries fill vacancies from the back: thread i -> slot 3 - i
    mfp0 = step.get('interaction_mfp')
    xs = step.get('macro_xs')
    step_len = step.get('step_length')
    for slot in range(4):
        r = restate.Xorwow(init[slot])
        mfp = -math.log(r.canonical())
        # mfp after the first step = mfp - step * xs (TrackUpdater)
        assert mfp0[slot] == pytest.approx(mfp - step_len[slot] * xs[slot], rel=1e-14)


PRIMARY_OPTIONS = {
    '_format': 'primary-generator', 'seed': 12345, 'pdg': [11, 22, -11],
    'num_events': 5, 'primaries_per_event': 7,
    'energy': {'distribution': 'delta', 'params': [150.0]},
    'position': {'distribution': 'box', 'params': [-22, -5, -5, -21, 5, 5]},
    'direction': {'distribution': 'isotropic'},
}


def test_primary_generator_restatement_properties():
    """test/celeritas/phys/PrimaryGenerator.test.cc:56-132: particle/event id pattern,
    box bounds and unit directions."""
    opts = dict(PRIMARY_OPTIONS, pdg=[22, 11], num_events=2, primaries_per_event=3,
                energy=10.0, position=[1, 2, 3], seed=0)
    prim = restate.generate_primaries(opts, [0, 1])
    assert [p['particle_id'] for p in prim] == [0, 1, 0, 0, 1, 0]
    assert [p['event_id'] for p in prim] == [0, 0, 0, 1, 1, 1]
    assert all(p['energy'] == 10 and p['pos'] == [1, 2, 3] and p['time'] == 0 for p in prim)
    opts = dict(PRIMARY_OPTIONS, pdg=[22], num_events=1, primaries_per_event=10, energy=1.0,
                position={'distribution': 'box', 'params': [-3, -3, -3, 3, 3, 3]}, seed=0)
    for p in restate.generate_primaries(opts, [0]):
        assert all(-3 <= x <= 3 for x in p['pos'])
        assert abs(sum(x * x for x in p['dir']) - 1) < 1e-14


@needs_ref
def test_ref_primary_generator_matches_restatement():
    """The reference's PrimaryGenerator (compiled from its sources) vs the restatement:
    identical doubles (the same mt19937 stream and the same arithmetic)."""
    import json
    cfg = json.load(open(data_path('images', 'testem3-small.json')))
    refp = celerref.Problem(cfg)
    for opts in (PRIMARY_OPTIONS, dict(PRIMARY_OPTIONS, seed=2**31 + 7, num_events=40)):
        got = refp.generate_primaries(opts)
        ids = [{11: 0, 22: 1, -11: 2}[p] for p in opts['pdg']]
        # particle ids of this problem: e-, gamma, e+ in the order of the image
        want = restate.generate_primaries(opts, ids)
        assert len(got) == len(want)
        pid_of = {}
        for g, w in zip(got, want):
            pid_of.setdefault(w['particle_id'], int(g['particle_id']))
            assert pid_of[w['particle_id']] == int(g['particle_id'])
            assert int(g['event_id']) == w['event_id']
            assert float(g['energy']) == w['energy']
            assert g['pos'].tolist() == w['pos']
            assert g['dir'].tolist() == w['dir']


@needs_ref
def test_golden_fixtures_match_the_reference():
    """tests/golden/*.json are what oracle/_ref produces today (regenerate with
    tests/golden/make_golden.py): re-run the two smallest cases and compare everything."""
    import json
    import sys
    sys.path.insert(0, os.path.join(REPO, 'tests', 'golden'))
    from make_golden import primaries_for, state_crcs
    for name in ('simple-cms-em-field', 'testem3-small-initcharge'):
        gold = json.load(open(os.path.join(REPO, 'tests', 'golden', name + '.json')))
        case = gold['case']
        cfg = json.load(open(data_path('images', case['image'] + '.json')))
        problem = celerref.Problem(cfg)
        stepper = problem.stepper(case['slots'])
        ids = {int(k): v for k, v in gold['particle_ids'].items()}
        prim = primaries_for(case, lambda pdg: ids[pdg], celerref.make_primaries)
        c = stepper.step(prim)
        for it, want in enumerate(gold['steps']):
            assert c == {k: want[k] for k in c}, (name, it)
            if 'crc' in want:
                assert state_crcs(stepper) == want['crc'], (name, it)
            if it + 1 < len(gold['steps']):
                c = stepper.step()
        assert np.allclose(problem.calo(len(gold['calo'])), gold['calo'], rtol=0, atol=0)


@needs_ref
def test_ref_orange_tracking_golden():
    """The oracle's ray trace (the reference's OrangeTrackView through oracle/ref_harness)
    reproduces the golden tracks of the reference's own ORANGE tests
    (test/orange/OrangeJson.test.cc:105-153, 622-638): volume names and segment lengths."""
    import celerref
    from orange_golden import GOLDEN, check_trace
    for geometry, pos, direction, names, dist in GOLDEN:
        ref = celerref.Problem({'problem': 'geometry',
                                'geometry_file': 'data/geometry/%s.org.json' % geometry})
        check_trace(ref.trace, geometry, pos, direction, names, dist)
