"""Device-resident step loop (csrc/tail.cu, Stepper::advance): several step iterations per
launch must leave exactly the state, the per-iteration counters and the tallies that one
host round trip per iteration leaves, i.e. what the reference's Stepper produces.

The reference (oracle/_ref, host Stepper) is stepped one iteration at a time; the CUDA stepper
takes its first iteration with the primaries through the per-action path and then advances K
iterations per call. After every call all per-slot fields are compared (integers and the six
RNG words exactly, reals to 1e-7), and every iteration's counters with the reference's.
"""
import json

import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu

NEVER = 0xffffffff


def setup(name, slots, **kw):
    import celeritas_b200 as cb
    import celerref
    cfg = json.load(open(data_path('images', name + '.json')))
    refp = celerref.Problem(cfg)
    ref = refp.stepper(slots)
    params = cb.Params(data_path('images', name + '.b2img'))
    gpu = cb.Stepper(params, slots, **kw)
    return refp, ref, params, gpu


def advance_lockstep(ref, gpu, prim, chunk, rtol=1e-7, max_iters=100000):
    from parity import compare_states
    cr, cg = ref.step(prim), gpu.step(prim)
    assert cr == cg
    hist = [cr]
    compare_states(ref, gpu, 0, rtol=rtol, atol=rtol)
    while (cr['alive'] or cr['queued']) and len(hist) < max_iters:
        got = gpu.advance(chunk)
        assert 1 <= len(got) <= chunk
        for cg in got:
            cr = ref.step()
            assert cr == cg, 'iteration %d: ref=%s gpu=%s' % (len(hist), cr, cg)
            hist.append(cr)
        compare_states(ref, gpu, len(hist), rtol=rtol, atol=rtol)
        if len(got) < chunk:
            assert not (cr['alive'] or cr['queued'])
    return hist


@pytest.mark.parametrize('chunk', [1, 5, 64])
@pytest.mark.parametrize('name', ['testem3-small', 'testem3-small-initcharge'])
def test_advance_matches_reference(name, chunk):
    """Whole 1 GeV showers; both slot-assignment orders (none: a dead parent's first
    secondary reuses its slot; init_charge: neutral tracks take the lowest vacancies)."""
    import celeritas_b200 as cb
    refp, ref, params, gpu = setup(name, 4096, tail_threshold=4096)
    prim = cb.make_primaries(2, particle_id=params.find_particle(11), energy=1000.0,
                             pos=(-22, 0, 0), direction=(1, 0, 0))
    hist = advance_lockstep(ref, gpu, prim, chunk)
    assert len(hist) > 100
    assert gpu.tail_iterations == len(hist) - 1
    assert np.allclose(refp.calo(100), gpu.calo(), rtol=1e-9, atol=1e-9)


def test_advance_switches_between_loop_and_per_action_path():
    """A threshold inside the shower's size range: the loop hands over to the per-action
    kernels when the shower grows (exit reason 'too many') and takes over again in the tail;
    the sorted vacancy array is rebuilt at every hand-over."""
    import celeritas_b200 as cb
    refp, ref, params, gpu = setup('testem3-small-initcharge', 4096, tail_threshold=100)
    prim = cb.make_primaries(3, particle_id=params.find_particle(11), energy=1000.0,
                             pos=(-22, 0, 0), direction=(1, 0, 0))
    hist = advance_lockstep(ref, gpu, prim, 9)
    sizes = [h['active'] for h in hist]
    assert max(sizes) > 300
    assert 0 < gpu.tail_iterations < len(hist) - 1


def test_advance_queued_initializers():
    """More tracks than slots: starts from the queue inside the loop (per-run vacancy
    search), initializer-queue counters identical."""
    import celeritas_b200 as cb
    refp, ref, params, gpu = setup('testem3-small-initcharge', 64, tail_threshold=4096)
    prim = cb.make_primaries(4, particle_id=params.find_particle(11), energy=200.0,
                             pos=(-22, 0, 0), direction=(1, 0, 0))
    hist = advance_lockstep(ref, gpu, prim, 16)
    assert max(h['queued'] for h in hist) > 20
    assert gpu.tail_iterations == len(hist) - 1


def test_advance_field_multilevel():
    """CMS-scale stand-in: four universe levels, 1 T field, looping tracks (reals at 1e-5 as
    in test_gpu_cms_scale.py: the gyration phase amplifies libm ulp differences)."""
    import celeritas_b200 as cb
    refp, ref, params, gpu = setup('cms-scale-small', 4096, tail_threshold=4096)
    opts = {'seed': 20220904, 'pdg': [11, 22], 'num_events': 1, 'primaries_per_event': 4,
            'energy': 200.0, 'position': [0, 0, 0], 'direction': {'distribution': 'isotropic'}}
    prim = refp.generate_primaries(opts)
    hist = advance_lockstep(ref, gpu, prim, 32, rtol=1e-5, max_iters=600)
    assert len(hist) > 100
    assert gpu.tail_iterations > 0


def test_run_events_same_with_and_without_loop():
    """The Transporter loop (b200_run_events) with the device-resident loop on and off:
    identical step, iteration and track counts and tallies."""
    import celeritas_b200 as cb
    params = cb.Params(data_path('images', 'testem3-small-initcharge.b2img'))
    prim = cb.make_primaries(8, particle_id=params.find_particle(11), energy=500.0,
                             pos=(-22, 0, 0), direction=(1, 0, 0),
                             event_of=lambda i: i // 2)
    offsets = np.arange(0, 9, 2, dtype=np.uint32)
    results = []
    for tail in (4096, NEVER):
        st = cb.Stepper(params, 8192, tail_threshold=tail)
        r = st.run_events(prim, offsets, merge_events=False)
        results.append((r, st.calo(), st.tail_iterations))
    (ra, ca, ta), (rb, cb_, tb) = results
    for k in ('num_steps', 'num_step_iterations', 'num_tracks', 'num_primaries', 'max_queued'):
        assert ra[k] == rb[k], k
    assert ta > 0 and tb == 0
    assert np.allclose(ca, cb_, rtol=1e-12, atol=0)
