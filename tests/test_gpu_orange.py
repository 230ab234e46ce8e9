"""GPU parity of ORANGE navigation on fixed ray sets: the sequence of volume ids, crossed
surface ids (bit-exact) and segment lengths of every ray, and the safety distance at the
origin, against the reference's OrangeTrackView on the host. Covers single- and multi-level
simple units (planes, cylinders, spheres, general quadrics), background volumes, complex
(internal-surface) volumes, translated daughters and rect arrays -- the reference's own
ORANGE test geometries (test/orange/data, test/geocel/data)."""
import json

import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu

GEOMETRIES = ['two-boxes', 'testem3-flat', 'testem3', 'simple-cms', 'five-volumes', 'universes',
              'rect-array', 'nested-rect-arrays', 'hex-array', 'three-spheres', 'testem15',
              'lar-sphere', 'four-steel-slabs', 'one-steel-sphere', 'cms-scale', 'many-faces']


def ray_set(name, n, seed):
    """Seeded rays: origins in the bounding region of the geometry, isotropic directions
    (generic, so no tangencies), plus a few axis-aligned rays."""
    g = json.load(open(data_path('geometry', name + '.org.json')))
    bbox = g['universes'][0].get('bbox') or [[-10, -10, -10], [10, 10, 10]]
    lo, hi = np.array(bbox[0], dtype=float), np.array(bbox[1], dtype=float)
    lo, hi = np.maximum(lo, -1e3), np.minimum(hi, 1e3)
    rng = np.random.default_rng(seed)
    pos = lo + (hi - lo) * (0.02 + 0.96 * rng.random((n, 3)))
    d = rng.normal(size=(n, 3))
    # half of the rays aim at a random point near the centre so small inner volumes are hit
    centre = 0.5 * (lo + hi)
    target = centre + 0.05 * (hi - lo) * rng.normal(size=(n, 3))
    aimed = rng.random(n) < 0.5
    d[aimed] = (target - pos)[aimed]
    d /= np.linalg.norm(d, axis=1)[:, None]
    return pos, d


@pytest.mark.parametrize('name', GEOMETRIES)
def test_trace_matches_reference(name):
    import celeritas_b200 as cb
    import celerref
    ref = celerref.Problem({'problem': 'geometry',
                            'geometry_file': 'data/geometry/%s.org.json' % name})
    gpu = cb.Params(data_path('images', 'geo-%s.b2img' % name))
    pos, d = ray_set(name, 4096, 42)
    rv, rs, rd, rc, rsafe = ref.trace(pos, d, 256)
    gv, gs, gd, gc, gsafe = gpu.trace(pos, d, 256)
    assert np.array_equal(rc, gc), 'segment counts differ for rays %s' % np.nonzero(rc != gc)[0][:8]
    located = rc != 0xffffffff
    assert located.sum() > 1000
    assert np.array_equal(rv, gv), 'volume id sequences differ'
    assert np.array_equal(rs, gs), 'surface id sequences differ'
    # distances: same IEEE operation sequence (only + - * / sqrt), so identical
    assert np.array_equal(rd, gd)
    assert np.array_equal(rsafe, gsafe)
    # the ray set actually exercises the geometry
    assert (rc[located] & 0x7fffffff).max() >= 2


def test_reference_golden_tracks():
    """The reference's own golden tracks (test/orange/OrangeJson.test.cc:105-153, 622-638)
    on the GPU: volume names and segment lengths (tests/orange_golden.py)."""
    import celeritas_b200 as cb
    from orange_golden import GOLDEN, check_trace
    for geometry, pos, direction, names, dist in GOLDEN:
        gpu = cb.Params(data_path('images', 'geo-%s.b2img' % geometry))
        check_trace(gpu.trace, geometry, pos, direction, names, dist)


@pytest.mark.parametrize('name', ['cms-scale', 'many-faces', 'universes', 'nested-rect-arrays',
                                  'hex-array', 'testem3'])
def test_trace_with_natively_built_geometry(name):
    """The geometry built by the library's own ORANGE construction from the .org.json file
    (celeritas_b200/host/OrangeBuilder.cpp; tests/test_cpu_orange_builder.py shows it equals
    the reference-built image column for column) traced against the reference."""
    import celeritas_b200 as cb
    import celerref
    ref = celerref.Problem({'problem': 'geometry',
                            'geometry_file': 'data/geometry/%s.org.json' % name})
    gpu = cb.Params(org_json=data_path('geometry', name + '.org.json'))
    pos, d = ray_set(name, 2048, 11)
    rv, rs, rd, rc, rsafe = ref.trace(pos, d, 256)
    gv, gs, gd, gc, gsafe = gpu.trace(pos, d, 256)
    assert np.array_equal(rc, gc) and np.array_equal(rv, gv) and np.array_equal(rs, gs)
    assert np.array_equal(rd, gd) and np.array_equal(rsafe, gsafe)
    assert (rc != 0xffffffff).sum() > 500
