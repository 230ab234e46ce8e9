"""Volumes beyond the register path's 32 faces / 32 intersections (csrc/orange.cuh, "big
volumes": warp-cooperative distance search). The reference sizes its per-track scratch from
the geometry's max_faces / max_intersections (orange/OrangeData.hh:348-544,
OrangeTrackView.hh:1042-1067), so real detector mother volumes with hundreds of faces load;
so must this library.

Geometry (tools/make_many_faces.py, built by the reference's OrangeParams): max_faces = 113,
max_intersections = 180. A background volume whose faces are all 113 surfaces of its unit
(SimpleUnitTracker::background_intersect), a convex polyhedron of 40 general planes
(simple_intersect), a box with 30 spherical holes as one internal-surface volume with 36 faces
and up to 66 intersections (complex_intersect)."""
import json

import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu

NEVER = 0xffffffff


def test_image_exceeds_register_path_limits():
    import celeritas_b200 as cb
    gpu = cb.Params(data_path('images', 'geo-many-faces.b2img'))
    import sys
    sys.path.insert(0, data_path('..', 'tools'))
    import b2img
    img = b2img.read_image(data_path('images', 'geo-many-faces.b2img'))
    assert list(img['geo.scalars'][1:3]) == [113, 180]
    assert gpu.max_depth == 1


def test_rays_cross_every_kind_of_big_volume():
    """Bit-identical traces (volume ids, surface ids, distances, safeties) on rays aimed
    through the polyhedron, the cheese and across the background volume."""
    import celeritas_b200 as cb
    import celerref
    ref = celerref.Problem({'problem': 'geometry',
                            'geometry_file': 'data/geometry/many-faces.org.json'})
    gpu = cb.Params(data_path('images', 'geo-many-faces.b2img'))
    labels = gpu.volume_labels
    rng = np.random.default_rng(7)
    n = 6144
    pos = rng.uniform(-17, 17, size=(n, 3))
    pos[: n // 3] = np.array([16.0, 0, 0]) + rng.uniform(-2.0, 2.0, size=(n // 3, 3))  # polyhedron
    pos[n // 3: 2 * n // 3] = np.array([0, 0, -17.0]) + rng.uniform(-4.5, 4.5,
                                                                   size=(n // 3, 3))  # cheese
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    rv, rs, rd, rc, rsafe = ref.trace(pos, d, 64)
    gv, gs, gd, gc, gsafe = gpu.trace(pos, d, 64)
    assert np.array_equal(rc, gc)
    assert np.array_equal(rv, gv)
    assert np.array_equal(rs, gs)
    assert np.array_equal(rd, gd)
    assert np.array_equal(rsafe, gsafe)
    first = gv[:, 0]
    names = [labels[v] for v in np.unique(first[first != 0xffffffff])]
    for want in ('polyhedron', 'cheese', 'mother'):
        assert want in names, names
    # rays that start in the cheese pass holes (internal crossings that do not leave it)
    visited = set(labels[v] for v in np.unique(gv[gv != 0xffffffff]))
    assert any(v.startswith('hole') for v in visited) and any(v.startswith('ball') for v in visited)


@pytest.mark.parametrize('fuse', [0, NEVER], ids=['fused', 'per-action'])
def test_showers_inside_the_113_face_volume(fuse):
    """Full-EM 200 MeV showers that develop INSIDE the background volume (liquid argon):
    every step's distance-to-boundary is a big-volume search. Lock-step with the reference."""
    import celeritas_b200 as cb
    import celerref
    from parity import lockstep
    cfg = json.load(open(data_path('images', 'many-faces.json')))
    ref = celerref.Problem(cfg).stepper(2048)
    params = cb.Params(data_path('images', 'many-faces.b2img'))
    gpu = cb.Stepper(params, 2048, fuse_threshold=fuse)
    e = params.find_particle(11)
    prim = np.concatenate([
        cb.make_primaries(3, particle_id=e, energy=200.0, pos=(0.3, 0.2, -28), direction=(0, 0, 1)),
        cb.make_primaries(3, particle_id=e, energy=200.0, pos=(-25, 0.5, 3.1), direction=(1, 0, 0)),
        cb.make_primaries(2, particle_id=params.find_particle(22), energy=200.0,
                          pos=(25, 0.1, 0.2), direction=(-1, 0, 0))])
    hist = lockstep(ref, gpu, prim, compare_every=3)
    assert len(hist) > 50 and max(h['active'] for h in hist) > 100


def test_device_resident_loop_with_big_volumes():
    """The tail loop's one-warp-per-track mode meets the big-volume search (a group of 32
    lanes holding the same track)."""
    import celeritas_b200 as cb
    import celerref
    from test_gpu_tail import advance_lockstep
    cfg = json.load(open(data_path('images', 'many-faces.json')))
    ref = celerref.Problem(cfg).stepper(1024)
    params = cb.Params(data_path('images', 'many-faces.b2img'))
    gpu = cb.Stepper(params, 1024, tail_threshold=1024)
    prim = cb.make_primaries(2, particle_id=params.find_particle(11), energy=50.0,
                             pos=(0.3, 0.2, -28), direction=(0, 0, 1))
    hist = advance_lockstep(ref, gpu, prim, 16)
    assert gpu.tail_iterations > 0 and len(hist) > 20
