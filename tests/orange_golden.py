"""Golden tracking vectors of the reference's own ORANGE tests, restated as data:
/root/reference/test/orange/OrangeJson.test.cc:105-153 (UniversesTest.tracking) and :622-638
(HexArrayTest.track_out). Each entry: geometry, start position, direction, the expected volume
NAMES along the ray and the expected segment lengths."""
import json

from conftest import data_path

GOLDEN = [
    ('universes', (-1.0, -3.75, 0.75), (1, 0, 0), ['johnny', 'patty', 'c', 'johnny'],
     [1, 0.5, 5.5, 2]),
    ('universes', (-1, -2, 1.0), (1, 0, 0), ['johnny', 'c', 'a', 'b', 'c', 'johnny'],
     [1, 1, 2, 2, 1, 2]),
    ('universes', (4, -5, 1.0), (0, 1, 0), ['johnny', 'c', 'b', 'c', 'bobby', 'johnny'],
     [1, 1, 2, 1, 2, 2]),
    ('universes', (4, -2, -0.75), (0, 0, 1), ['johnny', 'b', 'b', 'johnny'], [0.25, 1, 1, 0.5]),
    ('hex-array', (-6.9258369494022292, -4.9982766629573767, -10.8378536157757495),
     (0.6750034206933703, -0.3679917428721818, 0.6394939086732125),
     ['interior', 'cfill', 'dfill', 'interior'],
     [1.9914318088046, 5.3060674310398, 0.30636846908014, 5.9880767678838]),
]


def volume_names(geometry):
    """Global volume id -> name: universes in file order, local volumes in order (rect
    arrays contribute one volume per cell), as UniverseInserter numbers them
    (/root/reference/src/orange/detail/UniverseInserter.cc)."""
    g = json.load(open(data_path('geometry', geometry + '.org.json')))
    names = []
    for u in g['universes']:
        if u['_type'] in ('unit', 'simple unit'):
            labels = u.get('volume_labels') or u.get('cell_names')
            names += [s.split('@')[0] for s in labels]
        else:
            n = (len(u['x']) - 1) * (len(u['y']) - 1) * (len(u['z']) - 1)
            names += ['{array cell}'] * n
    return names


def check_trace(trace, geometry, pos, direction, exp_names, exp_dist):
    import numpy as np
    vol, surf, dist, count, safety = trace([pos], [direction], 64)
    n = int(count[0]) & 0x7fffffff
    names = volume_names(geometry)
    got = [names[v] for v in vol[0, :n]]
    # the reference's test helper stops when the track leaves the geometry
    assert got[:len(exp_names)] == exp_names, got
    assert np.allclose(dist[0, :len(exp_dist)], exp_dist, rtol=1e-11, atol=0)
    assert n == len(exp_names) or got[len(exp_names)] == '[EXTERIOR]'
