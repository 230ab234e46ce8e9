"""Native ORANGE construction (celeritas_b200/host/OrangeBuilder.cpp): the
geometry image built from an .org.json file by this library must be, column for column, the
image the reference's OrangeParams / UnitInserter / RectArrayInserter / BIHBuilder produced
for the same file (data/images/geo-*.b2img, exported through the oracle harness). Host-only:
runs without a GPU."""
import glob
import os
import sys

import numpy as np
import pytest

from conftest import REPO, data_path

sys.path.insert(0, os.path.join(REPO, 'tools'))

GEOMETRIES = sorted(os.path.basename(p)[4:-6]
                    for p in glob.glob(data_path('images', 'geo-*.b2img')))


def parse(raw):
    import struct
    assert raw[:8] == b'B2IMG\0\0\1'
    n, = struct.unpack_from('<I', raw, 8)
    pos, out = 12, {}
    dtypes = {0: 'u1', 1: '<u4', 2: '<i4', 3: '<f4', 4: '<f8', 5: '<u8'}
    for _ in range(n):
        ln, = struct.unpack_from('<I', raw, pos)
        name = raw[pos + 4:pos + 4 + ln].decode()
        dt, cnt = struct.unpack_from('<IQ', raw, pos + 4 + ln)
        dtype = np.dtype(dtypes[dt])
        start = pos + 4 + ln + 12
        out[name] = np.frombuffer(raw, dtype=dtype, count=cnt, offset=start)
        nbytes = cnt * dtype.itemsize
        pos = start + nbytes + (8 - nbytes % 8) % 8
    return out


def test_there_are_geometries():
    assert len(GEOMETRIES) >= 16 and 'cms-scale' in GEOMETRIES and 'many-faces' in GEOMETRIES


@pytest.mark.parametrize('name', GEOMETRIES)
def test_native_image_equals_reference_image(name):
    import celeritas_b200 as cb
    want = parse(open(data_path('images', 'geo-%s.b2img' % name), 'rb').read())
    got = parse(cb.orange_build_image(data_path('geometry', name + '.org.json')))
    columns = sorted(k for k in want if k.startswith('geo.'))
    assert len(columns) >= 38
    for key in columns:
        assert key in got, key
        a, b = want[key], got[key]
        assert a.dtype == b.dtype and a.shape == b.shape, '%s: %s %s vs %s %s' % (
            key, a.dtype, a.shape, b.dtype, b.shape)
        # bit patterns: -0.0 / 0.0 and infinities in the grids must match too
        assert a.tobytes() == b.tobytes(), '%s differs at %s' % (
            key, np.nonzero(a != b)[0][:8])


def test_builder_rejects_bad_input(tmp_path):
    import json
    import celeritas_b200 as cb
    good = json.load(open(data_path('geometry', 'many-faces.org.json')))
    cases = []
    bad = json.loads(json.dumps(good))
    bad['_format'] = 'other'
    cases.append((bad, 'format'))
    bad = json.loads(json.dumps(good))
    bad['universes'][0]['volumes'][1]['logic'] = '0 1 &  &'
    cases.append((bad, 'logic'))
    bad = json.loads(json.dumps(good))
    bad['universes'][0]['surfaces']['types'][0] = 'zz'
    cases.append((bad, 'surface type'))
    bad = json.loads(json.dumps(good))
    bad['universes'][0]['volumes'][1]['faces'] = [999]
    cases.append((bad, 'faces'))
    for j, fragment in cases:
        path = tmp_path / 'bad.org.json'
        path.write_text(json.dumps(j))
        with pytest.raises(cb.B200Error) as e:
            cb.orange_build_image(str(path))
        assert fragment in str(e.value), str(e.value)
    with pytest.raises(cb.B200Error):
        cb.orange_build_image(str(tmp_path / 'missing.org.json'))
