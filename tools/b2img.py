"""Reader of the .b2img problem-image container (celeritas_b200/host/Image.hh)."""
import struct

import numpy as np

_DTYPES = {0: 'u1', 1: '<u4', 2: '<i4', 3: '<f4', 4: '<f8', 5: '<u8'}


def read_image(path):
    raw = open(path, 'rb').read()
    assert raw[:8] == b'B2IMG\0\0\1'
    n, = struct.unpack_from('<I', raw, 8)
    pos, out = 12, {}
    for _ in range(n):
        ln, = struct.unpack_from('<I', raw, pos)
        name = raw[pos + 4:pos + 4 + ln].decode()
        dt, cnt = struct.unpack_from('<IQ', raw, pos + 4 + ln)
        dtype = np.dtype(_DTYPES[dt])
        nbytes = cnt * dtype.itemsize
        start = pos + 4 + ln + 12
        out[name] = np.frombuffer(raw, dtype=dtype, count=cnt, offset=start)
        pos = start + nbytes + (8 - nbytes % 8) % 8
    return out
