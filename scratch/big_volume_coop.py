"""Time the ray trace of the many-faces geometry (all rays cross the 113-face background
volume) with the library given by CELERITAS_B200_LIB (default: the cooperative build).
usage: python scratch/big_volume_coop.py [nrays]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import celeritas_b200 as cb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
gpu = cb.Params('data/images/geo-many-faces.b2img')
rng = np.random.default_rng(3)
pos = rng.uniform(-15, 15, size=(n, 3))
d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
gpu.trace(pos[:1024], d[:1024], 64)
best = 1e9
for _ in range(5):
    t = time.perf_counter(); out = gpu.trace(pos, d, 64); best = min(best, time.perf_counter() - t)
segs = int((out[3][out[3] != 0xffffffff] & 0x7fffffff).sum())
print('%s: %d rays, %d segments, best of 5: %.2f ms (host wall incl. copies), checksum %d'
      % (os.environ.get('CELERITAS_B200_LIB', 'default'), n, segs, best * 1e3, int(out[0].astype(np.uint64).sum())))
