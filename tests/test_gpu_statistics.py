"""Statistical agreement of the per-layer energy-deposition tallies when the RNG streams
are NOT the reference's: the GPU transports merged events on two concurrent streams (as
bench.py does), the reference transports event by event on its host Steppers. The tallies
cannot match bin by bin; they must be two samples of the same distribution.

Test: batches of showers on both sides; per bin the batch means are compared with Welch's
t statistic; the sum of t^2 over the bins is the chi-square. (The same-stream comparison,
per-bin rtol 1e-9, is test_gpu_testem3.py::test_calo_tally_matches_reference.)
"""
import json

import numpy as np
import pytest

from conftest import data_path

pytestmark = pytest.mark.gpu

NUM_BATCHES = 16
PRIMARIES_PER_BATCH = 64
ENERGY = 1000.0


def test_calo_tallies_chi_square_vs_reference():
    import celeritas_b200 as cb
    import celerref
    cfg = json.load(open(data_path('images', 'testem3.json')))
    cfg['max_streams'] = 8
    cfg['initializer_capacity'] = 1 << 22
    refp = celerref.Problem(cfg)
    params = cb.Params(data_path('images', 'testem3.b2img'))
    ndet = params.num_detectors
    electron = params.find_particle(11)
    steppers = [cb.Stepper(params, 1 << 16, stream_id=k) for k in range(2)]

    def batch(b):
        # 16 events of 4 primaries; distinct event ids per batch (RNG reseed key)
        n = PRIMARIES_PER_BATCH
        prim = cb.make_primaries(n, particle_id=electron, energy=ENERGY, pos=(-22, 0, 0),
                                 direction=(1, 0, 0), event_of=lambda i: b * 16 + i // 4)
        return prim, np.arange(0, n + 1, 4, dtype=np.uint32)

    gpu, ref = [], []
    gpu_steps = ref_steps = 0
    for b in range(NUM_BATCHES):
        prim, offsets = batch(b)
        for st in steppers:
            st.calo_clear()
        res, _ = cb.run_events_streams(steppers, prim, offsets, merge_events=True)
        gpu.append(sum(st.calo() for st in steppers))
        gpu_steps += sum(r['num_steps'] for r in res)
        refp.calo_clear()
        r = refp.run_events(prim, offsets, 4096, 8)
        ref.append(refp.calo(ndet))
        ref_steps += r['num_steps']
    gpu, ref = np.array(gpu), np.array(ref)

    # total deposited energy: both within 1% of each other and most of the beam energy
    total_g, total_r = gpu.sum(), ref.sum()
    beam = NUM_BATCHES * PRIMARIES_PER_BATCH * ENERGY
    assert total_r > 0.9 * beam and total_g > 0.9 * beam
    assert abs(total_g - total_r) / total_r < 0.01
    # track-steps per primary agree within 1%
    assert abs(gpu_steps - ref_steps) / ref_steps < 0.01

    mg, mr = gpu.mean(axis=0), ref.mean(axis=0)
    vg = gpu.var(axis=0, ddof=1) / NUM_BATCHES
    vr = ref.var(axis=0, ddof=1) / NUM_BATCHES
    use = (mr > 1e-3 * mr.max()) & (vg + vr > 0)
    assert use.sum() >= 60
    t2 = (mg[use] - mr[use]) ** 2 / (vg[use] + vr[use])
    ndf = int(use.sum())
    chi2 = float(t2.sum())
    # Welch t^2 with ~2(B-1) degrees of freedom has mean ~1.07; for ndf ~ 100 bins the
    # sum exceeds 1.6 ndf with probability < 1e-4 if the distributions are the same
    assert chi2 / ndf < 1.6, 'chi2/ndf = %.2f over %d bins' % (chi2 / ndf, ndf)
    # per-bin relative tolerance on the batch means: 5 standard errors
    assert np.all(np.abs(mg[use] - mr[use]) < 5 * np.sqrt(vg[use] + vr[use]) + 1e-9)
