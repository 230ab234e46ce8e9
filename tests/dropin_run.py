"""Child process of tests/test_gpu_dropin.py: drives the COMPILED drop-in
(oracle/_ref/libcelerref_dropin.so = celeritas_b200/adapter/B200Actions.cc built against the
reference's headers and CUDA build) and writes what it saw to an .npz file.

Runs in its own process because the reference's CUDA build (libcelerref_cuda.so) and its host
build (libcelerref.so, the checker the parent uses) define the same C++ symbols.

usage: dropin_run.py <config.json> <out.npz> <slots> <nprim> <particle id> <energy> <max_iters>
           <compare_every>
"""
import ctypes as C
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'oracle'))
os.environ['CELERREF_CUDA'] = '1'

STATE_FIELDS = ['status', 'track_id', 'particle_id', 'volume_id', 'post_step_action', 'energy',
                'rng']


def main():
    cfg_path, out_path = sys.argv[1], sys.argv[2]
    slots, nprim, particle_id = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    energy, max_iters, every = float(sys.argv[6]), int(sys.argv[7]), int(sys.argv[8])
    import celerref
    import celeritas_b200 as cb
    from celeritas_b200.lib import FIELDS
    celerref.lib()  # the reference's own CUDA build (CELERREF_CUDA=1)
    D = C.CDLL(os.path.join(REPO, 'oracle', '_ref', 'libcelerref_dropin.so'))
    D.celerref_dropin_last_error.restype = C.c_char_p
    D.celerref_dropin_create.restype = C.c_void_p
    D.celerref_dropin_create.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    D.celerref_dropin_destroy.argtypes = [C.c_void_p]
    D.celerref_dropin_step.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    D.celerref_dropin_warm_up.argtypes = [C.c_void_p]
    D.celerref_dropin_sequence_labels.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
    D.celerref_dropin_launch_count.argtypes = [C.c_void_p]
    D.celerref_dropin_launch_count.restype = C.c_uint64
    D.celerref_dropin_state_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    D.celerref_dropin_counters.argtypes = [C.c_void_p, C.c_void_p]

    def check(rc):
        if rc != 0:
            raise RuntimeError(D.celerref_dropin_last_error().decode())

    problem = celerref.Problem(json.load(open(cfg_path)))
    h = D.celerref_dropin_create(problem.h, slots, 0)
    if not h:
        raise RuntimeError(D.celerref_dropin_last_error().decode())
    buf = C.create_string_buffer(4096)
    D.celerref_dropin_sequence_labels(h, buf, 4096)
    sequence = buf.value.decode().split()

    registry = problem.action_labels()  # the reference's own ActionRegistry
    prim = cb.make_primaries(nprim, particle_id=particle_id, energy=energy, pos=(-22, 0, 0),
                             direction=(1, 0, 0))
    check(D.celerref_dropin_warm_up(h))
    counts = np.zeros(4, dtype=np.uint32)
    history = []
    snaps = {}

    def get(field):
        dt, w = FIELDS[field]
        out = np.zeros((slots, w) if w > 1 else slots, dtype=dt)
        check(D.celerref_dropin_state_get(h, field.encode(), out.ctypes.data))
        return out

    check(D.celerref_dropin_step(h, prim.ctypes.data, len(prim), counts.ctypes.data))
    it = 0
    while True:
        history.append(counts.copy())
        if every and it % every == 0:
            for f in STATE_FIELDS:
                snaps['%s_%d' % (f, it)] = get(f)
        if not (counts[1] > 0 or counts[3] > 0) or it >= max_iters:
            break
        check(D.celerref_dropin_step(h, None, 0, counts.ctypes.data))
        it += 1
    ref_counters = np.zeros(5, dtype=np.uint32)
    check(D.celerref_dropin_counters(h, ref_counters.ctypes.data))
    launches = int(D.celerref_dropin_launch_count(h))
    np.savez(out_path, history=np.array(history), ref_counters=ref_counters,
             launches=np.array([launches]), sequence=np.array(sequence),
             registry=np.array(registry), **snaps)
    D.celerref_dropin_destroy(h)
    print('dropin ok: %d iterations, %d launches' % (len(history), launches))


if __name__ == '__main__':
    main()
