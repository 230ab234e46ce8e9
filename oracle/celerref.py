"""ctypes wrapper over oracle/_ref/libcelerref.so (the reference's own host code).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline leg. Never by the product package.
"""
import ctypes as C
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, '_ref', 'libcelerref.so')


class Primary(C.Structure):
    """Same layout as B200Primary (include/celeritas_b200.h)."""
    _fields_ = [('particle_id', C.c_uint32), ('event_id', C.c_uint32),
                ('energy', C.c_double), ('pos', C.c_double * 3),
                ('dir', C.c_double * 3), ('time', C.c_double)]


PRIMARY_DTYPE = np.dtype([('particle_id', '<u4'), ('event_id', '<u4'), ('energy', '<f8'),
                          ('pos', '<f8', 3), ('dir', '<f8', 3), ('time', '<f8')])

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        # CELERREF_CUDA=1 selects the reference's own CUDA build (make -C oracle ref_cuda)
        path = LIB_PATH
        if os.environ.get('CELERREF_CUDA') == '1':
            path = LIB_PATH.replace('libcelerref.so', 'libcelerref_cuda.so')
        L = C.CDLL(path, mode=C.RTLD_LOCAL)
        L.celerref_last_error.restype = C.c_char_p
        L.celerref_problem_create.restype = C.c_void_p
        L.celerref_problem_create.argtypes = [C.c_char_p]
        L.celerref_problem_destroy.argtypes = [C.c_void_p]
        L.celerref_export_image.argtypes = [C.c_void_p, C.c_char_p]
        L.celerref_stepper_create.restype = C.c_void_p
        L.celerref_stepper_create.argtypes = [C.c_void_p, C.c_uint32]
        L.celerref_stepper_create_stream.restype = C.c_void_p
        L.celerref_stepper_create_stream.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.celerref_stepper_destroy.argtypes = [C.c_void_p]
        L.celerref_step.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.celerref_reseed.argtypes = [C.c_void_p, C.c_uint64]
        L.celerref_kill_active.argtypes = [C.c_void_p]
        L.celerref_state_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.celerref_calo_get.argtypes = [C.c_void_p, C.c_void_p]
        L.celerref_calo_clear.argtypes = [C.c_void_p]
        L.celerref_hits_count.argtypes = [C.c_void_p, C.c_uint32]
        L.celerref_hits_count.restype = C.c_uint32
        L.celerref_hits_get.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_void_p]
        L.celerref_geo_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                         C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
        L.celerref_diagnostic_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p,
                                              C.POINTER(C.c_uint32)]
        L.celerref_num_particles.argtypes = [C.c_void_p]
        L.celerref_action_labels.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
        L.celerref_generate_primaries.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p,
                                                  C.c_uint64, C.POINTER(C.c_uint64)]
        L.celerref_run_events.restype = C.c_double
        L.celerref_run_events.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                          C.c_uint32, C.c_int, C.c_void_p]
        L.celerref_run_merged_device.restype = C.c_double
        L.celerref_run_merged_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                                 C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().celerref_last_error().decode())


FIELDS = {
    'status': ('u1', 1), 'track_id': ('<u4', 1), 'parent_id': ('<u4', 1), 'event_id': ('<u4', 1),
    'num_steps': ('<u4', 1), 'num_looping_steps': ('<u4', 1), 'time': ('<f8', 1),
    'step_length': ('<f8', 1), 'post_step_action': ('<u4', 1), 'along_step_action': ('<u4', 1),
    'particle_id': ('<u4', 1), 'energy': ('<f8', 1), 'material_id': ('<u4', 1),
    'interaction_mfp': ('<f8', 1), 'macro_xs': ('<f8', 1), 'energy_deposition': ('<f8', 1),
    'dedx_range': ('<f8', 1), 'rng': ('<u4', 6), 'pos': ('<f8', 3), 'dir': ('<f8', 3),
    'volume_id': ('<u4', 1), 'surface_id': ('<u4', 1), 'geo_level': ('<u4', 1),
    'track_slots': ('<u4', 1),
}


HIT_FIELDS = {
    'detector': ('<u4', 1), 'track_id': ('<u4', 1), 'event_id': ('<u4', 1),
    'parent_id': ('<u4', 1), 'track_step_count': ('<u4', 1), 'particle': ('<u4', 1),
    'step_length': ('<f8', 1), 'energy_deposition': ('<f8', 1),
    'pre_time': ('<f8', 1), 'pre_energy': ('<f8', 1), 'pre_pos': ('<f8', 3), 'pre_dir': ('<f8', 3),
    'post_time': ('<f8', 1), 'post_energy': ('<f8', 1), 'post_pos': ('<f8', 3),
    'post_dir': ('<f8', 3),
}


class Problem:
    def __init__(self, config):
        config = dict(config)
        config.setdefault('base_dir', REPO)
        self.config = config
        self.h = lib().celerref_problem_create(json.dumps(config).encode())
        if not self.h:
            raise RuntimeError(lib().celerref_last_error().decode())

    def export_image(self, path):
        _check(lib().celerref_export_image(self.h, path.encode()))

    def stepper(self, num_track_slots, stream_id=0):
        return Stepper(self, num_track_slots, stream_id)

    def calo(self, n):
        out = np.zeros(n)
        _check(lib().celerref_calo_get(self.h, out.ctypes.data))
        return out

    def calo_clear(self):
        _check(lib().celerref_calo_clear(self.h))

    def hits(self, stream=0):
        """Step/hit output of the last step (the reference's DetectorStepOutput) as a dict
        of arrays; requires 'hit_volumes' in the problem configuration."""
        n = lib().celerref_hits_count(self.h, stream)
        out = {}
        for name, (dt, w) in HIT_FIELDS.items():
            a = np.zeros((n, w) if w > 1 else n, dtype=dt)
            if n:
                _check(lib().celerref_hits_get(self.h, stream, name.encode(), a.ctypes.data))
            out[name] = a
        return out

    def trace(self, pos, direction, max_segments=64):
        """Ray-trace with the reference's OrangeTrackView (same outputs as Params.trace)."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        direction = np.ascontiguousarray(direction, dtype=np.float64).reshape(-1, 3)
        n = len(pos)
        vol = np.full((n, max_segments), 0xffffffff, dtype=np.uint32)
        surf = np.full((n, max_segments), 0xffffffff, dtype=np.uint32)
        dist = np.zeros((n, max_segments))
        count = np.zeros(n, dtype=np.uint32)
        safety = np.zeros(n)
        _check(lib().celerref_geo_trace(self.h, pos.ctypes.data, direction.ctypes.data, n,
                                        max_segments, vol.ctypes.data, surf.ctypes.data,
                                        dist.ctypes.data, count.ctypes.data,
                                        safety.ctypes.data))
        return vol, surf, dist, count, safety

    def diagnostic(self, steps=False):
        """ActionDiagnostic (steps=False) or StepDiagnostic tallies: counts[particle][bin]."""
        nb = C.c_uint32()
        _check(lib().celerref_diagnostic_get(self.h, int(steps), None, C.byref(nb)))
        out = np.zeros((lib().celerref_num_particles(self.h), nb.value), dtype=np.uint32)
        _check(lib().celerref_diagnostic_get(self.h, int(steps), out.ctypes.data, C.byref(nb)))
        return out

    def action_labels(self):
        buf = C.create_string_buffer(1 << 16)
        _check(lib().celerref_action_labels(self.h, buf, len(buf)))
        return buf.value.decode().split('\n')[:-1]

    def generate_primaries(self, primary_options):
        """The reference's PrimaryGenerator on celer-sim `primary_options`."""
        text = json.dumps(primary_options).encode()
        count = C.c_uint64()
        _check(lib().celerref_generate_primaries(self.h, text, None, 0, C.byref(count)))
        out = np.zeros(count.value, dtype=PRIMARY_DTYPE)
        _check(lib().celerref_generate_primaries(self.h, text, out.ctypes.data, len(out),
                                                 C.byref(count)))
        return out

    def run_events(self, primaries, offsets, num_track_slots, num_threads=0):
        primaries = np.ascontiguousarray(primaries, dtype=PRIMARY_DTYPE)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        res = np.zeros(4, dtype=np.uint64)
        t = lib().celerref_run_events(self.h, primaries.ctypes.data, offsets.ctypes.data,
                                      len(offsets) - 1, num_track_slots, num_threads,
                                      res.ctypes.data)
        if t < 0:
            raise RuntimeError(lib().celerref_last_error().decode())
        return dict(seconds=t, num_steps=int(res[0]), num_step_iterations=int(res[1]),
                    num_primaries=int(res[2]), max_queued=int(res[3]))


def run_merged_device(problem, primaries, num_track_slots):
    """All primaries merged onto one device state, the reference's own CUDA kernels."""
    primaries = np.ascontiguousarray(primaries, dtype=PRIMARY_DTYPE)
    res = np.zeros(4, dtype=np.uint64)
    t = lib().celerref_run_merged_device(problem.h, primaries.ctypes.data, len(primaries),
                                         num_track_slots, res.ctypes.data)
    if t < 0:
        raise RuntimeError(lib().celerref_last_error().decode())
    return dict(seconds=t, num_steps=int(res[0]), num_step_iterations=int(res[1]),
                num_primaries=int(res[2]), max_queued=int(res[3]))


class Stepper:
    def __init__(self, problem, num_track_slots, stream_id=0):
        self.problem = problem
        self.n = num_track_slots
        self.h = lib().celerref_stepper_create_stream(problem.h, num_track_slots, stream_id)
        if not self.h:
            raise RuntimeError(lib().celerref_last_error().decode())

    def step(self, primaries=None):
        counts = np.zeros(4, dtype=np.uint32)
        if primaries is not None and len(primaries):
            primaries = np.ascontiguousarray(primaries, dtype=PRIMARY_DTYPE)
            _check(lib().celerref_step(self.h, primaries.ctypes.data, len(primaries),
                                       counts.ctypes.data))
        else:
            _check(lib().celerref_step(self.h, None, 0, counts.ctypes.data))
        return dict(generated=int(counts[0]), queued=int(counts[1]), active=int(counts[2]),
                    alive=int(counts[3]))

    def reseed(self, event_id):
        _check(lib().celerref_reseed(self.h, event_id))

    def kill_active(self):
        _check(lib().celerref_kill_active(self.h))

    def get(self, field):
        dt, w = FIELDS[field]
        out = np.zeros((self.n, w) if w > 1 else self.n, dtype=dt)
        _check(lib().celerref_state_get(self.h, field.encode(), out.ctypes.data))
        return out

    def __del__(self):
        try:
            lib().celerref_stepper_destroy(self.h)
        except Exception:
            pass


def make_primaries(n, particle_id=0, energy=100.0, pos=(0, 0, 0), direction=(1, 0, 0),
                   event_of=lambda i: 0):
    p = np.zeros(n, dtype=PRIMARY_DTYPE)
    p['particle_id'] = particle_id
    p['energy'] = energy
    p['pos'] = pos
    p['dir'] = direction
    p['event_id'] = [event_of(i) for i in range(n)]
    return p
