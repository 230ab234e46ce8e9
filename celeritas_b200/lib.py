"""ctypes binding of include/celeritas_b200.h."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class B200Error(RuntimeError):
    pass


class Primary(C.Structure):
    _fields_ = [('particle_id', C.c_uint32), ('event_id', C.c_uint32),
                ('energy', C.c_double), ('pos', C.c_double * 3),
                ('dir', C.c_double * 3), ('time', C.c_double)]


PRIMARY_DTYPE = np.dtype([('particle_id', '<u4'), ('event_id', '<u4'), ('energy', '<f8'),
                          ('pos', '<f8', 3), ('dir', '<f8', 3), ('time', '<f8')])
assert PRIMARY_DTYPE.itemsize == C.sizeof(Primary)


class StepperResult(C.Structure):
    _fields_ = [('generated', C.c_uint32), ('queued', C.c_uint32), ('active', C.c_uint32),
                ('alive', C.c_uint32)]


class RunResult(C.Structure):
    _fields_ = [('num_steps', C.c_uint64), ('num_step_iterations', C.c_uint64),
                ('num_primaries', C.c_uint64), ('max_queued', C.c_uint64),
                ('seconds', C.c_double), ('num_tracks', C.c_uint64),
                ('num_aborted', C.c_uint64)]


class StepperOptions(C.Structure):
    _fields_ = [('stream_id', C.c_uint32), ('num_track_slots', C.c_uint32),
                ('action_times', C.c_int), ('action_diagnostic', C.c_int),
                ('step_diagnostic_bins', C.c_uint32), ('fuse_threshold', C.c_uint32),
                ('tail_threshold', C.c_uint32)]


# Every symbol declared in include/celeritas_b200.h
EXPORTS = [
    'b200_last_error', 'b200_device_count', 'b200_params_create_from_image',
    'b200_params_destroy', 'b200_params_view', 'b200_params_num_actions',
    'b200_params_action_label', 'b200_params_num_volumes', 'b200_params_volume_label',
    'b200_params_num_detectors', 'b200_params_find_particle', 'b200_state_create',
    'b200_state_destroy', 'b200_state_view', 'b200_state_get', 'b200_state_calo_get',
    'b200_state_calo_clear', 'b200_step_extend_from_primaries', 'b200_step_initialize_tracks',
    'b200_step_pre_step', 'b200_step_along_step', 'b200_step_discrete_select',
    'b200_step_interact', 'b200_step_boundary', 'b200_step_tracking_cut', 'b200_step_tally',
    'b200_step_extend_from_secondaries', 'b200_reseed', 'b200_reset_generated',
    'b200_kill_active', 'b200_launch_count', 'b200_stepper_create', 'b200_stepper_destroy',
    'b200_stepper_state', 'b200_stepper_step', 'b200_stepper_warm_up', 'b200_stepper_reseed',
    'b200_stepper_kill_active', 'b200_stepper_num_step_actions',
    'b200_stepper_step_action_label', 'b200_stepper_launch_count', 'b200_run_events',
    'b200_stepper_set_action_times', 'b200_stepper_action_time', 'b200_set_device',
    'b200_geo_trace', 'b200_geo_trace_host',
    'b200_step_action_diagnostic', 'b200_step_step_diagnostic', 'b200_stepper_create_opts',
    'b200_stepper_num_actions', 'b200_stepper_action_label',
    'b200_stepper_action_diagnostic_get', 'b200_stepper_step_diagnostic_get',
    'b200_stepper_step_diagnostic_bins', 'b200_stepper_diagnostics_clear',
    'b200_primaries_generate', 'b200_celer_sim_run', 'b200_string_free',
    'b200_params_num_particles', 'b200_run_events_streams', 'b200_step_fused',
    'b200_step_post_tail', 'b200_step_along_select', 'b200_params_num_models', 'b200_params_model_action_begin',
    'b200_params_max_depth', 'b200_step_tail_loop', 'b200_tail_max_blocks',
    'b200_stepper_advance', 'b200_stepper_tail_iterations',
    'b200_params_create_from_memory', 'b200_stepper_insert', 'b200_stepper_begin_iteration',
    'b200_stepper_end_iteration', 'b200_stepper_stream', 'b200_step_sort_tracks',
    'b200_step_gather_hits', 'b200_stepper_hits_count', 'b200_stepper_hits_get',
    'b200_orange_build_image', 'b200_params_create_from_org_json', 'b200_import_root',
]

_lib = None


def library_path():
    # CELERITAS_B200_LIB selects another build of the same CUDA library (kernel tuning
    # experiments); it is never a fallback implementation
    return os.environ.get('CELERITAS_B200_LIB') or os.path.join(HERE, 'libceleritas_b200.so')


def load_library():
    """Load the CUDA extension; raise (never fall back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise B200Error('%s is missing: run `make -C celeritas_b200` (or '
                        '__graft_entry__.build()); there is no fallback path' % path)
    L = C.CDLL(path)
    vp = C.c_void_p
    L.b200_last_error.restype = C.c_char_p
    L.b200_device_count.restype = C.c_int
    L.b200_params_create_from_image.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.b200_params_create_from_memory.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(vp)]
    L.b200_params_destroy.argtypes = [vp]
    L.b200_orange_build_image.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.b200_import_root.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.b200_params_create_from_org_json.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.b200_string_free.argtypes = [vp]
    L.b200_stepper_insert.argtypes = [vp, vp, C.c_uint32]
    L.b200_stepper_begin_iteration.argtypes = [vp]
    L.b200_stepper_end_iteration.argtypes = [vp, C.POINTER(StepperResult)]
    L.b200_stepper_stream.argtypes = [vp]
    L.b200_stepper_stream.restype = vp
    L.b200_step_sort_tracks.argtypes = [vp, vp, C.c_uint32, vp]
    L.b200_stepper_hits_count.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.b200_stepper_hits_get.argtypes = [vp, C.c_char_p, vp]
    L.b200_params_view.restype = vp
    L.b200_params_view.argtypes = [vp]
    L.b200_params_num_actions.argtypes = [vp]
    L.b200_params_num_actions.restype = C.c_uint32
    L.b200_params_action_label.argtypes = [vp, C.c_uint32]
    L.b200_params_action_label.restype = C.c_char_p
    L.b200_params_num_volumes.argtypes = [vp]
    L.b200_params_num_volumes.restype = C.c_uint32
    L.b200_params_volume_label.argtypes = [vp, C.c_uint32]
    L.b200_params_volume_label.restype = C.c_char_p
    L.b200_params_num_detectors.argtypes = [vp]
    L.b200_params_num_detectors.restype = C.c_uint32
    L.b200_params_find_particle.argtypes = [vp, C.c_int]
    L.b200_params_find_particle.restype = C.c_uint32
    L.b200_params_num_particles.argtypes = [vp]
    L.b200_params_num_particles.restype = C.c_uint32
    L.b200_params_num_models.argtypes = [vp]
    L.b200_params_num_models.restype = C.c_uint32
    L.b200_params_model_action_begin.argtypes = [vp]
    L.b200_params_model_action_begin.restype = C.c_uint32
    L.b200_params_max_depth.argtypes = [vp]
    L.b200_params_max_depth.restype = C.c_uint32
    L.b200_state_create.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    L.b200_state_destroy.argtypes = [vp]
    L.b200_state_view.argtypes = [vp]
    L.b200_state_view.restype = vp
    L.b200_state_get.argtypes = [vp, C.c_char_p, vp]
    L.b200_state_calo_get.argtypes = [vp, vp]
    L.b200_state_calo_clear.argtypes = [vp]
    L.b200_launch_count.restype = C.c_uint64
    L.b200_stepper_create.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    L.b200_stepper_destroy.argtypes = [vp]
    L.b200_stepper_state.argtypes = [vp]
    L.b200_stepper_state.restype = vp
    L.b200_stepper_step.argtypes = [vp, vp, C.c_uint32, C.POINTER(StepperResult)]
    L.b200_stepper_warm_up.argtypes = [vp]
    L.b200_stepper_reseed.argtypes = [vp, C.c_uint64]
    L.b200_stepper_kill_active.argtypes = [vp]
    L.b200_stepper_num_step_actions.argtypes = [vp]
    L.b200_stepper_num_step_actions.restype = C.c_uint32
    L.b200_stepper_step_action_label.argtypes = [vp, C.c_uint32]
    L.b200_stepper_step_action_label.restype = C.c_char_p
    L.b200_stepper_advance.argtypes = [vp, C.c_uint32, C.POINTER(StepperResult),
                                       C.POINTER(C.c_uint32)]
    L.b200_stepper_tail_iterations.argtypes = [vp]
    L.b200_stepper_tail_iterations.restype = C.c_uint64
    L.b200_stepper_launch_count.argtypes = [vp]
    L.b200_stepper_launch_count.restype = C.c_uint64
    L.b200_stepper_set_action_times.argtypes = [vp, C.c_int]
    L.b200_stepper_action_time.argtypes = [vp, C.c_uint32]
    L.b200_stepper_action_time.restype = C.c_double
    L.b200_set_device.argtypes = [C.c_int]
    L.b200_geo_trace_host.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp]
    L.b200_run_events.argtypes = [vp, vp, vp, C.c_uint32, C.c_int, C.c_uint64,
                                  C.POINTER(RunResult)]
    L.b200_run_events_streams.argtypes = [vp, C.c_uint32, vp, vp, C.c_uint32, C.c_int,
                                          C.c_uint64, vp, C.POINTER(C.c_double)]
    L.b200_stepper_create_opts.argtypes = [vp, C.POINTER(StepperOptions), C.POINTER(vp)]
    L.b200_stepper_num_actions.argtypes = [vp]
    L.b200_stepper_num_actions.restype = C.c_uint32
    L.b200_stepper_action_label.argtypes = [vp, C.c_uint32]
    L.b200_stepper_action_label.restype = C.c_char_p
    L.b200_stepper_action_diagnostic_get.argtypes = [vp, vp]
    L.b200_stepper_step_diagnostic_get.argtypes = [vp, vp]
    L.b200_stepper_step_diagnostic_bins.argtypes = [vp]
    L.b200_stepper_step_diagnostic_bins.restype = C.c_uint32
    L.b200_stepper_diagnostics_clear.argtypes = [vp]
    L.b200_primaries_generate.argtypes = [vp, C.c_char_p, vp, C.c_uint64,
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    L.b200_celer_sim_run.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.b200_string_free.argtypes = [vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise B200Error('celeritas_b200 error %d: %s'
                        % (rc, load_library().b200_last_error().decode()))


def device_count():
    return load_library().b200_device_count()


def set_device(device):
    _check(load_library().b200_set_device(device))


def launch_count():
    return int(load_library().b200_launch_count())


FIELDS = {
    'status': ('u1', 1), 'track_id': ('<u4', 1), 'parent_id': ('<u4', 1), 'event_id': ('<u4', 1),
    'num_steps': ('<u4', 1), 'num_looping_steps': ('<u4', 1), 'time': ('<f8', 1),
    'step_length': ('<f8', 1), 'post_step_action': ('<u4', 1), 'along_step_action': ('<u4', 1),
    'particle_id': ('<u4', 1), 'energy': ('<f8', 1), 'material_id': ('<u4', 1),
    'interaction_mfp': ('<f8', 1), 'macro_xs': ('<f8', 1), 'energy_deposition': ('<f8', 1),
    'dedx_range': ('<f8', 1), 'rng': ('<u4', 6), 'pos': ('<f8', 3), 'dir': ('<f8', 3),
    'volume_id': ('<u4', 1), 'surface_id': ('<u4', 1), 'geo_level': ('<u4', 1),
    'sort_slots': ('<u4', 1),
}


HIT_FIELDS = {
    'detector': ('<u4', 1), 'track_id': ('<u4', 1), 'event_id': ('<u4', 1),
    'parent_id': ('<u4', 1), 'track_step_count': ('<u4', 1), 'particle': ('<u4', 1),
    'step_length': ('<f8', 1), 'energy_deposition': ('<f8', 1),
    'pre_time': ('<f8', 1), 'pre_energy': ('<f8', 1), 'pre_pos': ('<f8', 3), 'pre_dir': ('<f8', 3),
    'post_time': ('<f8', 1), 'post_energy': ('<f8', 1), 'post_pos': ('<f8', 3),
    'post_dir': ('<f8', 3),
}


def orange_build_image(org_json_path):
    """The geometry image of an .org.json file built by the library's own ORANGE
    construction (host only): bytes of a .b2img container."""
    L = load_library()
    ptr, size = C.c_void_p(), C.c_size_t()
    _check(L.b200_orange_build_image(os.fspath(org_json_path).encode(), C.byref(ptr),
                                     C.byref(size)))
    try:
        return C.string_at(ptr, size.value)
    finally:
        L.b200_string_free(ptr)


def import_root(root_path):
    """`celeritas::ImportData` of a reference physics export (.root) as a dict with the
    reference's member names, decoded by the library's own reader (host only)."""
    import json
    L = load_library()
    ptr = C.c_void_p()
    _check(L.b200_import_root(os.fspath(root_path).encode(), C.byref(ptr)))
    try:
        return json.loads(C.string_at(ptr))
    finally:
        L.b200_string_free(ptr)


class Params:
    """Problem parameters in HBM (reference: CoreParams)."""

    def __init__(self, image_path=None, image_bytes=None, org_json=None):
        L = load_library()
        h = C.c_void_p()
        if org_json is not None:
            # geometry built natively from the reference's ORANGE JSON input
            _check(L.b200_params_create_from_org_json(os.fspath(org_json).encode(), C.byref(h)))
        elif image_bytes is not None:
            # hand-off in memory (b200_params_create_from_memory)
            buf = bytes(image_bytes)
            _check(L.b200_params_create_from_memory(buf, C.c_size_t(len(buf)), C.byref(h)))
        else:
            _check(L.b200_params_create_from_image(os.fspath(image_path).encode(), C.byref(h)))
        self.h = h
        self.image_path = image_path

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                load_library().b200_params_destroy(self.h)
        except Exception:
            pass

    @property
    def action_labels(self):
        L = load_library()
        return [L.b200_params_action_label(self.h, i).decode()
                for i in range(L.b200_params_num_actions(self.h))]

    @property
    def volume_labels(self):
        L = load_library()
        return [L.b200_params_volume_label(self.h, i).decode()
                for i in range(L.b200_params_num_volumes(self.h))]

    @property
    def num_detectors(self):
        return load_library().b200_params_num_detectors(self.h)

    def find_particle(self, pdg):
        r = load_library().b200_params_find_particle(self.h, pdg)
        return None if r == 0xffffffff else r

    @property
    def num_particles(self):
        return load_library().b200_params_num_particles(self.h)

    @property
    def max_depth(self):
        return load_library().b200_params_max_depth(self.h)

    def generate_primaries(self, primary_options):
        """Primaries of a celer-sim `primary_options` dict; returns (primaries, offsets)."""
        import json
        L = load_library()
        text = json.dumps(primary_options).encode()
        count, per_event = C.c_uint64(), C.c_uint32()
        _check(L.b200_primaries_generate(self.h, text, None, 0, C.byref(count),
                                         C.byref(per_event)))
        out = np.zeros(count.value, dtype=PRIMARY_DTYPE)
        _check(L.b200_primaries_generate(self.h, text, out.ctypes.data, len(out),
                                         C.byref(count), C.byref(per_event)))
        offsets = np.arange(0, len(out) + 1, per_event.value, dtype=np.uint32)
        return out, offsets

    def trace(self, pos, direction, max_segments=64):
        """Ray-trace through the geometry on the GPU.

        Returns (volume[n, max_segments], surface[n, max_segments], distance[n, max_segments],
        count[n], safety[n]).
        """
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        direction = np.ascontiguousarray(direction, dtype=np.float64).reshape(-1, 3)
        n = len(pos)
        vol = np.full((n, max_segments), 0xffffffff, dtype=np.uint32)
        surf = np.full((n, max_segments), 0xffffffff, dtype=np.uint32)
        dist = np.zeros((n, max_segments))
        count = np.zeros(n, dtype=np.uint32)
        safety = np.zeros(n)
        _check(load_library().b200_geo_trace_host(
            self.h, pos.ctypes.data, direction.ctypes.data, n, max_segments, vol.ctypes.data,
            surf.ctypes.data, dist.ctypes.data, count.ctypes.data, safety.ctypes.data))
        return vol, surf, dist, count, safety


class Stepper:
    """One stream's stepping loop (reference: Stepper<MemSpace::device>)."""

    def __init__(self, params, num_track_slots, stream_id=0, action_times=False,
                 action_diagnostic=False, step_diagnostic_bins=0, fuse_threshold=0,
                 tail_threshold=0):
        L = load_library()
        self.params = params
        self.n = num_track_slots
        h = C.c_void_p()
        opts = StepperOptions(stream_id, num_track_slots, int(action_times),
                              int(action_diagnostic), step_diagnostic_bins, fuse_threshold,
                              tail_threshold)
        _check(L.b200_stepper_create_opts(params.h, C.byref(opts), C.byref(h)))
        self.h = h

    @property
    def all_action_labels(self):
        """Labels of every action (explicit and implicit) by action id."""
        L = load_library()
        return [L.b200_stepper_action_label(self.h, i).decode()
                for i in range(L.b200_stepper_num_actions(self.h))]

    def action_diagnostic(self):
        """counts[particle][action] of the post-step action of every track-step."""
        L = load_library()
        out = np.zeros((self.params.num_particles, L.b200_stepper_num_actions(self.h)),
                       dtype=np.uint32)
        _check(L.b200_stepper_action_diagnostic_get(self.h, out.ctypes.data))
        return out

    def step_diagnostic(self):
        """counts[particle][num_steps] of the steps every killed track took."""
        L = load_library()
        out = np.zeros((self.params.num_particles,
                        L.b200_stepper_step_diagnostic_bins(self.h) + 2), dtype=np.uint32)
        _check(L.b200_stepper_step_diagnostic_get(self.h, out.ctypes.data))
        return out

    def diagnostics_clear(self):
        _check(load_library().b200_stepper_diagnostics_clear(self.h))

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                load_library().b200_stepper_destroy(self.h)
        except Exception:
            pass

    def step(self, primaries=None):
        L = load_library()
        r = StepperResult()
        if primaries is not None and len(primaries):
            primaries = np.ascontiguousarray(primaries, dtype=PRIMARY_DTYPE)
            _check(L.b200_stepper_step(self.h, primaries.ctypes.data, len(primaries), C.byref(r)))
        else:
            _check(L.b200_stepper_step(self.h, None, 0, C.byref(r)))
        return dict(generated=r.generated, queued=r.queued, active=r.active, alive=r.alive)

    def advance(self, max_iterations):
        """Up to max_iterations iterations without primaries (device-resident loop while few
        tracks are left); returns the list of per-iteration counts."""
        L = load_library()
        res = (StepperResult * max_iterations)()
        n = C.c_uint32()
        _check(L.b200_stepper_advance(self.h, max_iterations, res, C.byref(n)))
        return [dict(generated=r.generated, queued=r.queued, active=r.active, alive=r.alive)
                for r in res[:n.value]]

    @property
    def tail_iterations(self):
        return int(load_library().b200_stepper_tail_iterations(self.h))

    def warm_up(self):
        _check(load_library().b200_stepper_warm_up(self.h))

    def reseed(self, event_id):
        _check(load_library().b200_stepper_reseed(self.h, event_id))

    def kill_active(self):
        _check(load_library().b200_stepper_kill_active(self.h))

    @property
    def step_action_labels(self):
        L = load_library()
        return [L.b200_stepper_step_action_label(self.h, i).decode()
                for i in range(L.b200_stepper_num_step_actions(self.h))]

    def set_action_times(self, enable=True):
        _check(load_library().b200_stepper_set_action_times(self.h, int(enable)))

    @property
    def action_times(self):
        """Accumulated device seconds per step action (label -> seconds)."""
        L = load_library()
        return {lab: L.b200_stepper_action_time(self.h, i)
                for i, lab in enumerate(self.step_action_labels)}

    @property
    def launch_count(self):
        return int(load_library().b200_stepper_launch_count(self.h))

    def get(self, field):
        L = load_library()
        dt, w = FIELDS[field]
        out = np.zeros((self.n, w) if w > 1 else self.n, dtype=dt)
        _check(L.b200_state_get(L.b200_stepper_state(self.h), field.encode(), out.ctypes.data))
        return out

    def hits(self):
        """Step/hit output of the last step iteration (reference: DetectorStepOutput) as a
        dict of arrays, one entry per hit, in track-slot order."""
        L = load_library()
        n = C.c_uint32()
        _check(L.b200_stepper_hits_count(self.h, C.byref(n)))
        out = {}
        for name, (dt, w) in HIT_FIELDS.items():
            a = np.zeros((n.value, w) if w > 1 else n.value, dtype=dt)
            if n.value:
                _check(L.b200_stepper_hits_get(self.h, name.encode(), a.ctypes.data))
            out[name] = a
        return out

    def sort_offsets(self):
        """First index of every sort key in get('sort_slots') after the last sort action
        (TrackOrder::reindex_*); entry [number of keys + 1] is the number of slots."""
        L = load_library()
        n = max(L.b200_params_num_actions(self.params.h),
                L.b200_params_num_particles(self.params.h), 1) + 2
        out = np.zeros(n, dtype=np.uint32)
        _check(L.b200_state_get(L.b200_stepper_state(self.h), b'sort_offsets', out.ctypes.data))
        return out

    def interaction_lists(self):
        """Per-model slot lists of the last per-action step: {action id: sorted slots}."""
        L = load_library()
        nm = L.b200_params_num_models(self.params.h)
        first = L.b200_params_model_action_begin(self.params.h)
        counts = np.zeros(16, dtype=np.uint32)
        lists = np.zeros((nm, self.n), dtype=np.uint32)
        state = L.b200_stepper_state(self.h)
        _check(L.b200_state_get(state, b'interact_count', counts.ctypes.data))
        _check(L.b200_state_get(state, b'interact_list', lists.ctypes.data))
        return {first + m: np.sort(lists[m, :counts[m]]) for m in range(nm)}

    def dense_lists(self, counts):
        """(charged slots, neutral slots) of the dense active lists, given the sizes."""
        L = load_library()
        slots = np.zeros(self.n, dtype=np.uint32)
        _check(L.b200_state_get(L.b200_stepper_state(self.h), b'track_slots',
                                slots.ctypes.data))
        nc, nn = counts
        return slots[:nc], slots[self.n - nn:][::-1]

    def calo(self):
        L = load_library()
        out = np.zeros(self.params.num_detectors)
        _check(L.b200_state_calo_get(L.b200_stepper_state(self.h), out.ctypes.data))
        return out

    def calo_clear(self):
        L = load_library()
        _check(L.b200_state_calo_clear(L.b200_stepper_state(self.h)))

    def run_events(self, primaries, offsets, merge_events=False, max_steps=0):
        """Transport whole events from HOST buffers (celer-sim Transporter loop)."""
        L = load_library()
        primaries = np.ascontiguousarray(primaries, dtype=PRIMARY_DTYPE)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        r = RunResult()
        _check(L.b200_run_events(self.h, primaries.ctypes.data, offsets.ctypes.data,
                                 len(offsets) - 1, int(merge_events), max_steps, C.byref(r)))
        return dict(num_steps=int(r.num_steps), num_step_iterations=int(r.num_step_iterations),
                    num_primaries=int(r.num_primaries), max_queued=int(r.max_queued),
                    seconds=r.seconds, num_tracks=int(r.num_tracks),
                    num_aborted=int(r.num_aborted))


def run_events_streams(steppers, primaries, offsets, merge_events=False, max_steps=0):
    """Transport events over several steppers (streams) concurrently; event e goes to
    steppers[e % len(steppers)]. Returns (per-stream results, device seconds)."""
    L = load_library()
    primaries = np.ascontiguousarray(primaries, dtype=PRIMARY_DTYPE)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
    n = len(steppers)
    handles = (C.c_void_p * n)(*[s.h for s in steppers])
    results = (RunResult * n)()
    seconds = C.c_double()
    _check(L.b200_run_events_streams(handles, n, primaries.ctypes.data, offsets.ctypes.data,
                                     len(offsets) - 1, int(merge_events), max_steps, results,
                                     C.byref(seconds)))
    out = [dict(num_steps=int(r.num_steps), num_step_iterations=int(r.num_step_iterations),
                num_primaries=int(r.num_primaries), max_queued=int(r.max_queued),
                seconds=r.seconds, num_tracks=int(r.num_tracks),
                num_aborted=int(r.num_aborted)) for r in results]
    return out, seconds.value


def celer_sim_run(run_input):
    """Run a celer-sim input (dict or JSON text) and return the report as a dict."""
    import json
    L = load_library()
    text = run_input if isinstance(run_input, str) else json.dumps(run_input)
    report = C.c_void_p()
    _check(L.b200_celer_sim_run(text.encode(), C.byref(report)))
    try:
        return json.loads(C.string_at(report).decode())
    finally:
        L.b200_string_free(report)


def make_primaries(n, particle_id=0, energy=100.0, pos=(0, 0, 0), direction=(1, 0, 0),
                   event_of=lambda i: 0):
    p = np.zeros(n, dtype=PRIMARY_DTYPE)
    p['particle_id'] = particle_id
    p['energy'] = energy
    p['pos'] = pos
    p['dir'] = direction
    p['event_id'] = [event_of(i) for i in range(n)]
    return p
