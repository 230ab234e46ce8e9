"""celeritas_b200: B200-native per-step track loop behind the reference's Stepper surface.

This package is a thin ctypes binding over ``libceleritas_b200.so`` (C-ABI in
``include/celeritas_b200.h``). There is no Python or CPU implementation of any
step action: importing works anywhere, but creating params/state requires a GPU
and fails loudly otherwise.
"""
from .lib import (Params, Stepper, Primary, PRIMARY_DTYPE, make_primaries, library_path,
                  load_library, launch_count, device_count, set_device, B200Error,
                  celer_sim_run, run_events_streams, orange_build_image, import_root)

__all__ = ['Params', 'Stepper', 'Primary', 'PRIMARY_DTYPE', 'make_primaries', 'library_path',
           'load_library', 'launch_count', 'device_count', 'set_device', 'B200Error',
           'celer_sim_run', 'run_events_streams', 'orange_build_image', 'import_root']
