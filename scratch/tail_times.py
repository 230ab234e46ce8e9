"""Per-iteration times of one celer-sim run with the device-resident loop on / off."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import celeritas_b200 as cb
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
workload = sys.argv[1] if len(sys.argv) > 1 else 'testem3'
img = {'testem3': 'testem3-initcharge', 'cms-scale': 'cms-scale'}[workload]
cfg = json.load(open(os.path.join(REPO, 'data', 'images', img + '.json')))
if workload == 'testem3':
    opts = {'seed': 1, 'pdg': [11], 'num_events': 20, 'primaries_per_event': 100, 'energy': 1000.0,
            'position': [-22, 0, 0], 'direction': [1, 0, 0]}
else:
    opts = {'seed': 20220904, 'pdg': [11, 22], 'num_events': 100, 'primaries_per_event': 10,
            'energy': 10000.0, 'position': [0, 0, 0], 'direction': {'distribution': 'isotropic'}}
run_input = {'_format': 'celer-sim', 'use_device': True, 'image_file': 'data/images/%s.b2img' % img,
             'base_dir': REPO, 'geometry_file': cfg['geometry_file'], 'primary_options': opts,
             'seed': cfg['seed'], 'num_track_slots': 1 << 19, 'initializer_capacity': 1 << 22,
             'secondary_stack_factor': 3, 'simple_calo': cfg['simple_calo'], 'merge_events': True,
             'write_step_times': True, 'write_track_counts': True, 'warm_up': True,
             'track_order': 'init_charge'}
if 'field' in cfg:
    run_input['field'] = cfg['field']
for thr in ('4294967295', sys.argv[2] if len(sys.argv) > 2 else '16384'):
    os.environ['B200_TAIL_THRESHOLD'] = thr
    for rep in range(2):
        out = cb.celer_sim_run(run_input)
    r = out['result']['runner']
    t = np.array(r['time']['steps'][0]); a = np.array(r['active'][0])
    print('threshold', thr, 'iterations', len(a), 'total ms %.1f' % (t.sum() * 1e3))
    edges = [0, 16, 128, 1024, 4096, 16384, 65536, 1 << 21]
    for lo, hi in zip(edges[:-1], edges[1:]):
        m = (a >= lo) & (a < hi)
        if m.any():
            print('  active [%6d,%7d): iters %4d  time %7.2f ms  us/iter %6.1f (median %6.1f)' % (
                lo, hi, m.sum(), t[m].sum() * 1e3, t[m].mean() * 1e6, np.median(t[m]) * 1e6))
