"""Generate the golden fixtures of this directory from the reference itself (oracle/_ref,
the reference's host sources compiled by oracle/Makefile). Run where /root/reference exists:

    python tests/golden/make_golden.py

Each fixture records, for one problem and one set of primaries, the reference's result of
EVERY step iteration: the StepperResult counters, a CRC of every integer state field and of
the XORWOW words over all slots, and the final per-detector energy deposition. The GPU tests
(tests/test_gpu_golden.py) replay the same input and must reproduce the counters and CRCs
exactly and the tallies to 1e-9: the same check as the lock-step tests, without needing the
reference at run time.
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'oracle'))
sys.path.insert(0, os.path.join(REPO, 'tests'))

INT_FIELDS = ['status', 'post_step_action', 'along_step_action', 'volume_id', 'surface_id']
ACTIVE_INT_FIELDS = ['particle_id', 'material_id', 'num_steps', 'track_id', 'parent_id',
                     'event_id']

CASES = {
    'simple-compton': dict(image='simple-compton', slots=64, particle=22, nprim=32, energy=100.0),
    'testem3-small': dict(image='testem3-small', slots=2048, particle=11, nprim=2, energy=1000.0),
    'testem3-small-initcharge': dict(image='testem3-small-initcharge', slots=2048, particle=11,
                                     nprim=3, energy=500.0),
    'simple-cms-em-field': dict(image='simple-cms-em-field', slots=4096, particle=11, nprim=2,
                                energy=300.0, pos=(0, 0, 0), direction=(0.6, 0.0, 0.8)),
    # four-level CMS-scale stand-in (tools/make_cms_scale.py), 1 T: through the tracker
    # shells into the ECAL rect array
    'cms-scale-small': dict(image='cms-scale-small', slots=4096, particle=11, nprim=2,
                            energy=500.0, pos=(0, 0, 0), direction=(0.6, 0.48, 0.64)),
}


def state_crcs(stepper):
    """CRC32 of the integer fields (inactive slots masked where their content is
    unspecified) and of the RNG words."""
    status = stepper.get('status')
    active = status != 0
    out = {}
    for f in INT_FIELDS:
        out[f] = zlib.crc32(np.ascontiguousarray(stepper.get(f)).tobytes())
    for f in ACTIVE_INT_FIELDS:
        a = np.where(active, stepper.get(f), 0)
        out[f] = zlib.crc32(np.ascontiguousarray(a).tobytes())
    out['rng'] = zlib.crc32(np.ascontiguousarray(stepper.get('rng')).tobytes())
    return out


def primaries_for(case, find_particle, make_primaries):
    return make_primaries(case['nprim'], particle_id=find_particle(case['particle']),
                          energy=case['energy'], pos=case.get('pos', (-22, 0, 0)),
                          direction=case.get('direction', (1, 0, 0)))


def run(stepper, prim, record):
    c = stepper.step(prim)
    it = 0
    while True:
        record(it, c)
        if not (c['alive'] or c['queued']):
            break
        c = stepper.step()
        it += 1


def main():
    import celerref
    for name, case in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        cfg = json.load(open(os.path.join(REPO, 'data', 'images', case['image'] + '.json')))
        problem = celerref.Problem(cfg)
        stepper = problem.stepper(case['slots'])
        # particle ids as the reference's ParticleParams numbers them (asked through its
        # own PrimaryGenerator)
        pdg_to_id = {}
        for pdg in (11, 22, -11):
            try:
                one = problem.generate_primaries(
                    {'seed': 0, 'pdg': [pdg], 'num_events': 1, 'primaries_per_event': 1,
                     'energy': 1.0, 'position': [0, 0, 0], 'direction': [1, 0, 0]})
                pdg_to_id[pdg] = int(one[0]['particle_id'])
            except RuntimeError:
                pass
        prim = primaries_for(case, lambda pdg: pdg_to_id[pdg], celerref.make_primaries)
        steps = []

        def record(it, c):
            entry = dict(c)
            if it % 5 == 0:
                entry['crc'] = state_crcs(stepper)
            steps.append(entry)

        run(stepper, prim, record)
        steps[-1]['crc'] = state_crcs(stepper)
        out = {'case': case, 'particle_ids': {str(k): v for k, v in pdg_to_id.items()},
               'steps': steps}
        ndet = len(cfg.get('simple_calo', []))
        if ndet:
            out['calo'] = problem.calo(ndet).tolist()
        path = os.path.join(HERE, name + '.json')
        json.dump(out, open(path, 'w'), separators=(',', ':'))
        print(name, len(steps), 'iterations', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
