"""Export the problem images used by tests and bench (runs where /root/reference was built
into oracle/_ref; the images are committed so the GPU box needs neither).

Each image is the reference's own host-side CoreParams, flattened by the adapter in
oracle/ref_harness/ExportImage.cc.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(REPO, 'oracle'))
import celerref  # noqa: E402

GAPS = ['gap_%d' % i for i in range(50)]
ABSORBERS = ['absorber_%d' % i for i in range(50)]

PROBLEMS = {
    'simple-compton': {'problem': 'simple-compton',
                       'geometry_file': 'data/geometry/two-boxes.org.json', 'seed': 20220511},
    # BASELINE config 0: TestEm3, no field, MSC off, no fluctuations
    'testem3-nomsc': {'geometry_file': 'data/geometry/testem3-flat.org.json',
                      'physics_file': 'data/physics/testem3-steel-lar.json',
                      'seed': 20220904, 'initializer_capacity': 1 << 20, 'max_events': 1024,
                      'disable_msc': True, 'eloss_fluctuation': False,
                      'simple_calo': GAPS + ABSORBERS},
    # BASELINE config 1: TestEm3 full EM (Urban MSC + fluctuations)
    'testem3': {'geometry_file': 'data/geometry/testem3-flat.org.json',
                'physics_file': 'data/physics/testem3-steel-lar.json',
                'seed': 20220904, 'initializer_capacity': 1 << 25, 'max_events': 16384,
                'simple_calo': GAPS + ABSORBERS},
    # BASELINE config 2: simple-CMS nested cylinders in a 1 T uniform field (the reference's
    # bundled simple-cms export: Compton + e-/e+ ionisation + Urban MSC, no fluctuations)
    'simple-cms-field': {'geometry_file': 'data/geometry/simple-cms.org.json',
                         'physics_file': 'data/physics/simple-cms.json',
                         'seed': 20220904, 'initializer_capacity': 1 << 20, 'max_events': 1024,
                         'field': [0, 0, 1],
                         'simple_calo': ['si_tracker', 'em_calorimeter', 'had_calorimeter',
                                         'sc_solenoid', 'fe_muon_chambers']},
    'simple-cms': {'geometry_file': 'data/geometry/simple-cms.org.json',
                   'physics_file': 'data/physics/simple-cms.json',
                   'seed': 20220904, 'initializer_capacity': 1 << 20, 'max_events': 1024},
    # same geometry and field with the full-EM stand-in materials (10 GeV showers)
    'simple-cms-em-field': {'geometry_file': 'data/geometry/simple-cms.org.json',
                            'physics_file': 'data/physics/simple-cms-steel-lar.json',
                            'seed': 20220904, 'initializer_capacity': 1 << 25,
                            'max_events': 16384, 'field': [0, 0, 1],
                            'simple_calo': ['si_tracker', 'em_calorimeter', 'had_calorimeter',
                                            'sc_solenoid', 'fe_muon_chambers']},
    # two-level geometry (daughter universes with translations), same physics
    'testem3-nested': {'geometry_file': 'data/geometry/testem3.org.json',
                       'physics_file': 'data/physics/testem3-nested-steel-lar.json',
                       'seed': 20220904, 'initializer_capacity': 1 << 18, 'max_events': 64,
                       'simple_calo': ['gap', 'absorber']},
    # the reference's GPU default track order: new tracks partitioned by charge
    'testem3-small-initcharge': {'geometry_file': 'data/geometry/testem3-flat.org.json',
                                 'physics_file': 'data/physics/testem3-steel-lar.json',
                                 'seed': 20220904, 'initializer_capacity': 1 << 18,
                                 'max_events': 64, 'track_order': 'init_charge',
                                 'simple_calo': GAPS + ABSORBERS},
    'testem3-initcharge': {'geometry_file': 'data/geometry/testem3-flat.org.json',
                           'physics_file': 'data/physics/testem3-steel-lar.json',
                           'seed': 20220904, 'initializer_capacity': 1 << 25,
                           'max_events': 16384, 'track_order': 'init_charge',
                           'simple_calo': GAPS + ABSORBERS},
    # small-capacity variant of the same physics for lock-step tests
    'testem3-small': {'geometry_file': 'data/geometry/testem3-flat.org.json',
                      'physics_file': 'data/physics/testem3-steel-lar.json',
                      'seed': 20220904, 'initializer_capacity': 1 << 18, 'max_events': 64,
                      'simple_calo': GAPS + ABSORBERS},
    # celer-sim `brem_combined`: one bremsstrahlung model (Seltzer-Berger below 1 GeV,
    # relativistic above) instead of two. The reference allows it only with single-element
    # materials (PhysicsParams.cc:665): liquid-argon sphere in vacuum, full EM
    'lar-sphere-combined': {'geometry_file': 'data/geometry/lar-sphere.org.json',
                            'physics_file': 'data/physics/lar-sphere-em.json',
                            'seed': 20220904, 'initializer_capacity': 1 << 18,
                            'max_events': 64, 'brem_combined': True,
                            'simple_calo': ['sphere']},
}

# BASELINE configs 3/4: CMS-scale stand-in (tools/make_cms_scale.py): four levels, 2916 unit
# volumes, two rect arrays, BIH trees over 276 / 2304 volumes, 1 T uniform field
CMS_CALO = (['solenoid'] + ['ecal%s_%s_%d' % (a, k, i) for a in 'xy' for k in ('abs', 'gap')
                            for i in range(5)]
            + ['endcap_%s_%d' % (k, i) for k in ('abs', 'gap') for i in range(20)])
PROBLEMS['cms-scale'] = {'geometry_file': 'data/geometry/cms-scale.org.json',
                         'physics_file': 'data/physics/cms-scale-steel-lar.json',
                         'seed': 20220904, 'initializer_capacity': 1 << 25, 'max_events': 16384,
                         'field': [0, 0, 1], 'track_order': 'init_charge',
                         'simple_calo': CMS_CALO}
PROBLEMS['simple-cms-em-field-initcharge'] = dict(PROBLEMS['simple-cms-em-field'],
                                                     track_order='init_charge')
PROBLEMS['cms-scale-small'] = dict(PROBLEMS['cms-scale'], initializer_capacity=1 << 18,
                                   max_events=64, track_order='none')

# Volumes beyond the register path's 32 faces / intersections (tools/make_many_faces.py): a
# 113-face background volume filled with liquid argon, a 40-plane polyhedron, a box with 30
# holes as one internal-surface volume; full EM
PROBLEMS['many-faces'] = {'geometry_file': 'data/geometry/many-faces.org.json',
                          'physics_file': 'data/physics/many-faces-steel-lar.json',
                          'seed': 20220904, 'initializer_capacity': 1 << 18, 'max_events': 64,
                          'simple_calo': ['mother', 'polyhedron', 'cheese']}

# Rayleigh scattering on top of the full-EM list: the reference's
# four-steel-slabs export with its LivermoreRayleigh cross sections kept
PROBLEMS['four-steel-slabs-rayleigh'] = {
    'geometry_file': 'data/geometry/four-steel-slabs.org.json',
    'physics_file': 'data/physics/four-steel-slabs-em-rayleigh.json',
    'seed': 20220904, 'initializer_capacity': 1 << 18, 'max_events': 64,
    'simple_calo': ['box@1', 'box@2', 'box@3', 'box@4']}

# ... plus single Coulomb scattering (Wentzel model) for e-/e+ above 100 MeV
PROBLEMS['four-steel-slabs-coulomb'] = dict(
    PROBLEMS['four-steel-slabs-rayleigh'],
    physics_file='data/physics/four-steel-slabs-em-coulomb.json')

# ... and the export's muons: mu-/mu+ ionisation (four models) and muon bremsstrahlung
PROBLEMS['four-steel-slabs-muon'] = dict(
    PROBLEMS['four-steel-slabs-rayleigh'],
    physics_file='data/physics/four-steel-slabs-em-muon.json')

# Geometry-only images for the navigation (ray-trace) parity tests: the ORANGE test
# geometries of the reference (test/orange/data, test/geocel/data)
for _g in ('two-boxes', 'testem3-flat', 'testem3', 'simple-cms', 'five-volumes', 'universes',
           'rect-array', 'nested-rect-arrays', 'hex-array', 'three-spheres', 'testem15',
           'lar-sphere', 'four-steel-slabs', 'one-steel-sphere', 'cms-scale', 'many-faces'):
    PROBLEMS['geo-' + _g] = {'problem': 'geometry',
                             'geometry_file': 'data/geometry/%s.org.json' % _g}


def main(names):
    out = os.path.join(REPO, 'data', 'images')
    os.makedirs(out, exist_ok=True)
    for name in names or PROBLEMS:
        cfg = PROBLEMS[name]
        p = celerref.Problem(cfg)
        path = os.path.join(out, name + '.b2img')
        p.export_image(path)
        with open(os.path.join(out, name + '.json'), 'w') as f:
            json.dump(cfg, f, indent=1)
        print(name, os.path.getsize(path))


if __name__ == '__main__':
    main(sys.argv[1:])
