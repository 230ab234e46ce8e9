"""Compose the physics inputs for the benchmark problems from the reference's bundled exports.

The reference ships *no* physics export for TestEm3 (Pb + lAr) or CMS; its TestEm3 tests
build tables live from Geant4 (test/celeritas/TestEm3Base.hh:20-38), which is not available
here. What it does ship (test/celeritas/data/*.root, decoded by tools/rootlite.py):

  four-steel-slabs.root  stainless steel + vacuum, full EM, 7 bins/decade, 1e-4..1e8 MeV
  lar-sphere.root        liquid argon + vacuum,   full EM, 7 bins/decade, 1e-4..1e8 MeV
  br29 / pe-*-19.dat     Seltzer-Berger table for Z=29, Livermore photoelectric data for Z=19

SUBSTITUTIONS made here (identical inputs are given to the reference and to the CUDA code, so
parity is unaffected; absolute physics is *not* that of a Pb/lAr calorimeter):

  1. The TestEm3 absorber material is stainless steel (from four-steel-slabs) instead of Pb;
     the gap is liquid argon (from lar-sphere); the world is G4_Galactic.
  2. Every element uses the Z=29 Seltzer-Berger differential cross-section table and the
     Z=19 Livermore photoelectric subshell data (the only ones bundled).
  3. Processes kept: e-/e+ ionisation, bremsstrahlung, e+ annihilation, Compton,
     photoelectric, gamma conversion, plus Urban MSC. Coulomb and muon processes are
     dropped (outside the hot-path scope); Rayleigh is kept only
     in four-steel-slabs-em-rayleigh.json.

Outputs: data/physics/testem3-steel-lar.json (and lar/steel single-material variants used by
the unit tests).
"""
import copy
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
PHYS = os.path.join(REPO, 'data', 'physics')

KEEP_PDG = (-11, 11, 22)
# ImportProcessClass: e_ioni=7, e_brems=8, photoelectric=9, compton=10, conversion=11,
# annihilation=13
KEEP_PROCESS = (7, 8, 9, 10, 11, 13)


def load(name):
    return json.load(open(os.path.join(PHYS, name + '.json')))


def material_index(data, name):
    for i, m in enumerate(data['geo_materials']):
        if m['name'] == name:
            return i
    raise KeyError(name)


def phys_index_of_geo(data, geo_idx):
    for i, m in enumerate(data['phys_materials']):
        if m['geo_material_id'] == geo_idx:
            return i
    raise KeyError(geo_idx)


def add_element_data(data):
    ed = load('element-data-z29sb-z19pe')
    zs = sorted({e['atomic_number'] for e in data['elements']})
    data['sb_data'] = {str(z): ed['sb'] for z in zs}
    data['livermore_pe_data'] = {str(z): ed['livermore_pe'] for z in zs}


def extend_urban_msc(data):
    """Give Urban MSC one table over the full energy range.

    four-steel-slabs was exported with Urban MSC below 100 MeV and WentzelVI above; lar-sphere
    with Urban over the whole range. UrbanMscParams requires identical limits for all
    materials, and WentzelVI is out of scope, so the steel Urban table (model class 3) is
    extended above 100 MeV with the WentzelVI (class 5) transport cross sections of the same
    export (same grid spacing; the duplicated 100 MeV node keeps the Urban value).
    """
    by_key = {(m['particle_pdg'], m['model_class']): m for m in data['msc_models']}
    for (pdg, cls), m in list(by_key.items()):
        if cls != 3 or (pdg, 5) not in by_key:
            continue
        hi = by_key[(pdg, 5)]
        for v, vh in zip(m['xs_table']['physics_vectors'], hi['xs_table']['physics_vectors']):
            if v['x'][-1] == vh['x'][0]:
                v['x'] = v['x'] + vh['x'][1:]
                v['y'] = v['y'] + vh['y'][1:]
    data['msc_models'] = [m for m in data['msc_models'] if m['model_class'] == 3]


def extend_coulomb_down(data, emin=1e-4):
    """Give the Coulomb-scattering process a defined cross section below its 100 MeV limit.

    Geant4 exports eCoulombScattering from 100 MeV up (it is meant to go with Wentzel VI
    multiple scattering above that energy). The reference extrapolates a cross-section grid
    flat below its first node (grid/XsCalculator.hh:119-122) and then finds no applicable
    model for the selected process (CoulombScatteringModel.cc:44-66 notes the limitation), so
    the export as-is cannot be run with electrons below 100 MeV. Here the grids are continued
    down to `emin` on the same logarithmic spacing with ZERO macroscopic cross section (what
    Geant4 itself does below a model's limit); the per-element grids repeat their first
    value so that element selection stays defined on the sliver below 100 MeV where the
    linear interpolation towards the first exported node is nonzero.
    """
    import math
    for p in data['processes']:
        if p['process_class'] != 6:
            continue
        for t in p['tables']:
            for v in t['physics_vectors']:
                per_decade = round((len(v['x']) - 1) / math.log10(v['x'][-1] / v['x'][0]))
                n = round(math.log10(v['x'][0] / emin) * per_decade)
                low = [v['x'][0] * 10 ** (-(n - k) / per_decade) for k in range(n)]
                v['x'] = low + v['x']
                v['y'] = [0.0] * n + v['y']
        for m in p['models']:
            for mm in m['materials']:
                n = round(math.log10(mm['energy'][0] / emin))
                low = [mm['energy'][0] * 10 ** (-(n - k)) for k in range(n)]
                mm['energy'] = low + mm['energy']
                mm['micro_xs'] = [[xs[0]] * n + xs for xs in mm['micro_xs']]


def filter_physics(data, keep_process=KEEP_PROCESS, keep_pdg=KEEP_PDG):
    extend_urban_msc(data)
    data['particles'] = [p for p in data['particles'] if p['pdg'] in keep_pdg]
    data['processes'] = [p for p in data['processes']
                         if p['particle_pdg'] in keep_pdg and p['process_class'] in keep_process]
    data['msc_models'] = [m for m in data['msc_models'] if m['particle_pdg'] in keep_pdg]
    for pm in data['phys_materials']:
        pm['pdg_cutoffs'] = [c for c in pm['pdg_cutoffs'] if c['first'] in keep_pdg]
    tp = data['trans_params']
    tp['looping'] = {k: v for k, v in tp['looping'].items() if int(k) in keep_pdg}
    data['mu_pair_production_data'] = {'atomic_number': [], 'physics_vectors': []}


def merge(sources, volumes):
    """sources: list of (data, geo material name); merged material i = sources[i].
    volumes: list of (volume name, merged material index)."""
    out = copy.deepcopy(sources[0][0])
    out['isotopes'], out['elements'], out['geo_materials'], out['phys_materials'] = [], [], [], []
    elem_map = {}  # (source idx, old element id) -> new id
    for si, (data, matname) in enumerate(sources):
        gi = material_index(data, matname)
        gm = copy.deepcopy(data['geo_materials'][gi])
        for comp in gm['elements']:
            key = (si, comp['element_id'])
            el = data['elements'][comp['element_id']]
            # merge identical elements by name
            found = [i for i, e in enumerate(out['elements']) if e['name'] == el['name']]
            if found:
                elem_map[key] = found[0]
            else:
                el = copy.deepcopy(el)
                new_iso = []
                for f in el['isotopes_fractions']:
                    out['isotopes'].append(data['isotopes'][f['first']])
                    new_iso.append({'first': len(out['isotopes']) - 1, 'second': f['second']})
                el['isotopes_fractions'] = new_iso
                out['elements'].append(el)
                elem_map[key] = len(out['elements']) - 1
            comp['element_id'] = elem_map[key]
        out['geo_materials'].append(gm)
        pm = copy.deepcopy(data['phys_materials'][phys_index_of_geo(data, gi)])
        pm['geo_material_id'] = si
        out['phys_materials'].append(pm)

    def pick(data, si, lst):
        matname = sources[si][1]
        return lst[phys_index_of_geo(data, material_index(data, matname))]

    # Processes: must exist in every source
    procs = []
    for p0 in sources[0][0]['processes']:
        key = (p0['particle_pdg'], p0['process_class'])
        matches = []
        for data, _ in sources:
            m = [p for p in data['processes']
                 if (p['particle_pdg'], p['process_class']) == key]
            matches.append(m[0] if m else None)
        if any(m is None for m in matches):
            continue
        p = copy.deepcopy(p0)
        assert all([mm['model_class'] for mm in m['models']]
                   == [mm['model_class'] for mm in p0['models']] for m in matches), key
        for mi, model in enumerate(p['models']):
            model['materials'] = [copy.deepcopy(pick(sources[si][0], si,
                                                     matches[si]['models'][mi]['materials']))
                                  for si in range(len(sources))]
        assert all([t['table_type'] for t in m['tables']]
                   == [t['table_type'] for t in p0['tables']] for m in matches), key
        for ti, table in enumerate(p['tables']):
            table['physics_vectors'] = [copy.deepcopy(pick(sources[si][0], si,
                                                           matches[si]['tables'][ti]['physics_vectors']))
                                        for si in range(len(sources))]
        procs.append(p)
    out['processes'] = procs
    msc = []
    for m0 in sources[0][0]['msc_models']:
        key = (m0['particle_pdg'], m0['model_class'])
        matches = []
        for data, _ in sources:
            m = [x for x in data['msc_models'] if (x['particle_pdg'], x['model_class']) == key]
            matches.append(m[0] if m else None)
        if any(m is None for m in matches):
            continue
        mm = copy.deepcopy(m0)
        mm['xs_table']['physics_vectors'] = [
            copy.deepcopy(pick(sources[si][0], si, matches[si]['xs_table']['physics_vectors']))
            for si in range(len(sources))]
        msc.append(mm)
    out['msc_models'] = msc
    out['regions'] = [{'name': 'DefaultRegionForTheWorld', 'field_manager': False,
                       'production_cuts': True, 'user_limits': False}]
    out['volumes'] = [{'geo_material_id': mi, 'region_id': 0, 'phys_material_id': mi,
                       'name': name, 'solid_name': name} for name, mi in volumes]
    return out


def main():
    steel = load('four-steel-slabs')
    lar = load('lar-sphere')
    for d in (steel, lar):
        filter_physics(d)

    # TestEm3 stand-in: world = G4_Galactic, absorber = steel, gap = lAr
    vols = [('world', 0)]
    for i in range(50):
        vols.append(('gap_%d' % i, 2))
        vols.append(('absorber_%d' % i, 1))
    em3 = merge([(steel, 'G4_Galactic'), (steel, 'G4_STAINLESS-STEEL'), (lar, 'lAr')], vols)
    add_element_data(em3)
    json.dump(em3, open(os.path.join(PHYS, 'testem3-steel-lar.json'), 'w'), separators=(',', ':'))
    print('testem3-steel-lar:', [m['name'] for m in em3['geo_materials']],
          [e['name'] for e in em3['elements']],
          [(p['particle_pdg'], p['process_class']) for p in em3['processes']])

    # Two-level TestEm3 (data/geometry/testem3.org.json: 50 'pair' daughters of one
    # gap+absorber universe): same materials, volumes named as in that geometry
    nested = merge([(steel, 'G4_Galactic'), (steel, 'G4_STAINLESS-STEEL'), (lar, 'lAr')],
                   [('world', 0), ('gap', 2), ('absorber', 1)])
    add_element_data(nested)
    json.dump(nested, open(os.path.join(PHYS, 'testem3-nested-steel-lar.json'), 'w'),
              separators=(',', ':'))

    # simple-CMS geometry (data/geometry/simple-cms.org.json) with full-EM stand-in materials:
    # the bundled simple-cms.root has only Compton + ionisation + MSC below 1 GeV, which
    # cannot stop a 10 GeV shower. Tracker -> lAr, calorimeters/solenoid/muon iron -> steel.
    cms_vols = [('vacuum_tube', 0), ('si_tracker', 2), ('em_calorimeter', 1),
                ('had_calorimeter', 1), ('sc_solenoid', 1), ('fe_muon_chambers', 1),
                ('world', 0)]
    cms = merge([(steel, 'G4_Galactic'), (steel, 'G4_STAINLESS-STEEL'), (lar, 'lAr')], cms_vols)
    add_element_data(cms)
    json.dump(cms, open(os.path.join(PHYS, 'simple-cms-steel-lar.json'), 'w'),
              separators=(',', ':'))

    # lar-sphere as bundled (geometry data/geometry/lar-sphere.org.json): volumes keep names
    lar_full = load('lar-sphere')
    filter_physics(lar_full)
    add_element_data(lar_full)
    json.dump(lar_full, open(os.path.join(PHYS, 'lar-sphere-em.json'), 'w'), separators=(',', ':'))
    steel_full = load('four-steel-slabs')
    filter_physics(steel_full)
    add_element_data(steel_full)
    json.dump(steel_full, open(os.path.join(PHYS, 'four-steel-slabs-em.json'), 'w'),
              separators=(',', ':'))

    # the same with Rayleigh scattering kept (ImportProcessClass::rayleigh = 12,
    # LivermoreRayleigh cross sections as exported): SURVEY 8(f)4
    steel_ray = load('four-steel-slabs')
    filter_physics(steel_ray, KEEP_PROCESS + (12,))
    # volumes named as in data/geometry/four-steel-slabs.org.json ('box@1'..'box@4' match
    # the extension-free 'box', geo/GeoMaterialParams.cc:108-140)
    steel_ray = merge([(steel_ray, 'G4_Galactic'), (steel_ray, 'G4_STAINLESS-STEEL')],
                      [('box', 1), ('World', 0)])
    add_element_data(steel_ray)
    json.dump(steel_ray, open(os.path.join(PHYS, 'four-steel-slabs-em-rayleigh.json'), 'w'),
              separators=(',', ':'))

    # ... and with single Coulomb scattering kept (ImportProcessClass::coulomb_scat = 6,
    # eCoulombScattering tables above 100 MeV as exported; without the Wentzel VI
    # multiple-scattering model the reference samples every angle, WentzelOKVIParams.cc:40-56)
    steel_cs = load('four-steel-slabs')
    filter_physics(steel_cs, KEEP_PROCESS + (6, 12))
    extend_coulomb_down(steel_cs)
    steel_cs = merge([(steel_cs, 'G4_Galactic'), (steel_cs, 'G4_STAINLESS-STEEL')],
                     [('box', 1), ('World', 0)])
    add_element_data(steel_cs)
    json.dump(steel_cs, open(os.path.join(PHYS, 'four-steel-slabs-em-coulomb.json'), 'w'),
              separators=(',', ':'))

    # ... and with the muons of the export: mu-/mu+ ionisation (ICRU73QO / Bragg below
    # 200 keV, Bethe-Bloch to 1 GeV, muon Bethe-Bloch above; ImportProcessClass::mu_ioni = 14)
    # and muon bremsstrahlung above 1 GeV (mu_brems = 15). The export has no multiple
    # scattering or pair production for muons.
    steel_mu = load('four-steel-slabs')
    filter_physics(steel_mu, KEEP_PROCESS + (12, 14, 15), KEEP_PDG + (-13, 13))
    steel_mu = merge([(steel_mu, 'G4_Galactic'), (steel_mu, 'G4_STAINLESS-STEEL')],
                     [('box', 1), ('World', 0)])
    add_element_data(steel_mu)
    json.dump(steel_mu, open(os.path.join(PHYS, 'four-steel-slabs-em-muon.json'), 'w'),
              separators=(',', ':'))


if __name__ == '__main__':
    main()
