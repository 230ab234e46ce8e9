"""Physics-data reader: `b200_import_root` decodes the reference's ROOT export
of celeritas::ImportData (/root/reference/src/celeritas/ext/RootExporter.cc:47-76,
RootImporter.cc, io/ImportData.hh:55-112) without the ROOT library. Checked against

  * the independent Python decoder `tools/rootlite.py` (whose output, the committed
    data/physics/*.json, is what the reference itself transports in every lock-step test:
    the oracle harness feeds it to the reference's own PhysicsParams construction), and
  * values the reference's own importer test expects from the same file
    (/root/reference/test/celeritas/ext/RootImporter.test.cc).
"""
import json
import os
import struct

import pytest

from conftest import data_path

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT_DIR = os.path.join(HERE, 'golden', 'root')
NAMES = ['four-steel-slabs', 'lar-sphere', 'simple-cms']


@pytest.mark.parametrize('name', NAMES)
def test_reader_reproduces_decoded_fixture(name):
    import celeritas_b200 as cb
    got = cb.import_root(os.path.join(ROOT_DIR, name + '.root'))
    want = json.load(open(data_path('physics', name + '.json')))
    assert list(got.keys()) == list(want.keys())
    assert got == want


def test_reference_importer_expectations():
    """The reference's own importer test on the same file
    (/root/reference/test/celeritas/ext/RootImporter.test.cc:79-207): particles, elements,
    geo/phys materials, processes, volumes."""
    import celeritas_b200 as cb
    d = cb.import_root(os.path.join(ROOT_DIR, 'four-steel-slabs.root'))
    assert [p['name'] for p in d['particles']] == ['e+', 'e-', 'gamma', 'mu+', 'mu-']
    assert [p['pdg'] for p in d['particles']] == [-11, 11, 22, -13, 13]
    assert [e['name'] for e in d['elements']] == ['Fe', 'Cr', 'Ni', 'H']
    assert [m['name'] for m in d['geo_materials']] == ['G4_STAINLESS-STEEL', 'G4_Galactic']
    assert [m['geo_material_id'] for m in d['phys_materials']] == [1, 0]
    assert len(d['processes']) == 15
    ioni = [p for p in d['processes'] if p['particle_pdg'] == 11 and p['process_class'] == 7]
    assert len(ioni) == 1 and ioni[0]['process_type'] == 2  # e_ioni, electromagnetic
    assert [m['model_class'] for m in ioni[0]['models']] == [10]  # moller_bhabha
    assert [v['phys_material_id'] for v in d['volumes']] == [1, 1, 1, 1, 0]
    assert [v['name'] for v in d['volumes']] == [
        'box0x125555be0', 'box0x125556d20', 'box0x125557160', 'box0x1255575a0',
        'World0x125555f10']
    assert [v['solid_name'] for v in d['volumes']] == [
        'box0x125555b70', 'box0x125556c70', 'box0x1255570a0', 'box0x125557500',
        'World0x125555ea0']
    # physics vectors are strictly increasing grids with one value per node
    count = 0
    for proc in d['processes']:
        for t in proc['tables']:
            for v in t['physics_vectors']:
                assert len(v['x']) == len(v['y']) >= 2
                assert all(a < b for a, b in zip(v['x'], v['x'][1:]))
                count += 1
    assert count > 20


def test_reader_rejects_what_it_cannot_read(tmp_path):
    import celeritas_b200 as cb
    with pytest.raises(cb.B200Error, match='cannot open'):
        cb.import_root(str(tmp_path / 'missing.root'))
    bad = tmp_path / 'bad.root'
    bad.write_bytes(b'not a root file' * 10)
    with pytest.raises(cb.B200Error, match='not a ROOT file'):
        cb.import_root(str(bad))
    # a truncated export: header intact, records cut off
    raw = open(os.path.join(ROOT_DIR, 'lar-sphere.root'), 'rb').read()
    cut = tmp_path / 'cut.root'
    cut.write_bytes(raw[:len(raw) // 2])
    with pytest.raises(cb.B200Error, match='ROOT physics file'):
        cb.import_root(str(cut))
    # a damaged zlib stream inside an otherwise well-formed record
    at = raw.index(b'ZL\x08') + 9 + 16
    blob = bytearray(raw)
    for i in range(at, at + 64):
        blob[i] ^= 0x5a
    corrupt = tmp_path / 'corrupt.root'
    corrupt.write_bytes(bytes(blob))
    with pytest.raises(cb.B200Error, match='ROOT physics file'):
        cb.import_root(str(corrupt))
