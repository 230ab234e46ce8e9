"""Generate `data/geometry/many-faces.org.json`: one unit whose volumes exceed the register
path's 32 faces / 32 intersections in all three of SimpleUnitTracker's search modes
(univ/SimpleUnitTracker.hh:390-640), in the reference's ORANGE JSON input format
(src/orange/OrangeInputIO.json.cc), so that the reference's OrangeParams builds it:

  * background volume (implicit, zorder B): its faces are ALL 113 surfaces of the unit
    (background_intersect: ordered walk + neighbour tests) -- what a CMS-scale mother volume
    looks like to ORANGE;
  * "polyhedron": a convex solid bounded by 40 general planes (simple_intersect, 40 faces);
  * "cheese": a box with 30 spherical holes as ONE volume with internal surfaces
    (complex_intersect, 36 faces, up to 66 intersections);
  * 36 lattice spheres + the 30 holes: ordinary small volumes, the neighbours the background
    search has to find.
"""
import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from make_cms_scale import Unit, box  # noqa: E402

INTERNAL, IMPLICIT, SIMPLE_SAFETY = 1, 2, 4


def build():
    u = Unit('global', box((-30, -30, -30), (30, 30, 30)))
    world = u.surface('sc', [30.0 * 30.0], 'world')
    u.volume('[EXTERIOR]', [(world, +1)], flags=0)
    u.volumes[-1]['zorder'] = 'X'

    # 36 lattice spheres in the upper half
    for ix in range(3):
        for iy in range(3):
            for iz in range(4):
                c = (-6.0 + 6 * ix, -6.0 + 6 * iy, -4.0 + 5 * iz)
                r = 1.5 + 0.1 * ((ix + 2 * iy + 3 * iz) % 5)
                s = u.surface('s', [c[0], c[1], c[2], r * r], 'ball%d%d%d' % (ix, iy, iz))
                u.volume('ball%d%d%d' % (ix, iy, iz), [(s, -1)],
                         bbox=box([x - r for x in c], [x + r for x in c]),
                         flags=SIMPLE_SAFETY)

    # convex polyhedron: 40 general planes tangent to a sphere of radius 3 around (16, 0, 0)
    centre, inr = (16.0, 0.0, 0.0), 3.0
    planes = []
    for i in range(40):
        z = 1 - 2 * (i + 0.5) / 40
        rho = math.sqrt(1 - z * z)
        phi = i * math.pi * (3 - math.sqrt(5))
        nrm = (rho * math.cos(phi), rho * math.sin(phi), z)
        d = sum(a * b for a, b in zip(nrm, centre)) + inr
        planes.append((u.surface('p', [nrm[0], nrm[1], nrm[2], d], 'facet%d' % i), -1))
    u.volume('polyhedron', planes, bbox=box([x - 4.2 for x in centre], [x + 4.2 for x in centre]))

    # "cheese": box [-5,5]^2 x [-22,-12] minus 30 holes, one volume with internal surfaces
    lo, hi = (-5.0, -5.0, -22.0), (5.0, 5.0, -12.0)
    walls = [(u.surface('px', [lo[0]], 'cheese.mx'), +1), (u.surface('px', [hi[0]], 'cheese.px'), -1),
             (u.surface('py', [lo[1]], 'cheese.my'), +1), (u.surface('py', [hi[1]], 'cheese.py'), -1),
             (u.surface('pz', [lo[2]], 'cheese.mz'), +1), (u.surface('pz', [hi[2]], 'cheese.pz'), -1)]
    holes = []
    for ix in range(5):
        for iy in range(3):
            for iz in range(2):
                c = (-4.0 + 2 * ix, -3.0 + 3 * iy, -19.5 + 5 * iz)
                r = 0.6 + 0.05 * ((ix + iy + iz) % 3)
                s = u.surface('s', [c[0], c[1], c[2], r * r], 'hole%d%d%d' % (ix, iy, iz))
                u.volume('hole%d%d%d' % (ix, iy, iz), [(s, -1)],
                         bbox=box([x - r for x in c], [x + r for x in c]), flags=SIMPLE_SAFETY)
                holes.append((s, +1))
    u.volume('cheese', walls + holes, bbox=box(lo, hi), flags=INTERNAL)

    # background: everything else inside the world sphere; faces = all surfaces
    nsurf = len(u.surf_types)
    u.volumes.append({'faces': list(range(nsurf)), 'flags': IMPLICIT | SIMPLE_SAFETY,
                      'logic': '* ~', 'zorder': 'B', 'bbox': None})
    u.labels.append('mother')
    uj = u.to_json()
    return {'_format': 'ORANGE', '_version': 0, 'tol': {'abs': 1e-5, 'rel': 1e-5},
            'universes': [uj]}


def main():
    geo = build()
    out = os.path.join(REPO, 'data', 'geometry', 'many-faces.org.json')
    json.dump(geo, open(out, 'w'))
    print('wrote', out)
    # Full-EM stand-in materials (tools/make_physics.py): the background volume is liquid
    # argon so that showers develop INSIDE the 113-face volume; balls and holes are steel
    import make_physics as mp
    steel, lar = mp.load('four-steel-slabs'), mp.load('lar-sphere')
    for d in (steel, lar):
        mp.filter_physics(d)
    vols = []
    for name in geo['universes'][0]['volume_labels']:
        if name.startswith('[EXTERIOR]'):
            continue
        vols.append((name, 1 if name.startswith(('ball', 'hole')) else 2))
    phys = mp.merge([(steel, 'G4_Galactic'), (steel, 'G4_STAINLESS-STEEL'), (lar, 'lAr')], vols)
    mp.add_element_data(phys)
    ppath = os.path.join(REPO, 'data', 'physics', 'many-faces-steel-lar.json')
    json.dump(phys, open(ppath, 'w'), separators=(',', ':'))
    print('wrote', ppath, os.path.getsize(ppath))


if __name__ == '__main__':
    main()
