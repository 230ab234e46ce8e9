"""Generate the CMS-scale stand-in geometry (BASELINE configs 3 and 4).

The reference's CMS2018 problem is a GDML file converted to ORANGE by `src/orange/g4org`,
which needs Geant4; neither the GDML nor Geant4 exists in this image. This script writes a
detector of comparable *navigation* character in the reference's own ORANGE JSON input
format (`src/orange/OrangeInputIO.json.cc:216-330,382-432`), so that the reference's
`OrangeParams` builds it (surfaces, BIH trees, universe tables) and the CUDA code loads the
flattened result like every other problem:

  level 0  global        world box, beam pipe, solenoid + 6 daughter placements
  level 1  tracker_u     22 coaxial cylinders x 11 z planes              -> 276 volumes (BIH)
           hcal_u        23 cylinders x 11 z planes x 8 phi sectors
                         (general planes through the beam axis)           -> 2304 volumes (BIH)
           muon_u        7 cylinders x 7 z planes                         -> 64 volumes
           ecal_u        one fill volume holding a rect array
           endcap_u      (placed twice, +z and -z) one fill volume holding a rect array
  level 2  ecal_arr      18 x 18 x 30 rect array of 20 cm cells, clipped by the parent shell
           endcap_arr    20 x 20 x 1 rect array of 30 x 30 x 100 cm towers
  level 3  ecal_cell_x/y 10 absorber/gap slabs along x or y (checkerboard)
           endcap_cell   40 slabs along z

Materials (names matched by `GeoMaterialParams`, `src/celeritas/geo/GeoMaterialParams.cc:84-140`):
steel absorbers, liquid-argon gaps and tracker layers, vacuum elsewhere: the same stand-in
materials as every other problem here (tools/make_physics.py).

Outputs: data/geometry/cms-scale.org.json, data/physics/cms-scale-steel-lar.json
"""
import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)

VAC, STEEL, LAR = 0, 1, 2
EXTERIOR = {"faces": [], "flags": 2, "logic": "* ~", "zorder": "x"}


class Unit:
    def __init__(self, name, bbox):
        self.name = name
        self.bbox = bbox
        self.surf_types, self.surf_data, self.surf_sizes, self.surf_labels = [], [], [], []
        self.volumes, self.labels = [], []
        self.parent_cells, self.daughters, self.transforms = [], [], []

    def surface(self, type_, data, label):
        self.surf_types.append(type_)
        self.surf_data.extend(data)
        self.surf_sizes.append(len(data))
        self.surf_labels.append(label)
        return len(self.surf_types) - 1

    def volume(self, label, senses, bbox=None, flags=0, logic=None, daughter=None):
        """senses: list of (surface id, +1 outside / -1 inside); a pure intersection."""
        senses = sorted(senses)
        faces = [s for s, _ in senses]
        if logic is None:
            if not faces:
                logic = "*"
            else:
                toks = []
                for i, (_, sign) in enumerate(senses):
                    toks.append(str(i))
                    if sign < 0:
                        toks.append("~")
                    if i > 0:
                        toks.append("&")
                logic = " ".join(toks)
        v = {"faces": faces, "logic": logic}
        if flags:
            v["flags"] = flags
        if bbox is not None:
            v["bbox"] = bbox
        self.volumes.append(v)
        self.labels.append(label)
        if daughter is not None:
            univ, trans = daughter
            self.parent_cells.append(len(self.volumes) - 1)
            self.daughters.append(univ)
            self.transforms.append(trans)
        return len(self.volumes) - 1

    def to_json(self):
        j = {"_type": "unit", "md": {"name": self.name}, "bbox": self.bbox,
             "surfaces": {"types": self.surf_types, "data": self.surf_data,
                          "sizes": self.surf_sizes},
             "surface_labels": self.surf_labels,
             "volumes": self.volumes, "volume_labels": self.labels}
        if self.daughters:
            j["parent_cells"] = self.parent_cells
            j["daughters"] = self.daughters
            j["transforms"] = self.transforms
        return j


def box(lo, hi):
    return [[float(x) for x in lo], [float(x) for x in hi]]


def sector_bbox(r0, r1, a0, a1, z0, z1):
    """Bounding box of an annular sector (angles in radians, a0 < a1 <= a0 + pi)."""
    pts = []
    for r in (r0, r1):
        for a in (a0, a1):
            pts.append((r * math.cos(a), r * math.sin(a)))
    k = math.ceil(a0 / (math.pi / 2))
    while k * math.pi / 2 <= a1:
        pts.append((r1 * math.cos(k * math.pi / 2), r1 * math.sin(k * math.pi / 2)))
        k += 1
    xs, ys = [p[0] for p in pts], [p[1] for p in pts]
    eps = 1e-6 * r1
    return box((min(xs) - eps, min(ys) - eps, z0), (max(xs) + eps, max(ys) + eps, z1))


def shells_unit(name, radii, zplanes, rmax, zmax, namer, nphi=0):
    """Coaxial shells x z segments (x phi sectors): len(radii)+1 zones, len(zplanes)+1 segments.

    The innermost zone is a full cylinder and the outermost zone / segments are unbounded, so
    that no daughter surface coincides with the parent volume's own boundary."""
    u = Unit(name, box((-rmax, -rmax, -zmax), (rmax, rmax, zmax)))
    u.volumes.append(dict(EXTERIOR))
    u.labels.append("[EXTERIOR]")
    cyl = [u.surface("czc", [float(r) ** 2], "%s.r%d" % (name, i)) for i, r in enumerate(radii)]
    pz = [u.surface("pz", [float(z)], "%s.z%d" % (name, i)) for i, z in enumerate(zplanes)]
    pp = []
    if nphi:
        assert nphi % 2 == 0
        for j in range(nphi // 2):
            a = 2 * math.pi * j / nphi
            pp.append(u.surface("p", [-math.sin(a), math.cos(a), 0.0, 0.0], "%s.phi%d" % (name, j)))
    mats = {}
    for ir in range(len(radii) + 1):
        r0 = radii[ir - 1] if ir > 0 else 0.0
        r1 = radii[ir] if ir < len(radii) else rmax
        for iz in range(len(zplanes) + 1):
            z0 = zplanes[iz - 1] if iz > 0 else -zmax
            z1 = zplanes[iz] if iz < len(zplanes) else zmax
            senses = []
            if ir > 0:
                senses.append((cyl[ir - 1], +1))
            if ir < len(radii):
                senses.append((cyl[ir], -1))
            if iz > 0:
                senses.append((pz[iz - 1], +1))
            if iz < len(zplanes):
                senses.append((pz[iz], -1))
            for ip in range(max(nphi, 1)):
                s = list(senses)
                if nphi:
                    half = nphi // 2
                    amid = 2 * math.pi * (ip + 0.5) / nphi
                    for j in (ip % half, (ip + 1) % half):
                        # sense of the sector's interior w.r.t. plane j (normal at a_j + pi/2)
                        aj = 2 * math.pi * j / nphi
                        val = -math.sin(aj) * math.cos(amid) + math.cos(aj) * math.sin(amid)
                        s.append((pp[j], +1 if val > 0 else -1))
                    a0, a1 = 2 * math.pi * ip / nphi, 2 * math.pi * (ip + 1) / nphi
                    bb = sector_bbox(r0, r1, a0, a1, z0, z1)
                else:
                    bb = box((-r1, -r1, z0), (r1, r1, z1))
                label, mat = namer(ir, iz, ip)
                u.volume(label, s, bbox=bb)
                mats[label] = mat
    return u, mats


def slab_cell(name, size, axis, nslab, frac_abs, prefix):
    """Box [0,size] with nslab (absorber, gap) pairs along one axis."""
    u = Unit(name, box((0, 0, 0), size))
    u.volumes.append(dict(EXTERIOR))
    u.labels.append("[EXTERIOR]")
    length = size[axis]
    pitch = length / nslab
    cuts = []
    for i in range(nslab):
        cuts.append(i * pitch + frac_abs * pitch)
        if i + 1 < nslab:
            cuts.append((i + 1) * pitch)
    planes = [u.surface("p" + "xyz"[axis], [float(c)], "%s.c%d" % (name, i))
              for i, c in enumerate(cuts)]
    mats = {}
    for i in range(len(cuts) + 1):
        lo = cuts[i - 1] if i > 0 else 0.0
        hi = cuts[i] if i < len(cuts) else length
        senses = []
        if i > 0:
            senses.append((planes[i - 1], +1))
        if i < len(cuts):
            senses.append((planes[i], -1))
        blo, bhi = [0.0, 0.0, 0.0], list(size)
        blo[axis], bhi[axis] = lo, hi
        label = "%s_%s_%d" % (prefix, "abs" if i % 2 == 0 else "gap", i // 2)
        u.volume(label, senses, bbox=box(blo, bhi))
        mats[label] = STEEL if i % 2 == 0 else LAR
    return u, mats


def array_holder(name, lo, hi, arr_index):
    u = Unit(name, box(lo, hi))
    u.volumes.append(dict(EXTERIOR))
    u.labels.append("[EXTERIOR]")
    u.volume(name + "_fill", [], bbox=box(lo, hi), daughter=(arr_index, [float(x) for x in lo]))
    return u


def rect_array(name, grids, pick):
    nx, ny, nz = (len(g) - 1 for g in grids)
    daughters, translations = [], []
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                daughters.append(pick(i, j, k))
                translations.extend([grids[0][i], grids[1][j], grids[2][k]])
    return {"_type": "rectarray", "md": {"name": name}, "daughters": daughters,
            "translations": [float(t) for t in translations],
            "x": [float(v) for v in grids[0]], "y": [float(v) for v in grids[1]],
            "z": [float(v) for v in grids[2]]}


def frange(lo, hi, step):
    n = int(round((hi - lo) / step))
    return [lo + i * step for i in range(n + 1)]


def build():
    # universe indices
    GLOBAL, TRACKER, HCAL, MUON, ECAL, ECAL_ARR, CELL_X, CELL_Y, ENDCAP, ENDCAP_ARR, ENDCAP_CELL \
        = range(11)
    mats = {}

    g = Unit("global", box((-1000, -1000, -1500), (1000, 1000, 1500)))
    wb = [g.surface(t, [float(v)], "world_box." + n)
          for t, v, n in (("px", -1000, "mx"), ("px", 1000, "px"), ("py", -1000, "my"),
                          ("py", 1000, "py"), ("pz", -1500, "mz"), ("pz", 1500, "pz"))]
    z_m4 = g.surface("pz", [-400.0], "endcap.mz")
    z_m3 = g.surface("pz", [-300.0], "barrel.mz")
    z_p3 = g.surface("pz", [300.0], "barrel.pz")
    z_p4 = g.surface("pz", [400.0], "endcap.pz")
    r_pipe, r_trk, r_ecal, r_hcal, r_sol, r_mu = [
        g.surface("czc", [float(r) ** 2], n + ".coz")
        for r, n in ((3, "beam_pipe"), (120, "tracker"), (180, "ecal"), (300, "hcal"),
                     (380, "solenoid"), (700, "muon"))]
    g.volumes.append({"faces": wb, "flags": 1, "logic": "0 1 ~ & 2 & 3 ~ & 4 & 5 ~ & ~"})
    g.labels.append("[EXTERIOR]")

    def barrel(label, rin, rout, zlo, zhi, R, Z0, Z1, **kw):
        s = [(zlo, +1), (zhi, -1), (rout, -1)]
        if rin is not None:
            s.append((rin, +1))
        return g.volume(label, s, bbox=box((-R, -R, Z0), (R, R, Z1)), **kw)

    barrel("beam_pipe", None, r_pipe, z_m3, z_p3, 3, -300, 300)
    mats["beam_pipe"] = VAC
    barrel("tracker", r_pipe, r_trk, z_m3, z_p3, 120, -300, 300, daughter=(TRACKER, []))
    barrel("ecal", r_trk, r_ecal, z_m3, z_p3, 180, -300, 300, daughter=(ECAL, []))
    barrel("hcal", r_ecal, r_hcal, z_m3, z_p3, 300, -300, 300, daughter=(HCAL, []))
    barrel("solenoid", r_hcal, r_sol, z_m4, z_p4, 380, -400, 400)
    mats["solenoid"] = STEEL
    barrel("muon", r_sol, r_mu, z_m4, z_p4, 700, -400, 400, daughter=(MUON, []))
    barrel("endcap_p", None, r_hcal, z_p3, z_p4, 300, 300, 400,
           daughter=(ENDCAP, [0.0, 0.0, 350.0]))
    barrel("endcap_m", None, r_hcal, z_m4, z_m3, 300, -400, -300,
           daughter=(ENDCAP, [0.0, 0.0, -350.0]))
    g.volumes.append({"faces": wb + [z_m4, z_p4, r_mu], "flags": 1,
                      "bbox": box((-1000, -1000, -1500), (1000, 1000, 1500)),
                      "logic": "0 1 ~ & 2 & 3 ~ & 4 & 5 ~ & 6 7 ~ & 8 ~ & ~ &"})
    g.labels.append("world")
    mats["world"] = VAC

    # tracker: 11 layers of 0.5 cm 'silicon' (lAr stand-in) at r = 10 .. 110
    radii = []
    for k in range(1, 12):
        radii += [10.0 * k, 10.0 * k + 0.5]
    trk, m = shells_unit(
        "tracker_u", radii, frange(-250, 250, 50), 125.0, 310.0,
        lambda ir, iz, ip: ("trk_%s_%d_%d" % ("si" if ir % 2 == 1 else "gas", ir // 2, iz),
                            LAR if ir % 2 == 1 else VAC))
    mats.update(m)

    # hadron calorimeter: 12 x (7 cm absorber + 3 cm gap) from r = 180, 8 phi sectors
    radii = []
    for k in range(12):
        radii += [180.0 + 10 * k + 7.0, 180.0 + 10 * (k + 1)]
    radii = radii[:-1]
    hcal, m = shells_unit(
        "hcal_u", radii, frange(-275, 275, 50), 305.0, 310.0,
        lambda ir, iz, ip: ("hcal_%s_%d_%d_%d" % ("abs" if ir % 2 == 0 else "gap", ir // 2, iz, ip),
                            STEEL if ir % 2 == 0 else LAR),
        nphi=8)
    mats.update(m)

    # muon system: 4 x (60 cm iron + 20 cm gap)
    radii = []
    for k in range(4):
        radii += [380.0 + 80 * k + 60.0, 380.0 + 80 * (k + 1)]
    radii = radii[:-1]
    muon, m = shells_unit(
        "muon_u", radii, frange(-300, 300, 100), 705.0, 410.0,
        lambda ir, iz, ip: ("mu_%s_%d_%d" % ("fe" if ir % 2 == 0 else "gap", ir // 2, iz),
                            STEEL if ir % 2 == 0 else VAC))
    mats.update(m)

    ecal = array_holder("ecal_u", (-180, -180, -300), (180, 180, 300), ECAL_ARR)
    ecal_arr = rect_array("ecal_arr", (frange(0, 360, 20), frange(0, 360, 20), frange(0, 600, 20)),
                          lambda i, j, k: CELL_X if (i + j + k) % 2 == 0 else CELL_Y)
    cell_x, m = slab_cell("ecal_cell_x", (20.0, 20.0, 20.0), 0, 5, 0.6, "ecalx")
    mats.update(m)
    cell_y, m = slab_cell("ecal_cell_y", (20.0, 20.0, 20.0), 1, 5, 0.6, "ecaly")
    mats.update(m)

    endcap = array_holder("endcap_u", (-300, -300, -50), (300, 300, 50), ENDCAP_ARR)
    endcap_arr = rect_array("endcap_arr", (frange(0, 600, 30), frange(0, 600, 30), [0.0, 100.0]),
                            lambda i, j, k: ENDCAP_CELL)
    endcap_cell, m = slab_cell("endcap_cell", (30.0, 30.0, 100.0), 2, 20, 0.7, "endcap")
    mats.update(m)

    universes = [g.to_json(), trk.to_json(), hcal.to_json(), muon.to_json(), ecal.to_json(),
                 ecal_arr, cell_x.to_json(), cell_y.to_json(), endcap.to_json(), endcap_arr,
                 endcap_cell.to_json()]
    return {"_format": "ORANGE", "_version": 0, "universes": universes}, mats


def main():
    geo, mats = build()
    path = os.path.join(REPO, "data", "geometry", "cms-scale.org.json")
    json.dump(geo, open(path, "w"), separators=(",", ":"))
    nvol = sum(len(u.get("volumes", [])) for u in geo["universes"])
    print("cms-scale.org.json: %d universes, %d unit volumes, %d bytes"
          % (len(geo["universes"]), nvol, os.path.getsize(path)))

    import make_physics as mp
    steel, lar = mp.load("four-steel-slabs"), mp.load("lar-sphere")
    for d in (steel, lar):
        mp.filter_physics(d)
    vols = sorted(mats.items())
    phys = mp.merge([(steel, "G4_Galactic"), (steel, "G4_STAINLESS-STEEL"), (lar, "lAr")], vols)
    mp.add_element_data(phys)
    ppath = os.path.join(REPO, "data", "physics", "cms-scale-steel-lar.json")
    json.dump(phys, open(ppath, "w"), separators=(",", ":"))
    print("cms-scale-steel-lar.json: %d volumes, %d bytes" % (len(vols), os.path.getsize(ppath)))


if __name__ == "__main__":
    main()
