"""rootlite: a minimal, dependency-free reader for the reference's bundled physics exports.

The reference stores its Geant4-derived physics tables (`celeritas::ImportData`,
/root/reference/src/celeritas/io/ImportData.hh:55-112) in ROOT files written by
`RootExporter` (/root/reference/src/celeritas/ext/RootExporter.cc:47-76): one TTree
`geant4_data` with a single entry whose branch `ImportData` is split member-wise into
one TBasket per leaf member.  ROOT is not available in this image, so this module
decodes exactly the subset of the ROOT file format those files use:

  * TFile header and TKey records (big-endian; zlib "ZL" compressed blocks),
  * the embedded TStreamerInfo list (gives the member list of every Import* struct as
    it was written, so schema differences with the current headers are visible),
  * split-branch baskets holding member-wise streamed STL collections of structs.

It is a build-time/test-time tool: `python tools/rootlite.py file.root out.json` converts
one file to plain JSON (the committed fixtures under tests/golden/physics/).
"""
import json
import struct
import sys
import zlib

kByteCountMask = 0x40000000
kNewClassTag = 0xFFFFFFFF
kClassMask = 0x80000000
kMapOffset = 2
kMemberWise = 0x4000


class Buf:
    def __init__(self, data, pos=0, origin=0):
        self.d = data
        self.p = pos
        self.origin = origin  # displacement used for class/object references

    def u8(self):
        v = self.d[self.p]; self.p += 1; return v

    def _un(self, fmt, n):
        v, = struct.unpack_from(fmt, self.d, self.p); self.p += n; return v

    def i16(self): return self._un('>h', 2)
    def u16(self): return self._un('>H', 2)
    def i32(self): return self._un('>i', 4)
    def u32(self): return self._un('>I', 4)
    def i64(self): return self._un('>q', 8)
    def u64(self): return self._un('>Q', 8)
    def f32(self): return self._un('>f', 4)
    def f64(self): return self._un('>d', 8)

    def string(self):
        n = self.u8()
        if n == 255:
            n = self.u32()
        s = self.d[self.p:self.p + n].decode('latin1'); self.p += n
        return s

    def cstring(self):
        e = self.d.index(b'\0', self.p)
        s = self.d[self.p:e].decode('latin1'); self.p = e + 1
        return s

    def version(self):
        """Read [bytecount] version; return (version, end_pos or None)."""
        start = self.p
        bc = self.u32()
        if bc & kByteCountMask:
            end = start + 4 + (bc & ~kByteCountMask)
            v = self.u16()
        else:
            self.p = start
            end = None
            v = self.u16()
        return v, end


def read_keys(fn):
    f = open(fn, 'rb').read()
    assert f[:4] == b'root'
    fVersion, fBEGIN, fEND = struct.unpack('>iii', f[4:16])
    assert fVersion < 1000000, 'large files unsupported'
    pos = fBEGIN
    keys = []
    while pos < fEND:
        nbytes, = struct.unpack('>i', f[pos:pos + 4])
        if nbytes < 0:
            pos += -nbytes
            continue
        ver, objlen, datime, keylen, cycle = struct.unpack('>hiihh', f[pos + 4:pos + 18])
        b = Buf(f, pos + 18 + (16 if ver > 1000 else 8))
        cls = b.string(); name = b.string(); title = b.string()
        data = f[pos + keylen:pos + nbytes]
        if objlen != nbytes - keylen:
            out = b''
            p = 0
            while len(out) < objlen:
                assert data[p:p + 2] == b'ZL', data[p:p + 9]
                c = data[p + 3] | data[p + 4] << 8 | data[p + 5] << 16
                out += zlib.decompress(data[p + 9:p + 9 + c])
                p += 9 + c
            data = out
        keys.append(dict(pos=pos, cls=cls, name=name, title=title, keylen=keylen,
                         data=data, hdr=f[pos:pos + keylen]))
        pos += nbytes
    return keys


# --------------------------------------------------------------------------- #
# TStreamerInfo list
# --------------------------------------------------------------------------- #
class ObjReader:
    """Object-wise reader for the few TObject-derived classes in the StreamerInfo list."""

    def __init__(self, data, keylen):
        self.b = Buf(data)
        self.keylen = keylen
        self.classes = {}  # tag position -> class name

    def read_tobject(self):
        b = self.b
        v = b.u16()
        if v & kByteCountMask >> 16:
            b.p += 4
        b.u32()  # fUniqueID
        bits = b.u32()
        if bits & (1 << 4):  # kIsReferenced
            b.p += 2

    def read_tnamed(self):
        b = self.b
        v, end = b.version()
        self.read_tobject()
        name = b.string(); title = b.string()
        return name, title

    def read_object_any(self):
        b = self.b
        start = b.p
        bc = b.u32()
        if not (bc & kByteCountMask) or bc == kNewClassTag:
            tag = bc; bc = 0; end = None
            tagpos = start
        else:
            end = start + 4 + (bc & ~kByteCountMask)
            tagpos = b.p
            tag = b.u32()
        if tag == 0:
            return None
        if not (tag & kClassMask):
            # reference to an already-read object: not needed here
            return ('ref', tag)
        if tag == kNewClassTag:
            cname = b.cstring()
            self.classes[tagpos + self.keylen + kMapOffset] = cname
        else:
            ref = tag & ~kClassMask
            cname = self.classes[ref]
        obj = self.read_class(cname)
        if end is not None:
            assert b.p == end, (cname, b.p, end)
        return obj

    def read_class(self, cname):
        b = self.b
        if cname == 'TStreamerInfo':
            v, end = b.version()
            name, title = self.read_tnamed()
            checksum = b.u32(); clsver = b.i32()
            elements = self.read_object_any()
            assert b.p == end
            return dict(name=name, checksum=checksum, version=clsver, elements=elements)
        if cname == 'TObjArray':
            v, end = b.version()
            self.read_tobject()
            name = b.string()
            n = b.i32(); low = b.i32()
            items = [self.read_object_any() for _ in range(n)]
            assert b.p == end
            return items
        if cname == 'TList':
            v, end = b.version()
            self.read_tobject()
            name = b.string(); n = b.i32()
            items = []
            for _ in range(n):
                items.append(self.read_object_any())
                nopt = b.u8(); b.p += nopt
            return items
        if cname.startswith('TStreamer'):
            v, end = b.version()
            if cname == 'TStreamerSTLstring':
                b.version()  # nested TStreamerSTL header
            el = self.read_element()
            if cname == 'TStreamerBase':
                el['base_version'] = b.i32()
            elif cname in ('TStreamerSTL', 'TStreamerSTLstring'):
                el['stltype'] = b.i32(); el['ctype'] = b.i32()
            el['kind'] = cname
            b.p = end
            return el
        if cname == 'TObjString':
            v, end = b.version()
            self.read_tobject()
            s = b.string()
            b.p = end
            return s
        raise NotImplementedError(cname)

    def read_element(self):
        b = self.b
        v, end = b.version()
        name, title = self.read_tnamed()
        ftype = b.i32(); size = b.i32(); alen = b.i32(); adim = b.i32()
        if v == 1:
            n = b.i32(); b.p += 4 * n
        else:
            b.p += 20
        tname = b.string()
        b.p = end
        return dict(name=name, title=title, type=ftype, size=size, typename=tname)


def read_streamer_infos(keys):
    k = [k for k in keys if k['cls'] == 'TList' and k['name'] == 'StreamerInfo'][0]
    r = ObjReader(k['data'], k['keylen'])
    items = r.read_class('TList')
    return {it['name']: it for it in items if isinstance(it, dict)}



# --------------------------------------------------------------------------- #
# Split-branch decoding
# --------------------------------------------------------------------------- #
BASIC = {1: ('b', 1), 2: ('>h', 2), 3: ('>i', 4), 4: ('>q', 8), 5: ('>f', 4), 8: ('>d', 8),
         11: ('B', 1), 12: ('>H', 2), 13: ('>I', 4), 14: ('>Q', 8), 16: ('>q', 8),
         17: ('>Q', 8), 18: ('?', 1)}
BASIC_NAMES = {'int': 3, 'unsigned int': 13, 'double': 8, 'float': 5, 'bool': 18,
               'unsigned': 13, 'long': 4, 'unsigned long': 14, 'short': 2, 'char': 1}


def split_template(tname):
    """'vector<pair<unsigned int,double> >' -> ('vector', ['pair<unsigned int,double>'])"""
    tname = tname.strip()
    if '<' not in tname:
        return tname, []
    outer = tname[:tname.index('<')]
    inner = tname[tname.index('<') + 1:tname.rindex('>')]
    args = []
    depth = 0
    cur = ''
    for ch in inner:
        if ch == '<':
            depth += 1
        elif ch == '>':
            depth -= 1
        if ch == ',' and depth == 0:
            args.append(cur.strip()); cur = ''
        else:
            cur += ch
    args.append(cur.strip())
    return outer, args


class Decoder:
    def __init__(self, fn):
        self.keys = read_keys(fn)
        self.infos = read_streamer_infos(self.keys)
        self.baskets = {}
        for k in self.keys:
            if k['cls'] == 'TBasket':
                assert k['name'] not in self.baskets, 'multi-basket branches unsupported'
                self.baskets[k['name']] = k['data']

    # -- helpers -------------------------------------------------------------
    def basic(self, b, tcode):
        fmt, n = BASIC[tcode]
        v, = struct.unpack_from(fmt, b.d, b.p); b.p += n
        return v

    def value_kind(self, tname):
        """Classify a C++ type name: ('basic', code) | ('string',) | ('stl', elemtype) | ('class', name)."""
        tname = tname.strip()
        if tname in BASIC_NAMES:
            return ('basic', BASIC_NAMES[tname])
        if tname in ('string', 'std::string'):
            return ('string',)
        outer, args = split_template(tname)
        if outer == 'vector':
            return ('stl', args[0])
        if outer == 'map':
            return ('stl', 'pair<%s,%s>' % (args[0], args[1]))
        if tname in self.infos:
            return ('class', tname)
        # enums are stored as int
        return ('basic', 3)

    def elem_kind(self, e):
        if e['type'] in BASIC:
            return ('basic', e['type'])
        return self.value_kind(e['typename'])

    # -- object-wise ------------------------------------------------------------
    def read_object(self, b, cname):
        v, end = b.version()
        if v == 0:
            b.u32()  # checksum of a class without ClassDef
        out = {}
        for e in self.infos[cname]['elements']:
            out[e['name']] = self.read_value(b, self.elem_kind(e))
        assert end is None or b.p == end, (cname, b.p, end)
        return out

    def read_value(self, b, kind):
        if kind[0] == 'basic':
            return self.basic(b, kind[1])
        if kind[0] == 'string':
            return b.string()
        if kind[0] == 'class':
            return self.read_object(b, kind[1])
        if kind[0] == 'stl':
            hdr = self.read_stl_header(b)
            val = self.read_stl_body(b, kind[1], hdr)
            assert hdr[1] is None or b.p == hdr[1]
            return val
        raise NotImplementedError(kind)

    # -- STL collections ----------------------------------------------------------
    def read_stl_header(self, b):
        v, end = b.version()
        memberwise = bool(v & kMemberWise)
        if memberwise:
            cv = b.u16()
            if cv == 0:
                b.u32()
        return memberwise, end

    def read_stl_body(self, b, etype, hdr):
        """One collection instance: count + contents (header already consumed)."""
        memberwise, _ = hdr
        n = b.u32()
        if n == 0:
            return []
        ek = self.value_kind(etype)
        if ek[0] == 'class' and memberwise:
            return self.read_memberwise(b, ek[1], n)
        if ek[0] == 'stl':
            # nested collection of basic collections, e.g. vector<vector<double>>
            out = []
            for _ in range(n):
                out.append(self.read_stl_body(b, ek[1], (False, None)))
            return out
        return [self.read_value(b, ek) for _ in range(n)]

    def read_column(self, b, kind, n):
        """The same member of n consecutive objects (member-wise layout)."""
        if kind[0] == 'basic':
            return [self.basic(b, kind[1]) for _ in range(n)]
        if kind[0] == 'string':
            # strings inside a member-wise collection carry no extra header
            return [b.string() for _ in range(n)]
        if kind[0] == 'class':
            return [self.read_object(b, kind[1]) for _ in range(n)]
        if kind[0] == 'stl':
            hdr = self.read_stl_header(b)
            out = [self.read_stl_body(b, kind[1], hdr) for _ in range(n)]
            assert hdr[1] is None or b.p == hdr[1], (kind, b.p, hdr)
            return out
        raise NotImplementedError(kind)

    def read_memberwise(self, b, cname, n):
        cols = {}
        for e in self.infos[cname]['elements']:
            cols[e['name']] = self.read_column(b, self.elem_kind(e), n)
        return [{k: cols[k][i] for k in cols} for i in range(n)]

    # -- branches -------------------------------------------------------------------
    def branch_column(self, name, kind, n):
        """Top-level split branch `name` holding one member of n objects."""
        if kind[0] == 'class' and name not in self.baskets:
            # nested struct: split further into sub-branches
            cols = {}
            for e in self.infos[kind[1]]['elements']:
                cols[e['name']] = self.branch_column(name + '.' + e['name'], self.elem_kind(e), n)
            return [{k: cols[k][i] for k in cols} for i in range(n)]
        b = Buf(self.baskets[name])
        if kind[0] == 'string':
            v, end = b.version()
            return [b.string() for _ in range(n)]
        return self.read_column(b, kind, n)

    def branch_collection(self, name, etype):
        n = Buf(self.baskets[name]).u32()
        ek = self.value_kind(etype)
        assert ek[0] == 'class'
        cols = {}
        for e in self.infos[ek[1]]['elements']:
            cols[e['name']] = self.branch_column(name + '.' + e['name'], self.elem_kind(e), n)
        return [{k: cols[k][i] for k in cols} for i in range(n)]

    def branch_struct(self, prefix, cname):
        out = {}
        for e in self.infos[cname]['elements']:
            name = (prefix + '.' if prefix else '') + e['name']
            if name.startswith('optical_'):
                continue  # optical physics is outside the EM track loop
            kind = self.elem_kind(e)
            if kind[0] == 'stl' and name in self.baskets and (name + '.') in ''.join(
                    k + ' ' for k in self.baskets if k.startswith(name + '.')):
                items = self.branch_collection(name, kind[1])
                outer, _ = split_template(e['typename'])
                if outer == 'map':
                    items = {str(it['first']): it['second'] for it in items}
                out[e['name']] = items
            elif kind[0] == 'class':
                if name in self.baskets and not any(k.startswith(name + '.') for k in self.baskets):
                    out[e['name']] = self.read_object(Buf(self.baskets[name]), kind[1])
                else:
                    out[e['name']] = self.branch_struct(name, kind[1])
            elif kind[0] == 'stl':
                # unsplit collection (of basic type) stored in one basket
                b = Buf(self.baskets[name])
                out[e['name']] = self.read_value(b, kind)
            elif kind[0] == 'string':
                b = Buf(self.baskets[name]); b.version()
                out[e['name']] = b.string()
            else:
                out[e['name']] = self.basic(Buf(self.baskets[name]), kind[1])
        return out


def load_import_data(fn):
    """Decode the single `ImportData` entry of a reference physics export."""
    d = Decoder(fn)
    return d.branch_struct('', 'celeritas::ImportData')


if __name__ == '__main__':
    if len(sys.argv) == 2:
        keys = read_keys(sys.argv[1])
        infos = read_streamer_infos(keys)
        for name, it in infos.items():
            print(name, 'v%d' % it['version'], hex(it['checksum']))
            for e in it['elements']:
                print('    %-28s %-60s type=%d %s' % (e['name'], e['typename'], e['type'], e['kind']))
    else:
        data = load_import_data(sys.argv[1])
        with open(sys.argv[2], 'w') as f:
            json.dump(data, f, separators=(',', ':'))
