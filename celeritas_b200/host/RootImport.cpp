//---------------------------------------------------------------------------//
// Reader for the reference's physics-data wire format.
//
// The reference stores `celeritas::ImportData` (src/celeritas/io/ImportData.hh:55-112) in a
// ROOT file written by `RootExporter` (src/celeritas/ext/RootExporter.cc:47-76): one TTree
// `geant4_data` with ONE entry whose branch `ImportData` is split member-wise, one TBasket
// per leaf member. `RootImporter` (src/celeritas/ext/RootImporter.cc) reads it back through
// the ROOT library; this reader does it without ROOT, decoding the subset of the file
// format those exports use:
//   * the TFile header and the chain of TKey records (big-endian, zlib "ZL" blocks),
//   * the embedded TStreamerInfo list: the member list of every Import* struct AS WRITTEN
//     (the decoder is driven by it, so a schema change in the reference shows up as data,
//     not as a silent misread),
//   * split-branch baskets with object-wise and member-wise streamed STL collections.
// The result is a JSON document with the reference's member names (optical_* members are
// skipped: outside the EM track loop).
//---------------------------------------------------------------------------//
#include "RootImport.hh"

#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include <nlohmann/json.hpp>
#include <zlib.h>

namespace b200
{
namespace
{
using json = nlohmann::ordered_json;
using Bytes = std::vector<unsigned char>;

constexpr uint32_t kByteCountMask = 0x40000000u;
constexpr uint32_t kNewClassTag = 0xFFFFFFFFu;
constexpr uint32_t kClassMask = 0x80000000u;
constexpr uint32_t kMapOffset = 2;
constexpr uint32_t kMemberWise = 0x4000u;

[[noreturn]] void fail(std::string const& what)
{
    throw std::runtime_error("ROOT physics file: " + what);
}

//! Big-endian cursor over a byte range; every read is bounds-checked
struct Buf
{
    unsigned char const* d{nullptr};
    size_t n{0};
    size_t p{0};

    Buf() = default;
    explicit Buf(Bytes const& b, size_t pos = 0) : d(b.data()), n(b.size()), p(pos) {}

    void need(size_t k) const
    {
        if (p + k > n)
            fail("truncated record");
    }
    uint64_t be(size_t k)
    {
        need(k);
        uint64_t v = 0;
        for (size_t i = 0; i < k; ++i)
            v = (v << 8) | d[p + i];
        p += k;
        return v;
    }
    uint8_t u8() { return static_cast<uint8_t>(be(1)); }
    uint16_t u16() { return static_cast<uint16_t>(be(2)); }
    uint32_t u32() { return static_cast<uint32_t>(be(4)); }
    int32_t i32() { return static_cast<int32_t>(be(4)); }
    void skip(size_t k)
    {
        need(k);
        p += k;
    }
    std::string string()
    {
        size_t len = u8();
        if (len == 255)
            len = u32();
        need(len);
        std::string s(reinterpret_cast<char const*>(d + p), len);
        p += len;
        return s;
    }
    std::string cstring()
    {
        size_t e = p;
        while (e < n && d[e] != 0)
            ++e;
        if (e == n)
            fail("unterminated class name");
        std::string s(reinterpret_cast<char const*>(d + p), e - p);
        p = e + 1;
        return s;
    }
    //! [byte count] version; `end` = position after the object, or npos without a count
    uint16_t version(size_t* end)
    {
        size_t const start = p;
        uint32_t const bc = u32();
        if (bc & kByteCountMask)
        {
            *end = start + 4 + (bc & ~kByteCountMask);
            return u16();
        }
        p = start;
        *end = npos;
        return u16();
    }
    static constexpr size_t npos = static_cast<size_t>(-1);
};

//---------------------------------------------------------------------------//
// TFile / TKey
//---------------------------------------------------------------------------//
struct Key
{
    std::string cls, name, title;
    uint32_t keylen{0};
    Bytes data;  // object bytes, inflated
};

Bytes inflate_blocks(unsigned char const* src, size_t srclen, size_t objlen)
{
    Bytes out(objlen);
    size_t p = 0, o = 0;
    while (o < objlen)
    {
        if (p + 9 > srclen || src[p] != 'Z' || src[p + 1] != 'L')
            fail("unsupported compression (only zlib 'ZL' blocks are read)");
        size_t const c = src[p + 3] | (src[p + 4] << 8) | (size_t(src[p + 5]) << 16);
        if (p + 9 + c > srclen)
            fail("truncated compressed block");
        uLongf dst = static_cast<uLongf>(objlen - o);
        if (uncompress(out.data() + o, &dst, src + p + 9, static_cast<uLong>(c)) != Z_OK)
            fail("zlib inflate failed");
        o += dst;
        p += 9 + c;
    }
    return out;
}

std::vector<Key> read_keys(Bytes const& f)
{
    if (f.size() < 64 || std::memcmp(f.data(), "root", 4) != 0)
        fail("not a ROOT file");
    Buf h(f, 4);
    int32_t const version = h.i32();
    int32_t const begin = h.i32();
    int32_t const end = h.i32();
    if (version >= 1000000)
        fail("large-file (64-bit) layout is not supported");
    if (begin < 0 || end < begin || size_t(end) > f.size())
        fail("inconsistent header");
    std::vector<Key> keys;
    size_t pos = begin;
    while (pos < size_t(end))
    {
        Buf b(f, pos);
        int32_t const nbytes = b.i32();
        if (nbytes < 0)
        {
            pos += size_t(-int64_t(nbytes));  // free segment
            continue;
        }
        if (nbytes < 18 || pos + nbytes > f.size())
            fail("inconsistent key record");
        int16_t const ver = static_cast<int16_t>(b.u16());
        int32_t const objlen = b.i32();
        b.skip(4);  // date/time
        uint16_t const keylen = b.u16();
        b.skip(2);  // cycle
        b.skip(ver > 1000 ? 16 : 8);  // seek key, seek parent directory
        Key k;
        k.cls = b.string();
        k.name = b.string();
        k.title = b.string();
        k.keylen = keylen;
        if (keylen > size_t(nbytes) || objlen < 0)
            fail("inconsistent key record");
        unsigned char const* data = f.data() + pos + keylen;
        size_t const len = size_t(nbytes) - keylen;
        if (size_t(objlen) != len)
            k.data = inflate_blocks(data, len, size_t(objlen));
        else
            k.data.assign(data, data + len);
        keys.push_back(std::move(k));
        pos += size_t(nbytes);
    }
    return keys;
}

//---------------------------------------------------------------------------//
// TStreamerInfo list
//---------------------------------------------------------------------------//
struct Element
{
    std::string name, type_name;
    int type{0};
};
struct Info
{
    std::string name;
    std::vector<Element> elements;
};
//! What the object-wise reader hands back for the few TObject classes in the list
struct Obj
{
    enum Kind
    {
        none,
        info,
        array,
        element,
        other
    } kind{none};
    Info info_value;
    std::vector<Obj> items;
    Element element_value;
};

class InfoReader
{
  public:
    InfoReader(Bytes const& data, uint32_t keylen) : b_(data), keylen_(keylen) {}

    std::vector<Obj> read_tlist()
    {
        size_t end;
        b_.version(&end);
        read_tobject();
        b_.string();
        int32_t const n = b_.i32();
        std::vector<Obj> items;
        for (int32_t i = 0; i < n; ++i)
        {
            items.push_back(read_object_any());
            b_.skip(b_.u8());  // option string
        }
        return items;
    }

  private:
    Buf b_;
    uint32_t keylen_;
    std::map<uint32_t, std::string> classes_;  // tag position -> class name

    void read_tobject()
    {
        uint16_t const v = b_.u16();
        if (v & (kByteCountMask >> 16))
            b_.skip(4);
        b_.u32();  // fUniqueID
        uint32_t const bits = b_.u32();
        if (bits & (1u << 4))  // kIsReferenced
            b_.skip(2);
    }
    void read_tnamed(std::string* name)
    {
        size_t end;
        b_.version(&end);
        read_tobject();
        *name = b_.string();
        b_.string();  // title
    }
    Obj read_object_any()
    {
        size_t const start = b_.p;
        uint32_t bc = b_.u32();
        uint32_t tag;
        size_t end = Buf::npos, tagpos;
        if (!(bc & kByteCountMask) || bc == kNewClassTag)
        {
            tag = bc;
            tagpos = start;
        }
        else
        {
            end = start + 4 + (bc & ~kByteCountMask);
            tagpos = b_.p;
            tag = b_.u32();
        }
        if (tag == 0 || !(tag & kClassMask))
            return Obj{};  // null, or a reference to an object already read
        std::string cname;
        if (tag == kNewClassTag)
        {
            cname = b_.cstring();
            classes_[static_cast<uint32_t>(tagpos) + keylen_ + kMapOffset] = cname;
        }
        else
        {
            auto it = classes_.find(tag & ~kClassMask);
            if (it == classes_.end())
                fail("dangling class reference in the streamer-info list");
            cname = it->second;
        }
        Obj o = read_class(cname);
        if (end != Buf::npos && b_.p != end)
            fail("streamer-info record '" + cname + "' has an unexpected length");
        return o;
    }
    Obj read_class(std::string const& cname)
    {
        Obj o;
        size_t end;
        if (cname == "TStreamerInfo")
        {
            b_.version(&end);
            read_tnamed(&o.info_value.name);
            b_.u32();  // checksum
            b_.i32();  // class version
            Obj elements = read_object_any();
            for (Obj const& e : elements.items)
                if (e.kind == Obj::element)
                    o.info_value.elements.push_back(e.element_value);
            if (b_.p != end)
                fail("TStreamerInfo record has an unexpected length");
            o.kind = Obj::info;
        }
        else if (cname == "TObjArray")
        {
            b_.version(&end);
            read_tobject();
            b_.string();
            int32_t const n = b_.i32();
            b_.i32();  // lower bound
            for (int32_t i = 0; i < n; ++i)
                o.items.push_back(read_object_any());
            if (b_.p != end)
                fail("TObjArray record has an unexpected length");
            o.kind = Obj::array;
        }
        else if (cname == "TList")
        {
            o.items = read_tlist();
            o.kind = Obj::array;
        }
        else if (cname.rfind("TStreamer", 0) == 0)
        {
            b_.version(&end);
            size_t inner;
            if (cname == "TStreamerSTLstring")
                b_.version(&inner);  // nested TStreamerSTL header
            // TStreamerElement
            size_t el_end;
            uint16_t const v = b_.version(&el_end);
            read_tnamed(&o.element_value.name);
            o.element_value.type = b_.i32();
            b_.i32();  // size
            b_.i32();  // array length
            b_.i32();  // array dimension
            if (v == 1)
                b_.skip(4 * size_t(b_.i32()));
            else
                b_.skip(20);
            o.element_value.type_name = b_.string();
            if (end == Buf::npos)
                fail("streamer element without a byte count");
            b_.p = end;
            o.kind = Obj::element;
        }
        else if (cname == "TObjString")
        {
            b_.version(&end);
            if (end == Buf::npos)
                fail("TObjString without a byte count");
            b_.p = end;
            o.kind = Obj::other;
        }
        else
        {
            fail("class '" + cname + "' in the streamer-info list is not supported");
        }
        return o;
    }
};

//---------------------------------------------------------------------------//
// Split-branch decoding, driven by the streamer infos
//---------------------------------------------------------------------------//
struct Kind
{
    enum K
    {
        basic,
        string,
        stl,
        cls
    } k{basic};
    int code{3};       // basic: ROOT type code
    std::string name;  // stl: element type; cls: class name
};

std::string trim(std::string s)
{
    size_t a = s.find_first_not_of(' ');
    size_t b = s.find_last_not_of(' ');
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

//! "vector<pair<unsigned int,double> >" -> "vector", {"pair<unsigned int,double>"}
std::string split_template(std::string tname, std::vector<std::string>* args)
{
    tname = trim(tname);
    args->clear();
    size_t const lt = tname.find('<');
    if (lt == std::string::npos)
        return tname;
    size_t const gt = tname.rfind('>');
    std::string const inner = tname.substr(lt + 1, gt - lt - 1);
    int depth = 0;
    std::string cur;
    for (char ch : inner)
    {
        if (ch == '<')
            ++depth;
        else if (ch == '>')
            --depth;
        if (ch == ',' && depth == 0)
        {
            args->push_back(trim(cur));
            cur.clear();
        }
        else
            cur += ch;
    }
    args->push_back(trim(cur));
    return tname.substr(0, lt);
}

int basic_code_of(std::string const& t)
{
    static std::map<std::string, int> const names{{"int", 3},
                                                  {"unsigned int", 13},
                                                  {"double", 8},
                                                  {"float", 5},
                                                  {"bool", 18},
                                                  {"unsigned", 13},
                                                  {"long", 4},
                                                  {"unsigned long", 14},
                                                  {"short", 2},
                                                  {"char", 1}};
    auto it = names.find(t);
    return it == names.end() ? 0 : it->second;
}

bool is_basic_code(int c)
{
    switch (c)
    {
        case 1: case 2: case 3: case 4: case 5: case 8: case 11: case 12: case 13:
        case 14: case 16: case 17: case 18:
            return true;
        default:
            return false;
    }
}

class Decoder
{
  public:
    explicit Decoder(Bytes const& file)
    {
        std::vector<Key> keys = read_keys(file);
        bool found = false;
        for (Key const& k : keys)
        {
            if (k.cls == "TList" && k.name == "StreamerInfo" && !found)
            {
                found = true;
                for (Obj& o : InfoReader(k.data, k.keylen).read_tlist())
                    if (o.kind == Obj::info)
                        infos_[o.info_value.name] = std::move(o.info_value);
            }
        }
        if (!found)
            fail("no StreamerInfo record");
        for (Key& k : keys)
        {
            if (k.cls != "TBasket")
                continue;
            if (baskets_.count(k.name))
                fail("branch '" + k.name + "' has more than one basket (one entry expected)");
            baskets_[k.name] = std::move(k.data);
        }
        if (!infos_.count("celeritas::ImportData"))
            fail("no celeritas::ImportData streamer info: not a Celeritas physics export");
    }

    json import_data() { return branch_struct("", "celeritas::ImportData"); }

  private:
    std::map<std::string, Info> infos_;
    std::map<std::string, Bytes> baskets_;

    Info const& info(std::string const& cname) const
    {
        auto it = infos_.find(cname);
        if (it == infos_.end())
            fail("no streamer info for class '" + cname + "'");
        return it->second;
    }
    Bytes const& basket(std::string const& name) const
    {
        auto it = baskets_.find(name);
        if (it == baskets_.end())
            fail("missing branch '" + name + "'");
        return it->second;
    }
    bool has_sub_branches(std::string const& name) const
    {
        auto it = baskets_.lower_bound(name + ".");
        return it != baskets_.end() && it->first.compare(0, name.size() + 1, name + ".") == 0;
    }

    static json basic(Buf& b, int code)
    {
        switch (code)
        {
            case 1: return static_cast<int8_t>(b.u8());
            case 2: return static_cast<int16_t>(b.u16());
            case 3: return b.i32();
            case 4: case 16: return static_cast<int64_t>(b.be(8));
            case 5:
            {
                uint32_t u = b.u32();
                float f;
                std::memcpy(&f, &u, 4);
                return static_cast<double>(f);
            }
            case 8:
            {
                uint64_t u = b.be(8);
                double v;
                std::memcpy(&v, &u, 8);
                return v;
            }
            case 11: return b.u8();
            case 12: return b.u16();
            case 13: return b.u32();
            case 14: case 17: return b.be(8);
            case 18: return b.u8() != 0;
        }
        fail("unsupported basic type code " + std::to_string(code));
    }

    Kind value_kind(std::string tname) const
    {
        tname = trim(tname);
        Kind k;
        if (int c = basic_code_of(tname))
        {
            k.k = Kind::basic;
            k.code = c;
            return k;
        }
        if (tname == "string" || tname == "std::string")
        {
            k.k = Kind::string;
            return k;
        }
        std::vector<std::string> args;
        std::string const outer = split_template(tname, &args);
        if (outer == "vector" && !args.empty())
        {
            k.k = Kind::stl;
            k.name = args[0];
            return k;
        }
        if (outer == "map" && args.size() >= 2)
        {
            k.k = Kind::stl;
            k.name = "pair<" + args[0] + "," + args[1] + ">";
            return k;
        }
        if (infos_.count(tname))
        {
            k.k = Kind::cls;
            k.name = tname;
            return k;
        }
        k.k = Kind::basic;  // enums are stored as int
        k.code = 3;
        return k;
    }
    Kind elem_kind(Element const& e) const
    {
        if (is_basic_code(e.type))
        {
            Kind k;
            k.k = Kind::basic;
            k.code = e.type;
            return k;
        }
        return value_kind(e.type_name);
    }

    //// object-wise ////
    json read_object(Buf& b, std::string const& cname)
    {
        size_t end;
        uint16_t const v = b.version(&end);
        if (v == 0)
            b.u32();  // checksum of a class without ClassDef
        json out = json::object();
        for (Element const& e : info(cname).elements)
            out[e.name] = read_value(b, elem_kind(e));
        if (end != Buf::npos && b.p != end)
            fail("object of class '" + cname + "' has an unexpected length");
        return out;
    }
    json read_value(Buf& b, Kind const& kind)
    {
        switch (kind.k)
        {
            case Kind::basic: return basic(b, kind.code);
            case Kind::string: return b.string();
            case Kind::cls: return read_object(b, kind.name);
            case Kind::stl:
            {
                size_t end;
                bool const memberwise = read_stl_header(b, &end);
                json val = read_stl_body(b, kind.name, memberwise);
                if (end != Buf::npos && b.p != end)
                    fail("collection of '" + kind.name + "' has an unexpected length");
                return val;
            }
        }
        fail("unreachable");
    }

    //// STL collections ////
    static bool read_stl_header(Buf& b, size_t* end)
    {
        uint16_t const v = b.version(end);
        bool const memberwise = (v & kMemberWise) != 0;
        if (memberwise)
        {
            if (b.u16() == 0)
                b.u32();
        }
        return memberwise;
    }
    //! One collection instance: count + contents (header already consumed)
    json read_stl_body(Buf& b, std::string const& etype, bool memberwise)
    {
        uint32_t const n = b.u32();
        json out = json::array();
        if (n == 0)
            return out;
        Kind const ek = value_kind(etype);
        if (ek.k == Kind::cls && memberwise)
            return read_memberwise(b, ek.name, n);
        for (uint32_t i = 0; i < n; ++i)
        {
            if (ek.k == Kind::stl)
                out.push_back(read_stl_body(b, ek.name, false));  // e.g. vector<vector<double>>
            else
                out.push_back(read_value(b, ek));
        }
        return out;
    }
    //! The same member of n consecutive objects (member-wise layout)
    std::vector<json> read_column(Buf& b, Kind const& kind, uint32_t n)
    {
        std::vector<json> out;
        out.reserve(n);
        switch (kind.k)
        {
            case Kind::basic:
                for (uint32_t i = 0; i < n; ++i)
                    out.push_back(basic(b, kind.code));
                break;
            case Kind::string:
                for (uint32_t i = 0; i < n; ++i)
                    out.push_back(b.string());
                break;
            case Kind::cls:
                for (uint32_t i = 0; i < n; ++i)
                    out.push_back(read_object(b, kind.name));
                break;
            case Kind::stl:
            {
                size_t end;
                bool const memberwise = read_stl_header(b, &end);
                for (uint32_t i = 0; i < n; ++i)
                    out.push_back(read_stl_body(b, kind.name, memberwise));
                if (end != Buf::npos && b.p != end)
                    fail("column of '" + kind.name + "' collections has an unexpected length");
                break;
            }
        }
        return out;
    }
    static json zip_columns(std::vector<std::pair<std::string, std::vector<json>>>& cols, uint32_t n)
    {
        json out = json::array();
        for (uint32_t i = 0; i < n; ++i)
        {
            json row = json::object();
            for (auto& c : cols)
                row[c.first] = std::move(c.second[i]);
            out.push_back(std::move(row));
        }
        return out;
    }
    json read_memberwise(Buf& b, std::string const& cname, uint32_t n)
    {
        std::vector<std::pair<std::string, std::vector<json>>> cols;
        for (Element const& e : info(cname).elements)
            cols.emplace_back(e.name, read_column(b, elem_kind(e), n));
        return zip_columns(cols, n);
    }

    //// branches ////
    //! Top-level split branch `name` holding one member of n objects
    std::vector<json> branch_column(std::string const& name, Kind const& kind, uint32_t n)
    {
        if (kind.k == Kind::cls && !baskets_.count(name))
        {
            // nested struct: split further into sub-branches
            std::vector<std::pair<std::string, std::vector<json>>> cols;
            for (Element const& e : info(kind.name).elements)
                cols.emplace_back(e.name, branch_column(name + "." + e.name, elem_kind(e), n));
            json rows = zip_columns(cols, n);
            return std::vector<json>(rows.begin(), rows.end());
        }
        Buf b(basket(name));
        if (kind.k == Kind::string)
        {
            size_t end;
            b.version(&end);
            std::vector<json> out;
            for (uint32_t i = 0; i < n; ++i)
                out.push_back(b.string());
            return out;
        }
        return read_column(b, kind, n);
    }
    json branch_collection(std::string const& name, std::string const& etype)
    {
        uint32_t const n = Buf(basket(name)).u32();
        Kind const ek = value_kind(etype);
        if (ek.k != Kind::cls)
            fail("split branch '" + name + "' does not hold structs");
        std::vector<std::pair<std::string, std::vector<json>>> cols;
        for (Element const& e : info(ek.name).elements)
            cols.emplace_back(e.name, branch_column(name + "." + e.name, elem_kind(e), n));
        return zip_columns(cols, n);
    }
    json branch_struct(std::string const& prefix, std::string const& cname)
    {
        json out = json::object();
        for (Element const& e : info(cname).elements)
        {
            std::string const name = prefix.empty() ? e.name : prefix + "." + e.name;
            if (name.rfind("optical_", 0) == 0)
                continue;  // optical physics is outside the EM track loop
            Kind const kind = elem_kind(e);
            if (kind.k == Kind::stl && baskets_.count(name) && has_sub_branches(name))
            {
                json items = branch_collection(name, kind.name);
                std::vector<std::string> args;
                if (split_template(e.type_name, &args) == "map")
                {
                    // keyed by the (integer) key as a string, like a JSON object must be
                    json m = json::object();
                    for (auto& it : items)
                    {
                        json const& key = it.at("first");
                        m[key.is_string() ? key.get<std::string>() : key.dump()]
                            = std::move(it.at("second"));
                    }
                    items = std::move(m);
                }
                out[e.name] = std::move(items);
            }
            else if (kind.k == Kind::cls)
            {
                if (baskets_.count(name) && !has_sub_branches(name))
                {
                    Buf b(basket(name));
                    out[e.name] = read_object(b, kind.name);
                }
                else
                    out[e.name] = branch_struct(name, kind.name);
            }
            else if (kind.k == Kind::stl)
            {
                Buf b(basket(name));  // unsplit collection (of basic type) in one basket
                out[e.name] = read_value(b, kind);
            }
            else if (kind.k == Kind::string)
            {
                Buf b(basket(name));
                size_t end;
                b.version(&end);
                out[e.name] = b.string();
            }
            else
            {
                Buf b(basket(name));
                out[e.name] = basic(b, kind.code);
            }
        }
        return out;
    }
};

Bytes read_file(std::string const& path)
{
    std::ifstream in(path, std::ios::binary);
    if (!in)
        fail("cannot open '" + path + "'");
    Bytes data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    return data;
}
}  // namespace

//---------------------------------------------------------------------------//
std::string import_root_to_json(std::string const& path)
{
    Bytes const file = read_file(path);
    return Decoder(file).import_data().dump();
}

}  // namespace b200
