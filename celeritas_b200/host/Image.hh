//---------------------------------------------------------------------------//
// Problem image: a flat container of named, typed arrays.
//
// The image is what crosses the drop-in boundary on the *params* side: every
// array in it is a plain column (no pointers, no nested structs) that the
// loader re-lays out for HBM.  It is produced either by the reference-side
// adapter (oracle/ref_harness/export_image.cc walks CoreParams::host_ref(),
// /root/reference/src/celeritas/global/CoreParams.hh:155-172) or by any other
// front end that can fill the same columns; see INTEGRATION.md.
//
// File layout (little endian):
//   char[8]  magic  "B2IMG\0\0\1"
//   u32      number of entries
//   per entry: u32 name length, name bytes, u32 dtype, u64 element count,
//              raw data padded to a multiple of 8 bytes
//---------------------------------------------------------------------------//
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200
{
enum class DType : uint32_t
{
    u8 = 0,
    u32 = 1,
    i32 = 2,
    f32 = 3,
    f64 = 4,
    u64 = 5
};

inline size_t dtype_size(DType t)
{
    switch (t)
    {
        case DType::u8: return 1;
        case DType::u32: return 4;
        case DType::i32: return 4;
        case DType::f32: return 4;
        case DType::f64: return 8;
        case DType::u64: return 8;
    }
    throw std::runtime_error("bad dtype");
}

template<class T> struct DTypeOf;
template<> struct DTypeOf<uint8_t> { static constexpr DType value = DType::u8; };
template<> struct DTypeOf<uint32_t> { static constexpr DType value = DType::u32; };
template<> struct DTypeOf<int32_t> { static constexpr DType value = DType::i32; };
template<> struct DTypeOf<float> { static constexpr DType value = DType::f32; };
template<> struct DTypeOf<double> { static constexpr DType value = DType::f64; };
template<> struct DTypeOf<uint64_t> { static constexpr DType value = DType::u64; };

struct ImageEntry
{
    DType dtype{DType::u8};
    uint64_t count{0};
    std::vector<unsigned char> bytes;
};

class Image
{
  public:
    template<class T>
    void put(std::string const& name, std::vector<T> const& v)
    {
        ImageEntry e;
        e.dtype = DTypeOf<T>::value;
        e.count = v.size();
        e.bytes.resize(v.size() * sizeof(T));
        if (!v.empty())
            std::memcpy(e.bytes.data(), v.data(), e.bytes.size());
        entries_[name] = std::move(e);
    }
    template<class T>
    void put_scalar(std::string const& name, T v)
    {
        this->put(name, std::vector<T>{v});
    }
    void put_string(std::string const& name, std::string const& s)
    {
        this->put(name, std::vector<uint8_t>(s.begin(), s.end()));
    }

    //! Insert or replace a whole entry (e.g. a column taken from another image)
    void put_entry(std::string const& name, ImageEntry const& entry) { entries_[name] = entry; }

    bool has(std::string const& name) const { return entries_.count(name) != 0; }

    template<class T>
    std::vector<T> get(std::string const& name) const
    {
        auto it = entries_.find(name);
        if (it == entries_.end())
            throw std::runtime_error("image has no entry '" + name + "'");
        if (it->second.dtype != DTypeOf<T>::value)
            throw std::runtime_error("image entry '" + name + "' has wrong dtype");
        std::vector<T> v(it->second.count);
        if (!v.empty())
            std::memcpy(v.data(), it->second.bytes.data(), it->second.bytes.size());
        return v;
    }
    template<class T>
    T get_scalar(std::string const& name) const
    {
        auto v = this->get<T>(name);
        if (v.size() != 1)
            throw std::runtime_error("image entry '" + name + "' is not a scalar");
        return v[0];
    }
    std::string get_string(std::string const& name) const
    {
        auto v = this->get<uint8_t>(name);
        return std::string(v.begin(), v.end());
    }

    std::map<std::string, ImageEntry> const& entries() const { return entries_; }

    //! The image as one contiguous byte string (the file layout above)
    std::vector<unsigned char> serialize() const
    {
        std::vector<unsigned char> out;
        auto wr = [&out](void const* src, size_t n) {
            auto const* b = static_cast<unsigned char const*>(src);
            out.insert(out.end(), b, b + n);
        };
        wr(magic(), 8);
        uint32_t n = entries_.size();
        wr(&n, 4);
        for (auto const& kv : entries_)
        {
            uint32_t len = kv.first.size();
            wr(&len, 4);
            wr(kv.first.data(), len);
            uint32_t dt = static_cast<uint32_t>(kv.second.dtype);
            wr(&dt, 4);
            wr(&kv.second.count, 8);
            wr(kv.second.bytes.data(), kv.second.bytes.size());
            static char const zeros[8] = {0};
            wr(zeros, (8 - kv.second.bytes.size() % 8) % 8);
        }
        return out;
    }

    void write(std::string const& path) const
    {
        FILE* f = std::fopen(path.c_str(), "wb");
        if (!f)
            throw std::runtime_error("cannot open '" + path + "' for writing");
        std::vector<unsigned char> const bytes = this->serialize();
        size_t const written = std::fwrite(bytes.data(), 1, bytes.size(), f);
        std::fclose(f);
        if (written != bytes.size())
            throw std::runtime_error("short write to '" + path + "'");
    }

    //! Parse an image held in memory (`what` names it in error messages)
    static Image parse(void const* data, size_t size, std::string const& what = "image in memory")
    {
        auto const* cur = static_cast<unsigned char const*>(data);
        auto const* const end = cur + size;
        auto rd = [&](void* dst, size_t n) {
            if (n > static_cast<size_t>(end - cur))
                throw std::runtime_error("truncated image '" + what + "'");
            if (n)
                std::memcpy(dst, cur, n);
            cur += n;
        };
        if (!data)
            throw std::runtime_error("null image");
        char m[8];
        rd(m, 8);
        if (std::memcmp(m, magic(), 8) != 0)
            throw std::runtime_error("'" + what + "' is not a B2IMG v1 file");
        uint32_t n;
        rd(&n, 4);
        Image img;
        for (uint32_t i = 0; i < n; ++i)
        {
            uint32_t len;
            rd(&len, 4);
            if (len > static_cast<size_t>(end - cur))
                throw std::runtime_error("truncated image '" + what + "'");
            std::string name(len, '\0');
            rd(&name[0], len);
            uint32_t dt;
            rd(&dt, 4);
            ImageEntry e;
            e.dtype = static_cast<DType>(dt);
            rd(&e.count, 8);
            size_t const esize = dtype_size(e.dtype);
            if (e.count > static_cast<size_t>(end - cur) / esize)
                throw std::runtime_error("truncated image '" + what + "'");
            e.bytes.resize(e.count * esize);
            rd(e.bytes.data(), e.bytes.size());
            char pad[8];
            rd(pad, (8 - e.bytes.size() % 8) % 8);
            img.entries_[name] = std::move(e);
        }
        return img;
    }

    static Image read(std::string const& path)
    {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f)
            throw std::runtime_error("cannot open image '" + path + "'");
        std::vector<unsigned char> bytes;
        unsigned char buf[1 << 16];
        size_t got;
        while ((got = std::fread(buf, 1, sizeof(buf), f)) > 0)
            bytes.insert(bytes.end(), buf, buf + got);
        std::fclose(f);
        return parse(bytes.data(), bytes.size(), path);
    }

  private:
    static char const* magic() { return "B2IMG\0\0\1"; }
    std::map<std::string, ImageEntry> entries_;
};
}  // namespace b200
