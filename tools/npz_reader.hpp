// npz_reader.hpp -- the .npz container (zip of .npy members) for the C++ host side: what `npyz::npz::NpzArchive` does for
// the reference's NpzReader (src/io/npz.rs:29-56,226-300).  Stored members always; deflated members (np.savez_compressed)
// when built with zlib (-DWS_HAVE_ZLIB, decided by web-splat_b200/build.py).  Host I/O only: the arrays go to the GPU
// through ws_pointcloud_create_from_c3dgs.
#ifndef WS_NPZ_READER_HPP
#define WS_NPZ_READER_HPP

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef WS_HAVE_ZLIB
#include <zlib.h>
#endif

namespace npz {

struct Array {
    std::string descr;              // numpy dtype string, e.g. "<f2", "|i1", "<i4", "|b1"
    std::vector<size_t> shape;
    bool fortran_order = false;
    std::vector<uint8_t> data;
    size_t count() const { size_t n = 1; for (size_t s : shape) n *= s; return n; }
    size_t itemsize() const { return descr.size() >= 3 ? (size_t)std::stoul(descr.substr(2)) : 0; }
    char kind() const { return descr.size() >= 2 ? descr[1] : '?'; }
    /// element i as a double, for the scalar members (`*_scale`, `*_zero_point`, `kernel_size`, `mip_splatting`)
    double as_double(size_t i = 0) const
    {
        const uint8_t *p = data.data() + i * itemsize();
        switch (kind()) {
        case 'f':
            if (itemsize() == 4) { float v; std::memcpy(&v, p, 4); return v; }
            if (itemsize() == 8) { double v; std::memcpy(&v, p, 8); return v; }
            break;
        case 'i':
            if (itemsize() == 1) return (int8_t)p[0];
            if (itemsize() == 2) { int16_t v; std::memcpy(&v, p, 2); return v; }
            if (itemsize() == 4) { int32_t v; std::memcpy(&v, p, 4); return v; }
            if (itemsize() == 8) { int64_t v; std::memcpy(&v, p, 8); return (double)v; }
            break;
        case 'u':
            if (itemsize() == 1) return p[0];
            if (itemsize() == 4) { uint32_t v; std::memcpy(&v, p, 4); return v; }
            if (itemsize() == 8) { uint64_t v; std::memcpy(&v, p, 8); return (double)v; }
            break;
        case 'b': return p[0] != 0;
        default: break;
        }
        throw std::runtime_error("npz: unsupported scalar dtype '" + descr + "'");
    }
};

namespace detail {
inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

/// the python-dict header of a .npy member: {'descr': '<f2', 'fortran_order': False, 'shape': (N, 3), }
inline Array parse_npy(const std::vector<uint8_t> &raw, const std::string &name)
{
    if (raw.size() < 10 || std::memcmp(raw.data(), "\x93NUMPY", 6) != 0) throw std::runtime_error("npz: member '" + name + "' is not a .npy array");
    const int major = raw[6];
    size_t hlen, hoff;
    if (major == 1) { hlen = rd16(raw.data() + 8); hoff = 10; }
    else { if (raw.size() < 12) throw std::runtime_error("npz: truncated .npy header"); hlen = rd32(raw.data() + 8); hoff = 12; }
    if (hlen > raw.size() - hoff) throw std::runtime_error("npz: truncated .npy header in '" + name + "'");
    const std::string h(reinterpret_cast<const char *>(raw.data() + hoff), hlen);
    Array a;
    auto value_after = [&](const char *key) -> size_t {
        const size_t k = h.find(key);
        if (k == std::string::npos) throw std::runtime_error("npz: .npy header of '" + name + "' lacks " + key);
        const size_t c = h.find(':', k);
        if (c == std::string::npos) throw std::runtime_error("npz: malformed .npy header in '" + name + "'");
        return c + 1;
    };
    {
        size_t p = h.find('\'', value_after("'descr'"));
        const size_t e = h.find('\'', p + 1);
        if (p == std::string::npos || e == std::string::npos) throw std::runtime_error("npz: structured dtypes are not supported ('" + name + "')");
        a.descr = h.substr(p + 1, e - p - 1);
    }
    {
        const size_t f = h.find_first_not_of(' ', value_after("'fortran_order'"));
        a.fortran_order = f != std::string::npos && h.compare(f, 4, "True") == 0;
    }
    {
        const size_t p = h.find('(', value_after("'shape'")), e = h.find(')', p);
        if (p == std::string::npos || e == std::string::npos) throw std::runtime_error("npz: bad shape in '" + name + "'");
        size_t i = p + 1;
        while (i < e) {
            while (i < e && (h[i] == ' ' || h[i] == ',')) i++;
            if (i >= e) break;
            size_t j = i; while (j < e && h[j] >= '0' && h[j] <= '9') j++;
            if (j == i) throw std::runtime_error("npz: bad shape in '" + name + "'");
            a.shape.push_back((size_t)std::stoull(h.substr(i, j - i)));
            i = j;
        }
    }
    if (a.descr.size() < 3 || (a.descr[0] == '>' && a.itemsize() > 1)) throw std::runtime_error("npz: big-endian / unsupported dtype '" + a.descr + "' in '" + name + "'");
    // overflow-safe element count: a shape like (2^63, 2) must not wrap to a small byte count
    const size_t avail = raw.size() - hoff - hlen, isz = a.itemsize();
    if (isz == 0) throw std::runtime_error("npz: unsupported dtype '" + a.descr + "' in '" + name + "'");
    size_t cnt = 1;
    for (size_t d : a.shape) if (d == 0) cnt = 0;                    // an empty array holds no data whatever the other extents
    if (cnt) for (size_t d : a.shape) {
        if (cnt > avail / d) throw std::runtime_error("npz: member '" + name + "' is shorter than its shape says");
        cnt *= d;
    }
    if (cnt > avail / isz) throw std::runtime_error("npz: member '" + name + "' is shorter than its shape says");
    const size_t need = cnt * isz;
    a.data.assign(raw.begin() + hoff + hlen, raw.begin() + hoff + hlen + need);
    return a;
}

inline std::vector<uint8_t> inflate_raw(const uint8_t *src, size_t n, size_t out_size, const std::string &name)
{
#ifdef WS_HAVE_ZLIB
    std::vector<uint8_t> out(out_size);
    z_stream zs; std::memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) throw std::runtime_error("npz: inflateInit2 failed");
    size_t in_done = 0, out_done = 0;                     // members above 4 GB: zlib's counters are 32-bit, feed it in chunks
    int rc = Z_OK;
    do {
        if (zs.avail_in == 0 && in_done < n) {
            const size_t c = std::min<size_t>(n - in_done, (size_t)1 << 30);
            zs.next_in = const_cast<Bytef *>(src + in_done); zs.avail_in = (uInt)c; in_done += c;
        }
        if (zs.avail_out == 0 && out_done < out_size) {
            const size_t c = std::min<size_t>(out_size - out_done, (size_t)1 << 30);
            zs.next_out = out.data() + out_done; zs.avail_out = (uInt)c; out_done += c;
        }
        rc = inflate(&zs, Z_NO_FLUSH);
    } while (rc == Z_OK && (zs.avail_in > 0 || in_done < n));
    const size_t produced = out_done - zs.avail_out;
    inflateEnd(&zs);
    if (rc != Z_STREAM_END || produced != out_size) throw std::runtime_error("npz: inflate failed on '" + name + "'");
    return out;
#else
    (void)src; (void)n; (void)out_size;
    throw std::runtime_error("npz: member '" + name + "' is deflated and this build has no zlib (np.savez_compressed); re-save with np.savez");
#endif
}
}  // namespace detail

/// All members of a .npz image, keyed by name without the ".npy" suffix.
inline std::map<std::string, Array> read(const uint8_t *bytes, size_t len)
{
    using namespace detail;
    if (len < 22) throw std::runtime_error("npz: file too short");
    // end-of-central-directory record: scan back over a possible comment
    size_t eocd = std::string::npos;
    for (size_t back = 22; back <= len && back <= 22 + 65535; back++)
        if (rd32(bytes + len - back) == 0x06054b50u) { eocd = len - back; break; }
    if (eocd == std::string::npos) throw std::runtime_error("npz: not a zip archive (no end-of-central-directory record)");
    uint64_t entries = rd16(bytes + eocd + 10), cd_off = rd32(bytes + eocd + 16);
    if (entries == 0xffffu || cd_off == 0xffffffffu) {               // zip64: locator 20 B in front of the EOCD
        if (eocd < 20 || rd32(bytes + eocd - 20) != 0x07064b50u) throw std::runtime_error("npz: zip64 locator missing");
        const uint64_t e64 = rd64(bytes + eocd - 20 + 8);
        if (e64 > len || len - e64 < 56 || rd32(bytes + e64) != 0x06064b50u) throw std::runtime_error("npz: bad zip64 end-of-central-directory record");
        entries = rd64(bytes + e64 + 32); cd_off = rd64(bytes + e64 + 48);
    }
    // every offset / size below comes from the file: compare against what is LEFT (len - x), never x + size > len (wraps)
    if (cd_off > len) throw std::runtime_error("npz: central directory offset beyond the file");
    if (entries > (len - cd_off) / 46) throw std::runtime_error("npz: more central directory entries than the file can hold");
    std::map<std::string, Array> out;
    uint64_t p = cd_off;
    for (uint64_t i = 0; i < entries; i++) {
        if (p > len || len - p < 46 || rd32(bytes + p) != 0x02014b50u) throw std::runtime_error("npz: bad central directory entry");
        const uint16_t method = rd16(bytes + p + 10), nlen = rd16(bytes + p + 28), xlen = rd16(bytes + p + 30), clen = rd16(bytes + p + 32);
        uint64_t csize = rd32(bytes + p + 20), usize = rd32(bytes + p + 24), lho = rd32(bytes + p + 42);
        if (len - p - 46 < (uint64_t)nlen + xlen + clen) throw std::runtime_error("npz: central directory entry runs past the end of the file");
        std::string name(reinterpret_cast<const char *>(bytes + p + 46), nlen);
        // zip64 extended information (header id 1): the 0xffffffff fields, in this order, each inside the sub-field's size
        for (uint64_t x = p + 46 + nlen, xe = x + xlen; x + 4 <= xe; ) {
            const uint16_t id = rd16(bytes + x), sz = rd16(bytes + x + 2);
            if (x + 4 + sz > xe) throw std::runtime_error("npz: extra field of '" + name + "' runs past its record");
            if (id == 1) {
                uint64_t q = x + 4;
                const uint64_t qe = q + sz;
                auto take64 = [&](uint64_t &v) { if (qe - q < 8) throw std::runtime_error("npz: short zip64 extra field in '" + name + "'"); v = rd64(bytes + q); q += 8; };
                if (usize == 0xffffffffu) take64(usize);
                if (csize == 0xffffffffu) take64(csize);
                if (lho == 0xffffffffu) take64(lho);
            }
            x += 4 + sz;
        }
        if (lho > len || len - lho < 30 || rd32(bytes + lho) != 0x04034b50u) throw std::runtime_error("npz: bad local header for '" + name + "'");
        const uint64_t data = lho + 30 + rd16(bytes + lho + 26) + rd16(bytes + lho + 28);
        if (data > len || csize > len - data) throw std::runtime_error("npz: member '" + name + "' is truncated");
        // a stored member is as long as its data; a deflated one cannot expand beyond deflate's ~1032:1 bound
        if ((method == 0 && usize != csize) || (method == 8 && usize / 1040 > csize + 1))
            throw std::runtime_error("npz: implausible uncompressed size for '" + name + "'");
        std::vector<uint8_t> raw;
        if (method == 0) raw.assign(bytes + data, bytes + data + csize);
        else if (method == 8) raw = inflate_raw(bytes + data, csize, usize, name);
        else throw std::runtime_error("npz: unsupported compression method in '" + name + "'");
        if (name.size() > 4 && name.compare(name.size() - 4, 4, ".npy") == 0) name.resize(name.size() - 4);
        out.emplace(name, parse_npy(raw, name));
        p += 46 + (uint64_t)nlen + xlen + clen;
    }
    return out;
}

}  // namespace npz

#endif  // WS_NPZ_READER_HPP
