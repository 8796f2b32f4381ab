// triplet_match/pcd_io.hpp — PCD (Point Cloud Data, v0.7) reader / writer for the 48-byte
// PointSurfel record, so the CLI works without PCL (the reference loads clouds with
// pcl::io::loadPCDFile, apps/triplet_match.cpp:14-15,33-34 and include/impl/pointcloud.hpp:60-64).
// Fields are matched by name: x y z | normal_x normal_y normal_z | rgba (or rgb) | radius
// confidence curvature (the tangent overlays the last three, include/common:62-70);
// tangent_x/y/z are accepted as aliases.  Missing fields stay zero.  DATA ascii, binary and
// binary_compressed (PCL's layout: u32 compressed size, u32 raw size, LZF stream of the records
// transposed field by field) are read; the writer emits ascii or binary.  SIZE 4 / 8 floats,
// 1 / 2 / 4 byte integers.
#ifndef TRIPLET_MATCH_PCD_IO_HPP_
#define TRIPLET_MATCH_PCD_IO_HPP_

#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "compat.hpp"

namespace triplet_match {
namespace pcd {

struct field_t {
    std::string name;
    int size = 4;
    char type = 'F';
    int count = 1;
    int offset = 0;  // byte offset inside a binary record
};

inline int surfel_slot(const std::string& n) {  // float index inside the 12-float record, -1 = ignore
    static const char* names[] = {"x", "y", "z", nullptr, "normal_x", "normal_y", "normal_z", nullptr,
                                  "rgba", "radius", "confidence", "curvature"};
    for (int i = 0; i < 12; ++i)
        if (names[i] && n == names[i]) return i;
    if (n == "rgb") return 8;
    if (n == "tangent_x") return 9;
    if (n == "tangent_y") return 10;
    if (n == "tangent_z") return 11;
    return -1;
}

inline double read_scalar(const char* p, const field_t& f) {
    switch (f.type) {
        case 'F':
            if (f.size == 4) { float v; std::memcpy(&v, p, 4); return v; }
            if (f.size == 8) { double v; std::memcpy(&v, p, 8); return v; }
            break;
        case 'U':
            if (f.size == 1) { uint8_t v; std::memcpy(&v, p, 1); return v; }
            if (f.size == 2) { uint16_t v; std::memcpy(&v, p, 2); return v; }
            if (f.size == 4) { uint32_t v; std::memcpy(&v, p, 4); return v; }
            break;
        case 'I':
            if (f.size == 1) { int8_t v; std::memcpy(&v, p, 1); return v; }
            if (f.size == 2) { int16_t v; std::memcpy(&v, p, 2); return v; }
            if (f.size == 4) { int32_t v; std::memcpy(&v, p, 4); return v; }
            break;
    }
    throw std::runtime_error("pcd: unsupported field type/size for '" + f.name + "'");
}

// LZF (liblzf) stream -> exactly out.size() bytes.  Control byte c: c < 32 copies c + 1 literal bytes;
// otherwise a back-reference of length (c >> 5) + 2 (length field 7: + the next byte) at distance
// ((c & 31) << 8 | next byte) + 1, which may overlap its own output.
inline bool lzf_decompress(const unsigned char* in, size_t n_in, std::vector<char>& out) {
    size_t ip = 0, op = 0;
    const size_t n_out = out.size();
    while (ip < n_in) {
        const unsigned c = in[ip++];
        if (c < 32u) {
            const size_t run = c + 1u;
            if (ip + run > n_in || op + run > n_out) return false;
            std::memcpy(&out[op], in + ip, run);
            ip += run;
            op += run;
        } else {
            size_t len = c >> 5;
            if (len == 7u) {
                if (ip >= n_in) return false;
                len += in[ip++];
            }
            if (ip >= n_in) return false;
            const size_t dist = (static_cast<size_t>(c & 31u) << 8 | in[ip++]) + 1u;
            len += 2u;
            if (dist > op || op + len > n_out) return false;
            for (size_t k = 0; k < len; ++k, ++op) out[op] = out[op - dist];
        }
    }
    return op == n_out;
}

inline void load(const std::string& filename, std::vector<pcl::PointSurfel>& out) {
    std::ifstream in(filename, std::ios::binary);
    if (!in) throw std::runtime_error("pcd: cannot open '" + filename + "'");
    std::vector<field_t> fields;
    uint64_t width = 0, height = 1, points = 0;
    bool have_points = false;
    std::string data_kind, line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ls(line);
        std::string key;
        ls >> key;
        auto fill = [&](auto setter) {
            std::string tok;
            size_t i = 0;
            while (ls >> tok) {
                if (fields.size() <= i) fields.resize(i + 1);
                setter(fields[i], tok);
                ++i;
            }
        };
        if (key == "FIELDS") fill([](field_t& f, const std::string& t) { f.name = t; });
        else if (key == "SIZE") fill([](field_t& f, const std::string& t) { f.size = static_cast<int>(std::strtol(t.c_str(), nullptr, 10)); });
        else if (key == "TYPE") fill([](field_t& f, const std::string& t) { f.type = t.empty() ? '?' : t[0]; });
        else if (key == "COUNT") fill([](field_t& f, const std::string& t) { f.count = static_cast<int>(std::strtol(t.c_str(), nullptr, 10)); });
        else if (key == "WIDTH") ls >> width;
        else if (key == "HEIGHT") ls >> height;
        else if (key == "POINTS") { ls >> points; have_points = true; }
        else if (key == "DATA") { ls >> data_kind; break; }
        // VERSION, VIEWPOINT: ignored
    }
    if (fields.empty()) throw std::runtime_error("pcd: no FIELDS line in '" + filename + "'");
    if (!have_points) points = width * height;
    // the header comes from an untrusted file: validate everything a buffer size or an offset is derived from
    int64_t rec64 = 0;
    for (auto& f : fields) {
        if (!(f.size == 1 || f.size == 2 || f.size == 4 || f.size == 8) || f.count < 1 || f.count > (1 << 20))
            throw std::runtime_error("pcd: bad SIZE / COUNT in '" + filename + "'");
        if (!(f.type == 'F' || f.type == 'I' || f.type == 'U'))
            throw std::runtime_error("pcd: bad TYPE in '" + filename + "'");
        f.offset = static_cast<int>(rec64);
        rec64 += static_cast<int64_t>(f.size) * f.count;
        if (rec64 > (1 << 24)) throw std::runtime_error("pcd: record too large in '" + filename + "'");
    }
    const int rec = static_cast<int>(rec64);
    {  // POINTS cannot exceed what the rest of the file can hold: >= 2 bytes per ascii point, a record per binary
       // point; an LZF token of 3 bytes expands to at most 264 bytes
        const std::streampos here = in.tellg();
        in.seekg(0, std::ios::end);
        const std::streampos end = in.tellg();
        in.seekg(here);
        const uint64_t remaining = (here >= 0 && end >= here) ? static_cast<uint64_t>(end - here) : 0;
        const uint64_t need = data_kind == "binary" ? points * static_cast<uint64_t>(rec)
                              : data_kind == "ascii" ? points * 2ull
                                                     : (points * static_cast<uint64_t>(rec)) / 128ull;
        if (points > (1ull << 40) || need > remaining)
            throw std::runtime_error("pcd: POINTS exceeds the file size in '" + filename + "'");
    }
    out.assign(points, pcl::PointSurfel());
    auto store = [](pcl::PointSurfel& p, int slot, const field_t& f, double v, const char* raw) {
        float* rec12 = reinterpret_cast<float*>(&p);
        if (slot == 8) {  // rgba / rgb keep their bit pattern
            if (raw && f.size == 4) std::memcpy(&p.rgba, raw, 4);
            else p.rgba = static_cast<uint32_t>(v);
        } else {
            rec12[slot] = static_cast<float>(v);
        }
    };
    if (data_kind == "ascii") {
        for (uint64_t i = 0; i < points; ++i) {
            if (!std::getline(in, line)) throw std::runtime_error("pcd: truncated ascii data in '" + filename + "'");
            std::istringstream ls(line);
            for (const auto& f : fields) {
                const int slot = surfel_slot(f.name);
                for (int c = 0; c < f.count; ++c) {
                    std::string tok;
                    if (!(ls >> tok)) throw std::runtime_error("pcd: short ascii record in '" + filename + "'");
                    if (slot < 0 || c > 0) continue;
                    // strtof / strtod: a packed rgb with alpha 0 is a denormal literal (std::stof throws on those),
                    // and "nan" / "inf" are legal PCD values; out-of-range values clamp instead of throwing
                    const char* cs = tok.c_str();
                    char* endp = nullptr;
                    if (slot == 8 && f.type == 'F') {  // PCL writes packed rgb as a float literal
                        float fv = std::strtof(cs, &endp);
                        if (endp == cs) throw std::runtime_error("pcd: bad number '" + tok + "' in '" + filename + "'");
                        std::memcpy(&out[i].rgba, &fv, 4);
                    } else {
                        const double dv = f.type == 'F' ? std::strtod(cs, &endp) : static_cast<double>(std::strtoll(cs, &endp, 10));
                        if (endp == cs) throw std::runtime_error("pcd: bad number '" + tok + "' in '" + filename + "'");
                        store(out[i], slot, f, dv, nullptr);
                    }
                }
            }
        }
    } else if (data_kind == "binary") {
        std::vector<char> buf(static_cast<size_t>(rec) * points);
        in.read(buf.data(), static_cast<std::streamsize>(buf.size()));
        if (static_cast<size_t>(in.gcount()) != buf.size())
            throw std::runtime_error("pcd: truncated binary data in '" + filename + "'");
        for (const auto& f : fields) {
            const int slot = surfel_slot(f.name);
            if (slot < 0) continue;
            for (uint64_t i = 0; i < points; ++i) {
                const char* p = buf.data() + static_cast<size_t>(rec) * i + f.offset;
                store(out[i], slot, f, slot == 8 && f.size == 4 ? 0.0 : read_scalar(p, f), p);
            }
        }
    } else if (data_kind == "binary_compressed") {
        uint32_t sizes[2] = {0u, 0u};  // compressed, raw
        in.read(reinterpret_cast<char*>(sizes), 8);
        if (in.gcount() != 8) throw std::runtime_error("pcd: truncated binary_compressed header in '" + filename + "'");
        if (static_cast<uint64_t>(sizes[1]) != static_cast<uint64_t>(rec) * points)
            throw std::runtime_error("pcd: binary_compressed size does not match FIELDS x POINTS in '" + filename + "'");
        std::vector<unsigned char> packed(sizes[0]);
        in.read(reinterpret_cast<char*>(packed.data()), static_cast<std::streamsize>(packed.size()));
        if (static_cast<size_t>(in.gcount()) != packed.size())
            throw std::runtime_error("pcd: truncated binary_compressed data in '" + filename + "'");
        std::vector<char> buf(sizes[1]);
        if (!lzf_decompress(packed.data(), packed.size(), buf))
            throw std::runtime_error("pcd: corrupt LZF stream in '" + filename + "'");
        // field-major: all values of field 0, then field 1, ... (each value f.size * f.count bytes)
        size_t column = 0;
        for (const auto& f : fields) {
            const int slot = surfel_slot(f.name);
            const size_t stride = static_cast<size_t>(f.size) * f.count;
            if (slot >= 0)
                for (uint64_t i = 0; i < points; ++i) {
                    const char* p = buf.data() + column + stride * i;
                    store(out[i], slot, f, slot == 8 && f.size == 4 ? 0.0 : read_scalar(p, f), p);
                }
            column += stride * points;
        }
    } else {
        throw std::runtime_error("pcd: DATA '" + data_kind + "' is not supported (ascii, binary, binary_compressed are)");
    }
    for (auto& p : out) p.data[3] = 1.f;
}

// writes the full surfel record (x y z normal_x normal_y normal_z rgba radius confidence curvature)
inline void save(const std::string& filename, const std::vector<pcl::PointSurfel>& pts, bool binary = true) {
    std::ofstream o(filename, std::ios::binary);
    if (!o) throw std::runtime_error("pcd: cannot write '" + filename + "'");
    o << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n"
      << "FIELDS x y z normal_x normal_y normal_z rgba radius confidence curvature\n"
      << "SIZE 4 4 4 4 4 4 4 4 4 4\nTYPE F F F F F F U F F F\nCOUNT 1 1 1 1 1 1 1 1 1 1\n"
      << "WIDTH " << pts.size() << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << pts.size() << "\n"
      << "DATA " << (binary ? "binary" : "ascii") << "\n";
    for (const auto& p : pts) {
        if (binary) {
            float r[10] = {p.x, p.y, p.z, p.normal_x, p.normal_y, p.normal_z, 0.f, p.radius, p.confidence, p.curvature};
            std::memcpy(&r[6], &p.rgba, 4);
            o.write(reinterpret_cast<const char*>(r), sizeof(r));
        } else {
            char line[512];
            std::snprintf(line, sizeof(line), "%.9g %.9g %.9g %.9g %.9g %.9g %u %.9g %.9g %.9g\n", p.x, p.y, p.z,
                          p.normal_x, p.normal_y, p.normal_z, p.rgba, p.radius, p.confidence, p.curvature);
            o << line;
        }
    }
}

}  // namespace pcd
}  // namespace triplet_match

#endif  // TRIPLET_MATCH_PCD_IO_HPP_
