// assets.hpp — C++ host-side asset I/O around the core, header-only, no dependencies.
//
// Mirrors cr::asset_loader (src/util/asset_loader.cpp) on top of crender.hpp's model_data / image:
//   load_model          :182-303  OBJ + MTL (tinyobj with triangulate = true in the reference): Kd -> colour (alpha 1),
//                                 every material `smooth`, emission 0; map_Kd decoded to RGBA/255 and flipped
//                                 vertically (stbi_set_flip_vertically_on_load). The reference decodes with
//                                 stb_image; here an own PNG decoder (8-bit, non-interlaced) — any other texture
//                                 format fails to load and the material stays untextured, which is what the reference
//                                 does for a texture it cannot read (:270-271).
//   export_framebuffer  :348-377  ./out/<name><ext>, " (n)" suffix when the file exists
//   export_png          :89-101   byte = min(x*255, 255) on all four channels (own writer, stored deflate blocks)
//   export_hdr          :172-178  pow(x, 2.2), then Radiance RGBE (flat scanlines)
//   export_exr          :112-170  three HALF channels in B, G, R order, tinyexr's float->half rounding
//   (JPG export needs a JPEG encoder and is provided by the Python host only: crender_b200/assets.py)
// The same behaviour as crender_b200/assets.py; tests/test_assets.py compares the two on the same files.
#pragma once

#include "crender.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <sys/stat.h>

namespace crb
{
    namespace asset_loader
    {
        enum class image_type
        {
            PNG,
            JPG,
            EXR,
            HDR
        };

        namespace detail
        {
            inline bool file_exists(const std::string &p)
            {
                struct stat st;
                return ::stat(p.c_str(), &st) == 0;
            }
            inline bool read_file(const std::string &p, std::vector<uint8_t> &out)
            {
                std::ifstream f(p, std::ios::binary);
                if (!f) return false;
                out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
                return true;
            }

            // ---------------------------------------------------------------- inflate (RFC 1951) for PNG textures
            struct bit_reader
            {
                const uint8_t *p;
                size_t         n, pos = 0;
                uint32_t       acc = 0;
                int            cnt = 0;
                bool           bad = false;
                uint32_t       bits(int k)
                {
                    while (cnt < k)
                    {
                        if (pos >= n)
                        {
                            bad = true;
                            return 0;
                        }
                        acc |= uint32_t(p[pos++]) << cnt, cnt += 8;
                    }
                    const uint32_t v = acc & ((1u << k) - 1u);
                    acc >>= k, cnt -= k;
                    return k ? v : 0;
                }
            };
            struct huffman
            {
                uint16_t count[16] = {}, symbol[320] = {};
                void     build(const uint8_t *len, int n)
                {
                    for (int i = 0; i < 16; i++) count[i] = 0;
                    for (int i = 0; i < n; i++) count[len[i]]++;
                    count[0] = 0;
                    uint16_t offs[16] = {};
                    for (int i = 1; i < 16; i++) offs[i] = uint16_t(offs[i - 1] + count[i - 1]);
                    for (int i = 0; i < n; i++)
                        if (len[i]) symbol[offs[len[i]]++] = uint16_t(i);
                }
                int decode(bit_reader &br) const
                {
                    int code = 0, first = 0, index = 0;
                    for (int l = 1; l < 16; l++)
                    {
                        code |= int(br.bits(1));
                        if (br.bad) return -1;
                        const int c = count[l];
                        if (code - c < first) return symbol[index + (code - first)];
                        index += c, first += c, first <<= 1, code <<= 1;
                    }
                    return -1;
                }
            };
            inline bool inflate(const uint8_t *src, size_t n, std::vector<uint8_t> &out)
            {
                static const uint16_t lbase[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
                static const uint8_t  lext[29]  = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
                static const uint16_t dbase[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
                static const uint8_t  dext[30]  = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
                static const uint8_t  order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
                bit_reader br { src, n };
                for (;;)
                {
                    const uint32_t last = br.bits(1), type = br.bits(2);
                    if (br.bad || type == 3) return false;
                    if (type == 0)
                    {
                        br.acc = 0, br.cnt = 0;
                        if (br.pos + 4 > n) return false;
                        const uint32_t len = src[br.pos] | (src[br.pos + 1] << 8);
                        br.pos += 4;
                        if (br.pos + len > n) return false;
                        out.insert(out.end(), src + br.pos, src + br.pos + len);
                        br.pos += len;
                    }
                    else
                    {
                        huffman hl, hd;
                        uint8_t len[320];
                        if (type == 1)
                        {
                            for (int i = 0; i < 288; i++) len[i] = uint8_t(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
                            hl.build(len, 288);
                            for (int i = 0; i < 30; i++) len[i] = 5;
                            hd.build(len, 30);
                        }
                        else
                        {
                            const int nl = int(br.bits(5)) + 257, nd = int(br.bits(5)) + 1, nc = int(br.bits(4)) + 4;
                            uint8_t   cl[19] = {};
                            for (int i = 0; i < nc; i++) cl[order[i]] = uint8_t(br.bits(3));
                            huffman hc;
                            hc.build(cl, 19);
                            int i = 0;
                            while (i < nl + nd)
                            {
                                const int s = hc.decode(br);
                                if (s < 0) return false;
                                if (s < 16)
                                    len[i++] = uint8_t(s);
                                else
                                {
                                    uint8_t v = 0;
                                    int     rep;
                                    if (s == 16)
                                    {
                                        if (!i) return false;
                                        v = len[i - 1], rep = 3 + int(br.bits(2));
                                    }
                                    else
                                        rep = s == 17 ? 3 + int(br.bits(3)) : 11 + int(br.bits(7));
                                    if (i + rep > nl + nd) return false;
                                    while (rep--) len[i++] = v;
                                }
                            }
                            hl.build(len, nl), hd.build(len + nl, nd);
                        }
                        for (;;)
                        {
                            const int s = hl.decode(br);
                            if (s < 0) return false;
                            if (s < 256)
                                out.push_back(uint8_t(s));
                            else if (s == 256)
                                break;
                            else
                            {
                                if (s > 285) return false;
                                const int l  = lbase[s - 257] + int(br.bits(lext[s - 257]));
                                const int ds = hd.decode(br);
                                if (ds < 0 || ds > 29) return false;
                                const size_t d = dbase[ds] + br.bits(dext[ds]);
                                if (d > out.size()) return false;
                                for (int k = 0; k < l; k++) out.push_back(out[out.size() - d]);
                            }
                        }
                    }
                    if (br.bad) return false;
                    if (last) return true;
                }
            }

            inline uint32_t be32(const uint8_t *p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

            // 8-bit, non-interlaced PNG of colour type 0/2/3/4/6 -> RGBA bytes, top row first
            inline bool decode_png(const std::vector<uint8_t> &f, uint32_t &w, uint32_t &h, std::vector<uint8_t> &rgba)
            {
                static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 13, 10, 26, 10 };
                if (f.size() < 33 || memcmp(f.data(), sig, 8) != 0) return false;
                std::vector<uint8_t> idat, plte, trns;
                int                  depth = 0, ctype = 0, interlace = 0;
                w = h = 0;
                for (size_t pos = 8; pos + 12 <= f.size();)
                {
                    const uint32_t len = be32(&f[pos]);
                    const char    *tag = reinterpret_cast<const char *>(&f[pos + 4]);
                    if (pos + 12 + len > f.size()) return false;
                    const uint8_t *d = &f[pos + 8];
                    if (!memcmp(tag, "IHDR", 4) && len >= 13)
                        w = be32(d), h = be32(d + 4), depth = d[8], ctype = d[9], interlace = d[12];
                    else if (!memcmp(tag, "PLTE", 4))
                        plte.assign(d, d + len);
                    else if (!memcmp(tag, "tRNS", 4))
                        trns.assign(d, d + len);
                    else if (!memcmp(tag, "IDAT", 4))
                        idat.insert(idat.end(), d, d + len);
                    else if (!memcmp(tag, "IEND", 4))
                        break;
                    pos += 12 + len;
                }
                if (!w || !h || depth != 8 || interlace || idat.size() < 6) return false;
                const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
                if (!ch) return false;
                std::vector<uint8_t> raw;
                raw.reserve((size_t(w) * ch + 1) * h);
                if (!inflate(idat.data() + 2, idat.size() - 2, raw)) return false;
                const size_t stride = size_t(w) * ch;
                if (raw.size() < (stride + 1) * h) return false;
                std::vector<uint8_t> img(stride * h);
                for (uint32_t y = 0; y < h; y++)
                {
                    const uint8_t  ft  = raw[(stride + 1) * y];
                    const uint8_t *in  = &raw[(stride + 1) * y + 1];
                    uint8_t       *out = &img[stride * y];
                    const uint8_t *up  = y ? &img[stride * (y - 1)] : nullptr;
                    for (size_t i = 0; i < stride; i++)
                    {
                        const int a = i >= size_t(ch) ? out[i - ch] : 0, b = up ? up[i] : 0, c = (up && i >= size_t(ch)) ? up[i - ch] : 0;
                        int       pr = 0;
                        switch (ft)
                        {
                        case 0: pr = 0; break;
                        case 1: pr = a; break;
                        case 2: pr = b; break;
                        case 3: pr = (a + b) >> 1; break;
                        case 4:
                        {
                            const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                            pr          = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                            break;
                        }
                        default: return false;
                        }
                        out[i] = uint8_t(in[i] + pr);
                    }
                }
                rgba.resize(size_t(w) * h * 4);
                for (size_t i = 0; i < size_t(w) * h; i++)
                {
                    const uint8_t *s = &img[i * ch];
                    uint8_t       *o = &rgba[i * 4];
                    switch (ctype)
                    {
                    case 0: o[0] = o[1] = o[2] = s[0], o[3] = 255; break;
                    case 2: o[0] = s[0], o[1] = s[1], o[2] = s[2], o[3] = 255; break;
                    case 3:
                        if (size_t(s[0]) * 3 + 2 >= plte.size()) return false;
                        o[0] = plte[s[0] * 3], o[1] = plte[s[0] * 3 + 1], o[2] = plte[s[0] * 3 + 2];
                        o[3] = s[0] < trns.size() ? trns[s[0]] : 255;
                        break;
                    case 4: o[0] = o[1] = o[2] = s[0], o[3] = s[1]; break;
                    default: o[0] = s[0], o[1] = s[1], o[2] = s[2], o[3] = s[3]; break;
                    }
                }
                return true;
            }

            // ---------------------------------------------------------------- writers
            inline uint32_t crc32(const uint8_t *p, size_t n, uint32_t crc = 0)
            {
                static uint32_t table[256];
                static bool     init = false;
                if (!init)
                {
                    for (uint32_t i = 0; i < 256; i++)
                    {
                        uint32_t c = i;
                        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
                        table[i] = c;
                    }
                    init = true;
                }
                crc = ~crc;
                for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
                return ~crc;
            }
            inline void put_be32(std::vector<uint8_t> &v, uint32_t x)
            {
                for (int s = 24; s >= 0; s -= 8) v.push_back(uint8_t(x >> s));
            }
            // zlib stream of stored deflate blocks (valid for any inflater; no compression)
            inline std::vector<uint8_t> zlib_stored(const std::vector<uint8_t> &raw)
            {
                std::vector<uint8_t> z { 0x78, 0x01 };
                size_t               pos = 0;
                do
                {
                    const size_t n = std::min<size_t>(65535, raw.size() - pos);
                    z.push_back(pos + n == raw.size() ? 1 : 0);
                    z.push_back(uint8_t(n & 0xff)), z.push_back(uint8_t(n >> 8));
                    z.push_back(uint8_t(~n & 0xff)), z.push_back(uint8_t((~n >> 8) & 0xff));
                    z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
                    pos += n;
                } while (pos < raw.size());
                uint32_t a = 1, b = 0;
                for (uint8_t c : raw) a = (a + c) % 65521u, b = (b + a) % 65521u;
                put_be32(z, (b << 16) | a);
                return z;
            }
            inline void png_chunk(std::vector<uint8_t> &f, const char *tag, const std::vector<uint8_t> &d)
            {
                put_be32(f, uint32_t(d.size()));
                const size_t at = f.size();
                f.insert(f.end(), tag, tag + 4);
                f.insert(f.end(), d.begin(), d.end());
                put_be32(f, crc32(&f[at], 4 + d.size()));
            }
            inline bool write_file(const std::string &p, const std::vector<uint8_t> &d)
            {
                FILE *f = std::fopen(p.c_str(), "wb");
                if (!f) return false;
                const bool ok = std::fwrite(d.data(), 1, d.size(), f) == d.size();
                std::fclose(f);
                return ok;
            }
            // asset_loader.cpp:92: data[i] = glm::min(buffer[i] * 255.f, 255.f) stored to uint8_t (truncation)
            inline uint8_t to_byte(float x)
            {
                const float v = std::fmin(x * 255.0f, 255.0f);
                return (v != v || v <= 0.0f) ? 0 : uint8_t(v);
            }
            // tinyexr's float_to_half_full: mantissa truncated to 10 bits, +1 when the first dropped bit is set
            inline uint16_t float_to_half_bits(float x)
            {
                uint32_t f;
                memcpy(&f, &x, 4);
                const uint32_t sign = (f >> 31) & 1, exp = (f >> 23) & 0xff, man = f & 0x7fffff;
                uint32_t       out = 0;
                const int      ne  = int(exp) - 127 + 15;
                if (exp == 255)
                    out = (31u << 10) | (man ? 0x200u : 0u);
                else if (exp == 0)
                    out = 0;
                else if (ne >= 31)
                    out = 31u << 10;
                else if (ne > 0)
                    out = ((uint32_t(ne) << 10) | (man >> 13)) + ((man >> 12) & 1u);
                else if (14 - ne <= 24)
                {
                    const uint32_t m = man | 0x800000u;
                    const int      sh = 14 - ne;
                    out               = (m >> sh) + ((m >> (sh - 1)) & 1u);
                }
                return uint16_t((sign << 15) | (out & 0x7fffu));
            }
            template<typename T>
            inline void put_le(std::vector<uint8_t> &v, T x)
            {
                uint8_t b[sizeof(T)];
                memcpy(b, &x, sizeof(T));
                v.insert(v.end(), b, b + sizeof(T));
            }
            inline void exr_attr(std::vector<uint8_t> &v, const char *name, const char *type, const std::vector<uint8_t> &val)
            {
                v.insert(v.end(), name, name + strlen(name) + 1);
                v.insert(v.end(), type, type + strlen(type) + 1);
                put_le<int32_t>(v, int32_t(val.size()));
                v.insert(v.end(), val.begin(), val.end());
            }
        }    // namespace detail

        // stbi_load(..., 4) with stbi_set_flip_vertically_on_load(true), then /255 (asset_loader.cpp:237-262)
        inline bool load_texture(const std::string &path, image &out)
        {
            std::vector<uint8_t> file, rgba;
            uint32_t             w = 0, h = 0;
            if (!detail::read_file(path, file) || !detail::decode_png(file, w, h, rgba)) return false;
            out.width = w, out.height = h;
            out.data.resize(size_t(w) * h * 4);
            for (uint32_t y = 0; y < h; y++)
                for (size_t i = 0; i < size_t(w) * 4; i++) out.data[size_t(y) * w * 4 + i] = float(rgba[size_t(h - 1 - y) * w * 4 + i]) / 255.f;
            return true;
        }

        // cr::asset_loader::load_model (asset_loader.cpp:182-303). Quads are split along the shorter diagonal, larger polygons fan-triangulated (what tinyobj does for
        // convex faces). Faces without a material get a default material appended (tinyobj reports id -1, which the
        // reference would use to index materials[] out of bounds).
        inline model_data load_model(const std::string &file, std::string folder = std::string())
        {
            if (folder.empty())
            {
                const size_t s = file.find_last_of('/');
                folder         = s == std::string::npos ? "." : file.substr(0, s);
            }
            model_data md;
            {
                const size_t s = file.find_last_of('/'), b = s == std::string::npos ? 0 : s + 1, d = file.find_last_of('.');
                md.name        = file.substr(b, (d == std::string::npos || d < b) ? std::string::npos : d - b);
            }
            std::ifstream in(file);
            if (!in) throw error(CRB_ERR_INVALID_ARG, "Couldn't parse OBJ from file");    // cr::exit in the reference (:193-194)
            struct mtl
            {
                std::string name, map_kd;
                float       kd[3] = { 0.6f, 0.6f, 0.6f };    // tinyobj's diffuse default
            };
            std::vector<mtl>           mtls;
            std::map<std::string, int> by_name;
            std::vector<int64_t>       ti, mi;
            int                        cur = -1;
            auto rest = [](std::istringstream &ss) {
                std::string r;
                std::getline(ss, r);
                const size_t a = r.find_first_not_of(" \t"), b = r.find_last_not_of(" \t\r");
                return a == std::string::npos ? std::string() : r.substr(a, b - a + 1);
            };
            std::string line;
            while (std::getline(in, line))
            {
                std::istringstream ss(line);
                std::string        tag;
                if (!(ss >> tag) || tag[0] == '#') continue;
                if (tag == "v")
                {
                    vec3 v { 0, 0, 0 };
                    ss >> v[0] >> v[1] >> v[2];
                    md.vertices.push_back(v);
                }
                else if (tag == "vt")
                {
                    vec2 t { 0, 0 };
                    ss >> t[0] >> t[1];
                    md.texture_coords.push_back(t);
                }
                else if (tag == "mtllib")
                {
                    std::ifstream mf(folder + "/" + rest(ss));
                    std::string   ml;
                    while (std::getline(mf, ml))
                    {
                        std::istringstream ms(ml);
                        std::string        mt;
                        if (!(ms >> mt) || mt[0] == '#') continue;
                        if (mt == "newmtl")
                        {
                            mtl m;
                            m.name          = rest(ms);
                            by_name[m.name] = int(mtls.size());
                            mtls.push_back(m);
                        }
                        else if (!mtls.empty() && mt == "Kd")
                            ms >> mtls.back().kd[0] >> mtls.back().kd[1] >> mtls.back().kd[2];
                        else if (!mtls.empty() && mt == "map_Kd")
                        {
                            std::string tok, lasttok;
                            while (ms >> tok) lasttok = tok;
                            mtls.back().map_kd = lasttok;
                        }
                    }
                }
                else if (tag == "usemtl")
                {
                    const auto it = by_name.find(rest(ss));
                    cur           = it == by_name.end() ? -1 : it->second;
                }
                else if (tag == "f")
                {
                    std::vector<std::pair<int64_t, int64_t>> corners;
                    std::string                              c;
                    while (ss >> c)
                    {
                        int64_t      v = 0, t = 0;
                        const size_t s1 = c.find('/');
                        v               = std::atoll(c.substr(0, s1).c_str());
                        if (s1 != std::string::npos)
                        {
                            const size_t      s2 = c.find('/', s1 + 1);
                            const std::string ts = c.substr(s1 + 1, s2 == std::string::npos ? std::string::npos : s2 - s1 - 1);
                            if (!ts.empty()) t = std::atoll(ts.c_str());
                        }
                        corners.push_back({ v > 0 ? v - 1 : int64_t(md.vertices.size()) + v, t ? (t > 0 ? t - 1 : int64_t(md.texture_coords.size()) + t) : -1 });
                    }
                    auto emit = [&](size_t a, size_t b, size_t c3) {
                        for (const auto &cc : { corners[a], corners[b], corners[c3] }) md.vertex_indices.push_back(uint32_t(cc.first)), ti.push_back(cc.second);
                        mi.push_back(cur);
                    };
                    bool quad_done = false;
                    if (corners.size() == 4)
                    {
                        // tinyobj (the reference's vendored loader, external/tinyobj/tinobj.h:1394-1490) splits a quad along
                        // its SHORTER diagonal, in float: |v2-v0|^2 < |v3-v1|^2 -> (0,1,2)(0,2,3), else (0,1,3)(1,2,3)
                        bool in_range = true;
                        for (const auto &cc : corners) in_range = in_range && cc.first >= 0 && size_t(cc.first) < md.vertices.size();
                        if (in_range)
                        {
                            const vec3 &p0 = md.vertices[size_t(corners[0].first)], &p1 = md.vertices[size_t(corners[1].first)], &p2 = md.vertices[size_t(corners[2].first)],
                                       &p3 = md.vertices[size_t(corners[3].first)];
                            const float ax = p2[0] - p0[0], ay = p2[1] - p0[1], az = p2[2] - p0[2], bx = p3[0] - p1[0], by = p3[1] - p1[1], bz = p3[2] - p1[2];
                            const float sqr02 = ax * ax + ay * ay + az * az, sqr13 = bx * bx + by * by + bz * bz;
                            if (sqr02 < sqr13)
                                emit(0, 1, 2), emit(0, 2, 3);
                            else
                                emit(0, 1, 3), emit(1, 2, 3);
                            quad_done = true;
                        }
                    }
                    // larger polygons: a fan, which is what tinyobj's ear clipping produces for convex faces
                    if (!quad_done)
                        for (size_t k = 1; k + 1 < corners.size(); k++) emit(0, k, k + 1);
                }
            }
            std::map<std::string, uint32_t> already;
            for (const mtl &m : mtls)
            {
                material mat;
                mat.info.name = m.name, mat.info.colour = { m.kd[0], m.kd[1], m.kd[2], 1.0f };
                mat.info.shade_type = material::smooth, mat.info.emission = 0.0f;
                if (!m.map_kd.empty())
                {
                    const auto it = already.find(m.map_kd);
                    if (it != already.end())
                        mat.info.tex = it->second;
                    else
                    {
                        image tex;
                        if (load_texture(folder + "/" + m.map_kd, tex))
                        {
                            md.textures.push_back(std::move(tex));
                            mat.info.tex      = uint32_t(md.textures.size() - 1);
                            already[m.map_kd] = *mat.info.tex;
                        }
                        else
                            std::fprintf(stderr, "[warn] Failed to find texture [%s], defaulting to blank material\n", m.map_kd.c_str());
                    }
                }
                md.materials.push_back(mat);
            }
            bool missing = md.materials.empty();
            for (int64_t m : mi) missing = missing || m < 0;
            if (missing)
            {
                material d;
                d.info.name = "default";
                md.materials.push_back(d);
            }
            for (int64_t m : mi) md.material_indices.push_back(uint32_t(m < 0 ? int64_t(md.materials.size()) - 1 : m));
            bool all_uv = !ti.empty() && !md.texture_coords.empty();
            for (int64_t t : ti) all_uv = all_uv && t >= 0;
            if (all_uv)
                for (int64_t t : ti) md.texture_indices.push_back(uint32_t(t));
            return md;
        }

        inline bool export_png(const std::string &path, const float *rgba, uint32_t w, uint32_t h)
        {
            std::vector<uint8_t> raw;
            raw.reserve((size_t(w) * 4 + 1) * h);
            for (uint32_t y = 0; y < h; y++)
            {
                raw.push_back(0);    // filter: none
                for (size_t i = 0; i < size_t(w) * 4; i++) raw.push_back(detail::to_byte(rgba[size_t(y) * w * 4 + i]));
            }
            std::vector<uint8_t> f { 0x89, 'P', 'N', 'G', 13, 10, 26, 10 }, ihdr;
            detail::put_be32(ihdr, w), detail::put_be32(ihdr, h);
            ihdr.insert(ihdr.end(), { 8, 6, 0, 0, 0 });
            detail::png_chunk(f, "IHDR", ihdr);
            detail::png_chunk(f, "IDAT", detail::zlib_stored(raw));
            detail::png_chunk(f, "IEND", {});
            return detail::write_file(path, f);
        }

        inline bool export_hdr(const std::string &path, const float *rgba, uint32_t w, uint32_t h)
        {
            char head[160];
            const int n = std::snprintf(head, sizeof(head), "#?RADIANCE\n# Written by crender_b200 (flat RGBE)\nFORMAT=32-bit_rle_rgbe\n\n-Y %u +X %u\n", h, w);
            std::vector<uint8_t> f(head, head + n);
            for (size_t i = 0; i < size_t(w) * h; i++)
            {
                float c[3];
                for (int k = 0; k < 3; k++) c[k] = std::pow(rgba[i * 4 + k], 2.2f);    // asset_loader.cpp:175
                const float m = std::fmax(c[0], std::fmax(c[1], c[2]));
                uint8_t     px[4] = { 0, 0, 0, 0 };
                if (m >= 1e-32f)
                {
                    int         e    = 0;
                    const float norm = std::frexp(m, &e) * 256.0f / m;    // stbiw__linear_to_rgbe
                    for (int k = 0; k < 3; k++) px[k] = uint8_t(c[k] * norm);
                    px[3] = uint8_t(e + 128);
                }
                f.insert(f.end(), px, px + 4);
            }
            return detail::write_file(path, f);
        }

        // OpenEXR 2 single-part scanline file, HALF channels B, G, R (asset_loader.cpp:135-161), ZIP_COMPRESSION with
        // 16-line blocks; a block is stored raw when compression does not shrink it (OpenEXR's rule), which without a
        // deflate compressor here is every block. Decodes to exactly the pixels the reference's file decodes to.
        inline bool export_exr(const std::string &path, const float *rgba, uint32_t w, uint32_t h)
        {
            using namespace detail;
            std::vector<uint8_t> f { 0x76, 0x2f, 0x31, 0x01 }, chl, box, one, zero2(8, 0);
            put_le<int32_t>(f, 2);
            for (const char ch : { 'B', 'G', 'R' })
            {
                chl.push_back(uint8_t(ch)), chl.push_back(0);
                put_le<int32_t>(chl, 1);
                chl.insert(chl.end(), 4, 0);
                put_le<int32_t>(chl, 1), put_le<int32_t>(chl, 1);
            }
            chl.push_back(0);
            put_le<int32_t>(box, 0), put_le<int32_t>(box, 0), put_le<int32_t>(box, int32_t(w) - 1), put_le<int32_t>(box, int32_t(h) - 1);
            put_le<float>(one, 1.0f);
            exr_attr(f, "channels", "chlist", chl);
            exr_attr(f, "compression", "compression", { 3 });
            exr_attr(f, "dataWindow", "box2i", box);
            exr_attr(f, "displayWindow", "box2i", box);
            exr_attr(f, "lineOrder", "lineOrder", { 0 });
            exr_attr(f, "pixelAspectRatio", "float", one);
            exr_attr(f, "screenWindowCenter", "v2f", zero2);
            exr_attr(f, "screenWindowWidth", "float", one);
            f.push_back(0);
            const uint32_t nblk = (h + 15) / 16;
            uint64_t       off  = f.size() + 8ull * nblk;
            for (uint32_t b = 0; b < nblk; b++)
            {
                put_le<uint64_t>(f, off);
                off += 8 + uint64_t(std::min<uint32_t>(16, h - b * 16)) * w * 3 * 2;
            }
            for (uint32_t y0 = 0; y0 < h; y0 += 16)
            {
                const uint32_t lines = std::min<uint32_t>(16, h - y0);
                put_le<int32_t>(f, int32_t(y0)), put_le<int32_t>(f, int32_t(lines * w * 3 * 2));
                for (uint32_t y = y0; y < y0 + lines; y++)
                    for (int ch = 2; ch >= 0; ch--)    // B row, G row, R row
                        for (uint32_t x = 0; x < w; x++) put_le<uint16_t>(f, float_to_half_bits(rgba[(size_t(y) * w + x) * 4 + ch]));
            }
            return write_file(path, f);
        }

        // Baseline JPEG (ITU-T T.81) writer for export_framebuffer(JPG): the reference calls stbi_write_jpg(path, w, h, 4,
        // data, 100) (asset_loader.cpp:103-109), i.e. quality 100: every quantiser step is 1, no chroma subsampling,
        // Y/Cb/Cr from RGB (alpha ignored), 8x8 forward DCT, the standard's example Huffman tables (Annex K.3). This is an
        // own implementation of the standard; decoded pixels agree with the reference's file to within rounding
        // (tests/test_reference_anchor.py compares both through the reference's own decoder).
        namespace detail
        {
            struct jpeg_bits
            {
                std::vector<uint8_t> &out;
                uint32_t              acc = 0;
                int                   n   = 0;
                void put(uint32_t code, int len)
                {
                    acc = (acc << len) | (code & ((1u << len) - 1u));
                    n += len;
                    while (n >= 8)
                    {
                        const uint8_t b = uint8_t(acc >> (n - 8));
                        out.push_back(b);
                        if (b == 0xff) out.push_back(0);    // byte stuffing
                        n -= 8;
                    }
                }
                void flush()
                {
                    if (n) put(0x7f, 8 - n);    // pad with ones
                }
            };
            struct jpeg_huff
            {
                uint16_t code[256];
                uint8_t  len[256];
                jpeg_huff(const uint8_t *bits /*16*/, const uint8_t *vals)
                {
                    for (int i = 0; i < 256; i++) code[i] = 0, len[i] = 0;
                    uint16_t c = 0;
                    int      k = 0;
                    for (int l = 1; l <= 16; l++)
                    {
                        for (int i = 0; i < bits[l - 1]; i++, k++) code[vals[k]] = c++, len[vals[k]] = uint8_t(l);
                        c <<= 1;
                    }
                }
            };
            inline void jpeg_fdct8x8(float *b)
            {
                // separable 8-point DCT-II, orthonormal scaling of T.81 A.3.3
                static float C[8][8];
                static bool  init = false;
                if (!init)
                {
                    for (int u = 0; u < 8; u++)
                        for (int x = 0; x < 8; x++) C[u][x] = float((u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * M_PI / 16.0));
                    init = true;
                }
                float t[64];
                for (int y = 0; y < 8; y++)
                    for (int u = 0; u < 8; u++)
                    {
                        float s = 0;
                        for (int x = 0; x < 8; x++) s += C[u][x] * b[y * 8 + x];
                        t[y * 8 + u] = s;
                    }
                for (int v = 0; v < 8; v++)
                    for (int u = 0; u < 8; u++)
                    {
                        float s = 0;
                        for (int y = 0; y < 8; y++) s += C[v][y] * t[y * 8 + u];
                        b[v * 8 + u] = s;
                    }
            }
        }    // namespace detail

        inline bool export_jpg(const std::string &path, const float *rgba, uint32_t w, uint32_t h)
        {
            using namespace detail;
            static const uint8_t zz[64] = { 0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                            35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };
            // Annex K.3 tables
            static const uint8_t dc_l_bits[16] = { 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0 }, dc_c_bits[16] = { 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0 };
            static const uint8_t dc_vals[12]   = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 };
            static const uint8_t ac_l_bits[16] = { 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d }, ac_c_bits[16] = { 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77 };
            static const uint8_t ac_l_vals[162] = {
                0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1,
                0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
                0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a,
                0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
                0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3,
                0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa
            };
            static const uint8_t ac_c_vals[162] = {
                0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1,
                0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
                0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
                0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
                0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca,
                0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa
            };
            const jpeg_huff hdc[2] = { jpeg_huff(dc_l_bits, dc_vals), jpeg_huff(dc_c_bits, dc_vals) };
            const jpeg_huff hac[2] = { jpeg_huff(ac_l_bits, ac_l_vals), jpeg_huff(ac_c_bits, ac_c_vals) };

            std::vector<uint8_t> f { 0xff, 0xd8, 0xff, 0xe0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0 };
            for (int t = 0; t < 2; t++)    // two quantisation tables, every step 1 (quality 100)
            {
                f.insert(f.end(), { 0xff, 0xdb, 0, 67, uint8_t(t) });
                f.insert(f.end(), 64, 1);
            }
            f.insert(f.end(), { 0xff, 0xc0, 0, 17, 8, uint8_t(h >> 8), uint8_t(h), uint8_t(w >> 8), uint8_t(w), 3, 1, 0x11, 0, 2, 0x11, 1, 3, 0x11, 1 });
            auto dht = [&](int cls_id, const uint8_t *bits, const uint8_t *vals, int nvals) {
                f.insert(f.end(), { 0xff, 0xc4, uint8_t((19 + nvals) >> 8), uint8_t(19 + nvals), uint8_t(cls_id) });
                f.insert(f.end(), bits, bits + 16);
                f.insert(f.end(), vals, vals + nvals);
            };
            dht(0x00, dc_l_bits, dc_vals, 12), dht(0x10, ac_l_bits, ac_l_vals, 162), dht(0x01, dc_c_bits, dc_vals, 12), dht(0x11, ac_c_bits, ac_c_vals, 162);
            f.insert(f.end(), { 0xff, 0xda, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0 });

            jpeg_bits bw { f };
            int       pred[3] = { 0, 0, 0 };
            for (uint32_t by = 0; by < h; by += 8)
                for (uint32_t bx = 0; bx < w; bx += 8)
                {
                    float blk[3][64];
                    for (int y = 0; y < 8; y++)
                        for (int x = 0; x < 8; x++)
                        {
                            const uint32_t px = std::min(bx + uint32_t(x), w - 1), py = std::min(by + uint32_t(y), h - 1);    // edge replication
                            const float   *p  = rgba + (size_t(py) * w + px) * 4;
                            const float    r = to_byte(p[0]), g = to_byte(p[1]), b = to_byte(p[2]);
                            blk[0][y * 8 + x] = 0.299f * r + 0.587f * g + 0.114f * b - 128.0f;
                            blk[1][y * 8 + x] = -0.168736f * r - 0.331264f * g + 0.5f * b;
                            blk[2][y * 8 + x] = 0.5f * r - 0.418688f * g - 0.081312f * b;
                        }
                    for (int c = 0; c < 3; c++)
                    {
                        jpeg_fdct8x8(blk[c]);
                        int q[64];
                        for (int i = 0; i < 64; i++) q[i] = int(std::lround(blk[c][zz[i]]));
                        const int t = c ? 1 : 0;
                        auto      mag = [](int v, int &bits) {
                            int a = v < 0 ? -v : v, n = 0;
                            while (a) n++, a >>= 1;
                            bits = v < 0 ? v - 1 : v;
                            return n;
                        };
                        int       bits;
                        const int diff = q[0] - pred[c];
                        pred[c]        = q[0];
                        int n          = mag(diff, bits);
                        bw.put(hdc[t].code[n], hdc[t].len[n]);
                        if (n) bw.put(uint32_t(bits), n);
                        int run = 0, last = 63;
                        while (last > 0 && q[last] == 0) last--;
                        for (int i = 1; i <= last; i++)
                        {
                            if (q[i] == 0)
                            {
                                run++;
                                continue;
                            }
                            while (run > 15) bw.put(hac[t].code[0xf0], hac[t].len[0xf0]), run -= 16;
                            n = mag(q[i], bits);
                            bw.put(hac[t].code[(run << 4) | n], hac[t].len[(run << 4) | n]);
                            bw.put(uint32_t(bits), n);
                            run = 0;
                        }
                        if (last != 63) bw.put(hac[t].code[0], hac[t].len[0]);    // EOB
                    }
                }
            bw.flush();
            f.push_back(0xff), f.push_back(0xd9);
            return write_file(path, f);
        }

        // cr::asset_loader::export_framebuffer (asset_loader.cpp:348-377): writes out_dir/name.ext, or "name (n).ext"
        // when that exists; returns the file written
        inline std::string export_framebuffer(const image &buffer, const std::string &name, image_type type, const std::string &out_dir = "./out/")
        {
            const char *ext = type == image_type::PNG ? ".png" : type == image_type::JPG ? ".jpg" : type == image_type::EXR ? ".exr" : ".hdr";
            ::mkdir(out_dir.c_str(), 0777);
            const std::string base   = out_dir + (out_dir.empty() || out_dir.back() == '/' ? "" : "/") + name;
            std::string       target = base + ext;
            for (int n = 1; detail::file_exists(target); n++) target = base + " (" + std::to_string(n) + ")" + ext;
            const uint32_t w = uint32_t(buffer.width), h = uint32_t(buffer.height);
            bool           ok = false;
            switch (type)
            {
            case image_type::PNG: ok = export_png(target, buffer.data.data(), w, h); break;
            case image_type::HDR: ok = export_hdr(target, buffer.data.data(), w, h); break;
            case image_type::EXR: ok = export_exr(target, buffer.data.data(), w, h); break;
            case image_type::JPG: ok = export_jpg(target, buffer.data.data(), w, h); break;
            }
            if (!ok) throw error(CRB_ERR_GENERIC, ("cannot write " + target).c_str());
            return target;
        }
    }    // namespace asset_loader
}    // namespace crb
