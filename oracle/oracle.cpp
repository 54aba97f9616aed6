/*
 * oracle.cpp — CPU restatement of CRender's path-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h). How it is pinned: the reference ships no tests and no golden
 * vectors. Since round 2 this file is checked against THE REFERENCE'S OWN CODE: its translation units for
 * this path (renderer, scene, camera, model, registry, thread_pool, asset_loader ...) compile unmodified
 * from /root/reference against shim headers for the three absent third-party libraries (oracle/ref ->
 * oracle/_ref/ref_render); the oracle's reference-stream mode is bit-identical to that
 * binary on seven scenes, and its outputs are committed as golden fixtures (tests/golden,
 * tests/test_reference_anchor.py). What stays restated from published conventions rather than pinned is
 * what Embree and glm themselves compute (below): "parity unpinned" applies to those two only. The KATs of
 * SURVEY.md section 4 and a brute-force triangle loop pin the rest of the arithmetic independently.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Third-party arithmetic that the reference delegates to and that is NOT in /root/reference:
 *   - Intel Embree 3.x (find_package(embree 3.0), CMakeLists.txt:22; readme.md:26 -> v3.13.0):
 *     rtcCommitScene / rtcIntersect1 / rtcInterpolate0. Restated here as: own binned-SAH BVH +
 *     Moeller-Trumbore with Embree's published hit conventions: Ng = (v1-v0)x(v2-v0) unnormalised,
 *     barycentrics (u,v) weight v1 and v2, no back-face culling, hit iff tnear < t <= tfar.
 *   - glm 0.9.9.8 (external/CMakeLists.txt:29-33): normalize(v) = v * (1/sqrt(dot(v,v))),
 *     reflect(I,N) = I - N*dot(N,I)*2, mat*vec as a sum of scaled columns, rotate() = Rodrigues.
 *
 * Compiled with -ffp-contract=off so that every float operation below is a single IEEE-754
 * binary32 rounding; the product's triangle test follows the same operation sequence.
 */
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <random>
#include <thread>
#include <vector>

namespace
{
    constexpr float INF = std::numeric_limits<float>::infinity();

    // src/util/numbers.h:14-24
    constexpr float PI      = 3.14159265359f;
    constexpr float TAU     = 6.28318530717f;
    constexpr float INV_PI  = 1.0f / 3.14159265359f;
    constexpr float INV_TAU = 1.0f / 6.28318530717f;

    struct vec2
    {
        float x = 0, y = 0;
    };
    struct vec3
    {
        float x = 0, y = 0, z = 0;
    };
    struct vec4
    {
        float x = 0, y = 0, z = 0, w = 0;
    };
    inline vec3 V3(float x, float y, float z) { return vec3 { x, y, z }; }
    inline vec3 operator+(vec3 a, vec3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
    inline vec3 operator-(vec3 a, vec3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
    inline vec3 operator*(vec3 a, vec3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
    inline vec3 operator*(vec3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
    inline vec3 operator*(float s, vec3 a) { return V3(s * a.x, s * a.y, s * a.z); }
    inline vec3 operator/(vec3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
    inline vec3 operator-(vec3 a) { return V3(-a.x, -a.y, -a.z); }
    // glm::dot: tmp = a*b; tmp.x + tmp.y + tmp.z
    inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
    inline vec3  cross(vec3 a, vec3 b)
    {
        return V3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
    }
    inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
    // glm::normalize: v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x)
    inline vec3 normalize(vec3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }
    // glm::reflect: I - N * dot(N, I) * 2
    inline vec3  reflect(vec3 I, vec3 N) { return I - N * dot(N, I) * 2.0f; }
    inline float clampf(float x, float lo, float hi) { return std::min(std::max(x, lo), hi); }

    // column-major 4x4, m[c][r]
    struct mat4
    {
        float m[4][4];
    };
    inline mat4 identity4()
    {
        mat4 r {};
        for (int i = 0; i < 4; i++) r.m[i][i] = 1.0f;
        return r;
    }
    inline vec4 mul(const mat4 &M, vec4 v)
    {
        // glm: Mov0 = v[0], ...; Mul0 = m[0]*Mov0 ...; (Mul0+Mul1) + (Mul2+Mul3)
        vec4 r;
        r.x = (M.m[0][0] * v.x + M.m[1][0] * v.y) + (M.m[2][0] * v.z + M.m[3][0] * v.w);
        r.y = (M.m[0][1] * v.x + M.m[1][1] * v.y) + (M.m[2][1] * v.z + M.m[3][1] * v.w);
        r.z = (M.m[0][2] * v.x + M.m[1][2] * v.y) + (M.m[2][2] * v.z + M.m[3][2] * v.w);
        r.w = (M.m[0][3] * v.x + M.m[1][3] * v.y) + (M.m[2][3] * v.z + M.m[3][3] * v.w);
        return r;
    }
    inline vec3 mul_point(const mat4 &M, vec3 p)
    {
        vec4 r = mul(M, vec4 { p.x, p.y, p.z, 1.0f });
        return V3(r.x, r.y, r.z);
    }
    inline vec3 mul_dir(const mat4 &M, vec3 d)
    {
        vec4 r = mul(M, vec4 { d.x, d.y, d.z, 0.0f });
        return V3(r.x, r.y, r.z);
    }
    inline mat4 matmul(const mat4 &A, const mat4 &B)
    {
        mat4 R {};
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++)
            {
                float s = 0;
                for (int k = 0; k < 4; k++) s += A.m[k][r] * B.m[c][k];
                R.m[c][r] = s;
            }
        return R;
    }
    // glm::inverse(mat4) (model.cpp:109) = glm 0.9.9.8 func_matrix.inl compute_inverse<4,4>: 18 2x2 sub-determinants
    // (Coef..), six factor vectors, four cofactor columns with alternating signs, determinant from the first row. The
    // operation order is glm's, so that a general (scaled / sheared) instance transform inverts to the same bits.
    mat4 inverse(const mat4 &M)
    {
        const auto m = [&](int c, int r) { return M.m[c][r]; };
        const float Coef00 = m(2, 2) * m(3, 3) - m(3, 2) * m(2, 3), Coef02 = m(1, 2) * m(3, 3) - m(3, 2) * m(1, 3), Coef03 = m(1, 2) * m(2, 3) - m(2, 2) * m(1, 3);
        const float Coef04 = m(2, 1) * m(3, 3) - m(3, 1) * m(2, 3), Coef06 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3), Coef07 = m(1, 1) * m(2, 3) - m(2, 1) * m(1, 3);
        const float Coef08 = m(2, 1) * m(3, 2) - m(3, 1) * m(2, 2), Coef10 = m(1, 1) * m(3, 2) - m(3, 1) * m(1, 2), Coef11 = m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2);
        const float Coef12 = m(2, 0) * m(3, 3) - m(3, 0) * m(2, 3), Coef14 = m(1, 0) * m(3, 3) - m(3, 0) * m(1, 3), Coef15 = m(1, 0) * m(2, 3) - m(2, 0) * m(1, 3);
        const float Coef16 = m(2, 0) * m(3, 2) - m(3, 0) * m(2, 2), Coef18 = m(1, 0) * m(3, 2) - m(3, 0) * m(1, 2), Coef19 = m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2);
        const float Coef20 = m(2, 0) * m(3, 1) - m(3, 0) * m(2, 1), Coef22 = m(1, 0) * m(3, 1) - m(3, 0) * m(1, 1), Coef23 = m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1);
        struct v4
        {
            float x, y, z, w;
        };
        const auto mul4 = [](v4 a, v4 b) { return v4 { a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w }; };
        const auto sub4 = [](v4 a, v4 b) { return v4 { a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w }; };
        const auto add4 = [](v4 a, v4 b) { return v4 { a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w }; };
        const v4 Fac0 { Coef00, Coef00, Coef02, Coef03 }, Fac1 { Coef04, Coef04, Coef06, Coef07 }, Fac2 { Coef08, Coef08, Coef10, Coef11 };
        const v4 Fac3 { Coef12, Coef12, Coef14, Coef15 }, Fac4 { Coef16, Coef16, Coef18, Coef19 }, Fac5 { Coef20, Coef20, Coef22, Coef23 };
        const v4 Vec0 { m(1, 0), m(0, 0), m(0, 0), m(0, 0) }, Vec1 { m(1, 1), m(0, 1), m(0, 1), m(0, 1) };
        const v4 Vec2 { m(1, 2), m(0, 2), m(0, 2), m(0, 2) }, Vec3 { m(1, 3), m(0, 3), m(0, 3), m(0, 3) };
        const v4 Inv0 = add4(sub4(mul4(Vec1, Fac0), mul4(Vec2, Fac1)), mul4(Vec3, Fac2));
        const v4 Inv1 = add4(sub4(mul4(Vec0, Fac0), mul4(Vec2, Fac3)), mul4(Vec3, Fac4));
        const v4 Inv2 = add4(sub4(mul4(Vec0, Fac1), mul4(Vec1, Fac3)), mul4(Vec3, Fac5));
        const v4 Inv3 = add4(sub4(mul4(Vec0, Fac2), mul4(Vec1, Fac4)), mul4(Vec2, Fac5));
        const v4 SignA { +1, -1, +1, -1 }, SignB { -1, +1, -1, +1 };
        const v4 col[4] = { mul4(Inv0, SignA), mul4(Inv1, SignB), mul4(Inv2, SignA), mul4(Inv3, SignB) };
        const v4 Dot0 { m(0, 0) * col[0].x, m(0, 1) * col[1].x, m(0, 2) * col[2].x, m(0, 3) * col[3].x };
        const float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
        const float OneOverDeterminant = 1.0f / Dot1;
        mat4 R;
        for (int c = 0; c < 4; c++)
            R.m[c][0] = col[c].x * OneOverDeterminant, R.m[c][1] = col[c].y * OneOverDeterminant, R.m[c][2] = col[c].z * OneOverDeterminant,
            R.m[c][3] = col[c].w * OneOverDeterminant;
        return R;
    }
    inline bool is_identity(const mat4 &M)
    {
        mat4 I = identity4();
        return std::memcmp(&M, &I, sizeof(mat4)) == 0;
    }

    // ---------------------------------------------------------------- sampler
    // The reference sampler is a default-seeded thread_local mt19937 in every worker thread
    // (renderer.cpp:6-11) and is not reproducible. The oracle therefore DEFINES the
    // counter-based sampler that both it and the product use:
    //   k1 = mix(mix(mix(seed + 0x9e3779b9) ^ pixel) ^ sample),  k2 = mix(mix(mix(seed + 0x85ebca6b) ^ sample) ^ pixel),
    //   u(dim) = (mix(mix(k1 + dim*0x9e3779b9) ^ k2) >> 8) * 2^-24
    // with mix = the 32-bit "lowbias32" integer finaliser. The key is 64 bits wide (two independent hashes of the same
    // (seed, pixel, sample)): with one 32-bit key, 3 % of the 1.3e8 paths of a 1080p 64-spp frame shared their whole
    // random stream with another path, and keys differing by a multiple of the dimension stride shared shifted streams.
    // Dimensions are assigned to the draws the
    // reference actually consumes, in its order: 0,1 pixel jitter (renderer.cpp:261-262); per bounce i:
    // 2+4i+{0,1} scatter draw (renderer.cpp:81,94), 2+4i+{2,3} sun cone draw (sampling.h:76).
    // Draws whose value the reference never uses (metal's hemp_cos, process_hit on a shadow hit)
    // are not assigned dimensions.
    inline uint32_t mix32(uint32_t x)
    {
        x ^= x >> 16;
        x *= 0x7feb352du;
        x ^= x >> 15;
        x *= 0x846ca68bu;
        x ^= x >> 16;
        return x;
    }
    inline uint64_t path_key(uint32_t seed, uint32_t pixel, uint32_t sample)
    {
        const uint32_t k1 = mix32(mix32(mix32(seed + 0x9e3779b9u) ^ pixel) ^ sample);
        const uint32_t k2 = mix32(mix32(mix32(seed + 0x85ebca6bu) ^ sample) ^ pixel);
        return (uint64_t(k2) << 32) | k1;
    }
    inline float rnd(uint64_t key, uint32_t dim)
    {
        return float(mix32(mix32(uint32_t(key) + dim * 0x9e3779b9u) ^ uint32_t(key >> 32)) >> 8) * (1.0f / 16777216.0f);
    }

    // ---------------------------------------------------------------- sampling.h
    struct local_coords
    {
        vec3 normal, tangent, bi_tangent;
    };
    // src/util/sampling.h:21-33
    local_coords build_local(vec3 n)
    {
        local_coords c;
        const float  s = (n.z < 0.0) ? -1.0f : 1.0f;
        const float  a = -1.0f / (s + n.z);
        const float  b = n.x * n.y * a;
        c.normal       = n;
        c.tangent      = V3(1.0f + s * n.x * n.x * a, s * b, -s * n.x);
        c.bi_tangent   = V3(b, s + n.y * n.y * a, -n.y);
        return c;
    }
    // src/util/sampling.h:35-42
    vec3 map_to_solid_angle(float ux, float uy, float theta_max)
    {
        const float phi       = TAU * ux;
        const float cos_theta = 1.0f - uy * (1.0f - std::cos(theta_max));
        const float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
        return V3(std::cos(phi) * sin_theta, cos_theta, std::sin(phi) * sin_theta);
    }
    // src/util/sampling.h:44-47
    float solid_angle_mapping_pdf(float theta_max) { return 1.0f / (TAU * (1.0f - std::cos(theta_max))); }
    // src/util/sampling.h:156-166
    vec3 sphere(float ux, float uy)
    {
        const float cos_theta = 2.0f * ux - 1.0f;
        const float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
        const float phi       = TAU * uy;
        const float sin_phi   = std::sin(phi);
        const float cos_phi   = std::cos(phi);
        return V3(sin_theta * cos_phi, cos_theta, sin_theta * sin_phi);
    }
    // src/util/sampling.h:168-172
    vec3 hemp_cos(vec3 normal, float ux, float uy) { return normal + sphere(ux, uy); }

    struct mat3
    {
        vec3 c0, c1, c2;
    };
    inline vec3 mul(const mat3 &M, vec3 v) { return (M.c0 * v.x + M.c1 * v.y) + M.c2 * v.z; }
    // registry.cpp:44-48, 248-256: mat3(tangent, normal, bitangent) of -sun.direction
    mat3 sun_transform_of(vec3 sun_dir)
    {
        local_coords lc = build_local(-sun_dir);
        return mat3 { lc.tangent, lc.normal, lc.bi_tangent };
    }

    // ---------------------------------------------------------------- image.h
    struct image
    {
        uint32_t           w = 0, h = 0;
        std::vector<float> px;
        // src/objects/image.h:104-121. static_cast<uint64_t>(float) of a negative value is UB in
        // C++; GCC/x86-64 lowers it through the signed conversion for |x| < 2^63, which is restated
        // here explicitly.
        static uint64_t to_u64(float f)
        {
            if (!(f == f)) return 0x8000000000000000ull;    // cvttss2si "indefinite"
            if (f >= 9223372036854775808.0f) return uint64_t(f);
            return uint64_t(int64_t(f));
        }
        vec4 get_uv(float u, float v) const
        {
            const uint64_t x = to_u64(u * float(uint64_t(w))) % w;
            const uint64_t y = to_u64(v * float(uint64_t(h))) % h;
            const size_t   b = (x + y * w) * 4;
            return vec4 { px[b], px[b + 1], px[b + 2], px[b + 3] };
        }
    };

    // ---------------------------------------------------------------- geometry + BVH
    struct tri_hit
    {
        float    t = INF, u = 0, v = 0;
        uint32_t prim = 0xffffffffu;
    };

    // Moeller-Trumbore, one IEEE rounding per operation, in this exact order (the product's
    // device triangle test follows the same sequence so t,u,v agree bit for bit):
    //   e1=v1-v0, e2=v2-v0, p=d x e2, det=e1.p, inv=1/det, s=o-v0, u=(s.p)*inv, q=s x e1,
    //   v=(d.q)*inv, t=(e2.q)*inv; accept iff 0<=u<=1, v>=0, u+v<=1, tnear < t <= tfar.
    inline bool tri_test(vec3 v0, vec3 e1, vec3 e2, vec3 o, vec3 d, float tnear, float tfar, float &t, float &u, float &v)
    {
        const vec3  p   = cross(d, e2);
        const float det = dot(e1, p);
        if (det == 0.0f) return false;
        const float inv = 1.0f / det;
        const vec3  s   = o - v0;
        u               = dot(s, p) * inv;
        if (!(u >= 0.0f && u <= 1.0f)) return false;
        const vec3 q = cross(s, e1);
        v            = dot(d, q) * inv;
        if (!(v >= 0.0f && u + v <= 1.0f)) return false;
        t = dot(e2, q) * inv;
        return t > tnear && t <= tfar;
    }

    struct aabb
    {
        vec3 lo = V3(INF, INF, INF), hi = V3(-INF, -INF, -INF);
        void grow(vec3 p)
        {
            lo = V3(std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z));
            hi = V3(std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z));
        }
        void  grow(const aabb &b) { grow(b.lo), grow(b.hi); }
        float area() const
        {
            vec3 e = hi - lo;
            if (e.x < 0) return 0;
            return 2.0f * (e.x * e.y + e.y * e.z + e.z * e.x);
        }
    };
    inline float axis(vec3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

    struct bvh_node
    {
        aabb     box;
        uint32_t left  = 0;    // internal: left child index (right = left+1); leaf: first tri
        uint32_t count = 0;    // 0 = internal
        uint32_t axis  = 0;
    };

    struct model
    {
        uint32_t                  ntris = 0;
        std::vector<vec3>         verts;      // 3 per tri (registry.cpp:65)
        std::vector<vec2>         uvs;        // 3 per tri (registry.cpp:66)
        std::vector<uint32_t>     mat_idx;    // per tri (registry.cpp:92)
        std::vector<orc_material> materials;
        std::vector<mat4>         transforms { identity4() };    // registry.cpp:73-74
        std::vector<mat4>         inv_transforms { identity4() };
        // BVH (stands in for the Embree scene, components.h:65-70)
        std::vector<bvh_node> nodes;
        std::vector<uint32_t> order;             // bvh slot -> prim
        std::vector<vec3>     bv0, be1, be2;     // reordered triangle data

        void build();
        void intersect(vec3 o, vec3 d, float tnear, float tfar, tri_hit &best) const;
        void brute(vec3 o, vec3 d, float tnear, float tfar, tri_hit &best) const;
    };

    void model::build()
    {
        const uint32_t n = ntris;
        nodes.clear();
        order.resize(n);
        std::vector<aabb> tb(n);
        std::vector<vec3> cen(n);
        for (uint32_t i = 0; i < n; i++)
        {
            order[i] = i;
            tb[i].grow(verts[3 * i]), tb[i].grow(verts[3 * i + 1]), tb[i].grow(verts[3 * i + 2]);
            cen[i] = (tb[i].lo + tb[i].hi) * 0.5f;
        }
        nodes.reserve(2 * size_t(n) + 1);
        nodes.emplace_back();
        struct job
        {
            uint32_t node, first, count;
        };
        std::vector<job> stack { { 0, 0, n } };
        constexpr int    NB = 16;
        while (!stack.empty())
        {
            job j = stack.back();
            stack.pop_back();
            aabb box, cb;
            for (uint32_t i = j.first; i < j.first + j.count; i++) box.grow(tb[order[i]]), cb.grow(cen[order[i]]);
            nodes[j.node].box = box;
            auto make_leaf    = [&] { nodes[j.node].left = j.first, nodes[j.node].count = j.count; };
            if (j.count <= 2)
            {
                make_leaf();
                continue;
            }
            float best_cost = INF;
            int   best_axis = -1, best_bin = -1;
            for (int a = 0; a < 3; a++)
            {
                const float lo = axis(cb.lo, a), hi = axis(cb.hi, a);
                if (!(hi > lo)) continue;
                aabb        bb[NB];
                uint32_t    bc[NB] = {};
                const float k      = float(NB) * (1.0f - 1e-6f) / (hi - lo);
                for (uint32_t i = j.first; i < j.first + j.count; i++)
                {
                    int b = std::min(NB - 1, std::max(0, int((axis(cen[order[i]], a) - lo) * k)));
                    bb[b].grow(tb[order[i]]);
                    bc[b]++;
                }
                float    right_area[NB];
                aabb     acc;
                uint32_t cnt = 0;
                for (int b = NB - 1; b > 0; b--) acc.grow(bb[b]), right_area[b] = acc.area();
                acc = aabb();
                uint32_t right_cnt[NB];
                for (int b = NB - 1; b > 0; b--) cnt += bc[b], right_cnt[b] = cnt;
                cnt = 0;
                for (int b = 0; b < NB - 1; b++)
                {
                    acc.grow(bb[b]);
                    cnt += bc[b];
                    if (cnt == 0 || right_cnt[b + 1] == 0) continue;
                    const float c = acc.area() * float(cnt) + right_area[b + 1] * float(right_cnt[b + 1]);
                    if (c < best_cost) best_cost = c, best_axis = a, best_bin = b;
                }
            }
            uint32_t mid;
            if (best_axis < 0)
            {
                if (j.count <= 4)
                {
                    make_leaf();
                    continue;
                }
                mid = j.first + j.count / 2;    // all centroids coincide: median split by index
            }
            else
            {
                const float leaf_cost = box.area() * float(j.count);
                if (j.count <= 4 && leaf_cost <= best_cost + box.area() * 1.0f)
                {
                    make_leaf();
                    continue;
                }
                const float lo = axis(cb.lo, best_axis), hi = axis(cb.hi, best_axis);
                const float k  = float(NB) * (1.0f - 1e-6f) / (hi - lo);
                auto        it = std::partition(order.begin() + j.first, order.begin() + j.first + j.count, [&](uint32_t p) {
                    int b = std::min(NB - 1, std::max(0, int((axis(cen[p], best_axis) - lo) * k)));
                    return b <= best_bin;
                });
                mid            = uint32_t(it - order.begin());
                if (mid == j.first || mid == j.first + j.count) mid = j.first + j.count / 2;
            }
            const uint32_t l    = uint32_t(nodes.size());
            nodes[j.node].left  = l;
            nodes[j.node].count = 0;
            nodes[j.node].axis  = best_axis < 0 ? 0 : uint32_t(best_axis);
            nodes.emplace_back();
            nodes.emplace_back();
            stack.push_back({ l, j.first, mid - j.first });
            stack.push_back({ l + 1, mid, j.first + j.count - mid });
        }
        bv0.resize(n), be1.resize(n), be2.resize(n);
        for (uint32_t i = 0; i < n; i++)
        {
            const uint32_t p = order[i];
            bv0[i]           = verts[3 * p];
            be1[i]           = verts[3 * p + 1] - verts[3 * p];
            be2[i]           = verts[3 * p + 2] - verts[3 * p];
        }
    }

    inline float safe_inv(float d)
    {
        const float tiny = 1e-20f;
        if (std::fabs(d) > tiny) return 1.0f / d;
        return 1.0f / std::copysign(tiny, d);
    }

    // conservative slab test: culls only boxes that cannot contain an accepted hit
    inline bool box_test(const aabb &b, vec3 o, vec3 id, float tnear, float tbest, float &tentry)
    {
        float t1 = (b.lo.x - o.x) * id.x, t2 = (b.hi.x - o.x) * id.x;
        float tmin = std::min(t1, t2), tmax = std::max(t1, t2);
        t1 = (b.lo.y - o.y) * id.y, t2 = (b.hi.y - o.y) * id.y;
        tmin = std::max(tmin, std::min(t1, t2)), tmax = std::min(tmax, std::max(t1, t2));
        t1 = (b.lo.z - o.z) * id.z, t2 = (b.hi.z - o.z) * id.z;
        tmin = std::max(tmin, std::min(t1, t2)), tmax = std::min(tmax, std::max(t1, t2));
        tmax *= 1.0000005f;
        tmin *= 0.9999995f;
        tentry = tmin;
        return tmin <= tmax && tmax >= tnear * 0.5f && tmin <= tbest;
    }

    // closest hit; ties in t resolved to the lowest prim id so the result is BVH-independent
    void model::intersect(vec3 o, vec3 d, float tnear, float tfar, tri_hit &best) const
    {
        if (ntris == 0) return;
        const vec3 id = V3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
        uint32_t   stack[128];
        int        sp    = 0;
        float      limit = std::min(best.t, tfar);
        float      te;
        if (!box_test(nodes[0].box, o, id, tnear, limit, te)) return;
        stack[sp++] = 0;
        while (sp)
        {
            const bvh_node &n = nodes[stack[--sp]];
            if (n.count)
            {
                for (uint32_t i = n.left; i < n.left + n.count; i++)
                {
                    float t, u, v;
                    if (tri_test(bv0[i], be1[i], be2[i], o, d, tnear, limit, t, u, v))
                    {
                        const uint32_t p = order[i];
                        if (t < best.t || (t == best.t && p < best.prim)) best.t = t, best.u = u, best.v = v, best.prim = p, limit = t;
                    }
                }
                continue;
            }
            float      ta, tb2;
            const bool ha = box_test(nodes[n.left].box, o, id, tnear, limit, ta);
            const bool hb = box_test(nodes[n.left + 1].box, o, id, tnear, limit, tb2);
            if (ha && hb)
            {
                if (ta <= tb2)
                    stack[sp++] = n.left + 1, stack[sp++] = n.left;
                else
                    stack[sp++] = n.left, stack[sp++] = n.left + 1;
            }
            else if (ha)
                stack[sp++] = n.left;
            else if (hb)
                stack[sp++] = n.left + 1;
        }
    }

    void model::brute(vec3 o, vec3 d, float tnear, float tfar, tri_hit &best) const
    {
        float limit = std::min(best.t, tfar);
        for (uint32_t p = 0; p < ntris; p++)
        {
            float      t, u, v;
            const vec3 v0 = verts[3 * p], e1 = verts[3 * p + 1] - v0, e2 = verts[3 * p + 2] - v0;
            if (tri_test(v0, e1, e2, o, d, tnear, limit, t, u, v))
                if (t < best.t || (t == best.t && p < best.prim)) best.t = t, best.u = u, best.v = v, best.prim = p, limit = t;
        }
    }

    // ---------------------------------------------------------------- scene
    // cr::ray::intersection_record, src/render/ray.h:15-23 (prim_id is never written by the
    // reference, model.cpp:29-48; kept here for the primary-hit parity check)
    struct record
    {
        float               distance = INF;
        const orc_material *material = nullptr;
        vec2                uv;
        vec3                normal;
        vec3                point;
        uint32_t            prim = 0xffffffffu, model = 0xffffffffu, inst = 0;
        float               t = INF, bu = 0, bv = 0;
    };
    struct ray
    {
        vec3 origin, direction;
        vec3 at(float t) const { return origin + direction * t; }    // ray.cpp:13-16
    };

    struct camera
    {
        orc_camera p {};
        mat4       M = identity4();
        camera()
        {
            // camera.h:21 defaults
            p.position[0] = 5, p.position[1] = 5, p.position[2] = 0;
            p.fov = 75, p.scale = 1, p.mode = 0;
        }
        // glm::rotate(m, angle, axis) for a unit axis, glm/ext/matrix_transform.inl
        static mat4 rotate(const mat4 &m, float angle, vec3 ax)
        {
            const float c = std::cos(angle), s = std::sin(angle);
            const vec3  t = ax * (1.0f - c);
            float       R[3][3];
            R[0][0] = c + t.x * ax.x, R[0][1] = t.x * ax.y + s * ax.z, R[0][2] = t.x * ax.z - s * ax.y;
            R[1][0] = t.y * ax.x - s * ax.z, R[1][1] = c + t.y * ax.y, R[1][2] = t.y * ax.z + s * ax.x;
            R[2][0] = t.z * ax.x + s * ax.y, R[2][1] = t.z * ax.y - s * ax.x, R[2][2] = c + t.z * ax.z;
            mat4 r;
            for (int k = 0; k < 3; k++)
                for (int row = 0; row < 4; row++) r.m[k][row] = (m.m[0][row] * R[k][0] + m.m[1][row] * R[k][1]) + m.m[2][row] * R[k][2];
            for (int row = 0; row < 4; row++) r.m[3][row] = m.m[3][row];
            return r;
        }
        // camera.cpp:54-68: T(pos) * Ry(rot.x) * Rx(rot.y) * Rz(rot.z), degrees
        void update_cache()
        {
            mat4 m    = identity4();
            m.m[3][0] = p.position[0], m.m[3][1] = p.position[1], m.m[3][2] = p.position[2];
            const float rad = 0.01745329251994329576923690768489f;
            m               = rotate(m, p.rotation[0] * rad, V3(0, 1, 0));
            m               = rotate(m, p.rotation[1] * rad, V3(1, 0, 0));
            m               = rotate(m, p.rotation[2] * rad, V3(0, 0, 1));
            M               = m;
        }
        // camera.cpp:14-39
        ray get_ray(float x, float y, float aspect) const
        {
            if (p.mode == 0)
            {
                const float u   = (2.0f * x - 1.0f) * aspect;
                const float v   = 2.0f * y - 1.0f;
                const float w   = 1.0f / std::tan(0.5f * (p.fov * 0.01745329251994329576923690768489f));
                const vec3  dir = mul_dir(M, V3(u, v, w));
                return ray { V3(p.position[0], p.position[1], p.position[2]), normalize(dir) };
            }
            const float u = 2.0f * x - 1.0f;
            const float v = 2.0f * y - 1.0f;
            const vec3  o = mul_point(M, V3(p.scale * u, p.scale * v, 0.0f));
            const vec3  d = V3(M.m[2][0], M.m[2][1], M.m[2][2]);
            return ray { o, normalize(d) };
        }
    };

    struct scene
    {
        std::vector<model> models;
        std::vector<image> textures;
        bool               sun_enabled = true;    // scene.h:45
        orc_sun            sun {};
        mat3               sun_transform;
        bool               has_skybox = false;
        image              skybox;
        vec2               skybox_rot;
        camera             cam;

        scene()
        {
            // components.h:23-29
            sun.size      = PI / 48.0f;
            sun.intensity = 100.0f;
            vec3 d        = normalize(V3(0.8f, -1.0f, 0.0f));
            sun.direction[0] = d.x, sun.direction[1] = d.y, sun.direction[2] = d.z;
            sun.colour[0] = 1.0f, sun.colour[1] = 0.9f, sun.colour[2] = 0.7f;
            sun_transform = sun_transform_of(d);
        }

        // scene.cpp:67-77
        vec3 sample_skybox(float x, float y) const
        {
            if (has_skybox)
            {
                vec4 c = skybox.get_uv(x + skybox_rot.x, y + skybox_rot.y);
                return V3(c.x, c.y, c.z);
            }
            return V3(0, 0, 0);
        }

        // model.cpp:5-49 (_intersect): tnear=1e-5, tfar=inf; normal = normalize(Ng);
        // material by primID; uv by rtcInterpolate0 = (1-u-v)*t0 + u*t1 + v*t2
        record intersect_model(const model &m, uint32_t mi, const ray &r, bool brute) const
        {
            tri_hit h;
            if (brute)
                m.brute(r.origin, r.direction, 0.00001f, INF, h);
            else
                m.intersect(r.origin, r.direction, 0.00001f, INF, h);
            record rec;
            if (h.prim == 0xffffffffu) return rec;
            rec.distance  = h.t;
            rec.point     = r.at(h.t);
            const vec3 v0 = m.verts[3 * h.prim], v1 = m.verts[3 * h.prim + 1], v2 = m.verts[3 * h.prim + 2];
            rec.normal    = normalize(cross(v1 - v0, v2 - v0));
            rec.material  = &m.materials[m.mat_idx[h.prim]];
            if (!m.uvs.empty())
            {
                const vec2  a = m.uvs[3 * h.prim], b = m.uvs[3 * h.prim + 1], c = m.uvs[3 * h.prim + 2];
                const float w = 1.0f - h.u - h.v;
                rec.uv.x      = (w * a.x + h.u * b.x) + h.v * c.x;
                rec.uv.y      = (w * a.y + h.u * b.y) + h.v * c.y;
            }
            rec.prim = h.prim, rec.model = mi, rec.t = h.t, rec.bu = h.u, rec.bv = h.v;
            return rec;
        }

        // model.cpp:99-126 (cr::model::intersect) inside scene.cpp:79-98 (cr::scene::cast_ray)
        record cast_ray(const ray &r, bool brute = false) const
        {
            record best;
            for (uint32_t mi = 0; mi < models.size(); mi++)
            {
                const model &m = models[mi];
                record       mb;
                for (uint32_t ii = 0; ii < m.transforms.size(); ii++)
                {
                    const mat4 &inv = m.inv_transforms[ii];
                    const ray   tr { mul_point(inv, r.origin), normalize(mul_dir(inv, r.direction)) };
                    record      cur = intersect_model(m, mi, tr, brute);
                    cur.point       = mul_point(m.transforms[ii], cur.point);
                    cur.inst        = ii;
                    if (cur.distance != INF) cur.distance = length(cur.point - r.origin);
                    if (cur.distance < mb.distance) mb = cur;
                }
                if (mb.distance < best.distance) best = mb;
            }
            return best;
        }
    };

    // ---------------------------------------------------------------- renderer
    struct processed_hit
    {
        bool  is_alpha = false;
        float emission = 0;
        vec3  albedo;
        vec4  colour;
        ray   r;
    };

    // renderer.cpp:21-102. u0,u1 are the two scatter draws (consumed by smooth; drawn and
    // discarded by metal; not drawn by glass).
    template<typename Draw2>
    processed_hit process_hit_d(const record &rec, const ray &r, const scene &sc, Draw2 &&draw2)
    {
        processed_hit       out;
        const orc_material &mat = *rec.material;
        out.emission            = mat.emission;
        if (mat.tex >= 0)
            out.colour = sc.textures[size_t(mat.tex)].get_uv(rec.uv.x, rec.uv.y);
        else
            out.colour = vec4 { mat.colour[0], mat.colour[1], mat.colour[2], mat.colour[3] };
        if (out.colour.w == 0.0)
        {
            out.is_alpha = true;
            return out;
        }
        out.albedo = V3(out.colour.x, out.colour.y, out.colour.z);
        switch (mat.shade_type)
        {
        case ORC_GLASS:
        {
            vec3  refracted;
            vec3  out_normal = rec.normal;
            vec3  reflected  = reflect(r.direction, rec.normal);
            float ni_over_nt = 1.0f / mat.ior;
            if (dot(r.direction, rec.normal) > 0) out_normal = -rec.normal, ni_over_nt = mat.ior;
            const vec3  uv   = normalize(r.direction);
            const float dt   = dot(uv, out_normal);
            const float disc = 1.0f - ni_over_nt * ni_over_nt * (1 - dt * dt);
            bool        refract = false;
            if (disc > 0)
            {
                refracted = ni_over_nt * (uv - out_normal * dt) - out_normal * std::sqrt(disc);
                refract   = true;
            }
            out.r.origin    = rec.point + out_normal * -0.0001f;
            out.r.direction = refract ? refracted : reflected;
        }
        break;
        case ORC_METAL:
        {
            out.r.origin    = rec.point + rec.normal * 0.0001f;
            (void) draw2();    // renderer.cpp:81: hemp_cos(record.normal, vec2(randf(), randf())) is computed and never used
            out.r.direction = reflect(r.direction, rec.normal);
            out.albedo      = out.albedo * mat.reflectiveness;
            break;
        }
        default:    // smooth
        {
            const vec2 uu   = draw2();
            const float u0 = uu.x, u1 = uu.y;
            const vec3 h    = hemp_cos(rec.normal, u0, u1);
            out.r.origin    = rec.point + rec.normal * 0.0001f;
            out.r.direction = normalize(h);
            break;
        }
        }
        return out;
    }

    inline processed_hit process_hit(const record &rec, const ray &r, const scene &sc, float u0, float u1)
    {
        return process_hit_d(rec, r, sc, [&] { return vec2 { u0, u1 }; });
    }

    // sampling.h:53-57
    vec3 sky_colour(vec3 direction, const orc_sun &sun)
    {
        const vec3  sd        = V3(sun.direction[0], sun.direction[1], sun.direction[2]);
        const float sun_angle = std::acos(dot(direction, -sd));
        return (sun_angle < sun.size) ? V3(sun.colour[0], sun.colour[1], sun.colour[2]) * sun.intensity : V3(0, 0, 0);
    }

    // ================================================================ EXTENDED shading mode
    // north_star / BASELINE configs 2 and 5 name "Lambert / GGX metal / Fresnel dielectric with NEE shadow
    // rays" and "64 area lights with NEE". In the reference those are DEAD code (SURVEY.md D5:
    // cr::brdf::ggx, src/render/brdf.h:10-29, and cook_torrence::*, src/util/sampling.h:83-142, are never
    // called; there is no area-light NEE). The extended mode is therefore SPECIFIED HERE, by the oracle
    // ("parity unpinned" by the reference; the product is checked against this restatement only). It keeps
    // everything of the ref-exact path (camera, sampler, cast_ray, alpha skip, AOVs, accumulation) and
    // replaces process_hit + NEE by:
    //   * shading normal ns = geometric normal flipped against the ray (the reference never face-forwards);
    //   * smooth: Lambert, wi = normalize(ns + sphere(u0,u1)) (sampling.h:156-172), weight = colour;
    //   * metal:  GGX microfacet reflection with the reference's own G term (sampling.h:110-118; D of :94-100 cancels against the sampling pdf; a =
    //             clamp(roughness, 0.02, 1)), Schlick Fresnel with f0 = colour * reflectiveness
    //             (sampling.h:137-141 per channel); half vector sampled from D*cos, weight = F*Vis*4*VoH*NoL/NoH;
    //   * glass:  exact unpolarised Fresnel reflectance R on top of the reference's refraction arithmetic
    //             (renderer.cpp:47-77); reflect with probability R (u0 < R), else refract; weight = colour;
    //   * emission: Le = material colour * emission, added on a hit only if the previous vertex was
    //             specular / the camera, or if the scene has no light list (no double counting with NEE);
    //   * NEE at smooth hits only: the sun cone (sampling.h:35-57,72-80) and/or ONE uniformly picked emissive
    //             triangle (uniform point, two-sided emitter); when both exist one of the two strategies is
    //             picked with probability 1/2. BSDF = colour/pi. Shadow ray from p + ns*1e-3, any hit in
    //             (1e-5, tmax], tmax = 0.999*distance for area lights, inf for the sun; alpha cut-outs are
    //             marched through in 0.1 steps like the reference's shadow loop (renderer.cpp:335-345);
    //   * a specular or camera path that escapes also sees the sun disc (sky_colour, sampling.h:53-57).
    // Sampler dimensions per bounce i: 2+6i+{0,1} scatter, {2,3} light sample, {4} strategy, {5} light pick.
    struct area_light
    {
        vec3 v0, e1, e2, le;
    };

    inline float pow5(float m)
    {
        const float m2 = m * m;
        return (m2 * m2) * m;
    }
    // sampling.h:110-118
    inline float specular_g(float NoV, float NoL, float a)
    {
        const float a2   = a * a;
        const float ggxv = NoL * std::sqrt(NoV * NoV * (1.0f - a2) + a2);
        const float ggxl = NoV * std::sqrt(NoL * NoL * (1.0f - a2) + a2);
        return 0.5f / (ggxv + ggxl);
    }
    // half vector distributed like D(h) * (n.h), in the frame of build_local (y = normal)
    inline vec3 sample_ggx_h(vec3 ns, float a, float u0, float u1, float &cos_h)
    {
        const float a2   = a * a;
        const float cos2 = (1.0f - u0) / (1.0f + (a2 - 1.0f) * u0);
        cos_h            = std::sqrt(cos2);
        const float sin_h = std::sqrt(std::max(0.0f, 1.0f - cos2));
        const float phi   = TAU * u1;
        const local_coords lc = build_local(ns);
        return (lc.tangent * (std::cos(phi) * sin_h) + lc.normal * cos_h) + lc.bi_tangent * (std::sin(phi) * sin_h);
    }
    struct ext_scatter
    {
        bool absorbed = false, specular = false;
        vec3 weight;
        ray  r;
    };
    ext_scatter scatter_extended(const orc_material &mat, vec3 colour, vec3 n, vec3 point, const ray &r, float u0, float u1)
    {
        ext_scatter out;
        const vec3  dn = normalize(r.direction);
        const vec3  ns = (dot(r.direction, n) > 0) ? -n : n;
        switch (mat.shade_type)
        {
        case ORC_GLASS:
        {
            const float eta  = (dot(r.direction, n) > 0) ? mat.ior : 1.0f / mat.ior;    // renderer.cpp:52-57
            const float dt   = dot(dn, ns);
            const float disc = 1.0f - eta * eta * (1 - dt * dt);
            float       R    = 1.0f;
            if (disc > 0)
            {
                const float cos_i = -dt, cos_t = std::sqrt(disc);
                const float rs = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
                const float rp = (cos_i - eta * cos_t) / (cos_i + eta * cos_t);
                R              = 0.5f * (rs * rs + rp * rp);
            }
            if (u0 < R)
            {
                out.r.origin    = point + ns * 0.0001f;
                out.r.direction = reflect(dn, ns);
            }
            else
            {
                out.r.origin    = point + ns * -0.0001f;
                out.r.direction = eta * (dn - ns * dt) - ns * std::sqrt(disc);    // renderer.cpp:63-67
            }
            out.weight   = colour;
            out.specular = true;
            break;
        }
        case ORC_METAL:
        {
            const float a = clampf(mat.roughness, 0.02f, 1.0f);
            float       NoH;
            const vec3  h   = sample_ggx_h(ns, a, u0, u1, NoH);
            const vec3  wi  = reflect(dn, h);
            const float VoH = -dot(dn, h), NoL = dot(ns, wi), NoV = -dot(dn, ns);
            out.specular    = true;
            if (!(VoH > 0 && NoL > 0 && NoV > 0))
            {
                out.absorbed = true;
                break;
            }
            const vec3  f0 = colour * mat.reflectiveness;
            const float f  = pow5(1.0f - VoH);
            const vec3  F  = V3(f + f0.x * (1.0f - f), f + f0.y * (1.0f - f), f + f0.z * (1.0f - f));    // sampling.h:137-141
            const float g  = specular_g(NoV, NoL, a) * 4.0f * VoH * NoL / NoH;
            out.weight      = F * g;
            out.r.origin    = point + ns * 0.0001f;
            out.r.direction = wi;
            break;
        }
        default:
        {
            out.r.origin    = point + ns * 0.0001f;
            out.r.direction = normalize(hemp_cos(ns, u0, u1));
            out.weight      = colour;
            break;
        }
        }
        return out;
    }

    struct render
    {
        scene             *sc;
        uint32_t           w, h, max_bounces, seed;
        float              aspect;
        uint32_t           row0, row1;
        std::vector<float> raw;                                // renderer.h:85 float[W*H*3]
        std::vector<float> buffer, normals, albedo, depth;     // renderer.h:87-91 RGBA f32
        uint32_t           current_sample = 0;
        std::atomic<uint64_t> total_queries { 0 }, ref_rays { 0 }, pixel_samples { 0 };
        // sampler_mode 1 = the REFERENCE's own stream (renderer.cpp:6-11): one default-seeded std::mt19937 consumed in call
        // order by a single worker thread. Used only to compare this restatement with the reference's sources compiled
        // under oracle/ref (tests/test_reference_anchor.py); needs nthreads == 1. Where the reference leaves the order
        // of two randf() calls to the compiler (function / constructor arguments), the order is the one g++ produces:
        // right to left, i.e. the SECOND argument gets the first draw.
        int          sampler_mode = 0;
        std::mt19937 mt;
        // sample table: in reference-stream mode the draws that feed the estimator are recorded by (sample, pixel,
        // dimension) — dimensions as in the counter-based sampler: 0,1 jitter; 2+4i+{0,1} scatter of bounce i;
        // 2+4i+{2,3} sun sample of bounce i — so that the CUDA path can replay the reference's own numbers
        // (crb_render_set_sample_table). sampler_mode 2 reads the same table back.
        float   *table      = nullptr;
        uint32_t table_dims = 0, table_samples = 0;
        vec2     table_pair(uint32_t sample, uint64_t pixel, uint32_t dim, vec2 drawn)
        {
            if (!table || dim + 1 >= table_dims || sample >= table_samples) return drawn;
            float *e = table + (size_t(sample) * w * h + pixel) * table_dims + dim;
            if (sampler_mode == 2) return vec2 { e[0], e[1] };
            e[0] = drawn.x, e[1] = drawn.y;
            return drawn;
        }
        float        mt_randf()
        {
            std::uniform_real_distribution<float> dist(0.f, 1.f);
            return dist(mt);
        }
        vec2 mt_pair()
        {
            vec2 p;
            p.y = mt_randf();
            p.x = mt_randf();
            return p;
        }
        bool                    extended = false;    // see "EXTENDED shading mode" above
        bool                    light_nee = true;    // false: leave the light list empty (emitters found by hits only; estimator cross-check)
        std::vector<area_light> lights;              // emissive triangles in world space, (model, instance, triangle) order

        render(scene *s, uint32_t w_, uint32_t h_, uint32_t mb, uint32_t seed_) : sc(s), w(w_), h(h_), max_bounces(mb), seed(seed_)
        {
            // renderer.cpp:199 (set_resolution); the ctor leaves it at 1 (renderer.h:80), the UI always
            // goes through set_resolution, and so does the product's boundary.
            aspect = float(w) / float(h);
            row0 = 0, row1 = h;
            reset();
        }
        void reset()
        {
            // renderer.cpp:154-170 and image.h:30-33 (images are FLT_MAX-filled)
            raw.assign(size_t(w) * h * 3, 0.0f);
            const float mx = std::numeric_limits<float>::max();
            buffer.assign(size_t(w) * h * 4, mx), normals.assign(size_t(w) * h * 4, mx);
            albedo.assign(size_t(w) * h * 4, mx), depth.assign(size_t(w) * h * 4, mx);
            current_sample = 0;
            total_queries = 0, ref_rays = 0, pixel_samples = 0;
        }
        static void set_px(std::vector<float> &img, uint32_t w, uint64_t x, uint64_t y, vec3 c)
        {
            const size_t b = (x + y * w) * 4;    // image.h:137-145
            img[b] = c.x, img[b + 1] = c.y, img[b + 2] = c.z, img[b + 3] = 1.0f;
        }

        // renderer.cpp:258-384
        void sample_pixel(uint64_t x, uint64_t y, uint32_t sample, uint64_t &queries, uint64_t &fired, record *primary_out)
        {
            const uint64_t key = path_key(seed, uint32_t(x + y * w), sample);
            const bool     ref_stream = sampler_mode == 1;
            const uint64_t pixel      = x + y * w;
            auto           draw       = [&](uint32_t dim) {
                const vec2 d = ref_stream ? mt_pair() : vec2 { rnd(key, dim), rnd(key, dim + 1) };
                return sampler_mode ? table_pair(sample, pixel, dim, d) : d;
            };
            const vec2 jit = draw(0);
            ray r = sc->cam.get_ray((float(x) + jit.x) / float(uint64_t(w)), (float(y) + jit.y) / float(uint64_t(h)), aspect);

            vec3  throughput = V3(1, 1, 1), final = V3(0, 0, 0), albedo_ = V3(0, 0, 0), normal_ = V3(0, 0, 0);
            float depth_     = 0.0f;

            int total_bounces = 1;
            for (uint32_t i = 0; i < max_bounces; i++, total_bounces++)
            {
                record isect = sc->cast_ray(r);
                queries++;
                if (i == 0 && primary_out)
                {
                    *primary_out = isect;
                    return;
                }
                processed_hit ph;
                if (isect.distance == INF)
                {
                    const float mu = 0.5f + std::atan2(r.direction.z, r.direction.x) * INV_TAU;
                    const float mv = 0.5f - std::asin(r.direction.y) * INV_PI;
                    const vec3  ms = sc->sample_skybox(mu, mv);
                    if (i == 0) albedo_ = ms;
                    final = final + throughput * ms;
                    break;
                }
                else
                {
                    ph = process_hit_d(isect, r, *sc, [&] { return draw(2 + 4 * i); });
                    if (ph.is_alpha)
                    {
                        r.origin = isect.point + r.direction * 0.1f;
                        continue;
                    }
                    if (i == 0) albedo_ = ph.albedo, normal_ = isect.normal, depth_ = isect.distance;
                    throughput = throughput * ph.albedo;
                    final      = final + throughput * ph.emission;
                    r          = ph.r;
                }
                // Sun NEE, renderer.cpp:315-354
                if (sc->sun_enabled)
                {
                    ray out_ray { isect.point + isect.normal * 0.001f, V3(0, 0, 0) };
                    // sampling.h:72-80
                    const vec2  su     = draw(2 + 4 * i + 2);
                    const vec3  dir    = mul(sc->sun_transform, map_to_solid_angle(su.x, su.y, sc->sun.size));
                    const float pdf    = solid_angle_mapping_pdf(sc->sun.size);
                    const float cosine = clampf(dot(isect.normal, dir), 0.0f, 1.0f);
                    out_ray.direction  = dir;

                    record sun_isect = sc->cast_ray(out_ray);
                    queries++;
                    if (sun_isect.distance != INF)
                    {
                        // process_hit on the occluder draws (and discards) two numbers for metal / smooth (renderer.cpp:81,93)
                        auto          waste = [&] { return ref_stream ? mt_pair() : vec2 { 0.5f, 0.5f }; };
                        processed_hit pi    = process_hit_d(sun_isect, r, *sc, waste);
                        while (sun_isect.distance != INF && pi.is_alpha)
                        {
                            out_ray.origin = sun_isect.point + out_ray.direction * 0.1f;
                            sun_isect      = sc->cast_ray(out_ray);
                            queries++;
                            if (sun_isect.distance == INF) break;
                            pi = process_hit_d(sun_isect, r, *sc, waste);
                        }
                    }
                    if (sun_isect.distance == INF)
                        final = final + throughput * V3(ph.colour.x, ph.colour.y, ph.colour.z) * cosine * sky_colour(out_ray.direction, sc->sun) / pdf;
                }
            }
            fired += uint64_t(total_bounces);
            write_pixel(x, y, final, albedo_, normal_, depth_);
        }

        void write_pixel(uint64_t x, uint64_t y, vec3 final, vec3 albedo_, vec3 normal_, float depth_)
        {
            // flip, renderer.cpp:358-365
            y = h - 1 - y;
            x = w - 1 - x;
            const size_t base = (x + y * w) * 3;
            raw[base + 0] += final.x;
            raw[base + 1] += final.y;
            raw[base + 2] += final.z;
            set_px(albedo, w, x, y, albedo_);
            set_px(normals, w, x, y, normal_ * .5f + V3(.5f, .5f, .5f));
            const float dd = std::min(depth_, 200.0f) / 200.f;
            set_px(depth, w, x, y, V3(dd, dd, dd));
            // renderer.cpp:371-383
            const float n = float(current_sample + 1);
            set_px(buffer, w, x, y,
                   V3(std::pow(clampf(raw[base + 0] / n, 0.0f, 1.0f), 1.f / 2.2f), std::pow(clampf(raw[base + 1] / n, 0.0f, 1.0f), 1.f / 2.2f),
                      std::pow(clampf(raw[base + 2] / n, 0.0f, 1.0f), 1.f / 2.2f)));
        }

        void build_lights()
        {
            lights.clear();
            if (!light_nee) return;
            for (const model &m : sc->models)
                for (const mat4 &T : m.transforms)
                    for (uint32_t t = 0; t < m.ntris; t++)
                    {
                        const orc_material &mat = m.materials[m.mat_idx[t]];
                        if (!(mat.emission > 0.0f) || mat.colour[3] == 0.0f) continue;
                        const vec3 v0 = mul_point(T, m.verts[3 * t]), v1 = mul_point(T, m.verts[3 * t + 1]), v2 = mul_point(T, m.verts[3 * t + 2]);
                        lights.push_back(area_light { v0, v1 - v0, v2 - v0, V3(mat.colour[0], mat.colour[1], mat.colour[2]) * mat.emission });
                    }
        }

        // any non-alpha hit in (1e-5, tmax] along a unit direction; alpha cut-outs are stepped through
        bool visible_ext(vec3 o, vec3 dir, float tmax, uint64_t &queries)
        {
            ray   sr { o, dir };
            float remaining = tmax;
            for (int guard = 0; guard < 4096; guard++)
            {
                const record h = sc->cast_ray(sr);
                queries++;
                if (h.distance == INF || !(h.distance <= remaining)) return true;
                const orc_material &mat = *h.material;
                const float alpha = (mat.tex >= 0) ? sc->textures[size_t(mat.tex)].get_uv(h.uv.x, h.uv.y).w : mat.colour[3];
                if (alpha != 0.0f) return false;
                const vec3 next = h.point + dir * 0.1f;
                remaining       = remaining - length(next - sr.origin);
                sr.origin       = next;
            }
            return false;
        }

        // the extended-mode path loop (same skeleton as renderer.cpp:258-384)
        void sample_pixel_ext(uint64_t x, uint64_t y, uint32_t sample, uint64_t &queries, uint64_t &fired)
        {
            const uint64_t key = path_key(seed, uint32_t(x + y * w), sample);
            ray r = sc->cam.get_ray((float(x) + rnd(key, 0)) / float(uint64_t(w)), (float(y) + rnd(key, 1)) / float(uint64_t(h)), aspect);

            vec3  throughput = V3(1, 1, 1), final = V3(0, 0, 0), albedo_ = V3(0, 0, 0), normal_ = V3(0, 0, 0);
            float depth_     = 0.0f;
            bool  specular   = true;    // the camera counts as a specular vertex
            const uint32_t NL = uint32_t(lights.size());

            int total_bounces = 1;
            for (uint32_t i = 0; i < max_bounces; i++, total_bounces++)
            {
                const uint32_t dim   = 2 + 6 * i;
                record         isect = sc->cast_ray(r);
                queries++;
                if (isect.distance == INF)
                {
                    const float mu = 0.5f + std::atan2(r.direction.z, r.direction.x) * INV_TAU;
                    const float mv = 0.5f - std::asin(r.direction.y) * INV_PI;
                    vec3        ms = sc->sample_skybox(mu, mv);
                    if (i == 0) albedo_ = ms;
                    if (sc->sun_enabled && specular) ms = ms + sky_colour(normalize(r.direction), sc->sun);
                    final = final + throughput * ms;
                    break;
                }
                const orc_material &mat = *isect.material;
                vec4                col = vec4 { mat.colour[0], mat.colour[1], mat.colour[2], mat.colour[3] };
                if (mat.tex >= 0) col = sc->textures[size_t(mat.tex)].get_uv(isect.uv.x, isect.uv.y);
                if (col.w == 0.0)
                {
                    r.origin = isect.point + r.direction * 0.1f;    // renderer.cpp:294-301
                    continue;
                }
                const vec3 colour = V3(col.x, col.y, col.z);
                if (i == 0) albedo_ = colour, normal_ = isect.normal, depth_ = isect.distance;
                if (mat.emission > 0.0f && (specular || NL == 0))
                    final = final + throughput * (V3(mat.colour[0], mat.colour[1], mat.colour[2]) * mat.emission);

                const ext_scatter sc_out = scatter_extended(mat, colour, isect.normal, isect.point, r, rnd(key, dim), rnd(key, dim + 1));
                const vec3        ns     = (dot(r.direction, isect.normal) > 0) ? -isect.normal : isect.normal;

                // next-event estimation at diffuse vertices
                if (!sc_out.specular && (sc->sun_enabled || NL))
                {
                    const vec3  so   = isect.point + ns * 0.001f;
                    const vec3  bsdf = (throughput * colour) * INV_PI;
                    const bool  both = sc->sun_enabled && NL;
                    const bool  use_sun = sc->sun_enabled && (!NL || rnd(key, dim + 4) < 0.5f);
                    const float u2 = rnd(key, dim + 2), u3 = rnd(key, dim + 3);
                    vec3        contrib = V3(0, 0, 0), dir = V3(0, 1, 0);
                    float       tmax = INF;
                    if (use_sun)
                    {
                        dir                = mul(sc->sun_transform, map_to_solid_angle(u2, u3, sc->sun.size));
                        const float cosine = clampf(dot(ns, dir), 0.0f, 1.0f);
                        contrib            = bsdf * cosine * sky_colour(dir, sc->sun) / solid_angle_mapping_pdf(sc->sun.size);
                    }
                    else
                    {
                        uint32_t k = uint32_t(rnd(key, dim + 5) * float(NL));
                        if (k >= NL) k = NL - 1;
                        const area_light &L = lights[k];
                        float             b1 = u2, b2 = u3;
                        if (b1 + b2 > 1.0f) b1 = 1.0f - b1, b2 = 1.0f - b2;
                        const vec3  q  = (L.v0 + L.e1 * b1) + L.e2 * b2;
                        const vec3  wv = q - so;
                        const float d2 = dot(wv, wv);
                        if (d2 > 0.0f)
                        {
                            const float dist  = std::sqrt(d2);
                            dir               = wv * (1.0f / dist);
                            const float cos_s = clampf(dot(ns, dir), 0.0f, 1.0f);
                            const float g     = cos_s * (0.5f * std::fabs(dot(cross(L.e1, L.e2), dir))) / d2 * float(NL);
                            contrib           = bsdf * L.le * g;
                            tmax              = dist * 0.999f;
                        }
                    }
                    if (both) contrib = contrib * 2.0f;
                    if ((contrib.x != 0.0f || contrib.y != 0.0f || contrib.z != 0.0f) && visible_ext(so, dir, tmax, queries)) final = final + contrib;
                }
                if (sc_out.absorbed) break;
                throughput = throughput * sc_out.weight;
                r          = sc_out.r;
                specular   = sc_out.specular;
            }
            fired += uint64_t(total_bounces);
            write_pixel(x, y, final, albedo_, normal_, depth_);
        }

        // renderer.cpp:116-144, 240-256 + thread_pool.cpp:40-60: one task per scanline per pass,
        // the pass is a barrier.
        void run(uint32_t first_sample, uint32_t n, int nthreads)
        {
            if (nthreads < 1) nthreads = 1;
            if (extended) build_lights();
            for (uint32_t s = 0; s < n; s++)
            {
                const uint32_t        sample = first_sample + s;
                std::atomic<uint32_t> next_row { row0 };
                auto                  worker = [&] {
                    for (;;)
                    {
                        const uint32_t y = next_row.fetch_add(1);
                        if (y >= row1) break;
                        uint64_t q = 0, fired = 0;
                        for (uint32_t x = 0; x < w; x++)
                            if (extended)
                                sample_pixel_ext(x, y, sample, q, fired);
                            else
                                sample_pixel(x, y, sample, q, fired, nullptr);
                        total_queries += q;
                        ref_rays += fired;
                        pixel_samples += w;
                    }
                };
                if (nthreads == 1)
                    worker();
                else
                {
                    std::vector<std::thread> th;
                    for (int t = 0; t < nthreads; t++) th.emplace_back(worker);
                    for (auto &t : th) t.join();
                }
                current_sample++;
            }
        }
    };

    template<typename F>
    void parallel_for(uint64_t n, int nthreads, F f)
    {
        if (nthreads <= 1 || n < 1024)
        {
            for (uint64_t i = 0; i < n; i++) f(i);
            return;
        }
        std::atomic<uint64_t>    next { 0 };
        const uint64_t           chunk = 4096;
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++)
            th.emplace_back([&] {
                for (;;)
                {
                    uint64_t b = next.fetch_add(chunk);
                    if (b >= n) break;
                    uint64_t e = std::min(n, b + chunk);
                    for (uint64_t i = b; i < e; i++) f(i);
                }
            });
        for (auto &t : th) t.join();
    }

    // batch query in WORLD space: t parametrises o + t*d for the caller's (un-normalised) d; nearest
    // over all models and instances, first (model, instance) wins ties.
    orc_hit batch_query(const scene &sc, const orc_ray &q, bool brute)
    {
        orc_hit out { INF, 0, 0, 0xffffffffu, 0xffffffffu, 0 };
        for (uint32_t mi = 0; mi < sc.models.size(); mi++)
        {
            const model &m = sc.models[mi];
            for (uint32_t ii = 0; ii < m.transforms.size(); ii++)
            {
                const mat4 &inv = m.inv_transforms[ii];
                vec3        o = V3(q.o[0], q.o[1], q.o[2]), d = V3(q.d[0], q.d[1], q.d[2]);
                if (!is_identity(m.transforms[ii])) o = mul_point(inv, o), d = mul_dir(inv, d);
                tri_hit h;
                h.t = out.t;
                if (brute)
                    m.brute(o, d, q.tmin, q.tmax, h);
                else
                    m.intersect(o, d, q.tmin, q.tmax, h);
                if (h.prim != 0xffffffffu && h.t < out.t) out = orc_hit { h.t, h.u, h.v, h.prim, mi, ii };
            }
        }
        return out;
    }
}    // namespace

struct orc_scene
{
    scene s;
};
struct orc_render
{
    render r;
    orc_render(scene *s, uint32_t w, uint32_t h, uint32_t mb, uint32_t seed) : r(s, w, h, mb, seed) {}
};

extern "C" {

orc_scene *orc_scene_create(void) { return new orc_scene(); }
void       orc_scene_destroy(orc_scene *s) { delete s; }

int orc_scene_add_mesh(orc_scene *s, const float *verts, const float *uvs, const uint32_t *mat_idx, uint32_t ntris)
{
    model m;
    m.ntris = ntris;
    m.verts.resize(size_t(ntris) * 3);
    std::memcpy(static_cast<void *>(m.verts.data()), verts, sizeof(float) * 9 * size_t(ntris));
    if (uvs)
    {
        m.uvs.resize(size_t(ntris) * 3);
        std::memcpy(static_cast<void *>(m.uvs.data()), uvs, sizeof(float) * 6 * size_t(ntris));
    }
    m.mat_idx.assign(ntris, 0);
    if (mat_idx) std::memcpy(m.mat_idx.data(), mat_idx, sizeof(uint32_t) * size_t(ntris));
    orc_material def {};    // material.h:31-41
    def.shade_type = ORC_SMOOTH, def.ior = 1.5f, def.roughness = 0.5f, def.reflectiveness = 1.0f, def.emission = 0.0f;
    def.colour[0] = def.colour[1] = def.colour[2] = def.colour[3] = 1.0f;
    def.tex = -1;
    uint32_t maxm = 0;
    for (uint32_t i : m.mat_idx) maxm = std::max(maxm, i);
    m.materials.assign(maxm + 1, def);
    s->s.models.push_back(std::move(m));
    return int(s->s.models.size()) - 1;
}

int orc_scene_set_materials(orc_scene *s, int mi, const orc_material *mats, uint32_t n)
{
    if (mi < 0 || size_t(mi) >= s->s.models.size()) return 1;
    model &m = s->s.models[size_t(mi)];
    for (uint32_t i : m.mat_idx)
        if (i >= n) return 1;
    m.materials.assign(mats, mats + n);
    return 0;
}

int orc_scene_set_instances(orc_scene *s, int mi, const float *mats, uint32_t n)
{
    if (mi < 0 || size_t(mi) >= s->s.models.size()) return 1;
    model &m = s->s.models[size_t(mi)];
    m.transforms.resize(n), m.inv_transforms.resize(n);
    for (uint32_t i = 0; i < n; i++)
    {
        std::memcpy(&m.transforms[i], mats + 16 * i, sizeof(mat4));
        m.inv_transforms[i] = is_identity(m.transforms[i]) ? identity4() : inverse(m.transforms[i]);
    }
    return 0;
}

int orc_scene_add_texture(orc_scene *s, const float *rgba, uint32_t w, uint32_t h)
{
    image im;
    im.w = w, im.h = h;
    im.px.assign(rgba, rgba + size_t(w) * h * 4);
    s->s.textures.push_back(std::move(im));
    return int(s->s.textures.size()) - 1;
}

void orc_scene_set_sun(orc_scene *s, const orc_sun *sun, int enabled)
{
    if (sun)
    {
        s->s.sun           = *sun;
        s->s.sun_transform = sun_transform_of(V3(sun->direction[0], sun->direction[1], sun->direction[2]));
    }
    s->s.sun_enabled = enabled != 0;
}

void orc_scene_set_skybox(orc_scene *s, const float *rgba, uint32_t w, uint32_t h, float ru, float rv)
{
    s->s.has_skybox = rgba != nullptr && w && h;
    if (s->s.has_skybox)
    {
        s->s.skybox.w = w, s->s.skybox.h = h;
        s->s.skybox.px.assign(rgba, rgba + size_t(w) * h * 4);
    }
    s->s.skybox_rot = vec2 { ru, rv };
}

void orc_scene_set_camera(orc_scene *s, const orc_camera *c)
{
    s->s.cam.p = *c;
    s->s.cam.update_cache();
}

double orc_scene_commit(orc_scene *s)
{
    auto t0 = std::chrono::steady_clock::now();
    for (auto &m : s->s.models) m.build();
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

void orc_intersect_batch(orc_scene *s, const orc_ray *rays, orc_hit *hits, uint64_t n, int nthreads)
{
    parallel_for(n, nthreads, [&](uint64_t i) { hits[i] = batch_query(s->s, rays[i], false); });
}
void orc_intersect_brute(orc_scene *s, const orc_ray *rays, orc_hit *hits, uint64_t n, int nthreads)
{
    parallel_for(n, nthreads, [&](uint64_t i) { hits[i] = batch_query(s->s, rays[i], true); });
}
void orc_occluded_batch(orc_scene *s, const orc_ray *rays, uint8_t *occ, uint64_t n, int nthreads)
{
    parallel_for(n, nthreads, [&](uint64_t i) { occ[i] = batch_query(s->s, rays[i], false).prim != 0xffffffffu; });
}

orc_render *orc_render_create(orc_scene *s, uint32_t w, uint32_t h, uint32_t mb, uint32_t seed) { return new orc_render(&s->s, w, h, mb, seed); }
void        orc_render_destroy(orc_render *r) { delete r; }
void        orc_render_reset(orc_render *r) { r->r.reset(); }
void        orc_render_set_extended(orc_render *r, int on) { r->r.extended = on != 0, r->r.light_nee = on != 2; }
void        orc_render_set_rows(orc_render *r, uint32_t y0, uint32_t y1) { r->r.row0 = y0, r->r.row1 = std::min(y1, r->r.h); }
void        orc_render_samples(orc_render *r, uint32_t first, uint32_t n, int nthreads) { r->r.run(first, n, nthreads); }

void orc_render_set_sample_table(orc_render *r, float *table, uint32_t n_samples, uint32_t dims, int replay)
{
    r->r.table = table, r->r.table_samples = n_samples, r->r.table_dims = dims;
    if (replay) r->r.sampler_mode = 2;
}

void orc_render_set_reference_stream(orc_render *r, int on, uint64_t discard)
{
    r->r.sampler_mode = on ? 1 : 0;
    r->r.mt           = std::mt19937();    // default seed 5489, like `thread_local std::mt19937 gen;`
    r->r.mt.discard(discard);
}

// ---- the bare BVH + triangle test (stands in for the Embree scene of ONE model): used by oracle/ref's embree3 shim
struct orc_rawbvh
{
    model m;
};
orc_rawbvh *orc_rawbvh_create(const float *verts9, uint32_t ntris)
{
    orc_rawbvh *b = new orc_rawbvh();
    b->m.ntris    = ntris;
    b->m.verts.resize(size_t(ntris) * 3);
    if (ntris) std::memcpy(static_cast<void *>(b->m.verts.data()), verts9, sizeof(float) * 9 * size_t(ntris));
    b->m.build();
    return b;
}
void orc_rawbvh_destroy(orc_rawbvh *b) { delete b; }
int  orc_rawbvh_intersect(const orc_rawbvh *b, const float *o, const float *d, float tnear, float tfar, float *t, float *u, float *v, uint32_t *prim)
{
    tri_hit h;
    b->m.intersect(V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]), tnear, tfar, h);
    if (h.prim == 0xffffffffu) return 0;
    *t = h.t, *u = h.u, *v = h.v, *prim = h.prim;
    return 1;
}

void orc_render_read(orc_render *r, int kind, float *dst)
{
    render      &R = r->r;
    const size_t n = size_t(R.w) * R.h;
    switch (kind)
    {
    case ORC_RAW_SUM:
        for (size_t i = 0; i < n; i++) dst[4 * i] = R.raw[3 * i], dst[4 * i + 1] = R.raw[3 * i + 1], dst[4 * i + 2] = R.raw[3 * i + 2], dst[4 * i + 3] = float(R.current_sample);
        break;
    case ORC_PROGRESS: std::memcpy(dst, R.buffer.data(), n * 16); break;
    case ORC_ALBEDO: std::memcpy(dst, R.albedo.data(), n * 16); break;
    case ORC_NORMAL: std::memcpy(dst, R.normals.data(), n * 16); break;
    case ORC_DEPTH: std::memcpy(dst, R.depth.data(), n * 16); break;
    }
}

void orc_render_stats(orc_render *r, orc_stats *out)
{
    out->total_queries = r->r.total_queries, out->ref_rays = r->r.ref_rays;
    out->pixel_samples = r->r.pixel_samples, out->passes = r->r.current_sample;
}

void orc_render_primary_hits(orc_render *r, uint32_t sample, orc_hit *hits, int nthreads)
{
    render &R = r->r;
    parallel_for(uint64_t(R.w) * R.h, nthreads, [&](uint64_t i) {
        record   rec;
        uint64_t q = 0, f = 0;
        R.sample_pixel(i % R.w, i / R.w, sample, q, f, &rec);
        hits[i] = orc_hit { rec.t, rec.bu, rec.bv, rec.prim, rec.model, rec.inst };
    });
}

// ---- KAT probes
void orc_kat_mt19937_randf(uint32_t n, float *out)
{
    std::mt19937                          gen;    // renderer.cpp:8, default seed 5489
    std::uniform_real_distribution<float> dist(0.f, 1.f);
    for (uint32_t i = 0; i < n; i++) out[i] = dist(gen);
}
float orc_kat_rng(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t dim) { return rnd(path_key(seed, pixel, sample), dim); }
void  orc_kat_camera_ray(const orc_camera *c, float x, float y, float aspect, float *o3, float *d3)
{
    camera cam;
    cam.p = *c;
    // The reference leaves _cached_matrix at identity until translate()/rotate() is called
    // (camera.cpp:3-4,41-52); a zero rotation here reproduces that (direction ignores translation).
    cam.update_cache();
    ray r = cam.get_ray(x, y, aspect);
    o3[0] = r.origin.x, o3[1] = r.origin.y, o3[2] = r.origin.z;
    d3[0] = r.direction.x, d3[1] = r.direction.y, d3[2] = r.direction.z;
}
void orc_kat_build_local(const float *n3, float *t3, float *b3)
{
    local_coords lc = build_local(V3(n3[0], n3[1], n3[2]));
    t3[0] = lc.tangent.x, t3[1] = lc.tangent.y, t3[2] = lc.tangent.z;
    b3[0] = lc.bi_tangent.x, b3[1] = lc.bi_tangent.y, b3[2] = lc.bi_tangent.z;
}
void orc_kat_sun_transform(const float *d3, float *m9)
{
    mat3 M = sun_transform_of(V3(d3[0], d3[1], d3[2]));
    m9[0] = M.c0.x, m9[1] = M.c0.y, m9[2] = M.c0.z, m9[3] = M.c1.x, m9[4] = M.c1.y, m9[5] = M.c1.z, m9[6] = M.c2.x, m9[7] = M.c2.y, m9[8] = M.c2.z;
}
void orc_kat_map_to_solid_angle(float u, float v, float theta_max, float *out3, float *pdf)
{
    vec3 d = map_to_solid_angle(u, v, theta_max);
    out3[0] = d.x, out3[1] = d.y, out3[2] = d.z;
    *pdf    = solid_angle_mapping_pdf(theta_max);
}
void orc_kat_sphere(float u, float v, float *out3)
{
    vec3 d = sphere(u, v);
    out3[0] = d.x, out3[1] = d.y, out3[2] = d.z;
}
void orc_kat_process_hit(const orc_material *m, const float *n3, const float *p3, const float *d3, float u0, float u1, float *o_out, float *d_out,
                         float *albedo3, int *is_alpha)
{
    scene  sc;
    record rec;
    rec.distance = 1.0f, rec.material = m;
    rec.normal = V3(n3[0], n3[1], n3[2]), rec.point = V3(p3[0], p3[1], p3[2]);
    ray           r { V3(0, 0, 0), V3(d3[0], d3[1], d3[2]) };
    processed_hit ph = process_hit(rec, r, sc, u0, u1);
    o_out[0] = ph.r.origin.x, o_out[1] = ph.r.origin.y, o_out[2] = ph.r.origin.z;
    d_out[0] = ph.r.direction.x, d_out[1] = ph.r.direction.y, d_out[2] = ph.r.direction.z;
    albedo3[0] = ph.albedo.x, albedo3[1] = ph.albedo.y, albedo3[2] = ph.albedo.z;
    *is_alpha  = ph.is_alpha;
}
int orc_kat_scatter_extended(const orc_material *m, const float *n3, const float *p3, const float *d3, float u0, float u1, float *w3, float *o3, float *dir3)
{
    const ray         r { V3(0, 0, 0), V3(d3[0], d3[1], d3[2]) };
    const ext_scatter e = scatter_extended(*m, V3(m->colour[0], m->colour[1], m->colour[2]), V3(n3[0], n3[1], n3[2]), V3(p3[0], p3[1], p3[2]), r, u0, u1);
    w3[0] = e.weight.x, w3[1] = e.weight.y, w3[2] = e.weight.z;
    o3[0] = e.r.origin.x, o3[1] = e.r.origin.y, o3[2] = e.r.origin.z;
    dir3[0] = e.r.direction.x, dir3[1] = e.r.direction.y, dir3[2] = e.r.direction.z;
    return e.absorbed ? 2 : (e.specular ? 1 : 0);
}
float orc_kat_resolve(float sum, uint32_t n_plus_1) { return std::pow(clampf(sum / float(n_plus_1), 0.0f, 1.0f), 1.f / 2.2f); }
int   orc_kat_tri(const float *a, const float *b, const float *c, const float *o, const float *d, float tmin, float tmax, float *t, float *u, float *v)
{
    vec3 v0 = V3(a[0], a[1], a[2]), v1 = V3(b[0], b[1], b[2]), v2 = V3(c[0], c[1], c[2]);
    return tri_test(v0, v1 - v0, v2 - v0, V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]), tmin, tmax, *t, *u, *v) ? 1 : 0;
}
void orc_kat_sky_uv(const float *d, float *uv)
{
    uv[0] = 0.5f + std::atan2(d[2], d[0]) * INV_TAU;    // renderer.cpp:279-281
    uv[1] = 0.5f - std::asin(d[1]) * INV_PI;
}
void orc_kat_image_get_uv_index(float u, float v, uint32_t w, uint32_t h, uint32_t *xy)
{
    xy[0] = uint32_t(image::to_u64(u * float(uint64_t(w))) % w);
    xy[1] = uint32_t(image::to_u64(v * float(uint64_t(h))) % h);
}
}
