// render.cu — wavefront path tracer (K6-K9).
//
// Replaces cr::renderer's hot loop: the management thread that issues one pass at a time
// (src/render/renderer.cpp:116-144), the per-scanline tasks (:240-256), _sample_pixel (:258-384) and
// process_hit (:21-102). Instead of "one CPU task per scanline, one sample per pixel per pass", a batch
// of (pixels x samples) paths is resident in HBM as a structure of arrays and every bounce is a
// sequence of persistent kernels fed by warp-aggregated atomic queues:
//
//   k_raygen      camera::get_ray + jitter (camera.cpp:14-39, renderer.cpp:260-263)            K6
//   per bounce i:
//     k_trace     closest hit for every active path (scene::cast_ray semantics) and a material
//                 sort: each hit is pushed to the queue of its shade type (miss/metal/smooth/glass)
//     k_shade     process_hit + the body of the bounce loop (renderer.cpp:277-313) + sun-NEE sample
//                 generation (:316-329); pushes survivors to the next queue, shadow rays to theirs  K7
//     k_shadow    sun visibility (renderer.cpp:330-353): any-hit, or the reference's closest-hit
//                 march when alpha cut-outs exist; connects the contribution                        K8
//     k_advance   swaps queue counters, accumulates ray statistics
//   k_accumulate  float4 accumulation buffer += per-sample radiance in sample order, AOVs, resolve
//                 pow(clamp(sum/n,0,1),1/2.2) (renderer.cpp:358-383)                                K9
//
// All shading arithmetic is single-rounding float in the reference's operation order (vecmath.cuh);
// this file is compiled with -fmad=false. Sampler dimensions: see rnd() below and DESIGN.md.
#include "render.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <vector>

namespace crb
{
    namespace
    {
        constexpr float TAU_F     = 6.28318530717f;           // src/util/numbers.h:17
        constexpr float INV_PI_F  = 1.0f / 3.14159265359f;    // :22
        constexpr float INV_TAU_F = 1.0f / 6.28318530717f;    // :24

        // ---- counter-based sampler (replaces the thread_local mt19937 of renderer.cpp:6-11, which is
        // default-seeded per worker thread and not reproducible). The key is two independent 32-bit hashes of
        // (seed, pixel, sample) - no two paths of a frame share a stream; u(dim) = top 24 bits of
        // mix(mix(k1 + dim*golden) ^ k2). dims: 0,1 jitter; 2+4i+{0,1} scatter of bounce i; 2+4i+{2,3} sun cone
        // sample of bounce i.
        __device__ __forceinline__ uint32_t mix32(uint32_t x)
        {
            x ^= x >> 16;
            x *= 0x7feb352du;
            x ^= x >> 15;
            x *= 0x846ca68bu;
            x ^= x >> 16;
            return x;
        }
        struct PathKey
        {
            uint32_t k1, k2;
        };
        __device__ __forceinline__ PathKey path_key(uint32_t seed, uint32_t pixel, uint32_t sample)
        {
            return PathKey { mix32(mix32(mix32(seed + 0x9e3779b9u) ^ pixel) ^ sample), mix32(mix32(mix32(seed + 0x85ebca6bu) ^ sample) ^ pixel) };
        }
        __device__ __forceinline__ float rnd(PathKey key, uint32_t dim)
        {
            return float(mix32(mix32(key.k1 + dim * 0x9e3779b9u) ^ key.k2) >> 8) * (1.0f / 16777216.0f);
        }

        // a caller-supplied sample table (crb_render_set_sample_table) replaces the hash: [sample][pixel][dimension]
        __device__ __forceinline__ float rnd_dim(const RenderParams &rp, PathKey key, uint32_t pixel, uint32_t sample, uint32_t dim)
        {
            if (rp.table && sample < rp.table_samples && dim < rp.table_dims)
                return rp.table[(size_t(sample) * rp.w * rp.h + pixel) * rp.table_dims + dim];
            return rnd(key, dim);
        }

        __device__ __forceinline__ float inf_f() { return __int_as_float(0x7f800000); }

        // k_shade's path record: all slot-indexed 16-byte loads issued together right after the queue entry is known.
        // Written as `hit.w first, classify, then the rest`, the kernel ran a chain of five to six dependent DRAM round
        // trips per tile (queue entry -> hit -> ray/throughput -> triangle -> material -> radiance; ncu: 49 % of DRAM
        // peak at 31-43 % issue utilisation, stalls on exactly those first uses); the asm keeps ptxas from sinking the
        // loads below the miss/hit branch. `rad` is only prefetched (it is consumed last; no registers held).
#ifndef CRB_SHADE_GROUPED_LOADS
#define CRB_SHADE_GROUPED_LOADS 1    // k_shade 14.5 -> 13.3 ms per 3 steps, +1.0 % end to end (profiles/r1g_sweeps.md section 6)
#endif
        __device__ __forceinline__ float4 ld128_grouped(const float4 *p)
        {
#if defined(CRB_EMU) || !CRB_SHADE_GROUPED_LOADS
            return *p;
#else
            float4 v;
            asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
            return v;
#endif
        }
        __device__ __forceinline__ void prefetch_l2(const void *p)
        {
#if !defined(CRB_EMU) && CRB_SHADE_GROUPED_LOADS
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
            (void) p;
#endif
        }

        // Queue records are read once and written once per bounce: streaming loads / stores (evict-first) keep them from
        // displacing BVH lines in the L1 / L2 the traversal lives on
#ifndef CRB_STREAM_QUEUES
#define CRB_STREAM_QUEUES 1
#endif
        __device__ __forceinline__ float4 ld_stream(const float4 *p)
        {
#if defined(CRB_EMU) || !CRB_STREAM_QUEUES
            return *p;
#else
            return __ldcs(p);
#endif
        }
        __device__ __forceinline__ void st_stream(float4 *p, float4 v)
        {
#if defined(CRB_EMU) || !CRB_STREAM_QUEUES
            *p = v;
#else
            __stcs(p, v);
#endif
        }

        // Block-aggregated reservation of queue space in NQ queues at once: warps count with a ballot, add
        // into shared memory, and ONE thread per queue issues the global atomic for the whole block
        // (same-address global atomics serialise in L2; per-warp pushes of a full-frame wavefront are
        // hundreds of thousands of them). Must be called by every thread of the block.
        template<int NQ>
        __device__ __forceinline__ void block_reserve(const bool (&pred)[NQ], uint32_t *counters, const int (&cidx)[NQ], uint32_t (&at)[NQ])
        {
            __shared__ uint32_t s_cnt[NQ], s_base[NQ];
            const unsigned lane = crb_lane_id();
            for (int q = threadIdx.x; q < NQ; q += blockDim.x) s_cnt[q] = 0;
            __syncthreads();
            uint32_t woff[NQ];
            unsigned mask[NQ];
#pragma unroll
            for (int q = 0; q < NQ; q++)
            {
                mask[q] = __ballot_sync(0xffffffffu, pred[q]);
                woff[q] = 0;
                if (mask[q])
                {
                    if (lane == 0) woff[q] = atomicAdd(&s_cnt[q], uint32_t(__popc(mask[q])));
                    woff[q] = __shfl_sync(0xffffffffu, woff[q], 0);
                }
            }
            __syncthreads();
            for (int q = threadIdx.x; q < NQ; q += blockDim.x)
                if (s_cnt[q]) s_base[q] = atomicAdd(counters + cidx[q], s_cnt[q]);
            __syncthreads();
#pragma unroll
            for (int q = 0; q < NQ; q++) at[q] = s_base[q] + woff[q] + uint32_t(__popc(mask[q] & ((1u << lane) - 1u)));
        }

        // persistent work fetch: a warp takes CRB_WARP consecutive items at a time
        __device__ __forceinline__ uint32_t warp_fetch(uint32_t *cursor)
        {
            uint32_t base = 0;
            if (crb_lane_id() == 0) base = atomicAdd(cursor, uint32_t(CRB_WARP));
            return __shfl_sync(0xffffffffu, base, 0);
        }

        // ---- cr::image::get_uv, src/objects/image.h:104-121. static_cast<uint64_t>(float) of a negative
        // value goes through the signed conversion on x86-64 (what the reference's build computes).
        __device__ __forceinline__ unsigned long long to_u64(float f)
        {
            if (!(f == f)) return 0x8000000000000000ull;
            if (f >= 9223372036854775808.0f) return (unsigned long long) f;
            return (unsigned long long) (long long) f;
        }
        __device__ __forceinline__ float4 image_get_uv(const float4 *px, uint32_t w, uint32_t h, float u, float v)
        {
            const unsigned long long x = to_u64(u * float(w)) % w;
            const unsigned long long y = to_u64(v * float(h)) % h;
            return __ldg(px + (x + y * w));
        }

        // the (model, instance) range that holds a flat primitive id (two-level mode: range index = instance index)
        __device__ __forceinline__ uint32_t range_of(const DScene &sc, uint32_t flat)
        {
            uint32_t lo = 0, hi = sc.n_ranges;
            while (hi - lo > 1)
            {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(&sc.ranges[mid].start) <= flat) lo = mid; else hi = mid;
            }
            return lo;
        }
        __device__ __forceinline__ uint32_t src_tri(const DScene &sc, uint32_t flat)
        {
            if (sc.two_level)
            {
                const FlatRange r = sc.ranges[range_of(sc, flat)];
                return r.src_start + (flat - r.start);
            }
            return sc.flat_src ? __ldg(sc.flat_src + flat) : flat;
        }

        // local row r of this render call -> frame row y in sample space: a contiguous range [row0, row0+nrows), or the
        // interleaved bands of the tile partition (band rows each, every band_stride-th band starting at band_first):
        // all of a GPU's bands are one launch sequence, not one per band
        __device__ __forceinline__ uint32_t row_of(const RenderParams &rp, uint32_t r)
        {
            if (rp.band == 0) return rp.row0 + r;
            const uint32_t b = r / rp.band;
            return (b * rp.band_stride + band_slot(b, rp.band_first, rp.band_stride, rp.band_serp)) * rp.band + (r - b * rp.band);
        }

        __device__ __forceinline__ uint32_t flipped_index(const RenderParams &rp, uint32_t x, uint32_t y)
        {
            return (rp.w - 1 - x) + (rp.h - 1 - y) * rp.w;    // renderer.cpp:358-362
        }

        // ------------------------------------------------------------------ K6 ray generation
        __global__ void __launch_bounds__(256) k_raygen(DScene sc, RenderParams rp, PathState ps)
        {
            const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
            if (slot >= rp.npix * rp.batch) return;
            const uint32_t pix = slot % rp.npix, s = slot / rp.npix;
            const uint32_t x = pix % rp.w, y = row_of(rp, pix / rp.w);
            const uint32_t sample = rp.first_sample + s;
            const PathKey  key    = path_key(rp.seed, x + y * rp.w, sample);
            const float    fx = (float(x) + rnd_dim(rp, key, x + y * rp.w, sample, 0)) / float(rp.w),
                           fy = (float(y) + rnd_dim(rp, key, x + y * rp.w, sample, 1)) / float(rp.h);    // renderer.cpp:260-263
            const DCamera &c = sc.cam;
            V3             o, d;
            if (c.mode == 0)
            {
                // camera.cpp:18-27
                const float u = (2.0f * fx - 1.0f) * c.aspect;
                const float v = 2.0f * fy - 1.0f;
                const float w = c.w;
                float       r[3];
#pragma unroll
                for (int k = 0; k < 3; k++) r[k] = (c.m[0][k] * u + c.m[1][k] * v) + (c.m[2][k] * w + c.trans[k] * 0.0f);
                o = v3(c.position[0], c.position[1], c.position[2]);
                d = normalize(v3(r[0], r[1], r[2]));
            }
            else
            {
                // camera.cpp:28-37
                const float u = 2.0f * fx - 1.0f, v = 2.0f * fy - 1.0f;
                const float su = c.scale * u, sv = c.scale * v;
                float       r[3];
#pragma unroll
                for (int k = 0; k < 3; k++) r[k] = (c.m[0][k] * su + c.m[1][k] * sv) + (c.m[2][k] * 0.0f + c.trans[k] * 1.0f);
                o = v3(r[0], r[1], r[2]);
                d = normalize(v3(c.m[2][0], c.m[2][1], c.m[2][2]));
            }
            ps.ray_o[slot] = make_float4(o.x, o.y, o.z, 0.f);
            ps.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.f);
            ps.thr[slot]   = make_float4(1.f, 1.f, 1.f, 1.f);    // w: extended mode's "previous vertex was specular" flag (the camera counts)
            ps.rad[slot]   = make_float4(0.f, 0.f, 0.f, 0.f);
            ps.q_in[slot]  = slot;
            if (sample == rp.aov_sample)
            {
                // defaults of renderer.cpp:265-269 pushed through :367-369
                const uint32_t fi = flipped_index(rp, x, y);
                rp.albedo[fi]     = make_float4(0.f, 0.f, 0.f, 1.f);
                rp.normal[fi]     = make_float4(.5f, .5f, .5f, 1.f);
                rp.depth[fi]      = make_float4(0.f, 0.f, 0.f, 1.f);
            }
        }

#ifdef CRB_EMU
        // Kernel-logic harness only (compiled out of the product): per-RAY node visits and triangle tests of the counting
        // instantiations, by query kind and outcome - [kind][outcome] = { rays, node visits, triangle tests, rays that needed at
        // most one node visit }; kind 0 closest hit (k_trace), 1 shadow (k_shadow); outcome 0 nothing hit, 1 hit. The harness
        // runs one lane, so the counters' growth between two retirements belongs to one ray. Read by tools/tree_quality.py
        // through crb_emu_ray_classes; this is how the occluded share of the shadow rays and their cost were found
        // (profiles/r2_sweeps.md section 17).
        unsigned long long g_ray_classes[2][2][4];
        struct RayClassProbe
        {
            unsigned long long nodes = 0, tris = 0;
            void retire(int kind, bool hit, const TravCounters &tc)
            {
                unsigned long long *c = g_ray_classes[kind][hit ? 1 : 0];
                c[0]++, c[1] += tc.nodes - nodes, c[2] += tc.tris - tris, c[3] += (tc.nodes - nodes) <= 1 ? 1 : 0;
                nodes = tc.nodes, tris = tc.tris;
            }
        };
#endif

        // ------------------------------------------------------------------ trace + material sort
        constexpr int TRACE_STEPS = 4;    // node iterations between two refill points of the persistent trace loop

        // resident CTAs per SM and CTA size of the two traversal kernels
        // With the rarely-touched lane state in shared memory (bvh8.cuh CRB_TP_SMEM) the loop compiles to 56 registers
        // without spills: 9 CTAs of 128 threads per SM = 9 warps per scheduler. Measured on config 2 (profiles/r2_sweeps.md
        // section 7): 256 x 4 (64 regs) 3468, 128 x 8 3451, 128 x 9 (56 regs) 3542, 128 x 10 (48 regs) 3535, 128 x 12 (40 regs,
        // 72 B of spills) 3083 Mrays/s.
#ifndef CRB_TRACE_OCC
#define CRB_TRACE_OCC 9
#endif
#ifndef CRB_TRACE_BLOCK
#define CRB_TRACE_BLOCK 128    // threads per CTA of the two single-level traversal kernels (at most 256: bvh8.cuh CRB_TP_SMEM)
#endif
        static_assert(CRB_TRACE_BLOCK <= 256 && CRB_TRACE_BLOCK % 32 == 0, "the trace loops keep per-lane state in shared arrays of 256 entries (bvh8.cuh)");
        template<bool COUNT, int STEPS>
        __global__ void __launch_bounds__(CRB_TRACE_BLOCK, CRB_TRACE_OCC) k_trace(DScene sc, PathState ps)
        {
            const uint32_t n = ps.counters[CTR_IN];
            TravCounters   tc;
            auto source = [&](uint32_t idx, uint32_t &slot, V3 &o, V3 &d, float &tmin, float &tmax) {
                slot            = idx;    // the path records are in queue order: no indirection, coalesced loads
                const float4 ro = ld_stream(ps.ray_o + idx), rd = ld_stream(ps.ray_d + idx);
                o               = v3(ro.x, ro.y, ro.z);
                d               = normalize(v3(rd.x, rd.y, rd.z));    // model.cpp:107-112: the query direction is normalised
                tmin = 0.00001f, tmax = inf_f();                      // model.cpp:21-22
            };
            // retiring a ray is one 16-byte store; the material sort is a separate full-width pass
            // (k_classify) because only a few lanes of a warp retire at any refill point
#ifdef CRB_EMU
            RayClassProbe probe;
#endif
            auto sink = [&](bool valid, uint32_t slot, const Hit &h) {
#ifdef CRB_EMU
                if (valid && COUNT) probe.retire(0, h.prim != INVALID_PRIM, tc);
#endif
                if (valid) st_stream(ps.hit + slot, make_float4(h.t, h.u, h.v, __uint_as_float(h.prim)));
            };
            trace_persistent<COUNT, STEPS>(sc.bvh, ps.counters + CTR_CUR_TRACE, n, ps.trace_chunk, false, source, sink, &tc);
            if (COUNT)
            {
                atomicAdd(ps.stats + ST_NODES, tc.nodes);
                atomicAdd(ps.stats + ST_TRIS, tc.tris);
            }
        }

        // material sort: every traced path is pushed to the queue of its shade class (0 miss, 1 metal,
        // 2 smooth, 3 glass) with one warp-aggregated atomic per class per warp
        __global__ void __launch_bounds__(256) k_classify(DScene sc, PathState ps)
        {
            const uint32_t n = ps.counters[CTR_IN];
            // all threads of a block iterate together (the reservation is block-collective)
            for (uint32_t tile = blockIdx.x * blockDim.x; tile < n; tile += gridDim.x * blockDim.x)
            {
                const uint32_t idx  = tile + threadIdx.x;
                int            cls  = -1;
                uint32_t       slot = 0;
                if (idx < n)
                {
                    slot                = idx;    // the class queues hold RECORD indices
                    const uint32_t prim = __float_as_uint(ps.hit[idx].w);
                    if (prim == INVALID_PRIM)
                        cls = 0;
                    else
                    {
                        const uint32_t mat = __float_as_uint(__ldg(sc.shade_tri + src_tri(sc, prim)).w);
                        cls                = 1 + int(sc.materials[mat].shade_type);
                    }
                }
                const bool pred[4] = { cls == 0, cls == 1, cls == 2, cls == 3 };
                const int  cidx[4] = { CTR_CLASS0, CTR_CLASS0 + 1, CTR_CLASS0 + 2, CTR_CLASS0 + 3 };
                uint32_t   at[4];
                block_reserve<4>(pred, ps.counters, cidx, at);
                if (cls == 0) ps.q_class[0][at[0]] = slot;
                if (cls == 1) ps.q_class[1][at[1]] = slot;
                if (cls == 2) ps.q_class[2][at[2]] = slot;
                if (cls == 3) ps.q_class[3][at[3]] = slot;
            }
        }

        // ------------------------------------------------------------------ shading
        struct Surface
        {
            V3       normal;    // object-space unit geometric normal, not face-forwarded (model.cpp:35)
            V3       point;     // intersection_point (model.cpp:33,116)
            float    distance;  // re-measured |point - origin| (model.cpp:119-120)
            uint32_t mat;
            float    uvx, uvy;
        };

        // o, d: the cr::ray as the reference holds it (direction not re-normalised). model.cpp:107-120: per instance the ray is
        // taken to object space with the inverse transform, the direction renormalised, the hit point mapped back with the
        // transform and the distance re-measured in world space. Single-level scenes (identity instances) reduce to
        // point = o + normalize(d) * t.
        __device__ __forceinline__ Surface surface_at(const DScene &sc, V3 o, V3 d, float4 hit)
        {
            Surface        s;
            const uint32_t flat = __float_as_uint(hit.w);
            uint32_t       src;
            if (sc.two_level)
            {
                const uint32_t  k = range_of(sc, flat);
                const FlatRange r = sc.ranges[k];
                src               = r.src_start + (flat - r.start);
                const Instance &I = sc.bvh2.inst[k];
                const V3        oo = xf34(I.inv, o, 1.0f), dd = normalize(xf34(I.inv, d, 0.0f));
                s.point            = xf34(I.fwd, oo + dd * hit.x, 1.0f);
            }
            else
            {
                src     = sc.flat_src ? __ldg(sc.flat_src + flat) : flat;
                s.point = o + normalize(d) * hit.x;    // ray.at(tfar), ray.cpp:13-16
            }
            const float4 st = __ldg(sc.shade_tri + src);
            s.normal        = v3(st.x, st.y, st.z);
            s.mat           = __float_as_uint(st.w);
            s.distance      = length(s.point - o);       // glm::distance
            s.uvx = s.uvy = 0.f;
            if (sc.obj_uvs)
            {
                // rtcInterpolate0 (model.cpp:39-47): (1-u-v)*t0 + u*t1 + v*t2
                const float *t = sc.obj_uvs + size_t(src) * 6;
                const float  w = 1.0f - hit.y - hit.z;
                s.uvx          = (w * t[0] + hit.y * t[2]) + hit.z * t[4];
                s.uvy          = (w * t[1] + hit.y * t[3]) + hit.z * t[5];
            }
            return s;
        }

        __device__ __forceinline__ float4 surface_colour(const DScene &sc, const DMaterial &m, const Surface &s)
        {
            if (m.tex >= 0)
            {
                const DTexture t = sc.textures[m.tex];
                return image_get_uv(sc.texels + t.offset, t.w, t.h, s.uvx, s.uvy);    // renderer.cpp:30-33
            }
            return make_float4(m.colour[0], m.colour[1], m.colour[2], m.colour[3]);
        }

        // src/util/sampling.h:156-166
        __device__ __forceinline__ V3 sample_sphere(float ux, float uy)
        {
            const float cos_theta = 2.0f * ux - 1.0f;
            const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
            const float phi       = TAU_F * uy;
            const float sin_phi = sinf(phi), cos_phi = cosf(phi);
            return v3(sin_theta * cos_phi, cos_theta, sin_theta * sin_phi);
        }
        // src/util/sampling.h:35-42 with 1-cos(theta_max) hoisted to the host
        __device__ __forceinline__ V3 map_to_solid_angle(float ux, float uy, float one_minus_cos)
        {
            const float phi       = TAU_F * ux;
            const float cos_theta = 1.0f - uy * one_minus_cos;
            const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
            return v3(cosf(phi) * sin_theta, cos_theta, sinf(phi) * sin_theta);
        }

        // CTA size of k_shade: its queue push is CTA-collective (three barriers per tile), so a CTA waits for its slowest warp
#ifndef CRB_SHADE_BLOCK
#define CRB_SHADE_BLOCK 128    // 256: 13.4, 128: 13.1, 512: 13.9 ms per 3 steps (profiles/r1g_sweeps.md section 6)
#endif
        __global__ void __launch_bounds__(CRB_SHADE_BLOCK, 1024 / CRB_SHADE_BLOCK) k_shade(DScene sc, RenderParams rp, PathState ps)
        {
            const uint32_t c0 = ps.counters[CTR_CLASS0], c1 = c0 + ps.counters[CTR_CLASS0 + 1], c2 = c1 + ps.counters[CTR_CLASS0 + 2],
                           n = ps.sorted ? c2 + ps.counters[CTR_CLASS0 + 3] : ps.counters[CTR_IN];
            const uint32_t i = rp.bounce;
            for (uint32_t tile = blockIdx.x * blockDim.x; tile < n; tile += gridDim.x * blockDim.x)
            {
                const uint32_t idx      = tile + threadIdx.x;
                bool           survive  = false, want_shadow = false;
                uint32_t       slot     = 0;
                ShadowRay      sr;
                float4         out_o, out_d, out_t;    // the surviving path's next record, written at its place in the next queue
                if (idx < n)
                {
                    int      cls;
                    uint32_t rec;    // index of the path's record in the queue-ordered arrays
                    if (ps.sorted)
                    {
                        cls = idx < c0 ? 0 : (idx < c1 ? 1 : (idx < c2 ? 2 : 3));
                        rec = ps.q_class[cls][idx - (cls == 0 ? 0u : (cls == 1 ? c0 : (cls == 2 ? c1 : c2)))];
                    }
                    else
                    {
                        // unsorted mode: paths are shaded in queue order, every load below is coalesced
                        rec = idx;
                        cls = -1;
                    }
                    slot            = ps.q_in[rec];
                    const float4 h4 = ld128_grouped(ps.hit + rec), ro = ld128_grouped(ps.ray_o + rec), rd = ld128_grouped(ps.ray_d + rec),
                                 t4 = ld128_grouped(ps.thr + rec);
                    prefetch_l2(ps.rad + slot);
                    if (cls < 0) cls = __float_as_uint(h4.w) == INVALID_PRIM ? 0 : 1;
                    const V3     o = v3(ro.x, ro.y, ro.z), d = v3(rd.x, rd.y, rd.z);
                    V3           thr = v3(t4.x, t4.y, t4.z);
                    const uint32_t pix = slot % rp.npix, s = slot / rp.npix;
                    const uint32_t x = pix % rp.w, y = row_of(rp, pix / rp.w);
                    const uint32_t sample = rp.first_sample + s;
                    const bool     aov    = (i == 0) && (sample == rp.aov_sample);

                    if (cls == 0)
                    {
                        // renderer.cpp:277-289
                        V3 ms = v3(0.f, 0.f, 0.f);
                        if (sc.skybox)
                        {
                            const float mu = 0.5f + atan2f(d.z, d.x) * INV_TAU_F;
                            const float mv = 0.5f - asinf(d.y) * INV_PI_F;
                            const float4 c = image_get_uv(sc.skybox, sc.sky_w, sc.sky_h, mu + sc.sky_rot[0], mv + sc.sky_rot[1]);    // scene.cpp:67-77
                            ms             = v3(c.x, c.y, c.z);
                        }
                        if (aov) rp.albedo[flipped_index(rp, x, y)] = make_float4(ms.x, ms.y, ms.z, 1.f);
                        if (ms.x != 0.f || ms.y != 0.f || ms.z != 0.f)
                        {
                            // (adding thr * 0 = +0 changes no bit of the non-negative sum: the 32-byte read-modify-write of
                            // the radiance record is skipped for a black sky)
                            const float4 r4 = ps.rad[slot];
                            const V3     r  = v3(r4.x, r4.y, r4.z) + thr * ms;
                            ps.rad[slot]    = make_float4(r.x, r.y, r.z, 0.f);
                        }
                    }
                    else
                    {
                        const V3         dn  = normalize(d);
                        const Surface    sf  = surface_at(sc, o, d, h4);
                        const DMaterial  mat = sc.materials[sf.mat];
                        const float4     col = surface_colour(sc, mat, sf);
                        if (col.w == 0.0f)
                        {
                            // alpha cut-out: renderer.cpp:294-301 (steps 0.1 along the un-normalised direction)
                            const V3 p = sf.point + d * 0.1f;
                            out_o = make_float4(p.x, p.y, p.z, 0.f), out_d = rd, out_t = t4;
                            survive = true;
                        }
                        else
                        {
                            V3 albedo = v3(col.x, col.y, col.z);
                            V3 no, nd;
                            if (mat.shade_type == CRB_GLASS)
                            {
                                // renderer.cpp:47-77
                                V3       out_normal = sf.normal;
                                const V3 reflected  = reflect(d, sf.normal);
                                float    ni_over_nt = 1.0f / mat.ior;
                                if (dot(d, sf.normal) > 0) out_normal = -sf.normal, ni_over_nt = mat.ior;
                                const V3    uv   = dn;
                                const float dt   = dot(uv, out_normal);
                                const float disc = 1.0f - ni_over_nt * ni_over_nt * (1.0f - dt * dt);
                                no               = sf.point + out_normal * -0.0001f;
                                if (disc > 0)
                                    nd = ni_over_nt * (uv - out_normal * dt) - out_normal * sqrtf(disc);
                                else
                                    nd = reflected;
                            }
                            else if (mat.shade_type == CRB_METAL)
                            {
                                // renderer.cpp:78-91 (the hemp_cos draw is computed and discarded there)
                                no     = sf.point + sf.normal * 0.0001f;
                                nd     = reflect(d, sf.normal);
                                albedo = albedo * mat.reflectiveness;
                            }
                            else
                            {
                                // renderer.cpp:92-98, sampling.h:168-172
                                const PathKey  key = path_key(rp.seed, x + y * rp.w, sample);
                                const V3       h   = sf.normal + sample_sphere(rnd_dim(rp, key, x + y * rp.w, sample, 2 + 4 * i), rnd_dim(rp, key, x + y * rp.w, sample, 2 + 4 * i + 1));
                                no                 = sf.point + sf.normal * 0.0001f;
                                nd                 = normalize(h);
                            }
                            if (aov)
                            {
                                // renderer.cpp:303-308,367-369
                                const uint32_t fi = flipped_index(rp, x, y);
                                rp.albedo[fi]     = make_float4(albedo.x, albedo.y, albedo.z, 1.f);
                                const V3 nn       = sf.normal * .5f + v3(.5f, .5f, .5f);
                                rp.normal[fi]     = make_float4(nn.x, nn.y, nn.z, 1.f);
                                const float dd    = fminf(sf.distance, 200.0f) / 200.f;
                                rp.depth[fi]      = make_float4(dd, dd, dd, 1.f);
                            }
                            // renderer.cpp:310-312
                            thr             = thr * albedo;
                            if (mat.emission != 0.0f)
                            {
                                // (same: thr * 0 = +0; only emitters touch the radiance record)
                                const float4 r4 = ps.rad[slot];
                                const V3     r  = v3(r4.x, r4.y, r4.z) + thr * mat.emission;
                                ps.rad[slot]    = make_float4(r.x, r.y, r.z, 0.f);
                            }
                            out_t   = make_float4(thr.x, thr.y, thr.z, 0.f);
                            out_o   = make_float4(no.x, no.y, no.z, 0.f);
                            out_d   = make_float4(nd.x, nd.y, nd.z, 0.f);
                            survive = true;

                            if (sc.sun.enabled)
                            {
                                // renderer.cpp:316-329,348-353; sampling.h:53-57,72-80
                                const PathKey  key = path_key(rp.seed, x + y * rp.w, sample);
                                const V3       so  = sf.point + sf.normal * 0.001f;
                                const V3       l   = map_to_solid_angle(rnd_dim(rp, key, x + y * rp.w, sample, 2 + 4 * i + 2), rnd_dim(rp, key, x + y * rp.w, sample, 2 + 4 * i + 3), sc.sun.one_minus_cos);
                                const float   *T   = sc.sun.transform;
                                const V3       dir = (v3(T[0], T[1], T[2]) * l.x + v3(T[3], T[4], T[5]) * l.y) + v3(T[6], T[7], T[8]) * l.z;
                                const float    cosine    = clampf(dot(sf.normal, dir), 0.0f, 1.0f);
                                const float    sun_angle = acosf(dot(dir, -v3(sc.sun.dir[0], sc.sun.dir[1], sc.sun.dir[2])));
                                const V3       sky = (sun_angle < sc.sun.size) ? v3(sc.sun.colour[0], sc.sun.colour[1], sc.sun.colour[2]) * sc.sun.intensity
                                                                               : v3(0.f, 0.f, 0.f);
                                const V3       contrib = thr * v3(col.x, col.y, col.z) * cosine * sky / sc.sun.pdf;
                                if (contrib.x != 0.f || contrib.y != 0.f || contrib.z != 0.f)
                                {
                                    want_shadow = true;
                                    sr.o        = make_float4(so.x, so.y, so.z, __uint_as_float(slot));
                                    sr.d        = make_float4(dir.x, dir.y, dir.z, inf_f());
                                    sr.c        = make_float4(contrib.x, contrib.y, contrib.z, 0.f);
                                }
                            }
                        }
                    }
                }
                const bool pred[2] = { survive, want_shadow };
                const int  cidx[2] = { CTR_NEXT, CTR_SHADOW };
                uint32_t   at[2];
                block_reserve<2>(pred, ps.counters, cidx, at);
                if (survive)
                {
                    // the next bounce's record, compacted: k_trace and the next k_shade stream it without indirection
                    ps.q_next[at[0]]   = slot;
                    ps.ray_o_next[at[0]] = out_o, ps.ray_d_next[at[0]] = out_d, ps.thr_next[at[0]] = out_t;
                }
                if (want_shadow) ps.shadow[at[1]] = sr;
            }
        }

        // ------------------------------------------------------------------ extended shading mode
        // CRB_RENDER_FLAG_EXTENDED: GGX metal, Fresnel dielectric, Lambert with face-forwarded normals, NEE of
        // the sun cone and of uniformly picked emissive triangles at diffuse vertices. In the reference this is
        // dead code (cr::brdf::ggx, src/render/brdf.h:10-29; cook_torrence::*, src/util/sampling.h:83-142) and
        // BASELINE configs 2 / 5 name it, so the mode is specified by the oracle ("EXTENDED shading mode" in
        // oracle.cpp) and restated here operation by operation. Sampler dimensions per bounce i:
        // 2+6i+{0,1} scatter, {2,3} light sample, {4} strategy, {5} light pick. thr.w carries the
        // "previous vertex was specular" flag.
        __device__ __forceinline__ float pow5(float m)
        {
            const float m2 = m * m;
            return (m2 * m2) * m;
        }
        // src/util/sampling.h:110-118
        __device__ __forceinline__ float specular_g(float NoV, float NoL, float a)
        {
            const float a2   = a * a;
            const float ggxv = NoL * sqrtf(NoV * NoV * (1.0f - a2) + a2);
            const float ggxl = NoV * sqrtf(NoL * NoL * (1.0f - a2) + a2);
            return 0.5f / (ggxv + ggxl);
        }
        // src/util/sampling.h:21-33 (build_local), then the half vector distributed like D(h) * (n.h)
        __device__ __forceinline__ V3 sample_ggx_h(V3 n, float a, float u0, float u1, float &cos_h)
        {
            const float a2    = a * a;
            const float cos2  = (1.0f - u0) / (1.0f + (a2 - 1.0f) * u0);
            cos_h             = sqrtf(cos2);
            const float sin_h = sqrtf(fmaxf(0.0f, 1.0f - cos2));
            const float phi   = TAU_F * u1;
            const float sg    = (n.z < 0.0f) ? -1.0f : 1.0f;
            const float ka    = -1.0f / (sg + n.z);
            const float kb    = n.x * n.y * ka;
            const V3    tangent = v3(1.0f + sg * n.x * n.x * ka, sg * kb, -sg * n.x), bitangent = v3(kb, sg + n.y * n.y * ka, -n.y);
            return (tangent * (cosf(phi) * sin_h) + n * cos_h) + bitangent * (sinf(phi) * sin_h);
        }

        #ifndef CRB_SHADE_EXT_OCC
#define CRB_SHADE_EXT_OCC 4    // CTAs per SM: 64 registers (12 bytes of spills) so that the 148 x 4 persistent grid is resident
#endif
        __global__ void __launch_bounds__(256, CRB_SHADE_EXT_OCC) k_shade_ext(DScene sc, RenderParams rp, PathState ps)
        {
            const uint32_t c0 = ps.counters[CTR_CLASS0], c1 = c0 + ps.counters[CTR_CLASS0 + 1], c2 = c1 + ps.counters[CTR_CLASS0 + 2],
                           n = ps.sorted ? c2 + ps.counters[CTR_CLASS0 + 3] : ps.counters[CTR_IN];
            const uint32_t i = rp.bounce, NL = sc.n_lights;
            for (uint32_t tile = blockIdx.x * blockDim.x; tile < n; tile += gridDim.x * blockDim.x)
            {
                const uint32_t idx     = tile + threadIdx.x;
                bool           survive = false, want_shadow = false;
                uint32_t       slot    = 0;
                ShadowRay      sr;
                float4         out_o, out_d, out_t;
                if (idx < n)
                {
                    int      cls;
                    uint32_t rec;
                    if (ps.sorted)
                    {
                        cls = idx < c0 ? 0 : (idx < c1 ? 1 : (idx < c2 ? 2 : 3));
                        rec = ps.q_class[cls][idx - (cls == 0 ? 0u : (cls == 1 ? c0 : (cls == 2 ? c1 : c2)))];
                    }
                    else
                    {
                        rec = idx;
                        cls = -1;
                    }
                    slot            = ps.q_in[rec];
                    const float4 h4 = ld128_grouped(ps.hit + rec), ro = ld128_grouped(ps.ray_o + rec), rd = ld128_grouped(ps.ray_d + rec),
                                 t4 = ld128_grouped(ps.thr + rec);
                    prefetch_l2(ps.rad + slot);
                    if (cls < 0) cls = __float_as_uint(h4.w) == INVALID_PRIM ? 0 : 1;
                    const V3     o = v3(ro.x, ro.y, ro.z), d = v3(rd.x, rd.y, rd.z);
                    const V3     thr = v3(t4.x, t4.y, t4.z);
                    const bool   specular = t4.w != 0.0f;
                    const uint32_t pix = slot % rp.npix, s = slot / rp.npix;
                    const uint32_t x = pix % rp.w, y = row_of(rp, pix / rp.w);
                    const uint32_t sample = rp.first_sample + s;
                    const bool     aov    = (i == 0) && (sample == rp.aov_sample);
                    const PathKey  key    = path_key(rp.seed, x + y * rp.w, sample);
                    const uint32_t dim    = 2 + 6 * i;
                    const V3       dn     = normalize(d);

                    if (cls == 0)
                    {
                        V3 ms = v3(0.f, 0.f, 0.f);
                        if (sc.skybox)
                        {
                            const float mu = 0.5f + atan2f(d.z, d.x) * INV_TAU_F;
                            const float mv = 0.5f - asinf(d.y) * INV_PI_F;
                            const float4 c = image_get_uv(sc.skybox, sc.sky_w, sc.sky_h, mu + sc.sky_rot[0], mv + sc.sky_rot[1]);
                            ms             = v3(c.x, c.y, c.z);
                        }
                        if (aov) rp.albedo[flipped_index(rp, x, y)] = make_float4(ms.x, ms.y, ms.z, 1.f);
                        if (sc.sun.enabled && specular)
                        {
                            // a camera / specular path that escapes sees the sun disc (sampling.h:53-57)
                            const float sun_angle = acosf(dot(dn, -v3(sc.sun.dir[0], sc.sun.dir[1], sc.sun.dir[2])));
                            if (sun_angle < sc.sun.size) ms = ms + v3(sc.sun.colour[0], sc.sun.colour[1], sc.sun.colour[2]) * sc.sun.intensity;
                            else ms = ms + v3(0.f, 0.f, 0.f);
                        }
                        const float4 r4 = ps.rad[slot];
                        const V3     r  = v3(r4.x, r4.y, r4.z) + thr * ms;
                        ps.rad[slot]    = make_float4(r.x, r.y, r.z, 0.f);
                    }
                    else
                    {
                        const Surface   sf  = surface_at(sc, o, d, h4);
                        const DMaterial mat = sc.materials[sf.mat];
                        const float4    col = surface_colour(sc, mat, sf);
                        if (col.w == 0.0f)
                        {
                            const V3 p = sf.point + d * 0.1f;    // renderer.cpp:294-301
                            out_o = make_float4(p.x, p.y, p.z, 0.f), out_d = rd, out_t = t4;
                            survive = true;
                        }
                        else
                        {
                            const V3   colour = v3(col.x, col.y, col.z);
                            const bool back   = dot(d, sf.normal) > 0;
                            const V3   ns     = back ? -sf.normal : sf.normal;
                            if (aov)
                            {
                                const uint32_t fi = flipped_index(rp, x, y);
                                rp.albedo[fi]     = make_float4(colour.x, colour.y, colour.z, 1.f);
                                const V3 nn       = sf.normal * .5f + v3(.5f, .5f, .5f);
                                rp.normal[fi]     = make_float4(nn.x, nn.y, nn.z, 1.f);
                                const float dd    = fminf(sf.distance, 200.0f) / 200.f;
                                rp.depth[fi]      = make_float4(dd, dd, dd, 1.f);
                            }
                            if (mat.emission > 0.0f && (specular || NL == 0))
                            {
                                const float4 r4 = ps.rad[slot];
                                const V3     r  = v3(r4.x, r4.y, r4.z) + thr * (v3(mat.colour[0], mat.colour[1], mat.colour[2]) * mat.emission);
                                ps.rad[slot]    = make_float4(r.x, r.y, r.z, 0.f);
                            }
                            const float u0 = rnd(key, dim), u1 = rnd(key, dim + 1);
                            V3          no, nd, weight = colour;
                            bool        next_specular = true, absorbed = false;
                            if (mat.shade_type == CRB_GLASS)
                            {
                                const float eta  = back ? mat.ior : 1.0f / mat.ior;    // renderer.cpp:52-57
                                const float dt   = dot(dn, ns);
                                const float disc = 1.0f - eta * eta * (1.0f - dt * dt);
                                float       R    = 1.0f;
                                if (disc > 0)
                                {
                                    const float cos_i = -dt, cos_t = sqrtf(disc);
                                    const float rs = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
                                    const float rp_ = (cos_i - eta * cos_t) / (cos_i + eta * cos_t);
                                    R               = 0.5f * (rs * rs + rp_ * rp_);
                                }
                                if (u0 < R)
                                {
                                    no = sf.point + ns * 0.0001f;
                                    nd = reflect(dn, ns);
                                }
                                else
                                {
                                    no = sf.point + ns * -0.0001f;
                                    nd = eta * (dn - ns * dt) - ns * sqrtf(disc);    // renderer.cpp:63-67
                                }
                            }
                            else if (mat.shade_type == CRB_METAL)
                            {
                                const float a = clampf(mat.roughness, 0.02f, 1.0f);
                                float       NoH;
                                const V3    h   = sample_ggx_h(ns, a, u0, u1, NoH);
                                const V3    wi  = reflect(dn, h);
                                const float VoH = -dot(dn, h), NoL = dot(ns, wi), NoV = -dot(dn, ns);
                                if (!(VoH > 0 && NoL > 0 && NoV > 0))
                                    absorbed = true;
                                else
                                {
                                    const V3    f0 = colour * mat.reflectiveness;
                                    const float f  = pow5(1.0f - VoH);
                                    const V3    F  = v3(f + f0.x * (1.0f - f), f + f0.y * (1.0f - f), f + f0.z * (1.0f - f));    // sampling.h:137-141
                                    const float g  = specular_g(NoV, NoL, a) * 4.0f * VoH * NoL / NoH;
                                    weight         = F * g;
                                    no             = sf.point + ns * 0.0001f;
                                    nd             = wi;
                                }
                            }
                            else
                            {
                                no            = sf.point + ns * 0.0001f;
                                nd            = normalize(ns + sample_sphere(u0, u1));
                                next_specular = false;
                            }

                            // next-event estimation at diffuse vertices: the sun cone or one emissive triangle
                            if (!next_specular && (sc.sun.enabled || NL))
                            {
                                const V3    so   = sf.point + ns * 0.001f;
                                const V3    bsdf = (thr * colour) * INV_PI_F;
                                const bool  both = sc.sun.enabled && NL;
                                const bool  use_sun = sc.sun.enabled && (!NL || rnd(key, dim + 4) < 0.5f);
                                const float u2 = rnd(key, dim + 2), u3 = rnd(key, dim + 3);
                                V3          contrib = v3(0.f, 0.f, 0.f), dir = v3(0.f, 1.f, 0.f);
                                float       tmax = inf_f();
                                if (use_sun)
                                {
                                    const V3     l = map_to_solid_angle(u2, u3, sc.sun.one_minus_cos);
                                    const float *T = sc.sun.transform;
                                    dir            = (v3(T[0], T[1], T[2]) * l.x + v3(T[3], T[4], T[5]) * l.y) + v3(T[6], T[7], T[8]) * l.z;
                                    const float cosine    = clampf(dot(ns, dir), 0.0f, 1.0f);
                                    const float sun_angle = acosf(dot(dir, -v3(sc.sun.dir[0], sc.sun.dir[1], sc.sun.dir[2])));
                                    const V3    sky = (sun_angle < sc.sun.size) ? v3(sc.sun.colour[0], sc.sun.colour[1], sc.sun.colour[2]) * sc.sun.intensity
                                                                                : v3(0.f, 0.f, 0.f);
                                    contrib         = bsdf * cosine * sky / sc.sun.pdf;
                                }
                                else
                                {
                                    uint32_t k = uint32_t(rnd(key, dim + 5) * float(NL));
                                    if (k >= NL) k = NL - 1;
                                    const float4 l0 = __ldg(sc.lights + 3 * size_t(k)), l1 = __ldg(sc.lights + 3 * size_t(k) + 1), l2 = __ldg(sc.lights + 3 * size_t(k) + 2);
                                    const V3     v0 = v3(l0.x, l0.y, l0.z), e1 = v3(l1.x, l1.y, l1.z), e2 = v3(l2.x, l2.y, l2.z), le = v3(l0.w, l1.w, l2.w);
                                    float        b1 = u2, b2 = u3;
                                    if (b1 + b2 > 1.0f) b1 = 1.0f - b1, b2 = 1.0f - b2;
                                    const V3    q  = (v0 + e1 * b1) + e2 * b2;
                                    const V3    wv = q - so;
                                    const float d2 = dot(wv, wv);
                                    if (d2 > 0.0f)
                                    {
                                        const float dist  = sqrtf(d2);
                                        dir               = wv * (1.0f / dist);
                                        const float cos_s = clampf(dot(ns, dir), 0.0f, 1.0f);
                                        const float g     = cos_s * (0.5f * fabsf(dot(cross(e1, e2), dir))) / d2 * float(NL);
                                        contrib           = bsdf * le * g;
                                        tmax              = dist * 0.999f;
                                    }
                                }
                                if (both) contrib = contrib * 2.0f;
                                if (contrib.x != 0.f || contrib.y != 0.f || contrib.z != 0.f)
                                {
                                    want_shadow = true;
                                    sr.o        = make_float4(so.x, so.y, so.z, __uint_as_float(slot));
                                    sr.d        = make_float4(dir.x, dir.y, dir.z, tmax);
                                    sr.c        = make_float4(contrib.x, contrib.y, contrib.z, 0.f);
                                }
                            }
                            if (!absorbed)
                            {
                                const V3 t  = thr * weight;
                                out_t   = make_float4(t.x, t.y, t.z, next_specular ? 1.f : 0.f);
                                out_o   = make_float4(no.x, no.y, no.z, 0.f);
                                out_d   = make_float4(nd.x, nd.y, nd.z, 0.f);
                                survive = true;
                            }
                        }
                    }
                }
                const bool pred[2] = { survive, want_shadow };
                const int  cidx[2] = { CTR_NEXT, CTR_SHADOW };
                uint32_t   at[2];
                block_reserve<2>(pred, ps.counters, cidx, at);
                if (survive)
                {
                    // the next bounce's record, compacted: k_trace and the next k_shade stream it without indirection
                    ps.q_next[at[0]]   = slot;
                    ps.ray_o_next[at[0]] = out_o, ps.ray_d_next[at[0]] = out_d, ps.thr_next[at[0]] = out_t;
                }
                if (want_shadow) ps.shadow[at[1]] = sr;
            }
        }

        // scene::cast_ray for ONE ray (the alpha cut-out shadow march): the flat BVH, or — two-level scenes — the reference's
        // own loop over (model, instance) pairs (scene.cpp:79-98, model.cpp:99-126), nearest by re-measured world distance
        template<bool COUNT>
        __device__ __forceinline__ Hit cast_closest(const DScene &sc, V3 o, V3 d, V3 dn, TravCounters *tc)
        {
            if (!sc.two_level) return traverse<false, COUNT>(sc.bvh, o, dn, 0.00001f, inf_f(), tc);
            Hit   best { inf_f(), 0.f, 0.f, INVALID_PRIM };
            float best_dist = inf_f();
            for (uint32_t k = 0; k < sc.bvh2.n_inst; k++)
            {
                const Instance &I  = sc.bvh2.inst[k];
                const V3        oo = xf34(I.inv, o, 1.0f), dd = normalize(xf34(I.inv, d, 0.0f));
                const Blas      bl = sc.bvh2.blas[I.blas];
                const Bvh8      view { sc.bvh2.nodes + size_t(bl.node_base) * BVH8_NODE_U4, sc.bvh2.tris + size_t(bl.tri_base) * BVH8_TRI_F4, bl.n_nodes, bl.n_tris };
                const Hit       h = traverse<false, COUNT>(view, oo, dd, 0.00001f, inf_f(), tc);
                if (h.prim == INVALID_PRIM) continue;
                const float dist = length(xf34(I.fwd, oo + dd * h.t, 1.0f) - o);
                if (dist < best_dist) best = Hit { h.t, h.u, h.v, I.flat_start + h.prim }, best_dist = dist;
            }
            return best;
        }

        // ------------------------------------------------------------------ two-level variants of the two traversal kernels
#ifndef CRB_TRACE2_STEPS
#define CRB_TRACE2_STEPS TRACE_STEPS    // node iterations between two refill points of the two-level loop
#endif
#ifndef CRB_TRACE2_OCC
#define CRB_TRACE2_OCC 4    // resident CTAs per SM of the two-level traversal kernels: 64 registers without spills since the TLAS-level ray
                            // lives in shared memory (bvh8.cuh CRB_2L_SMEM); measured on config 4: 3 CTAs 2086-2115, 4 CTAs 2331 Mrays/s
#endif
        // FB = true: the rays the instance wavefront could not take (more than IW_K candidate instances), by their queue
        template<bool COUNT, bool FB = false>
        __global__ void __launch_bounds__(256, CRB_TRACE2_OCC) k_trace2(DScene sc, PathState ps)
        {
            const uint32_t n = ps.counters[FB ? CTR_IW_FB : CTR_IN];
            TravCounters   tc;
            auto source = [&](uint32_t idx, uint32_t &slot, V3 &o, V3 &d, float &tmin, float &tmax) {
                if (FB) idx = ps.iw_fb[idx];
                slot            = idx;
                const float4 ro = ld_stream(ps.ray_o + idx), rd = ld_stream(ps.ray_d + idx);
                o = v3(ro.x, ro.y, ro.z), d = v3(rd.x, rd.y, rd.z);    // as the reference holds it: every instance renormalises (model.cpp:110-112)
                tmin = 0.00001f, tmax = inf_f();
            };
            auto sink = [&](bool valid, uint32_t slot, const Hit &h) {
                if (valid) st_stream(ps.hit + slot, make_float4(h.t, h.u, h.v, __uint_as_float(h.prim)));
            };
            trace_persistent_2l<COUNT, CRB_TRACE2_STEPS, true>(sc.bvh2, ps.counters + (FB ? CTR_IW_FB_CUR : CTR_CUR_TRACE), n, false, source, sink, &tc);
            if (COUNT)
            {
                atomicAdd(ps.stats + ST_NODES, tc.nodes);
                atomicAdd(ps.stats + ST_TRIS, tc.tris);
            }
        }

        // ------------------------------------------------------------------ two-level closest hit as a wavefront over instance visits
        // The persistent two-level loop (k_trace2) pays for instance entries and exits inside its inner loop: ~290 + ~70
        // instructions that some lane of a warp needs in almost every iteration, executed at 2 - 7 of 32 lanes (ncu source page:
        // a third of the kernel's issue slots; 28 G node visits/s against 52 G in the single-level loop on the same geometry).
        // Here every instance visit is a work item of its own:
        //   k_iw_candidates  one thread per ray: walks the TLAS, keeps the (at most IW_K) instances whose world bounds the ray
        //                    enters, sorted by entry distance, takes the ray into the nearest one's object space and emits
        //                    the item; a ray with more candidates goes to the fallback queue (k_trace2 over that queue)
        //   k_iw_trace       the SINGLE-LEVEL persistent loop over the items (object-space rays against their BLAS)
        //   k_iw_next        one thread per item: re-measures the hit in world space and merges it into the ray's best
        //                    (model.cpp:116-123), then emits the ray's next candidate unless its bounds start beyond the best
        //                    hit, or writes the ray's final hit record
        // for IW_K rounds. Same per-instance arithmetic, same candidate comparison (world distance, ties to the first
        // (model, instance)) and the same conservative pruning as trace_persistent_2l: hits are bit-identical.
        constexpr int IW_K = 8;
#ifndef CRB_IW_DEFAULT
#define CRB_IW_DEFAULT 1    // measured on config 4 (profiles/r2_sweeps.md section 13): 2315 -> 2603 Mrays/s, k_trace2 2104 -> 2489 closest-hit Mrays/s
#endif

        // entry into instance k (model.cpp:107-112 and the pruning bound of trace_persistent_2l): the object-space item
        constexpr uint32_t IW_MORE = 0x80000000u;    // item flag: the ray has per-ray state (further candidates or a best hit so far)
        __device__ __forceinline__ void iw_emit(const DScene &sc, const PathState &ps, int q, uint32_t at, uint32_t rec, uint32_t k, V3 wo, V3 wd, float best_key, bool more)
        {
            const float4 *ip = reinterpret_cast<const float4 *>(sc.bvh2.inst + k);
            const float4  v0 = __ldg(ip), v1 = __ldg(ip + 1), v2 = __ldg(ip + 2);
            const float   inv[12] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w };
            const V3      o = xf34(inv, wo, 1.0f);
            const V3      d = normalize(xf34(inv, wd, 0.0f));
            float         bound = best_key;
            if (best_key < inf_f())
            {
                const float4 f0 = __ldg(ip + 3), f1 = __ldg(ip + 4), f2 = __ldg(ip + 5);
                const float  fwd[12] = { f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w };
                const V3     wdir = xf34(fwd, d, 0.0f);
                bound             = (best_key / __fsqrt_rn(dot(wdir, wdir))) * 1.0001f;
            }
            ps.iw_item_o[q][at] = make_float4(o.x, o.y, o.z, __uint_as_float(rec));
            ps.iw_item_d[q][at] = make_float4(d.x, d.y, d.z, bound);
            ps.iw_item_k[q][at] = k | (more ? IW_MORE : 0u);
        }

        // entry distance of the ray into a world-space box, or a negative value if it misses (same slab arithmetic and
        // slack as ray_box in bvh8.cuh; NaN - the origin on a face of a flat box - counts as an entry at distance 0)
        __device__ __forceinline__ float iw_box_entry(V3 o, V3 idir, const float *lo, const float *hi, float tmin, float tmax)
        {
            const float ax = (lo[0] - o.x) * idir.x, bx = (hi[0] - o.x) * idir.x, ay = (lo[1] - o.y) * idir.y, by = (hi[1] - o.y) * idir.y;
            const float az = (lo[2] - o.z) * idir.z, bz = (hi[2] - o.z) * idir.z;
            const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
            const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
            if (!(tn == tn) || !(tf == tf)) return 0.0f;
            return tn <= tf * 1.00001f + 1e-30f ? tn : -1.0f;
        }

        // sorted insertion by (entry distance, instance index), fully unrolled so that the lists stay in registers
        __device__ __forceinline__ void iw_insert(float (&ct)[IW_K], uint32_t (&ci)[IW_K], int &nc, float te, uint32_t k)
        {
#pragma unroll
            for (int j = 0; j < IW_K; j++)
            {
                if (j < nc)
                {
                    if (ct[j] > te || (ct[j] == te && ci[j] > k))
                    {
                        const float    t0 = ct[j];
                        const uint32_t k0 = ci[j];
                        ct[j] = te, ci[j] = k, te = t0, k = k0;
                    }
                }
                else if (j == nc)
                    ct[j] = te, ci[j] = k;
            }
            nc++;
        }
        // one instance against the ray: its world bounds, then the list (false: the list is full, the ray needs the fallback)
        __device__ __forceinline__ bool iw_consider(const DScene &sc, uint32_t k, V3 wo, V3 tidir, float (&ct)[IW_K], uint32_t (&ci)[IW_K], int &nc)
        {
            const float4 *ip = reinterpret_cast<const float4 *>(sc.bvh2.inst + k);
            const float4  i6 = __ldg(ip + 6), i7 = __ldg(ip + 7);
            const uint4   i8 = __ldg(reinterpret_cast<const uint4 *>(ip + 8));
            const float   lo[3] = { i6.x, i6.y, i6.z }, hi[3] = { i7.x, i7.y, i7.z };
            const float   te = iw_box_entry(wo, tidir, lo, hi, 0.0f, inf_f());
            if (te < 0.0f || i8.z == 0u) return true;    // missed, or an instance of an empty model
            if (nc == IW_K) return false;
            iw_insert(ct, ci, nc, te, k);
            return true;
        }
        constexpr uint32_t IW_BRUTE = 32;    // up to this many instances every ray tests every instance's bounds (uniform control flow) instead of walking the TLAS

        template<bool COUNT>
        __global__ void __launch_bounds__(256) k_iw_candidates(DScene sc, PathState ps)
        {
            const uint32_t n = ps.counters[CTR_IN];
            unsigned long long nodes = 0;
            // all threads of a block iterate together (the queue reservation is block-collective)
            for (uint32_t tile = blockIdx.x * blockDim.x; tile < n; tile += gridDim.x * blockDim.x)
            {
                const uint32_t rec = tile + threadIdx.x;
                bool           emit = false, fallback = false;
                V3             wo = v3(0, 0, 0), wd = v3(0, 0, 1);
                uint32_t       first = 0;
                bool           more  = false;
                if (rec < n)
                {
                    const float4 ro = ld_stream(ps.ray_o + rec), rd = ld_stream(ps.ray_d + rec);
                    wo = v3(ro.x, ro.y, ro.z), wd = v3(rd.x, rd.y, rd.z);
                    const V3 td = normalize(wd), tidir = v3(safe_rcp(td.x), safe_rcp(td.y), safe_rcp(td.z));
                    float    ct[IW_K];
                    uint32_t ci[IW_K];
                    int      nc = 0;
                    for (int j = 0; j < IW_K; j++) ct[j] = -1.0f, ci[j] = 0u;
                    if (sc.bvh2.n_inst <= IW_BRUTE)
                    {
                        for (uint32_t k = 0; k < sc.bvh2.n_inst; k++)
                            if (!iw_consider(sc, k, wo, tidir, ct, ci, nc)) fallback = true;
                    }
                    else if (sc.bvh2.tlas.n_nodes != 0)
                    {
                        // the TLAS is a few nodes deep: a plain per-thread walk, every leaf proxy = one instance
                        const unsigned oct4 = make_oct4(td);
                        uint2          stack[BVH8_STACK];
                        int            sp = 0;
                        uint2          group = make_uint2(0u, 0x80000000u), tgroup;
                        unsigned       occ;
                        for (;;)
                        {
                            const unsigned node_index = pop_inner(group, oct4);
                            if (group.y) stack[sp++] = group;
                            uint4 n0, n1, n2, n3, n4;
                            load_node(sc.bvh2.tlas.nodes, node_index, n0, n1, n2, n3, n4);
                            if (COUNT) nodes++;
                            node_visit(n0, n1, n2, n3, n4, wo, tidir, oct4, 0.0f, inf_f(), group, tgroup, occ);
                            while (tgroup.y)
                            {
                                const uint32_t k = __float_as_uint(__ldg(sc.bvh2.tlas.tris + size_t(pop_triangle(tgroup, occ)) * BVH8_TRI_F4).w);
                                if (!iw_consider(sc, k, wo, tidir, ct, ci, nc)) fallback = true;
                            }
                            if (group.y == 0u)
                            {
                                if (sp == 0) break;
                                group = stack[--sp];
                            }
                        }
                    }
                    if (!fallback)
                    {
                        static_assert(IW_K == 8, "the candidate lists are stored as two 16-byte vectors each");
                        float4 *pt = reinterpret_cast<float4 *>(ps.iw_cand_t + size_t(rec) * IW_K);
                        uint4  *pi = reinterpret_cast<uint4 *>(ps.iw_cand_i + size_t(rec) * IW_K);
                        // a ray with ONE candidate needs no per-ray state: its item's hit is its hit (k_iw_next); the others keep
                        // their list, and round 0 of k_iw_next initialises (best hit, its world distance, its instance, next candidate)
                        if (nc > 1)
                        {
                            pt[0] = make_float4(ct[0], ct[1], ct[2], ct[3]), pt[1] = make_float4(ct[4], ct[5], ct[6], ct[7]);
                            pi[0] = make_uint4(ci[0], ci[1], ci[2], ci[3]), pi[1] = make_uint4(ci[4], ci[5], ci[6], ci[7]);
                            more = true;
                        }
                        if (nc)
                            emit = true, first = ci[0];
                        else
                            st_stream(ps.hit + rec, make_float4(inf_f(), 0.f, 0.f, __uint_as_float(INVALID_PRIM)));
                    }
                }
                const bool pred[2] = { emit, fallback };
                const int  cidx[2] = { CTR_IW_ITEMS, CTR_IW_FB };
                uint32_t   at[2];
                block_reserve<2>(pred, ps.counters, cidx, at);
                if (emit) iw_emit(sc, ps, 0, at[0], rec, first, wo, wd, inf_f(), more);
                if (fallback) ps.iw_fb[at[1]] = rec;
            }
            if (COUNT) atomicAdd(ps.stats + ST_NODES, nodes);
        }

        template<bool COUNT>
        __global__ void __launch_bounds__(CRB_TRACE_BLOCK, CRB_TRACE_OCC) k_iw_trace(DScene sc, PathState ps, int q)
        {
            const uint32_t n = ps.counters[CTR_IW_ITEMS + q];
            TravCounters   tc;
            const Bvh8     all { sc.bvh2.nodes, sc.bvh2.tris, 1u, 1u };    // every BLAS, concatenated; the item names its own
            auto source = [&](uint32_t idx, uint32_t &item, V3 &o, V3 &d, float &tmin, float &tmax, uint32_t &node_off, uint32_t &tri_off) {
                item            = idx;
                const float4 io = ld_stream(ps.iw_item_o[q] + idx), id = ld_stream(ps.iw_item_d[q] + idx);
                const uint4  i8 = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const float4 *>(sc.bvh2.inst + (ps.iw_item_k[q][idx] & ~IW_MORE)) + 8));
                o = v3(io.x, io.y, io.z), d = v3(id.x, id.y, id.z);
                tmin = 0.00001f, tmax = id.w;    // model.cpp:21; the pruning bound of the entry
                node_off = i8.x, tri_off = i8.y;
            };
            auto sink = [&](bool valid, uint32_t item, const Hit &h) {
                if (valid) st_stream(ps.iw_item_hit + item, make_float4(h.t, h.u, h.v, __uint_as_float(h.prim)));
            };
            trace_persistent<COUNT, TRACE_STEPS, true>(all, ps.counters + CTR_IW_CUR, n, ps.trace_chunk, false, source, sink, &tc);
            if (COUNT)
            {
                atomicAdd(ps.stats + ST_NODES, tc.nodes);
                atomicAdd(ps.stats + ST_TRIS, tc.tris);
            }
        }

        __global__ void __launch_bounds__(256) k_iw_next(DScene sc, PathState ps, int q, int round)
        {
            const uint32_t n = ps.counters[CTR_IW_ITEMS + q];
            for (uint32_t tile = blockIdx.x * blockDim.x; tile < n; tile += gridDim.x * blockDim.x)
            {
                const uint32_t idx  = tile + threadIdx.x;
                bool           emit = false;
                uint32_t       rec = 0, next_k = 0;
                V3             wo = v3(0, 0, 0), wd = v3(0, 0, 1);
                float          best_key = inf_f();
                if (idx < n)
                {
                    const float4   io = ld_stream(ps.iw_item_o[q] + idx), h = ld_stream(ps.iw_item_hit + idx);
                    const uint32_t kf = ps.iw_item_k[q][idx], cur = kf & ~IW_MORE;
                    rec               = __float_as_uint(io.w);
                    const float4 *ip  = reinterpret_cast<const float4 *>(sc.bvh2.inst + cur);
                    if (!(kf & IW_MORE))
                    {
                        // the ray's only candidate: no comparison, no re-measuring — its hit is the hit
                        const uint32_t prim = __float_as_uint(h.w);
                        st_stream(ps.hit + rec, prim == INVALID_PRIM ? make_float4(inf_f(), 0.f, 0.f, h.w) : make_float4(h.x, h.y, h.z, __uint_as_float(__float_as_uint(__ldg(ip + 7).w) + prim)));
                    }
                    else
                    {
                    const float4 id = ld_stream(ps.iw_item_d[q] + idx);
                    const float4 ro = ps.ray_o[rec], rd = ps.ray_d[rec];
                    wo = v3(ro.x, ro.y, ro.z), wd = v3(rd.x, rd.y, rd.z);
                    float4   best = make_float4(inf_f(), 0.f, 0.f, __uint_as_float(INVALID_PRIM)), meta = make_float4(inf_f(), __uint_as_float(0xffffffffu), __uint_as_float(1u), 0.f);
                    if (round != 0) best = ps.iw_best[rec], meta = ps.iw_meta[rec];
                    uint32_t best_k = __float_as_uint(meta.y), nxt = __float_as_uint(meta.z);
                    best_key        = meta.x;
                    if (__float_as_uint(h.w) != INVALID_PRIM)
                    {
                        // leave the instance: model.cpp:116-123 (map the point back, re-measure, keep the nearest)
                        const float4  f0 = __ldg(ip + 3), f1 = __ldg(ip + 4), f2 = __ldg(ip + 5);
                        const float   fwd[12] = { f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w };
                        const V3      o = v3(io.x, io.y, io.z), d = v3(id.x, id.y, id.z);
                        const float   key = length(xf34(fwd, o + d * h.x, 1.0f) - wo);    // glm::distance(point, ray.origin)
                        if (key < best_key || (key == best_key && cur < best_k))
                        {
                            best     = make_float4(h.x, h.y, h.z, __uint_as_float(__float_as_uint(__ldg(ip + 7).w) + __float_as_uint(h.w)));
                            best_key = key, best_k = cur;
                        }
                    }
                    // the next candidate, unless its bounds begin beyond the best hit (the candidates are sorted by entry distance)
                    if (nxt < uint32_t(IW_K))
                    {
                        const float te = ps.iw_cand_t[size_t(rec) * IW_K + nxt];
                        if (te >= 0.0f && te <= best_key * 1.00001f + 1e-30f) emit = true, next_k = ps.iw_cand_i[size_t(rec) * IW_K + nxt], nxt++;
                    }
                    if (emit)
                    {
                        ps.iw_best[rec] = best;
                        ps.iw_meta[rec] = make_float4(best_key, __uint_as_float(best_k), __uint_as_float(nxt), 0.f);
                    }
                    else
                        st_stream(ps.hit + rec, __float_as_uint(best.w) == INVALID_PRIM ? make_float4(inf_f(), 0.f, 0.f, best.w) : best);
                    }
                }
                const bool pred[1] = { emit };
                const int  cidx[1] = { CTR_IW_ITEMS + (q ^ 1) };
                uint32_t   at[1];
                block_reserve<1>(pred, ps.counters, cidx, at);
                if (emit) iw_emit(sc, ps, q ^ 1, at[0], rec, next_k, wo, wd, best_key, true);
            }
        }

        // one thread: the consumed item queue becomes the next round's target, the cursor of the item loop is reset
        __global__ void k_iw_roll(PathState ps, int q)
        {
            ps.counters[CTR_IW_ITEMS + q] = 0;
            ps.counters[CTR_IW_CUR]       = 0;
        }

        template<bool COUNT>
        __global__ void __launch_bounds__(256, CRB_TRACE2_OCC) k_shadow2(DScene sc, PathState ps)
        {
            const uint32_t n = ps.counters[CTR_SHADOW];
            TravCounters   tc;
            auto source = [&](uint32_t idx, uint32_t &item, V3 &o, V3 &d, float &tmin, float &tmax) {
                item            = idx;
                const float4 so = ld_stream(&ps.shadow[idx].o), sd = ld_stream(&ps.shadow[idx].d);
                o = v3(so.x, so.y, so.z), d = v3(sd.x, sd.y, sd.z);
                tmin = 0.00001f, tmax = sd.w;    // inf for the sun, 0.999 * distance (world) for an area light
            };
            auto sink = [&](bool valid, uint32_t item, const Hit &h) {
                if (valid && h.prim == INVALID_PRIM)
                {
                    const uint32_t slot = __float_as_uint(ps.shadow[item].o.w);
                    const float4   c    = ps.shadow[item].c;
                    const float4   r4   = ps.rad[slot];
                    const V3       r    = v3(r4.x, r4.y, r4.z) + v3(c.x, c.y, c.z);
                    ps.rad[slot]        = make_float4(r.x, r.y, r.z, 0.f);
                }
            };
            trace_persistent_2l<COUNT, CRB_TRACE2_STEPS, true>(sc.bvh2, ps.counters + CTR_CUR_SHADOW, n, true, source, sink, &tc);
            if (COUNT)
            {
                atomicAdd(ps.stats + ST_NODES_SHADOW, tc.nodes);
                atomicAdd(ps.stats + ST_TRIS_SHADOW, tc.tris);
            }
        }

        // ------------------------------------------------------------------ K8 shadow rays
        // sun visibility without alpha cut-outs: any-hit through the persistent trace loop
        template<bool COUNT, int STEPS>
        __global__ void __launch_bounds__(CRB_TRACE_BLOCK, CRB_TRACE_OCC) k_shadow(DScene sc, PathState ps)
        {
            const uint32_t n = ps.counters[CTR_SHADOW];
            TravCounters   tc;
            auto source = [&](uint32_t idx, uint32_t &item, V3 &o, V3 &d, float &tmin, float &tmax) {
                item            = idx;
                const float4 so = ld_stream(&ps.shadow[idx].o), sd = ld_stream(&ps.shadow[idx].d);
                o               = v3(so.x, so.y, so.z);
                d               = normalize(v3(sd.x, sd.y, sd.z));    // model.cpp:110-112
                tmin = 0.00001f, tmax = sd.w;                         // inf for the sun, 0.999 * distance for an area light
            };
#ifdef CRB_EMU
            RayClassProbe probe;
#endif
            auto sink = [&](bool valid, uint32_t item, const Hit &h) {
#ifdef CRB_EMU
                if (valid && COUNT) probe.retire(1, h.prim != INVALID_PRIM, tc);
#endif
                if (valid && h.prim == INVALID_PRIM)
                {
                    // renderer.cpp:348-353: the sun is visible, connect
                    const uint32_t slot = __float_as_uint(ps.shadow[item].o.w);
                    const float4   c    = ps.shadow[item].c;
                    const float4   r4   = ps.rad[slot];
                    const V3       r    = v3(r4.x, r4.y, r4.z) + v3(c.x, c.y, c.z);
                    ps.rad[slot]        = make_float4(r.x, r.y, r.z, 0.f);
                }
            };
            trace_persistent<COUNT, STEPS>(sc.bvh, ps.counters + CTR_CUR_SHADOW, n, ps.trace_chunk, true, source, sink, &tc);
            if (COUNT)
            {
                atomicAdd(ps.stats + ST_NODES_SHADOW, tc.nodes);
                atomicAdd(ps.stats + ST_TRIS_SHADOW, tc.tris);
            }
        }

        // sun visibility with alpha cut-outs present: the reference's closest-hit march (renderer.cpp:330-345)
        template<bool COUNT>
        __global__ void __launch_bounds__(256) k_shadow_alpha(DScene sc, PathState ps)
        {
            const uint32_t n = ps.counters[CTR_SHADOW];
            TravCounters   tc;
            for (;;)
            {
                const uint32_t base = warp_fetch(ps.counters + CTR_CUR_SHADOW);
                if (base >= n) break;
                const uint32_t idx = base + crb_lane_id();
                if (idx < n)
                {
                    const ShadowRay sr = ps.shadow[idx];
                    V3              o  = v3(sr.o.x, sr.o.y, sr.o.z);
                    const V3        d  = v3(sr.d.x, sr.d.y, sr.d.z);
                    const V3        dn = normalize(d);    // model.cpp:110-112
                    bool            visible = false;
                    float           remaining = sr.d.w;    // inf for the sun (the reference's loop), finite for area lights (extended mode)
                    for (int guard = 0; guard < 4096; guard++)
                    {
                        const Hit h = cast_closest<COUNT>(sc, o, d, dn, &tc);
                        if (h.prim == INVALID_PRIM)
                        {
                            visible = true;
                            break;
                        }
                        const Surface   sf  = surface_at(sc, o, d, make_float4(h.t, h.u, h.v, __uint_as_float(h.prim)));
                        if (!(sf.distance <= remaining))
                        {
                            visible = true;
                            break;
                        }
                        const DMaterial mat = sc.materials[sf.mat];
                        if (surface_colour(sc, mat, sf).w != 0.0f) break;
                        const V3 next = sf.point + d * 0.1f;    // marches in 0.1 steps of the un-normalised direction
                        remaining     = remaining - length(next - o);
                        o             = next;
                    }
                    if (visible)
                    {
                        const uint32_t slot = __float_as_uint(sr.o.w);
                        const float4   r4   = ps.rad[slot];
                        const V3       r    = v3(r4.x, r4.y, r4.z) + v3(sr.c.x, sr.c.y, sr.c.z);
                        ps.rad[slot]        = make_float4(r.x, r.y, r.z, 0.f);
                    }
                }
            }
            if (COUNT)
            {
                atomicAdd(ps.stats + ST_NODES_SHADOW, tc.nodes);
                atomicAdd(ps.stats + ST_TRIS_SHADOW, tc.tris);
            }
        }

        // one thread: roll the queue counters over to the next bounce and keep the ray statistics
        __global__ void k_advance(PathState ps, int last)
        {
            if (blockIdx.x * blockDim.x + threadIdx.x != 0) return;
            uint32_t *c = ps.counters;
            ps.stats[ST_CLOSEST] += c[CTR_IN];
            ps.stats[ST_SHADOW] += c[CTR_SHADOW];
            if (last) ps.stats[ST_RANOUT] += c[CTR_NEXT];
            c[CTR_IN] = c[CTR_NEXT];
            c[CTR_NEXT] = 0, c[CTR_SHADOW] = 0;
            c[CTR_CLASS0] = c[CTR_CLASS0 + 1] = c[CTR_CLASS0 + 2] = c[CTR_CLASS0 + 3] = 0;
            c[CTR_CUR_TRACE] = c[CTR_CUR_SHADE] = c[CTR_CUR_SHADOW] = 0;
#ifdef CRB_EMU
            if (getenv("CRB_IW_DEBUG")) fprintf(stderr, "iw: %u rays of this bounce went to the fallback queue\n", unsigned(c[CTR_IW_FB]));    // kernel-logic harness only
#endif
            c[CTR_IW_ITEMS] = c[CTR_IW_ITEMS + 1] = c[CTR_IW_CUR] = c[CTR_IW_FB] = c[CTR_IW_FB_CUR] = 0;
        }

        // ------------------------------------------------------------------ K9 accumulate + resolve
        // accum.w is the PER-PIXEL pass count: a render call may cover a row band only (crb_render_set_rows, the tile
        // partition), so one counter per renderer would divide later bands by the wrong count and a RAW_SUM read
        // would not be a complete checkpoint. The display is resolved with the pixel's own count.
        __global__ void __launch_bounds__(256) k_accumulate(RenderParams rp, PathState ps)
        {
            const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
            if (pix >= rp.npix) return;
            const uint32_t x = pix % rp.w, y = row_of(rp, pix / rp.w);
            const uint32_t fi = flipped_index(rp, x, y);
            float4         a  = rp.accum[fi];
            for (uint32_t s = 0; s < rp.batch; s++)
            {
                const float4 r = ps.rad[size_t(s) * rp.npix + pix];    // coalesced float4, sample order = the reference's pass order
                a.x += r.x, a.y += r.y, a.z += r.z;
            }
            a.w += float(rp.batch);
            rp.accum[fi]   = a;
            rp.display[fi] = resolve_px(a, a.w);
        }

        __global__ void __launch_bounds__(256) k_resolve(const float4 *__restrict__ accum, float4 *__restrict__ display, uint32_t n)
        {
            const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= n) return;
            const float4 a = accum[i];
            if (a.w > 0.0f) display[i] = resolve_px(a, a.w);    // pixels without a sample keep cr::image's FLT_MAX fill
        }

        __global__ void __launch_bounds__(256) k_set_pass_count(float4 *__restrict__ accum, uint32_t n, float passes)
        {
            const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i < n) accum[i].w = passes;
        }

        __global__ void k_fill4(float4 *p, uint32_t n, float4 v)
        {
            const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i < n) p[i] = v;
        }
    }    // namespace

    // ====================================================================== host side
    Render::Render(Scene *s, uint32_t w_, uint32_t h_, uint32_t mb, uint32_t seed_, uint32_t flags_)
        : scene(s), w(w_), h(h_), max_bounces(mb), seed(seed_), flags(flags_), row0(0), row1(h_)
    {
#ifndef CRB_EMU
        // (the C ABI runs every call under the handle's device, capi.cu DeviceScope)
        n_sms = s->n_sms;    // (cudaGetDeviceProperties costs tens of milliseconds; the scene already asked)
        // experiment knob: keep the context's local-memory reservation at its high-water mark (the traversal stack lives in
        // local memory; by default the driver may shrink the reservation when the device idles and re-grow it at a launch)
        static const bool lmem_max = getenv("CRB_LMEM_MAX") && atoi(getenv("CRB_LMEM_MAX"));
        if (lmem_max && cudaSetDeviceFlags(cudaDeviceLmemResizeToMax) != cudaSuccess) cudaGetLastError();
#endif
        counters.alloc(CTR_COUNT);
        dstats.alloc(ST_COUNT);
        alloc_images();
        reset();
    }

    Render::~Render()
    {
#ifndef CRB_EMU
        cudaStreamSynchronize(stream());
        if (copy_stream) cudaStreamSynchronize(copy_stream), cudaStreamDestroy(copy_stream);
        if (snap_done) cudaEventDestroy(snap_done);
        for (cudaEvent_t e : copy_done)
            if (e) cudaEventDestroy(e);
        for (const Span &sp : spans) cudaEventDestroy(sp.a), cudaEventDestroy(sp.b);
        for (const Timed &t : timed) cudaEventDestroy(t.a), cudaEventDestroy(t.b);
        for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
#endif
    }

    void Render::alloc_images()
    {
        const size_t n = size_t(w) * h;
        accum.alloc(n), display.alloc(n), albedo.alloc(n), normal.alloc(n), depth.alloc(n);
    }

    void Render::reset()
    {
        // renderer::start (renderer.cpp:154-170) zeroes _raw_buffer and clears _buffer; cr::image is
        // FLT_MAX-filled on construction/clear (image.h:30-38)
        collect_time();
        const uint32_t n = w * h;
        const float    mx = 3.402823466e+38f;
        const float4   fm = make_float4(mx, mx, mx, mx);
        dev_zero(accum.p, size_t(n) * 16, stream());
        const unsigned g = (n + 255) / 256;
        CRB_LAUNCH(k_fill4, g, 256, stream(), display.p, n, fm);
        CRB_LAUNCH(k_fill4, g, 256, stream(), albedo.p, n, fm);
        CRB_LAUNCH(k_fill4, g, 256, stream(), normal.p, n, fm);
        CRB_LAUNCH(k_fill4, g, 256, stream(), depth.p, n, fm);
        dev_zero(dstats.p, ST_COUNT * 8, stream());
        dev_zero(counters.p, CTR_COUNT * 4, stream());
        passes = 0, pass_px = 0, device_ms = 0, launches = 0, pixel_samples = 0;
        for (int i = 0; i < 8; i++) kernel_ms[i] = 0, kernel_count[i] = 0;
    }

    void Render::set_resolution(uint32_t w_, uint32_t h_)
    {
        sync();
        w = w_, h = h_, row0 = 0, row1 = h_, band = 0;
        alloc_images();
        scene_version = ~0ull;
        reset();
    }

    void Render::set_rows(uint32_t y0, uint32_t y1)
    {
        if (y1 > h) y1 = h;
        if (y0 >= y1) throw Error(ERR_INVALID_ARG, "set_rows: empty row range");
        row0 = y0, row1 = y1;
        band = 0;
    }

    void Render::set_bands(uint32_t band_rows, uint32_t first, uint32_t stride, bool serpentine)
    {
        if (band_rows == 0 || stride == 0 || first >= stride) throw Error(ERR_INVALID_ARG, "set_bands: need band_rows > 0 and first < stride");
        // this handle's band of period p is p * stride + band_slot(p): only the last period can lack it, so the local band
        // index is the period and the (possibly partial) last band of the frame is its owner's last local band
        uint32_t rows = 0;
        for (uint32_t p = 0; uint64_t(p) * stride * band_rows < h; p++)
        {
            const uint64_t y0 = (uint64_t(p) * stride + band_slot(p, first, stride, serpentine ? 1u : 0u)) * band_rows;
            if (y0 < h) rows += uint32_t(std::min<uint64_t>(band_rows, h - y0));
        }
        // rows == 0 is legal: more ranks than bands, this rank renders nothing
        band = band_rows, band_first = first, band_stride = stride, band_nrows = rows, band_serp = serpentine ? 1u : 0u;
        row0 = 0, row1 = h;
    }

    void Render::refresh()
    {
        scene->require_committed();
        dscene        = scene->device_scene(w, h);
        scene_version = scene->version;
    }

    void Render::ensure_paths(size_t n)
    {
        if (n <= capacity) return;
        sync();
        ray_o.alloc(n), ray_d.alloc(n), thr.alloc(n), rad.alloc(n), hit.alloc(n);
        ray_o2.alloc(n), ray_d2.alloc(n), thr2.alloc(n);
        q_in.alloc(n), q_next.alloc(n);
        for (auto &q : q_class) q.alloc(n);
        shadow.alloc(n);
        capacity = n;
    }

    // state of the instance wavefront (two-level scenes only): 188 B per path on top of the 200
    void Render::ensure_iw(size_t n)
    {
        if (n <= iw_capacity) return;
        sync();
        iw_cand_t.alloc(n * 8), iw_cand_i.alloc(n * 8), iw_best.alloc(n), iw_meta.alloc(n), iw_item_hit.alloc(n), iw_fb.alloc(n);
        for (int q = 0; q < 2; q++) iw_item_o[q].alloc(n), iw_item_d[q].alloc(n), iw_item_k[q].alloc(n);
        iw_capacity = n;
    }

    void Render::collect_time(bool wait)
    {
#ifndef CRB_EMU
        // spans complete in stream order: resolve the finished prefix, and everything when asked to wait
        size_t done = 0;
        for (; done < spans.size(); done++)
        {
            const Span &sp = spans[done];
            if (wait)
                CRB_CUDA_CHECK(cudaEventSynchronize(sp.b));
            else
            {
                const cudaError_t q = cudaEventQuery(sp.b);
                if (q == cudaErrorNotReady) break;
                CRB_CUDA_CHECK(q);
            }
            float ms = 0;
            CRB_CUDA_CHECK(cudaEventElapsedTime(&ms, sp.a, sp.b));
            device_ms += ms;
            ev_pool.push_back(sp.a), ev_pool.push_back(sp.b);
        }
        spans.erase(spans.begin(), spans.begin() + done);
        if (!wait) return;
        for (const Timed &t : timed)
        {
            CRB_CUDA_CHECK(cudaEventSynchronize(t.b));
            float ms = 0;
            CRB_CUDA_CHECK(cudaEventElapsedTime(&ms, t.a, t.b));
            kernel_ms[t.cls] += ms;
            kernel_count[t.cls]++;
            ev_pool.push_back(t.a), ev_pool.push_back(t.b);
        }
        timed.clear();
#endif
    }

#ifndef CRB_EMU
    cudaEvent_t Render::take_event()
    {
        if (!ev_pool.empty())
        {
            cudaEvent_t e = ev_pool.back();
            ev_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        CRB_CUDA_CHECK(cudaEventCreate(&e));
        return e;
    }
#endif
    void Render::tick(int cls)
    {
#ifndef CRB_EMU
        if (!(flags & CRB_RENDER_FLAG_TIMERS)) return;
        Timed t { cls, take_event(), take_event() };
        CRB_CUDA_CHECK(cudaEventRecord(t.a, stream()));
        timed.push_back(t);
#else
        (void) cls;
#endif
    }
    void Render::tock()
    {
#ifndef CRB_EMU
        if (!(flags & CRB_RENDER_FLAG_TIMERS)) return;
        CRB_CUDA_CHECK(cudaEventRecord(timed.back().b, stream()));
#endif
    }

    void Render::render_samples(uint32_t first, uint32_t n)
    {
        if (n == 0) return;
        if (scene_version != scene->version) refresh();
        collect_time(false);
        const uint32_t nrows = band ? band_nrows : row1 - row0;
        const uint32_t npix  = w * nrows;
        if (npix == 0) return;
        // CRB_SUBMIT_DEBUG=1: host-clock breakdown of this call's submission on stderr (measurement hook)
        static const bool submit_debug = getenv("CRB_SUBMIT_DEBUG") && atoi(getenv("CRB_SUBMIT_DEBUG"));
        const auto        now          = [] { return std::chrono::steady_clock::now(); };
        const auto        ms_since     = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
        const auto        t_enter      = now();
        double            ms_meminfo = 0, ms_ensure = 0, ms_first_launch = 0;
        static const size_t tp_env = getenv("CRB_TARGET_PATHS") ? size_t(atoll(getenv("CRB_TARGET_PATHS"))) : 0;    // tuning knob
        size_t         tpaths    = tp_env ? tp_env : target_paths;
#ifndef CRB_EMU
        {
            // never plan for more than a quarter of the free device memory (200 B of state per path)
            const auto t_a = now();
            if (capacity == 0)
                if (const size_t avail = dev_available_bytes()) mem_path_cap = std::max<size_t>(size_t(1) << 20, avail / 4 / (dscene.two_level ? 400 : 200));    // + 188 B for the instance wavefront
            if (mem_path_cap) tpaths = std::min(tpaths, mem_path_cap);
            ms_meminfo = ms_since(t_a);
        }
#endif
        uint32_t       spp_batch = uint32_t(std::max<size_t>(1, tpaths / npix));
        spp_batch                = std::min(spp_batch, n);
        {
            const auto t_a = now();
            ensure_paths(size_t(npix) * spp_batch);
            ms_ensure = ms_since(t_a);
        }

        PathState ps {};
        ps.ray_o = ray_o.p, ps.ray_d = ray_d.p, ps.thr = thr.p, ps.rad = rad.p, ps.hit = hit.p;
        ps.ray_o_next = ray_o2.p, ps.ray_d_next = ray_d2.p, ps.thr_next = thr2.p;
        ps.q_in = q_in.p, ps.q_next = q_next.p;
        for (int c = 0; c < 4; c++) ps.q_class[c] = q_class[c].p;
        ps.shadow = shadow.p, ps.counters = counters.p, ps.stats = dstats.p;
        // closest hit of two-level scenes as a wavefront over instance visits (k_iw_*): CRB_INSTANCE_WAVEFRONT=0 keeps the
        // persistent two-level loop for everything
        static const int iw_env = getenv("CRB_INSTANCE_WAVEFRONT") ? atoi(getenv("CRB_INSTANCE_WAVEFRONT")) : CRB_IW_DEFAULT;
        const bool iw = dscene.two_level && iw_env != 0;
        if (iw) ensure_iw(capacity);
        ps.iw_cand_t = iw_cand_t.p, ps.iw_cand_i = iw_cand_i.p, ps.iw_best = iw_best.p, ps.iw_meta = iw_meta.p, ps.iw_item_hit = iw_item_hit.p, ps.iw_fb = iw_fb.p;
        for (int q = 0; q < 2; q++) ps.iw_item_o[q] = iw_item_o[q].p, ps.iw_item_d[q] = iw_item_d[q].p, ps.iw_item_k[q] = iw_item_k[q].p;
        static const uint32_t trace_chunk = getenv("CRB_TRACE_CHUNK") ? uint32_t(atoi(getenv("CRB_TRACE_CHUNK"))) : 0u;    // tuning knob
        ps.trace_chunk = trace_chunk;
        // material sort before shading: implemented (k_classify + per-class queues) but OFF by default — the
        // reference's shading is a few dozen instructions, k_shade is HBM-bound, and shading in queue (= screen)
        // order keeps its gathers coalesced: measured 2598 vs 2409 Mrays/s (profiles/r1c_sweeps.md §12)
        static const int sort_env = getenv("CRB_SORT") ? atoi(getenv("CRB_SORT")) : -1;
        ps.sorted = sort_env >= 0 ? sort_env : ((flags & CRB_RENDER_FLAG_MATERIAL_SORT) ? 1 : 0);

        RenderParams rp {};
        rp.w = w, rp.h = h, rp.row0 = row0, rp.nrows = nrows, rp.npix = npix, rp.seed = seed;
        rp.band = band, rp.band_first = band_first, rp.band_stride = band_stride, rp.band_serp = band_serp;
        rp.table = sample_table.p, rp.table_samples = sample_table.p ? table_samples : 0u, rp.table_dims = table_dims;
        rp.aov_sample = first + n - 1;
        rp.accum = accum.p, rp.display = display.p, rp.albedo = albedo.p, rp.normal = normal.p, rp.depth = depth.p;

        const bool count = (flags & CRB_RENDER_FLAG_COUNTERS) != 0;
        static const int steps = getenv("CRB_TRACE_STEPS") ? atoi(getenv("CRB_TRACE_STEPS")) : TRACE_STEPS;    // tuning knob
        static const int ssteps = getenv("CRB_SHADOW_STEPS") ? atoi(getenv("CRB_SHADOW_STEPS")) : steps;    // ... of the shadow-ray loop alone
#ifdef CRB_EMU
        const unsigned pgrid = 1, pblock = 1, tgrid = 1, tblock = 1, t2grid = 1, sgrid = 1, sblock = 1;
#else
        const unsigned pgrid = unsigned(n_sms) * 4, pblock = 256, tgrid = unsigned(n_sms) * CRB_TRACE_OCC, tblock = CRB_TRACE_BLOCK, t2grid = unsigned(n_sms) * CRB_TRACE2_OCC,
                       sgrid = unsigned(n_sms) * (1024 / CRB_SHADE_BLOCK), sblock = CRB_SHADE_BLOCK;
        const Span span { take_event(), take_event() };
        CRB_CUDA_CHECK(cudaEventRecord(span.a, stream()));
#endif
        cudaStream_t st = stream();
        for (uint32_t done = 0; done < n; done += spp_batch)
        {
            const uint32_t b  = std::min(spp_batch, n - done);
            const uint32_t np = npix * b;
            rp.first_sample = first + done, rp.batch = b;
            // queue counters for bounce 0: everything is active
            uint32_t init[CTR_COUNT] = {};
            init[CTR_IN]             = np;
            dev_upload(counters.p, init, sizeof(init), st);
#ifdef CRB_EMU
            CRB_LAUNCH(k_raygen, np, 1, st, dscene, rp, ps);
#else
            tick(CRB_K_RAYGEN);
            const auto t_a = now();
            CRB_LAUNCH(k_raygen, (np + 255) / 256, 256, st, dscene, rp, ps);
            if (done == 0) ms_first_launch = ms_since(t_a);
            tock();
#endif
            launches++;
            for (uint32_t i = 0; i < max_bounces; i++)
            {
                rp.bounce = i;
                tick(CRB_K_TRACE);
                if (dscene.two_level && iw)
                {
                    // closest hit as a wavefront over instance visits: candidates, then IW_K rounds of (single-level loop over the
                    // items, merge + next candidate), then the general loop over the rays with more than IW_K candidates
                    if (count)
                        CRB_LAUNCH((k_iw_candidates<true>), pgrid, pblock, st, dscene, ps);
                    else
                        CRB_LAUNCH((k_iw_candidates<false>), pgrid, pblock, st, dscene, ps);
                    for (int round = 0; round < IW_K; round++)
                    {
                        const int q = round & 1;
                        if (count)
                            CRB_LAUNCH((k_iw_trace<true>), tgrid, tblock, st, dscene, ps, q);
                        else
                            CRB_LAUNCH((k_iw_trace<false>), tgrid, tblock, st, dscene, ps, q);
                        CRB_LAUNCH(k_iw_next, pgrid, pblock, st, dscene, ps, q, round);
                        CRB_LAUNCH(k_iw_roll, 1, 1, st, ps, q);
                    }
                    if (count)
                        CRB_LAUNCH((k_trace2<true, true>), t2grid, pblock, st, dscene, ps);
                    else
                        CRB_LAUNCH((k_trace2<false, true>), t2grid, pblock, st, dscene, ps);
                    launches += 1 + 3 * IW_K;
                }
                else if (dscene.two_level)
                {
                    if (count)
                        CRB_LAUNCH((k_trace2<true>), t2grid, pblock, st, dscene, ps);
                    else
                        CRB_LAUNCH((k_trace2<false>), t2grid, pblock, st, dscene, ps);
                }
                else if (count)
                    CRB_LAUNCH((k_trace<true, TRACE_STEPS>), tgrid, tblock, st, dscene, ps);
                else if (steps == 1)
                    CRB_LAUNCH((k_trace<false, 1>), tgrid, tblock, st, dscene, ps);
                else if (steps == 2)
                    CRB_LAUNCH((k_trace<false, 2>), tgrid, tblock, st, dscene, ps);
                else if (steps == 8)
                    CRB_LAUNCH((k_trace<false, 8>), tgrid, tblock, st, dscene, ps);
                else
                    CRB_LAUNCH((k_trace<false, 4>), tgrid, tblock, st, dscene, ps);
                tock();
                tick(CRB_K_SHADE);
                if (ps.sorted) CRB_LAUNCH(k_classify, pgrid, pblock, st, dscene, ps);
                if (flags & CRB_RENDER_FLAG_EXTENDED)
                    CRB_LAUNCH(k_shade_ext, pgrid, pblock, st, dscene, rp, ps);
                else
                    CRB_LAUNCH(k_shade, sgrid, sblock, st, dscene, rp, ps);
                tock();
                launches += ps.sorted ? 3 : 2;
                if (dscene.sun.enabled || ((flags & CRB_RENDER_FLAG_EXTENDED) && dscene.n_lights))
                {
                    tick(CRB_K_SHADOW);
                    if (dscene.has_alpha)
                    {
                        if (count)
                            CRB_LAUNCH((k_shadow_alpha<true>), pgrid, pblock, st, dscene, ps);
                        else
                            CRB_LAUNCH((k_shadow_alpha<false>), pgrid, pblock, st, dscene, ps);
                    }
                    else if (dscene.two_level)
                    {
                        if (count)
                            CRB_LAUNCH((k_shadow2<true>), t2grid, pblock, st, dscene, ps);
                        else
                            CRB_LAUNCH((k_shadow2<false>), t2grid, pblock, st, dscene, ps);
                    }
                    else if (count)
                        CRB_LAUNCH((k_shadow<true, TRACE_STEPS>), tgrid, tblock, st, dscene, ps);
                    else if (ssteps == 1)
                        CRB_LAUNCH((k_shadow<false, 1>), tgrid, tblock, st, dscene, ps);
                    else if (ssteps == 2)
                        CRB_LAUNCH((k_shadow<false, 2>), tgrid, tblock, st, dscene, ps);
                    else if (ssteps == 8)
                        CRB_LAUNCH((k_shadow<false, 8>), tgrid, tblock, st, dscene, ps);
                    else
                        CRB_LAUNCH((k_shadow<false, 4>), tgrid, tblock, st, dscene, ps);
                    tock();
                    launches++;
                }
                tick(CRB_K_ADVANCE);
                CRB_LAUNCH(k_advance, 1, 1, st, ps, int(i + 1 == max_bounces));
                tock();
                launches++;
                std::swap(ps.q_in, ps.q_next);
                std::swap(ps.ray_o, ps.ray_o_next), std::swap(ps.ray_d, ps.ray_d_next), std::swap(ps.thr, ps.thr_next);
            }
#ifdef CRB_EMU
            CRB_LAUNCH(k_accumulate, npix, 1, st, rp, ps);
#else
            tick(CRB_K_ACCUMULATE);
            CRB_LAUNCH(k_accumulate, (npix + 255) / 256, 256, st, rp, ps);
            tock();
#endif
            launches++;
            pixel_samples += uint64_t(np);
            // _current_sample of the reference counts whole-frame passes; with row bands that is pixel-samples / frame
            pass_px += uint64_t(np);
            passes = uint32_t(pass_px / (uint64_t(w) * h));
        }
#ifndef CRB_EMU
        CRB_CUDA_CHECK(cudaEventRecord(span.b, stream()));
        spans.push_back(span);
#endif
        if (submit_debug)
            fprintf(stderr, "crb submit: total %.2f ms (memory query %.2f, ensure_paths %.2f, first launch %.2f), first_sample %u\n", ms_since(t_enter), ms_meminfo,
                    ms_ensure, ms_first_launch, first);
    }

    void Render::sync()
    {
        stream_sync(stream());
#ifndef CRB_EMU
        if (copy_stream) stream_sync(copy_stream);    // "paused" includes every read_async issued so far
#endif
        collect_time();
    }

    void Render::resolve()
    {
        const uint32_t n = w * h;
#ifdef CRB_EMU
        CRB_LAUNCH(k_resolve, n, 1, stream(), accum.p, display.p, n);
#else
        CRB_LAUNCH(k_resolve, (n + 255) / 256, 256, stream(), accum.p, display.p, n);
#endif
    }

    void Render::set_sample_table(const float *table_host, uint32_t n_samples, uint32_t dims)
    {
        sync();
        if (!table_host || !n_samples || !dims)
        {
            sample_table.release();
            table_samples = table_dims = 0;
            return;
        }
        if (flags & CRB_RENDER_FLAG_EXTENDED) throw Error(ERR_INVALID_ARG, "set_sample_table: the table layout is the ref-exact mode's (2 + 4 dimensions per bounce)");
        const size_t n = size_t(n_samples) * w * h * dims;
        sample_table.alloc(n);
        dev_upload(sample_table.p, table_host, n * sizeof(float), stream());
        stream_sync(stream());
        table_samples = n_samples, table_dims = dims;
    }

    void Render::set_pass_count(uint32_t passes_)
    {
        const uint32_t n = w * h;
#ifdef CRB_EMU
        CRB_LAUNCH(k_set_pass_count, n, 1, stream(), accum.p, n, float(passes_));
#else
        CRB_LAUNCH(k_set_pass_count, (n + 255) / 256, 256, stream(), accum.p, n, float(passes_));
#endif
        passes = passes_, pass_px = uint64_t(passes_) * w * h;
    }

    const float4 *Render::buffer_of(int kind) const
    {
        switch (kind)
        {
        case CRB_RAW_SUM: return accum.p;
        case CRB_PROGRESS: return display.p;
        case CRB_ALBEDO: return albedo.p;
        case CRB_NORMAL: return normal.p;
        case CRB_DEPTH: return depth.p;
        default: throw Error(ERR_INVALID_ARG, "read: unknown buffer kind");
        }
    }

    void Render::read(int kind, float *dst)
    {
        sync();
        dev_download(dst, buffer_of(kind), size_t(w) * h * 16, stream());
    }

    // The reference's UI thread reads the live image buffers while the workers keep rendering
    // (renderer.cpp:220-238, lock-free). Here: a consistent snapshot is taken on the render stream after the work
    // queued so far (device-to-device, 33 MB at 1080p = ~10 us), and its device->host copy runs on a second stream,
    // so the caller can queue the next render_samples right away and the PCIe transfer hides behind it.
    uint64_t Render::read_async(int kind, float *dst)
    {
        const float4 *src   = buffer_of(kind);
        const size_t  bytes = size_t(w) * h * 16;
#ifdef CRB_EMU
        dev_download(dst, src, bytes, stream());
        return next_ticket++;
#else
        if (!copy_stream)
        {
            CRB_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
            CRB_CUDA_CHECK(cudaEventCreateWithFlags(&snap_done, cudaEventDisableTiming));
            for (cudaEvent_t &e : copy_done) CRB_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        if (staging.n != size_t(w) * h)
        {
            stream_sync(copy_stream);
            staging.alloc(size_t(w) * h);
        }
        // the one staging buffer is free again once the previous device->host copy has finished
        if (next_ticket) CRB_CUDA_CHECK(cudaStreamWaitEvent(stream(), copy_done[(next_ticket - 1) % 8], 0));
        dev_copy(staging.p, src, bytes, stream());
        CRB_CUDA_CHECK(cudaEventRecord(snap_done, stream()));
        CRB_CUDA_CHECK(cudaStreamWaitEvent(copy_stream, snap_done, 0));
        CRB_CUDA_CHECK(cudaMemcpyAsync(dst, staging.p, bytes, cudaMemcpyDeviceToHost, copy_stream));
        CRB_CUDA_CHECK(cudaEventRecord(copy_done[next_ticket % 8], copy_stream));
        return next_ticket++;
#endif
    }

    void Render::read_wait(uint64_t ticket)
    {
        if (ticket >= next_ticket) throw Error(ERR_INVALID_ARG, "read_wait: unknown ticket");
#ifndef CRB_EMU
        // copies complete in order on one stream: a ticket older than the event ring is covered by any newer one
        const uint64_t t = next_ticket - ticket > 8 ? next_ticket - 8 : ticket;
        CRB_CUDA_CHECK(cudaEventSynchronize(copy_done[t % 8]));
#endif
    }

    void Render::restore(const float *raw_sum_rgba, uint32_t passes_)
    {
        sync();
        dev_upload(accum.p, raw_sum_rgba, size_t(w) * h * 16, stream());
        stream_sync(stream());
        passes = passes_, pass_px = uint64_t(passes_) * w * h;    // the per-pixel counts travel in the image's alpha
        resolve();
        sync();
    }

    void Render::stats(crb_stats &out)
    {
        sync();
        unsigned long long st[ST_COUNT];
        dev_download(st, dstats.p, sizeof(st), stream());
        out.total_queries   = st[ST_CLOSEST] + st[ST_SHADOW];
        out.ref_rays        = st[ST_CLOSEST] + st[ST_RANOUT];    // renderer.cpp:271-272,356
        out.pixel_samples   = pixel_samples;
        out.passes          = passes;
        out.device_ms       = device_ms;
        out.kernel_launches = launches;
        out.node_visits[0] = st[ST_NODES], out.node_visits[1] = st[ST_NODES_SHADOW];
        out.tri_tests[0] = st[ST_TRIS], out.tri_tests[1] = st[ST_TRIS_SHADOW];
        out.closest_queries = st[ST_CLOSEST], out.shadow_queries = st[ST_SHADOW];
        for (int i = 0; i < 8; i++) out.kernel_ms[i] = kernel_ms[i], out.kernel_count[i] = kernel_count[i];
    }
}    // namespace crb

#ifdef CRB_EMU
// kernel-logic harness only: the per-ray statistics of RayClassProbe (16 counters), and their reset
extern "C" void crb_emu_ray_classes(unsigned long long *out16, int reset)
{
    for (int i = 0; i < 16; i++) out16[i] = (&crb::g_ray_classes[0][0][0])[i];
    if (reset)
        for (int i = 0; i < 16; i++) (&crb::g_ray_classes[0][0][0])[i] = 0;
}
#endif
