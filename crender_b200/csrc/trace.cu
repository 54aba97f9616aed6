// trace.cu — K5: batch closest-hit / any-hit queries (BASELINE config 3).
//
// Replaces cr::scene::cast_ray -> cr::model::intersect -> rtcIntersect1 called once per ray
// (src/render/scene.cpp:79-98, src/objects/model.cpp:5-49,99-126) with one kernel over a ray array.
// One thread per ray; rays are two coalesced 16-byte loads, the hit record is written as 24 bytes.
#include "scene.cuh"

#include <algorithm>

namespace crb
{
    namespace
    {
        __device__ __forceinline__ void resolve_flat(const DScene &sc, uint32_t flat, uint32_t &prim, uint32_t &model, uint32_t &inst)
        {
            // ranges are sorted by start; binary search the range containing `flat`
            uint32_t lo = 0, hi = sc.n_ranges;
            while (hi - lo > 1)
            {
                const uint32_t mid = (lo + hi) >> 1;
                if (sc.ranges[mid].start <= flat) lo = mid; else hi = mid;
            }
            const FlatRange r = sc.ranges[lo];
            prim = flat - r.start, model = r.model, inst = r.inst;
        }

        constexpr int BATCH_STEPS = 4;

        // persistent kernels: each warp is a pool of 32 traversal lanes refilled from `cursor` (bvh8.cuh)
        template<bool COUNT>
        __global__ void __launch_bounds__(128, 9) k_intersect_batch(DScene sc, const float4 *__restrict__ rays, uint32_t n, crb_hit *__restrict__ hits,
                                                                 uint32_t *cursor, unsigned long long *ctr)
        {
            TravCounters tc;
            auto source = [&](uint32_t idx, uint32_t &item, V3 &o, V3 &d, float &tmin, float &tmax) {
                item           = idx;
                const float4 a = rays[2 * size_t(idx)], b = rays[2 * size_t(idx) + 1];
                o = v3(a.x, a.y, a.z), d = v3(b.x, b.y, b.z), tmin = a.w, tmax = b.w;
            };
            auto sink = [&](bool valid, uint32_t item, const Hit &h) {
                if (COUNT || !valid) return;
                crb_hit out;
                out.t = h.t, out.u = h.u, out.v = h.v;
                out.prim = INVALID_PRIM, out.model = INVALID_PRIM, out.inst = 0;
                if (h.prim != INVALID_PRIM) resolve_flat(sc, h.prim, out.prim, out.model, out.inst);
                hits[item] = out;
            };
            trace_persistent<COUNT, BATCH_STEPS>(sc.bvh, cursor, n, 0u, false, source, sink, &tc);
            if (COUNT)
            {
                atomicAdd(ctr + 0, tc.nodes);
                atomicAdd(ctr + 1, tc.tris);
            }
        }

        template<bool COUNT>
        __global__ void __launch_bounds__(128, 9) k_occluded_batch(DScene sc, const float4 *__restrict__ rays, uint32_t n, uint8_t *__restrict__ occ,
                                                                uint32_t *cursor, unsigned long long *ctr)
        {
            TravCounters tc;
            auto source = [&](uint32_t idx, uint32_t &item, V3 &o, V3 &d, float &tmin, float &tmax) {
                item           = idx;
                const float4 a = rays[2 * size_t(idx)], b = rays[2 * size_t(idx) + 1];
                o = v3(a.x, a.y, a.z), d = v3(b.x, b.y, b.z), tmin = a.w, tmax = b.w;
            };
            auto sink = [&](bool valid, uint32_t item, const Hit &h) {
                if (occ && valid) occ[item] = h.prim != INVALID_PRIM ? 1 : 0;
            };
            trace_persistent<COUNT, BATCH_STEPS>(sc.bvh, cursor, n, 0u, true, source, sink, &tc);
            if (COUNT)
            {
                atomicAdd(ctr + 0, tc.nodes);
                atomicAdd(ctr + 1, tc.tris);
            }
        }

        // two-level scenes: the ray parameter t is invariant under the instance's affine map (no renormalisation for a
        // batch query with the caller's tnear/tfar), candidates are compared by t, ties go to the first (model, instance)
        template<bool COUNT, bool ANY>
        __global__ void __launch_bounds__(256, 4) k_batch2(DScene sc, const float4 *__restrict__ rays, uint32_t n, crb_hit *__restrict__ hits, uint8_t *__restrict__ occ,
                                                         uint32_t *cursor, unsigned long long *ctr)
        {
            TravCounters tc;
            auto source = [&](uint32_t idx, uint32_t &item, V3 &o, V3 &d, float &tmin, float &tmax) {
                item           = idx;
                const float4 a = rays[2 * size_t(idx)], b = rays[2 * size_t(idx) + 1];
                o = v3(a.x, a.y, a.z), d = v3(b.x, b.y, b.z), tmin = a.w, tmax = b.w;
            };
            auto sink = [&](bool valid, uint32_t item, const Hit &h) {
                if (COUNT || !valid) return;
                if (ANY)
                {
                    if (occ) occ[item] = h.prim != INVALID_PRIM ? 1 : 0;
                    return;
                }
                crb_hit out;
                out.t = h.t, out.u = h.u, out.v = h.v;
                out.prim = INVALID_PRIM, out.model = INVALID_PRIM, out.inst = 0;
                if (h.prim != INVALID_PRIM) resolve_flat(sc, h.prim, out.prim, out.model, out.inst);
                hits[item] = out;
            };
            trace_persistent_2l<COUNT, BATCH_STEPS, false>(sc.bvh2, cursor, n, ANY, source, sink, &tc);
            if (COUNT)
            {
                atomicAdd(ctr + 0, tc.nodes);
                atomicAdd(ctr + 1, tc.tris);
            }
        }

        struct Timer
        {
#ifndef CRB_EMU
            cudaEvent_t  e0 = nullptr, e1 = nullptr;
            cudaStream_t s;
            explicit Timer(cudaStream_t st) : s(st)
            {
                CRB_CUDA_CHECK(cudaEventCreate(&e0));
                CRB_CUDA_CHECK(cudaEventCreate(&e1));
            }
            ~Timer()
            {
                if (e0) cudaEventDestroy(e0);
                if (e1) cudaEventDestroy(e1);
            }
            void   start() { CRB_CUDA_CHECK(cudaEventRecord(e0, s)); }
            double stop()
            {
                CRB_CUDA_CHECK(cudaEventRecord(e1, s));
                CRB_CUDA_CHECK(cudaEventSynchronize(e1));
                float ms = 0;
                CRB_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
                return ms;
            }
#else
            explicit Timer(cudaStream_t) {}
            void   start() {}
            double stop() { return 0; }
#endif
        };

        constexpr uint64_t CHUNK = 1ull << 24;    // rays per kernel launch for device-pointer calls
        constexpr uint64_t PIPE_CHUNK = 1ull << 22;    // rays per pipeline stage for host-pointer calls (128 MB in, 96 MB out)

        void launch_batch(Scene &s, const DScene &sc, int mode, const float4 *rp, char *op, uint32_t cnt32, uint32_t *cursor, unsigned long long *ctr,
                          cudaStream_t st)
        {
#ifdef CRB_EMU
            const unsigned g = 1, blk = 1;
#else
            const unsigned g = unsigned(s.n_sms) * 9, blk = 128;    // 56 registers: 9 warps per scheduler (render.cu CRB_TRACE_OCC)
#endif
            dev_zero(cursor, 4, st);
            if (sc.two_level)
            {
#ifdef CRB_EMU
                const unsigned g2 = 1, blk2 = 1;
#else
                const unsigned g2 = unsigned(s.n_sms) * 4, blk2 = 256;
#endif
                switch (mode)
                {
                case 0: CRB_LAUNCH((k_batch2<false, false>), g2, blk2, st, sc, rp, cnt32, reinterpret_cast<crb_hit *>(op), (uint8_t *) nullptr, cursor, ctr); break;
                case 1: CRB_LAUNCH((k_batch2<false, true>), g2, blk2, st, sc, rp, cnt32, (crb_hit *) nullptr, reinterpret_cast<uint8_t *>(op), cursor, ctr); break;
                case 2: CRB_LAUNCH((k_batch2<true, false>), g2, blk2, st, sc, rp, cnt32, (crb_hit *) nullptr, (uint8_t *) nullptr, cursor, ctr); break;
                default: CRB_LAUNCH((k_batch2<true, true>), g2, blk2, st, sc, rp, cnt32, (crb_hit *) nullptr, (uint8_t *) nullptr, cursor, ctr); break;
                }
                return;
            }
            switch (mode)
            {
            case 0: CRB_LAUNCH((k_intersect_batch<false>), g, blk, st, sc, rp, cnt32, reinterpret_cast<crb_hit *>(op), cursor, ctr); break;
            case 1: CRB_LAUNCH((k_occluded_batch<false>), g, blk, st, sc, rp, cnt32, reinterpret_cast<uint8_t *>(op), cursor, ctr); break;
            case 2: CRB_LAUNCH((k_intersect_batch<true>), g, blk, st, sc, rp, cnt32, (crb_hit *) nullptr, cursor, ctr); break;
            default: CRB_LAUNCH((k_occluded_batch<true>), g, blk, st, sc, rp, cnt32, (uint8_t *) nullptr, cursor, ctr); break;
            }
        }

#ifndef CRB_EMU
        // Host-pointer queries: upload, traversal and download of consecutive chunks overlap on three streams with two
        // device slots (the reference calls cast_ray one ray at a time on the CPU; a host caller of the batch form has
        // its rays in host memory). With pinned host memory the copies are truly asynchronous and the call runs at
        // PCIe speed (32 B in + 24 B out per ray against ~0.05 ns of traversal); pageable memory still works, the
        // driver then stages the copies.
        struct BatchPipe
        {
            cudaStream_t up = nullptr, down = nullptr;
            cudaEvent_t  up_done[2] = {}, comp_done[2] = {}, down_done[2] = {}, t0 = nullptr, t1 = nullptr;
            BatchPipe()
            {
                CRB_CUDA_CHECK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
                CRB_CUDA_CHECK(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
                for (int i = 0; i < 2; i++)
                {
                    CRB_CUDA_CHECK(cudaEventCreateWithFlags(&up_done[i], cudaEventDisableTiming));
                    CRB_CUDA_CHECK(cudaEventCreateWithFlags(&comp_done[i], cudaEventDisableTiming));
                    CRB_CUDA_CHECK(cudaEventCreateWithFlags(&down_done[i], cudaEventDisableTiming));
                }
                CRB_CUDA_CHECK(cudaEventCreate(&t0));
                CRB_CUDA_CHECK(cudaEventCreate(&t1));
            }
            ~BatchPipe()
            {
                for (int i = 0; i < 2; i++)
                {
                    if (up_done[i]) cudaEventDestroy(up_done[i]);
                    if (comp_done[i]) cudaEventDestroy(comp_done[i]);
                    if (down_done[i]) cudaEventDestroy(down_done[i]);
                }
                if (t0) cudaEventDestroy(t0);
                if (t1) cudaEventDestroy(t1);
                if (up) cudaStreamDestroy(up);
                if (down) cudaStreamDestroy(down);
            }
        };

        void run_batch_pipelined(Scene &s, const DScene &sc, const crb_ray *rays, void *out, uint64_t n, int mode)
        {
            const size_t out_elem = mode == 0 ? sizeof(crb_hit) : 1;
            BatchPipe    pp;
            DBuf<float4> d_rays[2];
            DBuf<char>   d_out[2];
            DBuf<uint32_t> d_cursor;
            DBuf<unsigned long long> d_ctr;
            d_cursor.alloc(2), d_ctr.alloc(2);
            const uint64_t cap = std::min<uint64_t>(PIPE_CHUNK, n);
            for (int i = 0; i < 2; i++) d_rays[i].alloc(cap * 2), d_out[i].alloc(cap * out_elem);
            CRB_CUDA_CHECK(cudaEventRecord(pp.t0, s.stream));
            uint64_t ci = 0;
            for (uint64_t off = 0; off < n; off += PIPE_CHUNK, ci++)
            {
                const int      k   = int(ci & 1);
                const uint64_t cnt = std::min<uint64_t>(PIPE_CHUNK, n - off);
                if (ci >= 2) CRB_CUDA_CHECK(cudaStreamWaitEvent(pp.up, pp.comp_done[k], 0));      // the slot's rays have been traced
                CRB_CUDA_CHECK(cudaMemcpyAsync(d_rays[k].p, rays + off, cnt * sizeof(crb_ray), cudaMemcpyHostToDevice, pp.up));
                CRB_CUDA_CHECK(cudaEventRecord(pp.up_done[k], pp.up));
                CRB_CUDA_CHECK(cudaStreamWaitEvent(s.stream, pp.up_done[k], 0));
                if (ci >= 2) CRB_CUDA_CHECK(cudaStreamWaitEvent(s.stream, pp.down_done[k], 0));   // the slot's results have left
                launch_batch(s, sc, mode, d_rays[k].p, d_out[k].p, uint32_t(cnt), d_cursor.p + k, d_ctr.p, s.stream);
                CRB_CUDA_CHECK(cudaEventRecord(pp.comp_done[k], s.stream));
                CRB_CUDA_CHECK(cudaStreamWaitEvent(pp.down, pp.comp_done[k], 0));
                CRB_CUDA_CHECK(cudaMemcpyAsync(static_cast<char *>(out) + off * out_elem, d_out[k].p, cnt * out_elem, cudaMemcpyDeviceToHost, pp.down));
                CRB_CUDA_CHECK(cudaEventRecord(pp.down_done[k], pp.down));
            }
            CRB_CUDA_CHECK(cudaStreamSynchronize(pp.down));
            CRB_CUDA_CHECK(cudaEventRecord(pp.t1, s.stream));
            CRB_CUDA_CHECK(cudaEventSynchronize(pp.t1));
            float ms = 0;
            CRB_CUDA_CHECK(cudaEventElapsedTime(&ms, pp.t0, pp.t1));
            s.last_query_ms = ms;    // whole pipeline: copies included
        }
#endif

        // mode 0 closest, 1 any-hit, 2 counters(closest), 3 counters(any)
        void run_batch(Scene &s, const crb_ray *rays, void *out, uint64_t n, bool on_device, int mode, uint64_t *ctr_out)
        {
            s.require_committed();
            if (n && (!rays || (mode < 2 && !out))) throw Error(ERR_INVALID_ARG, "batch query: null ray/result pointer");
            const DScene sc = s.device_scene(1, 1);
#ifndef CRB_EMU
            if (!on_device && mode < 2 && n > 0)
            {
                run_batch_pipelined(s, sc, rays, out, n, mode);
                return;
            }
#endif
            Timer        timer(s.stream);
            double       ms = 0;
            DBuf<unsigned long long> d_ctr;
            d_ctr.alloc(2);
            dev_zero(d_ctr.p, 16, s.stream);
            DBuf<uint32_t> d_cursor;
            d_cursor.alloc(1);
            const size_t out_elem = mode == 0 ? sizeof(crb_hit) : 1;
            DBuf<float4> d_rays;
            DBuf<char>   d_out;
            for (uint64_t off = 0; off < n; off += CHUNK)
            {
                const uint64_t cnt = std::min<uint64_t>(CHUNK, n - off);
                const float4  *rp;
                char          *op;
                if (on_device)
                {
                    rp = reinterpret_cast<const float4 *>(rays + off);
                    op = static_cast<char *>(out) + off * out_elem;
                }
                else
                {
                    d_rays.ensure(cnt * 2);
                    if (mode < 2) d_out.ensure(cnt * out_elem);
                    dev_upload(d_rays.p, rays + off, cnt * sizeof(crb_ray), s.stream);
                    rp = d_rays.p, op = d_out.p;
                }
                timer.start();
                launch_batch(s, sc, mode, rp, op, uint32_t(cnt), d_cursor.p, d_ctr.p, s.stream);
                ms += timer.stop();
                if (!on_device && mode < 2) dev_download(static_cast<char *>(out) + off * out_elem, d_out.p, cnt * out_elem, s.stream);
            }
            if (ctr_out)
            {
                unsigned long long c[2];
                dev_download(c, d_ctr.p, 16, s.stream);
                ctr_out[0] = c[0], ctr_out[1] = c[1];
            }
            stream_sync(s.stream);
            s.last_query_ms = ms;
        }
    }    // namespace

    void intersect_batch(Scene &s, const crb_ray *rays, crb_hit *hits, uint64_t n, bool on_device) { run_batch(s, rays, hits, n, on_device, 0, nullptr); }
    void occluded_batch(Scene &s, const crb_ray *rays, uint8_t *occ, uint64_t n, bool on_device) { run_batch(s, rays, occ, n, on_device, 1, nullptr); }
    void trace_counters(Scene &s, const crb_ray *rays, uint64_t n, bool on_device, bool any_hit, uint64_t *nodes, uint64_t *tris)
    {
        uint64_t c[2] = { 0, 0 };
        run_batch(s, rays, nullptr, n, on_device, any_hit ? 3 : 2, c);
        if (nodes) *nodes = c[0];
        if (tris) *tris = c[1];
    }
}    // namespace crb

// ---------------------------------------------------------------------------------------------------
// Memory-system micro-benchmark used for the roofline context: read `bytes` of device
// memory `iters` times with 16-byte loads. A working set below the L2 size measures L2 bandwidth, a large
// one measures HBM read bandwidth.
namespace crb
{
    namespace
    {
        __global__ void __launch_bounds__(256) k_read_bw(const uint4 *__restrict__ p, size_t n16, int iters, unsigned *sink)
        {
            unsigned     acc    = 0;
            const size_t stride = size_t(gridDim.x) * blockDim.x;
            for (int it = 0; it < iters; it++)
                for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride)
                {
                    const uint4 v = p[i];
                    acc ^= v.x ^ v.y ^ v.z ^ v.w;
                }
            if (acc == 0x12345678u) *sink = acc;    // never true in practice; keeps the loads alive
        }
    }    // namespace

    double read_bandwidth_gbs(Scene &s, size_t bytes, int iters)
    {
#ifdef CRB_EMU
        (void) s, (void) bytes, (void) iters;
        return 0.0;
#else
        DBuf<uint4>    buf;
        DBuf<unsigned> sink;
        const size_t   n16 = bytes / 16;
        buf.alloc(n16), sink.alloc(1);
        dev_fill_byte(buf.p, 1, n16 * 16, s.stream);
        CRB_LAUNCH(k_read_bw, unsigned(s.n_sms) * 8, 256, s.stream, buf.p, n16, 1, sink.p);    // warm
        cudaEvent_t e0, e1;
        CRB_CUDA_CHECK(cudaEventCreate(&e0));
        CRB_CUDA_CHECK(cudaEventCreate(&e1));
        CRB_CUDA_CHECK(cudaEventRecord(e0, s.stream));
        CRB_LAUNCH(k_read_bw, unsigned(s.n_sms) * 8, 256, s.stream, buf.p, n16, iters, sink.p);
        CRB_CUDA_CHECK(cudaEventRecord(e1, s.stream));
        CRB_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        CRB_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0), cudaEventDestroy(e1);
        return double(n16) * 16.0 * iters / (double(ms) * 1e-3) / 1e9;
#endif
    }
}    // namespace crb
