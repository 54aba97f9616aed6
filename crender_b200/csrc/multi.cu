// multi.cu — multi-GPU rendering behind the crb_render handle: replicas, partition, merge (K10). See multi.cuh.
#include "multi.cuh"

#include <algorithm>
#ifndef CRB_EMU
#include <dlfcn.h>
#endif

namespace crb
{
    // ================================================================================= NCCL, loaded at run time
    namespace
    {
#ifndef CRB_EMU
        // the handful of NCCL entry points used, declared here so that the build does not depend on nccl.h
        // (values from NCCL 2.x's public header: ncclFloat32 = 7, ncclSum = 0, ncclUniqueId = 128 bytes)
        struct NcclId
        {
            char internal[128];
        };
        struct NcclApi
        {
            void *lib = nullptr;
            int (*GetUniqueId)(NcclId *)                                                                        = nullptr;
            int (*CommInitRank)(void **, int, NcclId, int)                                                      = nullptr;
            int (*CommInitAll)(void **, int, const int *)                                                       = nullptr;
            int (*CommDestroy)(void *)                                                                          = nullptr;
            int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t)                      = nullptr;
            int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t)                      = nullptr;
            int (*GroupStart)()                                                                                 = nullptr;
            int (*GroupEnd)()                                                                                   = nullptr;
            const char *(*GetErrorString)(int)                                                                  = nullptr;
            int (*GetVersion)(int *)                                                                            = nullptr;
            std::string why;
        };
        NcclApi &nccl()
        {
            static NcclApi            api;
            static std::once_flag     once;
            std::call_once(once, [] {
                const char *names[] = { getenv("CRB_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
                for (const char *n : names)
                {
                    if (!n || !*n) continue;
                    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
                    if (api.lib) break;
                    api.why = dlerror();
                }
                if (!api.lib) return;
                bool ok = true;
                auto sym = [&](const char *name) {
                    void *p = dlsym(api.lib, name);
                    if (!p) ok = false, api.why = std::string("missing symbol ") + name;
                    return p;
                };
                *(void **) &api.GetUniqueId    = sym("ncclGetUniqueId");
                *(void **) &api.CommInitRank   = sym("ncclCommInitRank");
                *(void **) &api.CommInitAll    = sym("ncclCommInitAll");
                *(void **) &api.CommDestroy    = sym("ncclCommDestroy");
                *(void **) &api.AllReduce      = sym("ncclAllReduce");
                *(void **) &api.Broadcast      = sym("ncclBroadcast");
                *(void **) &api.GroupStart     = sym("ncclGroupStart");
                *(void **) &api.GroupEnd       = sym("ncclGroupEnd");
                *(void **) &api.GetErrorString = sym("ncclGetErrorString");
                *(void **) &api.GetVersion     = sym("ncclGetVersion");
                if (!ok)
                {
                    dlclose(api.lib);
                    api.lib = nullptr;
                }
            });
            return api;
        }
        NcclApi &nccl_required()
        {
            NcclApi &a = nccl();
            if (!a.lib) throw Error(ERR_NCCL, "NCCL is not loadable (dlopen libnccl.so.2: " + a.why + ")");
            return a;
        }
        void nccl_check(int r, const char *what)
        {
            if (r != 0) throw Error(ERR_NCCL, std::string(what) + ": " + nccl().GetErrorString(r));
        }
        constexpr int NCCL_FLOAT = 7, NCCL_SUM = 0;

#endif

        __device__ __forceinline__ float4 flt_max4()
        {
            const float mx = 3.402823466e+38f;    // cr::image's fill (image.h:30-38)
            return make_float4(mx, mx, mx, mx);
        }

        // resolve behind an NCCL collective: merged sums -> display
        __global__ void __launch_bounds__(256) k_resolve_merged(const float4 *__restrict__ merged, float4 *__restrict__ display, uint32_t n)
        {
            for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
            {
                const float4 a = merged[i];
                display[i]     = a.w > 0.0f ? resolve_px(a, a.w) : flt_max4();
            }
        }

        // K10 fused: the collective and the resolve as ONE kernel over peer memory. This GPU owns pixels [lo, hi); it
        // pulls them from every rank's snapshot (NVLink loads, 16 bytes per lane, coalesced), sums in rank order
        // (spp partition) or takes the owner's value (tile partition: flipped row -> sample row -> row band ->
        // band_owner()), resolves, and stores sum + display into the root's merged buffers.
        __global__ void __launch_bounds__(256) k_merge_peers(const float4 *const *__restrict__ stage, int world, int tile, uint32_t w, uint32_t h, uint32_t lo,
                                                             uint32_t hi, float4 *__restrict__ merged_root, float4 *__restrict__ display_root)
        {
            for (uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x)
            {
                float4 a;
                if (tile)
                {
                    const uint32_t y = h - 1 - i / w;
                    a                = stage[band_owner(y / TILE_BAND_ROWS, uint32_t(world), TILE_SERPENTINE)][i];
                }
                else
                {
                    a = stage[0][i];
                    for (int r = 1; r < world; r++)
                    {
                        const float4 b = stage[r][i];
                        a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
                    }
                }
                merged_root[i]  = a;
                display_root[i] = a.w > 0.0f ? resolve_px(a, a.w) : flt_max4();
            }
        }
    }    // namespace

    bool nccl_available(std::string *why)
    {
#ifdef CRB_EMU
        if (why) *why = "kernel-logic harness: no NCCL";
        return false;
#else
        NcclApi &a = nccl();
        if (why) *why = a.why;
        return a.lib != nullptr;
#endif
    }

    void nccl_unique_id(void *out128)
    {
#ifdef CRB_EMU
        throw Error(ERR_NCCL, "no NCCL in the kernel-logic harness");
#else
        NcclId id;
        nccl_check(nccl_required().GetUniqueId(&id), "ncclGetUniqueId");
        memcpy(out128, &id, 128);
#endif
    }

    // ================================================================================= worker
    void Worker::start(int dev)
    {
        device = dev;
#ifndef CRB_EMU
        th = std::thread([this] {
            cudaSetDevice(device);
            for (;;)
            {
                std::function<void()> f;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return stop || !q.empty(); });
                    if (q.empty()) return;
                    f = std::move(q.front());
                    q.pop_front();
                    busy = true;
                }
                try
                {
                    f();
                }
                catch (...)
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (!err) err = std::current_exception();
                }
                {
                    std::lock_guard<std::mutex> lk(mu);
                    busy = false;
                }
                idle_cv.notify_all();
            }
        });
#endif
    }
    void Worker::post(std::function<void()> f)
    {
#ifdef CRB_EMU
        f();    // the harness executes launches serially anyway
#else
        {
            std::lock_guard<std::mutex> lk(mu);
            q.push_back(std::move(f));
        }
        cv.notify_one();
#endif
    }
    void Worker::wait()
    {
#ifndef CRB_EMU
        std::unique_lock<std::mutex> lk(mu);
        idle_cv.wait(lk, [&] { return q.empty() && !busy; });
        if (err)
        {
            std::exception_ptr e = err;
            err                  = nullptr;
            std::rethrow_exception(e);
        }
#endif
    }
    void Worker::join()
    {
#ifndef CRB_EMU
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        if (th.joinable()) th.join();
#endif
    }

    // ================================================================================= construction
    MultiRender::MultiRender(Scene *scene, const int *devices, int n, int partition_, uint32_t w_, uint32_t h_, uint32_t mb, uint32_t seed_, uint32_t flags_)
        : primary(scene), world(n), partition(partition_), w(w_), h(h_), max_bounces(mb), seed(seed_), flags(flags_)
    {
        if (n < 1) throw Error(ERR_INVALID_ARG, "create_multi: ngpus must be >= 1");
        if (partition != PARTITION_SPP && partition != PARTITION_TILE) throw Error(ERR_INVALID_ARG, "create_multi: partition must be 0 (spp) or 1 (tile)");
        scene->require_committed();
        all_local = true;
#ifndef CRB_EMU
        int count = 0;
        CRB_CUDA_CHECK(cudaGetDeviceCount(&count));
        for (int i = 0; i < n; i++)
        {
            const int d = devices ? devices[i] : i;
            if (d < 0 || d >= count) throw Error(ERR_INVALID_ARG, "create_multi: device " + std::to_string(d) + " does not exist (" + std::to_string(count) + " visible)");
            for (int j = 0; j < i; j++)
                if ((devices ? devices[j] : j) == d) throw Error(ERR_INVALID_ARG, "create_multi: device listed twice");
        }
#endif
        for (int i = 0; i < n; i++)
        {
            auto l    = std::make_unique<Local>();
            l->rank   = i;
            l->device = devices ? devices[i] : i;
#ifdef CRB_EMU
            l->device = scene->device;
#endif
            locals.push_back(std::move(l));
        }
        // replicas: scene + BVH on every GPU, built in parallel by the workers
        for (auto &lp : locals)
        {
            lp->worker = std::make_unique<Worker>();
            lp->worker->start(lp->device);
        }
        for_locals([&](Local &l) {
#ifndef CRB_EMU
            if (l.device == primary->device)
#else
            if (l.rank == 0)
#endif
                l.scene = primary;
            else
            {
                l.owned = std::make_unique<Scene>();
                l.scene = l.owned.get();
            }
            sync_replica(l, true);
            l.render = std::make_unique<Render>(l.scene, w, h, max_bounces, seed, flags);
            setup_partition(l);
        });
        init_common();
    }

    MultiRender::MultiRender(Scene *scene, const void *id128, int rank, int nranks, int partition_, uint32_t w_, uint32_t h_, uint32_t mb, uint32_t seed_,
                             uint32_t flags_)
        : primary(scene), world(nranks), partition(partition_), w(w_), h(h_), max_bounces(mb), seed(seed_), flags(flags_)
    {
        if (nranks < 1 || rank < 0 || rank >= nranks) throw Error(ERR_INVALID_ARG, "create_rank: need 0 <= rank < nranks");
        if (partition != PARTITION_SPP && partition != PARTITION_TILE) throw Error(ERR_INVALID_ARG, "create_rank: partition must be 0 (spp) or 1 (tile)");
        scene->require_committed();
        all_local = nranks == 1;
        auto l    = std::make_unique<Local>();
        l->rank = rank, l->device = scene->device, l->scene = scene;
        l->worker = std::make_unique<Worker>();
        l->worker->start(l->device);
        l->render = std::make_unique<Render>(scene, w, h, max_bounces, seed, flags);
        locals.push_back(std::move(l));
        setup_partition(*locals[0]);
#ifndef CRB_EMU
        if (nranks > 1)
        {
            if (!id128) throw Error(ERR_INVALID_ARG, "create_rank: null NCCL id");
            NcclApi &api = nccl_required();
            NcclId   id;
            memcpy(&id, id128, 128);
            DeviceScope ds(scene->device);
            nccl_check(api.CommInitRank(&locals[0]->comm, nranks, id, rank), "ncclCommInitRank");
            use_nccl = true;
        }
#else
        if (nranks > 1) throw Error(ERR_NCCL, "rank mode needs NCCL (not available in the kernel-logic harness)");
#endif
        init_common();
    }

    void MultiRender::init_common()
    {
#ifndef CRB_EMU
        if (all_local && world > 1)
        {
            // peer access between every pair -> the fused peer-memory merge; else NCCL
            static const int force_nccl = getenv("CRB_MULTI_NCCL") ? atoi(getenv("CRB_MULTI_NCCL")) : 0;
            bool             peers      = !force_nccl;
            for (auto &a : locals)
                for (auto &b : locals)
                {
                    if (a->device == b->device || !peers) continue;
                    int can = 0;
                    if (cudaDeviceCanAccessPeer(&can, a->device, b->device) != cudaSuccess || !can) peers = false;
                }
            if (peers)
            {
                for (auto &a : locals)
                {
                    DeviceScope ds(a->device);
                    for (auto &b : locals)
                    {
                        if (a->device == b->device) continue;
                        const cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
                        if (e == cudaErrorPeerAccessAlreadyEnabled)
                            cudaGetLastError();
                        else if (e != cudaSuccess)
                        {
                            cudaGetLastError();
                            peers = false;
                        }
                    }
                }
            }
            fused_peers = peers;
            if (!fused_peers)
            {
                NcclApi         &api = nccl_required();
                std::vector<int> devs;
                for (auto &l : locals) devs.push_back(l->device);
                std::vector<void *> comms(locals.size(), nullptr);
                nccl_check(api.CommInitAll(comms.data(), int(devs.size()), devs.data()), "ncclCommInitAll");
                for (size_t i = 0; i < locals.size(); i++) locals[i]->comm = comms[i];
                use_nccl = true;
            }
        }
        for (auto &lp : locals)
        {
            DeviceScope ds(lp->device);
            CRB_CUDA_CHECK(cudaStreamCreateWithFlags(&lp->comm_stream, cudaStreamNonBlocking));
            for (int k = 0; k < 2; k++)
            {
                CRB_CUDA_CHECK(cudaEventCreateWithFlags(&lp->snap_ready[k], cudaEventDisableTiming));
                CRB_CUDA_CHECK(cudaEventCreateWithFlags(&lp->merge_done[k], cudaEventDisableTiming));
            }
        }
#else
        fused_peers = all_local;
#endif
        for (auto &lp : locals)
        {
            DeviceScope ds(lp->device);
            alloc_merge_buffers(*lp);
        }
    }

    void MultiRender::alloc_merge_buffers(Local &l)
    {
        const size_t n = size_t(w) * h;
        l.stage[0].alloc(n), l.stage[1].alloc(n);
        // the peer path writes the merged image into the root's buffers only; NCCL needs a receive buffer per rank
        if (l.rank == locals[0]->rank || !fused_peers) l.merged.alloc(n), l.merged_display.alloc(n);
        l.peer_ptrs.alloc(size_t(2) * size_t(world));
        table_ok[0] = table_ok[1] = false;
    }

    void MultiRender::setup_partition(Local &l)
    {
        if (partition == PARTITION_TILE && world > 1)
            l.render->set_bands(TILE_BAND_ROWS, uint32_t(l.rank), uint32_t(world), TILE_SERPENTINE != 0);
        else
            l.render->set_rows(0, h);
    }

    MultiRender::~MultiRender()
    {
        try
        {
            sync();
        }
        catch (...)
        {
        }
        for (auto &lp : locals)
        {
            if (lp->worker) lp->worker->join();
            DeviceScope ds(lp->device);
#ifndef CRB_EMU
            if (lp->comm) nccl().CommDestroy(lp->comm);
            if (lp.get() == locals[0].get() && read_done) cudaEventDestroy(read_done);
            if (lp->comm_stream) cudaStreamSynchronize(lp->comm_stream), cudaStreamDestroy(lp->comm_stream);
            for (int k = 0; k < 2; k++)
            {
                if (lp->snap_ready[k]) cudaEventDestroy(lp->snap_ready[k]);
                if (lp->merge_done[k]) cudaEventDestroy(lp->merge_done[k]);
            }
#endif
            lp->render.reset();
            lp->stage[0].release(), lp->stage[1].release(), lp->merged.release(), lp->merged_display.release(), lp->aov_stage.release(), lp->peer_ptrs.release();
            lp->owned.reset();
        }
    }

    void MultiRender::for_locals(const std::function<void(Local &)> &f)
    {
        for (auto &lp : locals)
        {
            Local *l = lp.get();
            l->worker->post([l, &f] {
                DeviceScope ds(l->device);
                f(*l);
            });
        }
        std::exception_ptr first;
        for (auto &lp : locals)
        {
            try
            {
                lp->worker->wait();
            }
            catch (...)
            {
                if (!first) first = std::current_exception();
            }
        }
        if (first) std::rethrow_exception(first);
    }

    // brings a replica up to date with the caller's scene: geometry changes re-copy and re-commit (the BVH is rebuilt
    // on the replica's own GPU), everything else (camera, sun, materials, skybox rotation) is a light copy
    void MultiRender::sync_replica(Local &l, bool force)
    {
        if (l.scene == primary)
        {
            primary->require_committed();
            return;
        }
        if (!force && l.src_version == primary->version) return;
        if (force || l.src_geom != primary->geom_version)
        {
            l.owned->copy_description_from(*primary);
            l.owned->commit();
        }
        else
            l.owned->copy_light_state_from(*primary);
        l.src_version = primary->version, l.src_geom = primary->geom_version;
    }

    // ================================================================================= control
    void MultiRender::reset()
    {
        for_locals([&](Local &l) { l.render->reset(); });
        dirty = true, restored_px = 0;
    }

    void MultiRender::set_resolution(uint32_t w_, uint32_t h_)
    {
        sync();
        w = w_, h = h_;
        for_locals([&](Local &l) {
            l.render->set_resolution(w, h);
            setup_partition(l);
            alloc_merge_buffers(l);
        });
        dirty = true, restored_px = 0;
    }

    void MultiRender::set_max_bounces(uint32_t b)
    {
        max_bounces = b;
        for (auto &lp : locals) lp->render->max_bounces = b;
    }

    void MultiRender::refresh()
    {
        sync();
        for_locals([&](Local &l) {
            sync_replica(l, false);
            l.render->refresh();
        });
    }

    void MultiRender::render_samples(uint32_t first, uint32_t n)
    {
        if (n == 0) return;
        last_first = first, last_n = n;
        for (auto &lp : locals)
        {
            Local *l = lp.get();
            l->worker->post([this, l, first, n] {
                DeviceScope ds(l->device);
                if (partition == PARTITION_SPP)
                {
                    uint32_t lo, hi;
                    sample_share(uint32_t(l->rank), uint32_t(world), first, n, lo, hi);
                    if (hi > lo) l->render->render_samples(lo, hi - lo);
                }
                else
                    l->render->render_samples(first, n);    // its own row bands, every sample
            });
        }
        dirty = true;    // (workers are waited for at the next flush / sync: submission overlaps the caller)
    }

    // ================================================================================= merge
    void MultiRender::collective_accum(int k)
    {
        const size_t npx = size_t(w) * h;
#ifdef CRB_EMU
        // harness: every rank is local and memory is host memory — the fused peer kernel, executed serially
        Local &r = root();
        std::vector<const float4 *> ptrs;
        for (auto &lp : locals) ptrs.push_back(lp->stage[k].p);
        const float4 *const *pp = ptrs.data();
        CRB_LAUNCH(k_merge_peers, 1, 1, nullptr, pp, world, partition == PARTITION_TILE ? 1 : 0, w, h, 0u, uint32_t(npx), r.merged.p, r.merged_display.p);
#else
        if (fused_peers || world == 1)
        {
            Local &r = root();
            for (auto &lp : locals)
            {
                Local      &l = *lp;
                DeviceScope ds(l.device);
                // this GPU reads every rank's snapshot: wait for all of them
                for (auto &o : locals) CRB_CUDA_CHECK(cudaStreamWaitEvent(l.comm_stream, o->snap_ready[k], 0));
                // ... and writes into the root's merged buffers: a device->host copy of the previous merge must be over
                if (read_pending && &l != &r) CRB_CUDA_CHECK(cudaStreamWaitEvent(l.comm_stream, read_done, 0));
                const uint32_t lo = uint32_t(npx * size_t(l.rank) / size_t(world)), hi = uint32_t(npx * size_t(l.rank + 1) / size_t(world));
                const float4 *const *pp = l.peer_ptrs.p + size_t(k) * size_t(world);
                const unsigned g = std::min<unsigned>(unsigned(l.scene->n_sms) * 4u, (hi - lo + 255u) / 256u);
                if (hi > lo)
                    CRB_LAUNCH(k_merge_peers, g ? g : 1u, 256, l.comm_stream, pp, world, partition == PARTITION_TILE ? 1 : 0, w, h, lo, hi, r.merged.p, r.merged_display.p);
                CRB_CUDA_CHECK(cudaEventRecord(l.merge_done[k], l.comm_stream));
            }
            read_pending = false;
            return;
        }
        NcclApi &api = nccl_required();
        for (auto &lp : locals)
        {
            DeviceScope ds(lp->device);
            CRB_CUDA_CHECK(cudaStreamWaitEvent(lp->comm_stream, lp->snap_ready[k], 0));
        }
        if (partition == PARTITION_SPP)
        {
            // one all-reduce (sum) of the w*h float4 accumulators: 33 MB at 1080p, 133 MB at 4K
            nccl_check(api.GroupStart(), "ncclGroupStart");
            for (auto &lp : locals) nccl_check(api.AllReduce(lp->stage[k].p, lp->merged.p, npx * 4, NCCL_FLOAT, NCCL_SUM, lp->comm, lp->comm_stream), "ncclAllReduce");
            nccl_check(api.GroupEnd(), "ncclGroupEnd");
        }
        else
        {
            // all-gather of the interleaved row bands: band b is broadcast by its owner (band_owner). In the x/y-flipped
            // buffer the sample rows [y0,y1) are the contiguous rows [h-y1, h-y0).
            nccl_check(api.GroupStart(), "ncclGroupStart");
            uint32_t b = 0;
            for (uint32_t y0 = 0; y0 < h; y0 += TILE_BAND_ROWS, b++)
            {
                const uint32_t y1 = std::min(h, y0 + TILE_BAND_ROWS);
                const size_t   off = size_t(h - y1) * w, cnt = size_t(y1 - y0) * w;
                for (auto &lp : locals)
                    nccl_check(api.Broadcast(lp->stage[k].p + off, lp->merged.p + off, cnt * 4, NCCL_FLOAT, int(band_owner(b, uint32_t(world), TILE_SERPENTINE)), lp->comm, lp->comm_stream),
                               "ncclBroadcast");
            }
            nccl_check(api.GroupEnd(), "ncclGroupEnd");
        }
        for (auto &lp : locals)
        {
            DeviceScope ds(lp->device);
            CRB_LAUNCH(k_resolve_merged, unsigned(lp->scene->n_sms) * 4u, 256, lp->comm_stream, lp->merged.p, lp->merged_display.p, uint32_t(npx));
            CRB_CUDA_CHECK(cudaEventRecord(lp->merge_done[k], lp->comm_stream));
        }
#endif
    }

    void MultiRender::flush()
    {
        if (!dirty) return;
        for (auto &lp : locals) lp->worker->wait();    // the render launches are in their streams
        const int    k   = int(flushes & 1);
        const size_t npx = size_t(w) * h;
        for (auto &lp : locals)
        {
            Local      &l = *lp;
            DeviceScope ds(l.device);
            cudaStream_t rs = l.render->stream();
#ifndef CRB_EMU
            // stage[k] was read by the collective of flush k-2 (by every GPU on the peer path)
            if (flushes >= 2)
            {
                if (fused_peers)
                    for (auto &o : locals) CRB_CUDA_CHECK(cudaStreamWaitEvent(rs, o->merge_done[k], 0));
                else
                    CRB_CUDA_CHECK(cudaStreamWaitEvent(rs, l.merge_done[k], 0));
            }
#endif
            dev_copy(l.stage[k].p, l.render->accum.p, npx * 16, rs);
#ifndef CRB_EMU
            CRB_CUDA_CHECK(cudaEventRecord(l.snap_ready[k], rs));
#endif
            if (!table_ok[k] && int(locals.size()) == world)
            {
                // the peer-pointer table of this parity (static until the buffers are re-allocated)
                std::vector<const float4 *> ptrs;
                for (auto &o : locals) ptrs.push_back(o->stage[k].p);
                dev_upload(l.peer_ptrs.p + size_t(k) * size_t(world), ptrs.data(), ptrs.size() * sizeof(void *), rs);
                stream_sync(rs);
            }
        }
        table_ok[k] = true;
        collective_accum(k);
        flushes++;
        dirty = false;
    }

    // makes every render stream wait for the last merge: an event the caller records on the handle's stream afterwards
    // (crb_render_stream) then covers the collective and the resolve, so a whole step can be timed on the device
    void MultiRender::join_flush()
    {
        if (flushes == 0) return;
#ifndef CRB_EMU
        const int k = int((flushes - 1) & 1);
        for (auto &lp : locals)
        {
            DeviceScope ds(lp->device);
            if (fused_peers)
                for (auto &o : locals) CRB_CUDA_CHECK(cudaStreamWaitEvent(lp->render->stream(), o->merge_done[k], 0));
            else
                CRB_CUDA_CHECK(cudaStreamWaitEvent(lp->render->stream(), lp->merge_done[k], 0));
        }
#endif
    }

    void MultiRender::sync()
    {
        for (auto &lp : locals) lp->worker->wait();
        for (auto &lp : locals)
        {
            DeviceScope ds(lp->device);
            lp->render->sync();
#ifndef CRB_EMU
            if (lp->comm_stream) stream_sync(lp->comm_stream);
#endif
        }
    }

    const float4 *MultiRender::merged_buffer(int kind)
    {
        flush();
        Local &r = root();
        return kind == CRB_RAW_SUM ? r.merged.p : r.merged_display.p;
    }

    // AOVs (first hit of the latest sample, renderer.cpp:303-308,367-369) are not summed: with the spp partition the
    // rank that rendered the globally last sample holds them; with the tile partition every rank holds its own bands.
    void MultiRender::gather_aov(int kind)
    {
        const size_t npx = size_t(w) * h;
        Local       &r   = root();
        {
            DeviceScope ds(r.device);
            r.aov_stage.alloc(npx);
        }
        auto buf = [&](Local &l) -> float4 * { return kind == CRB_ALBEDO ? l.render->albedo.p : (kind == CRB_NORMAL ? l.render->normal.p : l.render->depth.p); };
        int  owner = 0;
        if (partition == PARTITION_SPP)
            for (int g = 0; g < world; g++)
            {
                uint32_t lo, hi;
                sample_share(uint32_t(g), uint32_t(world), last_first, last_n, lo, hi);
                if (hi > lo && hi == last_first + last_n) owner = g;
            }
        sync();
        if (all_local)
        {
            if (partition == PARTITION_SPP || world == 1)
            {
                Local &o = *locals[size_t(owner)];
#ifdef CRB_EMU
                memcpy(r.aov_stage.p, buf(o), npx * 16);
#else
                CRB_CUDA_CHECK(cudaMemcpyPeerAsync(r.aov_stage.p, r.device, buf(o), o.device, npx * 16, r.render->stream()));
#endif
            }
            else
            {
                uint32_t b = 0;
                for (uint32_t y0 = 0; y0 < h; y0 += TILE_BAND_ROWS, b++)
                {
                    const uint32_t y1 = std::min(h, y0 + TILE_BAND_ROWS);
                    const size_t   off = size_t(h - y1) * w, cnt = size_t(y1 - y0) * w;
                    Local         &o   = *locals[band_owner(b, uint32_t(world), TILE_SERPENTINE)];
#ifdef CRB_EMU
                    memcpy(r.aov_stage.p + off, buf(o) + off, cnt * 16);
#else
                    CRB_CUDA_CHECK(cudaMemcpyPeerAsync(r.aov_stage.p + off, r.device, buf(o) + off, o.device, cnt * 16, r.render->stream()));
#endif
                }
            }
            stream_sync(r.render->stream());
            return;
        }
#ifndef CRB_EMU
        // rank mode: a collective call (every rank reads the same AOV at the same point)
        NcclApi    &api = nccl_required();
        DeviceScope ds(r.device);
        nccl_check(api.GroupStart(), "ncclGroupStart");
        if (partition == PARTITION_SPP)
            nccl_check(api.Broadcast(buf(r), r.aov_stage.p, npx * 4, NCCL_FLOAT, owner, r.comm, r.comm_stream), "ncclBroadcast");
        else
        {
            uint32_t b = 0;
            for (uint32_t y0 = 0; y0 < h; y0 += TILE_BAND_ROWS, b++)
            {
                const uint32_t y1 = std::min(h, y0 + TILE_BAND_ROWS);
                const size_t   off = size_t(h - y1) * w, cnt = size_t(y1 - y0) * w;
                nccl_check(api.Broadcast(buf(r) + off, r.aov_stage.p + off, cnt * 4, NCCL_FLOAT, int(band_owner(b, uint32_t(world), TILE_SERPENTINE)), r.comm, r.comm_stream), "ncclBroadcast");
            }
        }
        nccl_check(api.GroupEnd(), "ncclGroupEnd");
        stream_sync(r.comm_stream);
#endif
    }

    void MultiRender::read(int kind, float *dst)
    {
        const size_t bytes = size_t(w) * h * 16;
        Local       &r     = root();
        if (kind == CRB_RAW_SUM || kind == CRB_PROGRESS)
        {
            const float4 *src = merged_buffer(kind);
            sync();
            DeviceScope ds(r.device);
            dev_download(dst, src, bytes, r.render->stream());
            return;
        }
        if (kind != CRB_ALBEDO && kind != CRB_NORMAL && kind != CRB_DEPTH) throw Error(ERR_INVALID_ARG, "read: unknown buffer kind");
        gather_aov(kind);
        DeviceScope ds(r.device);
        dev_download(dst, r.aov_stage.p, bytes, r.render->stream());
    }

    // the merged image after the work submitted so far, copied to the host behind the collective on the root's side
    // stream: the caller can submit the next crb_render_samples at once
    uint64_t MultiRender::read_async(int kind, float *dst)
    {
        if (kind != CRB_RAW_SUM && kind != CRB_PROGRESS)
        {
            read(kind, dst);
            return next_ticket++;
        }
        const float4 *src = merged_buffer(kind);
        Local        &r   = root();
        DeviceScope   ds(r.device);
#ifdef CRB_EMU
        memcpy(dst, src, size_t(w) * h * 16);
#else
        // all slices of the merged image have landed once every GPU's merge kernel is done
        const int k = int((flushes - 1) & 1);
        if (fused_peers)
            for (auto &o : locals) CRB_CUDA_CHECK(cudaStreamWaitEvent(r.comm_stream, o->merge_done[k], 0));
        CRB_CUDA_CHECK(cudaMemcpyAsync(dst, src, size_t(w) * h * 16, cudaMemcpyDeviceToHost, r.comm_stream));
        // the next flush overwrites the merged buffers: on the root its collective runs on this same stream, i.e. behind
        // the copy; the other GPUs' merge kernels write into the root's buffers too, so they wait for read_done
        if (!read_done) CRB_CUDA_CHECK(cudaEventCreateWithFlags(&read_done, cudaEventDisableTiming));
        CRB_CUDA_CHECK(cudaEventRecord(read_done, r.comm_stream));
        read_pending = true;
#endif
        return next_ticket++;
    }

    void MultiRender::read_wait(uint64_t ticket)
    {
        if (ticket >= next_ticket) throw Error(ERR_INVALID_ARG, "read_wait: unknown ticket");
#ifndef CRB_EMU
        Local      &r = root();
        DeviceScope ds(r.device);
        stream_sync(r.comm_stream);
#endif
    }

    void MultiRender::stats(crb_stats &out)
    {
        sync();
        memset(&out, 0, sizeof(out));
        uint64_t px = restored_px;
        for (auto &lp : locals)
        {
            DeviceScope ds(lp->device);
            crb_stats   s {};
            lp->render->stats(s);
            out.total_queries += s.total_queries, out.ref_rays += s.ref_rays, out.pixel_samples += s.pixel_samples;
            out.kernel_launches += s.kernel_launches, out.closest_queries += s.closest_queries, out.shadow_queries += s.shadow_queries;
            out.device_ms = std::max(out.device_ms, s.device_ms);
            for (int i = 0; i < 2; i++) out.node_visits[i] += s.node_visits[i], out.tri_tests[i] += s.tri_tests[i];
            for (int i = 0; i < 8; i++) out.kernel_ms[i] = std::max(out.kernel_ms[i], s.kernel_ms[i]), out.kernel_count[i] += s.kernel_count[i];
            px += lp->render->pass_px;
        }
        out.passes = px / (uint64_t(w) * h);    // whole-frame passes over the LOCAL ranks (all of them in single-process mode)
    }

    void MultiRender::restore(const float *raw, uint32_t passes)
    {
        sync();
        // spp partition: the checkpoint goes to the first rank, the others restart from zero (the merge is a sum);
        // tile partition: every rank takes it, the merge only reads a band from its owner
        for_locals([&](Local &l) {
            if (partition == PARTITION_TILE || l.rank == 0)
                l.render->restore(raw, passes);
            else
                l.render->reset();
            l.render->passes = 0, l.render->pass_px = 0;    // the restored passes are counted once, below
        });
        restored_px = uint64_t(passes) * w * h;
        dirty       = true;
    }
}    // namespace crb
