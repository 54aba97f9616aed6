// multi.cuh — multi-GPU rendering behind the same crb_render handle (K10).
//
// The reference has one process, one CPU, no collective of any kind ("NCCL / MPI / any
// collective call site: NONE"); its only parallelism is one task per scanline inside a pass
// (src/render/renderer.cpp:240-256). The path shards without a data-path exchange (every pixel-sample touches only
// its own pixel, renderer.cpp:362-383, and the scene is read-only during a pass), so here:
//
//   scene + BVH     replicated: one crb::Scene per GPU, committed on that GPU (the device build is deterministic)
//   work            partitioned by SAMPLE INDEX (PARTITION_SPP: rank g renders a contiguous share of every
//                   crb_render_samples range, BASELINE config 4) or by interleaved 8-ROW BANDS (PARTITION_TILE,
//                   BASELINE config 5); the sampler is keyed by global pixel and sample, so the union over ranks is
//                   the single-GPU set of paths
//   merge ("flush") snapshot of every rank's float4 accumulator (device-to-device on the render stream), then on a
//                   SIDE stream either ncclAllReduce(sum) (spp) or an all-gather of the row bands as grouped
//                   ncclBroadcasts (tile), with the resolve pow(clamp(sum/n,0,1),1/2.2) (renderer.cpp:371-383)
//                   fused behind the collective. Two staging buffers alternate, so the collective of flush k
//                   overlaps the kernels of step k+1. When every rank lives in this process and the GPUs have peer
//                   access, the collective AND the resolve are ONE kernel per GPU that pulls its slice of every
//                   peer's snapshot over NVLink, sums in rank order (bit-reproducible) and writes sum + display
//                   into the root GPU's merged buffers (k_merge_peers).
//
// Two ways to get ranks: crb_render_create_multi (one process drives all GPUs, ncclCommInitAll) and
// crb_render_create_rank (one process per GPU — torchrun/MPI — ncclCommInitRank with an id the host broadcasts).
// NCCL is loaded with dlopen("libnccl.so.2") at first use: a process that already carries a copy (PyTorch bundles
// one) shares it, a C host gets the system library; without NCCL the single-process peer path still works.
#pragma once
#include "render.cuh"

#include <condition_variable>
#include <deque>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace crb
{
    enum { PARTITION_SPP = 0, PARTITION_TILE = 1 };
    // Height of the interleaved row bands of the tile partition, and the owner order. Measured on one GPU by rendering the
    // 8 ranks' shares of the 4K config-5 frame separately (tools/band_balance.py, profiles/r2_sweeps.md section 9), mean/max
    // of the ranks' times = the strong-scaling efficiency the partition can reach: 64-row bands (SURVEY.md's example: 34 bands,
    // 5 or 4 per GPU) 0.76, 16 rows 0.95, 16 rows serpentine 0.97, 8 rows serpentine 0.99. On the 8-GPU box config 5 scaled at
    // 0.82 with 64-row and 0.955 with 16-row plain bands.
    constexpr uint32_t TILE_BAND_ROWS = 8;
    constexpr uint32_t TILE_SERPENTINE = 1;    // owner order reversed in every other period of bands (render.cuh band_owner)

    // contiguous share [lo, hi) of `n` samples starting at `first` for `rank` of `world` (earlier ranks take the remainder)
    inline void sample_share(uint32_t rank, uint32_t world, uint32_t first, uint32_t n, uint32_t &lo, uint32_t &hi)
    {
        const uint32_t base = n / world, rem = n % world;
        lo = first + rank * base + (rank < rem ? rank : rem);
        hi = lo + base + (rank < rem ? 1u : 0u);
    }

    // one submission thread per local GPU: a render call is hundreds of kernel launches, and with the work of a step
    // split N ways the device time per call shrinks N-fold while a single host thread's launch time would grow N-fold
    struct Worker
    {
        std::thread                       th;
        std::mutex                        mu;
        std::condition_variable           cv, idle_cv;
        std::deque<std::function<void()>> q;
        bool                              stop = false, busy = false;
        std::exception_ptr                err;
        int                               device = 0;
        void start(int dev);
        void post(std::function<void()> f);
        void wait();    // rethrows the first exception of a posted task
        void join();
    };

    struct MultiRender
    {
        struct Local
        {
            int                     rank = 0, device = 0;
            Scene                  *scene = nullptr;
            std::unique_ptr<Scene>  owned;     // replica created by the library (null: the caller's own scene)
            uint64_t                src_version = ~0ull, src_geom = ~0ull;
            std::unique_ptr<Render> render;
            void                   *comm = nullptr;    // ncclComm_t
            cudaStream_t            comm_stream = nullptr;
#ifndef CRB_EMU
            cudaEvent_t snap_ready[2] = {}, merge_done[2] = {};
#endif
            DBuf<float4>            stage[2];    // snapshots of the local accumulator
            DBuf<float4>            merged;      // accumulators merged over all ranks (w = per-pixel pass count)
            DBuf<float4>            merged_display;
            DBuf<float4>            aov_stage;   // lazily: gathered AOV buffer
            DBuf<const float4 *>    peer_ptrs;   // fused peer path: stage[k] of every rank, [2][world]
            std::unique_ptr<Worker> worker;
        };
        Scene   *primary;
        int      world = 1, partition = PARTITION_SPP;
        uint32_t w, h, max_bounces, seed, flags;
        bool     all_local = true;     // every rank is a Local of this process
        bool     fused_peers = false;  // collective + resolve as one peer-memory kernel
        bool     use_nccl = false;
        std::vector<std::unique_ptr<Local>> locals;
        uint64_t flushes = 0;
        bool     dirty = true;         // samples rendered since the last flush
        uint32_t last_first = 0, last_n = 0;
        uint64_t next_ticket = 0;
        uint64_t restored_px = 0;
        bool     table_ok[2] = { false, false };
        bool     read_pending = false;
#ifndef CRB_EMU
        cudaEvent_t read_done = nullptr;    // root GPU: the last asynchronous device->host copy of the merged image
#endif

        // single process: ranks 0..n-1 on `devices`; rank mode: one local rank `rank` of `nranks` on the scene's device
        MultiRender(Scene *scene, const int *devices, int n, int partition, uint32_t w, uint32_t h, uint32_t mb, uint32_t seed, uint32_t flags);
        MultiRender(Scene *scene, const void *nccl_id128, int rank, int nranks, int partition, uint32_t w, uint32_t h, uint32_t mb, uint32_t seed,
                    uint32_t flags);
        ~MultiRender();

        Local &root() { return *locals[0]; }
        void   reset();
        void   set_resolution(uint32_t w, uint32_t h);
        void   set_max_bounces(uint32_t b);
        void   refresh();
        void   render_samples(uint32_t first, uint32_t n);
        void   flush();    // asynchronous merge; implied by the read calls
        void   join_flush();
        void   sync();
        void   read(int kind, float *dst);
        uint64_t read_async(int kind, float *dst);
        void     read_wait(uint64_t ticket);
        void   stats(crb_stats &out);
        void   restore(const float *raw_sum_rgba, uint32_t passes);
        void   resolve() { flush(); }
        const float4 *merged_buffer(int kind);    // device pointer on the root local, after a flush

    private:
        void init_common();
        void setup_partition(Local &l);
        void alloc_merge_buffers(Local &l);
        void for_locals(const std::function<void(Local &)> &f);    // on the workers, waits for all
        void sync_replica(Local &l, bool force);
        void gather_aov(int kind);
        void collective_accum(int k);
    };

    // nccl unique id for crb_render_create_rank (128 bytes)
    void nccl_unique_id(void *out128);
    bool nccl_available(std::string *why);
}    // namespace crb
