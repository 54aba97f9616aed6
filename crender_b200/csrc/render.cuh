// render.cuh — the wavefront path tracer that replaces cr::renderer's management thread, per-scanline
// thread-pool tasks and _sample_pixel (src/render/renderer.cpp:116-144,240-384).
#pragma once
#include "scene.cuh"

namespace crb
{
    struct ShadowRay    // 48 bytes
    {
        float4 o;    // xyz origin, w = path slot (bits)
        float4 d;    // xyz direction as sampled (un-normalised), w = tmax (inf for the sun)
        float4 c;    // rgb contribution added to the path if the light is visible
    };

    // device pointers of the path state. A path's identity is its slot = sample_in_batch * npix + pixel (radiance record,
    // pixel, sample index); its per-bounce RECORD (ray, throughput, hit) lives at its position in the bounce's queue, so
    // that every kernel streams the records of a bounce in order: k_shade writes the surviving paths' next records
    // compacted into the *_next arrays, and the two sets are swapped after every bounce.
    struct PathState
    {
        float4 *ray_o;    // [queue position] xyz origin
        float4 *ray_d;    // [queue position] xyz direction exactly as the reference's cr::ray::direction (not re-normalised)
        float4 *thr;      // [queue position] xyz throughput, w = extended mode: previous vertex was specular
        float4 *hit;      // [queue position] t (normalised-direction units), u, v, flat prim (bits)
        float4 *ray_o_next, *ray_d_next, *thr_next;    // the next bounce's records
        float4 *rad;      // [slot] xyz radiance ("final")
        uint32_t *q_in;         // [queue position] path slot of this bounce's records
        uint32_t *q_next;       // path slots of the next bounce's records
        uint32_t *q_class[4];   // after trace: record indices by class: 0 miss, 1 metal, 2 smooth, 3 glass (material sort)
        ShadowRay *shadow;
        // instance wavefront (two-level scenes, render.cu k_iw_*): per ray [queue position] the candidate instances sorted by
        // entry distance (8 each), the best hit so far (t, u, v, flat prim) and (its world distance, its instance, the next
        // candidate); two item queues (object-space origin + ray, direction + pruning bound, instance), the items' hits, the
        // fallback queue
        float    *iw_cand_t;
        uint32_t *iw_cand_i;
        float4   *iw_best, *iw_meta;
        float4   *iw_item_o[2], *iw_item_d[2];
        uint32_t *iw_item_k[2];
        float4   *iw_item_hit;
        uint32_t *iw_fb;
        uint32_t  *counters;    // see CTR_* below
        unsigned long long *stats;    // see ST_* below
        uint32_t  trace_chunk;        // work-reservation granularity of the persistent trace loop (0 = exact)
        int       sorted;             // 1: k_classify sorts paths by shade class before k_shade; 0: shade in queue order
    };
    enum { CTR_IN = 0, CTR_CLASS0 = 1, CTR_NEXT = 5, CTR_SHADOW = 6, CTR_CUR_TRACE = 7, CTR_CUR_SHADE = 8, CTR_CUR_SHADOW = 9,
           CTR_IW_ITEMS = 10 /* and 11: the two item queues */, CTR_IW_CUR = 12, CTR_IW_FB = 13, CTR_IW_FB_CUR = 14, CTR_COUNT = 16 };
    enum { ST_CLOSEST = 0, ST_SHADOW = 1, ST_RANOUT = 2, ST_NODES = 3, ST_TRIS = 4, ST_NODES_SHADOW = 5, ST_TRIS_SHADOW = 6, ST_COUNT = 8 };

    // display = pow(clamp(sum / n, 0, 1), 1/2.2), alpha 1 (renderer.cpp:371-383); shared by the single-GPU accumulate
    // and the multi-GPU merge kernels so that both resolve with the same operations
    __device__ __forceinline__ float4 resolve_px(float4 a, float n)
    {
        const float g = 1.f / 2.2f;
        return make_float4(powf(clampf(a.x / n, 0.0f, 1.0f), g), powf(clampf(a.y / n, 0.0f, 1.0f), g), powf(clampf(a.z / n, 0.0f, 1.0f), g), 1.0f);
    }

    // Tile partition, owner of the bands: period p = bands [p*stride, (p+1)*stride). Plain interleave gives slot j of every
    // period to rank j; the serpentine order reverses every other period, which cancels a cost gradient along the image
    // rows (measured on the 4K config-5 frame, 8 ranks, 16-row bands: max/mean of the ranks' times 1.049 plain — rank 0
    // always the slowest, rank 7 the fastest — see profiles/r2_sweeps.md section 9).
    __host__ __device__ inline uint32_t band_slot(uint32_t period, uint32_t first, uint32_t stride, uint32_t serpentine)
    {
        return (serpentine && (period & 1u)) ? stride - 1u - first : first;
    }
    __host__ __device__ inline uint32_t band_owner(uint32_t b, uint32_t world, uint32_t serpentine)
    {
        const uint32_t p = b / world, j = b % world;
        return (serpentine && (p & 1u)) ? world - 1u - j : j;
    }

    struct RenderParams
    {
        uint32_t w, h, row0, nrows, npix;    // npix = w * nrows (pixels rendered per pass)
        uint32_t seed;
        uint32_t first_sample, batch;        // this batch renders global samples first_sample .. +batch-1
        uint32_t aov_sample;                 // global sample index whose first hit is written to the AOVs
        uint32_t bounce;
        uint32_t band, band_first, band_stride;    // band != 0: local rows are interleaved bands (tile partition), see row_of()
        uint32_t band_serp;                        // 1: the owner order is reversed in every other period (band_slot())
        float4  *accum, *display, *albedo, *normal, *depth;
        const float *table;    // caller-supplied sample table [table_samples][w*h][table_dims] or nullptr
        uint32_t     table_samples, table_dims;
    };

    struct Render
    {
        Scene   *scene;
        uint32_t w, h, max_bounces, seed, flags;
        uint32_t row0, row1;
        uint32_t band = 0, band_first = 0, band_stride = 1, band_nrows = 0;    // set_bands(): interleaved row bands instead of [row0,row1)
        uint32_t band_serp = 0;
        uint32_t passes = 0;     // whole-frame passes completed (_current_sample): pass_px / (w*h)
        uint64_t pass_px = 0;    // pixel-samples accumulated (incl. a restored checkpoint's)
        uint64_t scene_version = ~0ull;
        DScene   dscene;

        DBuf<float4> accum, display, albedo, normal, depth;
        // path state
        size_t              capacity = 0;
        DBuf<float4>        ray_o, ray_d, thr, rad, hit, ray_o2, ray_d2, thr2;
        size_t              iw_capacity = 0;
        DBuf<float>         iw_cand_t;
        DBuf<uint32_t>      iw_cand_i, iw_item_k[2], iw_fb;
        DBuf<float4>        iw_best, iw_meta, iw_item_o[2], iw_item_d[2], iw_item_hit;
        DBuf<uint32_t>      q_in, q_next, q_class[4];
        DBuf<ShadowRay>     shadow;
        DBuf<uint32_t>      counters;
        DBuf<unsigned long long> dstats;
        int      n_sms = 1;
        double   device_ms = 0;
        uint64_t launches  = 0;
        uint64_t pixel_samples = 0;
        // paths in flight per batch: bigger batches amortise kernel tails and keep the late, thin bounces
        // wide enough for 148 SMs (swept in profiles/r1c_sweeps.md §10); 200 B of path state each
        size_t   target_paths = size_t(1) << 25;
        size_t   mem_path_cap = 0;
        double   kernel_ms[8]    = {};
        uint64_t kernel_count[8] = {};
#ifndef CRB_EMU
        // one event pair per render_samples call; resolved without blocking when the next call is queued
        // (a blocking wait here would serialise the host's launch work with the device), blocking at sync()
        struct Span
        {
            cudaEvent_t a, b;
        };
        std::vector<Span> spans;
        // asynchronous read-back (read_async / read_wait): snapshot on the render stream, device->host copy on
        // its own stream so that it overlaps the next render_samples call
        cudaStream_t copy_stream = nullptr;
        cudaEvent_t  snap_done   = nullptr;
        cudaEvent_t  copy_done[8] = {};
#endif
        DBuf<float4> staging;
        uint64_t     next_ticket = 0;
        DBuf<float>  sample_table;
        uint32_t     table_samples = 0, table_dims = 0;
#ifndef CRB_EMU
        // CRB_RENDER_FLAG_TIMERS: event pairs around every launch, resolved at sync()
        struct Timed
        {
            int         cls;
            cudaEvent_t a, b;
        };
        std::vector<Timed>       timed;
        std::vector<cudaEvent_t> ev_pool;
        cudaEvent_t              take_event();
#endif
        void tick(int cls);    // call before a launch
        void tock();           // call after it

        Render(Scene *s, uint32_t w, uint32_t h, uint32_t max_bounces, uint32_t seed, uint32_t flags);
        ~Render();
        cudaStream_t stream() const { return scene->stream; }
        void reset();
        void set_resolution(uint32_t w, uint32_t h);
        void set_rows(uint32_t y0, uint32_t y1);
        void set_bands(uint32_t band_rows, uint32_t first, uint32_t stride, bool serpentine = false);
        void refresh();
        void render_samples(uint32_t first, uint32_t n);
        void sync();
        void resolve();
        void set_pass_count(uint32_t passes_);
        void set_sample_table(const float *table_host, uint32_t n_samples, uint32_t dims);
        void read(int kind, float *dst);
        uint64_t read_async(int kind, float *dst);
        void     read_wait(uint64_t ticket);
        void restore(const float *raw_sum_rgba, uint32_t passes_);
        void stats(crb_stats &out);

    private:
        void alloc_images();
        void ensure_paths(size_t n);
        void ensure_iw(size_t n);
        void collect_time(bool wait = true);
        const float4 *buffer_of(int kind) const;
    };
}    // namespace crb
