/*
 * crender_b200.h — C ABI of the B200-native path-tracing core for CRender.
 *
 * The reference has no plugin/FFI interface: its hot path sits behind the in-process C++ classes
 * cr::scene and cr::renderer that the ImGui layer calls. This header re-exports exactly that class
 * surface as plain C (opaque handles, plain pointers and sizes, no C++/torch types) so that a
 * maintainer can put it behind cr::scene / cr::renderer (see INTEGRATION.md for the shim), and so that
 * tests/bench drive it through ctypes. Each entry point cites the reference interface it replaces
 * (paths relative to the CRender source tree).
 *
 * Conventions
 *  - every function returns 0 on success or a crb_status code; crb_last_error() gives the message of
 *    the last failure on the calling thread. Nothing here ever calls exit() (the reference's
 *    cr::exit(), src/util/exception.h:9-13, terminates the process; codes 30/31 are kept from
 *    data/errors.json for geometry-buffer failures).
 *  - the library owns all device memory; host arrays passed in are copied before the call returns.
 *  - one host control thread per handle; work is asynchronous on the handle's CUDA stream until
 *    crb_render_sync / crb_render_read (same contract as renderer::update = pause -> mutate -> start,
 *    src/render/renderer.cpp:185-192).
 *  - there is NO CPU fallback: without a CUDA device every call fails with CRB_ERR_NO_DEVICE.
 */
#ifndef CRENDER_B200_H
#define CRENDER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum crb_status
{
    CRB_OK                = 0,
    CRB_ERR_GENERIC       = 1,  /* data/errors.json "1" */
    CRB_ERR_INVALID_ARG   = 2,
    CRB_ERR_NO_DEVICE     = 10,
    CRB_ERR_CUDA          = 11,
    CRB_ERR_OOM           = 12,
    CRB_ERR_NCCL          = 13, /* NCCL not loadable / a collective failed (multi-GPU handles only) */
    CRB_ERR_BUILD_VERTS   = 30, /* data/errors.json "30": could not create vertex buffer */
    CRB_ERR_BUILD_INDEX   = 31, /* data/errors.json "31": could not create index buffer */
    CRB_ERR_BVH_DEPTH     = 32,
    CRB_ERR_NOT_COMMITTED = 33
} crb_status;

/* cr::material::type, src/render/material/material.h:14-19 (same numeric order) */
enum { CRB_METAL = 0, CRB_SMOOTH = 1, CRB_GLASS = 2 };

/* cr::material::information, src/render/material/material.h:31-41 */
typedef struct crb_material
{
    uint32_t shade_type;     /* default CRB_SMOOTH */
    float    ior;            /* 1.5 */
    float    roughness;      /* 0.5 (dead in the reference: renderer.cpp:84-86 is commented out; GGX alpha in extended mode) */
    float    reflectiveness; /* 1 */
    float    emission;       /* 0 */
    float    colour[4];      /* 1,1,1,1; colour[3]==0 is an alpha cut-out (renderer.cpp:37-41) */
    int32_t  tex;            /* -1 none, else id from crb_scene_add_texture */
} crb_material;

/* cr::entity::sun, src/render/entities/components.h:23-29 */
typedef struct crb_sun
{
    float size;         /* cone half-angle, default pi/48 */
    float intensity;    /* 100 */
    float direction[3]; /* normalize(0.8,-1,0) */
    float colour[3];    /* 1,.9,.7 */
} crb_sun;

/* cr::camera, src/render/camera.h:11-41 */
typedef struct crb_camera
{
    float    position[3]; /* default 5,5,0 */
    float    rotation[3]; /* degrees: x about UP, y about RIGHT, z about FORWARD (camera.cpp:54-68) */
    float    fov;         /* degrees, default 75 */
    float    scale;       /* orthographic half-extent */
    uint32_t mode;        /* 0 perspective, 1 orthographic */
} crb_camera;

/* RTCRay subset, src/objects/model.cpp:14-25 */
typedef struct crb_ray
{
    float o[3], tmin;
    float d[3], tmax;
} crb_ray;

/* RTCHit subset + which model/instance (cr::ray::intersection_record, src/render/ray.h:15-23) */
typedef struct crb_hit
{
    float    t; /* +inf on miss, in units of |d|: the batch query does NOT normalise d (rtcIntersect1's own convention);
                   scene::cast_ray's callers get world-space distances because model.cpp:110 normalises first — pass unit d */
    float    u, v;
    uint32_t prim;  /* triangle index inside the model, 0xffffffff on miss */
    uint32_t model; /* model id, 0xffffffff on miss */
    uint32_t inst;  /* instance index inside the model */
} crb_hit;

typedef struct crb_build_info
{
    double   build_ms;  /* device time of the BVH build (the rtcCommitScene replacement) */
    double   upload_ms; /* host->device copies + flattening */
    uint64_t n_triangles; /* triangles a ray can hit: instances expanded (resident only once per model in two-level scenes) */
    uint64_t n_nodes;     /* 8-wide nodes */
    uint64_t node_bytes;
    uint64_t tri_bytes;
    uint32_t max_depth;
    float    sah_cost;
} crb_build_info;

/* cr::renderer::renderer_stats, src/render/renderer.h:48-54, extended */
typedef struct crb_stats
{
    uint64_t total_queries; /* closest-hit segments + shadow queries actually traced */
    uint64_t ref_rays;      /* the reference's _total_rays rule (renderer.cpp:271-272,356) */
    uint64_t pixel_samples;
    uint64_t passes;        /* _current_sample */
    double   device_ms;     /* device time spent in render kernels since reset */
    uint64_t kernel_launches;
    /* only with CRB_RENDER_FLAG_COUNTERS: traversal work, closest-hit ([0]) and shadow ([1]) kernels */
    uint64_t node_visits[2];
    uint64_t tri_tests[2];
    uint64_t closest_queries; /* total_queries = closest_queries + shadow_queries */
    uint64_t shadow_queries;
    /* only with CRB_RENDER_FLAG_TIMERS: device ms and launch count per kernel class, CUDA events around
     * every launch on the handle's stream. index: CRB_K_* */
    double   kernel_ms[8];
    uint64_t kernel_count[8];
    /* cr::renderer::renderer_stats as the reference defines it (renderer.h:48-54, renderer.cpp:396-404), with the device
     * time spent in render kernels since the last reset in place of the wall clock: running_time, rays_per_second =
     * ref_rays / running_time (path segments; shadow rays are not counted by the reference), samples_per_second =
     * passes / running_time (whole-frame PASSES per second, which is what the reference calls samples) */
    double   running_time;
    double   rays_per_second;
    double   samples_per_second;
} crb_stats;
enum { CRB_K_RAYGEN = 0, CRB_K_TRACE = 1, CRB_K_SHADE = 2, CRB_K_SHADOW = 3, CRB_K_ADVANCE = 4, CRB_K_ACCUMULATE = 5 };

/* cr::post_processor::{bloom,gray_scale,tonemapping}_settings, src/render/post/post_processor.h:19-39 */
typedef struct crb_post_settings
{
    int32_t use_bloom;            /* false */
    float   bloom_threshold;      /* 0.7 */
    float   bloom_strength;       /* 1.0 */
    int32_t use_gray_scale;       /* false */
    int32_t use_tonemapping;      /* false */
    int32_t tonemapping_type;     /* 0 linear, 1 reinhard, 2 jim-richard, 3 uncharted */
    float   tonemapping_exposure; /* 1.0 */
    float   gamma_correction;     /* 2.2 */
} crb_post_settings;

typedef struct crb_scene  crb_scene;
typedef struct crb_render crb_render;

const char *crb_last_error(void);
/* selects the CUDA device for handles created afterwards by this thread; -1 = current */
int crb_set_device(int device);
int crb_device_info(char *name, int name_cap, int *sm_count, uint64_t *l2_bytes, uint64_t *hbm_bytes);

/* ---- scene: cr::scene (src/render/scene.h:19-52) + cr::registry (src/render/entities/registry.h) */
int crb_scene_create(crb_scene **out);
int crb_scene_destroy(crb_scene *);
/* scene::add_model -> registry::register_model (scene.cpp:31-34, registry.cpp:51-97): the de-indexed
 * triangle soup (9 floats/tri), per-corner uvs (6 floats/tri, may be NULL), per-triangle material
 * index. The default instance is the identity (registry.cpp:73-74). */
int crb_scene_add_mesh(crb_scene *, const float *verts, const float *uvs, const uint32_t *mat_idx, uint32_t ntris, int *model_id);
/* wholesale material replacement (src/ui/ui.h:924-929) */
int crb_scene_set_materials(crb_scene *, int model_id, const crb_material *mats, uint32_t n);
/* wholesale instance replacement, column-major glm::mat4 (src/ui/ui.h:1185-1191) */
int crb_scene_set_instances(crb_scene *, int model_id, const float *mat4_colmajor, uint32_t n);
/* cr::image RGBA f32 row-major (src/objects/image.h) registered as an entity (registry.cpp:77-90) */
int crb_scene_add_texture(crb_scene *, const float *rgba, uint32_t w, uint32_t h, int *tex_id);
/* registry::set_sun + scene::set_sun_enabled (registry.cpp:248-256, scene.cpp:115-123) */
int crb_scene_set_sun(crb_scene *, const crb_sun *sun_or_null, int enabled);
/* scene::set_skybox + set_skybox_rotation (scene.cpp:36-65); rgba NULL removes it */
int crb_scene_set_skybox(crb_scene *, const float *rgba, uint32_t w, uint32_t h, float rot_u, float rot_v);
/* whole-struct camera assignment (src/ui/ui.h:674-675) */
int crb_scene_set_camera(crb_scene *, const crb_camera *);
/* model::instance_geometry -> rtcCommitGeometry/rtcAttachGeometry/rtcCommitScene (src/objects/model.cpp:52-97): builds
 * the 8-wide BVH(s) on the device. A scene whose models all have the single identity instance gets one BVH. An instanced
 * scene gets one object-space BVH per model + a top-level BVH over the (model, instance) pairs, and rays are taken into
 * object space per instance with the reference's own arithmetic (glm::inverse of the transform, renormalised direction,
 * hit point mapped back, distance re-measured: model.cpp:107-120) — hits are bit-identical to the reference's loop, every
 * model is stored once, and an instance edit (crb_scene_set_instances + commit) rebuilds only the top level.
 * CRB_SCENE_OPT_FLATTEN_INSTANCES = 1 selects the alternative: instances expanded into world-space triangles under ONE
 * BVH (faster traversal on heavily overlapping instances, I x the memory, hits equal to the reference only up to the
 * rounding of the changed coordinate frame, every instance edit is a full rebuild). */
int crb_scene_commit(crb_scene *, crb_build_info *info_or_null);
enum { CRB_SCENE_OPT_FLATTEN_INSTANCES = 1 };
int crb_scene_set_option(crb_scene *, int option, int value);

/* ---- queries: scene::cast_ray -> model::intersect -> rtcIntersect1 (scene.cpp:79-98,
 * model.cpp:5-49,99-126). rays/hits are host pointers unless on_device != 0. */
int crb_intersect_batch(crb_scene *, const crb_ray *rays, crb_hit *hits, uint64_t n, int on_device);
/* any-hit (rtcOccluded1 equivalent; the reference only ever calls rtcIntersect1) */
int crb_occluded_batch(crb_scene *, const crb_ray *rays, uint8_t *occluded, uint64_t n, int on_device);
/* instrumented traversal: mean node visits / triangle tests per ray for the roofline */
int crb_trace_counters(crb_scene *, const crb_ray *rays, uint64_t n, int on_device, int any_hit, uint64_t *node_visits,
                       uint64_t *tri_tests);
/* memory-system micro-benchmark for the roofline context: reads `bytes` of device memory `iters` times
 * with 16-byte loads and returns GB/s (working set < L2 size: L2 bandwidth; large: HBM read bandwidth) */
int crb_microbench_read(crb_scene *, uint64_t bytes, int iters, double *gb_per_s);
/* device time of the last crb_intersect_batch / crb_occluded_batch kernel, ms */
int crb_last_query_ms(crb_scene *, double *ms);

/* ---- renderer: cr::renderer (src/render/renderer.h:24-105) */
/* CRB_RENDER_FLAG_MATERIAL_SORT: sort the traced paths by shade class (miss/metal/smooth/glass) before shading */
/* CRB_RENDER_FLAG_EXTENDED: the shading BASELINE configs 2 and 5 name — Lambert / GGX metal (crb_material.roughness)
 * / Fresnel dielectric, NEE of the sun and of emissive triangles at diffuse vertices. Dead code in the reference
 * (src/render/brdf.h:10-29, src/util/sampling.h:83-142), specified by the oracle's extended mode. */
enum { CRB_RENDER_FLAG_COUNTERS = 1, CRB_RENDER_FLAG_TIMERS = 2, CRB_RENDER_FLAG_MATERIAL_SORT = 4, CRB_RENDER_FLAG_EXTENDED = 8 };
/* renderer::renderer(res_x,res_y,bounces,pool,scene) (renderer.cpp:106-145) + set_resolution's aspect
 * (renderer.cpp:194-208). seed keys the counter-based sampler (DESIGN.md "Sampler"). */
int crb_render_create(crb_scene *, uint32_t w, uint32_t h, uint32_t max_bounces, uint32_t seed, uint32_t flags, crb_render **out);
int crb_render_destroy(crb_render *);
/* renderer::start()'s clearing (renderer.cpp:154-170) */
int crb_render_reset(crb_render *);
int crb_render_set_resolution(crb_render *, uint32_t w, uint32_t h); /* renderer.cpp:194-208 */
int crb_render_set_max_bounces(crb_render *, uint32_t bounces);      /* renderer.cpp:210-213 */
/* picks up scene-side changes made since create (camera, sun, materials, re-commit) */
int crb_render_refresh(crb_render *);
/* restrict rendering to pixel rows [y0,y1) in sample space; default full frame. The reference's unit of work is one
 * scanline task (renderer.cpp:246-253); a row range is a set of them. */
int crb_render_set_rows(crb_render *, uint32_t y0, uint32_t y1);
/* ... or to interleaved row bands: bands of band_rows rows, this handle renders band `first`, first+stride, ... in ONE
 * launch sequence (what a rank of the tile partition renders). Single-GPU handles only. */
int crb_render_set_bands(crb_render *, uint32_t band_rows, uint32_t first, uint32_t stride);
/* The same with the owner order the library's own tile partition uses: serpentine != 0 reverses the order of the owners in
 * every other period of `stride` bands (period p odd: this handle renders band p*stride + stride-1-first), which cancels
 * a cost gradient along the image rows (measured on the 4K config-5 frame: max/mean of 8 ranks' times 1.05 -> 1.01). */
int crb_render_set_bands_ordered(crb_render *, uint32_t band_rows, uint32_t first, uint32_t stride, int serpentine);
/* n progressive passes, global sample indices first_sample..first_sample+n-1
 * (management thread + _get_tasks + _sample_pixel, renderer.cpp:116-144,240-384); asynchronous */
int crb_render_samples(crb_render *, uint32_t first_sample, uint32_t n);
int crb_render_sync(crb_render *);
/* The reference's management thread renders passes until _spp_target is reached (0 = for ever), one pass per loop
 * iteration (renderer.cpp:116-144; set_target_spp :215-218). crb_render_set_target_spp stores the target;
 * crb_render_run submits the passes still missing — from the current pass count up to the target, passes_per_call at a
 * time (the library batches them into wavefronts) — and returns without waiting, like renderer::start(). With target 0
 * it submits one call of passes_per_call passes; a host loop around it is the reference's "render for ever". */
int crb_render_set_target_spp(crb_render *, uint64_t target);
int crb_render_run(crb_render *, uint32_t passes_per_call, uint64_t *passes_submitted_total);
enum { CRB_RAW_SUM = 0, CRB_PROGRESS = 1, CRB_ALBEDO = 2, CRB_NORMAL = 3, CRB_DEPTH = 4 };
/* current_progress/normals/albedos/depths (renderer.cpp:220-238): w*h*4 floats, row-major, x/y
 * flipped exactly as the reference stores them. CRB_RAW_SUM = _raw_buffer as RGBA with A = the pixel's own pass
 * count (a render call may cover a row range only), so a RAW_SUM read is a complete checkpoint. */
int crb_render_read(crb_render *, int kind, float *dst_host);
/* The reference's UI thread reads those buffers while the workers keep rendering (lock-free getters,
 * renderer.cpp:220-238; display loop src/display/display.cpp:200-220). crb_render_read_async queues a consistent
 * snapshot of the buffer after the work submitted so far and copies it to dst_host (pinned memory for a truly
 * asynchronous copy) on a second stream, so the caller can submit the next crb_render_samples at once;
 * crb_render_read_wait blocks until the copy of that ticket has landed. crb_render_sync also waits for all of them. */
int crb_render_read_async(crb_render *, int kind, float *dst_host, uint64_t *ticket);
int crb_render_read_wait(crb_render *, uint64_t ticket);
int crb_render_stats(crb_render *, crb_stats *out); /* renderer::current_stats, renderer.cpp:396-404 */
/* The reference draws its random numbers from a default-seeded thread_local std::mt19937 in call order
 * (renderer.cpp:6-11); the library's default is a counter-based hash of (seed, pixel, sample, dimension). A caller-supplied
 * table replaces the hash: table[(sample * w*h + pixel) * dims + d], pixel = x + y*w in sample space, dimensions 0,1 =
 * pixel jitter, 2+4i+{0,1} = scatter of bounce i, 2+4i+{2,3} = sun sample of bounce i (ref-exact mode only; samples or
 * dimensions beyond the table fall back to the hash). Uses: low-discrepancy sequences, and replaying the reference's own
 * stream so that an image can be compared with the reference's (tests/test_reference_anchor.py). NULL removes it. */
int crb_render_set_sample_table(crb_render *, const float *table_host, uint32_t n_samples, uint32_t dims);
/* checkpoint / resume (the reference restarts from 0 spp on every start(), renderer.cpp:154-170): a
 * CRB_RAW_SUM read is a complete checkpoint; restore uploads it (w*h*4 floats) with its pass count, after
 * which crb_render_samples(first_sample = passes, ...) continues the same progressive render bit for bit */
int crb_render_restore(crb_render *, const float *raw_sum_rgba_host, uint32_t passes);
/* cr::post_processor::process (src/render/post/post_processor.cpp:110-296 + post_process.comp, blur.comp)
 * as CUDA kernels, no GL context: bloom (threshold, 10 blur passes), gray scale, 4 tonemappers. The first
 * form takes any host RGBA image, the second the renderer's device-resident display buffer. */
int crb_post_process(crb_scene *, const float *rgba_host, uint32_t w, uint32_t h, const crb_post_settings *, float *out_host);
int crb_render_post_process(crb_render *, const crb_post_settings *, float *out_host);
/* ---- multi-GPU. The reference is one process on one CPU; its management thread
 * (src/render/renderer.cpp:116-144) issues passes and the UI reads the image (src/ui/ui.h:358-372). The same two calls
 * — crb_render_samples and crb_render_read[_async] — drive N GPUs through a handle made by one of:
 *   crb_render_create_multi  one process, N GPUs: the library replicates `scene` (host description copied, BVH built on
 *                            every GPU), runs one submission thread per GPU and merges on side streams. With peer
 *                            access the merge is one kernel per GPU over NVLink peer memory (sum in rank order +
 *                            resolve, CRB_MERGE_PEER_KERNEL), else ncclCommInitAll + ncclAllReduce/ncclBroadcast.
 *   crb_render_create_rank   one process per GPU (torchrun / MPI): rank 0 calls crb_comm_unique_id, the host
 *                            broadcasts the 128 bytes by its own means, every rank creates its handle on its own
 *                            committed copy of the scene; the merge is ncclAllReduce (spp) or an all-gather of the row
 *                            bands as grouped ncclBroadcasts (tile) + the resolve behind it, on a side stream.
 * partition: CRB_PARTITION_SPP — every crb_render_samples(first, n) range is split into contiguous shares, rank g
 * renders samples [first + g*n/N, ...) of every pixel (BASELINE config 4); CRB_PARTITION_TILE — rank g renders all n
 * samples of the interleaved 8-row bands g, g+N, ... (owner order reversed in every other period of N bands) (BASELINE config 5). The sampler is keyed by global pixel and sample index,
 * so the union over ranks is the single-GPU set of paths; tile merges are bit-identical to one GPU, spp merges differ
 * by float summation order only. crb_render_read / _read_async return the MERGED image (a flush is implied;
 * crb_render_flush starts one early). In rank mode reading an AOV buffer is a collective call and crb_render_stats
 * reports this rank's share. set_rows / set_bands / set_pass_count are single-GPU calls. NCCL is dlopen'ed
 * ("libnccl.so.2", override with CRB_NCCL_LIB) at first use: CRB_ERR_NCCL if it is needed and missing. */
enum { CRB_PARTITION_SPP = 0, CRB_PARTITION_TILE = 1 };
enum { CRB_MERGE_NONE = 0, CRB_MERGE_PEER_KERNEL = 1, CRB_MERGE_NCCL = 2 };
int crb_render_create_multi(crb_scene *, const int *devices_or_null, int ngpus, int partition, uint32_t w, uint32_t h, uint32_t max_bounces,
                            uint32_t seed, uint32_t flags, crb_render **out);
int crb_comm_unique_id(void *id128); /* ncclGetUniqueId */
int crb_render_create_rank(crb_scene *, const void *id128, int rank, int nranks, int partition, uint32_t w, uint32_t h, uint32_t max_bounces,
                           uint32_t seed, uint32_t flags, crb_render **out);
int crb_render_flush(crb_render *); /* start merging the accumulators now (asynchronous); no-op on a single-GPU handle */
/* the handle's render stream(s) wait for the last flush: an event recorded on crb_render_stream afterwards covers the
 * collective + resolve (device timing of a whole step); no-op on a single-GPU handle */
int crb_render_join_flush(crb_render *);
int crb_render_info(crb_render *, int *ngpus_local, int *nranks, int *partition, int *merge_kind);
/* plumbing for a host that brings its own collective: the float4 accumulation buffer (device pointer, w*h*4 floats;
 * A = per-pixel pass count) to reduce in place, then crb_render_resolve re-resolves the display buffer with the
 * per-pixel counts; crb_render_set_pass_count overwrites every pixel's count. */
int crb_render_accum_ptr(crb_render *, void **device_ptr, uint64_t *n_floats);
int crb_render_set_pass_count(crb_render *, uint32_t passes);
int crb_render_resolve(crb_render *);
/* the CUDA stream the handle launches on (cudaStream_t as void*) for event timing by the caller */
int crb_render_stream(crb_render *, void **stream);
int crb_scene_stream(crb_scene *, void **stream);

#ifdef __cplusplus
}
#endif
#endif
