/*
 * oracle.h — C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY. This is a CPU restatement of CRender's path-tracing hot path
 * (reference files cited per function in oracle.cpp). Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. The product (crender_b200/) never
 * links, imports or calls anything in oracle/.
 *
 * How it is pinned. The reference ships no tests, golden vectors or fixtures, and cannot be built as a whole here (Embree
 * 3.x, glm 0.9.9.8, fmt, GLFW and OIDN are neither vendored nor installable). But its OWN sources for this path —
 * renderer.cpp, scene.cpp, model.cpp, registry.cpp, camera.cpp, ray.cpp, asset_loader.cpp ... — do compile, unmodified,
 * against shim headers for glm / fmt / embree3 (oracle/ref, built into oracle/_ref/ref_render). Run with a one-thread pool
 * the reference is reproducible, and this oracle in reference-stream mode (orc_render_set_reference_stream) reproduces its
 * raw sums, display buffer, AOVs and ray count BIT FOR BIT on seven seeded scenes (tests/test_reference_anchor.py; the
 * reference's outputs are committed as tests/golden/reference_v1.npz). What remains restated rather than compiled is the
 * third-party arithmetic: glm's vector operations and Embree's BVH + triangle test (shim/glm/glm.hpp, embree_shim.cpp);
 * for those: the known-answer tests derivable from the reference source (tests/test_oracle_kat.py)
 * and an O(T) brute-force triangle loop that is ground truth for the oracle's own BVH (tests/test_oracle_bvh.py).
 */
#ifndef CRENDER_ORACLE_H
#define CRENDER_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cr::material::type order, src/render/material/material.h:14-19 */
enum { ORC_METAL = 0, ORC_SMOOTH = 1, ORC_GLASS = 2 };

/* cr::material::information, src/render/material/material.h:31-41 */
typedef struct orc_material
{
    uint32_t shade_type;     /* default smooth */
    float    ior;            /* 1.5 */
    float    roughness;      /* 0.5 */
    float    reflectiveness; /* 1 */
    float    emission;       /* 0 */
    float    colour[4];      /* 1,1,1,1 */
    int32_t  tex;            /* -1 = none, else texture id returned by orc_scene_add_texture */
} orc_material;

/* cr::entity::sun, src/render/entities/components.h:23-29 */
typedef struct orc_sun
{
    float size;
    float intensity;
    float direction[3];
    float colour[3];
} orc_sun;

/* cr::camera, src/render/camera.h:11-41 */
typedef struct orc_camera
{
    float    position[3];
    float    rotation[3]; /* degrees; x about UP, y about RIGHT, z about FORWARD */
    float    fov;         /* degrees */
    float    scale;       /* orthographic only */
    uint32_t mode;        /* 0 perspective, 1 orthographic */
} orc_camera;

typedef struct orc_ray
{
    float o[3], tmin;
    float d[3], tmax;
} orc_ray;

typedef struct orc_hit
{
    float    t; /* +inf on miss */
    float    u, v;
    uint32_t prim;  /* triangle index inside the model; 0xffffffff on miss */
    uint32_t model; /* model id; 0xffffffff on miss */
    uint32_t inst;  /* instance index inside the model */
} orc_hit;

typedef struct orc_stats
{
    uint64_t total_queries; /* every closest-hit query issued (segments + shadow + alpha marches) */
    uint64_t ref_rays;      /* the reference's `_total_rays` rule, renderer.cpp:271-272,356 */
    uint64_t pixel_samples;
    uint64_t passes; /* the reference's `_current_sample` */
} orc_stats;

typedef struct orc_scene  orc_scene;
typedef struct orc_render orc_render;

orc_scene *orc_scene_create(void);
void       orc_scene_destroy(orc_scene *);
/* verts: 9 floats per triangle (de-indexed, registry.cpp:51-97); uvs: 6 floats per triangle or NULL */
int  orc_scene_add_mesh(orc_scene *, const float *verts, const float *uvs, const uint32_t *mat_idx, uint32_t ntris);
int  orc_scene_set_materials(orc_scene *, int model, const orc_material *mats, uint32_t n);
int  orc_scene_set_instances(orc_scene *, int model, const float *mat4_colmajor, uint32_t n);
int  orc_scene_add_texture(orc_scene *, const float *rgba, uint32_t w, uint32_t h);
void orc_scene_set_sun(orc_scene *, const orc_sun *, int enabled);
void orc_scene_set_skybox(orc_scene *, const float *rgba, uint32_t w, uint32_t h, float rot_u, float rot_v);
void orc_scene_set_camera(orc_scene *, const orc_camera *);
/* builds one BVH per model (stands in for rtcCommitScene, model.cpp:92-94); returns build ms */
double orc_scene_commit(orc_scene *);

/* Embree-style single-model queries in WORLD space through scene::cast_ray semantics are in
 * orc_cast_batch; orc_intersect_batch is the raw rtcIntersect1-equivalent with caller tmin/tmax. */
void orc_intersect_batch(orc_scene *, const orc_ray *rays, orc_hit *hits, uint64_t n, int nthreads);
void orc_intersect_brute(orc_scene *, const orc_ray *rays, orc_hit *hits, uint64_t n, int nthreads);
void orc_occluded_batch(orc_scene *, const orc_ray *rays, uint8_t *occluded, uint64_t n, int nthreads);

orc_render *orc_render_create(orc_scene *, uint32_t w, uint32_t h, uint32_t max_bounces, uint32_t seed);
void        orc_render_destroy(orc_render *);
void        orc_render_reset(orc_render *);
/* extended shading mode (GGX metal, Fresnel dielectric, area-light + sun NEE at diffuse vertices): specified
 * by the oracle itself, see "EXTENDED shading mode" in oracle.cpp — the reference has only dead code for it.
 * on = 2: extended without the light list (emitters are found by path hits only), used to cross-check the NEE estimator */
void        orc_render_set_extended(orc_render *, int on);
/* rows [y0,y1) only (bounded CPU samples for benchmarks); full frame = 0,h */
void orc_render_set_rows(orc_render *, uint32_t y0, uint32_t y1);
/* renders passes first_sample .. first_sample+n-1, one task per scanline per pass (renderer.cpp:240-256) */
void orc_render_samples(orc_render *, uint32_t first_sample, uint32_t n, int nthreads);
enum { ORC_RAW_SUM = 0, ORC_PROGRESS = 1, ORC_ALBEDO = 2, ORC_NORMAL = 3, ORC_DEPTH = 4 };
/* dst: w*h*4 floats, row-major, already x/y-flipped as the reference stores them */
void orc_render_read(orc_render *, int kind, float *dst);
void orc_render_stats(orc_render *, orc_stats *out);
/* reference-stream sampler (on != 0): the reference's own default-seeded std::mt19937 consumed in call order by ONE
 * thread (renderer.cpp:6-11), after skipping `discard` draws; only for the comparison with the reference's sources
 * compiled under oracle/ref (render with nthreads = 1) */
void orc_render_set_reference_stream(orc_render *, int on, uint64_t discard);
/* sample table [n_samples][w*h][dims] floats (caller-owned): recorded while rendering with the reference stream, or — with
 * replay != 0 — the sampler itself (the CUDA path replays the same table: crb_render_set_sample_table) */
void orc_render_set_sample_table(orc_render *, float *table, uint32_t n_samples, uint32_t dims, int replay);
/* the bare BVH + triangle test of one model (what stands in for the Embree scene): for oracle/ref's embree3 shim */
typedef struct orc_rawbvh orc_rawbvh;
orc_rawbvh *orc_rawbvh_create(const float *verts9, uint32_t ntris);
void        orc_rawbvh_destroy(orc_rawbvh *);
int         orc_rawbvh_intersect(const orc_rawbvh *, const float *o3, const float *d3, float tnear, float tfar, float *t, float *u, float *v, uint32_t *prim);
/* primary-ray (bounce 0) hits for sample `sample` of every pixel, in sample-space pixel order */
void orc_render_primary_hits(orc_render *, uint32_t sample, orc_hit *hits, int nthreads);

/* ---- known-answer probes ---- */
void  orc_kat_mt19937_randf(uint32_t n, float *out);                        /* renderer.cpp:6-11 */
float orc_kat_rng(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t dim);
void  orc_kat_camera_ray(const orc_camera *, float x, float y, float aspect, float *o3, float *d3);
void  orc_kat_build_local(const float *n3, float *tangent3, float *bitangent3);
void  orc_kat_sun_transform(const float *sun_dir3, float *mat3_colmajor9);
void  orc_kat_map_to_solid_angle(float u, float v, float theta_max, float *out3, float *pdf);
void  orc_kat_sphere(float u, float v, float *out3);
/* process_hit on a synthetic record: in: shade material, normal, point, ray dir, u0,u1;
 * out: origin3, dir3, albedo3, is_alpha */
void  orc_kat_process_hit(const orc_material *, const float *normal3, const float *point3,
                          const float *raydir3, float u0, float u1, float *origin3, float *dir3,
                          float *albedo3, int *is_alpha);
/* extended-mode scatter on a synthetic record: out weight3, origin3, dir3; returns 0 scattered diffuse,
 * 1 scattered specular, 2 absorbed */
int   orc_kat_scatter_extended(const orc_material *, const float *normal3, const float *point3, const float *raydir3, float u0, float u1,
                               float *weight3, float *origin3, float *dir3);
float orc_kat_resolve(float sum, uint32_t n_plus_1);
int   orc_kat_tri(const float *v0, const float *v1, const float *v2, const float *o, const float *d,
                  float tmin, float tmax, float *t, float *u, float *v);
void  orc_kat_sky_uv(const float *d3, float *uv2);
void  orc_kat_image_get_uv_index(float u, float v, uint32_t w, uint32_t h, uint32_t *xy2);

#ifdef __cplusplus
}
#endif
#endif
