// scene.cuh — host-side scene container and the device-resident scene the kernels read.
//
// Mirrors cr::scene + cr::registry (src/render/scene.h:19-52, src/render/entities/registry.cpp:51-97,
// 248-256) and the component structs of src/render/entities/components.h:18-97. Where the reference
// keeps one Embree scene per model and loops over models and instance transforms per ray
// (scene.cpp:83-95, model.cpp:107-123), this design flattens every (model, instance) pair into one
// world-space triangle array at commit time — HBM is 180 GB, a 10M-triangle flattened scene is <1 GB —
// and builds ONE 8-wide BVH over it, so a ray does a single traversal with no per-ray matrix inverse.
#pragma once
#include "../../include/crender_b200.h"
#include "bvh8.cuh"
#include "bvh_build.cuh"
#include "platform.cuh"

#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

namespace crb
{
#ifndef CRB_DMAT_ALIGN
#define CRB_DMAT_ALIGN 16
#endif
    struct alignas(CRB_DMAT_ALIGN) DMaterial    // 48 bytes = three 16-byte vector loads (12 scalar loads without the alignment)
    {
        float    colour[4];
        uint32_t shade_type;
        float    ior, reflectiveness, emission;
        int32_t  tex;
        float    roughness;    // extended shading mode only (dead in the reference, renderer.cpp:84-86)
        uint32_t pad[2];
    };

    struct DTexture
    {
        uint32_t w, h;
        uint32_t offset;    // into the float4 texel pool
        uint32_t pad;
    };

    struct DCamera
    {
        float    position[3];
        float    m[3][3];     // rotation part of _cached_matrix, m[column][row]
        float    trans[3];    // translation column (orthographic origin)
        float    w;           // 1/tan(fov/2)
        float    scale;
        uint32_t mode;
        float    aspect;
    };

    struct DSun
    {
        float    dir[3];          // sun.direction
        float    transform[9];    // mat3(tangent, normal, bitangent) of -dir, column-major
        float    size, intensity;
        float    colour[3];
        float    one_minus_cos;   // 1 - cos(size)
        float    pdf;             // 1/(tau*(1-cos(size)))
        uint32_t enabled;
    };

    struct FlatRange    // flat primitive ids [start, start+ntris) belong to (model, inst); in two-level mode range k = Instance k
    {
        uint32_t start, ntris, model, inst, src_start, pad[3];
    };

    // Everything a kernel needs, passed by value.
    struct DScene
    {
        Bvh8            bvh;
        const float4   *shade_tri;    // per SOURCE triangle: object-space unit normal (model.cpp:35), w = global material index
        const float    *obj_uvs;      // 6 floats per source triangle, or nullptr
        const uint32_t *flat_src;     // flat prim -> source triangle, or nullptr when identity
        const DMaterial *materials;
        const DTexture *textures;
        const float4   *texels;
        const float4   *skybox;
        uint32_t        sky_w, sky_h;
        float           sky_rot[2];
        const FlatRange *ranges;
        uint32_t        n_ranges;
        uint32_t        has_alpha;    // any material that can produce colour.w == 0
        Bvh2            bvh2;         // two-level mode: TLAS over instances + one object-space BLAS per model (bvh8.cuh)
        uint32_t        two_level;    // 1: trace through bvh2 with the reference's per-instance arithmetic; 0: bvh (flat)
        const float4   *lights;       // extended mode: 3 per emissive world-space triangle: (v0, Le.r) (e1, Le.g) (e2, Le.b)
        uint32_t        n_lights;
        DSun            sun;
        DCamera         cam;
    };

    // Host copy of the caller's arrays (the library copies at add_mesh so the caller may free them, SURVEY.md 8b "Ownership").
    // The copy of a large mesh is bound by the first touch of fresh pages, not by memcpy (36 MB: 28 ms into fresh pages,
    // 7 ms into touched ones on the development host; touching from 4 threads was 4x SLOWER there, so no threads), and an
    // interactive host loads and drops models repeatedly. So: resize() does not zero-fill first (default-initialising
    // construct), and large blocks are recycled through a bounded, exact-size free list instead of going back to the OS.
    struct HostBlockCache
    {
        static constexpr size_t         MIN_BLOCK = size_t(1) << 20, MAX_CACHED = size_t(1) << 30;
        std::mutex                      mu;
        std::multimap<size_t, void *>   free_blocks;
        size_t                          cached = 0;
    };
    inline HostBlockCache &host_block_cache()
    {
        static HostBlockCache *c = new HostBlockCache;    // never destroyed: scenes may outlive static destruction
        return *c;
    }
    template<class T>
    struct NoInitAlloc
    {
        using value_type = T;
        NoInitAlloc() = default;
        template<class U>
        NoInitAlloc(const NoInitAlloc<U> &)
        {
        }
        T *allocate(size_t n)
        {
            const size_t bytes = n * sizeof(T);
            if (bytes >= HostBlockCache::MIN_BLOCK)
            {
                HostBlockCache             &c = host_block_cache();
                std::lock_guard<std::mutex> lk(c.mu);
                auto                        it = c.free_blocks.find(bytes);
                if (it != c.free_blocks.end())
                {
                    void *p = it->second;
                    c.free_blocks.erase(it);
                    c.cached -= bytes;
                    return static_cast<T *>(p);
                }
            }
            return static_cast<T *>(::operator new(bytes));
        }
        void deallocate(T *p, size_t n)
        {
            const size_t bytes = n * sizeof(T);
            if (bytes >= HostBlockCache::MIN_BLOCK)
            {
                HostBlockCache             &c = host_block_cache();
                std::lock_guard<std::mutex> lk(c.mu);
                if (c.cached + bytes <= HostBlockCache::MAX_CACHED)
                {
                    c.free_blocks.insert({ bytes, p });
                    c.cached += bytes;
                    return;
                }
            }
            ::operator delete(p);
        }
        template<class U, class... A>
        void construct(U *p, A &&...a)
        {
            if constexpr (sizeof...(A) == 0)
                ::new (static_cast<void *>(p)) U;    // default-initialised: trivial types stay untouched
            else
                ::new (static_cast<void *>(p)) U(std::forward<A>(a)...);
        }
        template<class U>
        bool operator==(const NoInitAlloc<U> &) const
        {
            return true;
        }
        template<class U>
        bool operator!=(const NoInitAlloc<U> &) const
        {
            return false;
        }
    };
    template<class T>
    using HostVec = std::vector<T, NoInitAlloc<T>>;
    template<class T>
    inline void host_copy(HostVec<T> &dst, const T *src, size_t n)
    {
        dst.resize(n);
        if (n) memcpy(dst.data(), src, n * sizeof(T));
    }

    struct HostModel
    {
        uint64_t                  geom_id = 0;    // unique per add_mesh: tells commit() which BLASes are still valid
        uint32_t                  ntris = 0;
        HostVec<float>            verts, uvs;
        HostVec<uint32_t>         mat_idx;
        std::vector<crb_material> materials;
        std::vector<float>        transforms;    // 16 floats each, column-major
    };
    struct HostTexture
    {
        uint32_t           w, h;
        std::vector<float> rgba;
    };

    struct Scene
    {
        int          device = 0;
        int          n_sms  = 1;
        cudaStream_t stream = nullptr;

        std::vector<HostModel>   models;
        std::vector<HostTexture> textures;
        crb_sun                  sun;
        bool                     sun_enabled = true;    // scene.h:45
        std::vector<float>       skybox;
        uint32_t                 sky_w = 0, sky_h = 0;
        float                    sky_rot[2] = { 0, 0 };
        crb_camera               camera;

        // device state (valid after commit)
        bool             committed = false;
        uint64_t         version   = 0;    // bumped by every mutation/commit; renderers refresh on change
        uint64_t         geom_version = 0; // bumped by mutations that need a commit (meshes, instances, textures, material count)
        uint64_t         sky_version  = 0; // bumped when the skybox image changes
        DBuf<float>      d_wverts;         // kept only during the build
        DBuf<float4>     d_shade_tri;
        DBuf<float>      d_obj_uvs;
        DBuf<uint32_t>   d_flat_src;
        DBuf<DMaterial>  d_materials;
        DBuf<DTexture>   d_textures;
        DBuf<float4>     d_texels;
        DBuf<float4>     d_skybox;
        DBuf<FlatRange>  d_ranges;
        DBuf<float4>     d_lights;         // emissive triangles for the extended mode's area-light NEE
        uint32_t         n_lights = 0;
        DBuf<uint4>      d_nodes;          // flat mode: the one BVH; two-level mode: the TLAS
        DBuf<float4>     d_tris;
        // two-level mode (instanced scenes unless flatten_instances): BLASes survive instance edits
        bool             flatten_instances = false;    // crb_scene_set_option(CRB_SCENE_OPT_FLATTEN_INSTANCES): the r1 fast path
        bool             two_level = false;
        DBuf<uint4>      d_blas_nodes;
        DBuf<float4>     d_blas_tris;
        DBuf<Blas>       d_blas;
        DBuf<Instance>   d_inst;
        std::vector<uint64_t> blas_geom_ids;           // geometry the BLAS set was built from
        std::vector<Blas>     blas_table;
        double           blas_build_ms = 0, tlas_build_ms = 0;
        uint32_t         blas_nodes_total = 0, blas_depth = 0, tlas_nodes = 0, n_instances = 0;
        uint64_t         stored_tris = 0;    // triangles actually resident (two-level: every model once)
        std::vector<FlatRange> ranges;
        BuildStats       build;
        double           upload_ms = 0;
        uint32_t         n_flat    = 0;
        bool             has_alpha = false;
        double           last_query_ms = 0;

        Scene();
        ~Scene();
        int  add_mesh(const float *verts, const float *uvs, const uint32_t *mat_idx, uint32_t ntris);
        void set_materials(int model, const crb_material *m, uint32_t n);
        void set_instances(int model, const float *mats, uint32_t n);
        int  add_texture(const float *rgba, uint32_t w, uint32_t h);
        void commit();
        void commit_two_level();
        void upload_materials();    // also recomputes has_alpha and the light list
        void upload_skybox();
        DScene device_scene(uint32_t w, uint32_t h) const;    // camera aspect from the render target
        void   require_committed() const;
        // multi-GPU replicas (multi.cu): the host-side description of another scene, to be committed on this scene's GPU
        void copy_description_from(const Scene &src);
        void copy_light_state_from(const Scene &src);    // camera, sun, materials, skybox: no rebuild
        uint64_t src_sky_version = ~0ull;
    };

    // batch queries (trace.cu)
    void intersect_batch(Scene &s, const crb_ray *rays, crb_hit *hits, uint64_t n, bool on_device);
    void occluded_batch(Scene &s, const crb_ray *rays, uint8_t *occ, uint64_t n, bool on_device);
    void trace_counters(Scene &s, const crb_ray *rays, uint64_t n, bool on_device, bool any_hit, uint64_t *nodes, uint64_t *tris);
    double read_bandwidth_gbs(Scene &s, size_t bytes, int iters);
    // export-time post chain (post.cu)
    void post_process(Scene &sc, const float4 *d_src, uint32_t w, uint32_t h, const crb_post_settings &s, float *out_host);
}    // namespace crb
