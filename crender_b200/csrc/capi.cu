// capi.cu — the extern "C" boundary declared in include/crender_b200.h.
// Exceptions never cross the boundary: every entry point returns a crb_status and records the
// message for crb_last_error(). (The reference terminates the process instead: util/exception.h:9-13.)
#include "../../include/crender_b200.h"
#include "render.cuh"
#include "scene.cuh"

#include <new>
#include <string>

struct crb_scene
{
    crb::Scene s;
};
struct crb_render
{
    crb::Render r;
    crb_render(crb::Scene *s, uint32_t w, uint32_t h, uint32_t mb, uint32_t seed, uint32_t flags) : r(s, w, h, mb, seed, flags) {}
};

namespace
{
    thread_local std::string g_last_error;

    template<typename F>
    int guarded(F f)
    {
        try
        {
            f();
            return CRB_OK;
        }
        catch (const crb::Error &e)
        {
            g_last_error = e.what();
            return e.code;
        }
        catch (const std::bad_alloc &)
        {
            g_last_error = "host allocation failed";
            return CRB_ERR_OOM;
        }
        catch (const std::exception &e)
        {
            g_last_error = e.what();
            return CRB_ERR_GENERIC;
        }
    }
    void need(const void *p, const char *what)
    {
        if (!p) throw crb::Error(crb::ERR_INVALID_ARG, std::string("null argument: ") + what);
    }
}    // namespace

extern "C" {

const char *crb_last_error(void) { return g_last_error.c_str(); }

int crb_set_device(int device)
{
    return guarded([&] {
#ifndef CRB_EMU
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) throw crb::Error(crb::ERR_NO_DEVICE, "no CUDA device: crender_b200 has no CPU path");
        if (device >= 0) CRB_CUDA_CHECK(cudaSetDevice(device));
#endif
    });
}

int crb_device_info(char *name, int name_cap, int *sm_count, uint64_t *l2_bytes, uint64_t *hbm_bytes)
{
    return guarded([&] {
#ifndef CRB_EMU
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) throw crb::Error(crb::ERR_NO_DEVICE, "no CUDA device: crender_b200 has no CPU path");
        cudaDeviceProp p;
        CRB_CUDA_CHECK(cudaGetDeviceProperties(&p, dev));
        if (name && name_cap > 0) snprintf(name, size_t(name_cap), "%s", p.name);
        if (sm_count) *sm_count = p.multiProcessorCount;
        if (l2_bytes) *l2_bytes = uint64_t(p.l2CacheSize);
        if (hbm_bytes) *hbm_bytes = uint64_t(p.totalGlobalMem);
#else
        if (name && name_cap > 0) snprintf(name, size_t(name_cap), "emu");
        if (sm_count) *sm_count = 1;
        if (l2_bytes) *l2_bytes = 0;
        if (hbm_bytes) *hbm_bytes = 0;
#endif
    });
}

int crb_scene_create(crb_scene **out)
{
    return guarded([&] {
        need(out, "out");
        *out = new crb_scene();
    });
}
int crb_scene_destroy(crb_scene *s)
{
    return guarded([&] { delete s; });
}
int crb_scene_add_mesh(crb_scene *s, const float *verts, const float *uvs, const uint32_t *mat_idx, uint32_t ntris, int *model_id)
{
    return guarded([&] {
        need(s, "scene");
        const int id = s->s.add_mesh(verts, uvs, mat_idx, ntris);
        if (model_id) *model_id = id;
    });
}
int crb_scene_set_materials(crb_scene *s, int model, const crb_material *m, uint32_t n)
{
    return guarded([&] {
        need(s, "scene");
        s->s.set_materials(model, m, n);
    });
}
int crb_scene_set_instances(crb_scene *s, int model, const float *mats, uint32_t n)
{
    return guarded([&] {
        need(s, "scene");
        s->s.set_instances(model, mats, n);
    });
}
int crb_scene_add_texture(crb_scene *s, const float *rgba, uint32_t w, uint32_t h, int *tex_id)
{
    return guarded([&] {
        need(s, "scene");
        const int id = s->s.add_texture(rgba, w, h);
        if (tex_id) *tex_id = id;
    });
}
int crb_scene_set_sun(crb_scene *s, const crb_sun *sun, int enabled)
{
    return guarded([&] {
        need(s, "scene");
        if (sun) s->s.sun = *sun;
        s->s.sun_enabled = enabled != 0;
        s->s.version++;
    });
}
int crb_scene_set_skybox(crb_scene *s, const float *rgba, uint32_t w, uint32_t h, float ru, float rv)
{
    return guarded([&] {
        need(s, "scene");
        crb::Scene &S = s->s;
        if (rgba && w && h)
        {
            S.skybox.assign(rgba, rgba + size_t(w) * h * 4);
            S.sky_w = w, S.sky_h = h;
            S.upload_skybox();
        }
        else
            S.sky_w = S.sky_h = 0;
        S.sky_rot[0] = ru, S.sky_rot[1] = rv;
        S.version++;
    });
}
int crb_scene_set_camera(crb_scene *s, const crb_camera *c)
{
    return guarded([&] {
        need(s, "scene");
        need(c, "camera");
        if (c->mode > 1) throw crb::Error(crb::ERR_INVALID_ARG, "camera mode must be 0 (perspective) or 1 (orthographic)");
        s->s.camera = *c;
        s->s.version++;
    });
}
int crb_scene_commit(crb_scene *s, crb_build_info *info)
{
    return guarded([&] {
        need(s, "scene");
        s->s.commit();
        if (info)
        {
            info->build_ms    = s->s.build.build_ms;
            info->upload_ms   = s->s.upload_ms;
            info->n_triangles = s->s.build.n_tris;
            info->n_nodes     = s->s.build.n_nodes;
            info->node_bytes  = uint64_t(s->s.build.n_nodes) * 80;
            info->tri_bytes   = uint64_t(s->s.build.n_tris) * 48;
            info->max_depth   = s->s.build.max_depth;
            info->sah_cost    = s->s.build.sah_cost;
        }
    });
}

int crb_intersect_batch(crb_scene *s, const crb_ray *rays, crb_hit *hits, uint64_t n, int on_device)
{
    return guarded([&] {
        need(s, "scene");
        crb::intersect_batch(s->s, rays, hits, n, on_device != 0);
    });
}
int crb_occluded_batch(crb_scene *s, const crb_ray *rays, uint8_t *occ, uint64_t n, int on_device)
{
    return guarded([&] {
        need(s, "scene");
        crb::occluded_batch(s->s, rays, occ, n, on_device != 0);
    });
}
int crb_trace_counters(crb_scene *s, const crb_ray *rays, uint64_t n, int on_device, int any_hit, uint64_t *nodes, uint64_t *tris)
{
    return guarded([&] {
        need(s, "scene");
        crb::trace_counters(s->s, rays, n, on_device != 0, any_hit != 0, nodes, tris);
    });
}
int crb_microbench_read(crb_scene *s, uint64_t bytes, int iters, double *gb_per_s)
{
    return guarded([&] {
        need(s, "scene");
        need(gb_per_s, "gb_per_s");
        *gb_per_s = crb::read_bandwidth_gbs(s->s, size_t(bytes), iters);
    });
}
int crb_last_query_ms(crb_scene *s, double *ms)
{
    return guarded([&] {
        need(s, "scene");
        need(ms, "ms");
        *ms = s->s.last_query_ms;
    });
}
int crb_scene_stream(crb_scene *s, void **stream)
{
    return guarded([&] {
        need(s, "scene");
        need(stream, "stream");
        *stream = (void *) s->s.stream;
    });
}

int crb_post_process(crb_scene *s, const float *rgba, uint32_t w, uint32_t h, const crb_post_settings *ps, float *out)
{
    return guarded([&] {
        need(s, "scene");
        need(rgba, "rgba_host");
        need(ps, "settings");
        need(out, "out_host");
        if (!w || !h) throw crb::Error(crb::ERR_INVALID_ARG, "post_process: empty image");
        crb::DBuf<float4> src;
        src.alloc(size_t(w) * h);
        crb::dev_upload(src.p, rgba, size_t(w) * h * 16, s->s.stream);
        crb::post_process(s->s, src.p, w, h, *ps, out);
        crb::stream_sync(s->s.stream);
    });
}
int crb_render_post_process(crb_render *r, const crb_post_settings *ps, float *out)
{
    return guarded([&] {
        need(r, "render");
        need(ps, "settings");
        need(out, "out_host");
        r->r.sync();
        crb::post_process(*r->r.scene, r->r.display.p, r->r.w, r->r.h, *ps, out);
    });
}

// ------------------------------------------------------------------ renderer
int crb_render_create(crb_scene *s, uint32_t w, uint32_t h, uint32_t max_bounces, uint32_t seed, uint32_t flags, crb_render **out)
{
    return guarded([&] {
        need(s, "scene");
        need(out, "out");
        if (!w || !h) throw crb::Error(crb::ERR_INVALID_ARG, "render target must be non-empty");
        *out = new crb_render(&s->s, w, h, max_bounces, seed, flags);
    });
}
int crb_render_destroy(crb_render *r)
{
    return guarded([&] { delete r; });
}
int crb_render_reset(crb_render *r)
{
    return guarded([&] {
        need(r, "render");
        r->r.reset();
    });
}
int crb_render_set_resolution(crb_render *r, uint32_t w, uint32_t h)
{
    return guarded([&] {
        need(r, "render");
        if (!w || !h) throw crb::Error(crb::ERR_INVALID_ARG, "render target must be non-empty");
        r->r.set_resolution(w, h);
    });
}
int crb_render_set_max_bounces(crb_render *r, uint32_t b)
{
    return guarded([&] {
        need(r, "render");
        r->r.max_bounces = b;
    });
}
int crb_render_refresh(crb_render *r)
{
    return guarded([&] {
        need(r, "render");
        r->r.refresh();
    });
}
int crb_render_set_rows(crb_render *r, uint32_t y0, uint32_t y1)
{
    return guarded([&] {
        need(r, "render");
        r->r.set_rows(y0, y1);
    });
}
int crb_render_samples(crb_render *r, uint32_t first, uint32_t n)
{
    return guarded([&] {
        need(r, "render");
        r->r.render_samples(first, n);
    });
}
int crb_render_sync(crb_render *r)
{
    return guarded([&] {
        need(r, "render");
        r->r.sync();
    });
}
int crb_render_read(crb_render *r, int kind, float *dst)
{
    return guarded([&] {
        need(r, "render");
        need(dst, "dst");
        r->r.read(kind, dst);
    });
}
int crb_render_read_async(crb_render *r, int kind, float *dst, uint64_t *ticket)
{
    return guarded([&] {
        need(r, "render");
        need(dst, "dst");
        need(ticket, "ticket");
        *ticket = r->r.read_async(kind, dst);
    });
}
int crb_render_read_wait(crb_render *r, uint64_t ticket)
{
    return guarded([&] {
        need(r, "render");
        r->r.read_wait(ticket);
    });
}
int crb_render_stats(crb_render *r, crb_stats *out)
{
    return guarded([&] {
        need(r, "render");
        need(out, "out");
        r->r.stats(*out);
    });
}
int crb_render_restore(crb_render *r, const float *raw, uint32_t passes)
{
    return guarded([&] {
        need(r, "render");
        need(raw, "raw_sum_rgba_host");
        r->r.restore(raw, passes);
    });
}
int crb_render_accum_ptr(crb_render *r, void **p, uint64_t *n)
{
    return guarded([&] {
        need(r, "render");
        need(p, "device_ptr");
        *p = r->r.accum.p;
        if (n) *n = uint64_t(r->r.w) * r->r.h * 4;
    });
}
int crb_render_set_pass_count(crb_render *r, uint32_t passes)
{
    return guarded([&] {
        need(r, "render");
        r->r.passes = passes;
    });
}
int crb_render_resolve(crb_render *r)
{
    return guarded([&] {
        need(r, "render");
        r->r.resolve();
    });
}
int crb_render_stream(crb_render *r, void **stream)
{
    return guarded([&] {
        need(r, "render");
        need(stream, "stream");
        *stream = (void *) r->r.stream();
    });
}
}
