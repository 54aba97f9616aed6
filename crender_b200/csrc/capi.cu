// capi.cu — the extern "C" boundary declared in include/crender_b200.h.
// Exceptions never cross the boundary: every entry point returns a crb_status and records the
// message for crb_last_error(). (The reference terminates the process instead: util/exception.h:9-13.)
// Every entry point that takes a handle runs under the handle's CUDA device (DeviceScope) and restores the
// caller's current device on return, so scenes on different GPUs can be driven from one host thread and a host
// whose other libraries switch devices (torch) cannot make the library allocate or launch on a foreign GPU.
#include "../../include/crender_b200.h"
#include "multi.cuh"
#include "render.cuh"
#include "scene.cuh"

#include <algorithm>
#include <memory>
#include <new>
#include <string>

struct crb_scene
{
    crb::Scene s;
};
// one handle type for the single-GPU renderer and the multi-GPU one (crb_render_create_multi / _rank)
struct crb_render
{
    std::unique_ptr<crb::Render>      r;
    std::unique_ptr<crb::MultiRender> m;
    crb::Scene                       *scene = nullptr;
    uint64_t                          target_spp = 0, submitted = 0;    // renderer::_spp_target and the passes handed to the device so far
    int device() const { return scene->device; }
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
    // entry points on a scene / renderer handle: null check + the handle's device
    template<typename F>
    int on_scene(crb_scene *s, F f)
    {
        return guarded([&] {
            need(s, "scene");
            crb::DeviceScope ds(s->s.device);
            f(s->s);
        });
    }
    template<typename F>
    int on_render(crb_render *r, F f)
    {
        return guarded([&] {
            need(r, "render");
            crb::DeviceScope ds(r->device());
            f(*r);
        });
    }
    void single_only(crb_render &h, const char *what)
    {
        if (h.m) throw crb::Error(crb::ERR_INVALID_ARG, std::string(what) + ": not available on a multi-GPU handle (the partition owns the rows / buffers)");
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
    return guarded([&] {
        if (!s) return;
        crb::DeviceScope ds(s->s.device);
        delete s;
    });
}
int crb_scene_add_mesh(crb_scene *s, const float *verts, const float *uvs, const uint32_t *mat_idx, uint32_t ntris, int *model_id)
{
    return on_scene(s, [&](crb::Scene &) {
        const int id = s->s.add_mesh(verts, uvs, mat_idx, ntris);
        if (model_id) *model_id = id;
    });
}
int crb_scene_set_materials(crb_scene *s, int model, const crb_material *m, uint32_t n)
{
    return on_scene(s, [&](crb::Scene &) {
        s->s.set_materials(model, m, n);
    });
}
int crb_scene_set_instances(crb_scene *s, int model, const float *mats, uint32_t n)
{
    return on_scene(s, [&](crb::Scene &) {
        s->s.set_instances(model, mats, n);
    });
}
int crb_scene_add_texture(crb_scene *s, const float *rgba, uint32_t w, uint32_t h, int *tex_id)
{
    return on_scene(s, [&](crb::Scene &) {
        const int id = s->s.add_texture(rgba, w, h);
        if (tex_id) *tex_id = id;
    });
}
int crb_scene_set_sun(crb_scene *s, const crb_sun *sun, int enabled)
{
    return on_scene(s, [&](crb::Scene &) {
        if (sun) s->s.sun = *sun;
        s->s.sun_enabled = enabled != 0;
        s->s.version++;
    });
}
int crb_scene_set_skybox(crb_scene *s, const float *rgba, uint32_t w, uint32_t h, float ru, float rv)
{
    return on_scene(s, [&](crb::Scene &) {
        crb::Scene &S = s->s;
        if (rgba && w && h)
        {
            S.skybox.assign(rgba, rgba + size_t(w) * h * 4);
            S.sky_w = w, S.sky_h = h;
            S.sky_version++;
            S.upload_skybox();
        }
        else
            S.sky_w = S.sky_h = 0, S.sky_version++;
        S.sky_rot[0] = ru, S.sky_rot[1] = rv;
        S.version++;
    });
}
int crb_scene_set_camera(crb_scene *s, const crb_camera *c)
{
    return on_scene(s, [&](crb::Scene &) {
        need(c, "camera");
        if (c->mode > 1) throw crb::Error(crb::ERR_INVALID_ARG, "camera mode must be 0 (perspective) or 1 (orthographic)");
        s->s.camera = *c;
        s->s.version++;
    });
}
int crb_scene_set_option(crb_scene *s, int option, int value)
{
    return on_scene(s, [&](crb::Scene &sc) {
        if (option != CRB_SCENE_OPT_FLATTEN_INSTANCES) throw crb::Error(crb::ERR_INVALID_ARG, "set_option: unknown option");
        if (sc.flatten_instances != (value != 0)) sc.flatten_instances = value != 0, sc.committed = false, sc.version++, sc.geom_version++;
    });
}
int crb_scene_commit(crb_scene *s, crb_build_info *info)
{
    return on_scene(s, [&](crb::Scene &) {
        s->s.commit();
        if (info)
        {
            info->build_ms    = s->s.build.build_ms;
            info->upload_ms   = s->s.upload_ms;
            info->n_triangles = s->s.build.n_tris;
            info->n_nodes     = s->s.build.n_nodes;
            info->node_bytes  = uint64_t(s->s.build.n_nodes) * crb::BVH8_NODE_U4 * 16;
            info->tri_bytes   = uint64_t(s->s.stored_tris) * crb::BVH8_TRI_F4 * 16 + (s->s.two_level ? uint64_t(s->s.n_instances) * (sizeof(crb::Instance) + crb::BVH8_TRI_F4 * 16) : 0);
            info->max_depth   = s->s.build.max_depth;
            info->sah_cost    = s->s.build.sah_cost;
        }
    });
}

int crb_intersect_batch(crb_scene *s, const crb_ray *rays, crb_hit *hits, uint64_t n, int on_device)
{
    return on_scene(s, [&](crb::Scene &) {
        crb::intersect_batch(s->s, rays, hits, n, on_device != 0);
    });
}
int crb_occluded_batch(crb_scene *s, const crb_ray *rays, uint8_t *occ, uint64_t n, int on_device)
{
    return on_scene(s, [&](crb::Scene &) {
        crb::occluded_batch(s->s, rays, occ, n, on_device != 0);
    });
}
int crb_trace_counters(crb_scene *s, const crb_ray *rays, uint64_t n, int on_device, int any_hit, uint64_t *nodes, uint64_t *tris)
{
    return on_scene(s, [&](crb::Scene &) {
        crb::trace_counters(s->s, rays, n, on_device != 0, any_hit != 0, nodes, tris);
    });
}
int crb_microbench_read(crb_scene *s, uint64_t bytes, int iters, double *gb_per_s)
{
    return on_scene(s, [&](crb::Scene &) {
        need(gb_per_s, "gb_per_s");
        *gb_per_s = crb::read_bandwidth_gbs(s->s, size_t(bytes), iters);
    });
}
int crb_last_query_ms(crb_scene *s, double *ms)
{
    return on_scene(s, [&](crb::Scene &) {
        need(ms, "ms");
        *ms = s->s.last_query_ms;
    });
}
int crb_scene_stream(crb_scene *s, void **stream)
{
    return on_scene(s, [&](crb::Scene &) {
        need(stream, "stream");
        *stream = (void *) s->s.stream;
    });
}

int crb_post_process(crb_scene *s, const float *rgba, uint32_t w, uint32_t h, const crb_post_settings *ps, float *out)
{
    return on_scene(s, [&](crb::Scene &) {
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
    return on_render(r, [&](crb_render &h) {
        need(ps, "settings");
        need(out, "out_host");
        if (h.m)
        {
            const float4 *src = h.m->merged_buffer(CRB_PROGRESS);
            h.m->sync();
            crb::post_process(*h.scene, src, h.m->w, h.m->h, *ps, out);
            return;
        }
        h.r->sync();
        crb::post_process(*h.scene, h.r->display.p, h.r->w, h.r->h, *ps, out);
    });
}

// ------------------------------------------------------------------ renderer
int crb_render_create(crb_scene *s, uint32_t w, uint32_t h, uint32_t max_bounces, uint32_t seed, uint32_t flags, crb_render **out)
{
    return on_scene(s, [&](crb::Scene &sc) {
        need(out, "out");
        if (!w || !h) throw crb::Error(crb::ERR_INVALID_ARG, "render target must be non-empty");
        auto hdl   = std::make_unique<crb_render>();
        hdl->scene = &sc;
        hdl->r     = std::make_unique<crb::Render>(&sc, w, h, max_bounces, seed, flags);
        *out       = hdl.release();
    });
}
int crb_render_create_multi(crb_scene *s, const int *devices, int ngpus, int partition, uint32_t w, uint32_t h, uint32_t max_bounces, uint32_t seed,
                            uint32_t flags, crb_render **out)
{
    return on_scene(s, [&](crb::Scene &sc) {
        need(out, "out");
        if (!w || !h) throw crb::Error(crb::ERR_INVALID_ARG, "render target must be non-empty");
        auto hdl   = std::make_unique<crb_render>();
        hdl->scene = &sc;
        hdl->m     = std::make_unique<crb::MultiRender>(&sc, devices, ngpus, partition, w, h, max_bounces, seed, flags);
        *out       = hdl.release();
    });
}
int crb_comm_unique_id(void *id128)
{
    return guarded([&] {
        need(id128, "id128");
        crb::nccl_unique_id(id128);
    });
}
int crb_render_create_rank(crb_scene *s, const void *id128, int rank, int nranks, int partition, uint32_t w, uint32_t h, uint32_t max_bounces,
                           uint32_t seed, uint32_t flags, crb_render **out)
{
    return on_scene(s, [&](crb::Scene &sc) {
        need(out, "out");
        if (!w || !h) throw crb::Error(crb::ERR_INVALID_ARG, "render target must be non-empty");
        auto hdl   = std::make_unique<crb_render>();
        hdl->scene = &sc;
        hdl->m     = std::make_unique<crb::MultiRender>(&sc, id128, rank, nranks, partition, w, h, max_bounces, seed, flags);
        *out       = hdl.release();
    });
}
int crb_render_info(crb_render *r, int *ngpus_local, int *nranks, int *partition, int *merge_kind)
{
    return on_render(r, [&](crb_render &h) {
        if (ngpus_local) *ngpus_local = h.m ? int(h.m->locals.size()) : 1;
        if (nranks) *nranks = h.m ? h.m->world : 1;
        if (partition) *partition = h.m ? h.m->partition : CRB_PARTITION_SPP;
        if (merge_kind) *merge_kind = !h.m || h.m->world == 1 ? CRB_MERGE_NONE : (h.m->fused_peers ? CRB_MERGE_PEER_KERNEL : CRB_MERGE_NCCL);
    });
}
int crb_render_destroy(crb_render *r)
{
    return guarded([&] {
        if (!r) return;
        crb::DeviceScope ds(r->device());
        delete r;
    });
}
int crb_render_reset(crb_render *r)
{
    return on_render(r, [&](crb_render &h) {
        h.m ? h.m->reset() : h.r->reset();
        h.submitted = 0;
    });
}
int crb_render_set_resolution(crb_render *r, uint32_t w, uint32_t hh)
{
    return on_render(r, [&](crb_render &h) {
        if (!w || !hh) throw crb::Error(crb::ERR_INVALID_ARG, "render target must be non-empty");
        h.m ? h.m->set_resolution(w, hh) : h.r->set_resolution(w, hh);
        h.submitted = 0;
    });
}
int crb_render_set_max_bounces(crb_render *r, uint32_t b)
{
    return on_render(r, [&](crb_render &h) {
        if (h.m)
            h.m->set_max_bounces(b);
        else
            h.r->max_bounces = b;
    });
}
int crb_render_refresh(crb_render *r)
{
    return on_render(r, [&](crb_render &h) { h.m ? h.m->refresh() : h.r->refresh(); });
}
int crb_render_set_rows(crb_render *r, uint32_t y0, uint32_t y1)
{
    return on_render(r, [&](crb_render &h) {
        single_only(h, "crb_render_set_rows");
        h.r->set_rows(y0, y1);
    });
}
int crb_render_set_bands(crb_render *r, uint32_t band_rows, uint32_t first, uint32_t stride)
{
    return on_render(r, [&](crb_render &h) {
        single_only(h, "crb_render_set_bands");
        h.r->set_bands(band_rows, first, stride);
    });
}
int crb_render_set_bands_ordered(crb_render *r, uint32_t band_rows, uint32_t first, uint32_t stride, int serpentine)
{
    return on_render(r, [&](crb_render &h) {
        single_only(h, "crb_render_set_bands_ordered");
        h.r->set_bands(band_rows, first, stride, serpentine != 0);
    });
}
int crb_render_samples(crb_render *r, uint32_t first, uint32_t n)
{
    return on_render(r, [&](crb_render &h) {
        h.m ? h.m->render_samples(first, n) : h.r->render_samples(first, n);
        h.submitted = std::max<uint64_t>(h.submitted, uint64_t(first) + n);
    });
}
int crb_render_set_target_spp(crb_render *r, uint64_t target)
{
    return on_render(r, [&](crb_render &h) { h.target_spp = target; });
}
int crb_render_run(crb_render *r, uint32_t passes_per_call, uint64_t *total)
{
    return on_render(r, [&](crb_render &h) {
        if (passes_per_call == 0) passes_per_call = 1;
        do
        {
            uint64_t n = passes_per_call;
            if (h.target_spp)
            {
                if (h.submitted >= h.target_spp) break;
                n = std::min<uint64_t>(n, h.target_spp - h.submitted);
            }
            h.m ? h.m->render_samples(uint32_t(h.submitted), uint32_t(n)) : h.r->render_samples(uint32_t(h.submitted), uint32_t(n));
            h.submitted += n;
        } while (h.target_spp);
        if (total) *total = h.submitted;
    });
}
int crb_render_flush(crb_render *r)
{
    return on_render(r, [&](crb_render &h) {
        if (h.m) h.m->flush();
    });
}
int crb_render_join_flush(crb_render *r)
{
    return on_render(r, [&](crb_render &h) {
        if (h.m) h.m->join_flush();
    });
}
int crb_render_sync(crb_render *r)
{
    return on_render(r, [&](crb_render &h) { h.m ? h.m->sync() : h.r->sync(); });
}
int crb_render_read(crb_render *r, int kind, float *dst)
{
    return on_render(r, [&](crb_render &h) {
        need(dst, "dst");
        h.m ? h.m->read(kind, dst) : h.r->read(kind, dst);
    });
}
int crb_render_read_async(crb_render *r, int kind, float *dst, uint64_t *ticket)
{
    return on_render(r, [&](crb_render &h) {
        need(dst, "dst");
        need(ticket, "ticket");
        *ticket = h.m ? h.m->read_async(kind, dst) : h.r->read_async(kind, dst);
    });
}
int crb_render_read_wait(crb_render *r, uint64_t ticket)
{
    return on_render(r, [&](crb_render &h) { h.m ? h.m->read_wait(ticket) : h.r->read_wait(ticket); });
}
int crb_render_stats(crb_render *r, crb_stats *out)
{
    return on_render(r, [&](crb_render &h) {
        need(out, "out");
        h.m ? h.m->stats(*out) : h.r->stats(*out);
        out->running_time       = out->device_ms * 1e-3;
        out->rays_per_second    = out->running_time > 0 ? double(out->ref_rays) / out->running_time : 0.0;
        out->samples_per_second = out->running_time > 0 ? double(out->passes) / out->running_time : 0.0;
    });
}
int crb_render_restore(crb_render *r, const float *raw, uint32_t passes)
{
    return on_render(r, [&](crb_render &h) {
        need(raw, "raw_sum_rgba_host");
        h.m ? h.m->restore(raw, passes) : h.r->restore(raw, passes);
        h.submitted = passes;
    });
}
int crb_render_accum_ptr(crb_render *r, void **p, uint64_t *n)
{
    return on_render(r, [&](crb_render &h) {
        need(p, "device_ptr");
        crb::Render &R = h.m ? *h.m->root().render : *h.r;    // multi handle: the first local rank's own accumulator
        *p             = R.accum.p;
        if (n) *n = uint64_t(R.w) * R.h * 4;
    });
}
int crb_render_set_pass_count(crb_render *r, uint32_t passes)
{
    return on_render(r, [&](crb_render &h) {
        single_only(h, "crb_render_set_pass_count");
        h.r->set_pass_count(passes);
    });
}
int crb_render_set_sample_table(crb_render *r, const float *table, uint32_t n_samples, uint32_t dims)
{
    return on_render(r, [&](crb_render &h) {
        single_only(h, "crb_render_set_sample_table");
        h.r->set_sample_table(table, n_samples, dims);
    });
}
int crb_render_resolve(crb_render *r)
{
    return on_render(r, [&](crb_render &h) { h.m ? h.m->resolve() : h.r->resolve(); });
}
int crb_render_stream(crb_render *r, void **stream)
{
    return on_render(r, [&](crb_render &h) {
        need(stream, "stream");
        *stream = (void *) (h.m ? h.m->root().render->stream() : h.r->stream());
    });
}
}
