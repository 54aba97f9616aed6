// crender.hpp — C++ host-side mirror of CRender's render-facing classes over the C ABI
// (include/crender_b200.h). Header-only; link with libcrender_b200.so.
//
// Same names, argument meaning and defaults as the reference so that code written against cr::scene /
// cr::renderer reads the same (glm types are replaced by std::array so the header has no dependencies):
//   cr::material::information   src/render/material/material.h:31-41
//   cr::camera                  src/render/camera.h:11-41
//   cr::entity::sun             src/render/entities/components.h:23-29
//   cr::asset_loader::model_data src/util/asset_loader.h:16-30
//   cr::scene                   src/render/scene.h:19-52
//   cr::renderer                src/render/renderer.h:24-105
// Error behaviour: the reference calls cr::exit() (print + std::exit, util/exception.h:9-13); here every
// failing C-ABI call throws crb::error carrying the crb_status code and message.
#pragma once

#include "../../include/crender_b200.h"

#include <array>
#include <cstdint>
#include <functional>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

namespace crb
{
    struct error : std::runtime_error
    {
        int code;
        error(int c, const char *msg) : std::runtime_error(msg), code(c) {}
    };
    inline void check(int rc)
    {
        if (rc != CRB_OK) throw error(rc, crb_last_error());
    }

    using vec2 = std::array<float, 2>;
    using vec3 = std::array<float, 3>;
    using vec4 = std::array<float, 4>;
    using mat4 = std::array<float, 16>;    // column-major, like glm::mat4

    struct material
    {
        enum type : unsigned char
        {
            metal,
            smooth,
            glass
        };
        struct information
        {
            type                    shade_type     = smooth;
            float                   ior            = 1.5f;
            float                   roughness      = 0.5f;
            float                   reflectiveness = 1.0f;
            float                   emission       = 0.0f;
            vec4                    colour         = { 1, 1, 1, 1 };
            std::string             name           = "ERROR - Report";
            std::optional<uint32_t> tex;
        } info;
    };

    struct camera
    {
        enum class mode
        {
            perspective,
            orthographic
        };
        vec3  position { 5, 5, 0 };
        vec3  rotation { 0, 0, 0 };
        float fov          = 75;
        float scale        = 1;
        mode  current_mode = mode::perspective;
    };

    struct sun
    {
        float size      = 3.14159265359f / 48.0f;
        float intensity = 100.0f;
        vec3  direction { 0.624695f, -0.780869f, 0.0f };    // normalize(0.8,-1,0)
        vec3  colour { 1.0f, 0.9f, 0.7f };
    };

    struct image
    {
        uint64_t           width = 0, height = 0;
        std::vector<float> data;    // RGBA f32, row-major (cr::image)
    };

    // cr::asset_loader::model_data
    struct model_data
    {
        std::string           name;
        std::vector<vec3>     vertices;
        std::vector<material> materials;
        std::vector<vec2>     texture_coords;
        std::vector<image>    textures;
        std::vector<uint32_t> vertex_indices, material_indices, texture_indices;
    };

    class scene
    {
    public:
        scene() { check(crb_scene_create(&_h)); }
        ~scene() { crb_scene_destroy(_h); }
        scene(const scene &)            = delete;
        scene &operator=(const scene &) = delete;

        // scene::add_model -> registry::register_model (registry.cpp:51-97): de-index, register textures,
        // remap material texture handles; default instance = identity. Returns the model id.
        int add_model(const model_data &m)
        {
            std::vector<float> verts(m.vertex_indices.size() * 3), uvs;
            for (size_t i = 0; i < m.vertex_indices.size(); i++)
                for (int k = 0; k < 3; k++) verts[3 * i + k] = m.vertices[m.vertex_indices[i]][k];
            if (!m.texture_coords.empty() && !m.texture_indices.empty())
            {
                uvs.resize(m.texture_indices.size() * 2);
                for (size_t i = 0; i < m.texture_indices.size(); i++)
                    for (int k = 0; k < 2; k++) uvs[2 * i + k] = m.texture_coords[m.texture_indices[i]][k];
            }
            std::vector<int> handles;
            for (const image &t : m.textures)
            {
                int id = -1;
                check(crb_scene_add_texture(_h, t.data.data(), uint32_t(t.width), uint32_t(t.height), &id));
                handles.push_back(id);
            }
            int id = -1;
            check(crb_scene_add_mesh(_h, verts.data(), uvs.empty() ? nullptr : uvs.data(), m.material_indices.empty() ? nullptr : m.material_indices.data(),
                                     uint32_t(m.vertex_indices.size() / 3), &id));
            if (!m.materials.empty())
            {
                std::vector<material> mats = m.materials;
                for (material &mm : mats)
                    if (mm.info.tex) mm.info.tex = uint32_t(handles[*mm.info.tex]);
                set_materials(id, mats);
            }
            return id;
        }
        void set_materials(int model, const std::vector<material> &mats)
        {
            std::vector<crb_material> c(mats.size());
            for (size_t i = 0; i < mats.size(); i++)
            {
                const auto &s = mats[i].info;
                c[i]          = crb_material { uint32_t(s.shade_type), s.ior, s.roughness, s.reflectiveness, s.emission,
                                      { s.colour[0], s.colour[1], s.colour[2], s.colour[3] }, s.tex ? int32_t(*s.tex) : -1 };
            }
            check(crb_scene_set_materials(_h, model, c.data(), uint32_t(c.size())));
        }
        void set_instances(int model, const std::vector<mat4> &transforms)
        {
            check(crb_scene_set_instances(_h, model, transforms.empty() ? nullptr : transforms[0].data(), uint32_t(transforms.size())));
        }
        void set_skybox(const image &sky) { check(crb_scene_set_skybox(_h, sky.data.data(), uint32_t(sky.width), uint32_t(sky.height), _sky_rot[0], _sky_rot[1])), _sky = sky; }
        void set_skybox_rotation(const vec2 &r)
        {
            _sky_rot = r;
            check(crb_scene_set_skybox(_h, _sky.data.empty() ? nullptr : _sky.data.data(), uint32_t(_sky.width), uint32_t(_sky.height), r[0], r[1]));
        }
        void set_sun(const sun &s)
        {
            crb_sun c { s.size, s.intensity, { s.direction[0], s.direction[1], s.direction[2] }, { s.colour[0], s.colour[1], s.colour[2] } };
            check(crb_scene_set_sun(_h, &c, _sun_enabled));
        }
        void set_sun_enabled(bool v)
        {
            _sun_enabled = v;
            check(crb_scene_set_sun(_h, nullptr, v));
        }
        bool is_sun_enabled() const noexcept { return _sun_enabled; }
        void set_camera(const camera &c)
        {
            crb_camera cc { { c.position[0], c.position[1], c.position[2] }, { c.rotation[0], c.rotation[1], c.rotation[2] }, c.fov, c.scale,
                            c.current_mode == camera::mode::perspective ? 0u : 1u };
            check(crb_scene_set_camera(_h, &cc));
        }
        // the rtcCommitScene point; returns build information (device build time etc.)
        crb_build_info commit()
        {
            crb_build_info info {};
            check(crb_scene_commit(_h, &info));
            return info;
        }
        // batch form of scene::cast_ray
        void cast_rays(const std::vector<crb_ray> &rays, std::vector<crb_hit> &hits)
        {
            hits.resize(rays.size());
            check(crb_intersect_batch(_h, rays.data(), hits.data(), rays.size(), 0));
        }
        crb_scene *handle() noexcept { return _h; }

    private:
        crb_scene *_h = nullptr;
        bool       _sun_enabled = true;    // scene.h:45
        image      _sky;
        vec2       _sky_rot { 0, 0 };
    };

    class renderer
    {
    public:
        // renderer(res_x, res_y, bounces, pool, scene): the thread pool has no equivalent (the GPU grid
        // replaces it); seed keys the counter-based sampler.
        renderer(uint64_t res_x, uint64_t res_y, uint64_t bounces, scene *scn, uint32_t seed = 0) : _scene(scn), _res_x(res_x), _res_y(res_y)
        {
            check(crb_render_create(scn->handle(), uint32_t(res_x), uint32_t(res_y), uint32_t(bounces), seed, 0, &_h));
        }
        // the same renderer over `ngpus` GPUs of this process (devices 0..ngpus-1): the library replicates the scene,
        // splits every render(n) by sample index (CRB_PARTITION_SPP) or by interleaved 8-row bands (CRB_PARTITION_TILE) and
        // merges the accumulators; all getters return the merged image
        renderer(uint64_t res_x, uint64_t res_y, uint64_t bounces, scene *scn, uint32_t seed, int ngpus, int partition)
            : _scene(scn), _res_x(res_x), _res_y(res_y)
        {
            check(crb_render_create_multi(scn->handle(), nullptr, ngpus, partition, uint32_t(res_x), uint32_t(res_y), uint32_t(bounces), seed, 0, &_h));
        }
        ~renderer() { crb_render_destroy(_h); }
        renderer(const renderer &)            = delete;
        renderer &operator=(const renderer &) = delete;

        // renderer::start (renderer.cpp:154-170): clears the accumulation; renders to the target spp if set
        bool start()
        {
            check(crb_render_reset(_h));
            _next = 0;
            if (_spp_target)
            {
                // the library's own target loop (renderer.cpp:116-144): passes up to the target, 16 per call
                uint64_t total = 0;
                check(crb_render_set_target_spp(_h, _spp_target));
                check(crb_render_run(_h, 16, &total));
                check(crb_render_sync(_h));
                _next = uint32_t(total);
            }
            return true;
        }
        bool pause()
        {
            check(crb_render_sync(_h));
            return true;
        }
        // pause -> mutate -> restart from sample 0 (renderer.cpp:185-192)
        void update(const std::function<void()> &fn)
        {
            pause();
            fn();
            check(crb_render_refresh(_h));
            start();
        }
        void set_resolution(int x, int y)
        {
            check(crb_render_set_resolution(_h, uint32_t(x), uint32_t(y)));
            _res_x = uint64_t(x), _res_y = uint64_t(y), _next = 0;
        }
        void set_max_bounces(int b) { check(crb_render_set_max_bounces(_h, uint32_t(b))); }
        void set_target_spp(uint64_t t) { _spp_target = t; }
        // n progressive passes (what the reference's management thread issues in a loop, renderer.cpp:116-144)
        void render(uint32_t n_spp)
        {
            check(crb_render_samples(_h, _next, n_spp));
            check(crb_render_sync(_h));
            _next += n_spp;
        }
        image current_progress() { return read(CRB_PROGRESS); }
        image current_normals() { return read(CRB_NORMAL); }
        image current_albedos() { return read(CRB_ALBEDO); }
        image current_depths() { return read(CRB_DEPTH); }
        uint64_t current_sample_count()
        {
            crb_stats s = current_stats();
            return s.passes;
        }
        crb_stats current_stats()
        {
            crb_stats s {};
            check(crb_render_stats(_h, &s));
            return s;
        }
        std::array<uint64_t, 2> current_resolution() const noexcept { return { _res_x, _res_y }; }
        // live read while rendering goes on (the reference's lock-free getters, renderer.cpp:220-238): snapshot after
        // the work submitted so far, copied into dst (w*h*4 floats, pinned memory for a truly asynchronous copy)
        uint64_t current_progress_async(float *dst, int kind = CRB_PROGRESS)
        {
            uint64_t ticket = 0;
            check(crb_render_read_async(_h, kind, dst, &ticket));
            return ticket;
        }
        void wait_read(uint64_t ticket) { check(crb_render_read_wait(_h, ticket)); }

    private:
        image read(int kind)
        {
            image im;
            im.width = _res_x, im.height = _res_y;
            im.data.resize(size_t(_res_x) * _res_y * 4);
            check(crb_render_read(_h, kind, im.data.data()));
            return im;
        }
        scene      *_scene;
        crb_render *_h = nullptr;
        uint64_t    _res_x, _res_y, _spp_target = 0;
        uint32_t    _next = 0;
    };
}    // namespace crb
