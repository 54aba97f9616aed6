// ref_harness.cpp — drives the REFERENCE'S OWN render path, compiled unmodified from /root/reference/src, headlessly.
// TEST INFRASTRUCTURE ONLY (built into oracle/_ref/ref_render by oracle/ref/Makefile; never part of the product).
//
// What is the reference here: src/render/{renderer,scene,camera,ray}.cpp, src/render/entities/registry.cpp,
// src/objects/{model,thread_pool}.cpp, src/render/material/material.cpp, src/render/timer.cpp, src/util/{logger,
// asset_loader}.cpp, src/glad/glad.c and every header they include (entt, stb, tinyexr, tinyobj are vendored under
// external/). What is NOT: glm, fmt and Embree are absent from the image (fetched from the network / system packages by
// the reference's CMake), so they are replaced by the shims in oracle/ref/shim (glm: restated operation order; fmt: log
// strings only; embree3: the oracle's BVH + triangle test). OpenGL: the reference uploads meshes and the skybox to GL for
// its draft mode (registry.cpp:99-209, scene.cpp:36-60); glad's function pointers are pointed at no-ops.
//
// The reference has no headless mode and no seeds. The harness does what the ImGui panels do
// (src/ui/ui.h:421,505,674,924,977,1023,1185: everything inside renderer::update) with a ONE-thread pool, so that the
// default-seeded thread_local std::mt19937 of renderer.cpp:6-11 is consumed by a single worker in call order. The
// renderer starts rendering the empty scene as soon as it is constructed (renderer.h:94); the number of draws those
// passes consumed before our update() is reported (2 per pixel per pass) so that the oracle can skip them.
//
//   ref_render render <scene.bin> <out.bin>
//   ref_render export <PNG|JPG|EXR|HDR> <w> <h> <rgba.f32> <name>      -> ./out/<name>.<ext> (asset_loader.cpp:348-377)
//   ref_render load_model <file.obj> <folder> <out.bin>                  (asset_loader.cpp:182-303)
//   ref_render load_picture <file> <out.bin>                             (asset_loader.cpp:305-319)
// everything renderer.h includes, first and untouched ...
#include <algorithm>
#include <array>
#include <cmath>
#include <filesystem>
#include <iostream>
#include <memory>
#include <random>
#include <type_traits>
#include <variant>

#include <objects/image.h>
#include <objects/thread_pool.h>
#include <render/brdf.h>
#include <render/camera.h>
#include <render/scene.h>
#include <render/timer.h>
#include <util/sampling.h>
// ... then the class itself with its private members readable: _raw_buffer (renderer.h:85) is the float3 sum the parity
// check compares bit for bit; access specifiers do not change the layout, the reference's own .cpp files are untouched
#define private public
#include <render/renderer.h>
#undef private
#include <util/asset_loader.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <unistd.h>

extern "C" unsigned long long ref_shim_intersect_calls();

namespace
{
    struct Reader
    {
        FILE *f;
        template<typename T>
        T get()
        {
            T v;
            if (fread(&v, sizeof(T), 1, f) != 1) fprintf(stderr, "ref_render: short read\n"), _exit(3);
            return v;
        }
        template<typename T>
        std::vector<T> vec(size_t n)
        {
            std::vector<T> v(n);
            if (n && fread(v.data(), sizeof(T), n, f) != n) fprintf(stderr, "ref_render: short read\n"), _exit(3);
            return v;
        }
    };
    template<typename T>
    void put(FILE *f, const T &v)
    {
        fwrite(&v, sizeof(T), 1, f);
    }
    template<typename T>
    void put(FILE *f, const T *p, size_t n)
    {
        fwrite(p, sizeof(T), n, f);
    }

    void gl_noops()
    {
        glad_glGenTextures             = [](GLsizei n, GLuint *t) { for (GLsizei i = 0; i < n; i++) t[i] = 1; };
        glad_glBindTexture             = [](GLenum, GLuint) {};
        glad_glTexParameteri           = [](GLenum, GLenum, GLint) {};
        glad_glTexImage2D              = [](GLenum, GLint, GLint, GLsizei, GLsizei, GLint, GLenum, GLenum, const void *) {};
        glad_glGenVertexArrays         = [](GLsizei n, GLuint *t) { for (GLsizei i = 0; i < n; i++) t[i] = 1; };
        glad_glGenBuffers              = [](GLsizei n, GLuint *t) { for (GLsizei i = 0; i < n; i++) t[i] = 1; };
        glad_glBindVertexArray         = [](GLuint) {};
        glad_glBindBuffer              = [](GLenum, GLuint) {};
        glad_glBufferData              = [](GLenum, GLsizeiptr, const void *, GLenum) {};
        glad_glVertexAttribPointer     = [](GLuint, GLint, GLenum, GLboolean, GLsizei, const void *) {};
        glad_glEnableVertexAttribArray = [](GLuint) {};
    }

    int render(const char *in, const char *out)
    {
        Reader r { fopen(in, "rb") };
        if (!r.f) return fprintf(stderr, "ref_render: cannot open %s\n", in), 2;
        const uint32_t w = r.get<uint32_t>(), h = r.get<uint32_t>(), bounces = r.get<uint32_t>(), spp = r.get<uint32_t>(), threads = r.get<uint32_t>();
        gl_noops();

        auto pool     = std::make_unique<cr::thread_pool>(threads);
        auto scene    = std::make_unique<cr::scene>();
        // main.cpp:5-10: the renderer renders the (still empty) scene from the moment it exists. It is created at the
        // reference's own default size 1024x1024 so that an empty pass takes milliseconds: the reference's thread-pool and
        // pause hand-shakes wait on condition variables without predicates (thread_pool.cpp:12-13,55-56,
        // renderer.cpp:172-183) and lose wake-ups when a pass finishes within microseconds.
        const uint32_t w0 = 1024, h0 = 1024;
        auto renderer = std::make_unique<cr::renderer>(w0, h0, bounces, &pool, &scene);
        uint64_t passes_before = 0;

        renderer->update([&] {
            passes_before = renderer->current_sample_count();    // empty-scene passes since construction: 2*w0*h0 draws each
            // ---- camera (ui.h:674-675: whole-struct assignment of a camera the UI moved with translate/rotate)
            const auto pos = r.vec<float>(3), rot = r.vec<float>(3);
            const float fov = r.get<float>(), scale = r.get<float>();
            const uint32_t mode = r.get<uint32_t>();
            auto cam     = cr::camera(glm::vec3(pos[0], pos[1], pos[2]), fov, mode ? cr::camera::mode::orthographic : cr::camera::mode::perspective);
            cam.scale    = scale;
            cam.rotation = glm::vec3(rot[0], rot[1], rot[2]);
            cam.rotate(glm::vec3(0.0f));    // -> _update_cache (camera.cpp:47-52,54-68)
            *scene->registry()->camera() = cam;
            // ---- sun (ui.h:505-510)
            const uint32_t sun_enabled = r.get<uint32_t>();
            auto           sun         = cr::entity::sun();
            sun.size = r.get<float>(), sun.intensity = r.get<float>();
            const auto sd = r.vec<float>(3), sc = r.vec<float>(3);
            sun.direction = glm::vec3(sd[0], sd[1], sd[2]), sun.colour = glm::vec3(sc[0], sc[1], sc[2]);
            scene->registry()->set_sun(sun);
            scene->set_sun_enabled(sun_enabled != 0);
            // ---- skybox (ui.h:1023-1045)
            const uint32_t sw = r.get<uint32_t>(), sh = r.get<uint32_t>();
            const auto     srot = r.vec<float>(2);
            if (sw && sh)
            {
                const auto px = r.vec<float>(size_t(sw) * sh * 4);
                scene->set_skybox(cr::image(px, sw, sh));
            }
            scene->set_skybox_rotation(glm::vec2(srot[0], srot[1]));
            // ---- textures + models (ui.h:969-981 -> scene::add_model)
            const uint32_t         ntex = r.get<uint32_t>();
            std::vector<cr::image> textures;
            for (uint32_t t = 0; t < ntex; t++)
            {
                const uint32_t tw = r.get<uint32_t>(), th = r.get<uint32_t>();
                textures.emplace_back(r.vec<float>(size_t(tw) * th * 4), tw, th);
            }
            const uint32_t nmodels = r.get<uint32_t>();
            std::vector<std::vector<glm::mat4>> all_instances;
            for (uint32_t m = 0; m < nmodels; m++)
            {
                const uint32_t ntris = r.get<uint32_t>();
                const auto     verts = r.vec<float>(size_t(ntris) * 9);
                const uint32_t has_uv = r.get<uint32_t>();
                const auto     uvs    = r.vec<float>(has_uv ? size_t(ntris) * 6 : 0);
                const auto     midx   = r.vec<uint32_t>(ntris);
                auto data = cr::asset_loader::model_data();
                data.name = "model" + std::to_string(m);
                data.vertices.resize(size_t(ntris) * 3), data.texture_coords.resize(size_t(ntris) * 3);
                for (size_t i = 0; i < size_t(ntris) * 3; i++)
                {
                    data.vertices[i]       = glm::vec3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
                    data.texture_coords[i] = has_uv ? glm::vec2(uvs[2 * i], uvs[2 * i + 1]) : glm::vec2(0.0f, 0.0f);
                    data.vertex_indices.push_back(uint32_t(i)), data.texture_indices.push_back(uint32_t(i)), data.normal_indices.push_back(uint32_t(-1));
                }
                data.material_indices = midx;
                const uint32_t nmats  = r.get<uint32_t>();
                for (uint32_t k = 0; k < nmats; k++)
                {
                    auto info           = cr::material::information();
                    info.shade_type     = cr::material::type(r.get<uint32_t>());
                    info.ior            = r.get<float>();
                    info.roughness      = r.get<float>();
                    info.reflectiveness = r.get<float>();
                    info.emission       = r.get<float>();
                    const auto c        = r.vec<float>(4);
                    info.colour         = glm::vec4(c[0], c[1], c[2], c[3]);
                    const int32_t tex   = r.get<int32_t>();
                    if (tex >= 0) info.tex = uint32_t(tex);
                    data.materials.emplace_back(info);
                }
                data.textures = textures;    // material.tex indexes the model's own texture list (registry.cpp:77-90)
                const uint32_t ninst = r.get<uint32_t>();
                std::vector<glm::mat4> inst(ninst);
                for (uint32_t i = 0; i < ninst; i++)
                {
                    const auto f = r.vec<float>(16);
                    for (int c = 0; c < 4; c++) inst[i][c] = glm::vec4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
                }
                all_instances.push_back(inst);
                scene->add_model(data);
            }
            // instance transforms (ui.h:1185-1191): wholesale replacement on the model's entity; the model entities are
            // the ones that carry cr::entity::instances, matched to our models by their name component
            {
                auto &reg = scene->registry()->entities;
                for (const auto entity : reg.view<std::string, cr::entity::instances>())
                {
                    const auto &name = reg.get<std::string>(entity);
                    if (name.rfind("model", 0) == 0) reg.get<cr::entity::instances>(entity).transforms = all_instances[size_t(std::stoul(name.substr(5)))];
                }
            }
            renderer->set_resolution(int(w), int(h));    // ui.h:421: also sets the aspect correction (renderer.cpp:199)
            renderer->set_max_bounces(int(bounces));
            renderer->set_target_spp(spp);
        });

        // wait for the target (the management thread then blocks on _start_cond_var, renderer.cpp:139-141)
        const unsigned long long calls0 = ref_shim_intersect_calls();
        const auto               t0     = std::chrono::steady_clock::now();
        while (renderer->current_sample_count() < spp)
        {
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(600)) return fprintf(stderr, "ref_render: timeout\n"), 4;
        }
        const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
        const unsigned long long calls = ref_shim_intersect_calls() - calls0;
        const auto stats = renderer->current_stats();

        FILE *o = fopen(out, "wb");
        if (!o) return 2;
        put(o, uint64_t(passes_before) * 2ull * w0 * h0);    // randf() draws consumed before the scene was handed over
        put(o, uint64_t(stats.total_rays));
        put(o, uint64_t(renderer->current_sample_count()));
        put(o, uint64_t(calls));    // rtcIntersect1 calls = (closest-hit + shadow queries) x (model, instance) pairs
        put(o, seconds);            // wall time from the end of update() to the last pass (all `threads` workers)
        put(o, renderer->_raw_buffer.data(), size_t(w) * h * 3);
        put(o, renderer->current_progress()->data(), size_t(w) * h * 4);
        put(o, renderer->current_albedos()->data(), size_t(w) * h * 4);
        put(o, renderer->current_normals()->data(), size_t(w) * h * 4);
        put(o, renderer->current_depths()->data(), size_t(w) * h * 4);
        fclose(o);
        fflush(stdout);
        _exit(0);    // ~renderer would dead-lock: start() does nothing unless paused (renderer.cpp:147-170)
    }

    int export_image(int argc, char **argv)
    {
        if (argc < 7) return 1;
        const std::string type = argv[2];
        const uint64_t    w = std::stoull(argv[3]), h = std::stoull(argv[4]);
        Reader            r { fopen(argv[5], "rb") };
        if (!r.f) return 2;
        const auto img = cr::image(r.vec<float>(w * h * 4), w, h);
        cr::asset_loader::export_framebuffer(img, argv[6],
                                             type == "PNG"   ? cr::asset_loader::image_type::PNG
                                             : type == "JPG" ? cr::asset_loader::image_type::JPG
                                             : type == "EXR" ? cr::asset_loader::image_type::EXR
                                                             : cr::asset_loader::image_type::HDR);
        return 0;
    }

    int load_model(int argc, char **argv)
    {
        if (argc < 5) return 1;
        const auto m = cr::asset_loader::load_model(argv[2], argv[3]);
        FILE      *o = fopen(argv[4], "wb");
        if (!o) return 2;
        auto putvec = [&](const auto &v) {
            put(o, uint64_t(v.size()));
            if (!v.empty()) put(o, v.data(), v.size());
        };
        put(o, uint64_t(m.vertices.size()));
        for (const auto &v : m.vertices) put(o, v.x), put(o, v.y), put(o, v.z);
        put(o, uint64_t(m.texture_coords.size()));
        for (const auto &v : m.texture_coords) put(o, v.x), put(o, v.y);
        put(o, uint64_t(m.normals.size()));
        for (const auto &v : m.normals) put(o, v.x), put(o, v.y), put(o, v.z);
        putvec(m.vertex_indices), putvec(m.material_indices), putvec(m.texture_indices), putvec(m.normal_indices);
        put(o, uint64_t(m.materials.size()));
        for (const auto &mm : m.materials)
        {
            put(o, uint32_t(mm.info.shade_type)), put(o, mm.info.ior), put(o, mm.info.roughness), put(o, mm.info.reflectiveness), put(o, mm.info.emission);
            put(o, mm.info.colour.x), put(o, mm.info.colour.y), put(o, mm.info.colour.z), put(o, mm.info.colour.w);
            put(o, int32_t(mm.info.tex.has_value() ? int32_t(mm.info.tex.value()) : -1));
            put(o, uint64_t(mm.info.name.size()));
            put(o, mm.info.name.data(), mm.info.name.size());
        }
        put(o, uint64_t(m.textures.size()));
        for (const auto &t : m.textures)
        {
            put(o, uint64_t(t.width())), put(o, uint64_t(t.height()));
            put(o, t.data(), size_t(t.width()) * t.height() * 4);
        }
        put(o, uint64_t(m.name.size()));
        put(o, m.name.data(), m.name.size());
        fclose(o);
        return 0;
    }

    int load_picture(int argc, char **argv)
    {
        if (argc < 4) return 1;
        const auto p = cr::asset_loader::load_picture(argv[2]);
        FILE      *o = fopen(argv[3], "wb");
        if (!o) return 2;
        put(o, int32_t(p.res.x)), put(o, int32_t(p.res.y));
        put(o, p.colour.data(), p.colour.size());
        fclose(o);
        return 0;
    }
}    // namespace

int main(int argc, char **argv)
{
    const std::string cmd = argc > 1 ? argv[1] : "";
    if (cmd == "render" && argc >= 4) return render(argv[2], argv[3]);
    if (cmd == "export") return export_image(argc, argv);
    if (cmd == "load_model") return load_model(argc, argv);
    if (cmd == "load_picture") return load_picture(argc, argv);
    fprintf(stderr, "usage: ref_render render|export|load_model|load_picture ...\n");
    return 1;
}
