// crender_cli.cpp — headless driver over crender.hpp (the reference has no headless mode: main() opens a
// GLFW window unconditionally, src/main.cpp:3-18). Renders the built-in Cornell box (BASELINE config 1) and
// writes the display buffer as a binary PFM-like dump (w h, then RGBA f32 rows) for the tests to compare
// with the Python API.
//
//   g++ -std=c++17 -O2 crender_cli.cpp -o crender_cli -L.. -lcrender_b200 -Wl,-rpath,'$ORIGIN/..'
//   ./crender_cli [--gpus N] [--partition spp|tile] out.bin [w h spp bounces seed]     N GPUs of this process (crb_render_create_multi)
//   ./crender_cli --obj model.obj name PNG|HDR|EXR [w h spp bounces]     what the reference's "load model" + "export"
//                 panels do (src/ui/ui.h:567-631, asset_loader.cpp:182-377): OBJ/MTL/PNG textures in, ./out/name.ext out,
//                 camera on the -Z side framing the model's bounds, default sun
#include "assets.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>

namespace
{
    using crb::vec3;

    void quad(crb::model_data &m, vec3 a, vec3 b, vec3 c, vec3 d, uint32_t mat)
    {
        const uint32_t base = uint32_t(m.vertices.size());
        m.vertices.insert(m.vertices.end(), { a, b, c, d });
        for (uint32_t i : { 0u, 1u, 2u, 0u, 2u, 3u }) m.vertex_indices.push_back(base + i);
        m.material_indices.push_back(mat), m.material_indices.push_back(mat);
    }

    void box(crb::model_data &m, vec3 ctr, vec3 half, float yaw_deg, uint32_t mat)
    {
        const double c = std::cos(yaw_deg * M_PI / 180.0), s = std::sin(yaw_deg * M_PI / 180.0);
        auto         P = [&](int x, int y, int z) {
            return vec3 { float(ctr[0] + c * x * half[0] + s * z * half[2]), float(ctr[1] + y * half[1]), float(ctr[2] - s * x * half[0] + c * z * half[2]) };
        };
        quad(m, P(-1, -1, -1), P(-1, 1, -1), P(1, 1, -1), P(1, -1, -1), mat);
        quad(m, P(-1, -1, 1), P(1, -1, 1), P(1, 1, 1), P(-1, 1, 1), mat);
        quad(m, P(-1, -1, -1), P(-1, -1, 1), P(-1, 1, 1), P(-1, 1, -1), mat);
        quad(m, P(1, -1, -1), P(1, 1, -1), P(1, 1, 1), P(1, -1, 1), mat);
        quad(m, P(-1, 1, -1), P(-1, 1, 1), P(1, 1, 1), P(1, 1, -1), mat);
        quad(m, P(-1, -1, -1), P(1, -1, -1), P(1, -1, 1), P(-1, -1, 1), mat);
    }

    // same geometry as crender_b200/scenes.py:cornell()
    crb::model_data cornell()
    {
        crb::model_data m;
        m.name = "cornell";
        auto mat = [](crb::vec4 colour, float emission) {
            crb::material x;
            x.info.colour = colour, x.info.emission = emission;
            return x;
        };
        m.materials = { mat({ 0.73f, 0.73f, 0.73f, 1 }, 0), mat({ 0.65f, 0.05f, 0.05f, 1 }, 0), mat({ 0.12f, 0.45f, 0.15f, 1 }, 0), mat({ 1, 1, 1, 1 }, 15.0f) };
        quad(m, { -1, -1, -1 }, { -1, -1, 1 }, { 1, -1, 1 }, { 1, -1, -1 }, 0);
        quad(m, { -1, 1, -1 }, { 1, 1, -1 }, { 1, 1, 1 }, { -1, 1, 1 }, 0);
        quad(m, { -1, -1, 1 }, { -1, 1, 1 }, { 1, 1, 1 }, { 1, -1, 1 }, 0);
        quad(m, { -1, -1, -1 }, { -1, 1, -1 }, { -1, 1, 1 }, { -1, -1, 1 }, 1);
        quad(m, { 1, -1, -1 }, { 1, -1, 1 }, { 1, 1, 1 }, { 1, 1, -1 }, 2);
        quad(m, { -0.3f, 0.998f, -0.3f }, { 0.3f, 0.998f, -0.3f }, { 0.3f, 0.998f, 0.3f }, { -0.3f, 0.998f, 0.3f }, 3);
        box(m, { -0.35f, -0.7f, -0.25f }, { 0.3f, 0.3f, 0.3f }, 18.0f, 0);
        box(m, { 0.35f, -0.4f, 0.35f }, { 0.3f, 0.6f, 0.3f }, -17.0f, 0);
        return m;
    }
}    // namespace

static int g_gpus = 1, g_partition = CRB_PARTITION_SPP;

static int render_obj(int argc, char **argv)
{
    namespace al = crb::asset_loader;
    const std::string file = argv[2], name = argc > 3 ? argv[3] : "render", t = argc > 4 ? argv[4] : "PNG";
    const int w = argc > 5 ? atoi(argv[5]) : 640, h = argc > 6 ? atoi(argv[6]) : 360, spp = argc > 7 ? atoi(argv[7]) : 64, bounces = argc > 8 ? atoi(argv[8]) : 5;
    const crb::model_data m = al::load_model(file);
    if (m.vertices.empty() || m.vertex_indices.empty()) throw crb::error(CRB_ERR_INVALID_ARG, "the OBJ file has no faces");
    vec3 lo = m.vertices[0], hi = m.vertices[0];
    for (const vec3 &v : m.vertices)
        for (int k = 0; k < 3; k++) lo[k] = std::fmin(lo[k], v[k]), hi[k] = std::fmax(hi[k], v[k]);
    crb::scene scn;
    scn.add_model(m);
    crb::camera cam;
    const float ext = std::fmax(hi[0] - lo[0], hi[1] - lo[1]);
    cam.fov         = 40.0f;
    cam.position    = { 0.5f * (lo[0] + hi[0]), 0.5f * (lo[1] + hi[1]), lo[2] - 0.75f * ext / std::tan(20.0f * float(M_PI) / 180.0f) - 0.05f * ext };
    scn.set_camera(cam);
    const crb_build_info info = scn.commit();
    std::unique_ptr<crb::renderer> rp(g_gpus > 1 ? new crb::renderer(uint64_t(w), uint64_t(h), uint64_t(bounces), &scn, 0, g_gpus, g_partition)
                                                 : new crb::renderer(uint64_t(w), uint64_t(h), uint64_t(bounces), &scn, 0));
    crb::renderer &r = *rp;
    r.set_target_spp(uint64_t(spp));
    r.start();
    const crb_stats   st   = r.current_stats();
    const std::string path = al::export_framebuffer(r.current_progress(), name, t == "EXR" ? al::image_type::EXR : t == "HDR" ? al::image_type::HDR : al::image_type::PNG);
    std::printf("%s: %llu triangles, %llu materials, %llu textures, build %.3f ms, %d spp at %dx%d in %.3f ms device time (%.1f Mrays/s) -> %s\n", m.name.c_str(),
                (unsigned long long) info.n_triangles, (unsigned long long) m.materials.size(), (unsigned long long) m.textures.size(), info.build_ms, spp, w, h,
                st.device_ms, st.device_ms > 0 ? double(st.total_queries) / st.device_ms / 1e3 : 0.0, path.c_str());
    return 0;
}

int main(int argc, char **argv)
{
    // leading options: --gpus N, --partition spp|tile
    while (argc > 2 && (std::string(argv[1]) == "--gpus" || std::string(argv[1]) == "--partition"))
    {
        if (std::string(argv[1]) == "--gpus")
            g_gpus = atoi(argv[2]);
        else
            g_partition = std::string(argv[2]) == "tile" ? CRB_PARTITION_TILE : CRB_PARTITION_SPP;
        argv[2] = argv[0];
        argv += 2, argc -= 2;
    }
    if (argc > 2 && std::string(argv[1]) == "--obj")
    {
        try
        {
            return render_obj(argc, argv);
        }
        catch (const crb::error &e)
        {
            std::fprintf(stderr, "crender_cli: error %d: %s\n", e.code, e.what());
            return e.code;
        }
    }
    const char *out = argc > 1 ? argv[1] : "cornell.bin";
    const int   w = argc > 2 ? atoi(argv[2]) : 256, h = argc > 3 ? atoi(argv[3]) : 256, spp = argc > 4 ? atoi(argv[4]) : 16;
    const int   bounces = argc > 5 ? atoi(argv[5]) : 8, seed = argc > 6 ? atoi(argv[6]) : 0;
    try
    {
        crb::scene scn;
        scn.add_model(cornell());
        scn.set_sun_enabled(false);
        crb::camera cam;
        cam.position = { 0.0f, 0.0f, -3.4f }, cam.fov = 40.0f;
        scn.set_camera(cam);
        const crb_build_info info = scn.commit();
        std::unique_ptr<crb::renderer> rp(g_gpus > 1 ? new crb::renderer(uint64_t(w), uint64_t(h), uint64_t(bounces), &scn, uint32_t(seed), g_gpus, g_partition)
                                                     : new crb::renderer(uint64_t(w), uint64_t(h), uint64_t(bounces), &scn, uint32_t(seed)));
        crb::renderer &r = *rp;
        r.set_target_spp(uint64_t(spp));
        r.start();
        const crb_stats  st = r.current_stats();
        const crb::image im = r.current_progress();
        if (g_gpus > 1) std::printf("%d GPUs, %s partition: ", g_gpus, g_partition == CRB_PARTITION_TILE ? "tile" : "spp");
        std::printf("cornell %dx%d %d spp: %llu triangles, %llu nodes, build %.3f ms, %llu queries in %.3f ms device time (%.1f Mrays/s)\n", w, h, spp,
                    (unsigned long long) info.n_triangles, (unsigned long long) info.n_nodes, info.build_ms, (unsigned long long) st.total_queries,
                    st.device_ms, st.device_ms > 0 ? double(st.total_queries) / st.device_ms / 1e3 : 0.0);
        if (FILE *f = std::fopen(out, "wb"))
        {
            const int32_t dims[2] = { w, h };
            std::fwrite(dims, sizeof(dims), 1, f);
            std::fwrite(im.data.data(), sizeof(float), im.data.size(), f);
            std::fclose(f);
        }
        return 0;
    }
    catch (const crb::error &e)
    {
        std::fprintf(stderr, "crender_cli: error %d: %s\n", e.code, e.what());
        return e.code;
    }
}
