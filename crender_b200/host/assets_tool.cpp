// assets_tool.cpp — command-line front end of assets.hpp for the tests (tests/test_assets.py compares it with
// crender_b200/assets.py on the same files). No GPU needed.
//   assets_tool load  model.obj dump.bin        model_data as: counts (7 x u64), then the arrays
//   assets_tool export image.bin name TYPE dir  image.bin = w h (i32) + RGBA f32 rows; prints the file written
#include "assets.hpp"

int main(int argc, char **argv)
{
    using namespace crb;
    try
    {
        if (argc >= 4 && std::string(argv[1]) == "load")
        {
            const model_data m = asset_loader::load_model(argv[2], argc > 4 ? argv[4] : "");
            FILE            *f = std::fopen(argv[3], "wb");
            if (!f) return 2;
            uint64_t ntex_floats = 0;
            for (const image &t : m.textures) ntex_floats += t.data.size();
            const uint64_t counts[7] = { m.vertices.size(), m.texture_coords.size(), m.vertex_indices.size(), m.texture_indices.size(),
                                         m.material_indices.size(), m.materials.size(), m.textures.size() };
            std::fwrite(counts, sizeof(counts), 1, f);
            std::fwrite(m.vertices.data(), sizeof(vec3), m.vertices.size(), f);
            std::fwrite(m.texture_coords.data(), sizeof(vec2), m.texture_coords.size(), f);
            std::fwrite(m.vertex_indices.data(), 4, m.vertex_indices.size(), f);
            std::fwrite(m.texture_indices.data(), 4, m.texture_indices.size(), f);
            std::fwrite(m.material_indices.data(), 4, m.material_indices.size(), f);
            for (const material &mm : m.materials)
            {
                const float rec[6] = { mm.info.colour[0], mm.info.colour[1], mm.info.colour[2], mm.info.colour[3], mm.info.emission,
                                       mm.info.tex ? float(*mm.info.tex) : -1.0f };
                std::fwrite(rec, sizeof(rec), 1, f);
                const uint32_t type = mm.info.shade_type, len = uint32_t(mm.info.name.size());
                std::fwrite(&type, 4, 1, f), std::fwrite(&len, 4, 1, f), std::fwrite(mm.info.name.data(), 1, len, f);
            }
            for (const image &t : m.textures)
            {
                const uint64_t wh[2] = { t.width, t.height };
                std::fwrite(wh, sizeof(wh), 1, f);
                std::fwrite(t.data.data(), 4, t.data.size(), f);
            }
            std::fclose(f);
            std::printf("%s\n", m.name.c_str());
            return 0;
        }
        if (argc >= 6 && std::string(argv[1]) == "export")
        {
            std::vector<uint8_t> raw;
            if (!asset_loader::detail::read_file(argv[2], raw) || raw.size() < 8) return 2;
            int32_t dims[2];
            memcpy(dims, raw.data(), 8);
            image im;
            im.width = uint64_t(dims[0]), im.height = uint64_t(dims[1]);
            im.data.resize(size_t(dims[0]) * dims[1] * 4);
            if (raw.size() < 8 + im.data.size() * 4) return 2;
            memcpy(im.data.data(), raw.data() + 8, im.data.size() * 4);
            const std::string           t    = argv[4];
            const asset_loader::image_type type = t == "PNG" ? asset_loader::image_type::PNG : t == "JPG" ? asset_loader::image_type::JPG
                                                  : t == "EXR" ? asset_loader::image_type::EXR : asset_loader::image_type::HDR;
            std::printf("%s\n", asset_loader::export_framebuffer(im, argv[3], type, argv[5]).c_str());
            return 0;
        }
        std::fprintf(stderr, "usage: assets_tool load model.obj dump.bin [folder] | export image.bin name PNG|HDR|EXR out_dir\n");
        return 2;
    }
    catch (const crb::error &e)
    {
        std::fprintf(stderr, "assets_tool: error %d: %s\n", e.code, e.what());
        return e.code;
    }
}
