// scene.cu — host-side scene container: upload, instance flattening, commit (BVH build).
// See scene.cuh for the mapping to cr::scene / cr::registry.
#include "scene.cuh"

#include <chrono>
#include <cmath>

namespace crb
{
    namespace
    {
        constexpr float PI_F  = 3.14159265359f;    // src/util/numbers.h:15
        constexpr float TAU_F = 6.28318530717f;    // src/util/numbers.h:17

        // world = M * (v,1), glm mat4*vec4 order: (m0*x + m1*y) + (m2*z + m3*1). Identity transforms are
        // copied bit-exactly so that a default-instanced model is traced in its own coordinates.
        __global__ void k_flatten(const float *__restrict__ src, uint32_t ntris, const float *__restrict__ M, int identity, float *__restrict__ dst,
                                  uint32_t *__restrict__ flat_src, uint32_t src_start)
        {
            const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= ntris) return;
            const float *v = src + size_t(i) * 9;
            float       *o = dst + size_t(i) * 9;
            if (identity)
            {
#pragma unroll
                for (int k = 0; k < 9; k++) o[k] = v[k];
            }
            else
            {
#pragma unroll
                for (int k = 0; k < 3; k++)
                {
                    const float x = v[3 * k], y = v[3 * k + 1], z = v[3 * k + 2];
#pragma unroll
                    for (int r = 0; r < 3; r++)
                        o[3 * k + r] = __fadd_rn(__fadd_rn(__fmul_rn(M[0 + r], x), __fmul_rn(M[4 + r], y)), __fadd_rn(__fmul_rn(M[8 + r], z), M[12 + r]));
                }
            }
            if (flat_src) flat_src[i] = src_start + i;
        }

        // per source triangle: normalize((v1-v0) x (v2-v0)) in OBJECT space (model.cpp:35; the
        // reference never maps the normal back to world space, model.cpp:116-120), w = material index
        __global__ void k_shade_tri(const float *__restrict__ src, const uint32_t *__restrict__ mat_idx, uint32_t mat_base, uint32_t ntris,
                                    float4 *__restrict__ out)
        {
            const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= ntris) return;
            const float *v  = src + size_t(i) * 9;
            const V3     v0 = ld3(v), v1 = ld3(v + 3), v2 = ld3(v + 6);
            const V3     n  = normalize(cross(v1 - v0, v2 - v0));
            out[i]          = make_float4(n.x, n.y, n.z, __uint_as_float(mat_base + mat_idx[i]));
        }

        bool is_identity16(const float *m)
        {
            static const float I[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
            return memcmp(m, I, sizeof(I)) == 0;
        }

        crb_material default_material()
        {
            crb_material m {};    // material.h:31-41
            m.shade_type = CRB_SMOOTH, m.ior = 1.5f, m.roughness = 0.5f, m.reflectiveness = 1.0f, m.emission = 0.0f;
            m.colour[0] = m.colour[1] = m.colour[2] = m.colour[3] = 1.0f;
            m.tex = -1;
            return m;
        }

        // glm::rotate(m, angle, unit axis) restricted to the 3x3 part, m[column][row]
        void rotate3(float m[3][3], float angle, const float ax[3])
        {
            const float c = std::cos(angle), s = std::sin(angle);
            const float t[3] = { ax[0] * (1.0f - c), ax[1] * (1.0f - c), ax[2] * (1.0f - c) };
            float       R[3][3];
            R[0][0] = c + t[0] * ax[0], R[0][1] = t[0] * ax[1] + s * ax[2], R[0][2] = t[0] * ax[2] - s * ax[1];
            R[1][0] = t[1] * ax[0] - s * ax[2], R[1][1] = c + t[1] * ax[1], R[1][2] = t[1] * ax[2] + s * ax[0];
            R[2][0] = t[2] * ax[0] + s * ax[1], R[2][1] = t[2] * ax[1] - s * ax[0], R[2][2] = c + t[2] * ax[2];
            float r[3][3];
            for (int k = 0; k < 3; k++)
                for (int row = 0; row < 3; row++) r[k][row] = (m[0][row] * R[k][0] + m[1][row] * R[k][1]) + m[2][row] * R[k][2];
            memcpy(m, r, sizeof(r));
        }
    }    // namespace

    Scene::Scene()
    {
#ifndef CRB_EMU
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
            throw Error(ERR_NO_DEVICE, "no CUDA device: crender_b200 has no CPU path");
        CRB_CUDA_CHECK(cudaGetDevice(&device));
        CRB_CUDA_CHECK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device));
        CRB_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
#endif
        // components.h:23-29
        sun.size      = PI_F / 48.0f;
        sun.intensity = 100.0f;
        const float x = 0.8f, y = -1.0f, z = 0.0f;
        const float inv = 1.0f / std::sqrt((x * x + y * y) + z * z);
        sun.direction[0] = x * inv, sun.direction[1] = y * inv, sun.direction[2] = z * inv;
        sun.colour[0] = 1.0f, sun.colour[1] = 0.9f, sun.colour[2] = 0.7f;
        // camera.h:21
        camera = crb_camera {};
        camera.position[0] = 5, camera.position[1] = 5, camera.position[2] = 0;
        camera.fov = 75, camera.scale = 1, camera.mode = 0;
    }

    Scene::~Scene()
    {
#ifndef CRB_EMU
        if (stream) cudaStreamDestroy(stream);
#endif
    }

    int Scene::add_mesh(const float *verts, const float *uvs, const uint32_t *mat_idx, uint32_t ntris)
    {
        if (!verts && ntris) throw Error(ERR_BUILD_VERTS, "add_mesh: null vertex buffer");
        static uint64_t next_geom_id = 1;
        HostModel m;
        m.geom_id = next_geom_id++;
        m.ntris = ntris;
        host_copy(m.verts, verts, size_t(ntris) * 9);
        if (uvs) host_copy(m.uvs, uvs, size_t(ntris) * 6);
        if (mat_idx)
            host_copy(m.mat_idx, mat_idx, size_t(ntris));
        else
            m.mat_idx.assign(ntris, 0);
        uint32_t maxm = 0;
        for (uint32_t i : m.mat_idx) maxm = i > maxm ? i : maxm;
        m.materials.assign(size_t(maxm) + 1, default_material());
        static const float I[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };    // registry.cpp:73-74
        m.transforms.assign(I, I + 16);
        models.push_back(std::move(m));
        committed = false;
        version++, geom_version++;
        return int(models.size()) - 1;
    }

    void Scene::set_materials(int model, const crb_material *mats, uint32_t n)
    {
        if (model < 0 || size_t(model) >= models.size() || !mats) throw Error(ERR_INVALID_ARG, "set_materials: bad model id");
        HostModel &m = models[size_t(model)];
        for (uint32_t i : m.mat_idx)
            if (i >= n) throw Error(ERR_INVALID_ARG, "set_materials: a triangle references material " + std::to_string(i) + " >= " + std::to_string(n));
        for (uint32_t i = 0; i < n; i++)
            if (mats[i].tex >= int32_t(textures.size())) throw Error(ERR_INVALID_ARG, "set_materials: unknown texture id");
        const bool same_count = m.materials.size() == n;
        m.materials.assign(mats, mats + n);
        version++;
        if (committed && same_count)
            upload_materials();    // material edits do not need a rebuild (ui.h:924-929 path)
        else
            committed = false, geom_version++;
    }

    void Scene::set_instances(int model, const float *mats, uint32_t n)
    {
        if (model < 0 || size_t(model) >= models.size() || (!mats && n)) throw Error(ERR_INVALID_ARG, "set_instances: bad arguments");
        models[size_t(model)].transforms.assign(mats, mats + size_t(n) * 16);
        committed = false;
        version++, geom_version++;
    }

    int Scene::add_texture(const float *rgba, uint32_t w, uint32_t h)
    {
        if (!rgba || !w || !h) throw Error(ERR_INVALID_ARG, "add_texture: empty image");
        HostTexture t;
        t.w = w, t.h = h;
        t.rgba.assign(rgba, rgba + size_t(w) * h * 4);
        textures.push_back(std::move(t));
        committed = false;
        version++, geom_version++;
        return int(textures.size()) - 1;
    }

    void Scene::upload_materials()
    {
        std::vector<DMaterial> dm;
        has_alpha = false;
        for (const HostModel &m : models)
            for (const crb_material &s : m.materials)
            {
                DMaterial d {};
                memcpy(d.colour, s.colour, 16);
                d.shade_type = s.shade_type > 2 ? uint32_t(CRB_SMOOTH) : s.shade_type;
                d.ior = s.ior, d.reflectiveness = s.reflectiveness, d.emission = s.emission, d.tex = s.tex, d.roughness = s.roughness;
                dm.push_back(d);
                if (s.tex >= 0)
                {
                    const HostTexture &t = textures[size_t(s.tex)];
                    for (size_t i = 3; i < t.rgba.size(); i += 4)
                        if (t.rgba[i] == 0.0f)
                        {
                            has_alpha = true;
                            break;
                        }
                }
                else if (s.colour[3] == 0.0f)
                    has_alpha = true;
            }
        d_materials.alloc(dm.size() ? dm.size() : 1);
        dev_upload(d_materials.p, dm.data(), dm.size() * sizeof(DMaterial), stream);

        // light list of the extended shading mode: every emissive, non-cut-out triangle in world space, in
        // (model, instance, triangle) order; same arithmetic as k_flatten
        std::vector<float4> lt;
        for (const HostModel &m : models)
        {
            bool any = false;
            for (const crb_material &s : m.materials) any = any || (s.emission > 0.0f && s.colour[3] != 0.0f);
            if (!any) continue;
            for (size_t ii = 0; ii < m.transforms.size() / 16; ii++)
            {
                const float *M = &m.transforms[16 * ii];
                for (uint32_t t = 0; t < m.ntris; t++)
                {
                    const crb_material &s = m.materials[m.mat_idx[t]];
                    if (!(s.emission > 0.0f) || s.colour[3] == 0.0f) continue;
                    float w[9];
                    for (int k = 0; k < 3; k++)
                    {
                        const float x = m.verts[size_t(t) * 9 + 3 * k], y = m.verts[size_t(t) * 9 + 3 * k + 1], z = m.verts[size_t(t) * 9 + 3 * k + 2];
                        for (int r = 0; r < 3; r++) w[3 * k + r] = (M[0 + r] * x + M[4 + r] * y) + (M[8 + r] * z + M[12 + r]);
                    }
                    lt.push_back(make_float4(w[0], w[1], w[2], s.colour[0] * s.emission));
                    lt.push_back(make_float4(w[3] - w[0], w[4] - w[1], w[5] - w[2], s.colour[1] * s.emission));
                    lt.push_back(make_float4(w[6] - w[0], w[7] - w[1], w[8] - w[2], s.colour[2] * s.emission));
                }
            }
        }
        n_lights = uint32_t(lt.size() / 3);
        d_lights.alloc(lt.size() ? lt.size() : 1);
        dev_upload(d_lights.p, lt.data(), lt.size() * sizeof(float4), stream);
        stream_sync(stream);
    }

    void Scene::upload_skybox()
    {
        if (sky_w && sky_h)
        {
            d_skybox.alloc(size_t(sky_w) * sky_h);
            dev_upload(d_skybox.p, skybox.data(), size_t(sky_w) * sky_h * 16, stream);
            stream_sync(stream);
        }
        version++;
    }

    namespace
    {
        // glm::inverse(mat4) (model.cpp:109), glm 0.9.9.8 func_matrix.inl compute_inverse<4,4>, float, column-major m[c*4+r];
        // compiled with -ffp-contract=off: one rounding per operation, the same bits as the reference's per-ray inverse
        void glm_inverse(const float *M, float *out)
        {
            const auto m = [&](int c, int r) { return M[c * 4 + r]; };
            const float Coef00 = m(2, 2) * m(3, 3) - m(3, 2) * m(2, 3), Coef02 = m(1, 2) * m(3, 3) - m(3, 2) * m(1, 3), Coef03 = m(1, 2) * m(2, 3) - m(2, 2) * m(1, 3);
            const float Coef04 = m(2, 1) * m(3, 3) - m(3, 1) * m(2, 3), Coef06 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3), Coef07 = m(1, 1) * m(2, 3) - m(2, 1) * m(1, 3);
            const float Coef08 = m(2, 1) * m(3, 2) - m(3, 1) * m(2, 2), Coef10 = m(1, 1) * m(3, 2) - m(3, 1) * m(1, 2), Coef11 = m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2);
            const float Coef12 = m(2, 0) * m(3, 3) - m(3, 0) * m(2, 3), Coef14 = m(1, 0) * m(3, 3) - m(3, 0) * m(1, 3), Coef15 = m(1, 0) * m(2, 3) - m(2, 0) * m(1, 3);
            const float Coef16 = m(2, 0) * m(3, 2) - m(3, 0) * m(2, 2), Coef18 = m(1, 0) * m(3, 2) - m(3, 0) * m(1, 2), Coef19 = m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2);
            const float Coef20 = m(2, 0) * m(3, 1) - m(3, 0) * m(2, 1), Coef22 = m(1, 0) * m(3, 1) - m(3, 0) * m(1, 1), Coef23 = m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1);
            const float F[6][4] = { { Coef00, Coef00, Coef02, Coef03 }, { Coef04, Coef04, Coef06, Coef07 }, { Coef08, Coef08, Coef10, Coef11 },
                                    { Coef12, Coef12, Coef14, Coef15 }, { Coef16, Coef16, Coef18, Coef19 }, { Coef20, Coef20, Coef22, Coef23 } };
            const float V[4][4] = { { m(1, 0), m(0, 0), m(0, 0), m(0, 0) }, { m(1, 1), m(0, 1), m(0, 1), m(0, 1) }, { m(1, 2), m(0, 2), m(0, 2), m(0, 2) },
                                    { m(1, 3), m(0, 3), m(0, 3), m(0, 3) } };
            float       col[4][4];
            const float sa[4] = { +1, -1, +1, -1 }, sb[4] = { -1, +1, -1, +1 };
            for (int i = 0; i < 4; i++)
            {
                col[0][i] = ((V[1][i] * F[0][i] - V[2][i] * F[1][i]) + V[3][i] * F[2][i]) * sa[i];
                col[1][i] = ((V[0][i] * F[0][i] - V[2][i] * F[3][i]) + V[3][i] * F[4][i]) * sb[i];
                col[2][i] = ((V[0][i] * F[1][i] - V[1][i] * F[3][i]) + V[3][i] * F[5][i]) * sa[i];
                col[3][i] = ((V[0][i] * F[2][i] - V[1][i] * F[4][i]) + V[2][i] * F[5][i]) * sb[i];
            }
            const float d0 = m(0, 0) * col[0][0], d1 = m(0, 1) * col[1][0], d2 = m(0, 2) * col[2][0], d3 = m(0, 3) * col[3][0];
            const float inv_det = 1.0f / ((d0 + d1) + (d2 + d3));
            for (int c = 0; c < 4; c++)
                for (int r = 0; r < 4; r++) out[c * 4 + r] = col[c][r] * inv_det;
        }
    }    // namespace

    // Two-level commit: one object-space BLAS per model (rebuilt only when the geometry set changed), a TLAS over the
    // (model, instance) pairs (rebuilt on every commit: an instance edit costs a 10..100-instance build, not a rebuild of
    // 18M flattened triangles), and the instance table with glm's inverse of every transform.
    void Scene::commit_two_level()
    {
        auto t0 = std::chrono::steady_clock::now();
        const int B = 256;
        size_t n_src = 0, n_flat64 = 0;
        bool   any_uv = false;
        for (const HostModel &m : models) n_src += m.ntris, n_flat64 += size_t(m.ntris) * (m.transforms.size() / 16), any_uv = any_uv || !m.uvs.empty();
        if (n_flat64 > 0x7ffffff0ull) throw Error(ERR_BUILD_INDEX, "scene exceeds 2^31 instanced triangles");
        n_flat = uint32_t(n_flat64);
        std::vector<uint64_t> ids;
        for (const HostModel &m : models) ids.push_back(m.geom_id);
        BuildOptions opt;
        if (const char *e = getenv("CRB_TREELET_PASSES")) opt.treelet_passes = atoi(e);
        if (const char *e = getenv("CRB_OPTIMAL_COLLAPSE")) opt.optimal_collapse = atoi(e) != 0;
        if (const char *e = getenv("CRB_COST_PRIM")) opt.cost_prim = float(atof(e));
        double upload = 0;

        bool rebuilt = false;
        if (ids != blas_geom_ids || !d_blas_nodes.p)
        {
            rebuilt = true;
            // ---- geometry changed: upload object-space data, per-triangle shading records, one BLAS per model
            DBuf<float>    d_obj_verts;
            DBuf<uint32_t> d_mat_idx;
            d_obj_verts.alloc(n_src * 9 + 1);
            d_mat_idx.alloc(n_src + 1);
            d_shade_tri.alloc(n_src + 1);
            if (any_uv) d_obj_uvs.alloc(n_src * 6); else d_obj_uvs.release();
            d_flat_src.release();
            size_t   src_off = 0;
            uint32_t mat_base = 0;
            for (const HostModel &m : models)
            {
                dev_upload(d_obj_verts.p + src_off * 9, m.verts.data(), size_t(m.ntris) * 9 * 4, stream);
                dev_upload(d_mat_idx.p + src_off, m.mat_idx.data(), size_t(m.ntris) * 4, stream);
                if (any_uv)
                {
                    if (!m.uvs.empty())
                        dev_upload(d_obj_uvs.p + src_off * 6, m.uvs.data(), size_t(m.ntris) * 6 * 4, stream);
                    else
                        dev_zero(d_obj_uvs.p + src_off * 6, size_t(m.ntris) * 6 * 4, stream);
                }
                if (m.ntris) CRB_LAUNCH(k_shade_tri, (m.ntris + B - 1) / B, B, stream, d_obj_verts.p + src_off * 9, d_mat_idx.p + src_off, mat_base, m.ntris, d_shade_tri.p + src_off);
                src_off += m.ntris;
                mat_base += uint32_t(m.materials.size());
            }
            stream_sync(stream);
            upload = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            std::vector<DBuf<uint4>>  bn(models.size());
            std::vector<DBuf<float4>> bt(models.size());
            blas_table.assign(models.size(), Blas {});
            blas_build_ms = 0, blas_nodes_total = 0, blas_depth = 0;
            size_t   node_total = 0, tri_total = 0;
            float    sah = 0;
            src_off = 0;
            for (size_t mi = 0; mi < models.size(); mi++)
            {
                BuildStats st;
                build_bvh8(d_obj_verts.p + src_off * 9, models[mi].ntris, stream, opt, bn[mi], bt[mi], st);
                blas_table[mi] = Blas { uint32_t(node_total), uint32_t(tri_total), st.n_nodes, st.n_tris };
                node_total += st.n_nodes, tri_total += st.n_tris;
                blas_build_ms += st.build_ms, blas_depth = std::max(blas_depth, st.max_depth), sah += st.sah_cost;
                src_off += models[mi].ntris;
            }
            d_blas_nodes.alloc(node_total * BVH8_NODE_U4 + BVH8_NODE_U4);
            d_blas_tris.alloc(tri_total * BVH8_TRI_F4 + BVH8_TRI_F4);
            for (size_t mi = 0; mi < models.size(); mi++)
            {
                dev_copy(d_blas_nodes.p + size_t(blas_table[mi].node_base) * BVH8_NODE_U4, bn[mi].p, size_t(blas_table[mi].n_nodes) * BVH8_NODE_U4 * 16, stream);
                dev_copy(d_blas_tris.p + size_t(blas_table[mi].tri_base) * BVH8_TRI_F4, bt[mi].p, size_t(blas_table[mi].n_tris) * BVH8_TRI_F4 * 16, stream);
            }
            d_blas.alloc(models.size() ? models.size() : 1);
            dev_upload(d_blas.p, blas_table.data(), blas_table.size() * sizeof(Blas), stream);
            stream_sync(stream);
            blas_nodes_total = uint32_t(node_total);
            build.sah_cost   = sah;
            blas_geom_ids    = ids;
        }
        auto t1 = std::chrono::steady_clock::now();

        // ---- instances: table, flat ranges, bounds, TLAS
        ranges.clear();
        std::vector<Instance> inst;
        std::vector<float>    proxy;    // one "triangle" per instance whose box is the instance's bounds
        size_t   src_off = 0, flat_off = 0;
        for (size_t mi = 0; mi < models.size(); mi++)
        {
            const HostModel &m = models[mi];
            // object-space bounds of the model (cached per geometry would save this O(T) host loop on instance edits; 2M: ~5 ms)
            float lo[3] = { 3e38f, 3e38f, 3e38f }, hi[3] = { -3e38f, -3e38f, -3e38f };
            for (size_t i = 0; i < size_t(m.ntris) * 3; i++)
                for (int a = 0; a < 3; a++) lo[a] = std::min(lo[a], m.verts[i * 3 + a]), hi[a] = std::max(hi[a], m.verts[i * 3 + a]);
            const size_t ni = m.transforms.size() / 16;
            for (size_t ii = 0; ii < ni && m.ntris; ii++)
            {
                const float *T = &m.transforms[16 * ii];
                float        inv[16];
                glm_inverse(T, inv);
                Instance I {};
                for (int c = 0; c < 4; c++)
                    for (int r = 0; r < 3; r++) I.inv[3 * c + r] = inv[4 * c + r], I.fwd[3 * c + r] = T[4 * c + r];
                float wlo[3] = { 3e38f, 3e38f, 3e38f }, whi[3] = { -3e38f, -3e38f, -3e38f };
                for (int corner = 0; corner < 8; corner++)
                {
                    const double p[3] = { (corner & 1) ? hi[0] : lo[0], (corner & 2) ? hi[1] : lo[1], (corner & 4) ? hi[2] : lo[2] };
                    for (int r = 0; r < 3; r++)
                    {
                        const double w = double(T[r]) * p[0] + double(T[4 + r]) * p[1] + double(T[8 + r]) * p[2] + double(T[12 + r]);
                        wlo[r] = std::min(wlo[r], float(w)), whi[r] = std::max(whi[r], float(w));
                    }
                }
                for (int r = 0; r < 3; r++)
                {
                    // the device maps hit points back in float: keep the box conservative by a few ulps of its size / position
                    const float pad = 1e-5f * std::max(std::max(std::fabs(wlo[r]), std::fabs(whi[r])), whi[r] - wlo[r]) + 1e-30f;
                    I.lo[r] = wlo[r] - pad, I.hi[r] = whi[r] + pad;
                }
                I.blas = uint32_t(mi), I.flat_start = uint32_t(flat_off);
                I.node_base = blas_table[mi].node_base, I.tri_base = blas_table[mi].tri_base, I.n_nodes = blas_table[mi].n_nodes;
                inst.push_back(I);
                proxy.insert(proxy.end(), { I.lo[0], I.lo[1], I.lo[2], I.hi[0], I.hi[1], I.hi[2], I.lo[0], I.hi[1], I.lo[2] });
                FlatRange r {};
                r.start = uint32_t(flat_off), r.ntris = m.ntris, r.model = uint32_t(mi), r.inst = uint32_t(ii), r.src_start = uint32_t(src_off);
                ranges.push_back(r);
                flat_off += m.ntris;
            }
            src_off += m.ntris;
        }
        n_flat = uint32_t(flat_off);
        d_ranges.alloc(ranges.size() ? ranges.size() : 1);
        dev_upload(d_ranges.p, ranges.data(), ranges.size() * sizeof(FlatRange), stream);
        d_inst.alloc(inst.size() ? inst.size() : 1);
        dev_upload(d_inst.p, inst.data(), inst.size() * sizeof(Instance), stream);
        DBuf<float> d_proxy;
        d_proxy.alloc(proxy.size() + 9);
        dev_upload(d_proxy.p, proxy.data(), proxy.size() * 4, stream);
        // ---- textures, materials, skybox (as in the flat path)
        {
            std::vector<DTexture> dt;
            size_t                total = 0;
            for (const HostTexture &t : textures)
            {
                dt.push_back(DTexture { t.w, t.h, uint32_t(total), 0 });
                total += size_t(t.w) * t.h;
            }
            d_textures.alloc(dt.size() ? dt.size() : 1);
            d_texels.alloc(total ? total : 1);
            dev_upload(d_textures.p, dt.data(), dt.size() * sizeof(DTexture), stream);
            for (size_t i = 0; i < textures.size(); i++) dev_upload(d_texels.p + dt[i].offset, textures[i].rgba.data(), textures[i].rgba.size() * 4, stream);
        }
        upload_materials();
        if (sky_w && sky_h && !d_skybox.p) upload_skybox();
        stream_sync(stream);
        upload_ms = upload + std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
        BuildStats tl;
        build_bvh8(d_proxy.p, uint32_t(inst.size()), stream, opt, d_nodes, d_tris, tl);
        tlas_build_ms   = tl.build_ms;
        build.build_ms  = tl.build_ms + (rebuilt ? blas_build_ms : 0.0);    // an instance edit pays for the TLAS only
        stored_tris     = 0;
        for (const Blas &b : blas_table) stored_tris += b.n_tris;
        build.n_nodes   = tl.n_nodes + blas_nodes_total;
        build.n_tris    = n_flat;
        build.max_depth = tl.max_depth + blas_depth;
        tlas_nodes      = tl.n_nodes;
        n_instances     = uint32_t(inst.size());
        two_level       = true;
        committed       = true;
        version++;
    }

    void Scene::commit()
    {
        {
            // instanced scenes trace through a TLAS with the reference's per-instance arithmetic (bit-identical hits);
            // flattening into world space stays as the fast path behind crb_scene_set_option / CRB_FLATTEN=1
            bool instanced = false;
            for (const HostModel &m : models)
                instanced = instanced || m.transforms.size() != 16 || !is_identity16(m.transforms.data());
            static const int flatten_env = getenv("CRB_FLATTEN") ? atoi(getenv("CRB_FLATTEN")) : -1;
            const bool       flatten     = flatten_env >= 0 ? flatten_env != 0 : flatten_instances;
            if (instanced && !flatten)
            {
                commit_two_level();
                return;
            }
            two_level = false;
            blas_geom_ids.clear();
            d_blas_nodes.release(), d_blas_tris.release();
        }
        auto t0 = std::chrono::steady_clock::now();
        // ---- sizes
        size_t n_src = 0, n_flat64 = 0;
        bool   all_identity = true, any_uv = false;
        for (const HostModel &m : models)
        {
            n_src += m.ntris;
            const size_t ni = m.transforms.size() / 16;
            n_flat64 += size_t(m.ntris) * ni;
            for (size_t i = 0; i < ni; i++) all_identity = all_identity && is_identity16(&m.transforms[16 * i]);
            if (ni != 1) all_identity = false;
            any_uv = any_uv || !m.uvs.empty();
        }
        if (n_flat64 > 0x7ffffff0ull) throw Error(ERR_BUILD_INDEX, "flattened scene exceeds 2^31 triangles");
        n_flat = uint32_t(n_flat64);

        // ---- upload object-space data
        DBuf<float>    d_obj_verts;
        DBuf<uint32_t> d_mat_idx;
        DBuf<float>    d_xf;
        d_obj_verts.alloc(n_src * 9 + 1);
        d_mat_idx.alloc(n_src + 1);
        d_shade_tri.alloc(n_src + 1);
        d_wverts.alloc(size_t(n_flat) * 9 + 1);
        if (any_uv) d_obj_uvs.alloc(n_src * 6); else d_obj_uvs.release();
        if (!all_identity) d_flat_src.alloc(n_flat + 1); else d_flat_src.release();
        size_t n_xf = 0;
        for (const HostModel &m : models) n_xf += m.transforms.size();
        d_xf.alloc(n_xf + 16);

        ranges.clear();
        size_t   src_off = 0, flat_off = 0, xf_off = 0;
        uint32_t mat_base = 0;
        const int B = 256;
        for (size_t mi = 0; mi < models.size(); mi++)
        {
            const HostModel &m = models[mi];
            dev_upload(d_obj_verts.p + src_off * 9, m.verts.data(), size_t(m.ntris) * 9 * 4, stream);
            dev_upload(d_mat_idx.p + src_off, m.mat_idx.data(), size_t(m.ntris) * 4, stream);
            if (any_uv)
            {
                if (!m.uvs.empty())
                    dev_upload(d_obj_uvs.p + src_off * 6, m.uvs.data(), size_t(m.ntris) * 6 * 4, stream);
                else
                    dev_zero(d_obj_uvs.p + src_off * 6, size_t(m.ntris) * 6 * 4, stream);
            }
            dev_upload(d_xf.p + xf_off, m.transforms.data(), m.transforms.size() * 4, stream);
            if (m.ntris)
            {
                const unsigned g = (m.ntris + B - 1) / B;
                CRB_LAUNCH(k_shade_tri, g, B, stream, d_obj_verts.p + src_off * 9, d_mat_idx.p + src_off, mat_base, m.ntris, d_shade_tri.p + src_off);
                const size_t ni = m.transforms.size() / 16;
                for (size_t ii = 0; ii < ni; ii++)
                {
                    const int ident = is_identity16(&m.transforms[16 * ii]) ? 1 : 0;
                    CRB_LAUNCH(k_flatten, g, B, stream, d_obj_verts.p + src_off * 9, m.ntris, d_xf.p + xf_off + 16 * ii, ident, d_wverts.p + flat_off * 9,
                               all_identity ? (uint32_t *) nullptr : d_flat_src.p + flat_off, uint32_t(src_off));
                    FlatRange r {};
                    r.start = uint32_t(flat_off), r.ntris = m.ntris, r.model = uint32_t(mi), r.inst = uint32_t(ii), r.src_start = uint32_t(src_off);
                    ranges.push_back(r);
                    flat_off += m.ntris;
                }
            }
            src_off += m.ntris;
            xf_off += m.transforms.size();
            mat_base += uint32_t(m.materials.size());
        }
        d_ranges.alloc(ranges.size() ? ranges.size() : 1);
        dev_upload(d_ranges.p, ranges.data(), ranges.size() * sizeof(FlatRange), stream);

        // ---- textures
        {
            std::vector<DTexture> dt;
            size_t                total = 0;
            for (const HostTexture &t : textures)
            {
                dt.push_back(DTexture { t.w, t.h, uint32_t(total), 0 });
                total += size_t(t.w) * t.h;
            }
            d_textures.alloc(dt.size() ? dt.size() : 1);
            d_texels.alloc(total ? total : 1);
            dev_upload(d_textures.p, dt.data(), dt.size() * sizeof(DTexture), stream);
            for (size_t i = 0; i < textures.size(); i++) dev_upload(d_texels.p + dt[i].offset, textures[i].rgba.data(), textures[i].rgba.size() * 4, stream);
        }
        upload_materials();
        if (sky_w && sky_h && !d_skybox.p) upload_skybox();
        stream_sync(stream);
        upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

        // ---- BVH
        BuildOptions opt;
        if (const char *e = getenv("CRB_TREELET_PASSES")) opt.treelet_passes = atoi(e);    // build-quality experiments
        if (const char *e = getenv("CRB_OPTIMAL_COLLAPSE")) opt.optimal_collapse = atoi(e) != 0;
        if (const char *e = getenv("CRB_COST_PRIM")) opt.cost_prim = float(atof(e));
        build_bvh8(d_wverts.p, n_flat, stream, opt, d_nodes, d_tris, build);
        stored_tris = build.n_tris;
        d_wverts.release();
        committed = true;
        version++;
    }

    void Scene::copy_description_from(const Scene &src)
    {
        models = src.models, textures = src.textures;
        skybox = src.skybox, sky_w = src.sky_w, sky_h = src.sky_h;
        d_skybox.release();    // commit() uploads it again
        src_sky_version = src.sky_version;
        copy_light_state_from(src);
        committed = false;
        version++, geom_version++;
    }

    void Scene::copy_light_state_from(const Scene &src)
    {
        sun = src.sun, sun_enabled = src.sun_enabled, camera = src.camera, flatten_instances = src.flatten_instances;
        sky_rot[0] = src.sky_rot[0], sky_rot[1] = src.sky_rot[1];
        if (src_sky_version != src.sky_version)
        {
            skybox = src.skybox, sky_w = src.sky_w, sky_h = src.sky_h;
            src_sky_version = src.sky_version;
            if (committed) upload_skybox();
        }
        bool mats_differ = models.size() != src.models.size();
        for (size_t i = 0; i < models.size() && !mats_differ; i++)
            mats_differ = models[i].materials.size() != src.models[i].materials.size() ||
                          memcmp(models[i].materials.data(), src.models[i].materials.data(), models[i].materials.size() * sizeof(crb_material)) != 0;
        if (mats_differ && models.size() == src.models.size())
        {
            for (size_t i = 0; i < models.size(); i++) models[i].materials = src.models[i].materials;
            if (committed) upload_materials();
        }
        version++;
    }

    void Scene::require_committed() const
    {
        if (!committed) throw Error(ERR_NOT_COMMITTED, "scene not committed (call crb_scene_commit after adding meshes/instances)");
    }

    DScene Scene::device_scene(uint32_t w, uint32_t h) const
    {
        DScene d {};
        d.bvh.nodes = d_nodes.p, d.bvh.tris = d_tris.p, d.bvh.n_nodes = build.n_nodes, d.bvh.n_tris = build.n_tris;
        d.shade_tri = d_shade_tri.p, d.obj_uvs = d_obj_uvs.p, d.flat_src = d_flat_src.p;
        d.materials = d_materials.p, d.textures = d_textures.p, d.texels = d_texels.p;
        d.skybox = (sky_w && sky_h) ? d_skybox.p : nullptr, d.sky_w = sky_w, d.sky_h = sky_h;
        d.sky_rot[0] = sky_rot[0], d.sky_rot[1] = sky_rot[1];
        d.ranges = d_ranges.p, d.n_ranges = uint32_t(ranges.size());
        d.has_alpha = has_alpha ? 1u : 0u;
        d.lights = d_lights.p, d.n_lights = n_lights;
        d.two_level = two_level ? 1u : 0u;
        if (two_level)
        {
            d.bvh2.tlas.nodes = d_nodes.p, d.bvh2.tlas.tris = d_tris.p, d.bvh2.tlas.n_nodes = tlas_nodes, d.bvh2.tlas.n_tris = n_instances;
            d.bvh2.nodes = d_blas_nodes.p, d.bvh2.tris = d_blas_tris.p, d.bvh2.blas = d_blas.p, d.bvh2.inst = d_inst.p, d.bvh2.n_inst = n_instances;
        }

        // ---- sun: registry.cpp:248-256 + sampling.h:21-47 (host libm, same as the reference's CPU)
        DSun &s = d.sun;
        memcpy(s.dir, sun.direction, 12), memcpy(s.colour, sun.colour, 12);
        s.size = sun.size, s.intensity = sun.intensity, s.enabled = sun_enabled ? 1u : 0u;
        {
            const float nx = -sun.direction[0], ny = -sun.direction[1], nz = -sun.direction[2];
            const float sg = (nz < 0.0) ? -1.0f : 1.0f;
            const float a  = -1.0f / (sg + nz);
            const float b  = nx * ny * a;
            const float tangent[3]   = { 1.0f + sg * nx * nx * a, sg * b, -sg * nx };
            const float bitangent[3] = { b, sg + ny * ny * a, -ny };
            s.transform[0] = tangent[0], s.transform[1] = tangent[1], s.transform[2] = tangent[2];
            s.transform[3] = nx, s.transform[4] = ny, s.transform[5] = nz;
            s.transform[6] = bitangent[0], s.transform[7] = bitangent[1], s.transform[8] = bitangent[2];
        }
        s.one_minus_cos = 1.0f - std::cos(sun.size);
        s.pdf           = 1.0f / (TAU_F * (1.0f - std::cos(sun.size)));

        // ---- camera: camera.cpp:54-68 (T * Ry(rot.x) * Rx(rot.y) * Rz(rot.z), degrees) and :22
        DCamera &c = d.cam;
        memcpy(c.position, camera.position, 12), memcpy(c.trans, camera.position, 12);
        float m[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
        const float rad = 0.01745329251994329576923690768489f;
        const float up[3] = { 0, 1, 0 }, right[3] = { 1, 0, 0 }, fwd[3] = { 0, 0, 1 };
        rotate3(m, camera.rotation[0] * rad, up);
        rotate3(m, camera.rotation[1] * rad, right);
        rotate3(m, camera.rotation[2] * rad, fwd);
        memcpy(c.m, m, sizeof(m));
        c.w      = 1.0f / std::tan(0.5f * (camera.fov * rad));
        c.scale  = camera.scale;
        c.mode   = camera.mode;
        c.aspect = float(w) / float(h);    // renderer.cpp:199
        return d;
    }
}    // namespace crb
