// post.cu — export-time post chain as CUDA kernels. Replaces the OpenGL compute
// path of cr::post_processor::process (src/render/post/post_processor.cpp:110-296) and its shaders
// assets/app/shaders/post_process.comp / blur.comp, so that export needs no GL context:
//   _brightness  (post_processor.cpp:298-311)  zero pixels darker than the bloom threshold
//   _blur        (post_processor.cpp:237-296, blur.comp:11-44)  10 ping-pong passes, 15-tap Gaussian,
//                horizontal on even passes, vertical on odd ones, texelFetch outside the image = 0
//   compose      (post_process.comp:119-136)  bloom add, gray scale, one of 4 tonemappers
// Reference quirks that are reproduced because they change the pixels:
//   * every dispatch covers ceil(int(w/8)) x ceil(int(h/8)) groups of 8x8 (integer division,
//     post_processor.cpp:220-223,282-285): pixels beyond 8*(w/8) / 8*(h/8) stay at the cleared value 0;
//   * _brightness addresses pixel i as (i % width, i / HEIGHT) (post_processor.cpp:304);
//   * the gray-scale weights are 0.2126, 0.7162, 0.0722 (sic, post_process.comp:114).
#include "../../include/crender_b200.h"
#include "platform.cuh"
#include "scene.cuh"

namespace crb
{
    namespace
    {
        __constant__ float c_blur_w[8] = { 0.19744746769063704f, 0.1746973469158936f,  0.12099884565428047f,  0.06560233156931679f,
                                           0.027839605612666265f, 0.009246250740395456f, 0.002403157286908872f, 0.00048872837522002f };
#ifdef CRB_EMU
        const float *blur_w() { return c_blur_w; }
#else
        __device__ __forceinline__ const float *blur_w() { return c_blur_w; }
#endif

        __global__ void k_brightness(const float4 *__restrict__ src, float4 *__restrict__ dst, int w, int h, float threshold)
        {
            const int i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= w * h) return;
            dst[i] = src[i];
        }
        // second pass because the reference reads and writes pixel (i % w, i / h), not pixel i
        __global__ void k_brightness_apply(const float4 *__restrict__ src, float4 *__restrict__ dst, int w, int h, float threshold)
        {
            const int i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= w * h) return;
            const int x = i % w, y = i / h;
            if (y >= h) return;    // (only reachable when w > h; the reference would read out of bounds there)
            const float4 at = src[x + y * w];
            const float  b  = (at.x * 0.2126f + at.y * 0.7152f) + at.z * 0.0722f;
            if (b < threshold) dst[x + y * w] = make_float4(0.f, 0.f, 0.f, 1.f);
        }

        __device__ __forceinline__ float4 fetch0(const float4 *img, int w, int h, int x, int y)
        {
            return (x >= 0 && y >= 0 && x < w && y < h) ? img[x + y * w] : make_float4(0.f, 0.f, 0.f, 0.f);
        }

        __global__ void k_blur(const float4 *__restrict__ src, float4 *__restrict__ dst, int w, int h, int horizontal)
        {
            const int x = blockIdx.x * 8 + (threadIdx.x & 7), y = blockIdx.y * 8 + (threadIdx.x >> 3);
            if (!(x < w && y < h)) return;
            const float *wt = blur_w();
            float4       c  = src[x + y * w];
            float        r = c.x * wt[0], g = c.y * wt[0], b = c.z * wt[0];
            for (int i = 1; i < 8; ++i)
            {
                const float4 p = horizontal ? fetch0(src, w, h, x + i, y) : fetch0(src, w, h, x, y + i);
                r += p.x * wt[i], g += p.y * wt[i], b += p.z * wt[i];
                const float4 q = horizontal ? fetch0(src, w, h, x - i, y) : fetch0(src, w, h, x, y - i);
                r += q.x * wt[i], g += q.y * wt[i], b += q.z * wt[i];
            }
            dst[x + y * w] = make_float4(r, g, b, 1.0f);
        }

        __device__ __forceinline__ float um(float x)
        {
            const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
            return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F;
        }

        __global__ void k_compose(const float4 *__restrict__ src, const float4 *__restrict__ bloom, float4 *__restrict__ dst, int w, int h,
                                  crb_post_settings s)
        {
            const int x = blockIdx.x * 8 + (threadIdx.x & 7), y = blockIdx.y * 8 + (threadIdx.x >> 3);
            if (!(x < w && y < h)) return;
            const float4 c = src[x + y * w];
            float        r = c.x, g = c.y, b = c.z;
            if (s.use_bloom)
            {
                const float4 bl = bloom[x + y * w];
                r += bl.x * s.bloom_strength, g += bl.y * s.bloom_strength, b += bl.z * s.bloom_strength;
            }
            if (s.use_gray_scale)
            {
                const float gs = (0.2126f * r + 0.7162f * g) + 0.0722f * b;
                r = g = b = gs;
            }
            if (s.use_tonemapping)
            {
                const float e = s.tonemapping_exposure, ig = 1.0f / s.gamma_correction;
                switch (s.tonemapping_type)
                {
                case 0: r = powf(r * e, ig), g = powf(g * e, ig), b = powf(b * e, ig); break;
                case 1:
                    r *= e, g *= e, b *= e;
                    r = powf(r / (r + 1.0f), ig), g = powf(g / (g + 1.0f), ig), b = powf(b / (b + 1.0f), ig);
                    break;
                case 2:
                {
                    const float xr = fmaxf(0.0f, r * e - 0.004f), xg = fmaxf(0.0f, g * e - 0.004f), xb = fmaxf(0.0f, b * e - 0.004f);
                    r = (xr * (6.2f * xr + 0.5f)) / (xr * (6.2f * xr + 1.7f) + 0.06f);
                    g = (xg * (6.2f * xg + 0.5f)) / (xg * (6.2f * xg + 1.7f) + 0.06f);
                    b = (xb * (6.2f * xb + 0.5f)) / (xb * (6.2f * xb + 1.7f) + 0.06f);
                    break;
                }
                case 3:
                {
                    const float ws = 1.0f / um(11.2f);
                    r = powf(um(2.0f * (r * e)) * ws, ig), g = powf(um(2.0f * (g * e)) * ws, ig), b = powf(um(2.0f * (b * e)) * ws, ig);
                    break;
                }
                default: break;
                }
            }
            dst[x + y * w] = make_float4(r, g, b, 1.0f);
        }
    }    // namespace

    // d_src: device RGBA image (w*h float4). out_host receives the processed image. Mirrors
    // post_processor::process, including "no effect enabled -> the image is returned unchanged".
    void post_process(Scene &sc, const float4 *d_src, uint32_t w, uint32_t h, const crb_post_settings &s, float *out_host)
    {
        cudaStream_t st = sc.stream;
        const size_t n  = size_t(w) * h;
        if (!s.use_bloom && !s.use_gray_scale && !s.use_tonemapping)
        {
            dev_download(out_host, d_src, n * 16, st);
            return;
        }
        // groups exactly as glDispatchCompute(ceil(w / 8), ceil(h / 8)) with integer division
        const unsigned gx = w / 8, gy = h / 8;
        DBuf<float4>   a, b, out;
        out.alloc(n);
        dev_zero(out.p, n * 16, st);    // glClearTexImage(target, ..., nullptr)
        const float4 *bloom = nullptr;
        if (s.use_bloom)
        {
            a.alloc(n), b.alloc(n);
            const unsigned g1 = unsigned((n + 255) / 256);
            CRB_LAUNCH(k_brightness, g1, 256, st, d_src, a.p, int(w), int(h), s.bloom_threshold);
            CRB_LAUNCH(k_brightness_apply, g1, 256, st, d_src, a.p, int(w), int(h), s.bloom_threshold);
            float4 *cur = a.p, *nxt = b.p;
            for (int i = 0; i < 10; i++)
            {
                dev_zero(nxt, n * 16, st);
                if (gx && gy)
                {
#ifdef CRB_EMU
                    for (unsigned by = 0; by < gy; by++)
                    {
                        // the emulated launcher is 1-D: run each row of groups as its own launch
                        crb_emu::blockIdx_.y = by;
                        CRB_LAUNCH(k_blur, gx, 64, st, cur, nxt, int(w), int(h), int(i % 2 == 0));
                    }
                    crb_emu::blockIdx_.y = 0;
#else
                    k_blur<<<dim3(gx, gy), 64, 0, st>>>(cur, nxt, int(w), int(h), int(i % 2 == 0));
                    CRB_CUDA_CHECK(cudaGetLastError());
#endif
                }
                float4 *t = cur;
                cur = nxt, nxt = t;
            }
            bloom = cur;
        }
        if (gx && gy)
        {
#ifdef CRB_EMU
            for (unsigned by = 0; by < gy; by++)
            {
                crb_emu::blockIdx_.y = by;
                CRB_LAUNCH(k_compose, gx, 64, st, d_src, bloom, out.p, int(w), int(h), s);
            }
            crb_emu::blockIdx_.y = 0;
#else
            k_compose<<<dim3(gx, gy), 64, 0, st>>>(d_src, bloom, out.p, int(w), int(h), s);
            CRB_CUDA_CHECK(cudaGetLastError());
#endif
        }
        dev_download(out_host, out.p, n * 16, st);
    }
}    // namespace crb
