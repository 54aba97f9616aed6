// vecmath.cuh — 3-vector helpers for device code.
//
// Every operator here is ONE IEEE-754 binary32 rounding per written operation (the __f*_rn intrinsics
// are never contracted into FMAs), in the same operation order as glm 0.9.9.8 evaluates them, which is
// what the reference computes with (e.g. src/render/renderer.cpp:51,60-67; src/objects/model.cpp:35).
// That keeps ray generation, the triangle test and shading bit-comparable with the CPU oracle; only
// libm-vs-libdevice transcendentals (sinf, cosf, acosf, ...) can differ, by ulps.
#pragma once
#include "platform.cuh"

namespace crb
{
    struct V3
    {
        float x, y, z;
    };
    __device__ __forceinline__ V3 v3(float x, float y, float z) { return V3 { x, y, z }; }
    __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
    __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
    __device__ __forceinline__ V3 operator*(V3 a, V3 b) { return v3(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z)); }
    __device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
    __device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(__fmul_rn(s, a.x), __fmul_rn(s, a.y), __fmul_rn(s, a.z)); }
    __device__ __forceinline__ V3 operator/(V3 a, float s) { return v3(__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s)); }
    __device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
    // glm::dot(vec3): tmp = a*b; (tmp.x + tmp.y) + tmp.z
    __device__ __forceinline__ float dot(V3 a, V3 b)
    {
        return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
    }
    // glm::cross
    __device__ __forceinline__ V3 cross(V3 a, V3 b)
    {
        return v3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(b.y, a.z)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(b.z, a.x)),
                  __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(b.x, a.y)));
    }
    __device__ __forceinline__ float length(V3 a) { return __fsqrt_rn(dot(a, a)); }
    // glm::normalize = v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x)
    __device__ __forceinline__ V3 normalize(V3 a) { return a * __frcp_rn(__fsqrt_rn(dot(a, a))); }
    // glm::reflect = I - N * dot(N, I) * 2
    __device__ __forceinline__ V3    reflect(V3 I, V3 N) { return I - N * dot(N, I) * 2.0f; }
    // glm::clamp = min(max(x,lo),hi) with glm::max(a,b) = (a<b)?b:a and glm::min(a,b) = (b<a)?b:a (NaN stays NaN)
    __device__ __forceinline__ float clampf(float x, float lo, float hi)
    {
        const float m = (x < lo) ? lo : x;
        return (hi < m) ? hi : m;
    }
    __device__ __forceinline__ V3    ld3(const float *p) { return v3(p[0], p[1], p[2]); }
}    // namespace crb
