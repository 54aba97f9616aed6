// cuda_emu.h — CUDA built-ins for the kernel-logic test harness (tests/emu). TEST TOOL ONLY.
//
// The product's .cu sources are compiled by g++ with -DCRB_EMU=1 and this header on the include path: every "thread" of a
// launch is executed serially with a warp width of 1, so that per-thread kernel logic (hierarchy emission, refit,
// wide-node collapse, traversal, shading, the multi-GPU orchestration) can be checked against the oracle in a container
// without a GPU. It is never shipped, never loaded by the crender_b200 package and never a fallback: the product is
// built by nvcc for sm_100a only and fails loudly without a CUDA device. Block-cooperative code (shared memory,
// __syncthreads) is not emulated; such kernels have an explicitly serial body under CRB_EMU or are excluded.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

// ------------------------------------------------------------------ emulation (tests only)
#define CRB_WARP 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__

struct float2
{
    float x, y;
};
struct float3
{
    float x, y, z;
};
struct alignas(16) float4
{
    float x, y, z, w;
};
struct uint2
{
    unsigned x, y;
};
struct alignas(16) uint4
{
    unsigned x, y, z, w;
};
struct int2
{
    int x, y;
};
struct dim3
{
    unsigned x = 1, y = 1, z = 1;
};
inline float2 make_float2(float x, float y) { return float2 { x, y }; }
inline float3 make_float3(float x, float y, float z) { return float3 { x, y, z }; }
inline float4 make_float4(float x, float y, float z, float w) { return float4 { x, y, z, w }; }
inline uint2  make_uint2(unsigned x, unsigned y) { return uint2 { x, y }; }
inline uint4  make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4 { x, y, z, w }; }

namespace crb_emu
{
    extern dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;
}
#define threadIdx crb_emu::threadIdx_
#define blockIdx crb_emu::blockIdx_
#define blockDim crb_emu::blockDim_
#define gridDim crb_emu::gridDim_
typedef void *cudaStream_t;

template<typename F>
inline void crb_emu_launch(unsigned grid, unsigned block, F f)
{
    gridDim.x = grid, blockDim.x = block;
    for (unsigned b = 0; b < grid; b++)
        for (unsigned t = 0; t < block; t++)
        {
            blockIdx.x = b, threadIdx.x = t;
            f();
        }
}
#define CRB_LAUNCH(kernel, grid, block, stream, ...) crb_emu_launch((unsigned) (grid), (unsigned) (block), [&] { kernel(__VA_ARGS__); })

template<typename T>
inline T atomicAdd(T *p, T v)
{
    T o = *p;
    *p  = o + v;
    return o;
}
template<typename T>
inline T atomicMin(T *p, T v)
{
    T o = *p;
    if (v < o) *p = v;
    return o;
}
template<typename T>
inline T atomicMax(T *p, T v)
{
    T o = *p;
    if (v > o) *p = v;
    return o;
}
template<typename T>
inline T atomicOr(T *p, T v)
{
    T o = *p;
    *p  = o | v;
    return o;
}
template<typename T>
inline T atomicExch(T *p, T v)
{
    T o = *p;
    *p  = v;
    return o;
}
inline unsigned __float_as_uint(float f)
{
    unsigned u;
    memcpy(&u, &f, 4);
    return u;
}
inline float __uint_as_float(unsigned u)
{
    float f;
    memcpy(&f, &u, 4);
    return f;
}
inline int   __float_as_int(float f) { return (int) __float_as_uint(f); }
inline float __int_as_float(int i) { return __uint_as_float((unsigned) i); }
// the emu build is compiled with -ffp-contract=off, so plain operators are single roundings
inline float    __fmul_rn(float a, float b) { return a * b; }
inline float    __fadd_rn(float a, float b) { return a + b; }
inline float    __fsub_rn(float a, float b) { return a - b; }
inline float    __fdiv_rn(float a, float b) { return a / b; }
inline float    __frcp_rn(float a) { return 1.0f / a; }
inline float    __fsqrt_rn(float a) { return sqrtf(a); }
inline float    __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline int      __popc(unsigned x) { return __builtin_popcount(x); }
inline int      __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned) x); }
inline int      __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long) x); }
inline int      __ffs(int x) { return __builtin_ffs(x); }
inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
inline unsigned __activemask() { return 1u; }
template<typename T>
inline T __shfl_sync(unsigned, T v, int)
{
    return v;
}
template<typename T>
inline T __shfl_xor_sync(unsigned, T v, int)
{
    return v;
}
template<typename T>
inline T __shfl_down_sync(unsigned, T v, int)
{
    return v;
}
inline void __syncwarp(unsigned = 0xffffffffu) {}
inline void __syncthreads() {}    // emulated launches use one thread per block
#define __shared__ static
#define __constant__
inline void __threadfence() {}
template<typename T>
inline T __ldg(const T *p)
{
    return *p;
}
inline unsigned crb_lane_id() { return 0; }
inline float    __saturatef(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }

