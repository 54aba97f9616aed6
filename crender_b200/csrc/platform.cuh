// platform.cuh — compile-target shim for the crender_b200 kernels.
//
// The product is built by nvcc for sm_100a only (CRB_EMU undefined). There is no CPU fallback in the
// product: libcrender_b200.so fails loudly when no CUDA device is present.
//
// CRB_EMU is a TEST-HARNESS mode (tests/emu/, never shipped, never loaded by the package): the same
// kernel bodies are compiled by g++ against tests/emu/cuda_emu.h (serial execution, warp width 1) so that
// per-thread kernel logic can be checked against the oracle in the GPU-less development container.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <stdexcept>
#include <string>
#include <map>
#include <mutex>
#include <unordered_map>
#include <utility>

#ifndef CRB_EMU
#include <cuda_runtime.h>
#endif

namespace crb
{
    // error codes returned through the C ABI (include/crender_b200.h)
    enum : int
    {
        OK              = 0,
        ERR_GENERIC     = 1,     // data/errors.json "1"
        ERR_INVALID_ARG = 2,
        ERR_NO_DEVICE   = 10,
        ERR_CUDA        = 11,
        ERR_OOM         = 12,
        ERR_NCCL        = 13,
        ERR_BUILD_VERTS = 30,    // data/errors.json "30" (vertex buffer)
        ERR_BUILD_INDEX = 31,    // data/errors.json "31" (index buffer)
        ERR_BVH_DEPTH   = 32,
        ERR_NOT_COMMITTED = 33,
    };

    struct Error : std::runtime_error
    {
        int code;
        Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
    };
}    // namespace crb

#ifdef CRB_EMU
// the kernel-logic test harness (tests/emu, g++, serial execution): CUDA built-ins come from tests/emu/cuda_emu.h
#include "cuda_emu.h"
#else
// ------------------------------------------------------------------ CUDA (the product)
#define CRB_WARP 32
#define CRB_CUDA_CHECK(expr)                                                                                             \
    do {                                                                                                                 \
        cudaError_t _e = (expr);                                                                                         \
        if (_e != cudaSuccess)                                                                                           \
            throw crb::Error(_e == cudaErrorMemoryAllocation ? crb::ERR_OOM : crb::ERR_CUDA,                             \
                             std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    } while (0)
#define CRB_LAUNCH(kernel, grid, block, stream, ...)                  \
    do {                                                              \
        kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);        \
        CRB_CUDA_CHECK(cudaGetLastError());                           \
    } while (0)
__device__ __forceinline__ unsigned crb_lane_id()
{
    unsigned l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
#endif

namespace crb
{
    // Makes `dev` the calling thread's current CUDA device for the lifetime of the object and restores the previous one:
    // every C-ABI entry point that takes a handle runs under the handle's device, whatever the host thread (or another
    // library in the process, e.g. torch) had selected.
    struct DeviceScope
    {
#ifndef CRB_EMU
        int  prev = -1;
        bool changed = false;
        explicit DeviceScope(int dev)
        {
            if (cudaGetDevice(&prev) == cudaSuccess && prev != dev)
            {
                CRB_CUDA_CHECK(cudaSetDevice(dev));
                changed = true;
            }
        }
        ~DeviceScope()
        {
            if (changed) cudaSetDevice(prev);
        }
#else
        explicit DeviceScope(int) {}
#endif
        DeviceScope(const DeviceScope &)            = delete;
        DeviceScope &operator=(const DeviceScope &) = delete;
    };

    // ---- device memory helpers (cudaMalloc on the product; malloc in the test harness)
#ifndef CRB_EMU
    // Large device blocks are recycled instead of going back to the driver: cudaFree / cudaMalloc of the
    // multi-gigabyte path state costs tens of milliseconds, and an interactive host re-creates renderers on
    // every resolution change (renderer.cpp:194-201). Exact-size reuse per device, bounded, trimmed when an
    // allocation fails.
    struct DevBlockCache
    {
        static constexpr size_t MIN_BLOCK = size_t(1) << 20, MAX_CACHED = size_t(24) << 30;
        std::mutex                                          mu;
        std::multimap<std::pair<int, size_t>, void *>       free_blocks;    // (device, bytes) -> block
        std::unordered_map<void *, std::pair<int, size_t>>  live;           // blocks handed out that are cacheable
        size_t                                              cached = 0;
        std::unordered_map<int, size_t>                     live_bytes;     // per device: bytes of the blocks in `live`
        // per device: (total bytes, bytes held by anything that is not one of our blocks) as of the last driver query;
        // dropped when an allocation fails
        std::unordered_map<int, std::pair<size_t, size_t>>  budget;
        void trim_locked()
        {
            budget.clear();
            int cur = 0;
            cudaGetDevice(&cur);
            for (auto &kv : free_blocks)
            {
                cudaSetDevice(kv.first.first);
                cudaFree(kv.second);
            }
            cudaSetDevice(cur);
            free_blocks.clear();
            cached = 0;
        }
    };
    inline DevBlockCache &dev_block_cache()
    {
        static DevBlockCache c;
        return c;
    }
    inline size_t dev_cached_bytes()
    {
        DevBlockCache &c = dev_block_cache();
        std::lock_guard<std::mutex> lk(c.mu);
        return c.cached;
    }
    // Free device memory plus our recycled blocks on the current device, for planning. The driver is asked once per
    // device (and again after a failed allocation): cudaMemGetInfo was measured at 0.1-34 ms per call on B200
    // (profiles/r1k_submit_debug.txt), billed to the first render call of every new renderer. Afterwards the figure follows
    // our own large blocks; memory taken by others in between shows up as a failed allocation, which re-asks.
    inline size_t dev_available_bytes()
    {
        DevBlockCache &c = dev_block_cache();
        int            dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        std::lock_guard<std::mutex> lk(c.mu);
        const size_t                live = c.live_bytes[dev];
        auto                        it = c.budget.find(dev);
        if (it == c.budget.end())
        {
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
            {
                cudaGetLastError();
                return 0;
            }
            size_t cached_here = 0;
            for (const auto &kv : c.free_blocks)
                if (kv.first.first == dev) cached_here += kv.first.second;
            const size_t used = total_b - free_b, ours = live + cached_here;
            it = c.budget.insert({ dev, { total_b, used > ours ? used - ours : 0 } }).first;
        }
        const size_t total_b = it->second.first, foreign = it->second.second;
        return total_b > foreign + live ? total_b - foreign - live : 0;
    }
#endif
    inline void *dev_alloc(size_t bytes)
    {
        if (bytes == 0) bytes = 16;
#ifdef CRB_EMU
        void *p = aligned_alloc(256, (bytes + 255) & ~size_t(255));
        if (!p) throw Error(ERR_OOM, "alloc failed");
        return p;
#else
        void          *p = nullptr;
        DevBlockCache &c = dev_block_cache();
        int            dev = 0;
        CRB_CUDA_CHECK(cudaGetDevice(&dev));
        if (bytes >= DevBlockCache::MIN_BLOCK)
        {
            std::lock_guard<std::mutex> lk(c.mu);
            auto                        it = c.free_blocks.find({ dev, bytes });
            if (it != c.free_blocks.end())
            {
                p = it->second;
                c.free_blocks.erase(it);
                c.cached -= bytes;
                c.live[p] = { dev, bytes };
                c.live_bytes[dev] += bytes;
                return p;
            }
        }
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaErrorMemoryAllocation)
        {
            cudaGetLastError();
            {
                std::lock_guard<std::mutex> lk(c.mu);
                c.trim_locked();
            }
            e = cudaMalloc(&p, bytes);
        }
        CRB_CUDA_CHECK(e);
        if (bytes >= DevBlockCache::MIN_BLOCK)
        {
            std::lock_guard<std::mutex> lk(c.mu);
            c.live[p] = { dev, bytes };
            c.live_bytes[dev] += bytes;
        }
        return p;
#endif
    }
    inline void dev_free(void *p)
    {
        if (!p) return;
#ifdef CRB_EMU
        free(p);
#else
        DevBlockCache &c = dev_block_cache();
        {
            std::lock_guard<std::mutex> lk(c.mu);
            auto                        it = c.live.find(p);
            if (it != c.live.end())
            {
                const std::pair<int, size_t> key = it->second;
                c.live.erase(it);
                c.live_bytes[key.first] -= key.second;
                if (c.cached + key.second <= DevBlockCache::MAX_CACHED)
                {
                    // same guarantee as cudaFree: nothing on the device still uses the block when it is reused
                    int cur = 0;
                    cudaGetDevice(&cur);
                    if (cur != key.first) cudaSetDevice(key.first);
                    cudaDeviceSynchronize();
                    if (cur != key.first) cudaSetDevice(cur);
                    c.free_blocks.insert({ key, p });
                    c.cached += key.second;
                    return;
                }
            }
        }
        cudaFree(p);
#endif
    }
    inline void dev_upload(void *dst, const void *src, size_t bytes, cudaStream_t s)
    {
        if (!bytes) return;
#ifdef CRB_EMU
        memcpy(dst, src, bytes);
#else
        CRB_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
#endif
    }
    inline void dev_download(void *dst, const void *src, size_t bytes, cudaStream_t s)
    {
        if (!bytes) return;
#ifdef CRB_EMU
        memcpy(dst, src, bytes);
#else
        CRB_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
        CRB_CUDA_CHECK(cudaStreamSynchronize(s));
#endif
    }
    inline void dev_copy(void *dst, const void *src, size_t bytes, cudaStream_t s)
    {
        if (!bytes) return;
#ifdef CRB_EMU
        memcpy(dst, src, bytes);
#else
        CRB_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
#endif
    }
    inline void dev_zero(void *dst, size_t bytes, cudaStream_t s)
    {
        if (!bytes) return;
#ifdef CRB_EMU
        memset(dst, 0, bytes);
#else
        CRB_CUDA_CHECK(cudaMemsetAsync(dst, 0, bytes, s));
#endif
    }
    inline void dev_fill_byte(void *dst, int byte, size_t bytes, cudaStream_t s)
    {
        if (!bytes) return;
#ifdef CRB_EMU
        memset(dst, byte, bytes);
#else
        CRB_CUDA_CHECK(cudaMemsetAsync(dst, byte, bytes, s));
#endif
    }
    inline void stream_sync(cudaStream_t s)
    {
#ifndef CRB_EMU
        CRB_CUDA_CHECK(cudaStreamSynchronize(s));
#endif
    }

    template<typename T>
    struct DBuf
    {
        T     *p = nullptr;
        size_t n = 0;
        DBuf()   = default;
        DBuf(const DBuf &)            = delete;
        DBuf &operator=(const DBuf &) = delete;
        DBuf(DBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr, o.n = 0; }
        DBuf &operator=(DBuf &&o) noexcept
        {
            if (this != &o)
            {
                release();
                p = o.p, n = o.n, o.p = nullptr, o.n = 0;
            }
            return *this;
        }
        ~DBuf() { release(); }
        void release()
        {
            dev_free(p);
            p = nullptr, n = 0;
        }
        void alloc(size_t count)
        {
            if (count == n && p) return;
            release();
            p = static_cast<T *>(dev_alloc(count * sizeof(T)));
            n = count;
        }
        void ensure(size_t count)
        {
            if (count > n || !p) alloc(count);
        }
        size_t bytes() const { return n * sizeof(T); }
    };
}    // namespace crb
