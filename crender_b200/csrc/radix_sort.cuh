// radix_sort.cuh — hand-written stable LSD radix sort of (64-bit key, 32-bit value) pairs for the Morton
// ordering step of the BVH build (K1). 8 bits per pass; a pass is three kernels:
//
//   k_rs_hist     per-block digit histogram of a 2048-key tile (shared-memory counters), written bin-major
//   k_rs_scan     exclusive scan over the [256 bins x nblocks] table (one CTA, sequential over chunks)
//   k_rs_scatter  each warp owns a contiguous 256-key slice of the tile and walks it in order, so the
//                 sort is stable: per-warp digit counts -> warp base offsets (block offset from the scan +
//                 counts of the warps before it) -> rank inside a 32-key chunk by __match_any_sync
//
// Keys/values ping-pong between two buffers; the function returns which pair holds the result. The
// builder verifies the output order with k_check_sorted and fails loudly if it is not sorted.
#pragma once
#include "platform.cuh"

#ifndef CRB_EMU
namespace crb
{
    namespace radix
    {
        constexpr int RS_THREADS = 256;                          // 8 warps
        constexpr int RS_ITEMS   = 8;                            // keys per thread
        constexpr int RS_TILE    = RS_THREADS * RS_ITEMS;        // 2048 keys per CTA
        constexpr int RS_BINS    = 256;

        __device__ __forceinline__ unsigned digit_of(unsigned long long k, int shift) { return unsigned(k >> shift) & 0xffu; }

        __global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const unsigned long long *__restrict__ keys, uint32_t n, int shift, uint32_t *__restrict__ block_hist,
                                                                uint32_t nblocks)
        {
            __shared__ uint32_t hist[RS_BINS];
            hist[threadIdx.x] = 0;
            __syncthreads();
            const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
            for (int k = 0; k < RS_ITEMS; k++)
            {
                const uint32_t i = base + k * RS_THREADS + threadIdx.x;
                if (i < n) atomicAdd(&hist[digit_of(keys[i], shift)], 1u);
            }
            __syncthreads();
            block_hist[size_t(threadIdx.x) * nblocks + blockIdx.x] = hist[threadIdx.x];
        }

        // exclusive scan of `count` entries in place (count is a multiple of 256); one CTA of 1024 threads, every thread scans
        // 8 consecutive entries serially (two 16-byte loads) before the warp / CTA scan of the thread totals: 8192 entries per
        // round of four barriers (one entry per thread: 122 rounds for 1 M keys, 0.114 ms per pass x 8 passes of a 5.2 ms build)
        constexpr int RS_SCAN_ITEMS = 8;
        __global__ void __launch_bounds__(1024) k_rs_scan(uint32_t *__restrict__ data, uint32_t count)
        {
            __shared__ uint32_t warp_sums[32];
            __shared__ uint32_t carry;
            if (threadIdx.x == 0) carry = 0;
            __syncthreads();
            const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
            for (uint32_t base = 0; base < count; base += 1024 * RS_SCAN_ITEMS)
            {
                const uint32_t i  = base + threadIdx.x * RS_SCAN_ITEMS;
                const bool     ok = i < count;    // count % 8 == 0: a thread's 8 entries are all inside or all outside
                uint32_t       v[RS_SCAN_ITEMS];
                if (ok)
                {
                    const uint4 a = *reinterpret_cast<const uint4 *>(data + i), b = *reinterpret_cast<const uint4 *>(data + i + 4);
                    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < RS_SCAN_ITEMS; k++) v[k] = 0u;
                }
                uint32_t total = 0;
#pragma unroll
                for (int k = 0; k < RS_SCAN_ITEMS; k++)
                {
                    const uint32_t t = v[k];
                    v[k]             = total;    // exclusive inside the thread
                    total += t;
                }
                uint32_t x = total;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                    if (int(lane) >= o) x += y;
                }
                if (lane == 31) warp_sums[warp] = x;
                __syncthreads();
                if (warp == 0)
                {
                    uint32_t s = warp_sums[lane];
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1)
                    {
                        const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
                        if (int(lane) >= o) s += y;
                    }
                    warp_sums[lane] = s;    // inclusive over warps
                }
                __syncthreads();
                const uint32_t before = carry + (warp ? warp_sums[warp - 1] : 0u) + (x - total);
                if (ok)
                {
                    *reinterpret_cast<uint4 *>(data + i)     = make_uint4(before + v[0], before + v[1], before + v[2], before + v[3]);
                    *reinterpret_cast<uint4 *>(data + i + 4) = make_uint4(before + v[4], before + v[5], before + v[6], before + v[7]);
                }
                __syncthreads();
                if (threadIdx.x == 1023) carry = before + total;
                __syncthreads();
            }
        }

        __global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const unsigned long long *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                                   unsigned long long *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t n,
                                                                   int shift, const uint32_t *__restrict__ block_offs, uint32_t nblocks)
        {
            constexpr int   WARPS = RS_THREADS / 32;
            __shared__ uint32_t wcnt[WARPS][RS_BINS];
            for (int i = threadIdx.x; i < WARPS * RS_BINS; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
            __syncthreads();
            const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
            // warp `warp` owns keys [tile + warp*256, +256), chunk k = 32 consecutive keys
            const uint32_t     wbase = blockIdx.x * RS_TILE + warp * (32 * RS_ITEMS);
            unsigned long long key[RS_ITEMS];
            uint32_t           val[RS_ITEMS];
            unsigned           dig[RS_ITEMS];
#pragma unroll
            for (int k = 0; k < RS_ITEMS; k++)
            {
                const uint32_t i  = wbase + k * 32 + lane;
                const bool     ok = i < n;
                key[k]            = ok ? keys_in[i] : 0ull;
                val[k]            = ok ? vals_in[i] : 0u;
                dig[k]            = ok ? digit_of(key[k], shift) : 0xffffffffu;    // invalid lanes form their own match group
                const unsigned peers = __match_any_sync(0xffffffffu, dig[k]);
                if (ok && lane == unsigned(__ffs(int(peers)) - 1)) wcnt[warp][dig[k]] += uint32_t(__popc(peers));
                __syncwarp();
            }
            __syncthreads();
            // exclusive prefix over the warps of this CTA, seeded with the CTA's global offset per bin
            {
                const int bin     = threadIdx.x;
                uint32_t  running = block_offs[size_t(bin) * nblocks + blockIdx.x];
#pragma unroll
                for (int w = 0; w < WARPS; w++)
                {
                    const uint32_t c = wcnt[w][bin];
                    wcnt[w][bin]     = running;
                    running += c;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < RS_ITEMS; k++)
            {
                const uint32_t i     = wbase + k * 32 + lane;
                const bool     ok    = i < n;
                const unsigned peers = __match_any_sync(0xffffffffu, dig[k]);
                uint32_t       base  = 0;
                if (ok) base = wcnt[warp][dig[k]];
                __syncwarp();
                if (ok)
                {
                    const uint32_t dst = base + uint32_t(__popc(peers & ((1u << lane) - 1u)));
                    keys_out[dst]      = key[k];
                    vals_out[dst]      = val[k];
                    if (lane == unsigned(__ffs(int(peers)) - 1)) wcnt[warp][dig[k]] = base + uint32_t(__popc(peers));
                }
                __syncwarp();
            }
        }
    }    // namespace radix

    // Sorts n pairs by the low `bits` bits of the key. (k0,v0) holds the input; (k1,v1) is scratch of the
    // same size; hist is scratch of 256 * ceil(n/2048) uint32. Returns 0 if the result is in (k0,v0), 1 if
    // it is in (k1,v1).
    inline int radix_sort_pairs(unsigned long long *k0, uint32_t *v0, unsigned long long *k1, uint32_t *v1, uint32_t n, int bits, uint32_t *hist,
                                cudaStream_t stream)
    {
        using namespace radix;
        const uint32_t nblocks = (n + RS_TILE - 1) / RS_TILE;
        int            cur     = 0;
        for (int shift = 0; shift < bits; shift += 8)
        {
            unsigned long long *ki = cur ? k1 : k0, *ko = cur ? k0 : k1;
            uint32_t           *vi = cur ? v1 : v0, *vo = cur ? v0 : v1;
            CRB_LAUNCH(k_rs_hist, nblocks, RS_THREADS, stream, ki, n, shift, hist, nblocks);
            CRB_LAUNCH(k_rs_scan, 1, 1024, stream, hist, uint32_t(RS_BINS) * nblocks);
            CRB_LAUNCH(k_rs_scatter, nblocks, RS_THREADS, stream, ki, vi, ko, vo, n, shift, hist, nblocks);
            cur ^= 1;
        }
        return cur;
    }

    inline size_t radix_sort_hist_entries(uint32_t n) { return size_t(radix::RS_BINS) * ((n + radix::RS_TILE - 1) / radix::RS_TILE); }
}    // namespace crb
#endif
