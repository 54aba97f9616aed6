// bvh_build.cu — device-side BVH build for sm_100a.
//
// Replaces rtcCommitGeometry/rtcAttachGeometry/rtcCommitScene (src/objects/model.cpp:92-94), which in
// the reference is Embree's CPU binned-SAH builder. Pipeline, all on the device:
//   K1  k_prim_bounds  per-triangle AABBs + centroid bounds (warp-reduced ordered-int atomics)
//       k_morton       63-bit Morton keys of the centroids; hand-written stable LSD radix sort of (key, prim) pairs (radix_sort.cuh)
//   K2  k_hierarchy    LBVH topology from sorted keys (Karras 2012, one thread per inner node)
//       k_refit        bottom-up AABBs + subtree SAH cost with per-node arrival counters
//   K3  k_treelet      SAH treelet restructuring (Karras & Aila 2013, 7-leaf treelets, exact DP)
//   K4  k_collapse     level-synchronous collapse of the binary tree into 8-wide compressed nodes:
//                      greedy surface-area expansion to 8 children, leaves of <= 3 triangles,
//                      conservative 8-bit child-box quantisation (in double), octant slot assignment,
//                      triangles re-packed as (v0,e1,e2,prim) in node order.
// The published algorithms are followed as described in their papers; the code is original.
#include "bvh_build.cuh"

#include <algorithm>
#include <chrono>
#include <vector>

#include "radix_sort.cuh"

namespace crb
{
    namespace
    {
        // ---- ordered-int encoding so float min/max can use integer atomics
        __device__ __forceinline__ int   f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
        __device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

        struct BinTree
        {
            // inner node i in [0, n-1); leaf j in [0, n) is referenced as ~j (negative)
            int    *left, *right;          // children refs
            int    *parent;                // parent of inner node (-1 for root)
            int    *leaf_parent;           // parent of leaf (sorted position)
            int    *count;                 // triangles below the inner node (set by refit, kept by treelets)
            float4 *lo, *hi;               // inner node AABB (w of lo = subtree SAH cost, w of hi = area)
            int    *flags;                 // arrival counters
            float  *dpc;                   // collapse DP: C(n,1..7), 7 floats per inner node
            unsigned char *dpk;            // collapse DP: best left budget for j = 2..8 at [j-1], 8 bytes per inner node
        };

        __global__ void k_prim_bounds(const float *__restrict__ wv, uint32_t n, float4 *__restrict__ plo, float4 *__restrict__ phi, int *bounds)
        {
            const uint32_t i   = blockIdx.x * blockDim.x + threadIdx.x;
            const float    big = 3.0e38f;
            float          cx0 = big, cy0 = big, cz0 = big, cx1 = -big, cy1 = -big, cz1 = -big;
            if (i < n)
            {
                const float *v = wv + size_t(i) * 9;
                const float  lx = fminf(v[0], fminf(v[3], v[6])), ly = fminf(v[1], fminf(v[4], v[7])), lz = fminf(v[2], fminf(v[5], v[8]));
                const float  hx = fmaxf(v[0], fmaxf(v[3], v[6])), hy = fmaxf(v[1], fmaxf(v[4], v[7])), hz = fmaxf(v[2], fmaxf(v[5], v[8]));
                plo[i]          = make_float4(lx, ly, lz, 0.f);
                phi[i]          = make_float4(hx, hy, hz, 0.f);
                cx0 = cx1 = 0.5f * (lx + hx), cy0 = cy1 = 0.5f * (ly + hy), cz0 = cz1 = 0.5f * (lz + hz);
            }
#pragma unroll
            for (int o = CRB_WARP / 2; o > 0; o >>= 1)
            {
                cx0 = fminf(cx0, __shfl_xor_sync(0xffffffffu, cx0, o)), cy0 = fminf(cy0, __shfl_xor_sync(0xffffffffu, cy0, o));
                cz0 = fminf(cz0, __shfl_xor_sync(0xffffffffu, cz0, o)), cx1 = fmaxf(cx1, __shfl_xor_sync(0xffffffffu, cx1, o));
                cy1 = fmaxf(cy1, __shfl_xor_sync(0xffffffffu, cy1, o)), cz1 = fmaxf(cz1, __shfl_xor_sync(0xffffffffu, cz1, o));
            }
            if (crb_lane_id() == 0 && cx0 <= cx1)
            {
                atomicMin(bounds + 0, f2ord(cx0)), atomicMin(bounds + 1, f2ord(cy0)), atomicMin(bounds + 2, f2ord(cz0));
                atomicMax(bounds + 3, f2ord(cx1)), atomicMax(bounds + 4, f2ord(cy1)), atomicMax(bounds + 5, f2ord(cz1));
            }
        }

        __device__ __forceinline__ unsigned long long spread21(unsigned long long x)
        {
            x &= 0x1fffffull;
            x = (x | (x << 32)) & 0x1f00000000ffffull;
            x = (x | (x << 16)) & 0x1f0000ff0000ffull;
            x = (x | (x << 8)) & 0x100f00f00f00f00full;
            x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
            x = (x | (x << 2)) & 0x1249249249249249ull;
            return x;
        }

        __global__ void k_morton(const float4 *__restrict__ plo, const float4 *__restrict__ phi, uint32_t n, const int *__restrict__ bounds,
                                 unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals)
        {
            const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= n) return;
            const float  x0 = ord2f(bounds[0]), y0 = ord2f(bounds[1]), z0 = ord2f(bounds[2]);
            const float  ex = ord2f(bounds[3]) - x0, ey = ord2f(bounds[4]) - y0, ez = ord2f(bounds[5]) - z0;
            const float4 a = plo[i], b = phi[i];
            const float  cx = 0.5f * (a.x + b.x), cy = 0.5f * (a.y + b.y), cz = 0.5f * (a.z + b.z);
            const float  k  = 2097152.0f;    // 2^21
            const float  fx = ex > 0.f ? (cx - x0) / ex * k : 0.f, fy = ey > 0.f ? (cy - y0) / ey * k : 0.f, fz = ez > 0.f ? (cz - z0) / ez * k : 0.f;
            const unsigned long long qx = (unsigned long long) fminf(fmaxf(fx, 0.f), k - 1.f), qy = (unsigned long long) fminf(fmaxf(fy, 0.f), k - 1.f),
                                     qz = (unsigned long long) fminf(fmaxf(fz, 0.f), k - 1.f);
            keys[i] = (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
            vals[i] = i;
        }

        __global__ void k_check_sorted(const unsigned long long *__restrict__ keys, uint32_t n, uint32_t *bad)
        {
            const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i + 1 < n && keys[i] > keys[i + 1]) atomicAdd(bad, 1u);
        }

        // common-prefix length of sorted keys i and j, ties broken by position (Karras 2012 §4)
        __device__ __forceinline__ int delta(const unsigned long long *__restrict__ keys, int n, int i, int j)
        {
            if (j < 0 || j >= n) return -1;
            const unsigned long long a = keys[i], b = keys[j];
            if (a == b) return 64 + __clz(i ^ j);
            return __clzll((long long) (a ^ b));
        }

        __global__ void k_hierarchy(const unsigned long long *__restrict__ keys, int n, BinTree t)
        {
            const int i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= n - 1) return;
            const int d    = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
            const int dmin = delta(keys, n, i, i - d);
            int       lmax = 2;
            while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
            int l = 0;
            for (int s = lmax / 2; s >= 1; s /= 2)
                if (delta(keys, n, i, i + (l + s) * d) > dmin) l += s;
            const int j     = i + l * d;
            const int dnode = delta(keys, n, i, j);
            int       s     = 0;
            int       step  = l;
            do {
                step = (step + 1) >> 1;
                if (delta(keys, n, i, i + (s + step) * d) > dnode) s += step;
            } while (step > 1);
            const int gamma = i + s * d + (d < 0 ? -1 : 0);
            const int lo = i < j ? i : j, hi = i < j ? j : i;
            const int lref = (lo == gamma) ? ~gamma : gamma;
            const int rref = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
            t.left[i] = lref, t.right[i] = rref;
            if (lref < 0) t.leaf_parent[~lref] = i; else t.parent[lref] = i;
            if (rref < 0) t.leaf_parent[~rref] = i; else t.parent[rref] = i;
            if (i == 0) t.parent[0] = -1;
        }

        __device__ __forceinline__ float box_area(float4 lo, float4 hi)
        {
            const float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
            return 2.0f * (ex * ey + ey * ez + ez * ex);
        }

        constexpr float SAH_CI = 1.2f;    // cost of an inner node relative to a triangle test (Karras & Aila 2013)
        constexpr float SAH_CT = 1.0f;

        __device__ __forceinline__ void child_box(const BinTree &t, const float4 *plo, const float4 *phi, const uint32_t *vals, int ref, float4 &lo,
                                                  float4 &hi, float &cost)
        {
            if (ref < 0)
            {
                const uint32_t p = vals[~ref];
                lo = plo[p], hi = phi[p];
                cost = SAH_CT * box_area(lo, hi);
            }
            else
            {
                lo = t.lo[ref], hi = t.hi[ref];
                cost = lo.w;
            }
        }

        // one thread per leaf walks up; the second thread to arrive at an inner node computes it
        __global__ void k_refit(int n, BinTree t, const float4 *__restrict__ plo, const float4 *__restrict__ phi, const uint32_t *__restrict__ vals)
        {
            const int i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= n) return;
            int node = t.leaf_parent[i];
            while (node >= 0)
            {
                __threadfence();
                if (atomicAdd(t.flags + node, 1) == 0) return;
                __threadfence();
                float4 alo, ahi, blo, bhi;
                float  ca, cb;
                child_box(t, plo, phi, vals, t.left[node], alo, ahi, ca);
                child_box(t, plo, phi, vals, t.right[node], blo, bhi, cb);
                float4 lo   = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.f);
                float4 hi   = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.f);
                const float area = box_area(lo, hi);
                // SAH cost of the subtree: either an inner node over both children, or (small subtrees)
                // a flat leaf of all its triangles
                const int   cnt   = (t.left[node] < 0 ? 1 : t.count[t.left[node]]) + (t.right[node] < 0 ? 1 : t.count[t.right[node]]);
                t.count[node]     = cnt;
                float       cost  = SAH_CI * area + ca + cb;
                const float cleaf = SAH_CT * area * float(cnt);
                if (cnt <= BVH8_LEAF_TRIS && cleaf < cost) cost = cleaf;
                lo.w = cost, hi.w = area;
                t.lo[node] = lo, t.hi[node] = hi;
                node = t.parent[node];
            }
        }

        // ------------------------------------------------------------------ K4 collapse
        // Cost tables for the SAH-optimal binary -> 8-wide collapse (the dynamic programme of Ylitie,
        // Karras & Laine 2017 §4, as published): C(n,i) = cheapest way to represent the binary subtree n as
        // at most i children of one wide node.
        //   C(n,1) = leaf cost A(n) P(n) c_prim           if P(n) <= 3
        //          = A(n) c_node + D(n,8)                  otherwise (n becomes a wide node)
        //   C(n,i) = min(D(n,i), C(n,i-1)),  D(n,j) = min_{0<k<j} C(left,k) + C(right,j-k)
        // c_node : c_prim follows the measured instruction cost of a node test vs a triangle test (~3:1).
        constexpr float DP_CNODE = 1.0f;

        __global__ void k_collapse_dp(int n, BinTree t, const float4 *__restrict__ plo, const float4 *__restrict__ phi, const uint32_t *__restrict__ vals,
                                      float DP_CPRIM)
        {
            const int i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= n) return;
            int node = t.leaf_parent[i];
            while (node >= 0)
            {
                __threadfence();
                if (atomicAdd(t.flags + node, 1) == 0) return;
                __threadfence();
                float cl[8], cr[8];
                for (int side = 0; side < 2; side++)
                {
                    const int ref = side ? t.right[node] : t.left[node];
                    float    *dst = side ? cr : cl;
                    if (ref < 0)
                    {
                        const uint32_t p = vals[~ref];
                        const float    a = box_area(plo[p], phi[p]) * DP_CPRIM;
                        for (int k = 1; k < 8; k++) dst[k] = a;
                    }
                    else
                        for (int k = 1; k < 8; k++) dst[k] = t.dpc[size_t(ref) * 7 + (k - 1)];
                }
                float         dist[9];
                unsigned char kb[9];
                for (int j = 2; j <= 8; j++)
                {
                    float best = 3e38f;
                    int   bk   = 1;
                    for (int k = 1; k < j; k++)
                    {
                        if (k > 7 || j - k > 7) continue;
                        const float c = cl[k] + cr[j - k];
                        if (c < best) best = c, bk = k;
                    }
                    dist[j] = best, kb[j] = (unsigned char) bk;
                }
                const float area = t.hi[node].w;
                const int   cnt  = t.count[node];
                float      *out  = t.dpc + size_t(node) * 7;
                out[0]           = cnt <= BVH8_LEAF_TRIS ? area * float(cnt) * DP_CPRIM : area * DP_CNODE + dist[8];
                for (int k = 2; k <= 7; k++) out[k - 1] = fminf(dist[k], out[k - 2]);
                unsigned char *ko = t.dpk + size_t(node) * 8;
                ko[0]             = 0;
                for (int j = 2; j <= 8; j++) ko[j - 1] = kb[j];
                node = t.parent[node];
            }
        }

        struct CollapseCtx
        {
            BinTree         t;
            const float4   *plo, *phi;
            const uint32_t *vals;
            const float    *wv;
            uint4          *nodes;
            float4         *tris;
            uint32_t       *counters;    // [0] nodes allocated, [1] tris allocated, [8 + L] queue size of level L (zeroed per build)
            float          *sah;         // accumulated wide-tree SAH cost (area-weighted), informational
            const float4   *root_hi;     // hi[0] of the binary tree: .w = the root's surface area (null: a single leaf)
            int             use_dp;
            uint32_t        max_nodes;   // capacity of the node pool and of both collapse queues; overflow sets counters[7]
        };

        __device__ __forceinline__ int ref_count(const BinTree &t, int ref) { return ref < 0 ? 1 : t.count[ref]; }
        // sorted positions of the (<= BVH8_LEAF_TRIS) leaves below ref, left to right
        __device__ __forceinline__ int ref_leaves(const BinTree &t, int ref, int *out)
        {
            int n = 0, sp = 0, stack[BVH8_LEAF_TRIS + 1];
            stack[sp++] = ref;
            while (sp)
            {
                const int r = stack[--sp];
                if (r < 0)
                    out[n++] = ~r;
                else
                    stack[sp++] = t.right[r], stack[sp++] = t.left[r];
            }
            return n;
        }

        // One level of the wide tree per launch. The level's queue size is read from the device (counters[8 + level]) and the
        // launch covers the level's upper bound (min(8^level, pool)), so the host can queue a whole group of levels without
        // reading anything back (round 1 and most of round 2: one blocking 32-byte read per level, ~0.3 ms of a 5 ms build).
        constexpr uint32_t COLLAPSE_MAX_LEVELS = 72, COLLAPSE_GROUP = 12;
        __global__ void k_collapse(CollapseCtx c, const uint2 *__restrict__ in, uint32_t level, uint2 *__restrict__ out)
        {
            const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
            uint32_t       n_in = c.counters[8 + level];
            n_in                = n_in < c.max_nodes ? n_in : c.max_nodes;
            // an overflow flagged by an earlier level: the host re-runs the stage, nothing below is worth writing
            if (w >= n_in || *(volatile uint32_t *) (c.counters + 7)) return;
            const int      root = int(in[w].x);
            const uint32_t self = in[w].y;

            int ch[8];
            int k = 0;
            if (root >= 0)
                ch[0] = c.t.left[root], ch[1] = c.t.right[root], k = 2;
            else
                ch[0] = root, k = 1;
            if (c.use_dp && root >= 0)
            {
                // SAH-optimal choice of the (at most 8) children from the DP tables of k_collapse_dp:
                // distribute 8 slots over the two binary children, recursively, until a budget of one slot
                // (or a budget whose extra slots buy nothing) turns a binary subtree into a child.
                int sref[16], sbud[16], sp = 0;
                k                = 0;
                const int k8     = c.t.dpk[size_t(root) * 8 + 7];
                sref[sp] = c.t.right[root], sbud[sp] = 8 - k8, sp++;
                sref[sp] = c.t.left[root], sbud[sp] = k8, sp++;
                while (sp)
                {
                    --sp;
                    const int m = sref[sp];
                    int       j = sbud[sp];
                    if (m < 0)
                    {
                        ch[k++] = m;
                        continue;
                    }
                    const float *cm = c.t.dpc + size_t(m) * 7;
                    while (j > 1 && cm[j - 1] >= cm[j - 2]) j--;    // C(m,j) == C(m,j-1): the extra slot is useless
                    if (j == 1)
                    {
                        ch[k++] = m;
                        continue;
                    }
                    const int kl = c.t.dpk[size_t(m) * 8 + (j - 1)];
                    sref[sp] = c.t.right[m], sbud[sp] = j - kl, sp++;
                    sref[sp] = c.t.left[m], sbud[sp] = kl, sp++;
                }
            }
            else
            {
                // greedy fallback: open the inner child with the largest surface area until 8 children
                while (k < 8)
                {
                    int   best  = -1;
                    float besta = -1.f;
                    for (int j = 0; j < k; j++)
                        if (ch[j] >= 0)
                        {
                            const float a = c.t.hi[ch[j]].w;
                            if (a > besta) besta = a, best = j;
                        }
                    if (best < 0) break;
                    const int r = ch[best];
                    ch[best]    = c.t.left[r];
                    ch[k++]     = c.t.right[r];
                }
            }

            float4 clo[8], chi[8];
            float  nlo[3] = { 3e38f, 3e38f, 3e38f }, nhi[3] = { -3e38f, -3e38f, -3e38f };
            for (int j = 0; j < k; j++)
            {
                float cost;
                child_box(c.t, c.plo, c.phi, c.vals, ch[j], clo[j], chi[j], cost);
                nlo[0] = fminf(nlo[0], clo[j].x), nlo[1] = fminf(nlo[1], clo[j].y), nlo[2] = fminf(nlo[2], clo[j].z);
                nhi[0] = fmaxf(nhi[0], chi[j].x), nhi[1] = fmaxf(nhi[1], chi[j].y), nhi[2] = fmaxf(nhi[2], chi[j].z);
            }

            // slot assignment: child j -> slot s maximising (centroid_j - centroid_node) . sign(s), greedily
            int  child_in_slot[8];
            bool slot_used[8], child_done[8];
            for (int j = 0; j < 8; j++) slot_used[j] = false, child_done[j] = false, child_in_slot[j] = -1;
            const float ncx = 0.5f * (nlo[0] + nhi[0]), ncy = 0.5f * (nlo[1] + nhi[1]), ncz = 0.5f * (nlo[2] + nhi[2]);
            for (int round = 0; round < k; round++)
            {
                float bestc = -3e38f;
                int   bj = -1, bs = -1;
                for (int j = 0; j < k; j++)
                {
                    if (child_done[j]) continue;
                    const float dx = 0.5f * (clo[j].x + chi[j].x) - ncx, dy = 0.5f * (clo[j].y + chi[j].y) - ncy, dz = 0.5f * (clo[j].z + chi[j].z) - ncz;
                    for (int s = 0; s < 8; s++)
                    {
                        if (slot_used[s]) continue;
                        const float cost = ((s & 1) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 4) ? dz : -dz);
                        if (cost > bestc) bestc = cost, bj = j, bs = s;
                    }
                }
                child_in_slot[bs] = bj, slot_used[bs] = true, child_done[bj] = true;
            }

            // classify and allocate
            int n_inner = 0, n_ltris = 0;
            for (int s = 0; s < 8; s++)
            {
                const int j = child_in_slot[s];
                if (j < 0) continue;
                const int cnt = ref_count(c.t, ch[j]);
                if (cnt <= BVH8_LEAF_TRIS) n_ltris += cnt; else n_inner++;
            }
            const uint32_t child_base = n_inner ? atomicAdd(c.counters + 0, uint32_t(n_inner)) : 0u;
            const uint32_t tri_base   = n_ltris ? atomicAdd(c.counters + 1, uint32_t(n_ltris)) : 0u;
            const uint32_t out_base   = n_inner ? atomicAdd(c.counters + 9 + level, uint32_t(n_inner)) : 0u;
            // n/2+8 nodes hold for the greedy collapse (a non-full node has only leaf children); the DP collapse can
            // plateau (C(m,j) == C(m,j-1)) and emit thinner nodes. Never write past the pool / queues: flag, and the
            // host rebuilds this stage with the true bound (every wide inner node consumes a binary inner node: <= n)
            if (n_inner && (child_base + uint32_t(n_inner) > c.max_nodes || out_base + uint32_t(n_inner) > c.max_nodes))
            {
                atomicOr(c.counters + 7, 1u);
                return;
            }

            // quantisation frame (double: the grid must contain every child box exactly-conservatively)
            unsigned eb[3];
            double   scale[3];
            for (int a = 0; a < 3; a++)
            {
                const double ext = double(nhi[a]) - double(nlo[a]);
                int          e   = -126;
                if (ext > 0.0)
                {
                    int x;
                    frexp(ext / 255.0, &x);    // ext/255 = m * 2^x, m in [0.5,1)  ->  2^x > ext/255
                    e = x;
                }
                int b = e + 127;
                b     = b < 1 ? 1 : (b > 254 ? 254 : b);
                eb[a] = unsigned(b);
                scale[a] = ldexp(1.0, b - 127);
            }

            unsigned meta[8], qlo[3][8], qhi[3][8];
            unsigned imask = 0;
            int      inner_i = 0;
            float    leaf_area_tris = 0.f;
            // leaf triangles are stored in NIBBLE order of the slots (0,4,1,5,2,6,3,7): the order of the bits of the
            // traversal's occupancy word ((meta[4..7] << 4) | meta[0..3]) & 0x77777777 (bvh8.cuh)
            int tri_off_of[8];
            {
                int off = 0;
                for (int nb = 0; nb < 8; nb++)
                {
                    const int s = ((nb & 1) << 2) | (nb >> 1), j = child_in_slot[s];
                    tri_off_of[s] = off;
                    if (j < 0) continue;
                    const int cnt = ref_count(c.t, ch[j]);
                    if (cnt <= BVH8_LEAF_TRIS) off += cnt;
                }
            }
            for (int s = 0; s < 8; s++)
            {
                meta[s] = 0;
                for (int a = 0; a < 3; a++) qlo[a][s] = 255u, qhi[a][s] = 0u;
                const int j = child_in_slot[s];
                if (j < 0) continue;
                const int tri_off = tri_off_of[s];
                const float lo3[3] = { clo[j].x, clo[j].y, clo[j].z }, hi3[3] = { chi[j].x, chi[j].y, chi[j].z };
                for (int a = 0; a < 3; a++)
                {
                    const double p = double(nlo[a]);
                    double       q = floor((double(lo3[a]) - p) / scale[a]);
                    q              = q < 0.0 ? 0.0 : (q > 255.0 ? 255.0 : q);
                    while (q > 0.0 && p + q * scale[a] > double(lo3[a])) q -= 1.0;
                    qlo[a][s] = unsigned(q);
                    double r  = ceil((double(hi3[a]) - p) / scale[a]);
                    r         = r < 0.0 ? 0.0 : (r > 255.0 ? 255.0 : r);
                    while (r < 255.0 && p + r * scale[a] < double(hi3[a])) r += 1.0;
                    qhi[a][s] = unsigned(r);
                }
                const int cnt = ref_count(c.t, ch[j]);
                if (cnt <= BVH8_LEAF_TRIS)
                {
                    meta[s]         = (1u << cnt) - 1u;
                    int leaves[BVH8_LEAF_TRIS];
                    ref_leaves(c.t, ch[j], leaves);
                    for (int q = 0; q < cnt; q++)
                    {
                        const uint32_t prim = c.vals[leaves[q]];
                        const float   *v    = c.wv + size_t(prim) * 9;
                        float4        *dst  = c.tris + size_t(tri_base + uint32_t(tri_off + q)) * BVH8_TRI_F4;
                        dst[0]              = make_float4(v[0], v[1], v[2], __uint_as_float(prim));
                        dst[1]              = make_float4(__fsub_rn(v[3], v[0]), __fsub_rn(v[4], v[1]), __fsub_rn(v[5], v[2]), 0.f);
                        dst[2]              = make_float4(__fsub_rn(v[6], v[0]), __fsub_rn(v[7], v[1]), __fsub_rn(v[8], v[2]), 0.f);
                        if (BVH8_TRI_F4 > 3) dst[3] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    leaf_area_tris += box_area(clo[j], chi[j]) * float(cnt);
                }
                else
                {
                    meta[s] = 0x80u;
                    imask |= 1u << s;
                    out[out_base + uint32_t(inner_i)] = make_uint2(unsigned(ch[j]), child_base + uint32_t(inner_i));
                    inner_i++;
                }
            }

            auto pack4 = [](const unsigned *b) { return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24); };
            uint4 *np  = c.nodes + size_t(self) * BVH8_NODE_U4;
            np[0]      = make_uint4(__float_as_uint(nlo[0]), __float_as_uint(nlo[1]), __float_as_uint(nlo[2]), eb[0] | (eb[1] << 8) | (eb[2] << 16) | (imask << 24));
            np[1]      = make_uint4(child_base, tri_base, pack4(meta), pack4(meta + 4));
            np[2]      = make_uint4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4));
            np[3]      = make_uint4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
            np[4]      = make_uint4(pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
            if (BVH8_NODE_U4 > 5) np[5] = make_uint4(0u, 0u, 0u, 0u);    // padding of the 32-byte-aligned slot

            const float root_area = c.root_hi ? c.root_hi->w : 0.f;
            if (root_area > 0.f)
            {
                const float na = 2.0f * ((nhi[0] - nlo[0]) * (nhi[1] - nlo[1]) + (nhi[1] - nlo[1]) * (nhi[2] - nlo[2]) + (nhi[2] - nlo[2]) * (nhi[0] - nlo[0]));
                atomicAdd(c.sah, (SAH_CI * na + SAH_CT * leaf_area_tris) / root_area);
            }
        }

        // Models of at most 8 triangles (area-light quads, proxies of a small TLAS ...) are ONE node with one triangle per
        // slot: written by a single thread, no sort, no hierarchy, no host round trips (the full pipeline costs ~0.3 ms of
        // launch and synchronisation latency whatever the size). Same conservative quantisation as k_collapse; triangle j sits in
        // slot nibble_order[j], so the triangle array is in the bit order of the occupancy word (bvh8.cuh).
        constexpr uint32_t TINY_BVH_MAX = 8;
        __global__ void k_tiny_bvh(const float *__restrict__ wv, uint32_t n, uint4 *__restrict__ nodes, float4 *__restrict__ tris)
        {
            if (blockIdx.x * blockDim.x + threadIdx.x != 0) return;
            float lo[8][3], hi[8][3];
            float nlo[3] = { 3e38f, 3e38f, 3e38f }, nhi[3] = { -3e38f, -3e38f, -3e38f };
            for (uint32_t j = 0; j < n; j++)
            {
                const float *v = wv + size_t(j) * 9;
                for (int a = 0; a < 3; a++)
                {
                    lo[j][a] = fminf(v[a], fminf(v[3 + a], v[6 + a])), hi[j][a] = fmaxf(v[a], fmaxf(v[3 + a], v[6 + a]));
                    nlo[a] = fminf(nlo[a], lo[j][a]), nhi[a] = fmaxf(nhi[a], hi[j][a]);
                }
            }
            unsigned eb[3];
            double   scale[3];
            for (int a = 0; a < 3; a++)
            {
                const double ext = double(nhi[a]) - double(nlo[a]);
                int          e   = -126;
                if (ext > 0.0)
                {
                    int x;
                    frexp(ext / 255.0, &x);
                    e = x;
                }
                int b = e + 127;
                b     = b < 1 ? 1 : (b > 254 ? 254 : b);
                eb[a] = unsigned(b);
                scale[a] = ldexp(1.0, b - 127);
            }
            unsigned meta[8], qlo[3][8], qhi[3][8];
            for (int s = 0; s < 8; s++)
            {
                meta[s] = 0;
                for (int a = 0; a < 3; a++) qlo[a][s] = 255u, qhi[a][s] = 0u;
            }
            for (uint32_t j = 0; j < n; j++)
            {
                const int s = int(((j & 1u) << 2) | (j >> 1));    // nibble j of the occupancy word <-> slot
                for (int a = 0; a < 3; a++)
                {
                    const double p = double(nlo[a]);
                    double       q = floor((double(lo[j][a]) - p) / scale[a]);
                    q              = q < 0.0 ? 0.0 : (q > 255.0 ? 255.0 : q);
                    while (q > 0.0 && p + q * scale[a] > double(lo[j][a])) q -= 1.0;
                    qlo[a][s] = unsigned(q);
                    double r  = ceil((double(hi[j][a]) - p) / scale[a]);
                    r         = r < 0.0 ? 0.0 : (r > 255.0 ? 255.0 : r);
                    while (r < 255.0 && p + r * scale[a] < double(hi[j][a])) r += 1.0;
                    qhi[a][s] = unsigned(r);
                }
                meta[s]        = 1u;    // a leaf of one triangle
                const float *v = wv + size_t(j) * 9;
                tris[j * BVH8_TRI_F4 + 0] = make_float4(v[0], v[1], v[2], __uint_as_float(j));
                tris[j * BVH8_TRI_F4 + 1] = make_float4(__fsub_rn(v[3], v[0]), __fsub_rn(v[4], v[1]), __fsub_rn(v[5], v[2]), 0.f);
                tris[j * BVH8_TRI_F4 + 2] = make_float4(__fsub_rn(v[6], v[0]), __fsub_rn(v[7], v[1]), __fsub_rn(v[8], v[2]), 0.f);
                if (BVH8_TRI_F4 > 3) tris[j * BVH8_TRI_F4 + 3] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            auto pack4 = [](const unsigned *b) { return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24); };
            nodes[0]   = make_uint4(__float_as_uint(nlo[0]), __float_as_uint(nlo[1]), __float_as_uint(nlo[2]), eb[0] | (eb[1] << 8) | (eb[2] << 16));
            nodes[1]   = make_uint4(0u, 0u, pack4(meta), pack4(meta + 4));
            nodes[2]   = make_uint4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4));
            nodes[3]   = make_uint4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
            nodes[4]   = make_uint4(pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
            if (BVH8_NODE_U4 > 5) nodes[5] = make_uint4(0u, 0u, 0u, 0u);
        }

#include "bvh_treelet.inl"

        template<typename T>
        T *carve(char *&p, size_t count)
        {
            T *r = reinterpret_cast<T *>(p);
            p += (count * sizeof(T) + 255) & ~size_t(255);
            return r;
        }
    }    // namespace

    void build_bvh8(const float *d_wverts, uint32_t n, cudaStream_t stream, const BuildOptions &opt, DBuf<uint4> &nodes, DBuf<float4> &tris,
                    BuildStats &stats)
    {
        stats = BuildStats();
        if (n > 0x7ffffff0u) throw Error(ERR_BUILD_INDEX, "too many triangles for 32-bit primitive ids");
        const size_t max_nodes = size_t(n) / 2 + 8;
        nodes.alloc(max_nodes * BVH8_NODE_U4);
        tris.alloc(size_t(n ? n : 1) * BVH8_TRI_F4);
        if (n > 0 && n <= TINY_BVH_MAX && !getenv("CRB_NO_TINY_BVH"))
        {
            // one node, one triangle per slot: a single launch, no host round trip (the stream orders it before any use)
            CRB_LAUNCH(k_tiny_bvh, 1, 1, stream, d_wverts, n, nodes.p, tris.p);
            stats.n_nodes = 1, stats.n_tris = n, stats.max_depth = 1, stats.build_ms = 0.0;
            return;
        }

        if (n == 0)
        {
            // empty scene: one root with no children
            std::vector<uint4> root(5, make_uint4(0, 0, 0, 0));
            root[0] = make_uint4(0, 0, 0, 127u | (127u << 8) | (127u << 16));
            dev_upload(nodes.p, root.data(), 80, stream);
            stream_sync(stream);
            stats.n_nodes = 1, stats.max_depth = 1;
            return;
        }

        // ---- scratch
        const size_t ni = n > 1 ? n - 1 : 1;
        size_t       bytes = 0;
        auto         sz    = [&](size_t count, size_t elem) { bytes += (count * elem + 255) & ~size_t(255); };
        sz(n, 16), sz(n, 16);                 // plo, phi
        sz(8, 4);                             // bounds
        sz(n, 8), sz(n, 8), sz(n, 4), sz(n, 4);    // keys x2, vals x2
        sz(ni, 4), sz(ni, 4), sz(ni, 4), sz(n, 4), sz(ni, 4), sz(ni, 16), sz(ni, 16), sz(ni, 4), sz(ni, 28), sz(ni, 8);    // tree + collapse DP tables
        sz(max_nodes, 8), sz(max_nodes, 8);   // collapse queues
        sz(8 + COLLAPSE_MAX_LEVELS + 8, 4);   // counters + sah + per-level queue sizes of the collapse + treelet counter
#ifndef CRB_EMU
        sz(radix_sort_hist_entries(n), 4);    // radix sort histograms
#endif
        DBuf<char> scratch;
        scratch.alloc(bytes + 4096);
        // build_ms is the device pipeline from the first kernel to the last: the three allocations above are recycled blocks
        // in steady state (DevBlockCache) but cost whatever the driver's allocator costs the first time (measured: 20 - 160 ms
        // of jitter on the config-4 commit); the end-to-end figures (bench.py e2e.setup_ms) include them
#ifdef CRB_EMU
        auto t0 = std::chrono::steady_clock::now();
#else
        cudaEvent_t ev0, ev1;
        CRB_CUDA_CHECK(cudaEventCreate(&ev0));
        CRB_CUDA_CHECK(cudaEventCreate(&ev1));
        CRB_CUDA_CHECK(cudaEventRecord(ev0, stream));
#endif
        char   *p    = scratch.p;
        float4 *plo  = carve<float4>(p, n), *phi = carve<float4>(p, n);
        int    *bounds = carve<int>(p, 8);
        unsigned long long *keys0 = carve<unsigned long long>(p, n), *keys1 = carve<unsigned long long>(p, n);
        uint32_t *vals0 = carve<uint32_t>(p, n), *vals1 = carve<uint32_t>(p, n);
        BinTree   t;
        t.left = carve<int>(p, ni), t.right = carve<int>(p, ni), t.parent = carve<int>(p, ni), t.leaf_parent = carve<int>(p, n);
        t.count = carve<int>(p, ni);
        t.lo = carve<float4>(p, ni), t.hi = carve<float4>(p, ni);
        t.flags      = carve<int>(p, ni);
        t.dpc        = carve<float>(p, ni * 7);
        t.dpk        = carve<unsigned char>(p, ni * 8);
        uint2    *q0_small = carve<uint2>(p, max_nodes), *q1_small = carve<uint2>(p, max_nodes);
        uint32_t *counters = carve<uint32_t>(p, 8 + COLLAPSE_MAX_LEVELS + 8);
#ifndef CRB_EMU
        uint32_t *rs_hist = carve<uint32_t>(p, radix_sort_hist_entries(n));
#endif

        const int      B  = 256;
        const unsigned gn = (n + B - 1) / B;

        // ---- K1
        {
            const int big = 0x7f7fffff;
            int       init[8] = { big, big, big, ~big, ~big, ~big, 0, 0 };
            // ordered encoding of +max float is `big`; of -max float is (0xff7fffff ^ 0x7fffffff) = 0x80800000
            init[3] = init[4] = init[5] = int(0x80800000u);
            dev_upload(bounds, init, sizeof(init), stream);
        }
        CRB_LAUNCH(k_prim_bounds, gn, B, stream, d_wverts, n, plo, phi, bounds);
        CRB_LAUNCH(k_morton, gn, B, stream, plo, phi, n, bounds, keys0, vals0);

        const unsigned long long *keys = keys1;
        const uint32_t           *vals = vals1;
#ifdef CRB_EMU
        {
            std::vector<std::pair<unsigned long long, uint32_t>> kv(n);
            for (uint32_t i = 0; i < n; i++) kv[i] = { keys0[i], vals0[i] };
            std::stable_sort(kv.begin(), kv.end(), [](auto &a, auto &b) { return a.first < b.first; });
            for (uint32_t i = 0; i < n; i++) keys1[i] = kv[i].first, vals1[i] = kv[i].second;
        }
#else
        {
            // hand-written stable LSD radix sort (radix_sort.cuh): 8 passes of 8 bits over the 63-bit keys
            const int where = radix_sort_pairs(keys0, vals0, keys1, vals1, n, 63, rs_hist, stream);
            keys            = where ? keys1 : keys0;
            vals            = where ? vals1 : vals0;
            // loud failure instead of a silently broken tree
            dev_zero(counters + 6, 4, stream);
            CRB_LAUNCH(k_check_sorted, gn, B, stream, keys, n, counters + 6);
            uint32_t bad = 0;
            dev_download(&bad, counters + 6, 4, stream);
            if (bad) throw Error(ERR_GENERIC, "internal: Morton key sort produced an unsorted sequence");
        }
#endif

        // ---- K2
        int   root_ref  = ~0;    // single triangle: the root reference is leaf 0
        bool  treelets_ran = false;
        if (n > 1)
        {
            dev_zero(t.flags, ni * sizeof(int), stream);
            CRB_LAUNCH(k_hierarchy, (unsigned(ni) + B - 1) / B, B, stream, keys, int(n), t);
            CRB_LAUNCH(k_refit, gn, B, stream, int(n), t, plo, phi, vals);
            // ---- K3
            if (opt.treelet_passes > 0 && n >= 16)
            {
                // the count of changed treelets is read back with the collapse's counters (no round trip of its own)
                dev_zero(counters + 8 + COLLAPSE_MAX_LEVELS, 4, stream);
                treelet_optimize(t, int(n), plo, phi, vals, stream, opt.treelet_passes, counters + 8 + COLLAPSE_MAX_LEVELS);
                treelets_ran = true;
            }
            root_ref = 0;
        }

        // ---- K4
        size_t   pool_nodes = max_nodes;
        uint2   *q0 = q0_small, *q1 = q1_small;
        DBuf<uint2> q_big;
        uint32_t depth = 0;
        uint32_t cnt[8 + COLLAPSE_MAX_LEVELS + 8] = {};    // the collapse's counters as of its last group of levels (+ the treelet counter)
        for (int attempt = 0;; attempt++)
        {
            {
                uint32_t init[8 + COLLAPSE_MAX_LEVELS] = { 1, 0, 0, 0, 0, 0, 0, 0, 1 };    // node 0 = root is pre-allocated; level 0 = the root
                dev_upload(counters, init, sizeof(init), stream);
                uint2 first = make_uint2(unsigned(root_ref), 0u);
                dev_upload(q0, &first, sizeof(first), stream);
            }
            CollapseCtx c;
            c.t = t, c.plo = plo, c.phi = phi, c.vals = vals, c.wv = d_wverts, c.nodes = nodes.p, c.tris = tris.p;
            c.counters = counters, c.sah = reinterpret_cast<float *>(counters + 4), c.root_hi = n > 1 ? t.hi : nullptr;
            c.use_dp    = (opt.optimal_collapse && n > 1) ? 1 : 0;
            c.max_nodes = uint32_t(pool_nodes);
            if (c.use_dp && attempt == 0)
            {
                dev_zero(t.flags, ni * sizeof(int), stream);
                CRB_LAUNCH(k_collapse_dp, gn, B, stream, int(n), t, plo, phi, vals, opt.cost_prim);
            }
            bool     overflow = false;
            uint2   *qin = q0, *qout = q1;
            uint32_t level = 0;
            depth          = 0;
            uint32_t group = COLLAPSE_GROUP;
            if (const char *e = getenv("CRB_COLLAPSE_GROUP")) group = uint32_t(std::max(1, atoi(e)));    // tests: trees deeper than one group
            for (bool done = false; !done;)
            {
                // a group of levels without a host round trip: level L has at most min(8^L, pool) entries
                for (uint32_t g = 0; g < group && level + 1 < COLLAPSE_MAX_LEVELS; g++, level++)
                {
                    const size_t bound = level < 8 ? std::min<size_t>(size_t(1) << (3 * level), pool_nodes) : pool_nodes;
                    CRB_LAUNCH(k_collapse, unsigned((bound + 127) / 128), 128, stream, c, qin, level, qout);
                    std::swap(qin, qout);
                }
                dev_download(cnt, counters, sizeof(cnt), stream);
                if (cnt[7] || cnt[0] > pool_nodes)
                {
                    overflow = true;
                    break;
                }
                depth = 0;
                while (depth < level && cnt[8 + depth]) depth++;
                done = cnt[8 + level] == 0;    // the next level's queue is empty
                if (!done && level + 1 >= COLLAPSE_MAX_LEVELS) throw Error(ERR_BVH_DEPTH, "BVH deeper than " + std::to_string(COLLAPSE_MAX_LEVELS) + " levels");
            }
            if (!overflow) break;
            if (attempt > 0) throw Error(ERR_GENERIC, "internal: wide node pool overflow");
            // the true bound: every wide inner node consumes at least one binary inner node
            pool_nodes = size_t(n) + 8;
            nodes.alloc(pool_nodes * BVH8_NODE_U4);
            q_big.alloc(pool_nodes * 2);
            q0 = q_big.p, q1 = q_big.p + pool_nodes;
        }
        stats.n_nodes = cnt[0], stats.n_tris = cnt[1], stats.max_depth = depth;
        if (treelets_ran) stats.treelets_changed = cnt[8 + COLLAPSE_MAX_LEVELS];
        memcpy(&stats.sah_cost, &cnt[4], 4);
        if (stats.n_tris != n) throw Error(ERR_GENERIC, "internal: triangle count mismatch after collapse");
        if (stats.n_nodes > 0x00ffffffu) throw Error(ERR_BUILD_INDEX, "too many BVH nodes for the 24-bit child index of the traversal's stack entry");
        if (depth > uint32_t(BVH8_STACK)) throw Error(ERR_BVH_DEPTH, "BVH depth " + std::to_string(depth) + " exceeds the traversal stack");
#ifdef CRB_EMU
        stats.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (getenv("CRB_BUILD_STATS"))
        {
            // kernel-logic harness only: occupancy of the 8 child slots (tools/tree_quality.py)
            unsigned long long hc[9] = {}, hi_[9] = {}, hl[9] = {}, halves[4] = {};
            for (uint32_t i = 0; i < stats.n_nodes; i++)
            {
                const uint4 n1 = nodes.p[size_t(i) * BVH8_NODE_U4 + 1];
                int nc = 0, nin = 0, lo4 = 0, hi4 = 0;
                for (int s = 0; s < 8; s++)
                {
                    const unsigned m = ((s < 4 ? n1.z : n1.w) >> (8 * (s & 3))) & 0xffu;
                    if (!m) continue;
                    nc++, (s < 4 ? lo4 : hi4)++;
                    if (m == 0x80u) nin++;
                }
                hc[nc]++, hi_[nin]++, hl[nc - nin]++;
                halves[(lo4 ? 1 : 0) | (hi4 ? 2 : 0)]++;
            }
            fprintf(stderr, "children/node:");
            for (int k = 0; k <= 8; k++) fprintf(stderr, " %d:%.3f", k, double(hc[k]) / stats.n_nodes);
            fprintf(stderr, "\ninner/node:");
            for (int k = 0; k <= 8; k++) fprintf(stderr, " %d:%.3f", k, double(hi_[k]) / stats.n_nodes);
            fprintf(stderr, "\nleaves/node:");
            for (int k = 0; k <= 8; k++) fprintf(stderr, " %d:%.3f", k, double(hl[k]) / stats.n_nodes);
            fprintf(stderr, "\n");
        }
#else
        CRB_CUDA_CHECK(cudaEventRecord(ev1, stream));
        CRB_CUDA_CHECK(cudaEventSynchronize(ev1));
        float ms = 0;
        CRB_CUDA_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
        stats.build_ms = ms;
        cudaEventDestroy(ev0), cudaEventDestroy(ev1);
#endif
    }
}    // namespace crb
