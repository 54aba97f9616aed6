// bvh_treelet.inl — K3: SAH treelet restructuring of the binary LBVH (included by bvh_build.cu, inside
// namespace crb::<anon>).
//
// After Karras & Aila, "Fast Parallel Construction of High-Quality Bounding Volume Hierarchies" (HPG
// 2013), as published: walk the tree bottom-up with per-node arrival counters (same scheme as k_refit);
// at every inner node whose subtree holds at least TREELET_GAMMA triangles, grow a treelet of 7 leaves by
// repeatedly opening the treelet leaf with the largest surface area, find the SAH-optimal binary topology
// over those 7 leaves with dynamic programming over all 2^7 subsets, and re-link the treelet's 6 inner
// nodes in place if that is cheaper. One WARP per treelet (the DP tables live in shared memory, the subsets
// are spread over the lanes); the bottom-up order guarantees that nobody else touches the treelet's nodes
// while it is rewritten.
// The code is original.

constexpr int TREELET_N     = 7;
constexpr int TREELET_GAMMA = 7;    // minimum subtree size for a treelet root

__device__ __forceinline__ float ref_area(const BinTree &t, const float4 *plo, const float4 *phi, const uint32_t *vals, int ref)
{
    if (ref >= 0) return t.hi[ref].w;
    const uint32_t p = vals[~ref];
    return box_area(plo[p], phi[p]);
}

constexpr int TREELET_BLOCK = 128;    // threads per CTA of k_treelet

// Per-warp scratch of the cooperative treelet optimiser.
struct TreeletScratch
{
    float         area[128], copt[128];
    int           cnt[128];
    unsigned char part[128];
    float4        llo[TREELET_N], lhi[TREELET_N];
    float         lcost[TREELET_N];
    int           lcnt[TREELET_N], leaf[TREELET_N], inner[TREELET_N - 1];
    int           nl;
};

// All lanes of the warp restructure the treelet rooted at `node` (warp-uniform). The leader lane forms the
// treelet and re-links it; the subset tables (2^7 areas, optimal costs, best partitions) are spread over the
// lanes, the dynamic programme runs in rounds of equal subset size (a subset only needs smaller ones).
__device__ __forceinline__ void treelet_cooperative(int node, bool leader, unsigned lane, TreeletScratch &w, BinTree &t, const float4 *plo,
                                                    const float4 *phi, const uint32_t *vals, unsigned int *n_changed)
{
    const unsigned FULL = 0xffffffffu;
    if (leader)
    {
        // ---- form the treelet: 7 leaves (refs), 5 inner nodes besides the root
        int nl = 2, ni = 1;
        w.inner[0] = node;
        w.leaf[0] = t.left[node], w.leaf[1] = t.right[node];
        while (nl < TREELET_N)
        {
            int   best  = -1;
            float besta = -1.f;
            for (int k = 0; k < nl; k++)
                if (w.leaf[k] >= 0)
                {
                    const float a = t.hi[w.leaf[k]].w;
                    if (a > besta) besta = a, best = k;
                }
            if (best < 0) break;
            const int r   = w.leaf[best];
            w.inner[ni++] = r;
            w.leaf[best]  = t.left[r];
            w.leaf[nl++]  = t.right[r];
        }
        w.nl = nl;
    }
    __syncwarp(FULL);
    if (w.nl != TREELET_N) return;    // warp-uniform
    for (int k = int(lane); k < TREELET_N; k += CRB_WARP)
    {
        child_box(t, plo, phi, vals, w.leaf[k], w.llo[k], w.lhi[k], w.lcost[k]);
        w.lcnt[k] = w.leaf[k] < 0 ? 1 : t.count[w.leaf[k]];
    }
    __syncwarp(FULL);
    // ---- surface area and triangle count of every subset; singletons are the base case of the DP
    for (int s = int(lane); s < 128; s += CRB_WARP)
    {
        if (s == 0) continue;
        float lx = 3e38f, ly = 3e38f, lz = 3e38f, hx = -3e38f, hy = -3e38f, hz = -3e38f;
        int   cnt = 0;
        for (int k = 0; k < TREELET_N; k++)
            if (s & (1 << k))
            {
                lx = fminf(lx, w.llo[k].x), ly = fminf(ly, w.llo[k].y), lz = fminf(lz, w.llo[k].z);
                hx = fmaxf(hx, w.lhi[k].x), hy = fmaxf(hy, w.lhi[k].y), hz = fmaxf(hz, w.lhi[k].z);
                cnt += w.lcnt[k];
            }
        const float ex = hx - lx, ey = hy - ly, ez = hz - lz;
        w.area[s] = 2.0f * (ex * ey + ey * ez + ez * ex);
        w.cnt[s]  = cnt;
        if ((s & (s - 1)) == 0)
        {
            w.copt[s] = w.lcost[__ffs(s) - 1];
            w.part[s] = 0;
        }
    }
    __syncwarp(FULL);
    // ---- optimal cost of every subset, by subset size
    for (int size = 2; size <= TREELET_N; size++)
    {
        for (int s = int(lane); s < 128; s += CRB_WARP)
        {
            if (__popc(unsigned(s)) != size) continue;
            float     best  = 3e38f;
            int       bestp = 0;
            const int delta = (s - 1) & s;    // s without its lowest bit: enumerate each split once
            int       p     = (-delta) & s;
            do {
                const float c = w.copt[p] + w.copt[s ^ p];
                if (c < best) best = c, bestp = p;
                p = (p - delta) & s;
            } while (p != 0);
            float       c     = SAH_CI * w.area[s] + best;
            const float cleaf = SAH_CT * w.area[s] * float(w.cnt[s]);
            if (w.cnt[s] <= BVH8_LEAF_TRIS && cleaf < c) c = cleaf;    // same leaf rule as k_refit
            w.copt[s] = c;
            w.part[s] = (unsigned char) bestp;
        }
        __syncwarp(FULL);
    }
    // ---- rebuild if it pays
    if (leader && w.copt[127] < t.lo[node].w * 0.9999f)
    {
        atomicAdd(n_changed, 1u);
        int used = 1;    // inner[0] stays the root
        int stack_s[TREELET_N], stack_n[TREELET_N], sp = 0;
        stack_s[sp] = 127, stack_n[sp] = node, sp++;
        const int root_parent = t.parent[node];
        while (sp)
        {
            --sp;
            const int s = stack_s[sp], nd = stack_n[sp];
            const int ps[2] = { int(w.part[s]), s ^ int(w.part[s]) };
            int       refs[2];
            for (int c = 0; c < 2; c++)
            {
                const int q = ps[c];
                if ((q & (q - 1)) == 0)
                {
                    refs[c] = w.leaf[__ffs(q) - 1];
                    if (refs[c] < 0) t.leaf_parent[~refs[c]] = nd; else t.parent[refs[c]] = nd;
                }
                else
                {
                    refs[c]           = w.inner[used++];
                    t.parent[refs[c]] = nd;
                    stack_s[sp] = q, stack_n[sp] = refs[c], sp++;
                }
            }
            t.left[nd] = refs[0], t.right[nd] = refs[1];
            float lx = 3e38f, ly = 3e38f, lz = 3e38f, hx = -3e38f, hy = -3e38f, hz = -3e38f;
            for (int k = 0; k < TREELET_N; k++)
                if (s & (1 << k))
                {
                    lx = fminf(lx, w.llo[k].x), ly = fminf(ly, w.llo[k].y), lz = fminf(lz, w.llo[k].z);
                    hx = fmaxf(hx, w.lhi[k].x), hy = fmaxf(hy, w.lhi[k].y), hz = fmaxf(hz, w.lhi[k].z);
                }
            t.lo[nd]    = make_float4(lx, ly, lz, w.copt[s]);
            t.hi[nd]    = make_float4(hx, hy, hz, w.area[s]);
            t.count[nd] = w.cnt[s];
        }
        t.parent[node] = root_parent;
    }
    __syncwarp(FULL);    // the scratch is reused by the next treelet of this warp
}

// One thread per leaf walks up with arrival counters (as k_refit); the lanes of a warp that have reached a
// treelet root are served one after the other by the whole warp.
__global__ void __launch_bounds__(TREELET_BLOCK) k_treelet(int n, BinTree t, const float4 *__restrict__ plo, const float4 *__restrict__ phi,
                                                           const uint32_t *__restrict__ vals, unsigned int *n_changed)
{
    __shared__ TreeletScratch scratch[TREELET_BLOCK / CRB_WARP > 0 ? (CRB_WARP == 1 ? 1 : TREELET_BLOCK / CRB_WARP) : 1];
    const unsigned  FULL = 0xffffffffu;
    const unsigned  lane = crb_lane_id();
    TreeletScratch &w    = scratch[CRB_WARP == 1 ? 0 : threadIdx.x / CRB_WARP];
    const int       i    = blockIdx.x * blockDim.x + threadIdx.x;
    int             node = i < n ? t.leaf_parent[i] : -1;
    for (;;)
    {
        // climb: the second thread to arrive at an inner node owns it
        bool owner = false;
        if (node >= 0)
        {
            __threadfence();
            if (atomicAdd(t.flags + node, 1) == 0)
                node = -1;
            else
            {
                __threadfence();
                owner = true;
            }
        }
        if (__ballot_sync(FULL, owner) == 0u) break;    // every lane of the warp has stopped
        unsigned todo = __ballot_sync(FULL, owner && t.count[node] >= TREELET_GAMMA);
        while (todo)
        {
            const int leader = __ffs(int(todo)) - 1;
            todo &= todo - 1;
            const int root = __shfl_sync(FULL, node, leader);
            treelet_cooperative(root, int(lane) == leader, lane, w, t, plo, phi, vals, n_changed);
        }
        if (owner)
        {
            // refresh this node's cost from its (possibly restructured) children; the box is unchanged
            float4 alo, ahi, blo, bhi;
            float  ca, cb;
            child_box(t, plo, phi, vals, t.left[node], alo, ahi, ca);
            child_box(t, plo, phi, vals, t.right[node], blo, bhi, cb);
            float4      lo    = t.lo[node];
            const float area  = t.hi[node].w;
            const int   cnt   = t.count[node];
            float       cost  = SAH_CI * area + ca + cb;
            const float cleaf = SAH_CT * area * float(cnt);
            if (cnt <= BVH8_LEAF_TRIS && cleaf < cost) cost = cleaf;
            lo.w       = cost;
            t.lo[node] = lo;
            node       = t.parent[node];
        }
    }
}

// Runs `passes` bottom-up restructuring sweeps over the whole tree.
static void treelet_optimize(BinTree &t, int n, const float4 *plo, const float4 *phi, const uint32_t *vals, cudaStream_t stream, int passes,
                             unsigned int *d_changed)
{
    const int B = TREELET_BLOCK;
    for (int pass = 0; pass < passes; pass++)
    {
        dev_zero(t.flags, size_t(n - 1) * sizeof(int), stream);
        CRB_LAUNCH(k_treelet, (unsigned(n) + B - 1) / B, B, stream, n, t, plo, phi, vals, d_changed);
    }
}
