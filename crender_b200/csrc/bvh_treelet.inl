// bvh_treelet.inl — K3: SAH treelet restructuring of the binary LBVH (included by bvh_build.cu, inside
// namespace crb::<anon>).
//
// After Karras & Aila, "Fast Parallel Construction of High-Quality Bounding Volume Hierarchies" (HPG
// 2013), as published: walk the tree bottom-up with per-node arrival counters (same scheme as k_refit);
// at every inner node whose subtree holds at least TREELET_GAMMA triangles, grow a treelet of 7 leaves by
// repeatedly opening the treelet leaf with the largest surface area, find the SAH-optimal binary topology
// over those 7 leaves with dynamic programming over all 2^7 subsets, and re-link the treelet's 6 inner
// nodes in place if that is cheaper. One thread per treelet (the DP tables live in local memory); the
// bottom-up order guarantees that nobody else touches the treelet's nodes while it is rewritten.
// The code is original.

constexpr int TREELET_N     = 7;
constexpr int TREELET_GAMMA = 7;    // minimum subtree size for a treelet root

__device__ __forceinline__ float ref_area(const BinTree &t, const float4 *plo, const float4 *phi, const uint32_t *vals, int ref)
{
    if (ref >= 0) return t.hi[ref].w;
    const uint32_t p = vals[~ref];
    return box_area(plo[p], phi[p]);
}

__global__ void k_treelet(int n, BinTree t, const float4 *__restrict__ plo, const float4 *__restrict__ phi, const uint32_t *__restrict__ vals,
                          unsigned int *n_changed)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = t.leaf_parent[i];
    while (node >= 0)
    {
        __threadfence();
        if (atomicAdd(t.flags + node, 1) == 0) return;
        __threadfence();

        if (t.count[node] >= TREELET_GAMMA)
        {
            // ---- form the treelet: 7 leaves (refs), 5 inner nodes besides the root
            int leaf[TREELET_N], inner[TREELET_N - 1];
            int nl = 2, ni = 1;
            inner[0] = node;
            leaf[0] = t.left[node], leaf[1] = t.right[node];
            while (nl < TREELET_N)
            {
                int   best  = -1;
                float besta = -1.f;
                for (int k = 0; k < nl; k++)
                    if (leaf[k] >= 0)
                    {
                        const float a = t.hi[leaf[k]].w;
                        if (a > besta) besta = a, best = k;
                    }
                if (best < 0) break;
                const int r  = leaf[best];
                inner[ni++]  = r;
                leaf[best]   = t.left[r];
                leaf[nl++]   = t.right[r];
            }
            if (nl == TREELET_N)
            {
                float4 llo[TREELET_N], lhi[TREELET_N];
                float  lcost[TREELET_N];
                int    lcnt[TREELET_N];
                for (int k = 0; k < TREELET_N; k++)
                {
                    child_box(t, plo, phi, vals, leaf[k], llo[k], lhi[k], lcost[k]);
                    lcnt[k] = leaf[k] < 0 ? 1 : t.count[leaf[k]];
                }
                // ---- surface area of every subset
                float         area[128];
                float         copt[128];
                unsigned char part[128];
                for (int s = 1; s < 128; s++)
                {
                    float lx = 3e38f, ly = 3e38f, lz = 3e38f, hx = -3e38f, hy = -3e38f, hz = -3e38f;
                    for (int k = 0; k < TREELET_N; k++)
                        if (s & (1 << k))
                        {
                            lx = fminf(lx, llo[k].x), ly = fminf(ly, llo[k].y), lz = fminf(lz, llo[k].z);
                            hx = fmaxf(hx, lhi[k].x), hy = fmaxf(hy, lhi[k].y), hz = fmaxf(hz, lhi[k].z);
                        }
                    const float ex = hx - lx, ey = hy - ly, ez = hz - lz;
                    area[s]        = 2.0f * (ex * ey + ey * ez + ez * ex);
                }
                // ---- optimal cost of every subset (any proper subset is numerically smaller)
                for (int s = 1; s < 128; s++)
                {
                    if ((s & (s - 1)) == 0)
                    {
                        copt[s] = lcost[__ffs(s) - 1];
                        part[s] = 0;
                        continue;
                    }
                    float     best  = 3e38f;
                    int       bestp = 0;
                    const int delta = (s - 1) & s;    // s without its lowest bit: enumerate each split once
                    int       p     = (-delta) & s;
                    do {
                        const float c = copt[p] + copt[s ^ p];
                        if (c < best) best = c, bestp = p;
                        p = (p - delta) & s;
                    } while (p != 0);
                    int cnt = 0;
                    for (int k = 0; k < TREELET_N; k++)
                        if (s & (1 << k)) cnt += lcnt[k];
                    float       c     = SAH_CI * area[s] + best;
                    const float cleaf = SAH_CT * area[s] * float(cnt);
                    if (cnt <= BVH8_LEAF_TRIS && cleaf < c) c = cleaf;    // same leaf rule as k_refit
                    copt[s] = c;
                    part[s] = (unsigned char) bestp;
                }
                // ---- rebuild if it pays
                if (copt[127] < t.lo[node].w * 0.9999f)
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
                        const int ps[2] = { int(part[s]), s ^ int(part[s]) };
                        int       refs[2];
                        for (int c = 0; c < 2; c++)
                        {
                            const int q = ps[c];
                            if ((q & (q - 1)) == 0)
                            {
                                refs[c] = leaf[__ffs(q) - 1];
                                if (refs[c] < 0) t.leaf_parent[~refs[c]] = nd; else t.parent[refs[c]] = nd;
                            }
                            else
                            {
                                refs[c]            = inner[used++];
                                t.parent[refs[c]]  = nd;
                                stack_s[sp] = q, stack_n[sp] = refs[c], sp++;
                            }
                        }
                        t.left[nd] = refs[0], t.right[nd] = refs[1];
                        float lx = 3e38f, ly = 3e38f, lz = 3e38f, hx = -3e38f, hy = -3e38f, hz = -3e38f;
                        int   cnt = 0;
                        for (int k = 0; k < TREELET_N; k++)
                            if (s & (1 << k))
                            {
                                lx = fminf(lx, llo[k].x), ly = fminf(ly, llo[k].y), lz = fminf(lz, llo[k].z);
                                hx = fmaxf(hx, lhi[k].x), hy = fmaxf(hy, lhi[k].y), hz = fmaxf(hz, lhi[k].z);
                                cnt += lcnt[k];
                            }
                        t.lo[nd]    = make_float4(lx, ly, lz, copt[s]);
                        t.hi[nd]    = make_float4(hx, hy, hz, area[s]);
                        t.count[nd] = cnt;
                    }
                    t.parent[node] = root_parent;
                }
            }
        }
        else
        {
            // small subtree: nothing to restructure, but its cost must reflect restructured children
        }
        // refresh this node's cost from its (possibly restructured) children; the box is unchanged
        {
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
        }
        node = t.parent[node];
    }
}

// Runs `passes` bottom-up restructuring sweeps over the whole tree.
static void treelet_optimize(BinTree &t, int n, const float4 *plo, const float4 *phi, const uint32_t *vals, cudaStream_t stream, int passes,
                             unsigned int *d_changed)
{
    const int B = 128;
    for (int pass = 0; pass < passes; pass++)
    {
        dev_zero(t.flags, size_t(n - 1) * sizeof(int), stream);
        CRB_LAUNCH(k_treelet, (unsigned(n) + B - 1) / B, B, stream, n, t, plo, phi, vals, d_changed);
    }
}
