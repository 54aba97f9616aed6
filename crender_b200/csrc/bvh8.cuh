// bvh8.cuh — the 8-wide compressed BVH: node layout and per-ray traversal.
//
// Replaces the opaque Embree BVH built by rtcCommitScene (src/objects/model.cpp:92-94) and the
// rtcIntersect1 query (src/objects/model.cpp:27). B200 has no RT cores; this is a software BVH laid
// out for 16-byte vector loads from L2/HBM:
//
//   node  = 80 bytes in a 32-byte-aligned 96-byte slot = three 256-bit loads (CRB_NODE_U4; five 16-byte words used):
//     n0: origin.xyz (f32) | ex,ey,ez (biased power-of-two exponents) , imask (bit s: slot s is an inner node)
//     n1: child_base (index of first inner child, < 2^24) | tri_base (index of first leaf triangle) | meta[8]
//     n2: qlo.x[8] | qlo.y[8]        8-bit child boxes on the grid origin + q * 2^(e-127)
//     n3: qlo.z[8] | qhi.x[8]
//     n4: qhi.y[8] | qhi.z[8]
//     meta[s]: 0 = empty; inner: 0x80; leaf: unary triangle count 1/3/7
//   tri   = 48 bytes = three float4 loads: (v0, flat prim id) (e1=v1-v0, 0) (e2=v2-v0, 0)
//   A node's leaf triangles are stored from tri_base in NIBBLE ORDER of the slots (0,4,1,5,2,6,3,7): the node's
//   leaf occupancy word occ = ((meta[4..7] << 4) | meta[0..3]) & 0x77777777 has one bit per triangle, and the
//   triangle behind bit i is tri_base + popc(occ & ((1 << i) - 1)). Inner children are stored from child_base in
//   slot order (rank = popc(imask & ((1 << slot) - 1))).
//
// Children are assigned to slots at build time so that slot ^ (7 ^ ray octant) orders them front to
// back; a node's hit children are kept as one 8-byte stack entry (base index | imask, hit flags), following the
// compressed-wide-BVH scheme of Ylitie, Karras & Laine (HPG 2017) as published; the code is original.
//
// Hit-word assembly (round 2, see node_visit): ncu/SASS showed the node test bound by the ALU pipe, the conversion
// unit and the issue slots at once (136 ALU + 34 XU + 85 FMA of 263 instructions; ALU and FMA issue every other
// cycle per scheduler, the conversion unit every eighth). The per-child `bits = count << index` assembly of round 1
// (61 ALU instructions) is replaced by byte-parallel work on ALL children at once: the eight slab results become
// eight sign bits (one FFMA each), six PRMTs with sign replication turn them into two words of 0x00/0xff bytes,
// two LOP3 mask the meta bytes with them, two PRMTs with a per-ray selector put the bytes in front-to-back order
// and four shift/mask pairs split them into the inner-flag word, the leaf-triangle word and the occupancy word.
//
// The triangle test is Moeller-Trumbore with a fixed operation sequence of single roundings (see
// tri_test) so that t,u,v are bit-identical to the CPU oracle; ties in t go to the lowest flat
// primitive id, which makes closest-hit results independent of the tree and of traversal order.
#pragma once
#include "platform.cuh"
#include "vecmath.cuh"

// compile-time variants of the traversal loop (measured in profiles/r1g_sweeps.md)
#ifndef CRB_TRI_BRANCHFREE
#define CRB_TRI_BRANCHFREE 1    // +3.1 % (profiles/r1g_sweeps.md section 6)
#endif
#ifndef CRB_EARLY_POP
#define CRB_EARLY_POP 2    // 1: +0.7 %, 2 (predicated in-place loads): +4.3 % (profiles/r1g_sweeps.md section 6)
#endif

// Per-lane state of the persistent trace loop that is touched rarely (the direction: leaf phase only; the winner's
// barycentrics and the work item: at a hit / at retirement) lives in shared memory (CTAs of at most 256 threads, 8 KB):
// the loop then fits 56 registers, i.e. 9 instead of 8 warps per scheduler (CTAs of 128 threads, 9 per SM).
#ifndef CRB_TP_SMEM
#define CRB_TP_SMEM 1
#endif
// Any-hit queries visit the hit children of a node BACK to front. A shadow ray starts on a surface and leaves it: the boxes
// around its origin are the ones least likely to hold an occluder (the ray grazes its own neighbourhood), yet front-to-back
// order visits them first. On config 2, 71 % of the sun's shadow rays are occluded (NEE at every hit: inside the glass, in
// the sphere's shadow) and needed 18.2 node visits each to find an occluder front to back, 11.7 back to front: 13.9 -> 9.0
// node visits and 4.3 -> 2.8 triangle tests per shadow query (kernel-logic harness, tools/tree_quality.py). The answer of an
// any-hit query does not depend on the order.
#ifndef CRB_ANY_BACK_FIRST
#define CRB_ANY_BACK_FIRST 1
#endif

// Node slot in memory, in 16-byte words: 5 = the packed 80-byte record read by five 128-bit loads; 6 = the same record
// in a 32-byte-aligned 96-byte slot read by THREE 256-bit loads (sm_100's LDG.E.ENL2.256). The secondary-ray launches of
// the trace loop keep the L1 data pipe busier than the issue slots (l1tex__data_pipe_lsu_wavefronts 92 % against 73 %,
// profiles/r2p_k_trace.md): its load is one wavefront per distinct line of every load instruction of a divergent warp plus
// one per sector filled after a miss. Fewer, wider loads per node: 59.6 M -> 45.5 M load requests and 340 M -> 243 M
// wavefronts per launch (the fills stay: 0.43 G sectors, the algorithmic bytes), the pipe at 87 %; an aligned 96-byte slot
// also touches exactly three 32-byte sectors where an 80-byte record at a 16-byte boundary touches 3 or 4. Measured
// +1.5 % on config 2 (profiles/r2_sweeps.md section 14).
#ifndef CRB_NODE_U4
#define CRB_NODE_U4 6
#endif
// Triangle slot, in 16-byte words: 3 = the packed 48-byte record (three 128-bit loads); 4 = a 64-byte slot read by two
// 256-bit loads (same reasoning; +16 bytes per triangle of L2 footprint: measured 0.9 % slower, profiles/r2_sweeps.md
// section 14).
#ifndef CRB_TRI_F4
#define CRB_TRI_F4 3
#endif

namespace crb
{
    constexpr int      BVH8_STACK      = 48;     // entries; the builder rejects deeper trees loudly
    constexpr int      BVH8_LEAF_TRIS  = 3;      // max triangles per leaf child
    constexpr uint32_t INVALID_PRIM    = 0xffffffffu;
    constexpr float    BVH8_BOX_SLACK  = 1.000001f;    // relative loosening of the slab exit distance
    constexpr int      BVH8_NODE_U4    = CRB_NODE_U4;  // uint4 words per node slot (5 packed, 6 = 32-byte aligned)
    static_assert(BVH8_NODE_U4 == 5 || BVH8_NODE_U4 == 6, "node slot is 80 or 96 bytes");
    constexpr int      BVH8_TRI_F4     = CRB_TRI_F4;   // float4 words per triangle slot (3 packed, 4 = 64-byte aligned)
    static_assert(BVH8_TRI_F4 == 3 || BVH8_TRI_F4 == 4, "triangle slot is 48 or 64 bytes");

    struct Bvh8
    {
        const uint4  *nodes;    // BVH8_NODE_U4 per node (the record is the first five)
        const float4 *tris;     // BVH8_TRI_F4 per triangle (the record is the first three)
        uint32_t      n_nodes;
        uint32_t      n_tris;
    };

    struct Hit
    {
        float    t, u, v;
        uint32_t prim;
    };

    struct TravCounters
    {
        unsigned long long nodes = 0, tris = 0;
    };

    // Moeller-Trumbore; operation sequence is part of the parity contract (oracle.cpp tri_test):
    //   p = d x e2, det = e1.p, inv = 1/det, s = o - v0, u = (s.p)*inv, q = s x e1, v = (d.q)*inv,
    //   t = (e2.q)*inv; accept iff 0<=u<=1, v>=0, u+v<=1, tnear < t <= tfar (Embree's convention).
    __device__ __forceinline__ bool tri_test(V3 v0, V3 e1, V3 e2, V3 o, V3 d, float tnear, float tfar, float &t, float &u, float &v)
    {
#if CRB_TRI_BRANCHFREE
        // Same operations and roundings, no early exits: in the lock-step leaf phase some lane always runs the whole
        // test, so the exits only cost branch overhead — and ptxas sank the v0 load below the det branch, i.e. a second
        // serialised memory wait per leaf test (ncu: 13 % + 5 % of k_trace's stall samples on the two waits).
        // det == 0 gives inv = inf and NaN/inf in u, v, t; the det test is part of the final predicate.
        const V3    p   = cross(d, e2);
        const float det = dot(e1, p);
        const float inv = __frcp_rn(det);
        const V3    s   = o - v0;
        u               = __fmul_rn(dot(s, p), inv);
        const V3 q      = cross(s, e1);
        v               = __fmul_rn(dot(d, q), inv);
        t               = __fmul_rn(dot(e2, q), inv);
        return (det != 0.0f) & (u >= 0.0f) & (u <= 1.0f) & (v >= 0.0f) & (__fadd_rn(u, v) <= 1.0f) & (t > tnear) & (t <= tfar);
#else
        const V3    p   = cross(d, e2);
        const float det = dot(e1, p);
        if (det == 0.0f) return false;
        const float inv = __frcp_rn(det);
        const V3    s   = o - v0;
        u               = __fmul_rn(dot(s, p), inv);
        if (!(u >= 0.0f && u <= 1.0f)) return false;
        const V3 q = cross(s, e1);
        v          = __fmul_rn(dot(d, q), inv);
        if (!(v >= 0.0f && __fadd_rn(u, v) <= 1.0f)) return false;
        t = __fmul_rn(dot(e2, q), inv);
        return t > tnear && t <= tfar;
#endif
    }

    __device__ __forceinline__ float safe_rcp(float d)
    {
        const float tiny = 1e-20f;
        return 1.0f / (fabsf(d) > tiny ? d : copysignf(tiny, d));
    }

    __device__ __forceinline__ unsigned byte_of(unsigned w, int i) { return (w >> (8 * i)) & 0xffu; }
// Of the 48 byte->float conversions of a node test, the 16 of CRB_NODE_PRMT_AXES axes go through PRMT + a folded
// slab FMA on the ALU pipe instead of I2F on the conversion unit (4 lanes/clk/SMSP, the busiest pipe of k_trace at
// 65 %). Measured on config 2 (profiles/r1g_sweeps.md): 0 axes 2883, 1 axis 2992-3013, 2 axes 2970-2999, 3 axes
// 2944-2949 Mrays/s: one axis balances the two pipes.
#ifndef CRB_NODE_PRMT_AXES
#define CRB_NODE_PRMT_AXES 1
#endif
    // 1 + q * 2^-15 for byte i of w: the byte lands in bits 8..15 of the bit pattern of 1.0f (one PRMT)
    // `one` must hold 0x3f800000 in a REGISTER the compiler cannot fold (see opaque_one): SASS PRMT has one
    // immediate slot, and with a literal constant ptxas spends it on the constant and re-materialises the four
    // selectors from uniform registers before every PRMT (+48 moves per node)
    __device__ __forceinline__ float byte_as_float(unsigned w, int i, unsigned one)
    {
#ifdef CRB_EMU
        return __uint_as_float(one | (byte_of(w, i) << 8));
#else
        return __uint_as_float(__byte_perm(w, one, 0x7604u | (unsigned(i) << 4)));
#endif
    }

    // 16-byte read-only load that the compiler may not sink below later branches: the three loads of a
    // triangle record must be in flight together (ncu showed the v0 load issued after the det test,
    // i.e. two serialised L2 round trips per leaf test)
    __device__ __forceinline__ float4 ldg128_pinned(const float4 *p)
    {
#ifdef CRB_EMU
        return __ldg(p);
#else
        float4 v;
        asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
        return v;
#endif
    }

    // The three words of triangle `index` (see CRB_TRI_F4); the loads stay together like ldg128_pinned's.
    __device__ __forceinline__ void load_tri(const float4 *tris, size_t index, float4 &a, float4 &b, float4 &c)
    {
        const float4 *tp = tris + index * BVH8_TRI_F4;
#if CRB_TRI_F4 == 4 && !defined(CRB_EMU)
        float p0, p1, p2, p3;
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                     : "l"(tp));
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w), "=f"(p0), "=f"(p1), "=f"(p2), "=f"(p3)
                     : "l"(tp + 2));
#else
        a = ldg128_pinned(tp), b = ldg128_pinned(tp + 1), c = ldg128_pinned(tp + 2);
#endif
    }

    // The five words of node `index` (see CRB_NODE_U4).
    __device__ __forceinline__ void load_node(const uint4 *nodes, size_t index, uint4 &n0, uint4 &n1, uint4 &n2, uint4 &n3, uint4 &n4)
    {
        const uint4 *np = nodes + index * BVH8_NODE_U4;
#if CRB_NODE_U4 == 6 && !defined(CRB_EMU)
        unsigned pad0, pad1, pad2, pad3;
        asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=r"(n0.x), "=r"(n0.y), "=r"(n0.z), "=r"(n0.w), "=r"(n1.x), "=r"(n1.y), "=r"(n1.z), "=r"(n1.w)
            : "l"(np));
        asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=r"(n2.x), "=r"(n2.y), "=r"(n2.z), "=r"(n2.w), "=r"(n3.x), "=r"(n3.y), "=r"(n3.z), "=r"(n3.w)
            : "l"(np + 2));
        asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=r"(n4.x), "=r"(n4.y), "=r"(n4.z), "=r"(n4.w), "=r"(pad0), "=r"(pad1), "=r"(pad2), "=r"(pad3)
            : "l"(np + 4));
#else
        n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
#endif
    }

    // PTX prmt.b32 in its default mode: result byte i = byte (sel nibble i & 7) of {b,a}; nibble bit 3 set = the byte's
    // sign bit replicated over all 8 bits. (__byte_perm only exposes the 3-bit form.)
    __device__ __forceinline__ unsigned crb_prmt(unsigned a, unsigned b, unsigned sel)
    {
#ifdef CRB_EMU
        const unsigned long long src = ((unsigned long long) b << 32) | a;
        unsigned                 r   = 0;
        for (int i = 0; i < 4; i++)
        {
            const unsigned n    = (sel >> (4 * i)) & 0xfu;
            unsigned       byte = unsigned(src >> (8 * (n & 7u))) & 0xffu;
            if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
            r |= byte << (8 * i);
        }
        return r;
#else
        unsigned d;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
        return d;
#endif
    }

    // bit pattern whose SIGN BIT says "slab test failed": tf * slack - tn, one FFMA (on the GPU an inf - inf gives the
    // canonical positive NaN = hit, conservative; the harness mirrors that, x86 would produce a negative NaN)
    __device__ __forceinline__ unsigned slab_miss_sign(float tn, float tf)
    {
        const float diff = fmaf(tf, BVH8_BOX_SLACK, -tn);
#ifdef CRB_EMU
        if (diff != diff) return 0u;
#endif
        return __float_as_uint(diff);
    }

    // per-ray octant word: octinv = 7 ^ octant replicated in four nibbles (the low nibble is octinv itself). The PRMT
    // selectors that put a node's child bytes in front-to-back order are 0x7531 ^ oct4 and 0x6420 ^ oct4.
    __device__ __forceinline__ unsigned make_oct4(V3 d)
    {
        return (7u ^ ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u))) * 0x1111u;
    }

    // What a node visit leaves behind (see the header comment):
    //   group  = (child_base | imask << 24, inner-hit flags: bit 4p+3 = the hit inner child of front-to-back priority p)
    //   tgroup = (tri_base, hit leaf triangles: bits of occ)
    //   occ    = the node's leaf occupancy word (all its leaf triangles, hit or not): triangle of bit i = tri_base +
    //            popc(occ & ((1 << i) - 1))
    // Intersects the ray with the 8 child boxes of one node.
    __device__ __forceinline__ void node_visit(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4, V3 o, V3 idir, unsigned oct4,
                                               float tmin, float tmax, uint2 &group, uint2 &tgroup, unsigned &occ)
    {
        const float px = __uint_as_float(n0.x), py = __uint_as_float(n0.y), pz = __uint_as_float(n0.z);
        const float sx = __uint_as_float((n0.w & 0xffu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xffu) << 23),
                    sz = __uint_as_float(((n0.w >> 16) & 0xffu) << 23);
        const float    ax = sx * idir.x, ay = sy * idir.y, az = sz * idir.z;
        const float    bx = (px - o.x) * idir.x, by = (py - o.y) * idir.y, bz = (pz - o.z) * idir.z;
        const bool     nx = idir.x < 0.0f, ny = idir.y < 0.0f, nz = idir.z < 0.0f;
#if CRB_NODE_PRMT_AXES
        const unsigned one = 0x3f800000u ^ (n1.x >> 31);    // child_base < 2^24: always 1.0f, but not a literal (see byte_as_float)
        // folded slab constants; the offset b - a*2^15 is rounded once (<= |a| * 2^-9, i.e. 1/512 of a grid cell),
        // so the entry side is loosened downwards and the exit side upwards by 1/256 of a cell
        const float Ax = ax * 32768.0f, Bnx = (bx - Ax) - fabsf(ax) * 0.00390625f, Bfx = (bx - Ax) + fabsf(ax) * 0.00390625f;
#if CRB_NODE_PRMT_AXES >= 2
        const float Ay = ay * 32768.0f, Bny = (by - Ay) - fabsf(ay) * 0.00390625f, Bfy = (by - Ay) + fabsf(ay) * 0.00390625f;
#endif
#if CRB_NODE_PRMT_AXES >= 3
        const float Az = az * 32768.0f, Bnz = (bz - Az) - fabsf(az) * 0.00390625f, Bfz = (bz - Az) + fabsf(az) * 0.00390625f;
#endif
#endif
        unsigned miss[2];    // per half: byte j = 0xff if child 4*half + j is missed
#pragma unroll
        for (int half = 0; half < 2; half++)
        {
            const unsigned lox = half ? n2.y : n2.x, loy = half ? n2.w : n2.z, loz = half ? n3.y : n3.x;
            const unsigned hix = half ? n3.w : n3.z, hiy = half ? n4.y : n4.x, hiz = half ? n4.w : n4.z;
            const unsigned nearx = nx ? hix : lox, farx = nx ? lox : hix;
            const unsigned neary = ny ? hiy : loy, fary = ny ? loy : hiy;
            const unsigned nearz = nz ? hiz : loz, farz = nz ? loz : hiz;
            unsigned       sgn[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                // byte -> float without the conversion pipe: PRMT drops the byte into mantissa bits 8..15 of 1.0f
                // (value 1 + q * 2^-15), and the slab FMA absorbs the offset and the scale: q*a + b ==
                // (1 + q*2^-15) * (a*2^15) + (b - a*2^15). Same instruction count as I2F + FFMA.
#if CRB_NODE_PRMT_AXES >= 1
                const float t0x = fmaf(byte_as_float(nearx, j, one), Ax, Bnx), t1x = fmaf(byte_as_float(farx, j, one), Ax, Bfx);
#else
                const float t0x = fmaf(float(byte_of(nearx, j)), ax, bx), t1x = fmaf(float(byte_of(farx, j)), ax, bx);
#endif
#if CRB_NODE_PRMT_AXES >= 2
                const float t0y = fmaf(byte_as_float(neary, j, one), Ay, Bny), t1y = fmaf(byte_as_float(fary, j, one), Ay, Bfy);
#else
                const float t0y = fmaf(float(byte_of(neary, j)), ay, by), t1y = fmaf(float(byte_of(fary, j)), ay, by);
#endif
#if CRB_NODE_PRMT_AXES >= 3
                const float t0z = fmaf(byte_as_float(nearz, j, one), Az, Bnz), t1z = fmaf(byte_as_float(farz, j, one), Az, Bfz);
#else
                const float t0z = fmaf(float(byte_of(nearz, j)), az, bz), t1z = fmaf(float(byte_of(farz, j)), az, bz);
#endif
                const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
                const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
                sgn[j]         = slab_miss_sign(tn, tf);    // hit <=> tn <= tf * BVH8_BOX_SLACK
            }
            // four sign bits -> four bytes 0x00 / 0xff (PRMT sign replication of the top byte of each difference)
            const unsigned m01 = crb_prmt(sgn[0], sgn[1], 0xfbfbu), m23 = crb_prmt(sgn[2], sgn[3], 0xfbfbu);
            miss[half]         = crb_prmt(m01, m23, 0x5410u);
        }
        // hit children keep their meta byte (inner 0x80, leaf unary count, empty 0)
        const unsigned hlo = n1.z & ~miss[0], hhi = n1.w & ~miss[1];
        // front-to-back order: byte k of pa = the child of priority 2k+1, of pb = priority 2k (priority p <-> slot p ^ octinv)
        const unsigned pa = crb_prmt(hlo, hhi, 0x7531u ^ oct4), pb = crb_prmt(hlo, hhi, 0x6420u ^ oct4);
        group             = make_uint2(n1.x | (n0.w & 0xff000000u), (pa | (pb >> 4)) & 0x88888888u);
        tgroup            = make_uint2(n1.y, ((hhi << 4) | hlo) & 0x77777777u);
        occ               = ((n1.w << 4) | n1.z) & 0x77777777u;
    }

    // the front-most unvisited inner child of a node group: removes it from the group, returns its node index
    // back = true takes the BACK-most one instead (any-hit queries, CRB_ANY_BACK_FIRST).
    __device__ __forceinline__ unsigned pop_inner(uint2 &group, unsigned oct4, bool back = false)
    {
        const int bit = back ? __ffs(int(group.y)) - 1 : 31 - __clz(int(group.y));
        group.y &= ~(1u << bit);
        const unsigned slot = (unsigned(bit >> 2) ^ oct4) & 7u;
        return (group.x & 0x00ffffffu) + __popc((group.x >> 24) & ((1u << slot) - 1u));
    }

    // the next waiting leaf triangle of a node: removes its bit, returns its index in the triangle array
    __device__ __forceinline__ unsigned pop_triangle(uint2 &tgroup, unsigned occ)
    {
        const int i = __ffs(int(tgroup.y)) - 1;
        tgroup.y &= tgroup.y - 1;
        return tgroup.x + __popc(occ & ((1u << i) - 1u));
    }

    // Closest hit (ANY=false) or any hit (ANY=true) along o + t*d, t in (tmin, tmax].
    template<bool ANY, bool COUNT>
    __device__ __forceinline__ Hit traverse(const Bvh8 &bvh, V3 o, V3 d, float tmin, float tmax, TravCounters *ctr)
    {
        Hit best { tmax, 0.0f, 0.0f, INVALID_PRIM };
        if (bvh.n_nodes == 0)
        {
            best.t = __int_as_float(0x7f800000);
            return best;
        }
        const V3       idir = v3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
        const unsigned oct4 = make_oct4(d);

        uint2 stack[BVH8_STACK];
        int   sp = 0;
        // node group: x = base node index | imask << 24, y = hit flags of the inner children (bit 4p+3, priority p). The
        // root is entered as the single hit child of a pseudo group with base 0 and an empty imask (rank 0).
        uint2    group = make_uint2(0u, 0x80000000u), tgroup;
        unsigned occ;

        for (;;)
        {
            // pop the front-most (any-hit: back-most) unvisited inner child of the current group
            const unsigned node_index = pop_inner(group, oct4, CRB_ANY_BACK_FIRST && ANY);
            if (group.y) stack[sp++] = group;

            uint4 n0, n1, n2, n3, n4;
            load_node(bvh.nodes, node_index, n0, n1, n2, n3, n4);
            if (COUNT) ctr->nodes++;
            node_visit(n0, n1, n2, n3, n4, o, idir, oct4, tmin, best.t, group, tgroup, occ);

            while (tgroup.y)
            {
                float4 a, b, c;
                load_tri(bvh.tris, pop_triangle(tgroup, occ), a, b, c);
                if (COUNT) ctr->tris++;
                float t, u, v;
                if (tri_test(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), o, d, tmin, best.t, t, u, v))
                {
                    const unsigned prim = __float_as_uint(a.w);
                    if (ANY) return Hit { t, u, v, prim };
                    if (t < best.t || prim < best.prim) best = Hit { t, u, v, prim };
                }
            }

            if (group.y == 0u)
            {
                if (sp == 0) break;
                group = stack[--sp];
            }
        }
        if (best.prim == INVALID_PRIM) best.t = __int_as_float(0x7f800000);
        return best;
    }

    // ------------------------------------------------------------------------------------------------
    // Persistent trace loop with per-lane dynamic refill.
    //
    // Incoherent rays have wildly different traversal lengths (a ray that misses the scene box needs one
    // node, a grazing ray 50+), so "one warp = 32 fixed rays" leaves most lanes idle while the longest ray
    // finishes (measured: 6 of 32 lanes active per issued instruction, round-1 history in DESIGN.md section 4). Here a
    // warp is a pool of 32 traversal lanes: whenever lanes are idle, the warp takes exactly that many new
    // work items from the global cursor with ONE atomic (warp-aggregated), and every lane keeps its own
    // traversal state in registers. Lanes reconverge every STEPS node iterations, where finished rays are
    // handed to `sink` (which may itself use warp-aggregated queue pushes) and idle lanes are refilled.
    //
    // Every iteration is a node phase (lanes without waiting triangles visit one node) followed by a leaf phase in
    // lock step (every lane with waiting triangles tests ONE of them). A warp-cooperative leaf phase (all waiting
    // triangles of the warp redistributed over the 32 lanes through shared memory) was built, kept bit-identical and
    // measured 12 % slower (profiles/r2_sweeps.md section 1); it is no longer in the source.
    //
    //   source(idx, item, o, d, tmin, tmax)  loads work item idx (called by the lane that owns it)
    //   sink(valid, item, hit)               called by ALL lanes at a convergent point; valid lanes retire
    // `any` (any-hit vs closest-hit) is a RUN-TIME argument on purpose: both query kinds execute the same
    // compiled loop, so the bit-exact closest-hit parity tests cover the code any-hit queries run.
    // OFFSETS = true: every work item is a query against ITS OWN tree inside the concatenated arrays (a BLAS of a two-level
    // scene): source(idx, item, o, d, tmin, tmax, node_off, tri_off) also returns the tree's node / triangle base (child and
    // triangle indices inside a tree are relative to it); tmax < 0 marks an item with nothing to traverse.
    template<bool COUNT, int STEPS, bool OFFSETS = false, typename Source, typename Sink>
    __device__ __forceinline__ void trace_persistent(const Bvh8 &bvh, uint32_t *cursor, uint32_t n, uint32_t chunk_max, bool any, Source source, Sink sink,
                                                     TravCounters *ctr)
    {
        const unsigned FULL = 0xffffffffu;
        const unsigned lane = crb_lane_id();
        uint2          stack[BVH8_STACK];
        int            sp = 0;
        bool           active = false, finished = false, exhausted = false;
#if CRB_TP_SMEM && !defined(CRB_EMU)
        __shared__ float4 s_dir[256], s_win[256];    // (d, -) and (u, v, work item, -) of every lane of the CTA
        float4 &r_dir = s_dir[threadIdx.x], &r_win = s_win[threadIdx.x];
#else
        float4 r_dir = make_float4(0, 0, 1, 0), r_win = make_float4(0, 0, 0, 0);
#endif
        V3             o = v3(0, 0, 0), idir = v3(0, 0, 0);
        unsigned       oct4 = 0, occ = 0;
        float          tmin = 0.f, best_t = 0.f;
        unsigned       best_prim = INVALID_PRIM;
        uint2          group = make_uint2(0u, 0u), tgroup = make_uint2(0u, 0u);
        uint32_t       local_next = 0, local_end = 0;
        uint32_t       node_off = 0;    // OFFSETS: base of the item's tree in bvh.nodes (its triangle base lives in r_dir.w)
        // reservation granularity: large enough to keep the cursor atomic rare, small enough that the last
        // ranges are spread over all warps of the grid
        const uint32_t total_warps = (gridDim.x * blockDim.x + CRB_WARP - 1) / CRB_WARP;
        uint32_t       chunk       = n / (total_warps * 4u);
        chunk                      = chunk < uint32_t(CRB_WARP) ? uint32_t(CRB_WARP) : (chunk > chunk_max ? chunk_max : chunk);
        if (chunk_max && chunk_max < uint32_t(CRB_WARP)) chunk = chunk_max;    // small fixed look-ahead (experiments)

        for (;;)
        {
            {
                const float4 w = r_win;
                sink(finished, __float_as_uint(w.z), Hit { best_t, w.x, w.y, best_prim });
            }
            finished = false;

            const unsigned idle = __ballot_sync(FULL, !active);
            if (idle != 0u && !exhausted)
            {
                // the warp owns a reserved range [local_next, local_end) of work items; a new range is
                // reserved with one atomic only when it runs dry (same-address atomics serialise in L2)
                const uint32_t need = uint32_t(__popc(idle));
                if (local_next == local_end)
                {
                    const uint32_t want = chunk_max ? (chunk > need ? chunk : need) : need;    // chunk_max == 0: reserve exactly what is idle
                    uint32_t       b    = 0;
                    if (lane == 0) b = atomicAdd(cursor, want);
                    b          = __shfl_sync(FULL, b, 0);
                    local_next = b < n ? b : n;
                    local_end  = b + want < n ? b + want : n;
                    if (local_next == local_end) exhausted = true;    // warp-uniform
                }
                const uint32_t avail = local_end - local_next;
                const uint32_t rank  = uint32_t(__popc(idle & ((1u << lane) - 1u)));
                const uint32_t base  = local_next;
                local_next += need < avail ? need : avail;
                if (!active)
                {
                    const uint32_t idx = base + rank;
                    if (rank < avail)
                    {
                        float    tmax;
                        uint32_t item;
                        V3       d;
                        uint32_t tri_off = 0;
                        if constexpr (OFFSETS)
                            source(idx, item, o, d, tmin, tmax, node_off, tri_off);
                        else
                            source(idx, item, o, d, tmin, tmax);
                        r_dir  = make_float4(d.x, d.y, d.z, __uint_as_float(tri_off));
                        r_win  = make_float4(0.0f, 0.0f, __uint_as_float(item), 0.0f);
                        best_t = tmax, best_prim = INVALID_PRIM;
                        if (bvh.n_nodes == 0 || (OFFSETS && tmax < 0.0f))
                        {
                            best_t   = __int_as_float(0x7f800000);
                            finished = true;
                        }
                        else
                        {
                            idir   = v3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
                            oct4   = make_oct4(d);
                            group  = make_uint2(0u, 0x80000000u);
                            tgroup = make_uint2(0u, 0u);
                            sp     = 0;
                            active = true;
                        }
                    }
                }
            }
            if (__ballot_sync(FULL, active || finished) == 0u) break;

#pragma unroll 1
            for (int it = 0; it < STEPS; it++)
            {
                // ---- node phase: lanes without pending triangles visit one node
                if (active && tgroup.y == 0u)
                {
                    const unsigned node_index = pop_inner(group, oct4, CRB_ANY_BACK_FIRST && any);
                    if (group.y) stack[sp++] = group;

                    uint4 n0, n1, n2, n3, n4;
                    load_node(bvh.nodes, OFFSETS ? size_t(node_off) + node_index : size_t(node_index), n0, n1, n2, n3, n4);
                    if (COUNT) ctr->nodes++;
                    node_visit(n0, n1, n2, n3, n4, o, idir, oct4, tmin, best_t, group, tgroup, occ);
                }
#if CRB_EARLY_POP
                // ---- pop BEFORE the leaf phase: a lane whose node group has no inner child left will need the next
                // stack entry as soon as its waiting triangles are done, and `group` is dead until then, so the
                // local-memory load is issued here and its latency hides behind the leaf phase (ncu: 20 % of k_trace's
                // long-scoreboard stalls sat on a pop whose value the next instruction used). Visit order unchanged:
                // the node phase runs only once tgroup is empty.
#if CRB_EARLY_POP == 2 && !defined(CRB_EMU)
                {
                    // predicated local loads straight into group's registers: written in C++ (or as one 64-bit
                    // load), ptxas loads into a temporary pair and moves it into `group` at once, i.e. waits for the
                    // load right here; two 32-bit loads have no register-pair constraint and need no move
                    const bool pop = active && group.y == 0u && sp > 0;
                    sp -= pop ? 1 : 0;
                    const size_t a = __cvta_generic_to_local(&stack[sp]);
                    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p ld.local.u32 %0, [%3];\n @p ld.local.u32 %1, [%3+4];\n}"
                                 : "+r"(group.x), "+r"(group.y)
                                 : "r"(unsigned(pop)), "l"(a)
                                 : "memory");
                }
#else
                if (active && group.y == 0u && sp > 0) group = stack[--sp];
#endif
#endif
                // ---- leaf phase in lock step: ONE triangle per lane that has triangles waiting (an inner
                // per-lane triangle loop was 52 % of k_trace's instructions at 2.7 active lanes; waiting for
                // more lanes to have triangles was measured and is slower, profiles/r1c_sweeps.md §6)
                const bool pending = active && tgroup.y != 0u;
                if (__ballot_sync(FULL, pending) != 0u)
                {
                    if (pending)
                    {
                        const float4  dq = r_dir;
                        float4 a, b, c;
                        load_tri(bvh.tris, OFFSETS ? size_t(__float_as_uint(dq.w)) + pop_triangle(tgroup, occ) : size_t(pop_triangle(tgroup, occ)), a, b, c);
                        if (COUNT) ctr->tris++;
                        float t, u, v;
                        if (tri_test(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), o, v3(dq.x, dq.y, dq.z), tmin, best_t, t, u, v))
                        {
                            const unsigned prim = __float_as_uint(a.w);
                            if (t < best_t || prim < best_prim)
                            {
                                best_t = t, best_prim = prim;
                                r_win.x = u, r_win.y = v;
                            }
                            if (any)
                            {
                                // any hit ends the query: drop all remaining work, the advance step below
                                // retires the ray (single retire point for both query kinds)
                                group.y  = 0u;
                                tgroup.y = 0u;
                                sp       = 0;
                            }
                        }
                    }
                }
                // ---- advance: nothing left in this node group -> pop, or retire the ray
                if (active && tgroup.y == 0u && group.y == 0u)
                {
#if CRB_EARLY_POP
                    // the early pop found the stack empty (or an any-hit dropped everything): retire
                    if (best_prim == INVALID_PRIM) best_t = __int_as_float(0x7f800000);
                    active = false, finished = true;
#else
                    if (sp == 0)
                    {
                        if (best_prim == INVALID_PRIM) best_t = __int_as_float(0x7f800000);
                        active = false, finished = true;
                    }
                    else
                        group = stack[--sp];
#endif
                }
            }
        }
    }
}    // namespace crb

// ====================================================================================================
// Two-level traversal: one BLAS per model in OBJECT space + a TLAS over the (model, instance) pairs.
//
// The reference has no instance geometry in Embree: cr::model::intersect loops over the model's glm::mat4 transforms,
// inverts each per ray, transforms the ray into object space, RE-NORMALISES the direction, queries the model's own
// Embree scene with tnear = 1e-5 / tfar = inf, maps the hit point back and RE-MEASURES the distance in world space
// (src/objects/model.cpp:99-126), and cr::scene::cast_ray loops over the models (src/render/scene.cpp:79-98); the
// nearest by strict '<' in loop order wins. Flattening the instances into world space (r1) changes the coordinates the
// triangle test runs in, so hits on shared edges of small triangles can land on the neighbouring triangle (measured:
// 99.985 % id agreement at the 18M-triangle config 4, below the 99.99 % bar). Here the per-instance query is the
// reference's own arithmetic — pre-inverted matrix (glm's cofactor inverse, computed on the host), same mat*vec order,
// same renormalisation, same re-measuring — so t,u,v,prim are bit-identical to the oracle for instanced models too,
// the scene stores every model once, and an instance edit rebuilds only the TLAS.
namespace crb
{
    struct Instance    // 144 bytes = nine 16-byte loads
    {
        float    inv[12];    // object <- world: rows 0..2 of glm::inverse(transform), column-major: inv[3*c + r]
        float    fwd[12];    // world <- object: rows 0..2 of the transform
        float    lo[3];      // world-space bounds of the instance
        uint32_t blas;       // model index
        float    hi[3];
        uint32_t flat_start; // flat primitive id of the instance's triangle 0 (FlatRange::start)
        uint32_t node_base, tri_base, n_nodes, pad;    // the model's BLAS (copied from Blas: one dependent load less per entry)
    };
    static_assert(sizeof(Instance) == 144, "Instance is read as nine float4");
    struct Blas
    {
        uint32_t node_base, tri_base, n_nodes, n_tris;
    };
    struct Bvh2
    {
        Bvh8            tlas;    // leaves hold one proxy "triangle" per instance; its prim id is the instance index
        const uint4    *nodes;   // all BLAS node arrays, concatenated (child/triangle bases are BLAS-relative)
        const float4   *tris;    // all BLAS triangle arrays, concatenated (prim id = triangle index inside the model)
        const Blas     *blas;
        const Instance *inst;
        uint32_t        n_inst;
    };

    // glm mat4 * vec4 restricted to rows 0..2: (m0*x + m1*y) + (m2*z + m3*w), one rounding per operation
    __device__ __forceinline__ V3 xf34(const float *M, V3 p, float w)
    {
        return v3(__fadd_rn(__fadd_rn(__fmul_rn(M[0], p.x), __fmul_rn(M[3], p.y)), __fadd_rn(__fmul_rn(M[6], p.z), __fmul_rn(M[9], w))),
                  __fadd_rn(__fadd_rn(__fmul_rn(M[1], p.x), __fmul_rn(M[4], p.y)), __fadd_rn(__fmul_rn(M[7], p.z), __fmul_rn(M[10], w))),
                  __fadd_rn(__fadd_rn(__fmul_rn(M[2], p.x), __fmul_rn(M[5], p.y)), __fadd_rn(__fmul_rn(M[8], p.z), __fmul_rn(M[11], w))));
    }

    // exact slab test of a ray against a world-space box (instance bounds), conservative like the node test
    __device__ __forceinline__ bool ray_box(V3 o, V3 idir, const float *lo, const float *hi, float tmin, float tmax)
    {
        const float ax = (lo[0] - o.x) * idir.x, bx = (hi[0] - o.x) * idir.x, ay = (lo[1] - o.y) * idir.y, by = (hi[1] - o.y) * idir.y;
        const float az = (lo[2] - o.z) * idir.z, bz = (hi[2] - o.z) * idir.z;
        const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
        const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
        return tn <= tf * 1.00001f + 1e-30f || !(tn == tn) || !(tf == tf);    // NaN (origin on a face of a flat box): enter
    }

    constexpr int BVH2_STACK = 80;    // TLAS levels + marker and pending-leaf entries of one instance entry + BLAS levels

    // Persistent two-level trace loop (same per-lane refill scheme as trace_persistent).
    //   RENORM = true : scene::cast_ray semantics (the renderer): per instance the direction is renormalised, the BLAS is
    //                   queried with tnear 1e-5 / tfar inf, candidates are compared by re-measured WORLD distance, ties
    //                   go to the first (model, instance) in the reference's loop order; tmax (finite for area-light
    //                   shadow rays) bounds the world distance. sink gets Hit{ t in OBJECT space of the winning
    //                   instance, u, v, flat prim } (the shader recomputes the world point like model.cpp:116-120).
    //   RENORM = false: batch queries (rtcIntersect1 with the caller's tnear/tfar): the ray parameter t is invariant
    //                   under the affine map, d is NOT renormalised, candidates are compared by t.
    // Every iteration is a node phase (TLAS or BLAS nodes, one code path), a triangle phase (BLAS leaves) and an ENTRY
    // phase in which lanes standing on a TLAS leaf take their ray into the instance's object space. The entry is ~150
    // instructions (two matrix products, a normalisation, three reciprocals) and almost every iteration has SOME lane
    // that wants it, so run eagerly it was paid per iteration at 1-3 active lanes (measured: k_trace2 1.74 against 3.48
    // Grays/s for the flattened scene at equal node visits); it now runs when CRB_ENTRY_MIN lanes wait for it or nothing
    // else in the warp can progress. The TLAS ray is kept per lane, so leaving an instance is a register move.
#ifndef CRB_ENTRY_MIN
#define CRB_ENTRY_MIN 6
#endif
    // The query as given and the TLAS-level ray (13 values per lane) are only needed when a lane enters or leaves an
    // instance: kept in shared memory instead of registers (CTAs of 256 threads: 16 KB), which is what lets the loop fit
    // the register budget of one more resident CTA per SM.
#ifndef CRB_2L_SMEM
#define CRB_2L_SMEM 1
#endif
    template<bool COUNT, int STEPS, bool RENORM, typename Source, typename Sink>
    __device__ __forceinline__ void trace_persistent_2l(const Bvh2 &sc, uint32_t *cursor, uint32_t n, bool any, Source source, Sink sink, TravCounters *ctr)
    {
        const unsigned FULL = 0xffffffffu;
        const unsigned lane = crb_lane_id();
        constexpr uint32_t MARK = 0xffffffffu;    // stack entry {MARK, 0}: leave the current instance
        uint2    stack[BVH2_STACK];
        int      sp = 0;
        bool     active = false, finished = false, exhausted = false, in_blas = false;
        uint32_t item = 0;
#if CRB_2L_SMEM && !defined(CRB_EMU)
        __shared__ float4 s_wo[256], s_wd[256], s_td[256], s_ti[256];    // (wo, tmin_w) (wd, tmax_w) (td, toct) (tidir, -)
        float4 &r_wo = s_wo[threadIdx.x], &r_wd = s_wd[threadIdx.x], &r_td = s_td[threadIdx.x], &r_ti = s_ti[threadIdx.x];
#else
        float4 r_wo = make_float4(0, 0, 0, 0), r_wd = make_float4(0, 0, 1, 0), r_td = make_float4(0, 0, 1, 0), r_ti = make_float4(0, 0, 0, 0);
#endif
        bool unbounded = false;    // the query's tmax is infinite
        V3       o = v3(0, 0, 0), d = v3(0, 0, 1), idir = v3(0, 0, 0);    // the ray of the current level
        unsigned oct4 = 0, occ = 0;
        float    tmin = 0.f;
        uint32_t node_off = 0, tri_off = 0, cur = 0;
        Hit      loc { 0.f, 0.f, 0.f, INVALID_PRIM };     // best inside the current instance
        Hit      best { 0.f, 0.f, 0.f, INVALID_PRIM };    // best overall: t (object space if RENORM), u, v, flat prim
        float    best_key = 0.f;                          // what candidates are compared by: world distance (RENORM) or t
        uint32_t best_k = 0;
        uint2    group = make_uint2(0u, 0u), tgroup = make_uint2(0u, 0u);
        uint32_t local_next = 0, local_end = 0;

        for (;;)
        {
            sink(finished, item, best);
            finished = false;

            const unsigned idle = __ballot_sync(FULL, !active);
            if (idle != 0u && !exhausted)
            {
                const uint32_t need = uint32_t(__popc(idle));
                if (local_next == local_end)
                {
                    uint32_t bb = 0;
                    if (lane == 0) bb = atomicAdd(cursor, need);
                    bb         = __shfl_sync(FULL, bb, 0);
                    local_next = bb < n ? bb : n;
                    local_end  = bb + need < n ? bb + need : n;
                    if (local_next == local_end) exhausted = true;
                }
                const uint32_t avail = local_end - local_next;
                const uint32_t rank  = uint32_t(__popc(idle & ((1u << lane) - 1u)));
                const uint32_t base  = local_next;
                local_next += need < avail ? need : avail;
                if (!active && rank < avail)
                {
                    V3    wo, wd;
                    float tmin_w, tmax_w;
                    source(base + rank, item, wo, wd, tmin_w, tmax_w);
                    r_wo      = make_float4(wo.x, wo.y, wo.z, tmin_w);
                    r_wd      = make_float4(wd.x, wd.y, wd.z, tmax_w);
                    unbounded = !(tmax_w < __int_as_float(0x7f800000));
                    best      = Hit { tmax_w, 0.0f, 0.0f, INVALID_PRIM };
                    best_key  = tmax_w, best_k = MARK;
                    if (sc.tlas.n_nodes == 0 || sc.n_inst == 0)
                    {
                        best.t   = __int_as_float(0x7f800000);
                        finished = true;
                    }
                    else
                    {
                        // TLAS traversal runs on the world ray, with a unit direction (t = world distance) when RENORM
                        const V3 td = RENORM ? normalize(wd) : wd;
                        o = wo, d = td;
                        idir   = v3(safe_rcp(td.x), safe_rcp(td.y), safe_rcp(td.z));
                        oct4   = make_oct4(td);
                        r_td   = make_float4(td.x, td.y, td.z, __uint_as_float(oct4));
                        r_ti   = make_float4(idir.x, idir.y, idir.z, 0.0f);
                        tmin   = RENORM ? 0.0f : tmin_w;
                        node_off = 0, tri_off = 0, in_blas = false;
                        group  = make_uint2(0u, 0x80000000u);
                        tgroup = make_uint2(0u, 0u);
                        sp     = 0;
                        active = true;
                    }
                }
            }
            if (__ballot_sync(FULL, active || finished) == 0u) break;

#pragma unroll 1
            for (int it = 0; it < STEPS; it++)
            {
                // ---- node phase (TLAS or BLAS nodes: same layout)
                const bool do_node = active && tgroup.y == 0u && group.y != 0u;
                if (do_node)
                {
                    const unsigned node_index = pop_inner(group, oct4, CRB_ANY_BACK_FIRST && any);
                    if (group.y) stack[sp++] = group;
                    uint4 n0, n1, n2, n3, n4;
                    load_node(in_blas ? sc.nodes + size_t(node_off) * BVH8_NODE_U4 : sc.tlas.nodes, node_index, n0, n1, n2, n3, n4);
                    if (COUNT) ctr->nodes++;
                    // far limit of the slab test: inside an instance the local best; in the TLAS the best so far (a world
                    // distance when RENORM, the ray parameter otherwise), loosened so that candidates within rounding are kept
                    const float    lim = in_blas ? loc.t : (RENORM ? best_key * 1.00001f : best_key);
                    node_visit(n0, n1, n2, n3, n4, o, idir, oct4, tmin, lim, group, tgroup, occ);
                }
                // ---- triangle phase: one BLAS triangle per lane that has some waiting
                const bool pending = active && in_blas && tgroup.y != 0u;
                if (__ballot_sync(FULL, pending) != 0u)
                {
                    if (pending)
                    {
                        float4 a, b, c;
                        load_tri(sc.tris, size_t(tri_off) + pop_triangle(tgroup, occ), a, b, c);
                        if (COUNT) ctr->tris++;
                        float t, u, v;
                        if (tri_test(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), o, d, tmin, loc.t, t, u, v))
                        {
                            const unsigned prim = __float_as_uint(a.w);
                            if (t < loc.t || prim < loc.prim) loc = Hit { t, u, v, prim };
                            if (any && unbounded)
                            {
                                // an unbounded shadow ray: any triangle of any instance ends the query
                                best = Hit { t, u, v, sc.inst[cur].flat_start + prim };
                                group.y = 0u, tgroup.y = 0u, sp = 0;
                                in_blas = false;
                            }
                        }
                    }
                }
                // ---- entry phase: lanes standing on a TLAS leaf take ONE of its instances into object space. The entry is long
                // (two matrix products, a normalisation, three reciprocals) and rare per lane, so it runs when CRB_ENTRY_MIN
                // lanes wait for it or nothing else in the warp can progress.
                const bool     want  = active && !in_blas && tgroup.y != 0u;
                const unsigned wantm = __ballot_sync(FULL, want);
                if (wantm != 0u)
                {
                    // someone else can still do a node or a triangle next iteration?
                    const unsigned busy = __ballot_sync(FULL, active && !want && (tgroup.y != 0u || group.y != 0u || sp > 0));
                    if (__popc(wantm) >= CRB_ENTRY_MIN || busy == 0u)
                    {
                        if (want)
                        {
                            // the proxy triangle's id is the instance index; the record is nine 16-byte loads
                            const uint32_t k  = __float_as_uint(__ldg(sc.tlas.tris + size_t(pop_triangle(tgroup, occ)) * BVH8_TRI_F4).w);
                            const float4  *ip = reinterpret_cast<const float4 *>(sc.inst + k);
                            const float4   i6 = __ldg(ip + 6), i7 = __ldg(ip + 7);
                            const float    lo[3] = { i6.x, i6.y, i6.z }, hi[3] = { i7.x, i7.y, i7.z };
                            const float    lim = RENORM ? best_key * 1.00001f : best_key;
                            if (ray_box(o, idir, lo, hi, tmin, lim))
                            {
                                const float4 v0 = __ldg(ip), v1 = __ldg(ip + 1), v2 = __ldg(ip + 2);
                                const uint4  i8 = __ldg(reinterpret_cast<const uint4 *>(ip + 8));
                                const float  inv[12] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w };
                                if (group.y) stack[sp++] = group;
                                if (tgroup.y)
                                {
                                    // the other instances of this TLAS leaf: (occupancy word) below (triangle base, waiting bits);
                                    // waiting bits lie in 0x77777777, inner-hit flags in 0x88888888: the entry kinds cannot be confused
                                    stack[sp++] = make_uint2(occ, 0u);
                                    stack[sp++] = tgroup;
                                }
                                stack[sp++] = make_uint2(MARK, 0u);
                                cur = k;
                                // model.cpp:107-112: inv * vec4(origin, 1), normalize(inv * vec4(direction, 0))
                                const float4 qo = r_wo, qd = r_wd;
                                o = xf34(inv, v3(qo.x, qo.y, qo.z), 1.0f);
                                d = xf34(inv, v3(qd.x, qd.y, qd.z), 0.0f);
                                if (RENORM) d = normalize(d);
                                idir   = v3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
                                oct4   = make_oct4(d);
                                node_off = i8.x, tri_off = i8.y, in_blas = true;
                                float bound = best_key;
                                if (RENORM)
                                {
                                    tmin = 0.00001f;    // model.cpp:21
                                    if (best_key < __int_as_float(0x7f800000))
                                    {
                                        // world distance per unit of object-space t along this ray = |fwd * d'|: a hit beyond
                                        // best / that (loosened) cannot win the exact comparison made when the instance is left
                                        const float4 f0 = __ldg(ip + 3), f1 = __ldg(ip + 4), f2 = __ldg(ip + 5);
                                        const float  fwd[12] = { f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w };
                                        const V3     wdir = xf34(fwd, d, 0.0f);
                                        bound             = (best_key / __fsqrt_rn(dot(wdir, wdir))) * 1.0001f;
                                    }
                                }
                                loc    = Hit { bound, 0.0f, 0.0f, INVALID_PRIM };
                                group  = make_uint2(0u, i8.z ? 0x80000000u : 0u);
                                tgroup = make_uint2(0u, 0u);
                            }
                        }
                    }
                }
                // ---- advance: pop node groups / pending instances / instance markers, or retire
                while (active && tgroup.y == 0u && group.y == 0u)
                {
                    if (sp == 0)
                    {
                        if (best.prim == INVALID_PRIM) best.t = __int_as_float(0x7f800000);
                        active = false, finished = true;
                        break;
                    }
                    const uint2 e = stack[--sp];
                    if (e.x == MARK && e.y == 0u)
                    {
                        // leave the instance: model.cpp:116-123 (map the point back, re-measure, keep the nearest)
                        if (loc.prim != INVALID_PRIM)
                        {
                            const float4 *ip = reinterpret_cast<const float4 *>(sc.inst + cur);
                            float         key = loc.t;
                            const float4  qo = r_wo;
                            const V3      wo = v3(qo.x, qo.y, qo.z);
                            const float   tmax_w = r_wd.w;
                            if (RENORM)
                            {
                                const float4 f0 = __ldg(ip + 3), f1 = __ldg(ip + 4), f2 = __ldg(ip + 5);
                                const float  fwd[12] = { f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w };
                                key                  = length(xf34(fwd, o + d * loc.t, 1.0f) - wo);    // glm::distance(point, ray.origin)
                            }
                            if ((key < best_key || (key == best_key && cur < best_k)) && key <= tmax_w)
                            {
                                best     = Hit { loc.t, loc.u, loc.v, __float_as_uint(__ldg(ip + 7).w) + loc.prim };
                                best_key = key, best_k = cur;
                                if (any) sp = 0;    // a bounded shadow ray is blocked: done
                            }
                            loc.prim = INVALID_PRIM;
                        }
                        {
                            const float4 qo = r_wo, qt = r_td, qi = r_ti;
                            o = v3(qo.x, qo.y, qo.z), d = v3(qt.x, qt.y, qt.z), idir = v3(qi.x, qi.y, qi.z), oct4 = __float_as_uint(qt.w);
                            tmin = RENORM ? 0.0f : qo.w;
                        }
                        node_off = 0, tri_off = 0, in_blas = false;
                        continue;
                    }
                    if ((e.y & 0x88888888u) == 0u)
                    {
                        tgroup = e;    // pending instances of a TLAS leaf, their occupancy word below
                        occ    = stack[--sp].x;
                    }
                    else
                        group = e;
                }
            }
        }
    }
}    // namespace crb
