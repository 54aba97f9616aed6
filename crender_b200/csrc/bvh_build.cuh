// bvh_build.cuh — interface of the device-side BVH builder (the rtcCommitScene replacement,
// src/objects/model.cpp:92-94).
#pragma once
#include "platform.cuh"
#include "bvh8.cuh"

namespace crb
{
    struct BuildStats
    {
        double   build_ms  = 0;
        uint32_t n_nodes   = 0;
        uint32_t n_tris    = 0;
        uint32_t max_depth = 0;
        float    sah_cost  = 0;
        uint32_t treelets_changed = 0;
    };

    struct BuildOptions
    {
        // SAH treelet restructuring sweeps. Two: the second costs 2.7 ms at 1 M triangles (build 4.8 -> 7.5 ms) and buys 1.9 % of a
        // config-2 step (30.7 instead of 31.3 ms per 16 spp: 8.99 -> 8.31 node visits per shadow query, 11.83 -> 11.73 per
        // closest-hit query), i.e. it pays for itself after five steps; a third buys nothing (profiles/r2_sweeps.md section 19)
        int  treelet_passes = 2;
        bool optimal_collapse = true;  // SAH-optimal (dynamic programming) binary -> 8-wide collapse; false = greedy by area
        float cost_prim = 0.8f;        // collapse DP: cost of a triangle test relative to an 8-wide node test (swept: profiles/r1c_sweeps.md)
    };

    // wverts: device pointer, 9 floats per triangle (world space), n triangles.
    // Produces the node and triangle arrays of the 8-wide BVH. Throws crb::Error on failure.
    void build_bvh8(const float *d_wverts, uint32_t n, cudaStream_t stream, const BuildOptions &opt, DBuf<uint4> &nodes, DBuf<float4> &tris,
                    BuildStats &stats);
}    // namespace crb
