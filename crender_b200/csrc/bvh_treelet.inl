// bvh_treelet.inl — K3: SAH treelet restructuring of the binary LBVH (included by bvh_build.cu).
// Placeholder in this commit: the pass is a no-op; the collapse consumes the plain LBVH.
static void treelet_optimize(BinTree &, int, const float4 *, const float4 *, const uint32_t *, cudaStream_t) {}
