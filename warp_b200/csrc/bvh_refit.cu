// Bottom-up AABB refit over the sibling-pair layout.
//
// Behavioural contract = warp/native/bvh.cu:42-144 + mesh.cu:368-407: every visible node's box
// becomes the exact union (fminf/fmaxf) of the boxes of the items below it.  The reference starts
// one thread per ORIGINAL leaf and climbs through the muted nodes under each packed leaf, after an
// edge-length kernel, a CUB scan, compute_triangle_bounds and a memset.  Here ONE kernel
// (k_tree<REFIT>, merge.cuh) gathers every triangle straight from the vertex array (coalesced over the
// sorted positions), refreshes the packed-triangle cache the queries read, unions the items of each
// VISIBLE packed leaf (found through pos_parent[], written by the builder) and replays the static
// tree bottom-up: inside a block from shared memory, above the blocks with one atomic arrival counter
// per internal node.  The global counters are never cleared: they are even after a build and every
// refit adds exactly 0 or 2, so "second to arrive" == odd old value.
#include "state.h"
#include "merge.cuh"

const char* wb_refit(BvhState& s, cudaStream_t stream)
{
    if (s.n <= 0)
        return nullptr;
    cudaError_t e;
    if (s.is_mesh) {
        const MergeArgs<MeshSource, uint32_t> ma { MeshSource { s.points, s.indices }, s.n, s.leaf_size, nullptr, s.prim, s.pairs,
                                                   s.parent_int, s.pos_parent, s.counters, s.header, s.tris };
        e = wb_launch_tree<true, MeshSource, uint32_t, false>(ma, stream);
    } else {
        const MergeArgs<BoxSource, uint32_t> ma { BoxSource { s.item_lowers, s.item_uppers }, s.n, s.leaf_size, nullptr, s.prim,
                                                  s.pairs, s.parent_int, s.pos_parent, s.counters, s.header, s.tris };
        e = wb_launch_tree<true, BoxSource, uint32_t, false>(ma, stream);
    }
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
