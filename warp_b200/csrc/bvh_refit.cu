// Bottom-up AABB refit over the sibling-pair layout.
//
// Behavioural contract = warp/native/bvh.cu:42-144 + mesh.cu:368-407: every visible node's box
// becomes the exact union (fminf/fmaxf) of the boxes of the items below it.  The reference starts
// one thread per ORIGINAL leaf and climbs through the muted nodes under each packed leaf; here one
// thread per sorted position gathers its triangle straight from the vertex array (no lowers/uppers
// round trip, no edge-length / scan passes) and refreshes the packed-triangle cache the queries read;
// the first thread of each VISIBLE packed leaf (found through pos_parent[], written by the builder)
// unions the leaf's <= leaf_size boxes and the merge pass climbs with one atomic arrival counter per
// internal node.  The counters are never cleared: they are
// even after a build and every refit adds exactly 2, so "second to arrive" == odd old value.
#include "state.h"
#include "merge.cuh"  // wb_store_box

namespace {

constexpr int BT = 256;

// item (triangle / box) at sorted position k: bounds, and for meshes the refreshed packed-triangle record
template <class Src, bool WRITE>
__device__ __forceinline__ void refit_item(const Src& src, const int* __restrict__ prim, float4* __restrict__ tris, int k,
                                           float3& lo, float3& hi)
{
    const int item = __ldg(prim + k);
    if constexpr (Src::kIsMesh) {
        float3 p, q, r;
        src.tri(item, p, q, r);
        lo = wb_min3(wb_min3(p, q), r);
        hi = wb_max3(wb_max3(p, q), r);
        if (WRITE) {
            const float3 e0 = wb_sub(q, p), e1 = wb_sub(r, p), e2 = wb_sub(r, q);
            const float3 nrm = wb_cross(e0, e1);
            const float area2 = sqrtf(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z);
            const bool sliver = area2 / (wb_dot(e0, e0) + wb_dot(e1, e1) + wb_dot(e2, e2)) < 1.e-6f;
            float4* t = tris + 3 * (size_t)k;
            t[0] = make_float4(p.x, p.y, p.z, q.x);
            t[1] = make_float4(q.y, q.z, r.x, r.y);
            t[2] = make_float4(r.z, __int_as_float(item), __uint_as_float(sliver ? WB_TRI_SLIVER : 0u), 0.f);
        }
    } else {
        src.bounds(item, lo, hi);
    }
}

// One thread per sorted POSITION gathers its item and refreshes its triangle record (coalesced, like the
// builder's k_leaves); the item boxes are staged in shared memory, and the thread at the first position of each
// visible leaf unions its leaf's boxes from there.  Items of a leaf that continue past the block's last
// position are gathered again by the leaf's thread (their records are written by the block that owns them).
template <class Src>
__global__ void __launch_bounds__(BT)
k_refit_leaves(Src src, int n, const int* __restrict__ prim, const int* __restrict__ pos_parent, NodeRec* pairs,
               float4* __restrict__ tris, TreeHeader* hdr)
{
    __shared__ float sbox[6][BT];
    const int tid = threadIdx.x;
    const int i = blockIdx.x * BT + tid;
    float3 lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    int parent = WB_NO_PARENT;
    if (i < n) {
        parent = __ldg(pos_parent + i);
        refit_item<Src, true>(src, prim, tris, i, lo, hi);
        sbox[0][tid] = lo.x, sbox[1][tid] = lo.y, sbox[2][tid] = lo.z;
        sbox[3][tid] = hi.x, sbox[4][tid] = hi.y, sbox[5][tid] = hi.z;
    }
    __syncthreads();
    if (parent == WB_NO_PARENT)
        return;

    const int s = parent - n;  // internal slot of the parent
    int side = 0;
    int count;
    if (parent == WB_ROOT_PARENT) {
        count = n;
    } else if (i <= s) {
        side = 0;
        count = s - i + 1;
    } else {
        side = 1;
        count = (int)pairs[2 * (size_t)s + 1].aux - s;
    }

    int k = 1;
    for (; k < count && tid + k < BT; ++k) {
        lo = wb_min3(lo, make_float3(sbox[0][tid + k], sbox[1][tid + k], sbox[2][tid + k]));
        hi = wb_max3(hi, make_float3(sbox[3][tid + k], sbox[4][tid + k], sbox[5][tid + k]));
    }
    for (; k < count; ++k) {
        float3 a, b;
        refit_item<Src, false>(src, prim, tris, i + k, a, b);
        lo = wb_min3(lo, a);
        hi = wb_max3(hi, b);
    }

    if (parent == WB_ROOT_PARENT) {
        hdr->lx = lo.x, hdr->ly = lo.y, hdr->lz = lo.z;
        hdr->hx = hi.x, hdr->hy = hi.y, hdr->hz = hi.z;
        return;
    }

    // the merge pass (k_merge<true>, merge.cuh) unions these leaf boxes bottom-up
    wb_store_box(pairs + 2 * (size_t)s + side, lo, hi);
}

}  // namespace

#define WB_CUDA_TRY(expr)                  \
    do {                                   \
        cudaError_t _e = (expr);           \
        if (_e != cudaSuccess)             \
            return cudaGetErrorString(_e); \
    } while (0)

const char* wb_refit(BvhState& s, cudaStream_t stream)
{
    if (s.n <= 0)
        return nullptr;
    const int grid = wb_div_up(s.n, BT);
    if (s.is_mesh)
        k_refit_leaves<<<grid, BT, 0, stream>>>(MeshSource { s.points, s.indices }, s.n, s.prim, s.pos_parent, s.pairs,
                                                s.tris, s.header);
    else
        k_refit_leaves<<<grid, BT, 0, stream>>>(BoxSource { s.item_lowers, s.item_uppers }, s.n, s.prim, s.pos_parent,
                                                s.pairs, s.tris, s.header);
    WB_CUDA_TRY(cudaGetLastError());
    return wb_refit_merge(s, stream);
}
