// Shared device/host definitions for the B200 LBVH path.
//
// Tree layout in HBM ("sibling-pair" layout):
//   * every node is one 32-byte record (NodeRec): AABB + a reference word + one auxiliary word;
//   * the two children of internal node p = N + s (s = split position in Morton order, the
//     reference's numbering, warp/native/bvh.cu:337,349) are stored ADJACENT at pair[2s], pair[2s+1],
//     so one aligned 64-byte fetch yields both child boxes -- the traversal never loads a node's
//     own box, it arrives with the parent;
//   * the record of a child carries ref = reference node index (| WB_LEAF when the child is a packed
//     leaf) and aux = the FAR end of its key range (range start for a left child, range end for a
//     right child); with s this gives the [start, start+count) primitive range without a side table.
// The reference's two-array layout (bvh.h:161-207) is reproducible from this one bit-for-bit
// (k_export_reference_layout) and is what the parity tests diff against the oracle.
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#define WB_LEAF 0x80000000u
#define WB_IDX_MASK 0x7fffffffu
#define WB_NO_PARENT (-1)
#define WB_ROOT_PARENT (-2)      // pos_parent[] value of a leaf that is the root itself
#define WB_QUERY_STACK 32        // BVH_QUERY_STACK_SIZE, warp/native/bvh.h:18
#define WB_MAX_DEPTH 32          // packed-leaf depth rule, warp/native/bvh.cu:437
#define WB_HEIGHT_CAP 0xffffu
// heights[] entry of an internal node: height in the low 12 bits (capped), min(range length - 2, 15) in the high 4 --
// enough to decide "does this node hold at most leaf_size positions" for leaf_size <= 16 from two coalesced bytes
// instead of its 64-byte pair record (depth pass, bvh_build.cu)
#define WB_HEIGHT_MASK 0x0fffu

// one node record = two float4: (lo.xyz, ref) (hi.xyz, aux)
struct __align__(16) NodeRec {
    float lx, ly, lz;
    uint32_t ref;
    float hx, hy, hz;
    uint32_t aux;
};

// per-tree header kept in device memory (read by every query, written by build/refit)
struct __align__(16) TreeHeader {
    float lx, ly, lz;       // root AABB
    uint32_t root_ref;      // reference index of the root (| WB_LEAF if the root is a packed leaf)
    float hx, hy, hz;
    uint32_t root_count;    // number of items (range of the root is [0, n))
    int height;             // longest root-to-original-leaf path in edges (capped at WB_HEIGHT_CAP)
    int deep;               // 1 when the depth>=32 rule fired on at least one node
    int n;
    int leaf_size;
    // scene bounds used for the Morton grid (bvh.cu:449-488)
    float total_lo[3];
    float total_hi[3];
    float inv_edges[3];
    int pad[3];
};

// everything a kernel needs to know about one tree (passed by value)
struct TreeView {
    const NodeRec* pairs;        // 2*(n-1) records; children of internal node n+s at [2s], [2s+1]
    const TreeHeader* header;
    const float4* tris;          // 3 float4 per sorted position (mesh only): p, q, r, face, flags
    const int* prim;             // primitive_indices (sorted position -> item)
    const int* parent_int;       // parent (reference index) of internal node n+s, WB_NO_PARENT for the root
    const int* pos_parent;       // parent of the visible leaf that starts at a sorted position (state.h)
    int n;
};

#define WB_TRI_SLIVER 1u

// wavefront refit (bvh_refit.cu): positions per block, key layout of the plan
#define WB_WAVE_BP 1024
#define WB_PLAN_HEIGHT_BITS 12
#define WB_PLAN_TOP 0xFFFFFFFEu
#define WB_PLAN_SKIP 0xFFFFFFFFu
#define WB_PLAN_DST_ROOT 0xFFFFFFFFu

__host__ __device__ inline int wb_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
__host__ __device__ inline uint16_t wb_pack_height(unsigned height, int range_len)
{
    const unsigned h = height < WB_HEIGHT_MASK ? height : WB_HEIGHT_MASK;
    const unsigned r = range_len - 2 < 15 ? (unsigned)(range_len - 2) : 15u;
    return (uint16_t)(h | (r << 12));
}

#ifdef __CUDACC__

__device__ __forceinline__ float3 wb_min3(float3 a, float3 b)
{
    return make_float3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z));
}
__device__ __forceinline__ float3 wb_max3(float3 a, float3 b)
{
    return make_float3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z));
}
__device__ __forceinline__ float3 wb_sub(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 wb_add(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 wb_scale(float s, float3 a) { return make_float3(a.x * s, a.y * s, a.z * s); }
// a.x*b.x + a.y*b.y + a.z*b.z, summed left to right (vec.h:518-521); the library is built with
// -fmad=false so this is three roundings + two, like the host build of the reference
__device__ __forceinline__ float wb_dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 wb_cross(float3 a, float3 b)
{
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float wb_get(float3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

// arrival counter update with release (our record stores are visible before the count) and acquire (the
// sibling's record is visible after it) semantics in ONE instruction: MEMBAR.ALL.GPU + ATOMG + CCTL.IVALL,
// versus MEMBAR.SC + CCTL.IVALL twice for __threadfence(); atomicAdd(); __threadfence().  (A release-only arrival,
// legal here because every load after it bypasses L1 or reads static data, measured the same: the fences are not
// what bounds the climb.)
__device__ __forceinline__ unsigned wb_arrive(unsigned* counter, unsigned add)
{
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(counter), "r"(add) : "memory");
    return old;
}

// Start of a traversal at the caller's `root` (a reference node index, e.g. from bvh_get_group_root; bvh.h:504:
// root == -1 ? *bvh.root : root).  Internal node n+s that is not the tree root: its own box, leaf flag and range live
// in its record inside the parent's pair.  Returns false when r is not such a node (the caller keeps its default).
// a / b are the traversal entry: leaf = first sorted position | WB_LEAF and item count, inner = internal slot and 0.
__device__ __forceinline__ bool wb_root_entry(const TreeView& tv, int r, uint32_t& a, uint32_t& b, float3& lo, float3& hi)
{
    if (r < tv.n || r >= 2 * tv.n - 1)
        return false;
    const int s = r - tv.n;
    const int p = __ldg(tv.parent_int + s);
    if (p == WB_NO_PARENT)
        return false;  // the tree root itself
    const int ps = p - tv.n;
    const int side = ((int)tv.pairs[2 * (size_t)s + 1].aux == ps) ? 0 : 1;  // a left child's range ends at the split
    const NodeRec rec = tv.pairs[2 * (size_t)ps + side];
    lo = make_float3(rec.lx, rec.ly, rec.lz), hi = make_float3(rec.hx, rec.hy, rec.hz);
    if (rec.ref & WB_LEAF) {
        const uint32_t first = side ? (uint32_t)ps + 1u : rec.aux, last = side ? rec.aux : (uint32_t)ps;
        a = first | WB_LEAF, b = last - first + 1u;
    } else {
        a = (uint32_t)s, b = 0;
    }
    return true;
}

// item-bounds sources: a triangle mesh (bounds computed on the fly from vertices, replacing the
// lowers/uppers round trip of mesh.cu:16-36) or caller-provided boxes (wp.Bvh)
struct MeshSource {
    const float* points;
    const int* indices;
    __device__ __forceinline__ void tri(int t, float3& p, float3& q, float3& r) const
    {
        const int i = __ldg(indices + 3 * (size_t)t + 0);
        const int j = __ldg(indices + 3 * (size_t)t + 1);
        const int k = __ldg(indices + 3 * (size_t)t + 2);
        p = make_float3(__ldg(points + 3 * (size_t)i), __ldg(points + 3 * (size_t)i + 1), __ldg(points + 3 * (size_t)i + 2));
        q = make_float3(__ldg(points + 3 * (size_t)j), __ldg(points + 3 * (size_t)j + 1), __ldg(points + 3 * (size_t)j + 2));
        r = make_float3(__ldg(points + 3 * (size_t)k), __ldg(points + 3 * (size_t)k + 1), __ldg(points + 3 * (size_t)k + 2));
    }
    __device__ __forceinline__ void bounds(int t, float3& lo, float3& hi) const
    {
        float3 p, q, r;
        tri(t, p, q, r);
        lo = wb_min3(wb_min3(p, q), r);
        hi = wb_max3(wb_max3(p, q), r);
    }
    static constexpr bool kIsMesh = true;
};

struct BoxSource {
    const float* lowers;
    const float* uppers;
    __device__ __forceinline__ void bounds(int t, float3& lo, float3& hi) const
    {
        lo = make_float3(__ldg(lowers + 3 * (size_t)t), __ldg(lowers + 3 * (size_t)t + 1), __ldg(lowers + 3 * (size_t)t + 2));
        hi = make_float3(__ldg(uppers + 3 * (size_t)t), __ldg(uppers + 3 * (size_t)t + 1), __ldg(uppers + 3 * (size_t)t + 2));
    }
    __device__ __forceinline__ void tri(int, float3&, float3&, float3&) const { }
    static constexpr bool kIsMesh = false;
};

#endif  // __CUDACC__
