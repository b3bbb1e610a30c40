// Bottom-up node merging shared by the builder (topology + boxes) and the refit (boxes only).
//
// The reference climbs from every leaf with one atomic counter per internal node (bvh.cu:261-393,
// 42-144): 2 global atomics + 2 fences per node.  Here each thread owns MC consecutive sorted
// positions and replays the same bottom-up process SEQUENTIALLY inside its chunk with a small
// stack: a node that wants to merge to the right is parked until the thread itself produces its
// right sibling; merges whose two children are both in the thread's hands need no atomic, no fence
// and no waiting.  Only nodes whose sibling is carried by another thread (chunk boundaries and
// the spine above them) use the global arrival counter.  The parent of a node is a function of
// its key range alone, so the tree is bit-identical to the reference's.
//
// Key flavours: uint32 = the reference's 30-bit Morton code (ungrouped parity mode);
// uint64 + GROUPED = group << 32 | 30-bit code (bvh.cu:205-209); uint64 ungrouped = 63-bit code.
#pragma once

#include "common.cuh"

#ifndef WB_MC
#define WB_MC 8
#endif
constexpr int MC = WB_MC;  // sorted positions per merge thread

template <class KeyT> struct MergeArgs {
    int n;
    int leaf_size;
    const KeyT* keys;
    const int* prim;
    NodeRec* pairs;
    int* parent_int;
    int* pos_parent;
    unsigned* counters;
    TreeHeader* hdr;
};

__device__ __forceinline__ int wb_clz_key(uint32_t x) { return __clz((int)x); }
__device__ __forceinline__ int wb_clz_key(uint64_t x) { return __clzll((long long)x); }

// common-prefix length of keys i and i+1 (bvh.cu:218-226; equal keys give the full width, i.e. the
// reference's 64; only comparisons between deltas of the same width are ever made)
template <class KeyT> __device__ __forceinline__ int wb_key_delta(const KeyT* __restrict__ keys, int i)
{
    return wb_clz_key((KeyT)(__ldg(keys + i) ^ __ldg(keys + i + 1)));
}

template <class KeyT> __device__ __forceinline__ uint32_t wb_group_of(const KeyT* __restrict__ keys, int i)
{
    return (uint32_t)((uint64_t)__ldg(keys + i) >> 32);
}

// packed-leaf eligibility by size: fits leaf_size and (grouped trees) does not straddle groups (bvh.cu:431-437)
template <class KeyT, bool GROUPED>
__device__ __forceinline__ bool wb_size_leaf(const KeyT* __restrict__ keys, int leaf_size, int left, int right)
{
    if (right - left + 1 > leaf_size)
        return false;
    if (GROUPED)
        return wb_group_of(keys, left) == wb_group_of(keys, right);
    return true;
}

// parent choice of the node covering sorted positions [left, right] (bvh.cu:300-334):
// true = it becomes the LEFT child of node n+right, false = the RIGHT child of node n+left-1
template <class KeyT, bool GROUPED>
__device__ __forceinline__ bool wb_goes_right(const KeyT* __restrict__ keys, const int* __restrict__ prim, int n,
                                              int left, int right)
{
    if (left == 0)
        return true;
    if (GROUPED) {  // stay inside the group when exactly one neighbour allows it (bvh.cu:305-321)
        const uint32_t gl = wb_group_of(keys, left), gr = wb_group_of(keys, right);
        if (gl == gr) {
            const bool right_same = (right < n - 1) && wb_group_of(keys, right + 1) == gl;
            const bool left_same = wb_group_of(keys, left - 1) == gl;
            if (right_same != left_same)
                return right_same;
        }
    }
    if (right == n - 1)
        return false;
    const int dr = wb_key_delta(keys, right), dl = wb_key_delta(keys, left - 1);
    if (dr != dl)
        return dr > dl;
    return ((__ldg(prim + left - 1) % 2) ^ (__ldg(prim + right) % 2)) != 0;
}

__device__ __forceinline__ void wb_store_rec(NodeRec* dst, float3 lo, float3 hi, uint32_t ref, uint32_t aux)
{
    float4* d4 = reinterpret_cast<float4*>(dst);
    d4[0] = make_float4(lo.x, lo.y, lo.z, __uint_as_float(ref));
    d4[1] = make_float4(hi.x, hi.y, hi.z, __uint_as_float(aux));
}

__device__ __forceinline__ void wb_store_box(NodeRec* dst, float3 lo, float3 hi)
{
    dst->lx = lo.x, dst->ly = lo.y, dst->lz = lo.z;
    dst->hx = hi.x, dst->hy = hi.y, dst->hz = hi.z;
}

template <bool REFIT, class KeyT, bool GROUPED>
__global__ void __launch_bounds__(128)
k_merge(MergeArgs<KeyT> a)
{
    const int n = a.n;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long c0l = t * MC;
    if (c0l >= n)
        return;
    const int c0 = (int)c0l;
    const int c1 = min(c0 + MC - 1, n - 1);

    int rstack[MC];        // split positions of parked nodes (each is the LEFT child of n + rstack[k])
    unsigned hstack[MC];   // their heights (build only)
    int depth = 0;
    int pos = c0;

    // the node currently in hand
    bool have = false, fresh = false;  // fresh: a leaf unit whose record is already in memory
    int xl = 0, xr = 0;
    uint32_t xnode = 0;
    unsigned xh = 0;
    float3 lo = make_float3(0.f, 0.f, 0.f), hi = lo;
    int static_parent = WB_NO_PARENT;  // REFIT: parent of the node in hand

    for (;;) {
        bool go_right = false, merge_global = false;
        int s = 0;
        unsigned other_h = 0;
        bool resumed = false;  // true when a parked node was handed over and found its sibling waiting

        if (!have) {
            if (pos > c1) {
                if (depth == 0)
                    return;
                // our right neighbour is now in another thread's hands: hand the parked nodes over, top first
                --depth;
                s = rstack[depth];
                const unsigned h = min(hstack[depth], WB_HEIGHT_CAP);
                const unsigned old = wb_arrive(&a.counters[s], REFIT ? 1u : (1u | (h << 8)));
                const bool second = REFIT ? (old & 1u) != 0u : (old & 0xffu) != 0u;
                if (!second)
                    continue;
                // the right sibling was already there: take the parked node back in hand and merge below
                const NodeRec L = a.pairs[2 * (size_t)s];  // our own earlier store
                lo = make_float3(L.lx, L.ly, L.lz);
                hi = make_float3(L.hx, L.hy, L.hz);
                xl = (int)L.aux;
                xr = s;
                xnode = L.ref & WB_IDX_MASK;
                xh = h;
                go_right = true, merge_global = true, other_h = old >> 8;
                have = true, fresh = false, resumed = true;
            } else if (REFIT) {
                // next visible leaf of this chunk (its box was refreshed by the leaf pass)
                const int p = a.pos_parent[pos];
                if (p == WB_NO_PARENT) {
                    ++pos;
                    continue;
                }
                if (p == WB_ROOT_PARENT)
                    return;  // the root is a packed leaf: the leaf pass already wrote the header box
                const int ps = p - n;
                const NodeRec* rec = a.pairs + 2 * (size_t)ps + (pos <= ps ? 0 : 1);
                xl = pos;
                xr = (pos <= ps) ? ps : (int)rec->aux;
                lo = make_float3(rec->lx, rec->ly, rec->lz);
                hi = make_float3(rec->hx, rec->hy, rec->hz);
                static_parent = p;
                xnode = 0;  // unused for leaf units
                pos = xr + 1;
                have = true, fresh = true;
            } else {
                // next original leaf (its record was written by the leaf pass)
                const bool gr = wb_goes_right<KeyT, GROUPED>(a.keys, a.prim, n, pos, pos);
                const NodeRec* rec = a.pairs + 2 * (size_t)(gr ? pos : pos - 1) + (gr ? 0 : 1);
                xl = xr = pos;
                lo = make_float3(rec->lx, rec->ly, rec->lz);
                hi = make_float3(rec->hx, rec->hy, rec->hz);
                xnode = (uint32_t)pos;
                xh = 0;
                ++pos;
                have = true, fresh = true;
            }
        }

        if (!resumed) {
            // ---- the node in hand: root, or choose its parent
            if (REFIT) {
                if (static_parent == WB_NO_PARENT) {
                    a.hdr->lx = lo.x, a.hdr->ly = lo.y, a.hdr->lz = lo.z;
                    a.hdr->hx = hi.x, a.hdr->hy = hi.y, a.hdr->hz = hi.z;
                    return;
                }
                s = static_parent - n;
                go_right = (xr == s);  // a left child's range ends at the split
            } else {
                if (xl == 0 && xr == n - 1) {
                    const uint32_t self_ref =
                        xnode | (wb_size_leaf<KeyT, GROUPED>(a.keys, a.leaf_size, xl, xr) ? WB_LEAF : 0u);
                    a.hdr->lx = lo.x, a.hdr->ly = lo.y, a.hdr->lz = lo.z;
                    a.hdr->hx = hi.x, a.hdr->hy = hi.y, a.hdr->hz = hi.z;
                    a.hdr->root_ref = self_ref;
                    a.hdr->root_count = (uint32_t)n;
                    a.hdr->height = (int)xh;
                    a.hdr->deep = 0;
                    a.hdr->n = n;
                    a.hdr->leaf_size = a.leaf_size;
                    a.parent_int[xnode - n] = WB_NO_PARENT;  // n >= 2: the root is internal
                    if (self_ref & WB_LEAF)
                        a.pos_parent[0] = WB_ROOT_PARENT;
                    return;
                }
                go_right = wb_goes_right<KeyT, GROUPED>(a.keys, a.prim, n, xl, xr);
                s = go_right ? xr : xl - 1;
                if (xnode >= (uint32_t)n)
                    a.parent_int[xnode - n] = n + s;
            }

            NodeRec* mine = a.pairs + 2 * (size_t)s + (go_right ? 0 : 1);
            if (!fresh) {
                if (REFIT)
                    wb_store_box(mine, lo, hi);
                else
                    wb_store_rec(mine, lo, hi,
                                 xnode | (wb_size_leaf<KeyT, GROUPED>(a.keys, a.leaf_size, xl, xr) ? WB_LEAF : 0u),
                                 (uint32_t)(go_right ? xl : xr));
            }

            if (go_right && xr < c1) {  // our own next unit will become (part of) the right sibling: park
                rstack[depth] = s;
                hstack[depth] = xh;
                ++depth;
                have = false;
                continue;
            }
            if (!go_right && depth > 0) {
                // the parked top is exactly the left child of n+s: both children are in our hands
                --depth;
                other_h = hstack[depth];
                merge_global = false;
            } else {
                const unsigned h = min(xh, WB_HEIGHT_CAP);
                const unsigned old = wb_arrive(&a.counters[s], REFIT ? 1u : (1u | (h << 8)));
                const bool second = REFIT ? (old & 1u) != 0u : (old & 0xffu) != 0u;
                if (!second) {
                    have = false;  // the sibling's carrier continues; parked nodes (if any) are handed over above
                    continue;
                }
                other_h = old >> 8;
                merge_global = true;
            }
        }

        // ---- second to complete n+s: union with the sibling record and become the parent
        const NodeRec* sibling = a.pairs + 2 * (size_t)s + (go_right ? 1 : 0);
        float4 s0, s1;
        if (merge_global) {
            s0 = __ldcg(reinterpret_cast<const float4*>(sibling));
            s1 = __ldcg(reinterpret_cast<const float4*>(sibling) + 1);
        } else {
            s0 = reinterpret_cast<const float4*>(sibling)[0];  // our own earlier store
            s1 = reinterpret_cast<const float4*>(sibling)[1];
        }
        const int far_end = (int)__float_as_uint(s1.w);
        const int new_left = go_right ? xl : far_end;
        const int new_right = go_right ? far_end : xr;
        if (!REFIT) {
            // visible packed leaves: children that qualify by size while this parent does not
            if (!wb_size_leaf<KeyT, GROUPED>(a.keys, a.leaf_size, new_left, new_right)) {
                if (wb_size_leaf<KeyT, GROUPED>(a.keys, a.leaf_size, new_left, s))
                    a.pos_parent[new_left] = n + s;
                if (wb_size_leaf<KeyT, GROUPED>(a.keys, a.leaf_size, s + 1, new_right))
                    a.pos_parent[s + 1] = n + s;
            }
            xh = max(xh, other_h) + 1u;
        }
        lo = wb_min3(lo, make_float3(s0.x, s0.y, s0.z));
        hi = wb_max3(hi, make_float3(s1.x, s1.y, s1.z));
        xl = new_left, xr = new_right;
        xnode = (uint32_t)(n + s);
        static_parent = REFIT ? a.parent_int[s] : WB_NO_PARENT;
        fresh = false;
    }
}
