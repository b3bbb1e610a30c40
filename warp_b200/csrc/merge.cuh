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
    uint16_t* heights;  // builder: height of every internal node, for the refit plan
    const int* unit;    // builder: per sorted position, what k_small_nodes left there (WB_UNIT_*); nullptr = all plain leaves
};

// unit[] values (k_small_nodes -> k_merge<build>): a position is a plain leaf, lies inside a subtree that starts further
// left, or starts a finished subtree whose top node is n + (pos + (u & 7)), covering [pos, pos + ((u >> 3) & 7)], of
// height (u >> 6) & 7
#define WB_UNIT_LEAF 0
#define WB_UNIT_COVERED (-1)
#define WB_UNIT_TOP 0x200
#define WB_SMALL_MAX 8  // largest key range (sorted positions) of a node that k_small_nodes builds

__device__ __forceinline__ int wb_clz_key(uint32_t x) { return __clz((int)x); }
__device__ __forceinline__ int wb_clz_key(uint64_t x) { return __clzll((long long)x); }

// Key accessors: the functions below read keys (and the parity of primitive indices, for the tie-break) either
// straight from global memory or from a block's shared-memory copy of its own key range (k_merge phase A).
template <class KeyT> struct GlobalKeys {
    const KeyT* keys;
    const int* prim;
    __device__ __forceinline__ KeyT key(int i) const { return __ldg(keys + i); }
    __device__ __forceinline__ int parity(int i) const { return __ldg(prim + i) % 2; }
};

template <class KeyT> struct BlockKeys {
    const KeyT* skeys;          // keys[base ...]
    const unsigned char* spar;  // prim[base ...] % 2
    int base;
    __device__ __forceinline__ KeyT key(int i) const { return skeys[i - base]; }
    __device__ __forceinline__ int parity(int i) const { return spar[i - base]; }
};

// common-prefix length of keys i and i+1 (bvh.cu:218-226; equal keys give the full width, i.e. the
// reference's 64; only comparisons between deltas of the same width are ever made)
template <class K> __device__ __forceinline__ int wb_key_delta_k(const K& k, int i)
{
    return wb_clz_key((decltype(k.key(0)))(k.key(i) ^ k.key(i + 1)));
}

template <class K> __device__ __forceinline__ uint32_t wb_group_of_k(const K& k, int i)
{
    return (uint32_t)((uint64_t)k.key(i) >> 32);
}

template <class KeyT> __device__ __forceinline__ uint32_t wb_group_of(const KeyT* __restrict__ keys, int i)
{
    return (uint32_t)((uint64_t)__ldg(keys + i) >> 32);
}

// packed-leaf eligibility by size: fits leaf_size and (grouped trees) does not straddle groups (bvh.cu:431-437)
template <bool GROUPED, class K>
__device__ __forceinline__ bool wb_size_leaf_k(const K& k, int leaf_size, int left, int right)
{
    if (right - left + 1 > leaf_size)
        return false;
    if (GROUPED)
        return wb_group_of_k(k, left) == wb_group_of_k(k, right);
    return true;
}

template <class KeyT, bool GROUPED>
__device__ __forceinline__ bool wb_size_leaf(const KeyT* __restrict__ keys, int leaf_size, int left, int right)
{
    return wb_size_leaf_k<GROUPED>(GlobalKeys<KeyT> { keys, nullptr }, leaf_size, left, right);
}

// parent choice of the node covering sorted positions [left, right] (bvh.cu:300-334):
// true = it becomes the LEFT child of node n+right, false = the RIGHT child of node n+left-1
template <bool GROUPED, class K> __device__ __forceinline__ bool wb_goes_right_k(const K& k, int n, int left, int right)
{
    if (left == 0)
        return true;
    if (GROUPED) {  // stay inside the group when exactly one neighbour allows it (bvh.cu:305-321)
        const uint32_t gl = wb_group_of_k(k, left), gr = wb_group_of_k(k, right);
        if (gl == gr) {
            const bool right_same = (right < n - 1) && wb_group_of_k(k, right + 1) == gl;
            const bool left_same = wb_group_of_k(k, left - 1) == gl;
            if (right_same != left_same)
                return right_same;
        }
    }
    if (right == n - 1)
        return false;
    const int dr = wb_key_delta_k(k, right), dl = wb_key_delta_k(k, left - 1);
    if (dr != dl)
        return dr > dl;
    return (k.parity(left - 1) ^ k.parity(right)) != 0;
}

template <class KeyT, bool GROUPED>
__device__ __forceinline__ bool wb_goes_right(const KeyT* __restrict__ keys, const int* __restrict__ prim, int n,
                                              int left, int right)
{
    return wb_goes_right_k<GROUPED>(GlobalKeys<KeyT> { keys, prim }, n, left, right);
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


#ifndef WB_TBM
#define WB_TBM 256
#endif
constexpr int TBM = WB_TBM;   // merge threads per block
constexpr int BP = TBM * MC;  // sorted positions per block

// arrival at a block-private counter: release / acquire at CTA scope only (MEMBAR.CTA, no L1 invalidation)
__device__ __forceinline__ unsigned wb_arrive_cta(unsigned* counter, unsigned add)
{
    __threadfence_block();
    const unsigned old = atomicAdd(counter, add);
    __threadfence_block();
    return old;
}

// builder: the node in hand covers [0, n-1]
template <class KeyT, bool GROUPED>
__device__ __forceinline__ void wb_write_root(const MergeArgs<KeyT>& a, uint32_t xnode, unsigned xh, float3 lo, float3 hi)
{
    const int n = a.n;
    const uint32_t self_ref = xnode | (wb_size_leaf<KeyT, GROUPED>(a.keys, a.leaf_size, 0, n - 1) ? WB_LEAF : 0u);
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
}

// the node in hand -- a child of n+s, the left one when go_right -- absorbs its sibling's record (s0, s1) and
// becomes n+s
template <bool REFIT, class KeyT, bool GROUPED, class K>
__device__ __forceinline__ void wb_absorb(const MergeArgs<KeyT>& a, const K& kv, int s, bool go_right, float4 s0, float4 s1,
                                          unsigned other_h, float3& lo, float3& hi, int& xl, int& xr, unsigned& xh)
{
    const int far_end = (int)__float_as_uint(s1.w);
    const int new_left = go_right ? xl : far_end;
    const int new_right = go_right ? far_end : xr;
    if (!REFIT) {
        // visible packed leaves: children that qualify by size while this parent does not
        if (!wb_size_leaf_k<GROUPED>(kv, a.leaf_size, new_left, new_right)) {
            if (wb_size_leaf_k<GROUPED>(kv, a.leaf_size, new_left, s))
                a.pos_parent[new_left] = a.n + s;
            if (wb_size_leaf_k<GROUPED>(kv, a.leaf_size, s + 1, new_right))
                a.pos_parent[s + 1] = a.n + s;
        }
        xh = max(xh, other_h) + 1u;
        a.heights[s] = wb_pack_height(xh, new_right - new_left + 1);
    }
    lo = wb_min3(lo, make_float3(s0.x, s0.y, s0.z));
    hi = wb_max3(hi, make_float3(s1.x, s1.y, s1.z));
    xl = new_left, xr = new_right;
}

// Phase B: child `side` (0 left, 1 right) of n+s, whose record is in memory, arrives at the GLOBAL counter of
// n+s and, whenever it is the second child to do so, becomes the parent and climbs on (bvh.cu:261-393, 42-144)
template <bool REFIT, class KeyT, bool GROUPED>
__device__ __noinline__ void wb_climb(const MergeArgs<KeyT>& a, int s, int side, unsigned h)
{
    const int n = a.n;
    float3 lo = make_float3(0.f, 0.f, 0.f), hi = lo;
    int xl = 0, xr = 0;
    unsigned xh = h;
    bool loaded = false;
    for (;;) {
        const unsigned hc = min(xh, WB_HEIGHT_CAP);
        const unsigned old = wb_arrive(&a.counters[s], REFIT ? 1u : (1u | (hc << 8)));
        const bool second = REFIT ? (old & 1u) != 0u : (old & 0xffu) != 0u;
        if (!second)
            return;
        const float4* pair4 = reinterpret_cast<const float4*>(a.pairs + 2 * (size_t)s);
        if (!loaded) {  // the record this carrier starts from was written during phase A
            const float4 m0 = __ldcg(pair4 + 2 * side), m1 = __ldcg(pair4 + 2 * side + 1);
            lo = make_float3(m0.x, m0.y, m0.z);
            hi = make_float3(m1.x, m1.y, m1.z);
            const int far_end = (int)__float_as_uint(m1.w);
            xl = side ? s + 1 : far_end;
            xr = side ? far_end : s;
            loaded = true;
        }
        const float4 s0 = __ldcg(pair4 + 2 * (1 - side)), s1 = __ldcg(pair4 + 2 * (1 - side) + 1);
        wb_absorb<REFIT, KeyT, GROUPED>(a, GlobalKeys<KeyT> { a.keys, a.prim }, s, side == 0, s0, s1, old >> 8, lo, hi, xl,
                                        xr, xh);
        const uint32_t xnode = (uint32_t)(n + s);

        int ps;
        bool go_right;
        if (REFIT) {
            const int p = a.parent_int[s];
            if (p == WB_NO_PARENT) {
                a.hdr->lx = lo.x, a.hdr->ly = lo.y, a.hdr->lz = lo.z;
                a.hdr->hx = hi.x, a.hdr->hy = hi.y, a.hdr->hz = hi.z;
                return;
            }
            ps = p - n;
            go_right = (xr == ps);  // a left child's range ends at the split
            wb_store_box(a.pairs + 2 * (size_t)ps + (go_right ? 0 : 1), lo, hi);
        } else {
            if (xl == 0 && xr == n - 1) {
                wb_write_root<KeyT, GROUPED>(a, xnode, xh, lo, hi);
                return;
            }
            go_right = wb_goes_right<KeyT, GROUPED>(a.keys, a.prim, n, xl, xr);
            ps = go_right ? xr : xl - 1;
            a.parent_int[s] = n + ps;
            wb_store_rec(a.pairs + 2 * (size_t)ps + (go_right ? 0 : 1), lo, hi,
                         xnode | (wb_size_leaf<KeyT, GROUPED>(a.keys, a.leaf_size, xl, xr) ? WB_LEAF : 0u),
                         (uint32_t)(go_right ? xl : xr));
        }
        s = ps;
        side = go_right ? 0 : 1;
    }
}

// Phase A: a block owns BP consecutive sorted positions, a thread MC of them.  Every node whose range lies inside
// the block is produced here: inside a thread's chunk sequentially (no atomics), across the threads of the block
// through SHARED-memory arrival counters with CTA-scope fences.  A node that has to wait for a sibling spanning a
// block boundary stays "pending" in its counter; phase B re-announces the pending nodes (and the nodes that
// reach the splits shared with the neighbouring blocks) on the global counters and climbs the spine above.
template <bool REFIT, class KeyT, bool GROUPED>
__global__ void __launch_bounds__(TBM)
k_merge(MergeArgs<KeyT> a)
{
    // bit 0: arrival parity, bit 1: side of the first arrival, bits 8..: its height
    __shared__ unsigned scount[BP];

    const int n = a.n;
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * BP;
    const int b1 = min(b0 + BP - 1, n - 1);
    for (int k = tid; k < BP; k += TBM)
        scount[k] = 0u;
    // builder: the block's keys (one halo key each side) and primitive parities, loaded once, coalesced
    // (halo: one key to the left, WB_SMALL_MAX + 1 to the right -- a unit from k_small_nodes that starts in the block
    // may end up to WB_SMALL_MAX - 1 positions past it, and its parent choice looks one key further)
    __shared__ KeyT skeys[REFIT ? 1 : BP + 2 + WB_SMALL_MAX];
    __shared__ unsigned char spar[REFIT ? 1 : BP + 2 + WB_SMALL_MAX];
    if (!REFIT) {
        for (int k = tid; k < BP + 2 + WB_SMALL_MAX; k += TBM) {
            const long long g = (long long)b0 - 1 + k;
            if (g >= 0 && g < n) {
                skeys[k] = __ldg(a.keys + g);
                spar[k] = (unsigned char)(__ldg(a.prim + g) % 2);
            }
        }
    }
    const BlockKeys<KeyT> bk { skeys, spar, b0 - 1 };
    __syncthreads();

    const long long c0l = (long long)b0 + (long long)tid * MC;
    const bool active = c0l < n;
    const int c0 = active ? (int)c0l : n - 1;
    const int c1 = min(c0 + MC - 1, n - 1);

    // arrivals this thread owes to the GLOBAL counters in phase B: a node of the block reaching a split shared with
    // a neighbouring block (b0-1 or b1), or a leaf unit that itself spans the block boundary (refit: packed leaves)
    int dsplit[4];
    unsigned dinfo[4];  // bit 1: side, bits 8..: height
    int ndefer = 0;

    if (active) {
        int rstack[MC];       // split positions of parked nodes (each is the LEFT child of n + rstack[k])
        unsigned hstack[MC];  // their heights (build only)
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
            bool go_right = false;
            int s = 0;
            unsigned other_h = 0;
            bool resumed = false;  // true when a parked node was handed over and found its sibling waiting

            if (!have) {
                if (pos > c1) {
                    if (depth == 0)
                        break;
                    // our right neighbour is now in another thread's hands: hand the parked nodes over, top first
                    --depth;
                    s = rstack[depth];  // c0 <= s < c1: interior to the block
                    const unsigned h = min(hstack[depth], WB_HEIGHT_CAP);
                    const unsigned old = wb_arrive_cta(&scount[s - b0], 1u | (h << 8));
                    if (!(old & 1u))
                        continue;
                    // the right sibling was already there: take the parked node back in hand and merge below
                    const NodeRec L = a.pairs[2 * (size_t)s];  // our own earlier store
                    lo = make_float3(L.lx, L.ly, L.lz);
                    hi = make_float3(L.hx, L.hy, L.hz);
                    xl = (int)L.aux;
                    xr = s;
                    xnode = L.ref & WB_IDX_MASK;
                    xh = h;
                    go_right = true, other_h = old >> 8;
                    have = true, fresh = false, resumed = true;
                } else if (REFIT) {
                    // next visible leaf of this chunk (its box was refreshed by the leaf pass)
                    const int p = a.pos_parent[pos];
                    if (p == WB_NO_PARENT) {
                        ++pos;
                        continue;
                    }
                    if (p == WB_ROOT_PARENT)
                        break;  // the root is a packed leaf: the leaf pass already wrote the header box
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
                    // next unit: an original leaf (its record was written by the leaf pass), or the top of a subtree
                    // that k_small_nodes already finished (its record sits in ITS parent's pair, like a leaf's)
                    const int uw = a.unit ? a.unit[pos] : WB_UNIT_LEAF;
                    if (uw == WB_UNIT_COVERED) {
                        ++pos;
                        continue;
                    }
                    xl = pos;
                    xr = pos + ((uw >> 3) & 7);
                    xnode = uw == WB_UNIT_LEAF ? (uint32_t)pos : (uint32_t)(n + pos + (uw & 7));
                    xh = (unsigned)((uw >> 6) & 7);
                    const bool gr = wb_goes_right_k<GROUPED>(bk, n, xl, xr);
                    const NodeRec* rec = a.pairs + 2 * (size_t)(gr ? xr : xl - 1) + (gr ? 0 : 1);
                    lo = make_float3(rec->lx, rec->ly, rec->lz);
                    hi = make_float3(rec->hx, rec->hy, rec->hz);
                    pos = xr + 1;
                    have = true, fresh = true;
                }
            }

            if (!resumed) {
                // ---- the node in hand: root, or choose its parent
                if (REFIT) {
                    if (static_parent == WB_NO_PARENT) {
                        a.hdr->lx = lo.x, a.hdr->ly = lo.y, a.hdr->lz = lo.z;
                        a.hdr->hx = hi.x, a.hdr->hy = hi.y, a.hdr->hz = hi.z;
                        break;
                    }
                    s = static_parent - n;
                    go_right = (xr == s);  // a left child's range ends at the split
                } else {
                    if (xl == 0 && xr == n - 1) {
                        wb_write_root<KeyT, GROUPED>(a, xnode, xh, lo, hi);
                        break;
                    }
                    go_right = wb_goes_right_k<GROUPED>(bk, n, xl, xr);
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
                                     xnode | (wb_size_leaf_k<GROUPED>(bk, a.leaf_size, xl, xr) ? WB_LEAF : 0u),
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
                } else {
                    const unsigned h = min(xh, WB_HEIGHT_CAP);
                    if (xl < b0 || xr > b1 || s < b0 || s >= b1) {  // not a block-private merge: announced in phase B
                        dsplit[ndefer] = s;
                        dinfo[ndefer] = (go_right ? 0u : 2u) | (h << 8);
                        ++ndefer;
                        have = false;
                        continue;
                    }
                    const unsigned old = wb_arrive_cta(&scount[s - b0], 1u | (go_right ? 0u : 2u) | (h << 8));
                    if (!(old & 1u)) {
                        have = false;  // the sibling's carrier continues; parked nodes (if any) are handed over above
                        continue;
                    }
                    other_h = old >> 8;
                }
            }

            // ---- second to complete n+s: union with the sibling record and become the parent
            const float4* sibling = reinterpret_cast<const float4*>(a.pairs + 2 * (size_t)s + (go_right ? 1 : 0));
            const float4 s0 = sibling[0], s1 = sibling[1];  // written by this thread or published through scount
            wb_absorb<REFIT, KeyT, GROUPED>(a, bk, s, go_right, s0, s1, other_h, lo, hi, xl, xr, xh);
            xnode = (uint32_t)(n + s);
            static_parent = REFIT ? a.parent_int[s] : WB_NO_PARENT;
            fresh = false;
        }
    }

    // ---- phase B: everything written above becomes visible device-wide, then the pending nodes go global
    __threadfence();
    __syncthreads();
    unsigned pending = 0;  // bit k < MC: interior split c0 + k waits for a spanning sibling; bit MC + j: dsplit[j]
    if (active)
        for (int k = 0; k < MC; ++k)
            if (c0 + k < b1 && c0 + k <= c1 && (scount[c0 + k - b0] & 1u))
                pending |= 1u << k;
    pending |= ((1u << ndefer) - 1u) << MC;
    while (pending) {
        const int k = __ffs(pending) - 1;
        pending &= pending - 1;
        if (k < MC) {
            const unsigned v = scount[c0 + k - b0];
            wb_climb<REFIT, KeyT, GROUPED>(a, c0 + k, (int)((v >> 1) & 1u), v >> 8);
        } else {
            const unsigned v = dinfo[k - MC];
            wb_climb<REFIT, KeyT, GROUPED>(a, dsplit[k - MC], (int)((v >> 1) & 1u), v >> 8);
        }
    }
}
