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

template <class Src, class KeyT> struct MergeArgs {
    Src src;  // item source: triangles of a mesh or caller-provided boxes
    int n;
    int leaf_size;
    const KeyT* keys;  // builder only
    const int* prim;
    NodeRec* pairs;
    int* parent_int;
    int* pos_parent;
    unsigned* counters;
    TreeHeader* hdr;
    float4* tris;  // packed-triangle cache (meshes)
};

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
constexpr int TBM = WB_TBM;   // threads per block
constexpr int BP = TBM * MC;  // sorted positions per block
static_assert(BP <= 4096, "the shared arrival word keeps 12 bits of range offset");

// shared arrival word of an interior split: bit 0 arrival parity, bit 1 side of the first arrival (1 = right
// child), bits 2..13 far end of its range (offset from the block start), bits 14..29 its height
__device__ __forceinline__ unsigned wb_pack_arrival(bool right_child, int far_off, unsigned h)
{
    return 1u | (right_child ? 2u : 0u) | ((unsigned)far_off << 2) | (h << 14);
}

// arrival at a block-private counter: release / acquire at CTA scope only (MEMBAR.CTA, no L1 invalidation)
__device__ __forceinline__ unsigned wb_arrive_cta(unsigned* counter, unsigned add)
{
    __threadfence_block();
    const unsigned old = atomicAdd(counter, add);
    __threadfence_block();
    return old;
}

// item at sorted position k: its bounds and, for meshes, its packed-triangle record (sliver flag of the
// closest-point query, mesh.h:557-564, is a per-triangle constant)
template <class Src, bool WRITE>
__device__ __forceinline__ void wb_load_item(const Src& src, int item, float4* __restrict__ tris, int k, float3& lo, float3& hi)
{
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

// builder: the node in hand covers [0, n-1]
template <class A, class KeyT, bool GROUPED>
__device__ __forceinline__ void wb_write_root(const A& a, uint32_t xnode, unsigned xh, float3 lo, float3 hi)
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

// the node in hand -- a child of n+s, the left one when go_right -- absorbs its sibling (box slo/shi, range ending
// at far_end, height other_h) and becomes n+s
template <bool REFIT, bool GROUPED, class A, class K>
__device__ __forceinline__ void wb_absorb(const A& a, const K& kv, int s, bool go_right, float3 slo, float3 shi, int far_end,
                                          unsigned other_h, float3& lo, float3& hi, int& xl, int& xr, unsigned& xh)
{
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
    }
    lo = wb_min3(lo, slo);
    hi = wb_max3(hi, shi);
    xl = new_left, xr = new_right;
}

// Phase B: child `side` (0 left, 1 right) of n+s, whose record is in memory, arrives at the GLOBAL counter of
// n+s and, whenever it is the second child to do so, becomes the parent and climbs on (bvh.cu:261-393, 42-144)
template <bool REFIT, class KeyT, bool GROUPED, class A>
__device__ __noinline__ void wb_climb(const A& a, int s, int side, unsigned h)
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
        wb_absorb<REFIT, GROUPED>(a, GlobalKeys<KeyT> { a.keys, a.prim }, s, side == 0, make_float3(s0.x, s0.y, s0.z),
                                  make_float3(s1.x, s1.y, s1.z), (int)__float_as_uint(s1.w), old >> 8, lo, hi, xl, xr, xh);
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
                wb_write_root<A, KeyT, GROUPED>(a, xnode, xh, lo, hi);
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

// dynamic shared memory of k_tree
template <bool REFIT, class KeyT> constexpr size_t wb_tree_smem()
{
    return (size_t)BP * (6 * sizeof(float) + sizeof(unsigned))
        + (REFIT ? (size_t)BP * 2 * sizeof(int) : (size_t)(BP + 2) * (sizeof(KeyT) + 1) + 16);
}

// K4: leaves + hierarchy (builder) / leaf refresh + bottom-up union (refit) in one kernel.
//
// A block owns BP consecutive sorted positions, a thread MC of them.
//  stage 1  one thread per position, coalesced: gather the item, (re)write its packed-triangle record, put its box
//           in shared memory; stage the block's keys / primitive parities (builder) or parents (refit) there too.
//  stage 2  every node whose range lies inside the block is produced from shared memory alone: inside a thread's
//           chunk sequentially with a small stack of parked nodes (a node that wants to merge to the right waits
//           until the thread itself produces its right sibling -- no atomic, no fence), across the threads of the
//           block through SHARED arrival words with CTA-scope fences.  Boxes live in one slot per position: a
//           node that goes right sits in the slot of its last position, one that goes left in the slot of its
//           first; the two uses of a slot never overlap in time.  Global memory only receives the records.
//  stage 3  a node left waiting for a sibling that spans a block boundary, or that reached a split shared with
//           a neighbouring block, is announced on the GLOBAL counter and climbs the spine above the blocks.
// The parent of a node is a function of its key range alone, so the tree is bit-identical to
// the reference's whatever the arrival order.
template <bool REFIT, class Src, class KeyT, bool GROUPED>
__global__ void __launch_bounds__(TBM)
k_tree(MergeArgs<Src, KeyT> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sbox = reinterpret_cast<float*>(smem_raw);                    // [6][BP], one box slot per position
    unsigned* scount = reinterpret_cast<unsigned*>(sbox + 6 * BP);       // [BP] arrival words
    int* sparent = reinterpret_cast<int*>(scount + BP);                  // REFIT: parent_int of the block's splits
    int* spp = sparent + BP;                                             // REFIT: pos_parent of the block
    KeyT* skeys = reinterpret_cast<KeyT*>(reinterpret_cast<uintptr_t>(scount + BP + 3) & ~(uintptr_t)15);  // builder
    unsigned char* spar = reinterpret_cast<unsigned char*>(skeys + BP + 2);

    const int n = a.n;
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * BP;
    const int b1 = min(b0 + BP - 1, n - 1);

    // ---- stage 1
    for (int k = tid; k < BP; k += TBM)
        scount[k] = 0u;
    if (REFIT) {
        for (int k = tid; k < BP; k += TBM) {
            const int g = b0 + k;
            if (g <= b1) {
                float3 lo, hi;
                wb_load_item<Src, true>(a.src, __ldg(a.prim + g), a.tris, g, lo, hi);
                sbox[0 * BP + k] = lo.x, sbox[1 * BP + k] = lo.y, sbox[2 * BP + k] = lo.z;
                sbox[3 * BP + k] = hi.x, sbox[4 * BP + k] = hi.y, sbox[5 * BP + k] = hi.z;
                spp[k] = __ldg(a.pos_parent + g);
                sparent[k] = g < n - 1 ? __ldg(a.parent_int + g) : WB_NO_PARENT;
            }
        }
    } else {
        for (int k = tid; k < BP + 2; k += TBM) {  // one halo key each side
            const long long g = (long long)b0 - 1 + k;
            if (g >= 0 && g < n && g <= (long long)b1 + 1) {
                const int item = __ldg(a.prim + g);
                skeys[k] = __ldg(a.keys + g);
                spar[k] = (unsigned char)(item % 2);
                if (k >= 1 && g <= b1) {
                    float3 lo, hi;
                    wb_load_item<Src, true>(a.src, item, a.tris, (int)g, lo, hi);
                    const int q = k - 1;
                    sbox[0 * BP + q] = lo.x, sbox[1 * BP + q] = lo.y, sbox[2 * BP + q] = lo.z;
                    sbox[3 * BP + q] = hi.x, sbox[4 * BP + q] = hi.y, sbox[5 * BP + q] = hi.z;
                    a.pos_parent[g] = WB_NO_PARENT;
                }
            }
        }
    }
    const BlockKeys<KeyT> bk { skeys, spar, b0 - 1 };
    __syncthreads();

    const long long c0l = (long long)b0 + (long long)tid * MC;
    const bool active = c0l < n;
    const int c0 = active ? (int)c0l : n - 1;
    const int c1 = min(c0 + MC - 1, n - 1);

    // arrivals this thread owes to the GLOBAL counters in stage 3: a node of the block reaching a split shared with
    // a neighbouring block (b0-1 or b1), or a leaf unit that itself spans the block boundary (refit: packed leaves)
    int dsplit[4];
    unsigned dinfo[4];  // bit 1: side, bits 8..: height
    int ndefer = 0;

    // ---- stage 2
    if (active) {
        int rstack[MC];       // parked nodes: split position (each is the LEFT child of n + rstack[k]),
        int lstack[MC];       // first position of the range,
        unsigned hstack[MC];  // height (builder only)
        int depth = 0;
        int pos = c0;

        bool have = false;  // a node is in hand
        int xl = 0, xr = 0;
        uint32_t xnode = 0;
        unsigned xh = 0;
        float3 lo = make_float3(0.f, 0.f, 0.f), hi = lo;
        int static_parent = WB_NO_PARENT;  // REFIT: parent of the node in hand

        for (;;) {
            bool go_right = false;
            int s = 0, far_end = 0;
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
                    const unsigned old = wb_arrive_cta(&scount[s - b0], wb_pack_arrival(false, lstack[depth] - b0, h));
                    if (!(old & 1u))
                        continue;
                    // the right sibling was already there: take the parked node back in hand and merge below
                    const int q = s - b0;
                    lo = make_float3(sbox[0 * BP + q], sbox[1 * BP + q], sbox[2 * BP + q]);
                    hi = make_float3(sbox[3 * BP + q], sbox[4 * BP + q], sbox[5 * BP + q]);
                    xl = lstack[depth];
                    xr = s;
                    xh = h;
                    go_right = true, other_h = old >> 14, far_end = b0 + (int)((old >> 2) & 0xfffu);
                    have = true, resumed = true;
                } else if (REFIT) {
                    // next visible leaf of this chunk: union of its items' boxes
                    const int p = spp[pos - b0];
                    if (p == WB_NO_PARENT) {
                        ++pos;
                        continue;
                    }
                    xl = pos;
                    if (p == WB_ROOT_PARENT) {
                        xr = n - 1;
                    } else if (pos <= p - n) {
                        xr = p - n;  // a left child's range ends at the split
                    } else {
                        int q = pos + 1;  // the leaf ends where the next one starts
                        while (q <= b1 && spp[q - b0] == WB_NO_PARENT)
                            ++q;
                        xr = (q <= b1 || b1 == n - 1) ? q - 1 : (int)a.pairs[2 * (size_t)(p - n) + 1].aux;
                    }
                    lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
                    for (int k = xl; k <= min(xr, b1); ++k) {
                        const int q = k - b0;
                        lo = wb_min3(lo, make_float3(sbox[0 * BP + q], sbox[1 * BP + q], sbox[2 * BP + q]));
                        hi = wb_max3(hi, make_float3(sbox[3 * BP + q], sbox[4 * BP + q], sbox[5 * BP + q]));
                    }
                    for (int k = b1 + 1; k <= xr; ++k) {  // items past the block: gathered again (their block writes the cache)
                        float3 u, v;
                        wb_load_item<Src, false>(a.src, __ldg(a.prim + k), a.tris, k, u, v);
                        lo = wb_min3(lo, u);
                        hi = wb_max3(hi, v);
                    }
                    if (p == WB_ROOT_PARENT) {  // the root is a packed leaf
                        a.hdr->lx = lo.x, a.hdr->ly = lo.y, a.hdr->lz = lo.z;
                        a.hdr->hx = hi.x, a.hdr->hy = hi.y, a.hdr->hz = hi.z;
                        break;
                    }
                    static_parent = p;
                    pos = xr + 1;
                    have = true;
                } else {
                    // next original leaf
                    const int q = pos - b0;
                    xl = xr = pos;
                    lo = make_float3(sbox[0 * BP + q], sbox[1 * BP + q], sbox[2 * BP + q]);
                    hi = make_float3(sbox[3 * BP + q], sbox[4 * BP + q], sbox[5 * BP + q]);
                    xnode = (uint32_t)pos;
                    xh = 0;
                    ++pos;
                    have = true;
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
                        wb_write_root<MergeArgs<Src, KeyT>, KeyT, GROUPED>(a, xnode, xh, lo, hi);
                        break;
                    }
                    go_right = wb_goes_right_k<GROUPED>(bk, n, xl, xr);
                    s = go_right ? xr : xl - 1;
                    if (xnode >= (uint32_t)n)
                        a.parent_int[xnode - n] = n + s;
                }

                // the record goes to memory; the box also to the node's shared slot when a block-mate may need it
                NodeRec* mine = a.pairs + 2 * (size_t)s + (go_right ? 0 : 1);
                if (REFIT)
                    wb_store_box(mine, lo, hi);
                else
                    wb_store_rec(mine, lo, hi, xnode | (wb_size_leaf_k<GROUPED>(bk, a.leaf_size, xl, xr) ? WB_LEAF : 0u),
                                 (uint32_t)(go_right ? xl : xr));
                const bool inblock = xl >= b0 && xr <= b1;
                if (inblock && (REFIT || xl != xr)) {  // a builder leaf already sits in its slot
                    const int q = (go_right ? xr : xl) - b0;
                    sbox[0 * BP + q] = lo.x, sbox[1 * BP + q] = lo.y, sbox[2 * BP + q] = lo.z;
                    sbox[3 * BP + q] = hi.x, sbox[4 * BP + q] = hi.y, sbox[5 * BP + q] = hi.z;
                }

                if (go_right && xr < c1) {  // our own next unit will become (part of) the right sibling: park
                    rstack[depth] = s;
                    lstack[depth] = xl;
                    hstack[depth] = xh;
                    ++depth;
                    have = false;
                    continue;
                }
                if (!go_right && depth > 0) {
                    // the parked top is exactly the left child of n+s: both children are in our hands
                    --depth;
                    other_h = hstack[depth];
                    far_end = lstack[depth];
                } else {
                    const unsigned h = min(xh, WB_HEIGHT_CAP);
                    if (!inblock || s < b0 || s >= b1) {  // not a block-private merge: announced in stage 3
                        dsplit[ndefer] = s;
                        dinfo[ndefer] = (go_right ? 0u : 2u) | (h << 8);
                        ++ndefer;
                        have = false;
                        continue;
                    }
                    const unsigned old =
                        wb_arrive_cta(&scount[s - b0], wb_pack_arrival(!go_right, (go_right ? xl : xr) - b0, h));
                    if (!(old & 1u)) {
                        have = false;  // the sibling's carrier continues; parked nodes (if any) are handed over above
                        continue;
                    }
                    other_h = old >> 14;
                    far_end = b0 + (int)((old >> 2) & 0xfffu);
                }
            }

            // ---- second to complete n+s: union with the sibling's slot and become the parent
            {
                const int q = (go_right ? s + 1 : s) - b0;
                const float3 slo = make_float3(sbox[0 * BP + q], sbox[1 * BP + q], sbox[2 * BP + q]);
                const float3 shi = make_float3(sbox[3 * BP + q], sbox[4 * BP + q], sbox[5 * BP + q]);
                wb_absorb<REFIT, GROUPED>(a, bk, s, go_right, slo, shi, far_end, other_h, lo, hi, xl, xr, xh);
            }
            xnode = (uint32_t)(n + s);
            static_parent = REFIT ? sparent[s - b0] : WB_NO_PARENT;
        }
    }

    // ---- stage 3: everything written above becomes visible device-wide, then the pending nodes go global
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
            wb_climb<REFIT, KeyT, GROUPED>(a, c0 + k, (int)((v >> 1) & 1u), v >> 14);
        } else {
            const unsigned v = dinfo[k - MC];
            wb_climb<REFIT, KeyT, GROUPED>(a, dsplit[k - MC], (int)((v >> 1) & 1u), v >> 8);
        }
    }
}

// host side: one launch; opts the instantiation into its dynamic shared memory once
template <bool REFIT, class Src, class KeyT, bool GROUPED>
inline cudaError_t wb_launch_tree(const MergeArgs<Src, KeyT>& a, cudaStream_t stream)
{
    constexpr size_t smem = wb_tree_smem<REFIT, KeyT>();
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_tree<REFIT, Src, KeyT, GROUPED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess)
            return e;
        configured[dev] = true;
    }
    k_tree<REFIT, Src, KeyT, GROUPED><<<wb_div_up(a.n, BP), TBM, smem, stream>>>(a);
    return cudaGetLastError();
}
