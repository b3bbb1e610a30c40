// Batched generic BVH queries: all items whose AABB overlaps a query box, or is
// hit by a query ray, for a batch of queries against a wp.Bvh -- the broad phase of a collision loop.
//
// Behavioural contract = the reference's iterator (warp/native/bvh.h:494-600): depth-first, children pushed
// left then right (so the right subtree is reported first), node test on pop, a single-item leaf is reported
// without an item-level test (its box IS the item's box), items of a packed leaf are tested one by one in
// leaf order against item_lowers/item_uppers.  One call evaluates a whole batch; hits come back in CSR form
// (offsets[n+1], indices[total]) in exactly the order the reference iterator would yield them.
#include "query.h"
#include "state.h"
#include "merge.cuh"  // wb_goes_right

namespace {

constexpr int BQ = 128;

struct Entry2 {
    uint32_t a;  // leaf: first sorted position | WB_LEAF ; inner: internal slot s
    uint32_t b;  // leaf: item count
};

__device__ __forceinline__ bool overlap_aabb(float3 alo, float3 ahi, float3 blo, float3 bhi)
{
    // intersect_aabb_aabb (intersect.h:183-192)
    return !(alo.x > bhi.x || alo.y > bhi.y || alo.z > bhi.z || ahi.x < blo.x || ahi.y < blo.y || ahi.z < blo.z);
}

__device__ __forceinline__ bool ray_box(float3 pos, float3 rcp, float3 lo, float3 hi, float max_dist)
{
    // intersect_ray_aabb (intersect.h:127-152) + the half-open max_dist bound of bvh_query_test<RAY> (bvh.h:483-487)
    float l1 = (lo.x - pos.x) * rcp.x, l2 = (hi.x - pos.x) * rcp.x;
    float lmin = fminf(l1, l2), lmax = fmaxf(l1, l2);
    l1 = (lo.y - pos.y) * rcp.y, l2 = (hi.y - pos.y) * rcp.y;
    lmin = fmaxf(fminf(l1, l2), lmin), lmax = fminf(fmaxf(l1, l2), lmax);
    l1 = (lo.z - pos.z) * rcp.z, l2 = (hi.z - pos.z) * rcp.z;
    lmin = fmaxf(fminf(l1, l2), lmin), lmax = fminf(fmaxf(l1, l2), lmax);
    const bool hit = (lmax >= 0.f) & (lmax >= lmin);
    return hit && !(lmin >= max_dist);
}

// sphere-AABB overlap (intersect.h:197-205): squared distance from the centre to the box <= radius^2
__device__ __forceinline__ bool sphere_box(float3 c, float radius_sq, float3 lo, float3 hi)
{
    const float dx = fmaxf(fmaxf(lo.x - c.x, c.x - hi.x), 0.0f);
    const float dy = fmaxf(fmaxf(lo.y - c.y, c.y - hi.y), 0.0f);
    const float dz = fmaxf(fmaxf(lo.z - c.z, c.z - hi.z), 0.0f);
    return dx * dx + dy * dy + dz * dz <= radius_sq;
}

// capsule node test (bvh.h:472-482): the box inflated by the radius against the robust slab test (intersect.h:158-181),
// closed at max_dist.  `dir0` is what the reference passes as the direction: 1 / (1 / dir), only compared with zero.
__device__ __forceinline__ bool capsule_box(float3 pos, float3 rcp, float3 dir0, float radius, float3 lo, float3 hi,
                                            float max_dist)
{
    float lmin = -FLT_MAX, lmax = FLT_MAX;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float d = wb_get(dir0, k), o = wb_get(pos, k), l = wb_get(lo, k) - radius, u = wb_get(hi, k) + radius;
        if (d == 0.0f) {
            if (o < l || o > u)
                return false;
        } else {
            const float r = wb_get(rcp, k);
            const float l1 = (l - o) * r, l2 = (u - o) * r;
            lmin = fmaxf(fminf(l1, l2), lmin);
            lmax = fminf(fmaxf(l1, l2), lmax);
        }
    }
    const bool hit = (lmax >= 0.f) & (lmax >= lmin);
    return hit && !(lmin > max_dist);
}

// query kinds of the generic iterator (bvh.h:420-492)
constexpr int KIND_AABB = 0, KIND_RAY = 1, KIND_SPHERE = 2, KIND_CAPSULE = 3;

struct QueryArgs {
    float3 qa, qb;  // AABB: lower, upper; RAY / CAPSULE: start, 1 / dir; SPHERE: centre
    float3 dir0;    // CAPSULE: 1 / (1 / dir)
    float radius, radius_sq;
};

template <int KIND>
__device__ __forceinline__ bool test_box(const QueryArgs& q, float3 lo, float3 hi, float max_dist)
{
    if (KIND == KIND_RAY)
        return ray_box(q.qa, q.qb, lo, hi, max_dist);
    if (KIND == KIND_SPHERE)
        return sphere_box(q.qa, q.radius_sq, lo, hi);
    if (KIND == KIND_CAPSULE)
        return capsule_box(q.qa, q.qb, q.dir0, q.radius, lo, hi, max_dist);
    return overlap_aabb(q.qa, q.qb, lo, hi);
}

__device__ __forceinline__ float3 ld3(const float* __restrict__ p, size_t i)
{
    return make_float3(__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2));
}

// RAY: (qa, qb) = (start, 1/dir);  AABB: (qa, qb) = (lower, upper).  FILL = false counts, true writes indices.
// MESH: the tree is a wp.Mesh's (mesh_query_aabb, mesh.h:2476-2712): an item's box is its triangle's, taken from
// the packed-triangle cache (== mesh.lowers/uppers of the last build / refit, mesh.cu:16-36)
// SPHERE: (qa, radii) = (centre, radius);  CAPSULE: (qa, qb, radii) = (start, dir, radius), closed at max_dist.
template <int KIND, bool FILL, bool MESH>
__global__ void __launch_bounds__(BQ)
k_bvh_query(TreeView tv, const float* __restrict__ item_lowers, const float* __restrict__ item_uppers,
            const float* __restrict__ qa_in, const float* __restrict__ qb_in, const float* __restrict__ radii,
            const int* __restrict__ roots, long long nq, float max_dist, int* __restrict__ counts, const int* __restrict__ offsets, int* __restrict__ indices)
{
    const TreeHeader h = *tv.header;
    for (long long i = (long long)blockIdx.x * BQ + threadIdx.x; i < nq; i += (long long)gridDim.x * BQ) {
        QueryArgs q;
        q.qa = ld3(qa_in, (size_t)i);
        q.qb = KIND == KIND_SPHERE ? q.qa : ld3(qb_in, (size_t)i);
        q.dir0 = q.qb, q.radius = 0.f, q.radius_sq = 0.f;
        if (KIND == KIND_RAY || KIND == KIND_CAPSULE)
            q.qb = make_float3(1.0f / q.qb.x, 1.0f / q.qb.y, 1.0f / q.qb.z);  // bvh_query_ray stores 1/dir (bvh.h:526-531)
        if (KIND == KIND_CAPSULE)
            q.dir0 = make_float3(1.0f / q.qb.x, 1.0f / q.qb.y, 1.0f / q.qb.z);  // bvh.h:476-478
        if (KIND == KIND_SPHERE || KIND == KIND_CAPSULE) {
            q.radius = fmaxf(__ldg(radii + i), 0.0f);  // bvh.h:537-538, 548-549
            q.radius_sq = q.radius * q.radius;
        }
        int found = 0;
        int* out = FILL ? indices + offsets[i] : nullptr;

        Entry2 stack[WB_QUERY_STACK];
        int top = 0;
        // start node: the tree root, or the caller's `root` (bvh.h:504: root == -1 ? *bvh.root : root).  A node's own
        // box and leaf flag live in its parent's pair.
        float3 rlo = make_float3(h.lx, h.ly, h.lz), rhi = make_float3(h.hx, h.hy, h.hz);
        Entry2 start;
        if (h.root_ref & WB_LEAF)
            start.a = WB_LEAF | 0u, start.b = h.root_count;
        else
            start.a = (h.root_ref & WB_IDX_MASK) - (uint32_t)tv.n, start.b = 0;
        const int r = roots ? __ldg(roots + i) : -1;
        if (r >= 0 && r < tv.n) {  // an original leaf: one item
            start.a = (uint32_t)r | WB_LEAF, start.b = 1u;
            if (MESH) {
                const float4* t = tv.tris + 3 * (size_t)r;
                const float4 t0 = __ldg(t), t1 = __ldg(t + 1), t2 = __ldg(t + 2);
                const float3 p = make_float3(t0.x, t0.y, t0.z), q = make_float3(t0.w, t1.x, t1.y), w = make_float3(t1.z, t1.w, t2.x);
                rlo = wb_min3(wb_min3(p, q), w), rhi = wb_max3(wb_max3(p, q), w);
            } else {
                const int item = __ldg(tv.prim + r);
                rlo = ld3(item_lowers, (size_t)item), rhi = ld3(item_uppers, (size_t)item);
            }
        } else if (r >= tv.n) {
            wb_root_entry(tv, r, start.a, start.b, rlo, rhi);
        }
        if (test_box<KIND>(q, rlo, rhi, max_dist))
            stack[top++] = start;
        while (top) {
            const Entry2 cur = stack[--top];
            if (cur.a & WB_LEAF) {
                const uint32_t start = cur.a & WB_IDX_MASK;
                if (cur.b == 1u) {  // single-item leaf: reported without an item test (bvh.h:576-581)
                    if (FILL)
                        out[found] = __ldg(tv.prim + start);
                    ++found;
                } else {
                    for (uint32_t k = 0; k < cur.b; ++k) {
                        const int item = __ldg(tv.prim + start + k);
                        float3 ilo, ihi;
                        if (MESH) {
                            const float4* t = tv.tris + 3 * (size_t)(start + k);
                            const float4 t0 = __ldg(t), t1 = __ldg(t + 1), t2 = __ldg(t + 2);
                            const float3 p = make_float3(t0.x, t0.y, t0.z), q = make_float3(t0.w, t1.x, t1.y),
                                         r = make_float3(t1.z, t1.w, t2.x);
                            ilo = wb_min3(wb_min3(p, q), r);
                            ihi = wb_max3(wb_max3(p, q), r);
                        } else {
                            ilo = ld3(item_lowers, (size_t)item);
                            ihi = ld3(item_uppers, (size_t)item);
                        }
                        if (test_box<KIND>(q, ilo, ihi, max_dist)) {
                            if (FILL)
                                out[found] = item;
                            ++found;
                        }
                    }
                }
                continue;
            }
            const uint32_t s = cur.a;
            const float4* p4 = reinterpret_cast<const float4*>(tv.pairs + 2 * (size_t)s);
            const float4 a0 = __ldg(p4), a1 = __ldg(p4 + 1), b0 = __ldg(p4 + 2), b1 = __ldg(p4 + 3);
            const uint32_t lref = __float_as_uint(a0.w), laux = __float_as_uint(a1.w);
            const uint32_t rref = __float_as_uint(b0.w), raux = __float_as_uint(b1.w);
            // left is pushed first, right second => the right child is popped (reported) first (bvh.h:603-606);
            // a child whose box fails the test is simply not pushed (the reference pops and discards it)
            if (test_box<KIND>(q, make_float3(a0.x, a0.y, a0.z), make_float3(a1.x, a1.y, a1.z), max_dist)) {
                Entry2 e;
                if (lref & WB_LEAF)
                    e.a = laux | WB_LEAF, e.b = s - laux + 1;
                else
                    e.a = (lref & WB_IDX_MASK) - (uint32_t)tv.n, e.b = 0;
                stack[top++] = e;
            }
            if (test_box<KIND>(q, make_float3(b0.x, b0.y, b0.z), make_float3(b1.x, b1.y, b1.z), max_dist)) {
                Entry2 e;
                if (rref & WB_LEAF)
                    e.a = (s + 1) | WB_LEAF, e.b = raux - s;
                else
                    e.a = (rref & WB_IDX_MASK) - (uint32_t)tv.n, e.b = 0;
                stack[top++] = e;
            }
        }
        if (!FILL)
            counts[i] = found;
    }
}

// ---- exclusive scan of int32 counts into offsets[n+1] (three small kernels; n up to 2^31) ----
constexpr int SCAN_T = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_T * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_T)
k_scan_tiles(const int* __restrict__ in, int* __restrict__ out, long long n, long long* __restrict__ tile_sums)
{
    __shared__ int warp_sums[SCAN_T / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
            inc += t;
    }
    if (lane == 31)
        warp_sums[warp] = inc;
    __syncthreads();
    int off = 0;
    for (int w = 0; w < warp; ++w)
        off += warp_sums[w];
    int run = off + inc - sum;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n)
            out[base + k] = run;  // tile-local exclusive prefix; the tile offset is added by k_scan_add
        run += v[k];
    }
    if (threadIdx.x == SCAN_T - 1)
        tile_sums[blockIdx.x] = run;
}

__global__ void __launch_bounds__(1024)
k_scan_tile_sums(long long* tile_sums, int tiles, int* total_out)
{
    // one block: every thread owns a contiguous segment of the tile sums
    __shared__ long long seg[1024];
    const int per = (tiles + 1023) / 1024;
    const int t0 = min(tiles, (int)threadIdx.x * per), t1 = min(tiles, t0 + per);
    long long sum = 0;
    for (int t = t0; t < t1; ++t)
        sum += tile_sums[t];
    seg[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan of the 1024 segment sums
        const long long add = threadIdx.x >= (unsigned)o ? seg[threadIdx.x - o] : 0;
        __syncthreads();
        seg[threadIdx.x] += add;
        __syncthreads();
    }
    long long run = seg[threadIdx.x] - sum;
    for (int t = t0; t < t1; ++t) {
        const long long c = tile_sums[t];
        tile_sums[t] = run;
        run += c;
    }
    if (threadIdx.x == 1023) {
        const long long total = seg[1023];
        *total_out = (int)(total > 0x7fffffffll ? 0x7fffffffll : total);
    }
}

__global__ void __launch_bounds__(SCAN_T)
k_scan_add(int* __restrict__ out, long long n, const long long* __restrict__ tile_sums, const int* __restrict__ total)
{
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    const int off = (int)tile_sums[blockIdx.x];
    for (int k = threadIdx.x; k < SCAN_TILE; k += SCAN_T)
        if (base + k < n)
            out[base + k] += off;
    if (blockIdx.x == 0 && threadIdx.x == 0)
        out[n] = *total;
}

// bvh_get_group_root (bvh.h:287-390): binary search of the group's first / last sorted position (keys are sorted
// by group << 32 | code), then the lowest node covering [first, last] -- the reference walks node_parents from both
// leaves (lca); here the climb goes from the first leaf through parent_int until the range covers the last.
__global__ void __launch_bounds__(BQ)
k_group_roots(TreeView tv, const uint64_t* __restrict__ keys, const int* __restrict__ group_ids, long long nq,
              int* __restrict__ roots)
{
    const long long i = (long long)blockIdx.x * BQ + threadIdx.x;
    if (i >= nq)
        return;
    const TreeHeader h = *tv.header;
    const int n = tv.n, g = __ldg(group_ids + i);
    int first, last;
    if (!keys) {  // get_leaf_group == 0 for every leaf (bvh.h:287-292)
        first = (g == 0) ? 0 : -1;
        last = n - 1;
    } else {
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((int)(__ldg(keys + mid) >> 32) < g)
                lo = mid + 1;
            else
                hi = mid;
        }
        first = (lo == n || (int)(__ldg(keys + lo) >> 32) != g) ? -1 : lo;
        lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((int)(__ldg(keys + mid) >> 32) <= g)
                lo = mid + 1;
            else
                hi = mid;
        }
        last = lo - 1;
    }
    int root = -1;
    if (first >= 0) {
        if (first == last) {
            root = (n == 1) ? (int)(h.root_ref & WB_IDX_MASK) : first;
        } else if (first == 0 && last == n - 1) {
            root = (int)(h.root_ref & WB_IDX_MASK);
        } else {
            // parent of leaf `first`: the same rule the builder applied (merge.cuh)
            const bool gr = keys ? wb_goes_right<uint64_t, true>(keys, tv.prim, n, first, first) : true;
            int s = gr ? first : first - 1;
            // (bounded by the tree height -- the climb ends at the root at the latest; a tall subtree of many equal keys
            // is a legitimate input, so no constant cap)
            const int max_hops = h.height >= (int)WB_HEIGHT_CAP ? n : h.height + 1;  // a capped height says nothing: n bounds any climb
            for (int guard = 0; guard <= max_hops; ++guard) {
                const int l = (int)tv.pairs[2 * (size_t)s].aux, r = (int)tv.pairs[2 * (size_t)s + 1].aux;
                if (l <= first && r >= last) {
                    root = n + s;
                    break;
                }
                const int p = __ldg(tv.parent_int + s);
                if (p == WB_NO_PARENT)
                    break;
                s = p - n;
            }
        }
    }
    roots[i] = root;
}

int grid_for(long long nq)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (nq + BQ - 1) / BQ, cap = (long long)sms * 64;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

const char* wb_bvh_query(const TreeView& tv, const float* item_lowers, const float* item_uppers, int kind,
                         const float* qa, const float* qb, const float* radii, const int* roots, long long nq,
                         float max_dist, int* counts, const int* offsets, int* indices, cudaStream_t stream)
{
    if (nq <= 0)
        return nullptr;
    const int grid = grid_for(nq);
    const bool fill = offsets != nullptr;
    const bool mesh = item_lowers == nullptr;  // items are the triangles of tv.tris
#define WB_BQ_LAUNCH(K, F, M) \
    k_bvh_query<K, F, M><<<grid, BQ, 0, stream>>>(tv, item_lowers, item_uppers, qa, qb, radii, roots, nq, max_dist, counts, offsets, indices)
#define WB_BQ_KIND(K)                    \
    do {                                 \
        if (fill)                        \
            WB_BQ_LAUNCH(K, true, false);  \
        else                             \
            WB_BQ_LAUNCH(K, false, false); \
    } while (0)
    if (mesh) {
        if (kind != KIND_AABB)
            return "only AABB hit lists are defined for a wp.Mesh";
        if (fill)
            WB_BQ_LAUNCH(KIND_AABB, true, true);
        else
            WB_BQ_LAUNCH(KIND_AABB, false, true);
    } else if (kind == KIND_RAY) {
        WB_BQ_KIND(KIND_RAY);
    } else if (kind == KIND_SPHERE) {
        WB_BQ_KIND(KIND_SPHERE);
    } else if (kind == KIND_CAPSULE) {
        WB_BQ_KIND(KIND_CAPSULE);
    } else {
        WB_BQ_KIND(KIND_AABB);
    }
#undef WB_BQ_KIND
#undef WB_BQ_LAUNCH
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* wb_group_roots(const TreeView& tv, const void* keys, const int* group_ids, long long nq, int* roots,
                           cudaStream_t stream)
{
    if (nq <= 0)
        return nullptr;
    k_group_roots<<<(int)((nq + BQ - 1) / BQ), BQ, 0, stream>>>(tv, (const uint64_t*)keys, group_ids, nq, roots);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// offsets[0..n] = exclusive prefix sums of counts[0..n); scratch = ceil(n / 2048) + 1 int64 words
const char* wb_exclusive_scan(const int* counts, int* offsets, long long n, long long* scratch, cudaStream_t stream)
{
    if (n <= 0) {
        cudaMemsetAsync(offsets, 0, sizeof(int), stream);
        return nullptr;
    }
    const int tiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    int* total = reinterpret_cast<int*>(scratch + tiles);
    k_scan_tiles<<<tiles, SCAN_T, 0, stream>>>(counts, offsets, n, scratch);
    k_scan_tile_sums<<<1, 1024, 0, stream>>>(scratch, tiles, total);
    k_scan_add<<<tiles, SCAN_T, 0, stream>>>(offsets, n, scratch, total);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
