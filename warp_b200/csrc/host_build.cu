// Host-side top-down builders (constructor "sah" = 0, "median" = 1) uploaded into the sibling-pair layout.
//
// Behavioural contract = the reference's CUDA path for these constructors: item bounds are copied to the host, the tree is
// built there top-down and uploaded (warp/native/bvh.cu:625-652, 700-752; builder warp/native/bvh.cpp:216-572).  What has
// to match for the queries to be bit-identical is (1) which items end up under which node and (2) the order of the items
// inside a leaf (the first of two equally near triangles wins, mesh.h:590-597) -- both are decided by the same library
// calls on the same predicates as the reference makes (std::nth_element on the centre along the longest axis for
// "median"; std::partition at the split plane of a 16-bucket surface-area heuristic for "sah"), so they agree whenever
// both are built against the same C++ standard library.  Node NUMBERING is not reproduced: a node is identified, as in
// the LBVH, by the sorted position after which it splits its range -- any binary tree over a contiguous item order fits
// the pair layout (children of the node that splits after position s live at pairs[2s], pairs[2s + 1]; boundaries inside
// leaves simply stay unused).  Boxes are not computed here: the uploaded topology is refitted on the device
// (bvh_refit.cu), which also fills the packed-triangle cache -- exact min / max unions, the same values the reference's
// calc_bounds produces.  Purpose: tree-quality baselines -- queries/s on an LBVH against a SAH or
// median tree of the same mesh -- and wp.Mesh(..., bvh_constructor="sah") for callers that ask for it.  Not supported on
// host-built trees: groups, the reference-layout mirror (no reference numbering), the planned wavefront refit.
#include "state.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

namespace {

struct Box {
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX };
    float hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    void grow(const float* l, const float* h)
    {
        for (int k = 0; k < 3; ++k) {
            lo[k] = (l[k] < lo[k]) ? l[k] : lo[k];
            hi[k] = (h[k] > hi[k]) ? h[k] : hi[k];
        }
    }
    void grow(const Box& b) { grow(b.lo, b.hi); }
    // 2 (xy + xz + yz), in the reference's operation order (bvh.h:111-115)
    float area() const
    {
        const float e0 = hi[0] - lo[0], e1 = hi[1] - lo[1], e2 = hi[2] - lo[2];
        return 2.0f * (e0 * e1 + e0 * e2 + e1 * e2);
    }
};

// axis of the largest |extent|, the first one on ties (vec.h:1903-1915)
int widest_axis(const float* lo, const float* hi)
{
    int axis = 0;
    float best = std::fabs(hi[0] - lo[0]);
    for (int k = 1; k < 3; ++k) {
        const float e = std::fabs(hi[k] - lo[k]);
        if (e > best)
            axis = k, best = e;
    }
    return axis;
}

struct Items {
    const float* lowers;  // n x 3
    const float* uppers;
    float centre(int item, int axis) const { return 0.5f * (lowers[3 * (size_t)item + axis] + uppers[3 * (size_t)item + axis]); }
};

constexpr int kBuckets = 16;  // SAH_NUM_BUCKETS (bvh.h)

// split plane of the binned surface-area heuristic over order[start, end) (bvh.cpp:399-508); axis by reference
float sah_plane(const Items& it, const int* order, int start, int end, const Box& range, int& axis)
{
    float clo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, chi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int i = start; i < end; ++i)
        for (int k = 0; k < 3; ++k) {
            const float c = it.centre(order[i], k);
            clo[k] = (c < clo[k]) ? c : clo[k];
            chi[k] = (c > chi[k]) ? c : chi[k];
        }
    axis = widest_axis(clo, chi);
    const float a0 = clo[axis], a1 = chi[axis];
    if (a1 <= a0)
        return a0;  // no extent: the caller's partition comes out empty and it splits in the middle
    int count[kBuckets] = {};
    Box bucket[kBuckets];
    for (int i = start; i < end; ++i) {
        const int item = order[i];
        int b = (int)(kBuckets * (it.centre(item, axis) - a0) / (a1 - a0));
        b = b < 0 ? 0 : (b >= kBuckets ? kBuckets - 1 : b);
        bucket[b].grow(it.lowers + 3 * (size_t)item, it.uppers + 3 * (size_t)item);
        count[b]++;
    }
    // areas and counts of the kBuckets - 1 ways to cut between buckets, swept from both ends
    float area_below[kBuckets - 1], area_above[kBuckets - 1];
    int n_below[kBuckets - 1], n_above[kBuckets - 1];
    Box below, above;
    int cb = 0, ca = 0;
    for (int i = 0; i < kBuckets - 1; ++i) {
        below.grow(bucket[i]);
        above.grow(bucket[kBuckets - 1 - i]);
        area_below[i] = below.area();
        area_above[kBuckets - 2 - i] = above.area();
        cb += count[i];
        ca += count[kBuckets - 1 - i];
        n_below[i] = cb;
        n_above[kBuckets - 2 - i] = ca;
    }
    const float inv_total = 1.0f / range.area();
    int best = 0;
    float best_cost = FLT_MAX;
    for (int i = 0; i < kBuckets - 1; ++i) {
        const float p_below = area_below[i] * inv_total, p_above = area_above[i] * inv_total;
        const float cost = p_below * n_below[i] + p_above * n_above[i];
        if (cost < best_cost)
            best_cost = cost, best = i;
    }
    return a0 + (best + 1) * (a1 - a0) / kBuckets;
}

struct HostTree {
    std::vector<int> order;        // primitive_indices
    std::vector<NodeRec> pairs;    // 2 (n - 1) records: ref / aux filled, boxes left to the device refit
    std::vector<int> parent_int;   // n - 1
    std::vector<int> pos_parent;   // n
    uint32_t root_ref = 0;
    int depth = 0;
};

struct Todo {
    int start, end, depth, parent_slot, side;  // range [start, end), slot of the parent (-1: the root), 0 left / 1 right
};

// A bounds3 that nothing was added to is (FLT_MAX, -FLT_MAX): an empty bucket contributes that to the sweeps exactly as in
// the reference (bounds_union with a default-constructed bounds3 is the identity), so areas of one-sided sweeps match too.
void build_top_down(const Items& it, int n, int leaf_size, int constructor_type, HostTree& t)
{
    t.order.resize(n);
    for (int i = 0; i < n; ++i)
        t.order[i] = i;
    t.pairs.assign(n > 1 ? 2 * (size_t)(n - 1) : 0, NodeRec {});
    t.parent_int.assign(n > 1 ? n - 1 : 0, WB_NO_PARENT);
    t.pos_parent.assign(n, WB_NO_PARENT);
    std::vector<Todo> stack;
    stack.push_back({ 0, n, 0, -1, 0 });
    int* order = t.order.data();
    while (!stack.empty()) {
        const Todo w = stack.back();
        stack.pop_back();
        t.depth = std::max(t.depth, w.depth);
        const int count = w.end - w.start;
        // a leaf: small enough, or as deep as the query stack allows (bvh.cpp:533-541, BVH_QUERY_STACK_SIZE)
        const bool leaf = count <= leaf_size || w.depth >= WB_QUERY_STACK;
        int split = -1;
        if (!leaf) {
            Box range;
            for (int i = w.start; i < w.end; ++i)
                range.grow(it.lowers + 3 * (size_t)order[i], it.uppers + 3 * (size_t)order[i]);
            if (constructor_type == 0) {
                int axis = 0;
                const float plane = sah_plane(it, order, w.start, w.end, range, axis);
                int* mid = std::partition(order + w.start, order + w.end, [&](int item) { return it.centre(item, axis) < plane; });
                split = (int)(mid - order);
            } else {
                const int axis = widest_axis(range.lo, range.hi);
                split = (w.start + w.end) / 2;
                std::nth_element(order + w.start, order + split, order + w.end,
                                 [&](int a, int b) { return it.centre(a, axis) < it.centre(b, axis); });
            }
            if (split == w.start || split == w.end)
                split = (w.start + w.end) / 2;  // the partition failed: cut in the middle (bvh.cpp:558-561)
        }
        // this node's identity: an inner node is the slot of the position it splits after; a leaf is its first position
        const int slot = leaf ? -1 : split - 1;
        const uint32_t ref = leaf ? (WB_LEAF | (uint32_t)w.start) : (uint32_t)(n + slot);
        if (w.parent_slot < 0) {
            t.root_ref = ref;
            if (leaf)
                t.pos_parent[w.start] = WB_ROOT_PARENT;
        } else {
            NodeRec& rec = t.pairs[2 * (size_t)w.parent_slot + w.side];
            rec.ref = ref;
            rec.aux = (uint32_t)(w.side == 0 ? w.start : w.end - 1);  // the far end of the child's range
            if (leaf)
                t.pos_parent[w.start] = n + w.parent_slot;
            else
                t.parent_int[slot] = n + w.parent_slot;
        }
        if (!leaf) {
            stack.push_back({ split, w.end, w.depth + 1, slot, 1 });
            stack.push_back({ w.start, split, w.depth + 1, slot, 0 });
        }
    }
}

__global__ void k_triangle_bounds(const float* __restrict__ points, const int* __restrict__ indices, int n, float* __restrict__ lowers,
                                  float* __restrict__ uppers)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
        return;
    float3 lo, hi;
    MeshSource { points, indices }.bounds(t, lo, hi);
    lowers[3 * (size_t)t] = lo.x, lowers[3 * (size_t)t + 1] = lo.y, lowers[3 * (size_t)t + 2] = lo.z;
    uppers[3 * (size_t)t] = hi.x, uppers[3 * (size_t)t + 1] = hi.y, uppers[3 * (size_t)t + 2] = hi.z;
}

#define WB_TRY(expr)                       \
    do {                                   \
        cudaError_t _e = (expr);           \
        if (_e != cudaSuccess)             \
            return cudaGetErrorString(_e); \
    } while (0)

}  // namespace

// Host-only entry point (no device needed): the item order and the visible leaves of the sah / median tree over the given
// boxes -- what tests/test_host_builders_cpu.py pins against the reference's own host builder on every CPU-only run.
// order_out[n] = primitive_indices; leaf_start_out[n] = 1 where a leaf starts at that sorted position.  Returns the depth.
extern "C" __attribute__((visibility("default"))) int wp_b200_host_build_order(const float* lowers, const float* uppers, int n,
                                                                               int leaf_size, int constructor_type,
                                                                               int* order_out, unsigned char* leaf_start_out)
{
    if (n <= 0 || leaf_size < 1 || (constructor_type != 0 && constructor_type != 1))
        return -1;
    HostTree t;
    build_top_down(Items { lowers, uppers }, n, leaf_size, constructor_type, t);
    for (int i = 0; i < n; ++i) {
        order_out[i] = t.order[i];
        if (leaf_start_out)
            leaf_start_out[i] = t.pos_parent[i] != WB_NO_PARENT ? 1 : 0;
    }
    return t.depth;
}

// Builds the tree of an allocated BvhState (wb_alloc_tree) with the host constructor s.constructor_type (0 sah, 1 median),
// uploads it and refits it.  Synchronises `stream` (the item bounds have to reach the host first, as in the reference).
const char* wb_build_host(BvhState& s, cudaStream_t stream)
{
    const int n = s.n;
    if (n <= 0)
        return nullptr;
    if (s.groups)
        return "grouped trees are built with constructor 'lbvh' only";
    std::vector<float> lowers(3 * (size_t)n), uppers(3 * (size_t)n);
    if (s.is_mesh) {
        // triangle bounds on the device (min / max of the vertices, mesh.cu:16-36), then to the host; the sort's spare
        // buffers hold them meanwhile (keys_alt + prim_alt are 8 n bytes -- not enough: use a temporary allocation)
        float* d_bounds = nullptr;
        WB_TRY(cudaMallocAsync((void**)&d_bounds, 24 * (size_t)n, stream));
        k_triangle_bounds<<<wb_div_up(n, 256), 256, 0, stream>>>(s.points, s.indices, n, d_bounds, d_bounds + 3 * (size_t)n);
        WB_TRY(cudaMemcpyAsync(lowers.data(), d_bounds, 12 * (size_t)n, cudaMemcpyDeviceToHost, stream));
        WB_TRY(cudaMemcpyAsync(uppers.data(), d_bounds + 3 * (size_t)n, 12 * (size_t)n, cudaMemcpyDeviceToHost, stream));
        WB_TRY(cudaFreeAsync(d_bounds, stream));
    } else {
        WB_TRY(cudaMemcpyAsync(lowers.data(), s.item_lowers, 12 * (size_t)n, cudaMemcpyDeviceToHost, stream));
        WB_TRY(cudaMemcpyAsync(uppers.data(), s.item_uppers, 12 * (size_t)n, cudaMemcpyDeviceToHost, stream));
    }
    WB_TRY(cudaStreamSynchronize(stream));

    HostTree t;
    build_top_down(Items { lowers.data(), uppers.data() }, n, s.leaf_size, s.constructor_type, t);

    TreeHeader h;
    memset(&h, 0, sizeof(h));
    h.root_ref = t.root_ref, h.root_count = (uint32_t)n, h.height = t.depth, h.deep = 0, h.n = n, h.leaf_size = s.leaf_size;
    WB_TRY(cudaMemcpyAsync(s.header, &h, sizeof(h), cudaMemcpyHostToDevice, stream));
    WB_TRY(cudaMemcpyAsync(s.prim, t.order.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, stream));
    WB_TRY(cudaMemcpyAsync(s.pos_parent, t.pos_parent.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, stream));
    if (n > 1) {
        WB_TRY(cudaMemcpyAsync(s.pairs, t.pairs.data(), sizeof(NodeRec) * 2 * (size_t)(n - 1), cudaMemcpyHostToDevice, stream));
        WB_TRY(cudaMemcpyAsync(s.parent_int, t.parent_int.data(), 4 * (size_t)(n - 1), cudaMemcpyHostToDevice, stream));
        WB_TRY(cudaMemsetAsync(s.counters, 0, sizeof(unsigned) * (size_t)(n - 1), stream));
    }
    s.plan_valid = false;
    s.host_built = true;
    WB_TRY(cudaStreamSynchronize(stream));  // the host vectors go away with this frame
    return wb_refit(s, stream);             // boxes of every node + the packed-triangle cache
}
