/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the generic wp.Bvh queries.
 *
 * Restates the reference iterator warp/native/bvh.h:494-600 (bvh_query_aabb / bvh_query_ray +
 * bvh_query_next) over the reference's two-array node layout: depth-first, children pushed left then right,
 * node test on pop, single-item leaves reported without an item test, items of packed leaves tested one by
 * one in leaf order.  Returns, per query, the hit items in the order the iterator yields them.
 * Pinned against numpy brute force (exact set equality, the reference's own test criterion,
 * warp/tests/geometry/test_bvh.py:186-262) in tests/test_oracle.py.
 */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct {
    float x, y, z;
    uint32_t ib;
} half_t;

#define H_I(h) ((int)((h).ib & 0x7fffffffu))
#define H_B(h) ((int)((h).ib >> 31))

static inline float fmin_r(float a, float b) { return (a <= b) ? a : ((b == b) ? b : a); }
static inline float fmax_r(float a, float b) { return (a >= b) ? a : ((b == b) ? b : a); }

/* intersect_aabb_aabb, intersect.h:183-192 */
static int overlap(const float* alo, const float* ahi, const float* blo, const float* bhi)
{
    return !(alo[0] > bhi[0] || alo[1] > bhi[1] || alo[2] > bhi[2] || ahi[0] < blo[0] || ahi[1] < blo[1]
             || ahi[2] < blo[2]);
}

/* intersect_ray_aabb (intersect.h:127-152) + half-open max_dist (bvh.h:483-487) */
static int ray_box(const float* pos, const float* rcp, const float* lo, const float* hi, float max_dist)
{
    float l1 = (lo[0] - pos[0]) * rcp[0], l2 = (hi[0] - pos[0]) * rcp[0];
    float lmin = fmin_r(l1, l2), lmax = fmax_r(l1, l2);
    l1 = (lo[1] - pos[1]) * rcp[1], l2 = (hi[1] - pos[1]) * rcp[1];
    lmin = fmax_r(fmin_r(l1, l2), lmin), lmax = fmin_r(fmax_r(l1, l2), lmax);
    l1 = (lo[2] - pos[2]) * rcp[2], l2 = (hi[2] - pos[2]) * rcp[2];
    lmin = fmax_r(fmin_r(l1, l2), lmin), lmax = fmin_r(fmax_r(l1, l2), lmax);
    const int hit = (lmax >= 0.f) & (lmax >= lmin);
    return hit && !(lmin >= max_dist);
}

/* intersect_sphere_aabb, intersect.h:197-205 */
static int sphere_box(const float* c, float radius_sq, const float* lo, const float* hi)
{
    const float dx = fmax_r(fmax_r(lo[0] - c[0], c[0] - hi[0]), 0.0f);
    const float dy = fmax_r(fmax_r(lo[1] - c[1], c[1] - hi[1]), 0.0f);
    const float dz = fmax_r(fmax_r(lo[2] - c[2], c[2] - hi[2]), 0.0f);
    return dx * dx + dy * dy + dz * dz <= radius_sq;
}

/* capsule node test, bvh.h:472-482: intersect_ray_aabb_robust (intersect.h:158-181) on the box inflated by the
 * radius, called with dir = 1 / rcp_dir (only compared with zero), closed at max_dist */
static int capsule_box(const float* pos, const float* rcp, float radius, const float* lo, const float* hi, float max_dist)
{
    float lmin = -FLT_MAX, lmax = FLT_MAX;
    for (int k = 0; k < 3; ++k) {
        const float d = 1.0f / rcp[k], l = lo[k] - radius, u = hi[k] + radius;
        if (d == 0.0f) {
            if (pos[k] < l || pos[k] > u)
                return 0;
        } else {
            const float l1 = (l - pos[k]) * rcp[k], l2 = (u - pos[k]) * rcp[k];
            lmin = fmax_r(fmin_r(l1, l2), lmin);
            lmax = fmin_r(fmax_r(l1, l2), lmax);
        }
    }
    const int hit = (lmax >= 0.f) & (lmax >= lmin);
    return hit && !(lmin > max_dist);
}

/* kind: 0 aabb, 1 ray, 2 sphere, 3 capsule (BvhQueryKind, bvh.h:420-492) */
static int test(int kind, const float* qa, const float* qb, float radius, const float* lo, const float* hi, float max_dist)
{
    if (kind == 1)
        return ray_box(qa, qb, lo, hi, max_dist);
    if (kind == 2)
        return sphere_box(qa, radius * radius, lo, hi);
    if (kind == 3)
        return capsule_box(qa, qb, radius, lo, hi, max_dist);
    return overlap(qa, qb, lo, hi);
}

/* One query.  out may be NULL (count only).  Returns the number of hits. */
static int query_one(const half_t* lowers, const half_t* uppers, const int* prim, int root, const float* item_lowers,
                     const float* item_uppers, int kind, const float* qa, const float* qb_in, float radius, float max_dist,
                     int* out)
{
    float qb[3] = { qb_in[0], qb_in[1], qb_in[2] };
    if (kind == 1 || kind == 3)
        qb[0] = 1.0f / qb[0], qb[1] = 1.0f / qb[1], qb[2] = 1.0f / qb[2];
    radius = fmax_r(radius, 0.0f); /* bvh.h:537, 548 */
    int stack[64];
    int count = 1, found = 0;
    stack[0] = root;
    while (count) {
        const int node = stack[--count];
        const half_t lo = lowers[node], hi = uppers[node];
        if (!test(kind, qa, qb, radius, &lo.x, &hi.x, max_dist))
            continue;
        if (H_B(lo)) {
            const int start = H_I(lo), end = H_I(hi);
            if (end - start == 1) {
                if (out)
                    out[found] = prim[start];
                found++;
            } else {
                for (int k = start; k < end; ++k) {
                    const int item = prim[k];
                    if (test(kind, qa, qb, radius, item_lowers + 3 * item, item_uppers + 3 * item, max_dist)) {
                        if (out)
                            out[found] = item;
                        found++;
                    }
                }
            }
        } else {
            stack[count++] = H_I(lo);
            stack[count++] = H_I(hi);
        }
    }
    return found;
}

/* offsets[n+1] and (if indices != NULL) indices[offsets[n]]; call once with indices NULL to size the output.
 * roots (optional): per-query start node; -1 = the tree root (bvh.h:504).  radii: per query, kinds 2 and 3 only;
 * the sphere kind reads its centre from qa (qb is ignored) */
void orc_bvh_query_kind(const void* node_lowers, const void* node_uppers, const int* prim, int root,
                        const float* item_lowers, const float* item_uppers, int kind, const float* qa, const float* qb,
                        const float* radii, const int* roots, int64_t n, float max_dist, int* offsets, int* indices)
{
    int run = 0;
    for (int64_t i = 0; i < n; ++i) {
        offsets[i] = run;
        const int start = (roots && roots[i] != -1) ? roots[i] : root;
        run += query_one((const half_t*)node_lowers, (const half_t*)node_uppers, prim, start, item_lowers, item_uppers, kind,
                         qa + 3 * i, (kind == 2 ? qa : qb) + 3 * i, radii ? radii[i] : 0.0f, max_dist,
                         indices ? indices + run : NULL);
    }
    offsets[n] = run;
}

void orc_bvh_query(const void* node_lowers, const void* node_uppers, const int* prim, int root, const float* item_lowers,
                   const float* item_uppers, int ray, const float* qa, const float* qb, const int* roots, int64_t n,
                   float max_dist, int* offsets, int* indices)
{
    orc_bvh_query_kind(node_lowers, node_uppers, prim, root, item_lowers, item_uppers, ray ? 1 : 0, qa, qb, NULL, roots, n,
                       max_dist, offsets, indices);
}

/* get_leaf_group / lower_bound_group / upper_bound_group / lca / bvh_get_group_root, bvh.h:287-390 */
static int leaf_group(const half_t* lowers, const int* prim, const int* item_groups, int leaf)
{
    if (!item_groups)
        return 0;
    return item_groups[prim[H_I(lowers[leaf])]];
}

static int lca(int a, int b, const int* parent)
{
    int da = 0, db = 0;
    for (int t = a; t != -1; t = parent[t])
        ++da;
    for (int t = b; t != -1; t = parent[t])
        ++db;
    if (da > db) {
        int diff = da - db;
        while (diff-- && a != -1)
            a = parent[a];
    } else if (db > da) {
        int diff = db - da;
        while (diff-- && b != -1)
            b = parent[b];
    }
    while (a != b) {
        if (a == -1 || b == -1)
            return -1;
        a = parent[a];
        b = parent[b];
    }
    return a;
}

void orc_bvh_group_roots(const void* node_lowers, const int* prim, const int* parents, const int* item_groups,
                         int num_leaf_nodes, const int* group_ids, int64_t n, int* roots)
{
    const half_t* lowers = (const half_t*)node_lowers;
    for (int64_t i = 0; i < n; ++i) {
        const int group = group_ids[i];
        int lo = 0, hi = num_leaf_nodes;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (leaf_group(lowers, prim, item_groups, mid) < group)
                lo = mid + 1;
            else
                hi = mid;
        }
        if (lo == num_leaf_nodes || leaf_group(lowers, prim, item_groups, lo) != group) {
            roots[i] = -1;
            continue;
        }
        const int first = lo;
        lo = 0, hi = num_leaf_nodes;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (leaf_group(lowers, prim, item_groups, mid) <= group)
                lo = mid + 1;
            else
                hi = mid;
        }
        roots[i] = lca(first, lo - 1, parents);
    }
}
