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

static int test(int ray, const float* qa, const float* qb, const float* lo, const float* hi, float max_dist)
{
    return ray ? ray_box(qa, qb, lo, hi, max_dist) : overlap(qa, qb, lo, hi);
}

/* One query.  out may be NULL (count only).  Returns the number of hits. */
static int query_one(const half_t* lowers, const half_t* uppers, const int* prim, int root, const float* item_lowers,
                     const float* item_uppers, int ray, const float* qa, const float* qb_in, float max_dist, int* out)
{
    float qb[3] = { qb_in[0], qb_in[1], qb_in[2] };
    if (ray)
        qb[0] = 1.0f / qb[0], qb[1] = 1.0f / qb[1], qb[2] = 1.0f / qb[2];
    int stack[64];
    int count = 1, found = 0;
    stack[0] = root;
    while (count) {
        const int node = stack[--count];
        const half_t lo = lowers[node], hi = uppers[node];
        if (!test(ray, qa, qb, &lo.x, &hi.x, max_dist))
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
                    if (test(ray, qa, qb, item_lowers + 3 * item, item_uppers + 3 * item, max_dist)) {
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
 * roots (optional): per-query start node; -1 = the tree root (bvh.h:504) */
void orc_bvh_query(const void* node_lowers, const void* node_uppers, const int* prim, int root, const float* item_lowers,
                   const float* item_uppers, int ray, const float* qa, const float* qb, const int* roots, int64_t n,
                   float max_dist, int* offsets, int* indices)
{
    int run = 0;
    for (int64_t i = 0; i < n; ++i) {
        offsets[i] = run;
        const int start = (roots && roots[i] != -1) ? roots[i] : root;
        run += query_one((const half_t*)node_lowers, (const half_t*)node_uppers, prim, start, item_lowers, item_uppers, ray,
                         qa + 3 * i, qb + 3 * i, max_dist, indices ? indices + run : NULL);
    }
    offsets[n] = run;
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
