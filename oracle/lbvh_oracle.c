/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the mesh/BVH spatial-query path.
 *
 * A plain-C, single-purpose restatement of the reference algorithms (NVIDIA/warp 1.17.0.dev4),
 * written from their behaviour, not translated from their source.  Each function names the
 * reference lines whose results it must reproduce.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; nothing in warp_b200/
 * does, and the product fails loudly when its CUDA library is missing.
 *
 * Parity pins (see DESIGN.md "Oracle"):
 *   - queries / closest-point / ray-triangle / slab tests: pinned against the reference's own
 *     C++ (oracle/_ref/libwarp_ref_cpu.so, built from /root/reference/warp/native) on the golden
 *     vectors of warp/tests/geometry/test_mesh.py and on random trees (tests/test_oracle.py);
 *   - LBVH (Morton keys, sorted order, topology): the reference LBVH is CUDA-only
 *     (warp/native/bvh.cpp:226-233 rejects it on the host); pinned by traversing the tree this
 *     file builds with the reference's own query code and against arrays dumped on a B200 from the
 *     reference's own bvh.cu (full reference build, baseline/ref_cuda.py golden ->
 *     tests/golden/golden_ref_lbvh.npz): all 2N-1 nodes equal bit for bit.
 *
 * Float semantics: IEEE binary32, round-to-nearest, NO fused multiply-add except where the
 * reference calls fmaf() explicitly (intersect.h:334-341).  Build with -ffp-contract=off and no
 * -mfma so gcc cannot contract.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_STACK 32 /* BVH_QUERY_STACK_SIZE, bvh.h:18 */

/* one 16-byte half node, bvh.h:161-174: xyz + (31-bit index | leaf bit in the top bit) */
typedef struct {
    float x, y, z;
    uint32_t ib;
} orc_half;

#define HALF_I(h) ((int)((h).ib & 0x7fffffffu))
#define HALF_B(h) ((int)((h).ib >> 31))

typedef struct {
    float x, y, z;
} v3;

/* scalar min/max of the reference's host build (builtin.h:708-722): a<=b ? a : (b==b ? b : a) */
static inline float fmin_ref(float a, float b) { return (a <= b) ? a : ((b == b) ? b : a); }
static inline float fmax_ref(float a, float b) { return (a >= b) ? a : ((b == b) ? b : a); }

static inline v3 v3_make(float x, float y, float z)
{
    v3 r = { x, y, z };
    return r;
}
static inline v3 v3_ld(const float* p, int64_t i) { return v3_make(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_scale(float s, v3 a) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* vec.h:518-521 */
static inline v3 v3_cross(v3 a, v3 b)                                                /* vec.h:1121-1124 */
{
    return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline v3 v3_min(v3 a, v3 b) { return v3_make(fmin_ref(a.x, b.x), fmin_ref(a.y, b.y), fmin_ref(a.z, b.z)); }
static inline v3 v3_max(v3 a, v3 b) { return v3_make(fmax_ref(a.x, b.x), fmax_ref(a.y, b.y), fmax_ref(a.z, b.z)); }
static inline float v3_get(v3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

/* ------------------------------------------------------------------------------------------ */
/* build                                                                                       */
/* ------------------------------------------------------------------------------------------ */

/* per-triangle AABB -- mesh.cu:16-36 */
void orc_triangle_bounds(const float* points, const int* indices, int num_tris, float* lowers, float* uppers)
{
    for (int t = 0; t < num_tris; ++t) {
        v3 p = v3_ld(points, indices[3 * t + 0]);
        v3 q = v3_ld(points, indices[3 * t + 1]);
        v3 r = v3_ld(points, indices[3 * t + 2]);
        v3 lo = v3_min(v3_min(p, q), r);
        v3 hi = v3_max(v3_max(p, q), r);
        lowers[3 * t + 0] = lo.x, lowers[3 * t + 1] = lo.y, lowers[3 * t + 2] = lo.z;
        uppers[3 * t + 0] = hi.x, uppers[3 * t + 1] = hi.y, uppers[3 * t + 2] = hi.z;
    }
}

/* scene AABB and per-axis 1/(extent + 1e-4) -- bvh.cu:449-488 (min/max are order independent) */
void orc_total_bounds(const float* lowers, const float* uppers, int n, float* total_lower, float* total_upper,
                      float* inv_edges)
{
    v3 lo = v3_make(FLT_MAX, FLT_MAX, FLT_MAX);
    v3 hi = v3_make(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int i = 0; i < n; ++i) {
        lo = v3_min(lo, v3_ld(lowers, i));
        hi = v3_max(hi, v3_ld(uppers, i));
    }
    total_lower[0] = lo.x, total_lower[1] = lo.y, total_lower[2] = lo.z;
    total_upper[0] = hi.x, total_upper[1] = hi.y, total_upper[2] = hi.z;
    float ex = (hi.x - lo.x) + 0.0001f, ey = (hi.y - lo.y) + 0.0001f, ez = (hi.z - lo.z) + 0.0001f;
    inv_edges[0] = 1.0f / ex, inv_edges[1] = 1.0f / ey, inv_edges[2] = 1.0f / ez;
}

/* spread the low 10 bits of n to every third bit -- bvh.h:257-265 */
static inline uint32_t spread3(uint32_t n)
{
    n = (n ^ (n << 16)) & 0xff0000ffu;
    n = (n ^ (n << 8)) & 0x0300f00fu;
    n = (n ^ (n << 4)) & 0x030c30c3u;
    n = (n ^ (n << 2)) & 0x09249249u;
    return n;
}

static inline uint32_t quant1024(float x)
{
    int q = (int)(x * 1024.0f); /* C truncation, bvh.h:270 */
    if (q < 0)
        q = 0;
    if (q > 1023)
        q = 1023;
    return (uint32_t)q;
}

/* 30-bit Morton code on a 1024^3 grid -- bvh.h:268-275 */
uint32_t orc_morton3_1024(float x, float y, float z)
{
    return (spread3(quant1024(z)) << 2) | (spread3(quant1024(y)) << 1) | spread3(quant1024(x));
}

/* 21 bits per axis on a 2 097 152^3 grid: the quality option of the B200 builder (no reference counterpart;
 * same centroid / scale arithmetic as bvh.cu:198-207, only the quantisation and interleave are wider) */
static inline uint64_t spread3_21(uint64_t v)
{
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
static inline uint64_t quant2m(float x)
{
    int q = (int)(x * 2097152.0f);
    if (q < 0)
        q = 0;
    if (q > 2097151)
        q = 2097151;
    return (uint64_t)q;
}
uint64_t orc_morton3_63(float x, float y, float z)
{
    return (spread3_21(quant2m(z)) << 2) | (spread3_21(quant2m(y)) << 1) | spread3_21(quant2m(x));
}

/* key = group << 32 | morton(centroid) -- bvh.cu:184-214; morton_bits 63 selects the wide code (ungrouped only) */
void orc_morton_keys(const float* lowers, const float* uppers, int n, const float* grid_lower, const float* inv_edges,
                     const int* groups, int morton_bits, uint64_t* keys)
{
    for (int i = 0; i < n; ++i) {
        v3 lo = v3_ld(lowers, i), hi = v3_ld(uppers, i);
        float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
        float lx = (cx - grid_lower[0]) * inv_edges[0];
        float ly = (cy - grid_lower[1]) * inv_edges[1];
        float lz = (cz - grid_lower[2]) * inv_edges[2];
        uint64_t g = groups ? (uint64_t)(uint32_t)groups[i] : 0u;
        if (morton_bits == 63)
            keys[i] = orc_morton3_63(lx, ly, lz);
        else
            keys[i] = (g << 32) | (uint64_t)orc_morton3_1024(lx, ly, lz);
    }
}

/* stable ascending sort of (key, index) over all 64 key bits -- the only property the reference
 * takes from cub::DeviceRadixSort::SortPairs (sort.cu:273-300, call site bvh.cu:577). */
void orc_sort_pairs(uint64_t* keys, int* vals, int n)
{
    uint64_t* k2 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    int* v2 = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    uint64_t *ks = keys, *kd = k2;
    int *vs = vals, *vd = v2;
    for (int shift = 0; shift < 64; shift += 16) {
        size_t* count = (size_t*)calloc(65537, sizeof(size_t));
        for (int i = 0; i < n; ++i)
            count[((ks[i] >> shift) & 0xffff) + 1]++;
        int trivial = 0;
        for (int d = 0; d < 65536; ++d) {
            if (count[d + 1] == (size_t)n)
                trivial = 1;
            count[d + 1] += count[d];
        }
        if (!trivial) {
            for (int i = 0; i < n; ++i) {
                size_t dst = count[(ks[i] >> shift) & 0xffff]++;
                kd[dst] = ks[i];
                vd[dst] = vs[i];
            }
            uint64_t* tk = ks;
            ks = kd, kd = tk;
            int* tv = vs;
            vs = vd, vd = tv;
        }
        free(count);
    }
    if (ks != keys) {
        memcpy(keys, ks, sizeof(uint64_t) * (size_t)n);
        memcpy(vals, vs, sizeof(int) * (size_t)n);
    }
    free(k2);
    free(v2);
}

static inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }

/* Full LBVH build over item AABBs: bvh.cu:515-613 (stage order), 218-226 (deltas), 228-255
 * (leaves), 261-393 (bottom-up parent choice + AABB union), 402-443 (packed leaves).
 *
 * Outputs (all caller allocated): keys[n] sorted, primitive_indices[n], node_lowers/uppers
 * [2n-1] halves, parents[2n-1], *root.  Nodes are numbered as the reference numbers them:
 * leaves 0..n-1 in sorted order, the internal node that splits after sorted position s is n+s.
 * The bottom-up pass is replayed serially; the result does not depend on arrival order because
 * a node's parent is a function of its key range only.
 */
void orc_lbvh_build(const float* item_lowers, const float* item_uppers, int n, const int* groups, int leaf_size,
                    int morton_bits, uint64_t* keys, int* primitive_indices, orc_half* node_lowers, orc_half* node_uppers,
                    int* parents, int* root, float* total_lower, float* total_upper, float* inv_edges)
{
    if (n <= 0)
        return;
    const int max_nodes = 2 * n - 1;
    float tl[3], tu[3], inv[3];
    orc_total_bounds(item_lowers, item_uppers, n, tl, tu, inv);
    if (total_lower)
        memcpy(total_lower, tl, sizeof(tl));
    if (total_upper)
        memcpy(total_upper, tu, sizeof(tu));
    if (inv_edges)
        memcpy(inv_edges, inv, sizeof(inv));

    orc_morton_keys(item_lowers, item_uppers, n, tl, inv, groups, morton_bits, keys);
    const int grouped = groups != NULL; /* without groups the group rule is vacuous for 30-bit keys and must be off for 63-bit ones */
    for (int i = 0; i < n; ++i)
        primitive_indices[i] = i;
    orc_sort_pairs(keys, primitive_indices, n);

    int* deltas = (int*)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i + 1 < n; ++i)
        deltas[i] = clz64(keys[i] ^ keys[i + 1]);

    int* range_l = (int*)malloc(sizeof(int) * (size_t)max_nodes);
    int* range_r = (int*)malloc(sizeof(int) * (size_t)max_nodes);
    int* arrivals = (int*)calloc((size_t)max_nodes, sizeof(int));
    memset(node_lowers, 0, sizeof(orc_half) * (size_t)max_nodes); /* bvh.cu:720-730 */
    memset(node_uppers, 0, sizeof(orc_half) * (size_t)max_nodes);

    for (int i = 0; i < n; ++i) {
        const int item = primitive_indices[i];
        node_lowers[i].x = item_lowers[3 * item], node_lowers[i].y = item_lowers[3 * item + 1];
        node_lowers[i].z = item_lowers[3 * item + 2], node_lowers[i].ib = (uint32_t)i | 0x80000000u;
        node_uppers[i].x = item_uppers[3 * item], node_uppers[i].y = item_uppers[3 * item + 1];
        node_uppers[i].z = item_uppers[3 * item + 2], node_uppers[i].ib = (uint32_t)i;
        range_l[i] = range_r[i] = i;
    }

    for (int leaf = 0; leaf < n; ++leaf) {
        int index = leaf;
        for (;;) {
            const int left = range_l[index], right = range_r[index];
            if (left == 0 && right == n - 1) {
                *root = index;
                parents[index] = -1;
                break;
            }
            int go_right;
            if (left == 0) {
                go_right = 1;
            } else {
                int decided = 0;
                go_right = 0;
                const uint32_t gl = (uint32_t)(keys[left] >> 32), gr = (uint32_t)(keys[right] >> 32);
                if (grouped && gl == gr) { /* stay inside the group when exactly one neighbour allows it */
                    const int right_same = (right < n - 1) && ((uint32_t)(keys[right + 1] >> 32) == gl);
                    const int left_same = ((uint32_t)(keys[left - 1] >> 32) == gl);
                    if (right_same != left_same) {
                        go_right = right_same;
                        decided = 1;
                    }
                }
                if (!decided) { /* larger common prefix wins; equal -> parity of the two item ids */
                    if (right != n - 1 && deltas[right] >= deltas[left - 1]) {
                        if (deltas[right] == deltas[left - 1])
                            go_right = (primitive_indices[left - 1] % 2) ^ (primitive_indices[right] % 2);
                        else
                            go_right = 1;
                    } else {
                        go_right = 0;
                    }
                }
            }
            int parent;
            if (go_right) { /* we become the LEFT child of node n+right */
                parent = right + n;
                parents[index] = parent;
                node_lowers[parent].ib = (node_lowers[parent].ib & 0x80000000u) | (uint32_t)index;
                range_l[parent] = left;
            } else { /* we become the RIGHT child of node n+left-1 */
                parent = left + n - 1;
                parents[index] = parent;
                node_uppers[parent].ib = (node_uppers[parent].ib & 0x80000000u) | (uint32_t)index;
                range_r[parent] = right;
            }
            if (arrivals[parent]++ == 1) {
                const int lc = HALF_I(node_lowers[parent]), rc = HALF_I(node_uppers[parent]);
                node_lowers[parent].x = fmin_ref(node_lowers[lc].x, node_lowers[rc].x);
                node_lowers[parent].y = fmin_ref(node_lowers[lc].y, node_lowers[rc].y);
                node_lowers[parent].z = fmin_ref(node_lowers[lc].z, node_lowers[rc].z);
                node_lowers[parent].ib = (uint32_t)lc;
                node_uppers[parent].x = fmax_ref(node_uppers[lc].x, node_uppers[rc].x);
                node_uppers[parent].y = fmax_ref(node_uppers[lc].y, node_uppers[rc].y);
                node_uppers[parent].z = fmax_ref(node_uppers[lc].z, node_uppers[rc].z);
                node_uppers[parent].ib = (uint32_t)rc;
                index = parent;
            } else {
                break;
            }
        }
    }

    /* packed leaves: every node whose range fits leaf_size, or that sits at depth >= 32 */
    for (int node = 0; node < max_nodes; ++node) {
        int depth = 1;
        for (int p = parents[node]; p != -1; p = parents[p])
            depth++;
        const int left = range_l[node], right = range_r[node] + 1;
        const int single_group = !grouped || (keys[left] >> 32) == (keys[right - 1] >> 32);
        if (single_group && (right - left <= leaf_size || depth >= ORC_STACK)) {
            node_lowers[node].ib = 0x80000000u | (uint32_t)left;
            node_uppers[node].ib = (node_uppers[node].ib & 0x80000000u) | (uint32_t)right;
        }
    }

    free(deltas);
    free(range_l);
    free(range_r);
    free(arrivals);
}

/* bottom-up refit over an existing tree -- bvh.cu:42-144, 772-788 (serial replay) */
void orc_lbvh_refit(int n, const int* parents, const int* primitive_indices, orc_half* node_lowers,
                    orc_half* node_uppers, const float* item_lowers, const float* item_uppers)
{
    if (n <= 0)
        return;
    const int max_nodes = 2 * n - 1;
    int* arrivals = (int*)calloc((size_t)max_nodes, sizeof(int));
    for (int leaf = 0; leaf < n; ++leaf) {
        int index = leaf;
        if (!HALF_B(node_lowers[index]))
            continue;
        int parent = parents[index];
        if (parent == -1 || !HALF_B(node_lowers[parent])) {
            v3 lo = v3_make(FLT_MAX, FLT_MAX, FLT_MAX), hi = v3_make(-FLT_MAX, -FLT_MAX, -FLT_MAX);
            for (int c = HALF_I(node_lowers[index]); c < HALF_I(node_uppers[index]); ++c) {
                const int prim = primitive_indices[c];
                lo = v3_min(lo, v3_ld(item_lowers, prim));
                hi = v3_max(hi, v3_ld(item_uppers, prim));
            }
            node_lowers[index].x = lo.x, node_lowers[index].y = lo.y, node_lowers[index].z = lo.z;
            node_uppers[index].x = hi.x, node_uppers[index].y = hi.y, node_uppers[index].z = hi.z;
        }
        for (;;) {
            parent = parents[index];
            if (parent == -1)
                break;
            if (arrivals[parent]++ != 1)
                break;
            if (HALF_B(node_lowers[parent])) { /* packed leaf reached from below: rebuild from its items */
                const int pp = parents[parent];
                if (pp == -1 || !HALF_B(node_lowers[pp])) {
                    v3 lo = v3_make(FLT_MAX, FLT_MAX, FLT_MAX), hi = v3_make(-FLT_MAX, -FLT_MAX, -FLT_MAX);
                    for (int c = HALF_I(node_lowers[parent]); c < HALF_I(node_uppers[parent]); ++c) {
                        const int prim = primitive_indices[c];
                        lo = v3_min(lo, v3_ld(item_lowers, prim));
                        hi = v3_max(hi, v3_ld(item_uppers, prim));
                    }
                    node_lowers[parent].x = lo.x, node_lowers[parent].y = lo.y, node_lowers[parent].z = lo.z;
                    node_uppers[parent].x = hi.x, node_uppers[parent].y = hi.y, node_uppers[parent].z = hi.z;
                }
            } else {
                const int lc = HALF_I(node_lowers[parent]), rc = HALF_I(node_uppers[parent]);
                node_lowers[parent].x = fmin_ref(node_lowers[lc].x, node_lowers[rc].x);
                node_lowers[parent].y = fmin_ref(node_lowers[lc].y, node_lowers[rc].y);
                node_lowers[parent].z = fmin_ref(node_lowers[lc].z, node_lowers[rc].z);
                node_uppers[parent].x = fmax_ref(node_uppers[lc].x, node_uppers[rc].x);
                node_uppers[parent].y = fmax_ref(node_uppers[lc].y, node_uppers[rc].y);
                node_uppers[parent].z = fmax_ref(node_uppers[lc].z, node_uppers[rc].z);
            }
            index = parent;
        }
    }
    free(arrivals);
}

/* ------------------------------------------------------------------------------------------ */
/* geometric primitives                                                                        */
/* ------------------------------------------------------------------------------------------ */

/* squared distance point -> AABB, clamp form -- mesh.h:92-98 */
static inline float dist_aabb_sq(v3 p, const orc_half* lo, const orc_half* hi)
{
    const float dx = fmin_ref(hi->x, fmax_ref(lo->x, p.x)) - p.x;
    const float dy = fmin_ref(hi->y, fmax_ref(lo->y, p.y)) - p.y;
    const float dz = fmin_ref(hi->z, fmax_ref(lo->z, p.z)) - p.z;
    return dx * dx + dy * dy + dz * dz;
}

/* Voronoi-region closest point; returns barycentric (u, v) of vertices a, b -- intersect.h:44-109 */
void orc_closest_point_to_triangle(const float* pa, const float* pb, const float* pc, const float* pp, float* uv)
{
    const v3 a = v3_ld(pa, 0), b = v3_ld(pb, 0), c = v3_ld(pc, 0), p = v3_ld(pp, 0);
    const v3 ab = v3_sub(b, a), ac = v3_sub(c, a), ap = v3_sub(p, a);
    float v, w;
    const float d1 = v3_dot(ab, ap), d2 = v3_dot(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) { /* vertex a */
        v = 0.0f, w = 0.0f;
        goto done;
    }
    {
        const v3 bp = v3_sub(p, b);
        const float d3 = v3_dot(ab, bp), d4 = v3_dot(ac, bp);
        if (d3 >= 0.0f && d4 <= d3) { /* vertex b */
            v = 1.0f, w = 0.0f;
            goto done;
        }
        const float vc = d1 * d4 - d3 * d2;
        if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) { /* edge ab */
            v = d1 / (d1 - d3), w = 0.0f;
            goto done;
        }
        const v3 cp = v3_sub(p, c);
        const float d5 = v3_dot(ab, cp), d6 = v3_dot(ac, cp);
        if (d6 >= 0.0f && d5 <= d6) { /* vertex c */
            v = 0.0f, w = 1.0f;
            goto done;
        }
        const float vb = d5 * d2 - d1 * d6;
        if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) { /* edge ac */
            v = 0.0f, w = d2 / (d2 - d6);
            goto done;
        }
        const float va = d3 * d6 - d5 * d4;
        if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) { /* edge bc */
            w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
            v = 1.0f - w;
            goto done;
        }
        const float denom = 1.0f / (va + vb + vc); /* interior */
        v = vb * denom;
        w = vc * denom;
    }
done:
    uv[0] = 1.0f - v - w;
    uv[1] = v;
}

/* slab test with precomputed reciprocal direction -- intersect.h:127-152 */
static inline int ray_aabb_fast(v3 pos, v3 rcp, const orc_half* lo, const orc_half* hi, float* t)
{
    float l1 = (lo->x - pos.x) * rcp.x, l2 = (hi->x - pos.x) * rcp.x;
    float lmin = fmin_ref(l1, l2), lmax = fmax_ref(l1, l2);
    l1 = (lo->y - pos.y) * rcp.y, l2 = (hi->y - pos.y) * rcp.y;
    lmin = fmax_ref(fmin_ref(l1, l2), lmin), lmax = fmin_ref(fmax_ref(l1, l2), lmax);
    l1 = (lo->z - pos.z) * rcp.z, l2 = (hi->z - pos.z) * rcp.z;
    lmin = fmax_ref(fmin_ref(l1, l2), lmin), lmax = fmin_ref(fmax_ref(l1, l2), lmax);
    const int hit = (lmax >= 0.f) & (lmax >= lmin);
    if (hit)
        *t = lmin;
    return hit;
}

/* slab test that treats a zero direction component as "parallel to the slab" -- intersect.h:158-181 */
static inline int ray_aabb_robust(v3 pos, v3 dir, v3 rcp, const orc_half* lo, const orc_half* hi, float* t)
{
    float lmin = -FLT_MAX, lmax = FLT_MAX;
    const float lov[3] = { lo->x, lo->y, lo->z }, hiv[3] = { hi->x, hi->y, hi->z };
    for (int k = 0; k < 3; ++k) {
        const float d = v3_get(dir, k), o = v3_get(pos, k);
        if (d == 0.0f) {
            if (o < lov[k] || o > hiv[k])
                return 0;
        } else {
            const float r = v3_get(rcp, k);
            const float l1 = (lov[k] - o) * r, l2 = (hiv[k] - o) * r;
            lmin = fmax_ref(fmin_ref(l1, l2), lmin);
            lmax = fmin_ref(fmax_ref(l1, l2), lmax);
        }
    }
    const int hit = (lmax >= 0.f) & (lmax >= lmin);
    if (hit)
        *t = lmin;
    return hit;
}

/* a*b - c*d with the rounding error of c*d recovered by two explicit FMAs -- intersect.h:334-341 */
static inline float diff_of_products(float a, float b, float c, float d)
{
    const float cd = c * d;
    const float diff = fmaf(a, b, -cd);
    const float err = fmaf(-c, d, cd);
    return diff + err;
}

static inline float flip_sign(float x, uint32_t mask)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u ^= mask;
    memcpy(&x, &u, 4);
    return x;
}

/* watertight ray/triangle test (Woop et al.) -- intersect.h:344-444.  Returns hit; writes t, u, v,
 * sign (= determinant) and the un-normalised geometric normal cross(b-a, c-a). */
static int ray_tri_watertight(v3 p, v3 dir, v3 a, v3 b, v3 c, float* t, float* u, float* v, float* sign, v3* normal)
{
    /* dominant axis: first strictly larger |component| wins (vec.h:1903-1915) */
    int kz = 0;
    float best = fabsf(dir.x);
    if (fabsf(dir.y) > best)
        kz = 1, best = fabsf(dir.y);
    if (fabsf(dir.z) > best)
        kz = 2;
    int kx = (kz + 1) % 3, ky = (kx + 1) % 3;
    if (v3_get(dir, kz) < 0.0f) {
        const int tmp = kx;
        kx = ky, ky = tmp;
    }
    const float Sx = v3_get(dir, kx) / v3_get(dir, kz);
    const float Sy = v3_get(dir, ky) / v3_get(dir, kz);
    const float Sz = 1.0f / v3_get(dir, kz);

    const v3 A = v3_sub(a, p), B = v3_sub(b, p), C = v3_sub(c, p);
    const float Ax = v3_get(A, kx) - Sx * v3_get(A, kz), Ay = v3_get(A, ky) - Sy * v3_get(A, kz);
    const float Bx = v3_get(B, kx) - Sx * v3_get(B, kz), By = v3_get(B, ky) - Sy * v3_get(B, kz);
    const float Cx = v3_get(C, kx) - Sx * v3_get(C, kz), Cy = v3_get(C, ky) - Sy * v3_get(C, kz);

    float U = diff_of_products(Cx, By, Cy, Bx);
    float V = diff_of_products(Ax, Cy, Ay, Cx);
    float W = diff_of_products(Bx, Ay, By, Ax);
    if (U == 0.0f || V == 0.0f || W == 0.0f) { /* on an edge: redo in double */
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f))
        return 0;
    const float det = U + V + W;
    if (det == 0.0f)
        return 0;
    const float Az = Sz * v3_get(A, kz), Bz = Sz * v3_get(B, kz), Cz = Sz * v3_get(C, kz);
    const float T = U * Az + V * Bz + W * Cz;
    uint32_t det_bits;
    memcpy(&det_bits, &det, 4);
    if (flip_sign(T, det_bits & 0x80000000u) < 0.0f)
        return 0;
    const float rcp_det = 1.0f / det;
    *u = U * rcp_det;
    *v = V * rcp_det;
    *t = T * rcp_det;
    *sign = det;
    *normal = v3_cross(v3_sub(b, a), v3_sub(c, a));
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* queries                                                                                     */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const float* points;
    const int* indices;
    const orc_half* node_lowers;
    const orc_half* node_uppers;
    const int* primitive_indices;
    int root;
} orc_mesh;

typedef struct {
    uint64_t nodes_visited; /* nodes whose AABB halves were fetched */
    uint64_t tris_tested;
} orc_stats;

/* closest point without sign -- mesh.h:501-676.  Depth-first, near child popped first, strict '<'
 * updates, slivers skipped, outputs (u, v) = (1 - v - w, v) of the winning triangle. */
static int point_no_sign_one(const orc_mesh* m, v3 point, float max_dist, int* face, float* u, float* v,
                             orc_stats* st)
{
    int stack[ORC_STACK];
    int count = 1;
    stack[0] = m->root;
    float best = max_dist * max_dist;
    int best_face = 0;
    float best_v = 0.f, best_w = 0.f;

    while (count) {
        const int node = stack[--count];
        const orc_half lo = m->node_lowers[node], hi = m->node_uppers[node];
        if (st)
            st->nodes_visited++;
        if (dist_aabb_sq(point, &lo, &hi) > best)
            continue;
        const int li = HALF_I(lo), ri = HALF_I(hi);
        if (HALF_B(lo)) {
            for (int pc = li; pc < ri; ++pc) {
                const int prim = m->primitive_indices[pc];
                const v3 p = v3_ld(m->points, m->indices[3 * prim + 0]);
                const v3 q = v3_ld(m->points, m->indices[3 * prim + 1]);
                const v3 r = v3_ld(m->points, m->indices[3 * prim + 2]);
                const v3 e0 = v3_sub(q, p), e1 = v3_sub(r, p), e2 = v3_sub(r, q);
                const v3 nrm = v3_cross(e0, e1);
                if (st)
                    st->tris_tested++;
                if (sqrtf(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z)
                        / (v3_dot(e0, e0) + v3_dot(e1, e1) + v3_dot(e2, e2))
                    < 1.e-6f)
                    continue;
                float uv[2];
                orc_closest_point_to_triangle(&p.x, &q.x, &r.x, &point.x, uv);
                const float bu = uv[0], bv = uv[1], bw = 1.f - bu - bv;
                const v3 c = v3_add(v3_add(v3_scale(bu, p), v3_scale(bv, q)), v3_scale(bw, r));
                const v3 d = v3_sub(c, point);
                const float dsq = v3_dot(d, d);
                if (dsq < best) {
                    best = dsq;
                    best_v = bv;
                    best_w = bw;
                    best_face = prim;
                }
            }
        } else {
            const orc_half llo = m->node_lowers[li], lhi = m->node_uppers[li];
            const orc_half rlo = m->node_lowers[ri], rhi = m->node_uppers[ri];
            if (st)
                st->nodes_visited += 2;
            const float dl = dist_aabb_sq(point, &llo, &lhi), dr = dist_aabb_sq(point, &rlo, &rhi);
            int first, second;
            float dfirst, dsecond;
            if (dl < dr) { /* farther one is pushed first so the nearer is popped first */
                first = ri, second = li, dfirst = dr, dsecond = dl;
            } else {
                first = li, second = ri, dfirst = dl, dsecond = dr;
            }
            if (dfirst < best)
                stack[count++] = first;
            if (dsecond < best)
                stack[count++] = second;
        }
    }
    if (best < max_dist * max_dist) {
        *u = 1.0f - best_v - best_w;
        *v = best_v;
        *face = best_face;
        return 1;
    }
    return 0;
}

/* sign of the closest hit along one probe ray, push-both traversal -- mesh.h:2286-2339 */
static int ray_closest_sign(const orc_mesh* m, v3 start, v3 dir, float* out_sign, orc_stats* st)
{
    int stack[ORC_STACK];
    int size = 0;
    int node = m->root;
    const v3 rcp = v3_make(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    float min_t = FLT_MAX, tt;
    int hit = 0;
    for (;;) {
        const orc_half lo = m->node_lowers[node], hi = m->node_uppers[node];
        if (st)
            st->nodes_visited++;
        if (ray_aabb_robust(start, dir, rcp, &lo, &hi, &tt) && tt < min_t) {
            if (HALF_B(lo)) {
                for (int pc = HALF_I(lo); pc < HALF_I(hi); ++pc) {
                    const int prim = m->primitive_indices[pc];
                    const v3 p = v3_ld(m->points, m->indices[3 * prim + 0]);
                    const v3 q = v3_ld(m->points, m->indices[3 * prim + 1]);
                    const v3 r = v3_ld(m->points, m->indices[3 * prim + 2]);
                    float t, u, v, s;
                    v3 nrm;
                    if (st)
                        st->tris_tested++;
                    if (ray_tri_watertight(start, dir, p, q, r, &t, &u, &v, &s, &nrm)) {
                        if (t >= 0.0f && t < min_t) {
                            min_t = t;
                            *out_sign = s;
                            hit = 1;
                        }
                    }
                }
            } else {
                stack[size++] = HALF_I(lo);
                stack[size++] = HALF_I(hi);
            }
        }
        if (size == 0)
            break;
        node = stack[--size];
    }
    return hit;
}

/* inside/outside by majority of three axis probes -- mesh.h:2342-2359 */
static float inside_by_axis_rays(const orc_mesh* m, v3 p, orc_stats* st)
{
    int votes = 0;
    float s = 0.f;
    for (int axis = 0; axis < 3; ++axis) {
        const v3 dir = v3_make(axis == 0 ? 1.f : 0.f, axis == 1 ? 1.f : 0.f, axis == 2 ? 1.f : 0.f);
        if (ray_closest_sign(m, p, dir, &s, st) && s < 0)
            votes++;
    }
    return votes >= 2 ? -1.0f : 1.0f;
}

typedef struct {
    int leaf;
    int lo_payload;
    int hi_payload;
} ray_entry;

/* closest hit along a ray, near child first -- mesh.h:1735-1891 */
static int ray_one(const orc_mesh* m, v3 start, v3 dir, float max_t, float* t, float* u, float* v, float* sign,
                   v3* normal, int* face, orc_stats* st)
{
    ray_entry stack[ORC_STACK];
    int size = 0;
    ray_entry cur;
    cur.leaf = HALF_B(m->node_lowers[m->root]);
    cur.lo_payload = HALF_I(m->node_lowers[m->root]);
    cur.hi_payload = HALF_I(m->node_uppers[m->root]);
    if (st)
        st->nodes_visited++;

    v3 safe = dir; /* zero components patched only for the reciprocal */
    if (safe.x == 0.0f)
        safe.x = 1.0e-20f;
    if (safe.y == 0.0f)
        safe.y = 1.0e-20f;
    if (safe.z == 0.0f)
        safe.z = 1.0e-20f;
    const v3 rcp = v3_make(1.0f / safe.x, 1.0f / safe.y, 1.0f / safe.z);
    const int fast = dir.x != 0.0f && dir.y != 0.0f && dir.z != 0.0f;

    float min_t = max_t, min_u = 0.f, min_v = 0.f, min_sign = 1.0f;
    int min_face = 0, hit = 0;
    v3 min_n = v3_make(0, 0, 0);

    for (;;) {
        if (cur.leaf) {
            for (int pc = cur.lo_payload; pc < cur.hi_payload; ++pc) {
                const int prim = m->primitive_indices[pc];
                const v3 p = v3_ld(m->points, m->indices[3 * prim + 0]);
                const v3 q = v3_ld(m->points, m->indices[3 * prim + 1]);
                const v3 r = v3_ld(m->points, m->indices[3 * prim + 2]);
                float tt, tu, tv, ts;
                v3 n;
                if (st)
                    st->tris_tested++;
                if (ray_tri_watertight(start, dir, p, q, r, &tt, &tu, &tv, &ts, &n)) {
                    if (tt < min_t && tt >= 0.0f) {
                        min_t = tt, min_face = prim, min_u = tu, min_v = tv, min_sign = ts, min_n = n;
                        hit = 1;
                    }
                }
            }
            if (size == 0)
                break;
            cur = stack[--size];
            continue;
        }
        const int li = cur.lo_payload, ri = cur.hi_payload;
        const orc_half llo = m->node_lowers[li], lhi = m->node_uppers[li];
        const orc_half rlo = m->node_lowers[ri], rhi = m->node_uppers[ri];
        if (st)
            st->nodes_visited += 2;
        float t0 = FLT_MAX, t1 = FLT_MAX;
        const int h0 = (fast ? ray_aabb_fast(start, rcp, &llo, &lhi, &t0)
                             : ray_aabb_robust(start, dir, rcp, &llo, &lhi, &t0))
            && t0 < min_t;
        const int h1 = (fast ? ray_aabb_fast(start, rcp, &rlo, &rhi, &t1)
                             : ray_aabb_robust(start, dir, rcp, &rlo, &rhi, &t1))
            && t1 < min_t;
        ray_entry le = { HALF_B(llo), HALF_I(llo), HALF_I(lhi) };
        ray_entry re = { HALF_B(rlo), HALF_I(rlo), HALF_I(rhi) };
        if (h0 && h1) {
            const int near_left = t0 < t1;
            if (size >= ORC_STACK)
                break; /* overflow aborts with what was found, mesh.h:1860-1861 */
            stack[size++] = near_left ? re : le;
            cur = near_left ? le : re;
        } else if (h0) {
            cur = le;
        } else if (h1) {
            cur = re;
        } else {
            if (size == 0)
                break;
            cur = stack[--size];
        }
    }
    if (hit) {
        *u = min_u, *v = min_v, *sign = min_sign, *t = min_t, *face = min_face;
        const float l = sqrtf(min_n.x * min_n.x + min_n.y * min_n.y + min_n.z * min_n.z);
        if (l > 0.0f)
            *normal = v3_make(min_n.x / l, min_n.y / l, min_n.z / l);
        else
            *normal = v3_make(0, 0, 0);
        return 1;
    }
    return 0;
}

static orc_mesh make_mesh(const float* points, const int* indices, const orc_half* lowers, const orc_half* uppers,
                          const int* prims, int root)
{
    orc_mesh m = { points, indices, lowers, uppers, prims, root };
    return m;
}

/* batched drivers; result/face/u/v(/sign/t/normal) keep their zero defaults on a miss, like the
 * struct-returning overloads (mesh.h:1514-1540, 1583-1608, 2216-2257).  stats may be NULL. */
void orc_query_point_no_sign(const float* points, const int* indices, const orc_half* node_lowers,
                             const orc_half* node_uppers, const int* primitive_indices, int root,
                             const float* queries, int64_t n, float max_dist, uint8_t* result, int* face, float* u,
                             float* v, uint64_t* stats)
{
    const orc_mesh m = make_mesh(points, indices, node_lowers, node_uppers, primitive_indices, root);
    orc_stats st = { 0, 0 };
    for (int64_t i = 0; i < n; ++i) {
        int f = 0;
        float bu = 0.f, bv = 0.f;
        const int ok = point_no_sign_one(&m, v3_ld(queries, i), max_dist, &f, &bu, &bv, stats ? &st : NULL);
        result[i] = (uint8_t)ok;
        face[i] = ok ? f : 0;
        u[i] = ok ? bu : 0.f;
        v[i] = ok ? bv : 0.f;
    }
    if (stats)
        stats[0] = st.nodes_visited, stats[1] = st.tris_tested;
}

/* closest point + inside/outside sign -- mesh.h:128-307 (sign only evaluated when a point was found) */
void orc_query_point(const float* points, const int* indices, const orc_half* node_lowers,
                     const orc_half* node_uppers, const int* primitive_indices, int root, const float* queries,
                     int64_t n, float max_dist, uint8_t* result, float* sign, int* face, float* u, float* v,
                     uint64_t* stats)
{
    const orc_mesh m = make_mesh(points, indices, node_lowers, node_uppers, primitive_indices, root);
    orc_stats st = { 0, 0 };
    for (int64_t i = 0; i < n; ++i) {
        int f = 0;
        float bu = 0.f, bv = 0.f;
        const v3 q = v3_ld(queries, i);
        const int ok = point_no_sign_one(&m, q, max_dist, &f, &bu, &bv, stats ? &st : NULL);
        result[i] = (uint8_t)ok;
        face[i] = ok ? f : 0;
        u[i] = ok ? bu : 0.f;
        v[i] = ok ? bv : 0.f;
        sign[i] = ok ? inside_by_axis_rays(&m, q, stats ? &st : NULL) : 0.f;
    }
    if (stats)
        stats[0] = st.nodes_visited, stats[1] = st.tris_tested;
}

void orc_query_ray(const float* points, const int* indices, const orc_half* node_lowers, const orc_half* node_uppers,
                   const int* primitive_indices, int root, const float* starts, const float* dirs, int64_t n,
                   float max_t, uint8_t* result, float* sign, int* face, float* t, float* u, float* v, float* normal,
                   uint64_t* stats, const int* roots)
{
    orc_mesh m = make_mesh(points, indices, node_lowers, node_uppers, primitive_indices, root);
    orc_stats st = { 0, 0 };
    for (int64_t i = 0; i < n; ++i) {
        float tt = 0.f, tu = 0.f, tv = 0.f, ts = 0.f;
        int f = 0;
        v3 nrm = v3_make(0, 0, 0);
        m.root = (roots && roots[i] != -1) ? roots[i] : root; /* mesh_query_ray(..., root), mesh.h:1778 */
        const int ok = ray_one(&m, v3_ld(starts, i), v3_ld(dirs, i), max_t, &tt, &tu, &tv, &ts, &nrm, &f,
                               stats ? &st : NULL);
        result[i] = (uint8_t)ok;
        sign[i] = ok ? ts : 0.f;
        face[i] = ok ? f : 0;
        t[i] = ok ? tt : 0.f;
        u[i] = ok ? tu : 0.f;
        v[i] = ok ? tv : 0.f;
        normal[3 * i + 0] = ok ? nrm.x : 0.f;
        normal[3 * i + 1] = ok ? nrm.y : 0.f;
        normal[3 * i + 2] = ok ? nrm.z : 0.f;
    }
    if (stats)
        stats[0] = st.nodes_visited, stats[1] = st.tris_tested;
}

/* any hit with 0 <= t < max_t, near child first, fixed bound -- mesh.h:1893-1974 */
static int ray_anyhit_one(const orc_mesh* m, v3 start, v3 dir, float max_t)
{
    ray_entry stack[ORC_STACK];
    int size = 0;
    ray_entry cur = { HALF_B(m->node_lowers[m->root]), HALF_I(m->node_lowers[m->root]), HALF_I(m->node_uppers[m->root]) };
    v3 safe = dir;
    if (safe.x == 0.0f)
        safe.x = 1.0e-20f;
    if (safe.y == 0.0f)
        safe.y = 1.0e-20f;
    if (safe.z == 0.0f)
        safe.z = 1.0e-20f;
    const v3 rcp = v3_make(1.0f / safe.x, 1.0f / safe.y, 1.0f / safe.z);
    const int fast = dir.x != 0.0f && dir.y != 0.0f && dir.z != 0.0f;
    for (;;) {
        if (cur.leaf) {
            for (int pc = cur.lo_payload; pc < cur.hi_payload; ++pc) {
                const int prim = m->primitive_indices[pc];
                const v3 p = v3_ld(m->points, m->indices[3 * prim + 0]);
                const v3 q = v3_ld(m->points, m->indices[3 * prim + 1]);
                const v3 r = v3_ld(m->points, m->indices[3 * prim + 2]);
                float tt, tu, tv, ts;
                v3 n;
                if (ray_tri_watertight(start, dir, p, q, r, &tt, &tu, &tv, &ts, &n) && tt < max_t && tt >= 0.0f)
                    return 1;
            }
            if (size == 0)
                return 0;
            cur = stack[--size];
            continue;
        }
        const int li = cur.lo_payload, ri = cur.hi_payload;
        const orc_half llo = m->node_lowers[li], lhi = m->node_uppers[li];
        const orc_half rlo = m->node_lowers[ri], rhi = m->node_uppers[ri];
        float t0 = FLT_MAX, t1 = FLT_MAX;
        const int h0 = (fast ? ray_aabb_fast(start, rcp, &llo, &lhi, &t0)
                             : ray_aabb_robust(start, dir, rcp, &llo, &lhi, &t0))
            && t0 < max_t;
        const int h1 = (fast ? ray_aabb_fast(start, rcp, &rlo, &rhi, &t1)
                             : ray_aabb_robust(start, dir, rcp, &rlo, &rhi, &t1))
            && t1 < max_t;
        ray_entry le = { HALF_B(llo), HALF_I(llo), HALF_I(lhi) };
        ray_entry re = { HALF_B(rlo), HALF_I(rlo), HALF_I(rhi) };
        if (h0 && h1) {
            if (size >= ORC_STACK)
                return 0; /* mesh.h:1958-1959 */
            const int near_left = t0 < t1;
            stack[size++] = near_left ? re : le;
            cur = near_left ? le : re;
        } else if (h0) {
            cur = le;
        } else if (h1) {
            cur = re;
        } else {
            if (size == 0)
                return 0;
            cur = stack[--size];
        }
    }
}

/* number of triangles hit with t >= 0 along the whole ray, push-both traversal -- mesh.h:1976-2032.
 * The reference pushes without an overflow check; trees deeper than the stack are outside its defined
 * behaviour and this restatement stops pushing there. */
static int ray_count_one(const orc_mesh* m, v3 start, v3 dir)
{
    int stack[ORC_STACK + 2];
    int count = 1, hits = 0;
    stack[0] = m->root;
    const v3 rcp = v3_make(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    while (count) {
        const int node = stack[--count];
        const orc_half lo = m->node_lowers[node], hi = m->node_uppers[node];
        float tt;
        if (!ray_aabb_robust(start, dir, rcp, &lo, &hi, &tt))
            continue;
        if (HALF_B(lo)) {
            for (int pc = HALF_I(lo); pc < HALF_I(hi); ++pc) {
                const int prim = m->primitive_indices[pc];
                const v3 p = v3_ld(m->points, m->indices[3 * prim + 0]);
                const v3 q = v3_ld(m->points, m->indices[3 * prim + 1]);
                const v3 r = v3_ld(m->points, m->indices[3 * prim + 2]);
                float t, u, v, s;
                v3 n;
                if (ray_tri_watertight(start, dir, p, q, r, &t, &u, &v, &s, &n) && t >= 0.0f)
                    hits++;
            }
        } else if (count < ORC_STACK) {
            stack[count++] = HALF_I(lo);
            stack[count++] = HALF_I(hi);
        }
    }
    return hits;
}

void orc_query_ray_anyhit(const float* points, const int* indices, const orc_half* node_lowers,
                          const orc_half* node_uppers, const int* primitive_indices, int root, const float* starts,
                          const float* dirs, int64_t n, float max_t, uint8_t* result, const int* roots)
{
    orc_mesh m = make_mesh(points, indices, node_lowers, node_uppers, primitive_indices, root);
    for (int64_t i = 0; i < n; ++i) {
        m.root = (roots && roots[i] != -1) ? roots[i] : root;
        result[i] = (uint8_t)ray_anyhit_one(&m, v3_ld(starts, i), v3_ld(dirs, i), max_t);
    }
}

void orc_query_ray_count(const float* points, const int* indices, const orc_half* node_lowers,
                         const orc_half* node_uppers, const int* primitive_indices, int root, const float* starts,
                         const float* dirs, int64_t n, int* counts, const int* roots)
{
    orc_mesh m = make_mesh(points, indices, node_lowers, node_uppers, primitive_indices, root);
    for (int64_t i = 0; i < n; ++i) {
        m.root = (roots && roots[i] != -1) ? roots[i] : root;
        counts[i] = ray_count_one(&m, v3_ld(starts, i), v3_ld(dirs, i));
    }
}

/* p*u + q*v + r*(1 - u - v) -- mesh.h:2767-2805 (vec3 scale then add, left to right) */
void orc_mesh_eval(const float* attr, const int* indices, const int* face, const float* u, const float* v, int64_t n,
                   float* out)
{
    for (int64_t i = 0; i < n; ++i) {
        const int f = face[i];
        const v3 p = v3_ld(attr, indices[3 * f + 0]), q = v3_ld(attr, indices[3 * f + 1]), r = v3_ld(attr, indices[3 * f + 2]);
        const v3 x = v3_add(v3_add(v3_scale(u[i], p), v3_scale(v[i], q)), v3_scale(1.0f - u[i] - v[i], r));
        out[3 * i] = x.x, out[3 * i + 1] = x.y, out[3 * i + 2] = x.z;
    }
}

/* mesh_query_point_sign_parity (mesh.h:309-498) = the no-sign closest point + mesh_query_inside_parity
 * (mesh.h:2362-2392): n_sample rays along (1,1,1) + (randf, randf, randf), PCG stream seeded with 42 (rand.h:29-80),
 * inside when vote * 2 >= n_sample.  The three randf() calls are arguments of ONE constructor call in the reference,
 * so their order is the compiler's: the reference's device builds (nvcc / NVRTC, clang for the CPU JIT) draw x, y, z
 * (left to right; checked on nvcc PTX), g++ -- which builds oracle/_ref -- draws z, y, x.  rtl != 0 selects the
 * latter so the restatement can be pinned on _ref; the CUDA path is compared with rtl == 0. */
static uint32_t rand_pcg(uint32_t state)
{
    const uint32_t b = state * 747796405u + 2891336453u;
    const uint32_t c = ((b >> ((b >> 28u) + 4u)) ^ b) * 277803737u;
    return (c >> 22u) ^ c;
}

static float randf_range(uint32_t* state, float lo, float hi)
{
    *state = rand_pcg(*state);
    return (hi - lo) * ((float)(*state >> 8) * (1.0f / 16777216.0f)) + lo;
}

void orc_query_point_sign_parity(const float* points, const int* indices, const orc_half* node_lowers,
                                 const orc_half* node_uppers, const int* primitive_indices, int root, const float* queries,
                                 int64_t n, float max_dist, int n_sample, float scale, int rtl, uint8_t* result,
                                 float* sign, int* face, float* u, float* v)
{
    const orc_mesh m = make_mesh(points, indices, node_lowers, node_uppers, primitive_indices, root);
    for (int64_t i = 0; i < n; ++i) {
        int f = 0;
        float bu = 0.f, bv = 0.f;
        const v3 p = v3_ld(queries, i);
        const int ok = point_no_sign_one(&m, p, max_dist, &f, &bu, &bv, NULL);
        result[i] = (uint8_t)ok;
        face[i] = ok ? f : 0, u[i] = ok ? bu : 0.f, v[i] = ok ? bv : 0.f, sign[i] = 0.f;
        if (!ok)
            continue;
        uint32_t state = rand_pcg(42u);
        int vote = 0;
        for (int k = 0; k < n_sample; ++k) {
            v3 dir;
            do {
                float r0 = randf_range(&state, -scale, scale);
                float r1 = randf_range(&state, -scale, scale);
                float r2 = randf_range(&state, -scale, scale);
                dir = rtl ? v3_make(1.0f + r2, 1.0f + r1, 1.0f + r0) : v3_make(1.0f + r0, 1.0f + r1, 1.0f + r2);
            } while (v3_dot(dir, dir) < 1e-8f);
            if (ray_count_one(&m, p, dir) % 2)
                vote++;
        }
        sign[i] = (vote * 2 >= n_sample) ? -1.0f : 1.0f;
    }
}

/* mesh_query_point_sign_normal (mesh.h:860-1090): closest point on DISTANCES with an epsilon band
 * eps = average_edge_length * epsilon; every triangle within eps of the running minimum adds
 * weight * normalize(normal) (vertex: corner angle via acosf, edge: pi, interior: 2 pi); sign = +1 when the
 * accumulated normal points towards the query.  average_edge_length is an INPUT here: the reference computes it
 * with a float loop on the CPU (mesh.cpp:140-155, orc_average_edge_length mode 0) and with a CUB scan on the GPU
 * (mesh.cu:38-60, 299-307); the CUDA path of this repo reduces the same float terms in double (mode 1). */
float orc_average_edge_length(const float* points, const int* indices, int num_tris, int mode)
{
    float fsum = 0.0f;
    double dsum = 0.0;
    for (int t = 0; t < num_tris; ++t) {
        const v3 p = v3_ld(points, indices[3 * t + 0]), q = v3_ld(points, indices[3 * t + 1]),
                 r = v3_ld(points, indices[3 * t + 2]);
        if (mode == 0) { /* mesh.cpp:153: length(p0 - p1) + length(p0 - p2) + length(p2 - p1) */
            const v3 a = v3_sub(p, q), b = v3_sub(p, r), c = v3_sub(r, q);
            fsum += sqrtf(v3_dot(a, a)) + sqrtf(v3_dot(b, b)) + sqrtf(v3_dot(c, c));
        } else { /* mesh.cu:53: length(p - q) + length(p - r) + length(q - r) */
            const v3 a = v3_sub(p, q), b = v3_sub(p, r), c = v3_sub(q, r);
            dsum += (double)(sqrtf(v3_dot(a, a)) + sqrtf(v3_dot(b, b)) + sqrtf(v3_dot(c, c)));
        }
    }
    if (mode == 0)
        return fsum / (float)(num_tris * 3);
    return (float)dsum / (float)(3 * num_tris);
}

static inline v3 v3_normalize(v3 a) /* vec.h:1111-1118, kEps = 0 */
{
    const float l = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    if (l > 0.0f)
        return v3_make(a.x / l, a.y / l, a.z / l);
    return v3_make(0.f, 0.f, 0.f);
}

static int point_sign_normal_one(const orc_mesh* m, v3 point, float max_dist, float eps, float* inside, int* face,
                                 float* u, float* v)
{
    int stack[ORC_STACK];
    int count = 1;
    stack[0] = m->root;
    float min_dist = max_dist;
    int min_face = 0;
    float min_v = 0.f, min_w = 0.f;
    v3 acc = v3_make(0.f, 0.f, 0.f);
    const float eps_sq = eps * eps;

    while (count) {
        const int node = stack[--count];
        const orc_half lo = m->node_lowers[node], hi = m->node_uppers[node];
        if (dist_aabb_sq(point, &lo, &hi) > (min_dist + eps) * (min_dist + eps))
            continue;
        const int li = HALF_I(lo), ri = HALF_I(hi);
        if (HALF_B(lo)) {
            for (int pc = li; pc < ri; ++pc) {
                const int prim = m->primitive_indices[pc];
                const v3 p = v3_ld(m->points, m->indices[3 * prim + 0]);
                const v3 q = v3_ld(m->points, m->indices[3 * prim + 1]);
                const v3 r = v3_ld(m->points, m->indices[3 * prim + 2]);
                const v3 e0 = v3_sub(q, p), e1 = v3_sub(r, p), e2 = v3_sub(r, q);
                const v3 nrm = v3_cross(e0, e1);
                const float e0n = v3_dot(e0, e0), e1n = v3_dot(e1, e1), e2n = v3_dot(e2, e2);
                if (sqrtf(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z) / (e0n + e1n + e2n) < 1.e-6f)
                    continue;
                float uv[2];
                orc_closest_point_to_triangle(&p.x, &q.x, &r.x, &point.x, uv);
                const float bu = uv[0], bv = uv[1], bw = 1.f - bu - bv;
                const v3 c = v3_add(v3_add(v3_scale(bu, p), v3_scale(bv, q)), v3_scale(bw, r));
                const v3 d = v3_sub(c, point);
                const float dist = sqrtf(v3_dot(d, d));
                if (dist < min_dist + eps) {
                    float weight;
                    const v3 cp = v3_sub(c, p), cq = v3_sub(c, q), cr = v3_sub(c, r);
                    const float lcp = v3_dot(cp, cp), lcq = v3_dot(cq, cq), lcr = v3_dot(cr, cr);
                    const v3 ne0 = v3_make(-e0.x, -e0.y, -e0.z), ne1 = v3_make(-e1.x, -e1.y, -e1.z),
                             ne2 = v3_make(-e2.x, -e2.y, -e2.z);
                    if (lcp < eps_sq) {
                        weight = acosf(v3_dot(v3_normalize(e0), v3_normalize(e1)));
                    } else if (lcq < eps_sq) {
                        weight = acosf(v3_dot(v3_normalize(e2), v3_normalize(ne0)));
                    } else if (lcr < eps_sq) {
                        weight = acosf(v3_dot(v3_normalize(ne1), v3_normalize(ne2)));
                    } else {
                        const float e0cp = v3_dot(e0, cp), e2cq = v3_dot(e2, cq), e1cp = v3_dot(e1, cp);
                        if ((lcp * e0n - e0cp * e0cp < eps_sq * e0n) || (lcq * e2n - e2cq * e2cq < eps_sq * e2n)
                            || (lcp * e1n - e1cp * e1cp < eps_sq * e1n))
                            weight = 3.14159265359f;
                        else
                            weight = 2.0f * 3.14159265359f;
                    }
                    const v3 wn = v3_scale(weight, v3_normalize(nrm));
                    if (dist > min_dist - eps) {
                        acc = v3_add(acc, wn);
                        if (dist < min_dist)
                            min_dist = dist, min_v = bv, min_w = bw, min_face = prim;
                    } else {
                        min_dist = dist, min_v = bv, min_w = bw, min_face = prim;
                        acc = wn;
                    }
                }
            }
        } else {
            const orc_half llo = m->node_lowers[li], lhi = m->node_uppers[li];
            const orc_half rlo = m->node_lowers[ri], rhi = m->node_uppers[ri];
            const float dl = dist_aabb_sq(point, &llo, &lhi), dr = dist_aabb_sq(point, &rlo, &rhi);
            int first, second;
            float dfirst, dsecond;
            if (dl < dr) {
                first = ri, second = li, dfirst = dr, dsecond = dl;
            } else {
                first = li, second = ri, dfirst = dl, dsecond = dr;
            }
            if (dfirst < (min_dist + eps) * (min_dist + eps))
                stack[count++] = first;
            if (dsecond < (min_dist + eps) * (min_dist + eps))
                stack[count++] = second;
        }
    }
    if (min_dist < max_dist) {
        *u = 1.0f - min_v - min_w;
        *v = min_v;
        *face = min_face;
        const v3 p = v3_ld(m->points, m->indices[3 * min_face + 0]);
        const v3 q = v3_ld(m->points, m->indices[3 * min_face + 1]);
        const v3 r = v3_ld(m->points, m->indices[3 * min_face + 2]);
        const v3 cpt = v3_add(v3_add(v3_scale(*u, p), v3_scale(*v, q)), v3_scale(min_w, r));
        *inside = v3_dot(acc, v3_sub(point, cpt)) > 0.0f ? 1.0f : -1.0f;
        return 1;
    }
    return 0;
}

void orc_query_point_sign_normal(const float* points, const int* indices, const orc_half* node_lowers,
                                 const orc_half* node_uppers, const int* primitive_indices, int root, const float* queries,
                                 int64_t n, float max_dist, float average_edge_length, float epsilon, uint8_t* result,
                                 float* sign, int* face, float* u, float* v)
{
    const orc_mesh m = make_mesh(points, indices, node_lowers, node_uppers, primitive_indices, root);
    const float eps = average_edge_length * epsilon;
    for (int64_t i = 0; i < n; ++i) {
        int f = 0;
        float bu = 0.f, bv = 0.f, sg = 0.f;
        const int ok = point_sign_normal_one(&m, v3_ld(queries, i), max_dist, eps, &sg, &f, &bu, &bv);
        result[i] = (uint8_t)ok;
        face[i] = ok ? f : 0, u[i] = ok ? bu : 0.f, v[i] = ok ? bv : 0.f, sign[i] = ok ? sg : 0.f;
    }
}

/* mesh_query_furthest_point_no_sign (mesh.h:678-858) with furthest_distance_to_aabb_sq (mesh.h:100-119) and
 * furthest_point_to_triangle (intersect.h:111-125): farther child popped first, nodes culled when their farthest
 * corner is nearer than the best so far, candidates are triangle vertices, strict '>' updates. */
static inline float far_aabb_sq(v3 p, const orc_half* lo, const orc_half* hi)
{
    const float lx = fabsf(p.x - lo->x), ux = fabsf(p.x - hi->x), cx = (lx > ux) ? lx : ux;
    const float ly = fabsf(p.y - lo->y), uy = fabsf(p.y - hi->y), cy = (ly > uy) ? ly : uy;
    const float lz = fabsf(p.z - lo->z), uz = fabsf(p.z - hi->z), cz = (lz > uz) ? lz : uz;
    return cx * cx + cy * cy + cz * cz;
}

static int point_furthest_one(const orc_mesh* m, v3 point, float min_dist, int* face, float* u, float* v)
{
    int stack[ORC_STACK];
    int count = 1;
    stack[0] = m->root;
    float best = min_dist * min_dist;
    int best_face = 0;
    float best_v = 0.f, best_w = 0.f;
    while (count) {
        const int node = stack[--count];
        const orc_half lo = m->node_lowers[node], hi = m->node_uppers[node];
        if (far_aabb_sq(point, &lo, &hi) < best)
            continue;
        const int li = HALF_I(lo), ri = HALF_I(hi);
        if (HALF_B(lo)) {
            for (int pc = li; pc < ri; ++pc) {
                const int prim = m->primitive_indices[pc];
                const v3 p = v3_ld(m->points, m->indices[3 * prim + 0]);
                const v3 q = v3_ld(m->points, m->indices[3 * prim + 1]);
                const v3 r = v3_ld(m->points, m->indices[3 * prim + 2]);
                const v3 e0 = v3_sub(q, p), e1 = v3_sub(r, p), e2 = v3_sub(r, q);
                const v3 nrm = v3_cross(e0, e1);
                if (sqrtf(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z)
                        / (v3_dot(e0, e0) + v3_dot(e1, e1) + v3_dot(e2, e2))
                    < 1.e-6f)
                    continue;
                const v3 pa = v3_sub(point, p), pb = v3_sub(point, q), pcv = v3_sub(point, r);
                const float da = v3_dot(pa, pa), db = v3_dot(pb, pb), dc = v3_dot(pcv, pcv);
                float bu, bv;
                if (da > db && da > dc)
                    bu = 1.0f, bv = 0.0f;
                else if (db > dc)
                    bu = 0.0f, bv = 1.0f;
                else
                    bu = 0.0f, bv = 0.0f;
                const float bw = 1.f - bu - bv;
                const v3 c = v3_add(v3_add(v3_scale(bu, p), v3_scale(bv, q)), v3_scale(bw, r));
                const v3 d = v3_sub(c, point);
                const float dsq = v3_dot(d, d);
                if (dsq > best)
                    best = dsq, best_v = bv, best_w = bw, best_face = prim;
            }
        } else {
            const orc_half llo = m->node_lowers[li], lhi = m->node_uppers[li];
            const orc_half rlo = m->node_lowers[ri], rhi = m->node_uppers[ri];
            const float dl = far_aabb_sq(point, &llo, &lhi), dr = far_aabb_sq(point, &rlo, &rhi);
            int first, second;
            float dfirst, dsecond;
            if (dl > dr) {
                first = ri, second = li, dfirst = dr, dsecond = dl;
            } else {
                first = li, second = ri, dfirst = dl, dsecond = dr;
            }
            if (dfirst > best)
                stack[count++] = first;
            if (dsecond > best)
                stack[count++] = second;
        }
    }
    if (best > min_dist * min_dist) {
        *u = 1.0f - best_v - best_w;
        *v = best_v;
        *face = best_face;
        return 1;
    }
    return 0;
}

void orc_query_furthest_point_no_sign(const float* points, const int* indices, const orc_half* node_lowers,
                                      const orc_half* node_uppers, const int* primitive_indices, int root,
                                      const float* queries, int64_t n, float min_dist, uint8_t* result, int* face, float* u,
                                      float* v)
{
    const orc_mesh m = make_mesh(points, indices, node_lowers, node_uppers, primitive_indices, root);
    for (int64_t i = 0; i < n; ++i) {
        int f = 0;
        float bu = 0.f, bv = 0.f;
        const int ok = point_furthest_one(&m, v3_ld(queries, i), min_dist, &f, &bu, &bv);
        result[i] = (uint8_t)ok;
        face[i] = ok ? f : 0, u[i] = ok ? bu : 0.f, v[i] = ok ? bv : 0.f;
    }
}

/* mesh_eval_face_normal (mesh.h:2870-2888) */
void orc_mesh_face_normal(const float* points, const int* indices, const int* face, int64_t n, float* out)
{
    for (int64_t i = 0; i < n; ++i) {
        const int t = face[i];
        const v3 p = v3_ld(points, indices[3 * t + 0]), q = v3_ld(points, indices[3 * t + 1]),
                 r = v3_ld(points, indices[3 * t + 2]);
        const v3 nn = v3_normalize(v3_cross(v3_sub(q, p), v3_sub(r, p)));
        out[3 * i] = nn.x, out[3 * i + 1] = nn.y, out[3 * i + 2] = nn.z;
    }
}

/* mesh_query_sphere + mesh_query_sphere_next run to exhaustion (mesh.h:2457-2737): node test = exact sphere / box
 * (intersect.h:197-205); face test (also on single-face leaves) = the same test on the face's box (mesh.lowers /
 * uppers of the last build / refit, passed in), then the closest point of the triangle -- of its longest edge for a
 * zero-area face -- within the radius.  offsets[n+1]; indices may be NULL (count only). */
static int sphere_aabb(v3 c, float radius_sq, const float* lo, const float* hi)
{
    const float dx = fmax_ref(fmax_ref(lo[0] - c.x, c.x - hi[0]), 0.0f);
    const float dy = fmax_ref(fmax_ref(lo[1] - c.y, c.y - hi[1]), 0.0f);
    const float dz = fmax_ref(fmax_ref(lo[2] - c.z, c.z - hi[2]), 0.0f);
    return dx * dx + dy * dy + dz * dz <= radius_sq;
}

static int sphere_face(const orc_mesh* m, const float* tri_lowers, const float* tri_uppers, v3 center, float radius_sq,
                       int prim)
{
    if (!sphere_aabb(center, radius_sq, tri_lowers + 3 * prim, tri_uppers + 3 * prim))
        return 0;
    const v3 a = v3_ld(m->points, m->indices[3 * prim + 0]), b = v3_ld(m->points, m->indices[3 * prim + 1]),
             c = v3_ld(m->points, m->indices[3 * prim + 2]);
    const v3 ab = v3_sub(b, a), ac = v3_sub(c, a);
    const v3 n = v3_cross(ab, ac);
    v3 cp;
    if (v3_dot(n, n) == 0.0f) {
        const v3 bc = v3_sub(c, b);
        const float lab2 = v3_dot(ab, ab), lac2 = v3_dot(ac, ac), lbc2 = v3_dot(bc, bc);
        v3 p, q;
        float len2;
        if (lab2 >= lac2 && lab2 >= lbc2)
            p = a, q = b, len2 = lab2;
        else if (lac2 >= lbc2)
            p = a, q = c, len2 = lac2;
        else
            p = b, q = c, len2 = lbc2;
        const v3 pq = v3_sub(q, p);
        const float t = (len2 > 0.0f) ? fmin_ref(fmax_ref(0.0f, v3_dot(v3_sub(center, p), pq) / len2), 1.0f) : 0.0f;
        cp = v3_add(p, v3_scale(t, pq));
    } else {
        float uv[2];
        orc_closest_point_to_triangle(&a.x, &b.x, &c.x, &center.x, uv);
        cp = v3_add(v3_add(v3_scale(uv[0], a), v3_scale(uv[1], b)), v3_scale(1.0f - uv[0] - uv[1], c));
    }
    const v3 d = v3_sub(cp, center);
    return v3_dot(d, d) <= radius_sq;
}

void orc_mesh_query_sphere(const float* points, const int* indices, const orc_half* node_lowers, const orc_half* node_uppers,
                           const int* primitive_indices, int root, const float* tri_lowers, const float* tri_uppers,
                           const float* centers, const float* radii, int64_t n, int* offsets, int* out)
{
    const orc_mesh m = make_mesh(points, indices, node_lowers, node_uppers, primitive_indices, root);
    int run = 0;
    for (int64_t i = 0; i < n; ++i) {
        offsets[i] = run;
        const v3 c = v3_ld(centers, i);
        const float r = fmax_ref(radii[i], 0.0f);
        const float r2 = r * r;
        int stack[ORC_STACK * 2];
        int count = 1;
        stack[0] = root;
        while (count) {
            const int node = stack[--count];
            const orc_half lo = node_lowers[node], hi = node_uppers[node];
            if (!sphere_aabb(c, r2, &lo.x, &hi.x))
                continue;
            if (HALF_B(lo)) {
                for (int k = HALF_I(lo); k < HALF_I(hi); ++k) {
                    const int prim = primitive_indices[k];
                    if (sphere_face(&m, tri_lowers, tri_uppers, c, r2, prim)) {
                        if (out)
                            out[run] = prim;
                        ++run;
                    }
                }
            } else {
                stack[count++] = HALF_I(lo);
                stack[count++] = HALF_I(hi);
            }
        }
    }
    offsets[n] = run;
}
