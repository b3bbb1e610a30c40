// Batched mesh queries over the sibling-pair tree: closest point (with / without inside-outside
// sign) and closest ray hit.  One thread per query; every float operation is written in the order
// the reference evaluates it (warp/native/mesh.h, intersect.h) and the library is compiled with
// -fmad=false, so results are bit-identical to the reference's host build on the same tree --
// including which face wins an exact distance tie, because nodes are visited in the same order.
//
// What differs from the reference kernels (mesh.h:128-307, 501-676, 1768-1891, 2286-2359):
//   * one aligned 64-byte fetch per inner node (both child boxes) instead of 2 + 4 separate 16-byte
//     loads from two arrays; a node's own box is never re-fetched (its distance rides on the stack);
//   * leaf triangles come from the packed, Morton-ordered triangle cache (48 contiguous bytes per
//     triangle incl. face id and sliver flag) instead of the prim -> indices -> points chain;
//   * the nearer child is entered directly instead of being pushed and popped.
#include "state.h"
#include "query.h"

#include <cstdlib>

namespace {

#ifndef WB_QT
#define WB_QT 128
#endif
constexpr int QT = WB_QT;  // threads per block

struct Entry {
    uint32_t a;  // leaf: first sorted position | WB_LEAF ; inner: internal slot s
    uint32_t b;  // leaf: primitive count
};

struct Tri {
    float3 p, q, r;
    int face;
    uint32_t flags;
};

__device__ __forceinline__ Tri load_tri(const float4* __restrict__ tris, uint32_t pos)
{
    const float4 t0 = __ldg(tris + 3 * (size_t)pos), t1 = __ldg(tris + 3 * (size_t)pos + 1),
                 t2 = __ldg(tris + 3 * (size_t)pos + 2);
    Tri t;
    t.p = make_float3(t0.x, t0.y, t0.z);
    t.q = make_float3(t0.w, t1.x, t1.y);
    t.r = make_float3(t1.z, t1.w, t2.x);
    t.face = __float_as_int(t2.y);
    t.flags = __float_as_uint(t2.z);
    return t;
}

struct Pair {
    float3 llo, lhi, rlo, rhi;
    Entry left, right;
};

// fetch both children of internal slot s and decode their stack entries
__device__ __forceinline__ Pair load_pair(const NodeRec* __restrict__ pairs, uint32_t s, int n)
{
    const float4* p4 = reinterpret_cast<const float4*>(pairs + 2 * (size_t)s);
    const float4 a0 = __ldg(p4), a1 = __ldg(p4 + 1), b0 = __ldg(p4 + 2), b1 = __ldg(p4 + 3);
    Pair pr;
    pr.llo = make_float3(a0.x, a0.y, a0.z), pr.lhi = make_float3(a1.x, a1.y, a1.z);
    pr.rlo = make_float3(b0.x, b0.y, b0.z), pr.rhi = make_float3(b1.x, b1.y, b1.z);
    const uint32_t lref = __float_as_uint(a0.w), laux = __float_as_uint(a1.w);
    const uint32_t rref = __float_as_uint(b0.w), raux = __float_as_uint(b1.w);
    if (lref & WB_LEAF)
        pr.left.a = laux | WB_LEAF, pr.left.b = s - laux + 1;  // range [laux, s]
    else
        pr.left.a = (lref & WB_IDX_MASK) - (uint32_t)n, pr.left.b = 0;
    if (rref & WB_LEAF)
        pr.right.a = (s + 1) | WB_LEAF, pr.right.b = raux - s;  // range [s+1, raux]
    else
        pr.right.a = (rref & WB_IDX_MASK) - (uint32_t)n, pr.right.b = 0;
    return pr;
}

// squared distance point -> AABB (mesh.h:92-98)
__device__ __forceinline__ float dist_aabb_sq(float3 p, float3 lo, float3 hi)
{
    const float dx = fminf(hi.x, fmaxf(lo.x, p.x)) - p.x;
    const float dy = fminf(hi.y, fmaxf(lo.y, p.y)) - p.y;
    const float dz = fminf(hi.z, fmaxf(lo.z, p.z)) - p.z;
    return dx * dx + dy * dy + dz * dz;
}

// Voronoi-region closest point on a triangle (intersect.h:44-109); returns (v, w), u = 1 - v - w
__device__ __forceinline__ void closest_vw(float3 a, float3 b, float3 c, float3 p, float& v, float& w)
{
    const float3 ab = wb_sub(b, a), ac = wb_sub(c, a), ap = wb_sub(p, a);
    const float d1 = wb_dot(ab, ap), d2 = wb_dot(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) {
        v = 0.0f, w = 0.0f;
        return;
    }
    const float3 bp = wb_sub(p, b);
    const float d3 = wb_dot(ab, bp), d4 = wb_dot(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) {
        v = 1.0f, w = 0.0f;
        return;
    }
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        v = d1 / (d1 - d3), w = 0.0f;
        return;
    }
    const float3 cp = wb_sub(p, c);
    const float d5 = wb_dot(ab, cp), d6 = wb_dot(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) {
        v = 0.0f, w = 1.0f;
        return;
    }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        v = 0.0f, w = d2 / (d2 - d6);
        return;
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        v = 1.0f - w;
        return;
    }
    const float denom = 1.0f / (va + vb + vc);
    v = vb * denom;
    w = vc * denom;
}

// slab tests (intersect.h:127-181)
__device__ __forceinline__ bool ray_aabb_fast(float3 pos, float3 rcp, float3 lo, float3 hi, float& t)
{
    float l1 = (lo.x - pos.x) * rcp.x, l2 = (hi.x - pos.x) * rcp.x;
    float lmin = fminf(l1, l2), lmax = fmaxf(l1, l2);
    l1 = (lo.y - pos.y) * rcp.y, l2 = (hi.y - pos.y) * rcp.y;
    lmin = fmaxf(fminf(l1, l2), lmin), lmax = fminf(fmaxf(l1, l2), lmax);
    l1 = (lo.z - pos.z) * rcp.z, l2 = (hi.z - pos.z) * rcp.z;
    lmin = fmaxf(fminf(l1, l2), lmin), lmax = fminf(fmaxf(l1, l2), lmax);
    const bool hit = (lmax >= 0.f) & (lmax >= lmin);
    if (hit)
        t = lmin;
    return hit;
}

__device__ __forceinline__ bool ray_aabb_robust(float3 pos, float3 dir, float3 rcp, float3 lo, float3 hi, float& t)
{
    float lmin = -FLT_MAX, lmax = FLT_MAX;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float d = wb_get(dir, k), o = wb_get(pos, k), l = wb_get(lo, k), h = wb_get(hi, k);
        if (d == 0.0f) {
            if (o < l || o > h)
                return false;
        } else {
            const float r = wb_get(rcp, k);
            const float l1 = (l - o) * r, l2 = (h - o) * r;
            lmin = fmaxf(fminf(l1, l2), lmin);
            lmax = fminf(fmaxf(l1, l2), lmax);
        }
    }
    const bool hit = (lmax >= 0.f) & (lmax >= lmin);
    if (hit)
        t = lmin;
    return hit;
}

// per-ray constants of the watertight test (intersect.h:359-375)
struct WoopRay {
    int kx, ky, kz;
    float Sx, Sy, Sz;
};

__device__ __forceinline__ WoopRay woop_setup(float3 dir)
{
    WoopRay w;
    int kz = 0;
    float best = fabsf(dir.x);
    if (fabsf(dir.y) > best)
        kz = 1, best = fabsf(dir.y);
    if (fabsf(dir.z) > best)
        kz = 2;
    int kx = kz + 1 == 3 ? 0 : kz + 1;
    int ky = kx + 1 == 3 ? 0 : kx + 1;
    if (wb_get(dir, kz) < 0.0f) {
        const int tmp = kx;
        kx = ky, ky = tmp;
    }
    w.kx = kx, w.ky = ky, w.kz = kz;
    w.Sx = wb_get(dir, kx) / wb_get(dir, kz);
    w.Sy = wb_get(dir, ky) / wb_get(dir, kz);
    w.Sz = 1.0f / wb_get(dir, kz);
    return w;
}

// a*b - c*d with the error of c*d recovered by two explicit FMAs (intersect.h:334-341)
__device__ __forceinline__ float diff_of_products(float a, float b, float c, float d)
{
    const float cd = __fmul_rn(c, d);
    const float diff = __fmaf_rn(a, b, -cd);
    const float err = __fmaf_rn(-c, d, cd);
    return __fadd_rn(diff, err);
}

// watertight ray/triangle (intersect.h:377-444); t,u,v,sign written on a hit
__device__ __forceinline__ bool ray_tri(const WoopRay& w, float3 org, float3 a, float3 b, float3 c, float& t, float& u,
                                        float& v, float& sign)
{
    const float3 A = wb_sub(a, org), B = wb_sub(b, org), C = wb_sub(c, org);
    const float Akz = wb_get(A, w.kz), Bkz = wb_get(B, w.kz), Ckz = wb_get(C, w.kz);
    const float Ax = wb_get(A, w.kx) - w.Sx * Akz, Ay = wb_get(A, w.ky) - w.Sy * Akz;
    const float Bx = wb_get(B, w.kx) - w.Sx * Bkz, By = wb_get(B, w.ky) - w.Sy * Bkz;
    const float Cx = wb_get(C, w.kx) - w.Sx * Ckz, Cy = wb_get(C, w.ky) - w.Sy * Ckz;

    float U = diff_of_products(Cx, By, Cy, Bx);
    float V = diff_of_products(Ax, Cy, Ay, Cx);
    float W = diff_of_products(Bx, Ay, By, Ax);
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx));
        V = (float)__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx));
        W = (float)__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f))
        return false;
    const float det = U + V + W;
    if (det == 0.0f)
        return false;
    const float Az = w.Sz * Akz, Bz = w.Sz * Bkz, Cz = w.Sz * Ckz;
    const float T = U * Az + V * Bz + W * Cz;
    const uint32_t det_sign = __float_as_uint(det) & 0x80000000u;
    if (__uint_as_float(__float_as_uint(T) ^ det_sign) < 0.0f)
        return false;
    const float rcp_det = 1.0f / det;
    u = U * rcp_det;
    v = V * rcp_det;
    t = T * rcp_det;
    sign = det;
    return true;
}

#ifndef WB_LEAF_PRETEST
#define WB_LEAF_PRETEST 8u
#endif

struct Counters {
    unsigned long long pairs = 0;  // 64-byte sibling-pair fetches
    unsigned long long tris = 0;   // 48-byte packed-triangle fetches
};

// ------------------------------------------------------------------------------------------------
// closest point, mesh.h:501-676
// ------------------------------------------------------------------------------------------------
// STACK8 (the default for trees below 2^28 items): a stack entry is ONE 8-byte word {entry, distance}.  A leaf entry packs
// its first position and its item count into 31 bits -- start << 3 | count - 1 for counts up to 7; the code 7 stands for
// "8 or more" (depth-rule leaves) and sends the pop to the leaf's parent record for the count.  256 B of local memory per
// thread instead of 384 B, one STL.64 / LDL.64 per push / pop instead of three 4-byte accesses: -2.7 % on C2.
__device__ __forceinline__ uint32_t pack_entry(Entry e)
{
    if (!(e.a & WB_LEAF))
        return e.a;
    return WB_LEAF | ((e.a & WB_IDX_MASK) << 3) | (e.b - 1u < 7u ? e.b - 1u : 7u);
}

__device__ __forceinline__ Entry unpack_entry(const TreeView& tv, uint32_t w)
{
    Entry e;
    if (!(w & WB_LEAF)) {
        e.a = w, e.b = 0;
        return e;
    }
    const uint32_t start = (w & WB_IDX_MASK) >> 3, c = w & 7u;
    e.a = start | WB_LEAF;
    if (c < 7u) {
        e.b = c + 1u;
    } else {  // a large leaf: its range is in its record inside the parent's pair
        const int p = __ldg(tv.pos_parent + start);
        if (p == WB_ROOT_PARENT) {
            e.b = (uint32_t)tv.n;
        } else {
            const uint32_t ps = (uint32_t)(p - tv.n);
            e.b = start <= ps ? ps - start + 1u : tv.pairs[2 * (size_t)ps + 1].aux - ps;
        }
    }
    return e;
}

// STACK16: one 16-byte local-memory word {a, b, distance} per stack entry (one STL.128 / LDL.128 per push / pop) instead of
// three 4-byte words in two arrays -- the stack is ~40 % of the kernel's L1 transactions (126 pair fetches x 4 LDG.128 +
// 23 triangle fetches x 3 against ~60 pushes and pops x 3 per query)
template <bool COUNT, bool STACK16 = false, bool STACK8 = false>
__device__ __forceinline__ bool closest_point(const TreeView& tv, const TreeHeader& h, float3 point, float max_dist,
                                              int& face, float& u, float& v, Counters& cnt)
{
    Entry stack[(STACK16 || STACK8) ? 1 : WB_QUERY_STACK];
    float stack_d[(STACK16 || STACK8) ? 1 : WB_QUERY_STACK];
    uint4 stack16[STACK16 ? WB_QUERY_STACK : 1];
    uint2 stack8[STACK8 ? WB_QUERY_STACK : 1];  // {packed entry, distance}: see pack_entry
    int top = 0;

    float best = max_dist * max_dist;
    int best_face = 0;
    float best_v = 0.f, best_w = 0.f;

    Entry cur;
    if (h.root_ref & WB_LEAF)
        cur.a = WB_LEAF | 0u, cur.b = h.root_count;
    else
        cur.a = (h.root_ref & WB_IDX_MASK) - (uint32_t)tv.n, cur.b = 0;
    float cur_d = dist_aabb_sq(point, make_float3(h.lx, h.ly, h.lz), make_float3(h.hx, h.hy, h.hz));
    bool have = true;

    for (;;) {
        if (!have) {
            if (top == 0)
                break;
            --top;
            if (STACK8) {
                const uint2 e = stack8[top];
                cur = unpack_entry(tv, e.x), cur_d = __uint_as_float(e.y);
            } else if (STACK16) {
                const uint4 e = stack16[top];
                cur.a = e.x, cur.b = e.y, cur_d = __uint_as_float(e.z);
            } else {
                cur = stack[top];
                cur_d = stack_d[top];
            }
        }
        have = false;
        if (cur_d > best)
            continue;
        if (cur.a & WB_LEAF) {
            const uint32_t start = cur.a & WB_IDX_MASK;
            // leaves made by the depth rule (bvh.cu:419-441) hold dozens of triangles -- ~100 on a 100 M-triangle mesh,
            // where the 30-bit grid puts that many centroids in one cell.  There a triangle whose own box is already
            // farther than the best so far is skipped before the closest-point evaluation: the same cut the walk applies
            // to nodes (a triangle is never nearer than its box), at a quarter of the instructions.  Ordinary leaves
            // (<= leaf_size triangles) skip the pre-test: it costs more than it saves there (-1 % on C2).
            const bool big_leaf = cur.b > WB_LEAF_PRETEST;
            for (uint32_t pos = start; pos < start + cur.b; ++pos) {
                const Tri t = load_tri(tv.tris, pos);
                if (COUNT)
                    cnt.tris++;
                if (t.flags & WB_TRI_SLIVER)
                    continue;
                if (big_leaf
                    && dist_aabb_sq(point, wb_min3(wb_min3(t.p, t.q), t.r), wb_max3(wb_max3(t.p, t.q), t.r)) > best)
                    continue;
                float bv, bw;
                closest_vw(t.p, t.q, t.r, point, bv, bw);
                const float bu = 1.0f - bv - bw;       // what closest_point_to_triangle returns as u
                const float w = 1.f - bu - bv;         // mesh.h:569 recomputes w from (u, v)
                const float3 c = wb_add(wb_add(wb_scale(bu, t.p), wb_scale(bv, t.q)), wb_scale(w, t.r));
                const float3 d = wb_sub(c, point);
                const float dsq = wb_dot(d, d);
                if (dsq < best) {
                    best = dsq;
                    best_v = bv;
                    best_w = w;
                    best_face = t.face;
                }
            }
            continue;
        }
        const Pair pr = load_pair(tv.pairs, cur.a, tv.n);
        if (COUNT)
            cnt.pairs++;
        const float dl = dist_aabb_sq(point, pr.llo, pr.lhi), dr = dist_aabb_sq(point, pr.rlo, pr.rhi);
        Entry far_e, near_e;
        float far_d, near_d;
        if (dl < dr)
            far_e = pr.right, far_d = dr, near_e = pr.left, near_d = dl;
        else
            far_e = pr.left, far_d = dl, near_e = pr.right, near_d = dr;
        if (far_d < best) {
            if (STACK8) {
                stack8[top] = make_uint2(pack_entry(far_e), __float_as_uint(far_d));
            } else if (STACK16) {
                stack16[top] = make_uint4(far_e.a, far_e.b, __float_as_uint(far_d), 0u);
            } else {
                stack[top] = far_e;
                stack_d[top] = far_d;
            }
            ++top;
        }
        if (near_d < best) {
            cur = near_e;
            cur_d = near_d;
            have = true;
        }
    }
    if (best < max_dist * max_dist) {
        u = 1.0f - best_v - best_w;
        v = best_v;
        face = best_face;
        return true;
    }
    return false;
}

// robust slab test specialised for the probe direction e_axis (intersect.h:158-181 with dir = unit axis):
// the two other axes only ask "is the origin inside the slab", and on the probe axis rcp_dir = 1/1 = 1, so
// (bound - origin) * rcp_dir == bound - origin exactly.  Same value of `t`, same hit decision, fewer instructions.
__device__ __forceinline__ bool ray_aabb_axis(float oa, float o1, float o2, int axis, int a1, int a2, float3 lo, float3 hi,
                                              float& t)
{
    if (o1 < wb_get(lo, a1) || o1 > wb_get(hi, a1) || o2 < wb_get(lo, a2) || o2 > wb_get(hi, a2))
        return false;
    const float l1 = wb_get(lo, axis) - oa, l2 = wb_get(hi, axis) - oa;
    const float lmin = fmaxf(fminf(l1, l2), -FLT_MAX), lmax = fminf(fmaxf(l1, l2), FLT_MAX);
    const bool hit = (lmax >= 0.f) & (lmax >= lmin);
    if (hit)
        t = lmin;
    return hit;
}

// sign of the closest hit along an axis probe, push-both order (mesh.h:2286-2339)
template <bool COUNT>
__device__ __forceinline__ bool probe_sign(const TreeView& tv, const TreeHeader& h, float3 org, int axis, float& out_sign,
                                           Counters& cnt)
{
    const float3 dir = make_float3(axis == 0 ? 1.f : 0.f, axis == 1 ? 1.f : 0.f, axis == 2 ? 1.f : 0.f);
    const WoopRay wr = woop_setup(dir);
    const int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;  // the two axes the probe is parallel to
    const float oa = wb_get(org, axis), o1 = wb_get(org, a1), o2 = wb_get(org, a2);

    Entry stack[WB_QUERY_STACK];
    float stack_t[WB_QUERY_STACK];
    int top = 0;
    float min_t = FLT_MAX;
    bool hit = false;

    {
        float tt;
        if (ray_aabb_axis(oa, o1, o2, axis, a1, a2, make_float3(h.lx, h.ly, h.lz), make_float3(h.hx, h.hy, h.hz), tt)) {
            if (h.root_ref & WB_LEAF)
                stack[0].a = WB_LEAF | 0u, stack[0].b = h.root_count;
            else
                stack[0].a = (h.root_ref & WB_IDX_MASK) - (uint32_t)tv.n, stack[0].b = 0;
            stack_t[0] = tt;
            top = 1;
        }
    }
    while (top) {
        --top;
        const Entry cur = stack[top];
        if (!(stack_t[top] < min_t))
            continue;
        if (cur.a & WB_LEAF) {
            const uint32_t start = cur.a & WB_IDX_MASK;
            for (uint32_t pos = start; pos < start + cur.b; ++pos) {
                const Tri t = load_tri(tv.tris, pos);
                if (COUNT)
                    cnt.tris++;
                float tt, tu, tvv, ts;
                if (ray_tri(wr, org, t.p, t.q, t.r, tt, tu, tvv, ts)) {
                    if (tt >= 0.0f && tt < min_t) {
                        min_t = tt;
                        out_sign = ts;
                        hit = true;
                    }
                }
            }
        } else {
            const Pair pr = load_pair(tv.pairs, cur.a, tv.n);
            if (COUNT)
                cnt.pairs++;
            float tl, tr;
            if (ray_aabb_axis(oa, o1, o2, axis, a1, a2, pr.llo, pr.lhi, tl)) {
                stack[top] = pr.left;
                stack_t[top] = tl;
                ++top;
            }
            if (ray_aabb_axis(oa, o1, o2, axis, a1, a2, pr.rlo, pr.rhi, tr)) {
                stack[top] = pr.right;
                stack_t[top] = tr;
                ++top;
            }
        }
    }
    return hit;
}

// The same probe, walked NEAR CHILD FIRST.  The reference's walk (above) visits nodes in a fixed order -- right subtree
// before left, whatever the distance -- so its running bound tightens late.  Its ANSWER, though, does not depend on
// that order: as long as a box is entered no later than any triangle inside it (t_entry <= t_hit), the walk ends on
// the FIRST triangle, in its fixed order, among those with the smallest t.  That order is a function of sorted
// positions alone -- leaves with a larger start come first, positions inside a leaf ascend -- so any traversal that
// keeps subtrees whose entry is <= the bound (ties included) and breaks equal t by that rank returns the same
// triangle, hence the same sign.  Near-first with early pruning reaches the hit after a few dozen nodes instead of
// several hundred.  (-DWB_PROBE_REFERENCE_ORDER=1 builds the reference-order walk instead.)
template <bool COUNT>
__device__ __forceinline__ bool probe_sign_ordered(const TreeView& tv, const TreeHeader& h, float3 org, int axis,
                                                   float& out_sign, Counters& cnt)
{
    const float3 dir = make_float3(axis == 0 ? 1.f : 0.f, axis == 1 ? 1.f : 0.f, axis == 2 ? 1.f : 0.f);
    const WoopRay wr = woop_setup(dir);
    const int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;
    const float oa = wb_get(org, axis), o1 = wb_get(org, a1), o2 = wb_get(org, a2);

    Entry stack[WB_QUERY_STACK];
    float stack_t[WB_QUERY_STACK];
    int top = 0;
    float min_t = FLT_MAX;
    uint32_t best_leaf = 0, best_pos = 0;
    bool hit = false;

    Entry cur;
    float cur_t;
    if (!ray_aabb_axis(oa, o1, o2, axis, a1, a2, make_float3(h.lx, h.ly, h.lz), make_float3(h.hx, h.hy, h.hz), cur_t))
        return false;
    if (h.root_ref & WB_LEAF)
        cur.a = WB_LEAF | 0u, cur.b = h.root_count;
    else
        cur.a = (h.root_ref & WB_IDX_MASK) - (uint32_t)tv.n, cur.b = 0;
    bool have = true;
    for (;;) {
        if (!have) {
            if (top == 0)
                break;
            --top;
            cur = stack[top];
            cur_t = stack_t[top];
        }
        have = false;
        if (cur_t > min_t)
            continue;
        if (cur.a & WB_LEAF) {
            const uint32_t start = cur.a & WB_IDX_MASK;
            for (uint32_t pos = start; pos < start + cur.b; ++pos) {
                const Tri t = load_tri(tv.tris, pos);
                if (COUNT)
                    cnt.tris++;
                float tt, tu, tvv, ts;
                if (ray_tri(wr, org, t.p, t.q, t.r, tt, tu, tvv, ts) && tt >= 0.0f) {
                    // strictly nearer, or equally near and earlier in the reference's visiting order
                    const bool earlier = start > best_leaf || (start == best_leaf && pos < best_pos);
                    if (tt < min_t || (hit && tt == min_t && earlier)) {
                        min_t = tt;
                        out_sign = ts;
                        best_leaf = start, best_pos = pos;
                        hit = true;
                    }
                }
            }
            continue;
        }
        const Pair pr = load_pair(tv.pairs, cur.a, tv.n);
        if (COUNT)
            cnt.pairs++;
        float tl, tr;
        const bool hl = ray_aabb_axis(oa, o1, o2, axis, a1, a2, pr.llo, pr.lhi, tl) && tl <= min_t;
        const bool hr = ray_aabb_axis(oa, o1, o2, axis, a1, a2, pr.rlo, pr.rhi, tr) && tr <= min_t;
        if (hl && hr) {
            const bool near_left = tl < tr;
            stack[top] = near_left ? pr.right : pr.left;
            stack_t[top] = near_left ? tr : tl;
            ++top;
            cur = near_left ? pr.left : pr.right;
            cur_t = near_left ? tl : tr;
            have = true;
        } else if (hl) {
            cur = pr.left, cur_t = tl, have = true;
        } else if (hr) {
            cur = pr.right, cur_t = tr, have = true;
        }
    }
    return hit;
}

#ifndef WB_PROBE_REFERENCE_ORDER
#define WB_PROBE_REFERENCE_ORDER 0
#endif
#ifndef WB_SIGN_ALL_PROBES
#define WB_SIGN_ALL_PROBES 0
#endif

#ifndef WB_QP_MIN_BLOCKS
#define WB_QP_MIN_BLOCKS 10  // measured on C2: 9 (ptxas default, 52 registers) 653, 10: 668, 11: 617, 12: 617, 16: 464 M queries/s
#endif
// launch geometry of the signed variant (three extra probe traversals per query): see WB_QT_SIGN below
#ifndef WB_QT_SIGN
#define WB_QT_SIGN 128  // with the near-first probes (C2, M queries/s): 256 x 4 / 5 / 6: 194 / 208 / 186, 128 x 8 / 9 / 10 / 11 / 12:
#endif                  // 202 / 207 / 214 / 192 / 191, 64 x 20: 214
#ifndef WB_SIGN_MINB
#define WB_SIGN_MINB 10
#endif
constexpr int QT_SIGN = WB_QT_SIGN;
// MODE bits (A/B-measured variants of the memory side of the kernel):
//   1 QM_STACK16  16-byte stack entries (see closest_point)
//   2 QM_STREAM   evict-first loads of the permutation / points and streaming stores of the results (ld.global.cs /
//                 st.global.cs): 419 MB of once-touched I/O per batch should not displace the 147 MB tree from the 126 MB L2
//   4 QM_PACKED   ordered batches write ONE 16-byte {face, u, v, result} record per query at packed[perm[slot]] instead of
//                 four 4-byte scatters (sector-granular: 330 B of DRAM writes per query); k_unpack_results then streams the
//                 records into the caller's SoA arrays
//  16 QM_STACK8   8-byte stack entries (see pack_entry); combined with 0, 2 or 6 only
//   8 QM_STAGED   the batch is gathered into curve order once (k_gather_points) and every block stages its 128 points
//                 (1536 contiguous bytes) into shared memory with ONE TMA bulk copy (cp.async.bulk + mbarrier)
#define QM_STACK16 1
#define QM_STREAM 2
#define QM_PACKED 4
#define QM_STAGED 8

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool SIGN, bool COUNT, int MODE = 0>
__global__ void __launch_bounds__(SIGN ? QT_SIGN : QT, SIGN ? WB_SIGN_MINB : WB_QP_MIN_BLOCKS)
k_query_point(TreeView tv, const float* __restrict__ pts, const int* __restrict__ perm, long long nq, float max_dist,
              uint8_t* __restrict__ result, float* __restrict__ sign, int* __restrict__ face, float* __restrict__ u,
              float* __restrict__ v, unsigned long long* __restrict__ stats, uint4* __restrict__ packed = nullptr,
              const float* __restrict__ sorted_pts = nullptr)
{
    const TreeHeader h = *tv.header;
    Counters cnt;
    constexpr int T = SIGN ? QT_SIGN : QT;
    constexpr bool STAGED = (MODE & QM_STAGED) != 0;
    __shared__ __align__(16) float s_pts[STAGED ? 3 * T : 4];
    __shared__ __align__(8) unsigned long long s_bar[1];
    uint32_t phase = 0;
    if (STAGED) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    for (long long base = (long long)blockIdx.x * T; base < nq; base += (long long)gridDim.x * T) {
        const long long slot = base + threadIdx.x;
        if (STAGED) {
            // one elected thread issues the bulk copy of this block's 12 * T contiguous bytes (16-byte aligned: T % 4 == 0)
            const long long count = (nq - base < T) ? nq - base : T;
            const uint32_t bytes = (uint32_t)((12 * count + 15) & ~15ll);  // the staging buffer of the batch is padded to 16 B
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(s_bar)), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(s_pts)),
                             "l"(sorted_pts + 3 * base), "r"(bytes), "r"(smem_u32(s_bar))
                             : "memory");
            }
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done)
                             : "r"(smem_u32(s_bar)), "r"(phase)
                             : "memory");
            }
            phase ^= 1u;
        }
        if (slot < nq) {
        // `perm` (optional) is a Morton ordering of the batch: thread `slot` answers query perm[slot]
        const long long i = perm ? (long long)((MODE & QM_STREAM) ? __ldcs(perm + slot) : __ldg(perm + slot)) : slot;
        float3 p;
        if (STAGED)
            p = make_float3(s_pts[3 * threadIdx.x], s_pts[3 * threadIdx.x + 1], s_pts[3 * threadIdx.x + 2]);
        else if (MODE & QM_STREAM)
            p = make_float3(__ldcs(pts + 3 * i), __ldcs(pts + 3 * i + 1), __ldcs(pts + 3 * i + 2));
        else
            p = make_float3(__ldg(pts + 3 * i), __ldg(pts + 3 * i + 1), __ldg(pts + 3 * i + 2));
        int f = 0;
        float bu = 0.f, bv = 0.f, sg = 0.f;
        const bool ok = closest_point<COUNT, (MODE & QM_STACK16) != 0, (MODE & 16) != 0>(tv, h, p, max_dist, f, bu, bv, cnt);
        if (SIGN && ok) {  // majority of three axis probes, mesh.h:2342-2359
            int votes = 0;
            float s = 0.f;
            // the +z probe is only cast when +x and +y disagree: with 0 or 2 votes after two probes the majority is
            // settled (the reference casts all three, mesh.h:2349-2357; same answer)
#pragma unroll 1
            for (int axis = 0; axis < 3; ++axis) {
                if (!WB_SIGN_ALL_PROBES && axis == 2 && votes != 1)
                    break;
                if ((WB_PROBE_REFERENCE_ORDER ? probe_sign<COUNT>(tv, h, p, axis, s, cnt)
                                              : probe_sign_ordered<COUNT>(tv, h, p, axis, s, cnt))
                    && s < 0.f)
                    votes++;
            }
            sg = votes >= 2 ? -1.0f : 1.0f;
        }
        if ((MODE & QM_PACKED) && packed) {
            const uint4 rec = make_uint4(ok ? (uint32_t)f : 0u, __float_as_uint(ok ? bu : 0.f), __float_as_uint(ok ? bv : 0.f),
                                         (ok ? 1u : 0u) | (SIGN ? (__float_as_uint(sg) & 0x80000000u) | (ok ? 2u : 0u) : 0u));
            __stcs(packed + i, rec);
        } else if (MODE & QM_STREAM) {
            __stcs(result + i, (uint8_t)(ok ? 1 : 0));
            __stcs(face + i, ok ? f : 0);
            __stcs(u + i, ok ? bu : 0.f);
            __stcs(v + i, ok ? bv : 0.f);
            if (sign)
                __stcs(sign + i, sg);
        } else {
            result[i] = ok ? 1 : 0;
            face[i] = ok ? f : 0;
            u[i] = ok ? bu : 0.f;
            v[i] = ok ? bv : 0.f;
            if (sign)
                sign[i] = sg;
        }
        }
        if (STAGED)
            __syncthreads();  // everyone has read its point before the next bulk copy overwrites the buffer
    }
    if (COUNT) {
        atomicAdd(stats + 0, cnt.pairs);
        atomicAdd(stats + 1, cnt.tris);
    }
}

// QM_PACKED: records in ORIGINAL query order -> the caller's SoA arrays (coalesced reads and writes)
__global__ void __launch_bounds__(256)
k_unpack_results(const uint4* __restrict__ packed, long long nq, uint8_t* __restrict__ result, int* __restrict__ face,
                 float* __restrict__ u, float* __restrict__ v)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= nq)
        return;
    const uint4 r = __ldcs(packed + i);
    __stcs(result + i, (uint8_t)(r.w & 1u));
    __stcs(face + i, (int)r.x);
    __stcs(u + i, __uint_as_float(r.y));
    __stcs(v + i, __uint_as_float(r.z));
}

// QM_STAGED: the batch in curve order (12 B per point; the buffer is padded so that every block's slice may be read in
// whole 16-byte units)
__global__ void __launch_bounds__(256)
k_gather_points(const float* __restrict__ pts, const int* __restrict__ perm, long long nq, float* __restrict__ sorted)
{
    const long long s = (long long)blockIdx.x * 256 + threadIdx.x;
    if (s >= nq)
        return;
    const long long i = __ldcs(perm + s);
    sorted[3 * s + 0] = __ldcs(pts + 3 * i);
    sorted[3 * s + 1] = __ldcs(pts + 3 * i + 1);
    sorted[3 * s + 2] = __ldcs(pts + 3 * i + 2);
}

// start entry of a ray traversal: the tree root, or the caller's root node (mesh_query_ray(..., root), mesh.h:1768)
__device__ __forceinline__ Entry ray_start_entry(const TreeView& tv, const TreeHeader& h, const int* __restrict__ roots,
                                                 long long i, float3& lo, float3& hi)
{
    Entry e;
    if (h.root_ref & WB_LEAF)
        e.a = WB_LEAF | 0u, e.b = h.root_count;
    else
        e.a = (h.root_ref & WB_IDX_MASK) - (uint32_t)tv.n, e.b = 0;
    lo = make_float3(h.lx, h.ly, h.lz), hi = make_float3(h.hx, h.hy, h.hz);
    const int r = roots ? __ldg(roots + i) : -1;
    if (r >= 0 && r < tv.n) {  // an original leaf: one triangle
        e.a = (uint32_t)r | WB_LEAF, e.b = 1u;
        const Tri t = load_tri(tv.tris, (uint32_t)r);
        lo = wb_min3(wb_min3(t.p, t.q), t.r), hi = wb_max3(wb_max3(t.p, t.q), t.r);
    } else if (r >= tv.n) {
        wb_root_entry(tv, r, e.a, e.b, lo, hi);
    }
    return e;
}

// ------------------------------------------------------------------------------------------------
// closest ray hit, near child first (mesh.h:1735-1891)
// ------------------------------------------------------------------------------------------------
#ifndef WB_QR_MIN_BLOCKS
#define WB_QR_MIN_BLOCKS 10  // measured on C3: unhinted 2.98, (QT, 1) 2.88, 8: 2.86, 10: 3.01, 12: 2.83 G rays/s
#endif
// PACK: 4-byte stack entries (pack_entry) for trees below 2^28 items -- 128 B of local memory per ray instead of 256 B
template <bool COUNT, bool PACK>
__global__ void __launch_bounds__(QT, WB_QR_MIN_BLOCKS)
k_query_ray(TreeView tv, const float* __restrict__ starts, const float* __restrict__ dirs, const int* __restrict__ perm,
            const int* __restrict__ roots, long long nq, float max_t, uint8_t* __restrict__ result, float* __restrict__ sign, int* __restrict__ face,
            float* __restrict__ out_t, float* __restrict__ out_u, float* __restrict__ out_v, float* __restrict__ normal,
            unsigned long long* __restrict__ stats)
{
    const TreeHeader h = *tv.header;
    Counters cnt;
    for (long long slot = (long long)blockIdx.x * QT + threadIdx.x; slot < nq; slot += (long long)gridDim.x * QT) {
        // `perm` (optional): origin/direction ordering of the batch, thread `slot` traces ray perm[slot]
        const long long i = perm ? (long long)__ldg(perm + slot) : slot;
        const float3 org = make_float3(__ldg(starts + 3 * i), __ldg(starts + 3 * i + 1), __ldg(starts + 3 * i + 2));
        const float3 dir = make_float3(__ldg(dirs + 3 * i), __ldg(dirs + 3 * i + 1), __ldg(dirs + 3 * i + 2));
        float3 safe = dir;
        if (safe.x == 0.0f)
            safe.x = 1.0e-20f;
        if (safe.y == 0.0f)
            safe.y = 1.0e-20f;
        if (safe.z == 0.0f)
            safe.z = 1.0e-20f;
        const float3 rcp = make_float3(1.0f / safe.x, 1.0f / safe.y, 1.0f / safe.z);
        const bool fast = dir.x != 0.0f && dir.y != 0.0f && dir.z != 0.0f;
        const WoopRay wr = woop_setup(dir);

        Entry stack[PACK ? 1 : WB_QUERY_STACK];
        uint32_t stack4[PACK ? WB_QUERY_STACK : 1];
        int top = 0;
        float3 start_lo, start_hi;  // the start node's box is not tested (mesh.h:1779)
        Entry cur = ray_start_entry(tv, h, roots, i, start_lo, start_hi);

        float min_t = max_t, min_u = 0.f, min_v = 0.f, min_sign = 1.0f;
        int min_face = 0;
        uint32_t min_pos = 0;
        bool hit = false;

        for (;;) {
            if (cur.a & WB_LEAF) {
                const uint32_t start = cur.a & WB_IDX_MASK;
                for (uint32_t pos = start; pos < start + cur.b; ++pos) {
                    const Tri t = load_tri(tv.tris, pos);
                    if (COUNT)
                        cnt.tris++;
                    float tt, tu, tvv, ts;
                    if (ray_tri(wr, org, t.p, t.q, t.r, tt, tu, tvv, ts)) {
                        if (tt < min_t && tt >= 0.0f) {
                            min_t = tt, min_face = t.face, min_u = tu, min_v = tvv, min_sign = ts, min_pos = pos;
                            hit = true;
                        }
                    }
                }
                if (top == 0)
                    break;
                --top;
                cur = PACK ? unpack_entry(tv, stack4[top]) : stack[top];
                continue;
            }
            const Pair pr = load_pair(tv.pairs, cur.a, tv.n);
            if (COUNT)
                cnt.pairs++;
            float t0 = FLT_MAX, t1 = FLT_MAX;
            const bool h0 = (fast ? ray_aabb_fast(org, rcp, pr.llo, pr.lhi, t0)
                                  : ray_aabb_robust(org, dir, rcp, pr.llo, pr.lhi, t0))
                && t0 < min_t;
            const bool h1 = (fast ? ray_aabb_fast(org, rcp, pr.rlo, pr.rhi, t1)
                                  : ray_aabb_robust(org, dir, rcp, pr.rlo, pr.rhi, t1))
                && t1 < min_t;
            if (h0 && h1) {
                const bool near_left = t0 < t1;
                if (top >= WB_QUERY_STACK)
                    break;  // mesh.h:1860-1861
                if (PACK)
                    stack4[top++] = pack_entry(near_left ? pr.right : pr.left);
                else
                    stack[top++] = near_left ? pr.right : pr.left;
                cur = near_left ? pr.left : pr.right;
            } else if (h0) {
                cur = pr.left;
            } else if (h1) {
                cur = pr.right;
            } else {
                if (top == 0)
                    break;
                --top;
                cur = PACK ? unpack_entry(tv, stack4[top]) : stack[top];
            }
        }

        float3 nrm = make_float3(0.f, 0.f, 0.f);
        if (hit) {
            const Tri t = load_tri(tv.tris, min_pos);
            const float3 g = wb_cross(wb_sub(t.q, t.p), wb_sub(t.r, t.p));
            const float l = sqrtf(g.x * g.x + g.y * g.y + g.z * g.z);
            if (l > 0.0f)
                nrm = make_float3(g.x / l, g.y / l, g.z / l);
        }
        result[i] = hit ? 1 : 0;
        sign[i] = hit ? min_sign : 0.f;
        face[i] = hit ? min_face : 0;
        out_t[i] = hit ? min_t : 0.f;
        out_u[i] = hit ? min_u : 0.f;
        out_v[i] = hit ? min_v : 0.f;
        normal[3 * i + 0] = nrm.x;
        normal[3 * i + 1] = nrm.y;
        normal[3 * i + 2] = nrm.z;
    }
    if (COUNT) {
        atomicAdd(stats + 0, cnt.pairs);
        atomicAdd(stats + 1, cnt.tris);
    }
}

// mesh_query_ray_count_intersections (mesh.h:1976-2032): triangles hit with t >= 0 along the whole ray; push-both
// traversal with the robust slab test.  Children are tested before the push (the reference tests on pop: same count,
// shallower stack); the reference has no overflow check (undefined past 32 entries), subtrees that do not fit are dropped.
__device__ __forceinline__ int count_ray_hits(const TreeView& tv, float3 org, float3 dir, Entry root, float3 root_lo,
                                              float3 root_hi)
{
    const WoopRay wr = woop_setup(dir);
    Entry stack[WB_QUERY_STACK];
    int top = 0;
    const float3 rcp = make_float3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    int hits = 0;
    float tt;
    if (ray_aabb_robust(org, dir, rcp, root_lo, root_hi, tt))
        stack[top++] = root;
    while (top) {
        const Entry cur = stack[--top];
        if (cur.a & WB_LEAF) {
            const uint32_t start = cur.a & WB_IDX_MASK;
            for (uint32_t pos = start; pos < start + cur.b; ++pos) {
                const Tri t = load_tri(tv.tris, pos);
                float t_hit, tu, tvv, ts;
                if (ray_tri(wr, org, t.p, t.q, t.r, t_hit, tu, tvv, ts) && t_hit >= 0.0f)
                    hits++;
            }
        } else {
            const Pair pr = load_pair(tv.pairs, cur.a, tv.n);
            if (top < WB_QUERY_STACK && ray_aabb_robust(org, dir, rcp, pr.llo, pr.lhi, tt))
                stack[top++] = pr.left;
            if (top < WB_QUERY_STACK && ray_aabb_robust(org, dir, rcp, pr.rlo, pr.rhi, tt))
                stack[top++] = pr.right;
        }
    }
    return hits;
}

// ------------------------------------------------------------------------------------------------
// further ray queries sharing the traversal core
//   ANYHIT: mesh_query_ray_anyhit (mesh.h:1893-1974) -- is there any hit with 0 <= t < max_t
//   COUNT : mesh_query_ray_count_intersections (mesh.h:1976-2032) -- number of hits with t >= 0 along the
//           whole ray (robust slab test, push-both traversal, no distance bound)
// Both answers are order independent (a fixed bound, no running minimum), so they equal the reference's.
// ------------------------------------------------------------------------------------------------
template <bool COUNT_MODE>
__global__ void __launch_bounds__(QT)
k_query_ray_aux(TreeView tv, const float* __restrict__ starts, const float* __restrict__ dirs,
                const int* __restrict__ roots, long long nq, float max_t, uint8_t* __restrict__ any_out,
                int* __restrict__ count_out)
{
    const TreeHeader h = *tv.header;
    for (long long i = (long long)blockIdx.x * QT + threadIdx.x; i < nq; i += (long long)gridDim.x * QT) {
        const float3 org = make_float3(__ldg(starts + 3 * i), __ldg(starts + 3 * i + 1), __ldg(starts + 3 * i + 2));
        const float3 dir = make_float3(__ldg(dirs + 3 * i), __ldg(dirs + 3 * i + 1), __ldg(dirs + 3 * i + 2));
        const WoopRay wr = woop_setup(dir);
        Entry stack[WB_QUERY_STACK];
        int top = 0;
        float3 start_lo, start_hi;
        const Entry root = ray_start_entry(tv, h, roots, i, start_lo, start_hi);

        if (COUNT_MODE) {
            count_out[i] = count_ray_hits(tv, org, dir, root, start_lo, start_hi);
        } else {
            float3 safe = dir;
            if (safe.x == 0.0f)
                safe.x = 1.0e-20f;
            if (safe.y == 0.0f)
                safe.y = 1.0e-20f;
            if (safe.z == 0.0f)
                safe.z = 1.0e-20f;
            const float3 rcp = make_float3(1.0f / safe.x, 1.0f / safe.y, 1.0f / safe.z);
            const bool fast = dir.x != 0.0f && dir.y != 0.0f && dir.z != 0.0f;
            Entry cur = root;
            bool found = false;
            for (;;) {
                if (cur.a & WB_LEAF) {
                    const uint32_t start = cur.a & WB_IDX_MASK;
                    for (uint32_t pos = start; pos < start + cur.b && !found; ++pos) {
                        const Tri t = load_tri(tv.tris, pos);
                        float t_hit, tu, tvv, ts;
                        if (ray_tri(wr, org, t.p, t.q, t.r, t_hit, tu, tvv, ts) && t_hit < max_t && t_hit >= 0.0f)
                            found = true;
                    }
                    if (found || top == 0)
                        break;
                    cur = stack[--top];
                    continue;
                }
                const Pair pr = load_pair(tv.pairs, cur.a, tv.n);
                float t0 = FLT_MAX, t1 = FLT_MAX;
                const bool h0 = (fast ? ray_aabb_fast(org, rcp, pr.llo, pr.lhi, t0)
                                      : ray_aabb_robust(org, dir, rcp, pr.llo, pr.lhi, t0))
                    && t0 < max_t;
                const bool h1 = (fast ? ray_aabb_fast(org, rcp, pr.rlo, pr.rhi, t1)
                                      : ray_aabb_robust(org, dir, rcp, pr.rlo, pr.rhi, t1))
                    && t1 < max_t;
                if (h0 && h1) {
                    if (top >= WB_QUERY_STACK)
                        break;  // mesh.h:1957-1958 returns false
                    const bool near_left = t0 < t1;
                    stack[top++] = near_left ? pr.right : pr.left;
                    cur = near_left ? pr.left : pr.right;
                } else if (h0) {
                    cur = pr.left;
                } else if (h1) {
                    cur = pr.right;
                } else {
                    if (top == 0)
                        break;
                    cur = stack[--top];
                }
            }
            any_out[i] = found ? 1 : 0;
        }
    }
}

// sign of mesh_query_point_sign_parity (mesh.h:309-498, 2362-2392): n_sample rays from the point along
// (1,1,1) + U(-scale, scale)^3, drawn from the PCG stream seeded with 42 (rand.h:29-80: the same directions for every
// query), inside when at least half of them cross an odd number of faces.  The three offsets of a direction are
// separate randf() calls inside one constructor call in the reference; its device builds (nvcc / NVRTC, and clang for
// the CPU JIT) evaluate them left to right -- x first -- which is what this does.
__device__ __forceinline__ uint32_t rand_pcg(uint32_t state)
{
    const uint32_t b = state * 747796405u + 2891336453u;
    const uint32_t c = ((b >> ((b >> 28u) + 4u)) ^ b) * 277803737u;
    return (c >> 22u) ^ c;
}
__device__ __forceinline__ float randf_range(uint32_t& state, float lo, float hi)
{
    state = rand_pcg(state);
    return (hi - lo) * ((state >> 8) * (1.0f / 16777216.0f)) + lo;
}

__global__ void __launch_bounds__(QT)
k_sign_parity(TreeView tv, const float* __restrict__ pts, long long nq, int n_sample, float scale,
              const uint8_t* __restrict__ result, float* __restrict__ sign)
{
    const TreeHeader h = *tv.header;
    for (long long i = (long long)blockIdx.x * QT + threadIdx.x; i < nq; i += (long long)gridDim.x * QT) {
        if (!result[i]) {
            sign[i] = 0.0f;
            continue;
        }
        const float3 p = make_float3(__ldg(pts + 3 * i), __ldg(pts + 3 * i + 1), __ldg(pts + 3 * i + 2));
        uint32_t state = rand_pcg(42u);
        int vote = 0;
        for (int k = 0; k < n_sample; ++k) {
            float3 dir;
            do {
                const float rx = randf_range(state, -scale, scale);
                const float ry = randf_range(state, -scale, scale);
                const float rz = randf_range(state, -scale, scale);
                dir = make_float3(1.0f + rx, 1.0f + ry, 1.0f + rz);
            } while (dir.x * dir.x + dir.y * dir.y + dir.z * dir.z < 1e-8f);
            float3 rlo, rhi;
            const Entry root = ray_start_entry(tv, h, nullptr, i, rlo, rhi);
            if (count_ray_hits(tv, p, dir, root, rlo, rhi) % 2)
                vote++;
        }
        sign[i] = (vote * 2 >= n_sample) ? -1.0f : 1.0f;
    }
}

// ------------------------------------------------------------------------------------------------
// mesh_query_point_sign_normal (mesh.h:860-1090): closest point with an epsilon band -- every triangle whose distance
// is within eps = average_edge_length * epsilon of the running minimum adds its angle-weighted normal (vertex: the
// corner angle, edge: pi, interior: 2 pi); sign = +1 when the accumulated normal points towards the query.
// Works on distances, not squared distances, like the reference.  The average edge length (mesh.cu:38-60) is reduced
// in a fixed order (per-thread grid-stride sums in double, block tree, one finishing block), so it is run-to-run
// deterministic; the reference sums the same float terms with a CUB scan (GPU) or a float loop (CPU).
// ------------------------------------------------------------------------------------------------
constexpr int EDGE_BLOCKS = 296;

__global__ void __launch_bounds__(256)
k_edge_partials(const float* __restrict__ points, const int* __restrict__ indices, int n, double* __restrict__ partials)
{
    __shared__ double sm[256];
    double acc = 0.0;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long long)gridDim.x * 256) {
        float3 p, q, r;
        MeshSource { points, indices }.tri((int)t, p, q, r);
        const float3 pq = wb_sub(p, q), pr = wb_sub(p, r), qr = wb_sub(q, r);
        acc += (double)(sqrtf(wb_dot(pq, pq)) + sqrtf(wb_dot(pr, pr)) + sqrtf(wb_dot(qr, qr)));
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
            sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        partials[blockIdx.x] = sm[0];
}

__global__ void k_edge_finish(const double* __restrict__ partials, int nb, int n, float* __restrict__ out)
{
    double total = 0.0;
    for (int b = 0; b < nb; ++b)
        total += partials[b];
    *out = (float)total / (float)(3 * n);
}

__device__ __forceinline__ float3 normalize3(float3 a)  // vec.h:1111-1118 (kEps = 0)
{
    const float l = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    if (l > 0.0f)
        return make_float3(a.x / l, a.y / l, a.z / l);
    return make_float3(0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(QT)
k_query_point_sign_normal(TreeView tv, const float* __restrict__ pts, const int* __restrict__ perm, long long nq,
                          float max_dist, float epsilon, const float* __restrict__ avg_edge, uint8_t* __restrict__ result,
                          float* __restrict__ sign, int* __restrict__ face, float* __restrict__ u, float* __restrict__ v)
{
    const TreeHeader h = *tv.header;
    const float eps = *avg_edge * epsilon;
    const float eps_sq = eps * eps;
    for (long long slot = (long long)blockIdx.x * QT + threadIdx.x; slot < nq; slot += (long long)gridDim.x * QT) {
        const long long qi = perm ? (long long)__ldg(perm + slot) : slot;
        const float3 point = make_float3(__ldg(pts + 3 * qi), __ldg(pts + 3 * qi + 1), __ldg(pts + 3 * qi + 2));

        Entry stack[WB_QUERY_STACK];
        float stack_d[WB_QUERY_STACK];
        int top = 0;
        float min_dist = max_dist;
        int min_face = 0;
        float min_v = 0.f, min_w = 0.f;
        float3 min_p = make_float3(0.f, 0.f, 0.f), min_q = min_p, min_r = min_p;
        float3 acc = make_float3(0.f, 0.f, 0.f);

        Entry cur;
        if (h.root_ref & WB_LEAF)
            cur.a = WB_LEAF | 0u, cur.b = h.root_count;
        else
            cur.a = (h.root_ref & WB_IDX_MASK) - (uint32_t)tv.n, cur.b = 0;
        float cur_d = dist_aabb_sq(point, make_float3(h.lx, h.ly, h.lz), make_float3(h.hx, h.hy, h.hz));
        bool have = true;
        for (;;) {
            if (!have) {
                if (top == 0)
                    break;
                --top;
                cur = stack[top];
                cur_d = stack_d[top];
            }
            have = false;
            if (cur_d > (min_dist + eps) * (min_dist + eps))
                continue;
            if (cur.a & WB_LEAF) {
                const uint32_t start = cur.a & WB_IDX_MASK;
                for (uint32_t pos = start; pos < start + cur.b; ++pos) {
                    const Tri t = load_tri(tv.tris, pos);
                    if (t.flags & WB_TRI_SLIVER)
                        continue;
                    const float3 e0 = wb_sub(t.q, t.p), e1 = wb_sub(t.r, t.p), e2 = wb_sub(t.r, t.q);
                    const float3 normal = wb_cross(e0, e1);
                    const float e0n = wb_dot(e0, e0), e1n = wb_dot(e1, e1), e2n = wb_dot(e2, e2);
                    float bv, bw;
                    closest_vw(t.p, t.q, t.r, point, bv, bw);
                    const float bu = 1.0f - bv - bw;
                    const float w = 1.f - bu - bv;
                    const float3 c = wb_add(wb_add(wb_scale(bu, t.p), wb_scale(bv, t.q)), wb_scale(w, t.r));
                    const float3 d = wb_sub(c, point);
                    const float dist = sqrtf(wb_dot(d, d));
                    if (dist < min_dist + eps) {
                        float weight;
                        const float3 cp = wb_sub(c, t.p), cq = wb_sub(c, t.q), cr = wb_sub(c, t.r);
                        const float lcp = wb_dot(cp, cp), lcq = wb_dot(cq, cq), lcr = wb_dot(cr, cr);
                        const float3 neg_e0 = make_float3(-e0.x, -e0.y, -e0.z), neg_e1 = make_float3(-e1.x, -e1.y, -e1.z),
                                     neg_e2 = make_float3(-e2.x, -e2.y, -e2.z);
                        if (lcp < eps_sq) {
                            weight = acosf(wb_dot(normalize3(e0), normalize3(e1)));
                        } else if (lcq < eps_sq) {
                            weight = acosf(wb_dot(normalize3(e2), normalize3(neg_e0)));
                        } else if (lcr < eps_sq) {
                            weight = acosf(wb_dot(normalize3(neg_e1), normalize3(neg_e2)));
                        } else {
                            const float e0cp = wb_dot(e0, cp), e2cq = wb_dot(e2, cq), e1cp = wb_dot(e1, cp);
                            if ((lcp * e0n - e0cp * e0cp < eps_sq * e0n) || (lcq * e2n - e2cq * e2cq < eps_sq * e2n)
                                || (lcp * e1n - e1cp * e1cp < eps_sq * e1n))
                                weight = 3.14159265359f;
                            else
                                weight = 2.0f * 3.14159265359f;
                        }
                        const float3 wn = wb_scale(weight, normalize3(normal));
                        if (dist > min_dist - eps) {  // treated as equal: accumulate
                            acc = wb_add(acc, wn);
                            if (dist < min_dist)
                                min_dist = dist, min_v = bv, min_w = w, min_face = t.face, min_p = t.p, min_q = t.q, min_r = t.r;
                        } else {
                            min_dist = dist, min_v = bv, min_w = w, min_face = t.face, min_p = t.p, min_q = t.q, min_r = t.r;
                            acc = wn;
                        }
                    }
                }
                continue;
            }
            const Pair pr = load_pair(tv.pairs, cur.a, tv.n);
            const float dl = dist_aabb_sq(point, pr.llo, pr.lhi), dr = dist_aabb_sq(point, pr.rlo, pr.rhi);
            Entry far_e, near_e;
            float far_d, near_d;
            if (dl < dr)
                far_e = pr.right, far_d = dr, near_e = pr.left, near_d = dl;
            else
                far_e = pr.left, far_d = dl, near_e = pr.right, near_d = dr;
            const float bound = (min_dist + eps) * (min_dist + eps);
            if (far_d < bound) {
                stack[top] = far_e;
                stack_d[top] = far_d;
                ++top;
            }
            if (near_d < bound) {
                cur = near_e;
                cur_d = near_d;
                have = true;
            }
        }
        const bool ok = min_dist < max_dist;
        float bu = 0.f, sg = 0.f;
        if (ok) {
            bu = 1.0f - min_v - min_w;
            const float3 cpt = wb_add(wb_add(wb_scale(bu, min_p), wb_scale(min_v, min_q)), wb_scale(min_w, min_r));
            sg = wb_dot(acc, wb_sub(point, cpt)) > 0.0f ? 1.0f : -1.0f;
        }
        result[qi] = ok ? 1 : 0;
        sign[qi] = sg;
        face[qi] = ok ? min_face : 0;
        u[qi] = bu;
        v[qi] = ok ? min_v : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
// mesh_query_furthest_point_no_sign (mesh.h:678-858): the mirror image of the closest-point walk -- nodes are ranked
// by the distance to their farthest corner (mesh.h:100-119), the farther child is entered first, a node is skipped when
// even its farthest corner is nearer than the best so far, and the candidate on a triangle is its farthest VERTEX
// (intersect.h:111-125).  Updates on strict '>', result = best > min_dist^2.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float far_aabb_sq(float3 p, float3 lo, float3 hi)
{
    const float lx = fabsf(p.x - lo.x), ux = fabsf(p.x - hi.x), cx = (lx > ux) ? lx : ux;
    const float ly = fabsf(p.y - lo.y), uy = fabsf(p.y - hi.y), cy = (ly > uy) ? ly : uy;
    const float lz = fabsf(p.z - lo.z), uz = fabsf(p.z - hi.z), cz = (lz > uz) ? lz : uz;
    return cx * cx + cy * cy + cz * cz;
}

__global__ void __launch_bounds__(QT)
k_query_furthest(TreeView tv, const float* __restrict__ pts, long long nq, float min_dist, uint8_t* __restrict__ result,
                 int* __restrict__ face, float* __restrict__ u, float* __restrict__ v)
{
    const TreeHeader h = *tv.header;
    for (long long qi = (long long)blockIdx.x * QT + threadIdx.x; qi < nq; qi += (long long)gridDim.x * QT) {
        const float3 point = make_float3(__ldg(pts + 3 * qi), __ldg(pts + 3 * qi + 1), __ldg(pts + 3 * qi + 2));
        Entry stack[WB_QUERY_STACK];
        float stack_d[WB_QUERY_STACK];
        int top = 0;
        float best = min_dist * min_dist;
        int best_face = 0;
        float best_v = 0.f, best_w = 0.f;

        Entry cur;
        if (h.root_ref & WB_LEAF)
            cur.a = WB_LEAF | 0u, cur.b = h.root_count;
        else
            cur.a = (h.root_ref & WB_IDX_MASK) - (uint32_t)tv.n, cur.b = 0;
        float cur_d = far_aabb_sq(point, make_float3(h.lx, h.ly, h.lz), make_float3(h.hx, h.hy, h.hz));
        bool have = true;
        for (;;) {
            if (!have) {
                if (top == 0)
                    break;
                --top;
                cur = stack[top];
                cur_d = stack_d[top];
            }
            have = false;
            if (cur_d < best)
                continue;
            if (cur.a & WB_LEAF) {
                const uint32_t start = cur.a & WB_IDX_MASK;
                for (uint32_t pos = start; pos < start + cur.b; ++pos) {
                    const Tri t = load_tri(tv.tris, pos);
                    if (t.flags & WB_TRI_SLIVER)
                        continue;
                    const float3 pa = wb_sub(point, t.p), pb = wb_sub(point, t.q), pc = wb_sub(point, t.r);
                    const float da = wb_dot(pa, pa), db = wb_dot(pb, pb), dc = wb_dot(pc, pc);
                    float bu, bv;
                    if (da > db && da > dc)
                        bu = 1.0f, bv = 0.0f;
                    else if (db > dc)
                        bu = 0.0f, bv = 1.0f;
                    else
                        bu = 0.0f, bv = 0.0f;
                    const float w = 1.f - bu - bv;
                    const float3 c = wb_add(wb_add(wb_scale(bu, t.p), wb_scale(bv, t.q)), wb_scale(w, t.r));
                    const float3 d = wb_sub(c, point);
                    const float dsq = wb_dot(d, d);
                    if (dsq > best)
                        best = dsq, best_v = bv, best_w = w, best_face = t.face;
                }
                continue;
            }
            const Pair pr = load_pair(tv.pairs, cur.a, tv.n);
            const float dl = far_aabb_sq(point, pr.llo, pr.lhi), dr = far_aabb_sq(point, pr.rlo, pr.rhi);
            Entry first_e, second_e;  // second is entered now (the reference pushes it last and pops it first)
            float first_d, second_d;
            if (dl > dr)
                first_e = pr.right, first_d = dr, second_e = pr.left, second_d = dl;
            else
                first_e = pr.left, first_d = dl, second_e = pr.right, second_d = dr;
            if (first_d > best) {
                stack[top] = first_e;
                stack_d[top] = first_d;
                ++top;
            }
            if (second_d > best) {
                cur = second_e;
                cur_d = second_d;
                have = true;
            }
        }
        const bool ok = best > min_dist * min_dist;
        result[qi] = ok ? 1 : 0;
        face[qi] = ok ? best_face : 0;
        u[qi] = ok ? 1.0f - best_v - best_w : 0.f;
        v[qi] = ok ? best_v : 0.f;
    }
}

// mesh_eval_face_normal (mesh.h:2870-2888): normalize(cross(q - p, r - p)) from the caller's arrays
__global__ void __launch_bounds__(QT)
k_mesh_face_normal(const float* __restrict__ points, const int* __restrict__ indices, const int* __restrict__ face,
                   const uint8_t* __restrict__ mask, long long n, float* __restrict__ out)
{
    for (long long i = (long long)blockIdx.x * QT + threadIdx.x; i < n; i += (long long)gridDim.x * QT) {
        if (mask && !mask[i]) {  // a ray that missed: mesh_query_ray leaves the normal zero (mesh.h:2216-2248)
            out[3 * i + 0] = 0.f, out[3 * i + 1] = 0.f, out[3 * i + 2] = 0.f;
            continue;
        }
        float3 p, q, r;
        MeshSource { points, indices }.tri(face[i], p, q, r);
        const float3 nrm = wb_cross(wb_sub(q, p), wb_sub(r, p));
        const float l = sqrtf(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z);
        const bool nz = l > 0.0f;  // vec.h:1111-1118 with kEps = 0
        out[3 * i + 0] = nz ? nrm.x / l : 0.f;
        out[3 * i + 1] = nz ? nrm.y / l : 0.f;
        out[3 * i + 2] = nz ? nrm.z / l : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
// mesh_query_sphere + mesh_query_sphere_next (mesh.h:2457-2616, 2694-2737) for a batch: the faces that intersect the
// sphere, in the iterator's order (children pushed left then right, leaf items in leaf order).  Node test: exact
// sphere / box (intersect.h:197-205).  Face test, also on single-face leaves: the same test on the face's box, then the
// closest point of the triangle -- of its longest edge when the face has zero area -- within the radius.
// FILL = false counts, true writes the face indices at indices[offsets[i]...].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool sphere_box_sq(float3 c, float radius_sq, float3 lo, float3 hi)
{
    const float dx = fmaxf(fmaxf(lo.x - c.x, c.x - hi.x), 0.0f);
    const float dy = fmaxf(fmaxf(lo.y - c.y, c.y - hi.y), 0.0f);
    const float dz = fmaxf(fmaxf(lo.z - c.z, c.z - hi.z), 0.0f);
    return dx * dx + dy * dy + dz * dz <= radius_sq;
}

__device__ __forceinline__ bool sphere_face(float3 center, float radius_sq, const Tri& t)
{
    if (!sphere_box_sq(center, radius_sq, wb_min3(wb_min3(t.p, t.q), t.r), wb_max3(wb_max3(t.p, t.q), t.r)))
        return false;
    const float3 a = t.p, b = t.q, c = t.r;
    const float3 ab = wb_sub(b, a), ac = wb_sub(c, a);
    const float3 n = wb_cross(ab, ac);
    float3 cp;
    if (wb_dot(n, n) == 0.0f) {  // degenerate face: closest point of its longest edge (mesh.h:2589-2611)
        const float3 bc = wb_sub(c, b);
        const float lab2 = wb_dot(ab, ab), lac2 = wb_dot(ac, ac), lbc2 = wb_dot(bc, bc);
        float3 p, q;
        float len2;
        if (lab2 >= lac2 && lab2 >= lbc2)
            p = a, q = b, len2 = lab2;
        else if (lac2 >= lbc2)
            p = a, q = c, len2 = lac2;
        else
            p = b, q = c, len2 = lbc2;
        const float3 pq = wb_sub(q, p);
        const float tt = (len2 > 0.0f) ? fminf(fmaxf(0.0f, wb_dot(wb_sub(center, p), pq) / len2), 1.0f) : 0.0f;
        cp = wb_add(p, wb_scale(tt, pq));
    } else {
        float bv, bw;
        closest_vw(a, b, c, center, bv, bw);
        const float bu = 1.0f - bv - bw;
        cp = wb_add(wb_add(wb_scale(bu, a), wb_scale(bv, b)), wb_scale(1.0f - bu - bv, c));
    }
    const float3 d = wb_sub(cp, center);
    return wb_dot(d, d) <= radius_sq;
}

template <bool FILL>
__global__ void __launch_bounds__(QT)
k_mesh_query_sphere(TreeView tv, const float* __restrict__ centers, const float* __restrict__ radii, long long nq,
                    int* __restrict__ counts, const int* __restrict__ offsets, int* __restrict__ indices)
{
    const TreeHeader h = *tv.header;
    for (long long i = (long long)blockIdx.x * QT + threadIdx.x; i < nq; i += (long long)gridDim.x * QT) {
        const float3 c = make_float3(__ldg(centers + 3 * i), __ldg(centers + 3 * i + 1), __ldg(centers + 3 * i + 2));
        const float r = fmaxf(__ldg(radii + i), 0.0f);  // mesh.h:2525-2526
        const float r2 = r * r;
        int found = 0;
        int* out = FILL ? indices + offsets[i] : nullptr;
        Entry stack[WB_QUERY_STACK];
        int top = 0;
        if (sphere_box_sq(c, r2, make_float3(h.lx, h.ly, h.lz), make_float3(h.hx, h.hy, h.hz))) {
            if (h.root_ref & WB_LEAF)
                stack[0].a = WB_LEAF | 0u, stack[0].b = h.root_count;
            else
                stack[0].a = (h.root_ref & WB_IDX_MASK) - (uint32_t)tv.n, stack[0].b = 0;
            top = 1;
        }
        while (top) {
            const Entry cur = stack[--top];
            if (cur.a & WB_LEAF) {
                const uint32_t start = cur.a & WB_IDX_MASK;
                for (uint32_t pos = start; pos < start + cur.b; ++pos) {
                    const Tri t = load_tri(tv.tris, pos);
                    if (sphere_face(c, r2, t)) {
                        if (FILL)
                            out[found] = t.face;
                        ++found;
                    }
                }
                continue;
            }
            const Pair pr = load_pair(tv.pairs, cur.a, tv.n);
            if (sphere_box_sq(c, r2, pr.llo, pr.lhi))
                stack[top++] = pr.left;
            if (sphere_box_sq(c, r2, pr.rlo, pr.rhi))
                stack[top++] = pr.right;
        }
        if (!FILL)
            counts[i] = found;
    }
}

// mesh_eval_position / mesh_eval_velocity (mesh.h:2767-2807): p*u + q*v + r*(1-u-v) from the caller's arrays
__global__ void __launch_bounds__(QT)
k_mesh_eval(const float* __restrict__ attr, const int* __restrict__ indices, const int* __restrict__ face,
            const float* __restrict__ u, const float* __restrict__ v, long long n, float* __restrict__ out)
{
    for (long long i = (long long)blockIdx.x * QT + threadIdx.x; i < n; i += (long long)gridDim.x * QT) {
        const int f = face[i];
        const int a = __ldg(indices + 3 * (size_t)f), b = __ldg(indices + 3 * (size_t)f + 1), c = __ldg(indices + 3 * (size_t)f + 2);
        const float uu = u[i], vv = v[i], ww = 1.0f - uu - vv;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float p = __ldg(attr + 3 * (size_t)a + k), q = __ldg(attr + 3 * (size_t)b + k), r = __ldg(attr + 3 * (size_t)c + k);
            out[3 * i + k] = p * uu + q * vv + r * ww;
        }
    }
}

int query_grid(long long nq, int threads = QT)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (nq + threads - 1) / threads;
#ifndef WB_QGRID_CAP
#define WB_QGRID_CAP 4096  // blocks per SM before grid-striding: 16: 577, 64: 670, 256: 689, 4096: 694 M queries/s (C2)
#endif
    const long long cap = (long long)sms * WB_QGRID_CAP * QT / threads;  // grid-stride beyond that
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

namespace {
template <int MODE>
void launch_point_mode(const TreeView& tv, const float* pts, const int* perm, long long nq, float max_dist, uint8_t* result,
                       int* face, float* u, float* v, uint4* packed, float* sorted_pts, cudaStream_t stream)
{
    const int grid = query_grid(nq);
    if ((MODE & QM_STAGED) && perm && sorted_pts)
        k_gather_points<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>(pts, perm, nq, sorted_pts);
    const bool pack = (MODE & QM_PACKED) && perm && packed;
    k_query_point<false, false, MODE><<<grid, QT, 0, stream>>>(tv, pts, perm, nq, max_dist, result, nullptr, face, u, v, nullptr,
                                                              pack ? packed : nullptr, sorted_pts);
    if (pack)
        k_unpack_results<<<(unsigned)((nq + 255) / 256), 256, 0, stream>>>(packed, nq, result, face, u, v);
}
}  // namespace

const char* wb_query_point(const TreeView& tv, const float* pts, const int* perm, long long nq, float max_dist,
                           int with_sign, uint8_t* result, float* sign, int* face, float* u, float* v,
                           unsigned long long* stats, cudaStream_t stream, int mode, uint4* packed, float* sorted_pts)
{
    if (nq <= 0)
        return nullptr;
    const int grid = query_grid(nq);
    if (!with_sign && !stats && mode) {
        // the staged variant needs the ordered copy of the batch; without an ordering it reads the caller's points in place
        if ((mode & QM_STAGED) && !(perm && sorted_pts))
            mode &= ~QM_STAGED;
        if ((mode & 16) && tv.n >= (1 << 28))
            mode &= ~16;  // packed stack entries hold 28 bits of position
        if ((mode & ~15) == 16 && ((mode & 15) == 6 || (mode & 15) == 2 || (mode & 15) == 0)) {  // 8-byte stack entries (+ I/O bits)
            if ((mode & 15) == 6)
                launch_point_mode<22>(tv, pts, perm, nq, max_dist, result, face, u, v, packed, sorted_pts, stream);
            else if ((mode & 15) == 2)
                launch_point_mode<18>(tv, pts, perm, nq, max_dist, result, face, u, v, packed, sorted_pts, stream);
            else
                launch_point_mode<16>(tv, pts, perm, nq, max_dist, result, face, u, v, packed, sorted_pts, stream);
            cudaError_t e22 = cudaGetLastError();
            return e22 == cudaSuccess ? nullptr : cudaGetErrorString(e22);
        }
        switch (mode & 15) {
#define WB_MODE_CASE(M)                                                                                                       \
    case M:                                                                                                                   \
        launch_point_mode<M>(tv, pts, perm, nq, max_dist, result, face, u, v, packed, sorted_pts, stream);                  \
        break;
            WB_MODE_CASE(1) WB_MODE_CASE(2) WB_MODE_CASE(3) WB_MODE_CASE(4) WB_MODE_CASE(5) WB_MODE_CASE(6) WB_MODE_CASE(7)
            WB_MODE_CASE(8) WB_MODE_CASE(9) WB_MODE_CASE(10) WB_MODE_CASE(11) WB_MODE_CASE(12) WB_MODE_CASE(13)
            WB_MODE_CASE(14) WB_MODE_CASE(15)
#undef WB_MODE_CASE
        default:
            k_query_point<false, false><<<grid, QT, 0, stream>>>(tv, pts, perm, nq, max_dist, result, sign, face, u, v, stats);
        }
        cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
    }
    if (with_sign) {
        const int sgrid = query_grid(nq, QT_SIGN);
        if (stats)
            k_query_point<true, true><<<sgrid, QT_SIGN, 0, stream>>>(tv, pts, perm, nq, max_dist, result, sign, face, u, v, stats);
        else if ((mode & 16) && tv.n < (1 << 28))  // 8-byte stack entries for the closest-point part
            k_query_point<true, false, 16><<<sgrid, QT_SIGN, 0, stream>>>(tv, pts, perm, nq, max_dist, result, sign, face, u, v, stats);
        else
            k_query_point<true, false><<<sgrid, QT_SIGN, 0, stream>>>(tv, pts, perm, nq, max_dist, result, sign, face, u, v, stats);
    } else {
        if (stats)
            k_query_point<false, true><<<grid, QT, 0, stream>>>(tv, pts, perm, nq, max_dist, result, sign, face, u, v, stats);
        else
            k_query_point<false, false><<<grid, QT, 0, stream>>>(tv, pts, perm, nq, max_dist, result, sign, face, u, v, stats);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* wb_query_ray(const TreeView& tv, const float* starts, const float* dirs, const int* perm, const int* roots,
                         long long nq, float max_t, uint8_t* result, float* sign, int* face, float* t, float* u, float* v,
                         float* normal, unsigned long long* stats, cudaStream_t stream)
{
    if (nq <= 0)
        return nullptr;
    const int grid = query_grid(nq);
    // 4-byte packed stack entries for rays: measured SLOWER on C3 (2.96 vs 3.08 G primary rays/s: a primary ray's stack is
    // shallow, the unpack on every pop costs more than the smaller footprint saves) -- off unless WARP_B200_RAY_PACK=1
    static const bool pack_env = [] {
        const char* e = getenv("WARP_B200_RAY_PACK");
        return e && atoi(e) != 0;
    }();
    const bool pack = pack_env && tv.n < (1 << 28);
    if (stats)
        k_query_ray<true, false><<<grid, QT, 0, stream>>>(tv, starts, dirs, perm, roots, nq, max_t, result, sign, face, t, u, v, normal, stats);
    else if (pack)
        k_query_ray<false, true><<<grid, QT, 0, stream>>>(tv, starts, dirs, perm, roots, nq, max_t, result, sign, face, t, u, v, normal, stats);
    else
        k_query_ray<false, false><<<grid, QT, 0, stream>>>(tv, starts, dirs, perm, roots, nq, max_t, result, sign, face, t, u, v, normal, stats);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* wb_query_ray_anyhit(const TreeView& tv, const float* starts, const float* dirs, const int* roots, long long nq,
                                float max_t, uint8_t* result, cudaStream_t stream)
{
    if (nq <= 0)
        return nullptr;
    k_query_ray_aux<false><<<query_grid(nq), QT, 0, stream>>>(tv, starts, dirs, roots, nq, max_t, result, nullptr);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* wb_query_ray_count(const TreeView& tv, const float* starts, const float* dirs, const int* roots, long long nq,
                               int* counts, cudaStream_t stream)
{
    if (nq <= 0)
        return nullptr;
    k_query_ray_aux<true><<<query_grid(nq), QT, 0, stream>>>(tv, starts, dirs, roots, nq, 0.0f, nullptr, counts);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* wb_mesh_eval(const float* attr, const int* indices, const int* face, const float* u, const float* v,
                         long long n, float* out, cudaStream_t stream)
{
    if (n <= 0)
        return nullptr;
    k_mesh_eval<<<query_grid(n), QT, 0, stream>>>(attr, indices, face, u, v, n, out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* wb_sign_parity(const TreeView& tv, const float* pts, long long nq, int n_sample, float scale,
                           const uint8_t* result, float* sign, cudaStream_t stream)
{
    if (nq <= 0)
        return nullptr;
    k_sign_parity<<<query_grid(nq), QT, 0, stream>>>(tv, pts, nq, n_sample, scale, result, sign);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* wb_query_point_sign_normal(const TreeView& tv, const float* mesh_points, const int* mesh_indices,
                                       const float* pts, const int* perm, long long nq, float max_dist, float epsilon,
                                       double* partials, float* avg_edge, uint8_t* result, float* sign, int* face, float* u,
                                       float* v, cudaStream_t stream)
{
    const int nb = tv.n < EDGE_BLOCKS * 256 ? (tv.n + 255) / 256 : EDGE_BLOCKS;
    k_edge_partials<<<nb, 256, 0, stream>>>(mesh_points, mesh_indices, tv.n, partials);
    k_edge_finish<<<1, 1, 0, stream>>>(partials, nb, tv.n, avg_edge);
    if (nq > 0)
        k_query_point_sign_normal<<<query_grid(nq), QT, 0, stream>>>(tv, pts, perm, nq, max_dist, epsilon, avg_edge, result, sign,
                                                                face, u, v);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* wb_query_furthest(const TreeView& tv, const float* pts, long long nq, float min_dist, uint8_t* result, int* face,
                              float* u, float* v, cudaStream_t stream)
{
    if (nq <= 0)
        return nullptr;
    k_query_furthest<<<query_grid(nq), QT, 0, stream>>>(tv, pts, nq, min_dist, result, face, u, v);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* wb_mesh_face_normal(const float* points, const int* indices, const int* face, const uint8_t* mask, long long n,
                                float* out, cudaStream_t stream)
{
    if (n <= 0)
        return nullptr;
    k_mesh_face_normal<<<query_grid(n), QT, 0, stream>>>(points, indices, face, mask, n, out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* wb_mesh_query_sphere(const TreeView& tv, const float* centers, const float* radii, long long nq, int* counts,
                                 const int* offsets, int* indices, cudaStream_t stream)
{
    if (nq <= 0)
        return nullptr;
    if (offsets)
        k_mesh_query_sphere<true><<<query_grid(nq), QT, 0, stream>>>(tv, centers, radii, nq, counts, offsets, indices);
    else
        k_mesh_query_sphere<false><<<query_grid(nq), QT, 0, stream>>>(tv, centers, radii, nq, counts, offsets, indices);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
