// TEST INFRASTRUCTURE ONLY (oracle/_ref).  Thin C-ABI array harness around the UNMODIFIED
// reference sources compiled where they lie under /root/reference/warp/native (bvh.cpp,
// mesh.cpp + the header-only mesh.h / bvh.h / intersect.h).  Nothing in warp_b200/ links,
// loads or calls this library; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs do.
//
// What it exposes:
//   * the reference's own host builders (SAH / median; the reference has no CPU LBVH,
//     bvh.cpp:226-233) through wp_mesh_create_host (mesh.cpp:118-191);
//   * the reference's own query code (mesh.h:128-307, 501-676, 1768-1891, 2286-2359) run over
//     arrays, either on a reference-built tree or on a caller-supplied tree in the reference's
//     node layout (bvh.h:161-207) -- which is how an LBVH built elsewhere gets traversed by the
//     reference's traversal code on the CPU.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include <atomic>
#include <thread>
#include <vector>

#include "warp.h"
#include "mesh.h"

using namespace wp;

// dynamic chunked parallel loop over [0, n) on std::thread (libgomp is not in this image).
// nthreads <= 0 means "all hardware threads"; nthreads == 1 runs inline (the reference's own CPU
// launch is a serial loop, codegen.py:7004-7022).
template <typename F> static void parallel_for(int64_t n, int nthreads, F body)
{
    if (nthreads <= 0) {
        unsigned hw = std::thread::hardware_concurrency();
        nthreads = hw ? (int)hw : 1;
    }
    if (nthreads == 1 || n < 1024) {
        for (int64_t i = 0; i < n; ++i)
            body(i);
        return;
    }
    std::atomic<int64_t> next(0);
    const int64_t chunk = 256;
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
        pool.emplace_back([&]() {
            for (;;) {
                int64_t b = next.fetch_add(chunk);
                if (b >= n)
                    break;
                int64_t e = b + chunk < n ? b + chunk : n;
                for (int64_t i = b; i < e; ++i)
                    body(i);
            }
        });
    for (auto& th : pool)
        th.join();
}

extern "C" {

uint64_t ref_mesh_create(const float* points, int npts, const int* indices, int ntris, int constructor, int leaf_size)
{
    array_t<vec3> p((vec3*)points, npts);
    array_t<vec3> v;
    array_t<int> t((int*)indices, ntris * 3);
    return wp_mesh_create_host(p, v, t, npts, ntris, 0, constructor, nullptr, leaf_size);
}

void ref_mesh_destroy(uint64_t id) { wp_mesh_destroy_host(id); }
void ref_mesh_refit(uint64_t id) { wp_mesh_refit_host(id); }

// tree introspection: sizes, then raw copies of the five arrays
void ref_mesh_tree_info(uint64_t id, int* max_nodes, int* num_nodes, int* num_leaf_nodes, int* root, int* num_items)
{
    Mesh* m = (Mesh*)id;
    *max_nodes = m->bvh.max_nodes;
    *num_nodes = m->bvh.num_nodes;
    *num_leaf_nodes = m->bvh.num_leaf_nodes;
    *root = m->bvh.root ? *m->bvh.root : -1;
    *num_items = m->bvh.num_items;
}

void ref_mesh_tree_copy(uint64_t id, void* node_lowers, void* node_uppers, int* parents, int* primitive_indices)
{
    Mesh* m = (Mesh*)id;
    const BVH& b = m->bvh;
    if (node_lowers)
        memcpy(node_lowers, b.node_lowers, sizeof(BVHPackedNodeHalf) * b.max_nodes);
    if (node_uppers)
        memcpy(node_uppers, b.node_uppers, sizeof(BVHPackedNodeHalf) * b.max_nodes);
    if (parents && b.node_parents)
        memcpy(parents, b.node_parents, sizeof(int) * b.max_nodes);
    if (primitive_indices)
        memcpy(primitive_indices, b.primitive_indices, sizeof(int) * b.num_items);
}

// Wrap caller-owned arrays (reference node layout) into a wp::Mesh the reference query code
// can traverse.  All arrays are borrowed; keep them alive until ref_mesh_from_tree_destroy().
uint64_t ref_mesh_from_tree(
    const float* points,
    int npts,
    const int* indices,
    int ntris,
    const void* node_lowers,
    const void* node_uppers,
    const int* primitive_indices,
    int root,
    int max_nodes,
    int leaf_size
)
{
    Mesh* m = new Mesh();
    m->points = array_t<vec3>((vec3*)points, npts);
    m->indices = array_t<int>((int*)indices, ntris * 3);
    m->num_points = npts;
    m->num_tris = ntris;
    m->bvh.node_lowers = (BVHPackedNodeHalf*)node_lowers;
    m->bvh.node_uppers = (BVHPackedNodeHalf*)node_uppers;
    m->bvh.primitive_indices = (int*)primitive_indices;
    m->bvh.max_nodes = max_nodes;
    m->bvh.num_nodes = max_nodes;
    m->bvh.num_items = ntris;
    m->bvh.num_leaf_nodes = ntris;
    m->bvh.leaf_size = leaf_size;
    m->bvh.constructor_type = BVH_CONSTRUCTOR_LBVH;
    m->bvh.root = new int(root);
    return (uint64_t)m;
}

void ref_mesh_from_tree_destroy(uint64_t id)
{
    Mesh* m = (Mesh*)id;
    delete m->bvh.root;
    delete m;
}

int ref_max_threads()
{
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

// struct-returning overloads (mesh.h:1583-1608, 2250-2257): fields keep their zero init on a miss
void ref_query_point_no_sign(
    uint64_t id, const float* pts, int64_t n, float max_dist, uint8_t* result, int* face, float* u, float* v,
    int nthreads
)
{
    parallel_for(n, nthreads, [&](int64_t i) {
        mesh_query_point_t q = mesh_query_point_no_sign(id, vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), max_dist);
        result[i] = q.result ? 1 : 0;
        face[i] = q.face;
        u[i] = q.u;
        v[i] = q.v;
    });
}

void ref_query_point(
    uint64_t id, const float* pts, int64_t n, float max_dist, uint8_t* result, float* sign, int* face, float* u,
    float* v, int nthreads
)
{
    parallel_for(n, nthreads, [&](int64_t i) {
        mesh_query_point_t q = mesh_query_point(id, vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), max_dist);
        result[i] = q.result ? 1 : 0;
        sign[i] = q.sign;
        face[i] = q.face;
        u[i] = q.u;
        v[i] = q.v;
    });
}

void ref_query_point_sign_parity(
    uint64_t id, const float* pts, int64_t n, float max_dist, int n_sample, float scale, uint8_t* result, float* sign,
    int* face, float* u, float* v, int nthreads
)
{
    parallel_for(n, nthreads, [&](int64_t i) {
        mesh_query_point_t q
            = mesh_query_point_sign_parity(id, vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), max_dist, n_sample, scale);
        result[i] = q.result ? 1 : 0;
        sign[i] = q.sign;
        face[i] = q.face;
        u[i] = q.u;
        v[i] = q.v;
    });
}

// mesh_query_point_sign_normal (mesh.h:860-1090).  The reference's host constructor fills average_edge_length
// (mesh.cpp:140-155); a mesh wrapped around a caller's tree gets it through the setter below.
float ref_mesh_get_average_edge_length(uint64_t id) { return ((Mesh*)id)->average_edge_length; }
void ref_mesh_set_average_edge_length(uint64_t id, float v) { ((Mesh*)id)->average_edge_length = v; }

void ref_query_point_sign_normal(
    uint64_t id, const float* pts, int64_t n, float max_dist, float epsilon, uint8_t* result, float* sign, int* face,
    float* u, float* v, int nthreads
)
{
    parallel_for(n, nthreads, [&](int64_t i) {
        mesh_query_point_t q
            = mesh_query_point_sign_normal(id, vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), max_dist, epsilon);
        result[i] = q.result ? 1 : 0;
        sign[i] = q.sign;
        face[i] = q.face;
        u[i] = q.u;
        v[i] = q.v;
    });
}

void ref_query_furthest_point_no_sign(
    uint64_t id, const float* pts, int64_t n, float min_dist, uint8_t* result, int* face, float* u, float* v, int nthreads
)
{
    parallel_for(n, nthreads, [&](int64_t i) {
        mesh_query_point_t q
            = mesh_query_furthest_point_no_sign(id, vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), min_dist);
        result[i] = q.result ? 1 : 0;
        face[i] = q.face;
        u[i] = q.u;
        v[i] = q.v;
    });
}

void ref_mesh_eval_face_normal(uint64_t id, const int* face, int64_t n, float* out)
{
    for (int64_t i = 0; i < n; ++i) {
        vec3 nn = mesh_eval_face_normal(id, face[i]);
        out[3 * i + 0] = nn[0], out[3 * i + 1] = nn[1], out[3 * i + 2] = nn[2];
    }
}

void ref_query_ray(
    uint64_t id, const float* starts, const float* dirs, int64_t n, float max_t, uint8_t* result, float* sign,
    int* face, float* t, float* u, float* v, float* normal, int nthreads, const int* roots
)
{
    parallel_for(n, nthreads, [&](int64_t i) {
        mesh_query_ray_t q = mesh_query_ray(
            id, vec3(starts[3 * i], starts[3 * i + 1], starts[3 * i + 2]),
            vec3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), max_t, roots ? roots[i] : -1
        );
        result[i] = q.result ? 1 : 0;
        sign[i] = q.sign;
        face[i] = q.face;
        t[i] = q.t;
        u[i] = q.u;
        v[i] = q.v;
        normal[3 * i + 0] = q.normal[0];
        normal[3 * i + 1] = q.normal[1];
        normal[3 * i + 2] = q.normal[2];
    });
}

void ref_query_ray_anyhit(uint64_t id, const float* starts, const float* dirs, int64_t n, float max_t, uint8_t* result,
                          int nthreads, const int* roots)
{
    parallel_for(n, nthreads, [&](int64_t i) {
        result[i] = mesh_query_ray_anyhit(id, vec3(starts[3 * i], starts[3 * i + 1], starts[3 * i + 2]),
                                          vec3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), max_t, roots ? roots[i] : -1)
            ? 1
            : 0;
    });
}

void ref_query_ray_count(uint64_t id, const float* starts, const float* dirs, int64_t n, int* counts, int nthreads,
                         const int* roots)
{
    parallel_for(n, nthreads, [&](int64_t i) {
        counts[i] = mesh_query_ray_count_intersections(id, vec3(starts[3 * i], starts[3 * i + 1], starts[3 * i + 2]),
                                                       vec3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), roots ? roots[i] : -1);
    });
}

void ref_mesh_eval(uint64_t id, int velocity, const int* face, const float* u, const float* v, int64_t n, float* out)
{
    for (int64_t i = 0; i < n; ++i) {
        vec3 r = velocity ? mesh_eval_velocity(id, face[i], u[i], v[i]) : mesh_eval_position(id, face[i], u[i], v[i]);
        out[3 * i] = r[0], out[3 * i + 1] = r[1], out[3 * i + 2] = r[2];
    }
}

// from_tree meshes borrow per-triangle bounds (mesh.lowers / mesh.uppers, read by mesh_query_aabb's item test)
void ref_mesh_set_bounds(uint64_t id, const float* lowers, const float* uppers)
{
    Mesh* m = (Mesh*)id;
    m->lowers = (vec3*)lowers;
    m->uppers = (vec3*)uppers;
}

// mesh_query_aabb + mesh_query_aabb_next loop (mesh.h:2476-2712): offsets[n+1]; indices may be NULL (count only)
void ref_mesh_query_aabb(uint64_t id, const float* lowers, const float* uppers, int64_t n, int* offsets, int* indices)
{
    int run = 0;
    for (int64_t i = 0; i < n; ++i) {
        offsets[i] = run;
        mesh_query_aabb_t q = mesh_query_aabb(id, vec3(lowers[3 * i], lowers[3 * i + 1], lowers[3 * i + 2]),
                                              vec3(uppers[3 * i], uppers[3 * i + 1], uppers[3 * i + 2]));
        int face;
        while (mesh_query_aabb_next(q, face)) {
            if (indices)
                indices[run] = face;
            ++run;
        }
    }
    offsets[n] = run;
}

// mesh_query_sphere + mesh_query_sphere_next loop (mesh.h:2457-2737): offsets[n+1]; indices may be NULL (count only)
void ref_mesh_query_sphere(uint64_t id, const float* centers, const float* radii, int64_t n, int* offsets, int* indices)
{
    int run = 0;
    for (int64_t i = 0; i < n; ++i) {
        offsets[i] = run;
        mesh_query_aabb_t q = mesh_query_sphere(id, vec3(centers[3 * i], centers[3 * i + 1], centers[3 * i + 2]), radii[i]);
        int face;
        while (mesh_query_sphere_next(q, face)) {
            if (indices)
                indices[run] = face;
            ++run;
        }
    }
    offsets[n] = run;
}

// generic wp.Bvh iterator (bvh.h:494-664) and bvh_get_group_root (bvh.h:376-390) over a caller-supplied tree in the
// reference layout.  ray: (qa, qb) = (start, dir), else (lower, upper); roots optional (-1 / NULL = tree root).
static BVH make_bvh(const void* node_lowers, const void* node_uppers, const int* parents, const int* prim,
                    const float* item_lowers, const float* item_uppers, const int* item_groups, int num_items, int* root)
{
    BVH b;
    b.node_lowers = (BVHPackedNodeHalf*)node_lowers;
    b.node_uppers = (BVHPackedNodeHalf*)node_uppers;
    b.node_parents = (int*)parents;
    b.primitive_indices = (int*)prim;
    b.item_lowers = (vec3*)item_lowers;
    b.item_uppers = (vec3*)item_uppers;
    b.item_groups = (int*)item_groups;
    b.num_items = num_items;
    b.num_leaf_nodes = num_items;
    b.max_nodes = b.num_nodes = 2 * num_items - 1;
    b.root = root;
    return b;
}

void ref_bvh_query(const void* node_lowers, const void* node_uppers, const int* prim, int root, const float* item_lowers,
                   const float* item_uppers, int num_items, int ray, const float* qa, const float* qb, const int* roots,
                   int64_t n, float max_dist, int* offsets, int* indices)
{
    BVH b = make_bvh(node_lowers, node_uppers, nullptr, prim, item_lowers, item_uppers, nullptr, num_items, &root);
    const uint64_t id = (uint64_t)&b;
    int run = 0;
    for (int64_t i = 0; i < n; ++i) {
        offsets[i] = run;
        const vec3 a(qa[3 * i], qa[3 * i + 1], qa[3 * i + 2]), c(qb[3 * i], qb[3 * i + 1], qb[3 * i + 2]);
        const int r = roots ? roots[i] : -1;
        bvh_query_t q = ray ? bvh_query_ray(id, a, c, r) : bvh_query_aabb(id, a, c, r);
        int item;
        while (ray ? bvh_query_ray_next(q, item, max_dist) : bvh_query_next(q, item, max_dist)) {
            if (indices)
                indices[run] = item;
            ++run;
        }
    }
    offsets[n] = run;
}

// kind: 0 aabb, 1 ray, 2 sphere (qa = centre), 3 capsule; radii per query for kinds 2 and 3
void ref_bvh_query_kind(const void* node_lowers, const void* node_uppers, const int* prim, int root,
                        const float* item_lowers, const float* item_uppers, int num_items, int kind, const float* qa,
                        const float* qb, const float* radii, const int* roots, int64_t n, float max_dist, int* offsets,
                        int* indices)
{
    BVH b = make_bvh(node_lowers, node_uppers, nullptr, prim, item_lowers, item_uppers, nullptr, num_items, &root);
    const uint64_t id = (uint64_t)&b;
    int run = 0;
    for (int64_t i = 0; i < n; ++i) {
        offsets[i] = run;
        const vec3 a(qa[3 * i], qa[3 * i + 1], qa[3 * i + 2]);
        const float* qbp = kind == 2 ? qa : qb;
        const vec3 c(qbp[3 * i], qbp[3 * i + 1], qbp[3 * i + 2]);
        const int r = roots ? roots[i] : -1;
        const float rad = radii ? radii[i] : 0.0f;
        bvh_query_t q = kind == 1 ? bvh_query_ray(id, a, c, r)
            : kind == 2           ? bvh_query_sphere(id, a, rad, r)
            : kind == 3           ? bvh_query_capsule(id, a, c, rad, r)
                                  : bvh_query_aabb(id, a, c, r);
        int item;
        for (;;) {
            const bool more = kind == 1 ? bvh_query_ray_next(q, item, max_dist)
                : kind == 2             ? bvh_query_sphere_next(q, item, max_dist)
                : kind == 3             ? bvh_query_capsule_next(q, item, max_dist)
                                        : bvh_query_next(q, item, max_dist);
            if (!more)
                break;
            if (indices)
                indices[run] = item;
            ++run;
        }
    }
    offsets[n] = run;
}

void ref_bvh_group_roots(const void* node_lowers, const void* node_uppers, const int* parents, const int* prim,
                         const int* item_groups, int num_items, int root, const int* group_ids, int64_t n, int* roots)
{
    BVH b = make_bvh(node_lowers, node_uppers, parents, prim, nullptr, nullptr, item_groups, num_items, &root);
    for (int64_t i = 0; i < n; ++i)
        roots[i] = bvh_get_group_root((uint64_t)&b, group_ids[i]);
}

// standalone primitives for unit-level pinning of the restatement
void ref_closest_point_to_triangle(const float* a, const float* b, const float* c, const float* p, float* uv)
{
    vec2 r = closest_point_to_triangle(vec3(a[0], a[1], a[2]), vec3(b[0], b[1], b[2]), vec3(c[0], c[1], c[2]),
                                       vec3(p[0], p[1], p[2]));
    uv[0] = r[0];
    uv[1] = r[1];
}

uint32_t ref_morton3_1024(float x, float y, float z) { return morton3<1024>(x, y, z); }

}  // extern "C"
