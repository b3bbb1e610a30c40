// batched query drivers (query.cu); pointers are device pointers, work is enqueued on `stream`
#pragma once
#include "common.cuh"

// generic BVH queries (bvh_query.cu).  kind 0: (qa, qb) = (lower, upper); 1 ray: (start, dir), half-open max_dist;
// 2 sphere: qa = centre, radii[nq]; 3 capsule: (start, dir), radii[nq], closed max_dist (bvh.h:462-492).  offsets == NULL
// counts hits into counts[nq]; otherwise writes the hit items of query i at indices[offsets[i]...]
// roots: optional per-query start node (reference node index, e.g. from wb_group_roots; < 0 = the tree root)
const char* wb_bvh_query(const TreeView& tv, const float* item_lowers, const float* item_uppers, int kind,
                         const float* qa, const float* qb, const float* radii, const int* roots, long long nq,
                         float max_dist, int* counts, const int* offsets, int* indices, cudaStream_t stream);
// bvh_get_group_root for a batch; keys = the tree's 64-bit (group << 32 | code) keys, NULL for an ungrouped tree
const char* wb_group_roots(const TreeView& tv, const void* keys, const int* group_ids, long long nq, int* roots,
                           cudaStream_t stream);
// offsets[0..n] = exclusive prefix sums of counts[0..n); scratch = ceil(n / 2048) + 1 int64 words
const char* wb_exclusive_scan(const int* counts, int* offsets, long long n, long long* scratch, cudaStream_t stream);

// perm: optional permutation (thread slot -> query index), e.g. from wb_morton_order
// mode: memory-side variant bits of the unsigned kernel (query.cu QM_*: 1 16-byte stack entries, 2 streaming I/O hints,
// 4 packed 16-byte result records through `packed` (nq entries) + unpack pass, 8 curve-ordered copy of the batch in
// `sorted_pts` (12 * nq + 16 bytes) staged per block with one TMA bulk copy); 0 = the plain kernel
const char* wb_query_point(const TreeView& tv, const float* pts, const int* perm, long long nq, float max_dist,
                           int with_sign, uint8_t* result, float* sign, int* face, float* u, float* v,
                           unsigned long long* stats, cudaStream_t stream, int mode = 0, uint4* packed = nullptr,
                           float* sorted_pts = nullptr);
// roots: optional per-ray start node (reference node index; < 0 or NULL = the tree root)
const char* wb_query_ray(const TreeView& tv, const float* starts, const float* dirs, const int* perm, const int* roots,
                         long long nq, float max_t, uint8_t* result, float* sign, int* face, float* t, float* u, float* v,
                         float* normal, unsigned long long* stats, cudaStream_t stream);
// mesh_query_ray_anyhit / mesh_query_ray_count_intersections / mesh_eval_{position,velocity} (mesh.h:1893-2032, 2767-2807)
const char* wb_query_ray_anyhit(const TreeView& tv, const float* starts, const float* dirs, const int* roots, long long nq,
                                float max_t, uint8_t* result, cudaStream_t stream);
const char* wb_query_ray_count(const TreeView& tv, const float* starts, const float* dirs, const int* roots, long long nq,
                               int* counts, cudaStream_t stream);
const char* wb_mesh_eval(const float* attr, const int* indices, const int* face, const float* u, const float* v,
                         long long n, float* out, cudaStream_t stream);
// mesh_query_point_sign_normal (mesh.h:860-1090); partials: >= 296 doubles of scratch, avg_edge: device float that
// receives Mesh.average_edge_length (recomputed from mesh_points / mesh_indices first).  nq == 0 only refreshes it.
const char* wb_query_point_sign_normal(const TreeView& tv, const float* mesh_points, const int* mesh_indices,
                                       const float* pts, const int* perm, long long nq, float max_dist, float epsilon,
                                       double* partials, float* avg_edge, uint8_t* result, float* sign, int* face, float* u,
                                       float* v, cudaStream_t stream);
// sign of mesh_query_point_sign_parity (mesh.h:2362-2392) for the queries whose `result` is set; 0 elsewhere
const char* wb_sign_parity(const TreeView& tv, const float* pts, long long nq, int n_sample, float scale,
                           const uint8_t* result, float* sign, cudaStream_t stream);
// mesh_query_furthest_point_no_sign (mesh.h:678-858) and mesh_eval_face_normal (mesh.h:2870-2888)
const char* wb_query_furthest(const TreeView& tv, const float* pts, long long nq, float min_dist, uint8_t* result, int* face,
                              float* u, float* v, cudaStream_t stream);
// mask (optional): entries with mask[i] == 0 get the zero vector (the normal of a ray that missed)
const char* wb_mesh_face_normal(const float* points, const int* indices, const int* face, const uint8_t* mask, long long n,
                                float* out, cudaStream_t stream);
// mesh_query_sphere hit lists (mesh.h:2457-2737): offsets == NULL counts into counts[nq], else fills indices
const char* wb_mesh_query_sphere(const TreeView& tv, const float* centers, const float* radii, long long nq, int* counts,
                                 const int* offsets, int* indices, cudaStream_t stream);
