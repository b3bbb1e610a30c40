/*
 * warp_b200.h -- C ABI of libwarp_b200.so, the B200-native (sm_100a) replacement for the
 * mesh / BVH spatial-query path of NVIDIA/warp's native library (warp.so).
 *
 * Part 1 re-exports, with identical names, argument order and meaning, the entry points that the
 * reference's Python layer binds for this path (warp/_src/context.py:7000-7070 binding
 * warp/native/warp.h:93-140).  A maintainer can point `warp._src.context.runtime.core` at this
 * library for these symbols and `wp.Mesh(...)` / `wp.Bvh(...)` / `.refit()` / `.rebuild()` keep
 * working (INTEGRATION.md shows the stub).
 *
 * Part 2 is the small slice of the runtime API (allocation, copies, streams, events) this path
 * needs, again under the reference's names (warp/native/warp.h:38-84, 682-764).
 *
 * Part 3 is new: the reference has no C entry point for queries (they are header code inlined
 * into NVRTC-compiled user kernels, warp/native/mesh.h); here they are batched calls returning
 * the fields of MeshQueryPoint / MeshQueryRay (warp/_src/types.py:7786-7848) as SoA arrays.
 *
 * Conventions (same as the reference): an object id is the address of a device-resident
 * descriptor (layout below, identical to wp::BVH / wp::Mesh so kernels that dereference the id keep
 * working once wp_b200_bvh_sync_reference_layout() has been called); id 0 means failure and
 * wp_get_error_string() says why; points / indices / lowers / uppers / groups are BORROWED -- the
 * caller keeps them alive and may modify them in place before refit(); all work is enqueued on the
 * current stream of the calling thread's device (wp_cuda_context_set_stream) and is asynchronous.
 * `context` arguments accept NULL (= current device) or a handle from
 * wp_cuda_device_get_primary_context().  No entry point falls back to the CPU.
 */
#ifndef WARP_B200_H
#define WARP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WP_B200_API __attribute__((visibility("default")))

/* wp::vec3 -- three packed floats (warp/native/vec.h) */
typedef struct {
    float c[3];
} wp_vec3;

/* wp::array_t<T>, passed BY VALUE (warp/native/array.h:173-277, warp/_src/types.py:2417-2425) */
typedef struct {
    uint64_t data;
    uint64_t grad;
    int32_t shape[4];
    int32_t strides[4];
    uint16_t ndim;
    uint16_t flags;
} wp_array_t;

/* BVH_CONSTRUCTOR_* (warp/native/bvh.h:21-24) */
#define WP_BVH_CONSTRUCTOR_SAH 0
#define WP_BVH_CONSTRUCTOR_MEDIAN 1
#define WP_BVH_CONSTRUCTOR_LBVH 2
#define WP_BVH_CONSTRUCTOR_CUBQL (-1)

/* device-resident descriptors at `id`; field-for-field wp::BVH (bvh.h:176-207, 112 bytes) and
 * wp::Mesh (mesh.h:18-35, 328 bytes) */
typedef struct {
    void* node_lowers;  /* BVHPackedNodeHalf[max_nodes]; valid after wp_b200_bvh_sync_reference_layout */
    void* node_uppers;
    int* node_parents;
    int* node_counts;
    int* primitive_indices;
    int max_depth, max_nodes, num_nodes, num_leaf_nodes;
    int* root;
    wp_vec3* item_lowers;
    wp_vec3* item_uppers;
    int* item_groups;
    int num_items, leaf_size, constructor_type;
    void* context;
} wp_b200_bvh_desc;

typedef struct {
    wp_array_t points, velocities, indices;
    wp_vec3* lowers;
    wp_vec3* uppers;
    void* solid_angle_props;
    int num_points, num_tris;
    wp_b200_bvh_desc bvh;
    void* context;
    float average_edge_length;
} wp_b200_mesh_desc;

/* ---------------------------------------------------------------------------------------------
 * Part 1 -- drop-in entry points (same symbols as warp.so)
 * ------------------------------------------------------------------------------------------- */

/* replaces warp/native/bvh.cu:852-875 (decl. warp.h:99-101).  Only constructor_type LBVH builds on
 * the GPU; SAH / MEDIAN / CUBQL return 0 with an error string (this library has no CPU builder).
 * `groups` (optional, one int per item) builds a grouped tree: no subtree mixes groups until a group is complete. */
WP_B200_API uint64_t wp_bvh_create_device(void* context, wp_vec3* lowers, wp_vec3* uppers, int num_items,
                                          int constructor_type, int* groups, int leaf_size);
/* replaces bvh.cu:878-888 (warp.h:102) */
WP_B200_API void wp_bvh_destroy_device(uint64_t id);
/* replaces bvh.cu:809-817 (warp.h:103) */
WP_B200_API void wp_bvh_refit_device(uint64_t id);
/* replaces bvh.cu:819-843 (warp.h:104): in-place LBVH rebuild, no allocation, capture safe */
WP_B200_API void wp_bvh_rebuild_device(uint64_t id);

/* replaces warp/native/mesh.cu:258-345 (warp.h:123-134) */
WP_B200_API uint64_t wp_mesh_create_device(void* context, wp_array_t points, wp_array_t velocities, wp_array_t tris,
                                           int num_points, int num_tris, int support_winding_number,
                                           int constructor_type, int* groups, int bvh_leaf_size);
/* replaces mesh.cu:347-364 (warp.h:135) */
WP_B200_API void wp_mesh_destroy_device(uint64_t id);
/* replaces mesh.cu:368-407 (warp.h:136): 1 ok, 0 error */
WP_B200_API int wp_mesh_refit_device(uint64_t id);
/* replaces mesh.cu:409-433 (warp.h:139): shape must match; triggers a refit */
WP_B200_API int wp_mesh_set_points_device(uint64_t id, wp_array_t points);
/* replaces mesh.cu:435-456 (warp.h:142) */
WP_B200_API void wp_mesh_set_velocities_device(uint64_t id, wp_array_t velocities);
/* replaces warp/native/error.cpp:8-27 (warp.h:38) */
WP_B200_API const char* wp_get_error_string(void);

/* ---------------------------------------------------------------------------------------------
 * Part 2 -- runtime slice (names of warp.h:60-84, 682-764)
 * ------------------------------------------------------------------------------------------- */
WP_B200_API int wp_init(const char* expected_version);
WP_B200_API int wp_is_cuda_enabled(void);
WP_B200_API int wp_cuda_device_get_count(void);
WP_B200_API void* wp_cuda_device_get_primary_context(int ordinal);
WP_B200_API void* wp_cuda_context_get_current(void);
WP_B200_API void wp_cuda_context_set_current(void* context);
WP_B200_API void wp_cuda_context_synchronize(void* context);
WP_B200_API void* wp_cuda_context_get_stream(void* context);
WP_B200_API void wp_cuda_context_set_stream(void* context, void* stream, int sync);
WP_B200_API void* wp_cuda_stream_create(void* context, int priority);
WP_B200_API void wp_cuda_stream_destroy(void* context, void* stream);
WP_B200_API void wp_cuda_stream_synchronize(void* stream);
WP_B200_API void* wp_cuda_event_create(void* context, unsigned flags); /* flags: 1 = disable timing */
WP_B200_API void wp_cuda_event_destroy(void* event);
WP_B200_API void wp_cuda_event_record(void* event, void* stream, int external);
WP_B200_API void wp_cuda_event_synchronize(void* event);
WP_B200_API float wp_cuda_event_elapsed_time(void* start_event, void* end_event);
/* CUDA graph capture of the stream work of this path (warp.h:766-771; 1 ok / 0 error).  `stream` must be a created
 * stream; mode = cudaStreamCaptureMode (0 global, 1 thread local, 2 relaxed); external != 0: capture already active */
WP_B200_API int wp_cuda_graph_begin_capture(void* context, void* stream, int external, int mode);
WP_B200_API int wp_cuda_graph_end_capture(void* context, void* stream, void** graph_ret);
WP_B200_API int wp_cuda_graph_create_exec(void* context, void* stream, void* graph, void** graph_exec_ret);
WP_B200_API int wp_cuda_graph_launch(void* graph_exec, void* stream);
WP_B200_API int wp_cuda_graph_destroy(void* context, void* graph);
WP_B200_API int wp_cuda_graph_exec_destroy(void* context, void* graph_exec);
WP_B200_API void* wp_alloc_device(void* context, size_t s, const char* tag);
WP_B200_API void wp_free_device(void* context, void* ptr);
WP_B200_API void* wp_alloc_pinned(size_t s, const char* tag);
WP_B200_API void wp_free_pinned(void* ptr);
WP_B200_API int wp_memcpy_h2d(void* context, void* dest, void* src, size_t n, void* stream);
WP_B200_API int wp_memcpy_d2h(void* context, void* dest, void* src, size_t n, void* stream);
WP_B200_API int wp_memcpy_d2d(void* context, void* dest, void* src, size_t n, void* stream);
WP_B200_API int wp_memset_device(void* context, void* dest, int value, size_t n, void* stream);
/* device-wide query used by bench.py (L2 size, SM count, clocks, free memory) */
WP_B200_API int wp_b200_device_attr(int ordinal, const char* name, long long* value);
WP_B200_API int wp_b200_pointer_device(const void* ptr, long long* ordinal); /* which device owns a device pointer */
WP_B200_API int wp_b200_device_name(int ordinal, char* buf, int len);

/* ---------------------------------------------------------------------------------------------
 * Part 3 -- batched queries (new; semantics of warp/native/mesh.h)
 * All pointers are DEVICE pointers unless the name ends in _host.  Points / starts / dirs are n x 3
 * packed floats.  Outputs follow the struct-returning overloads (mesh.h:1514-1540, 1583-1608,
 * 2216-2257): on a miss result = 0 and every other field is 0.  Return 1 ok, 0 error.
 * ------------------------------------------------------------------------------------------- */

/* wp.mesh_query_point_no_sign (mesh.h:501-676) */
WP_B200_API int wp_b200_mesh_query_point_no_sign(uint64_t id, const float* points, int64_t n, float max_dist,
                                                 uint8_t* result, int32_t* face, float* u, float* v);
/* wp.mesh_query_point (mesh.h:128-307 + 2286-2359): sign = -1 inside / +1 outside */
WP_B200_API int wp_b200_mesh_query_point(uint64_t id, const float* points, int64_t n, float max_dist, uint8_t* result,
                                         float* sign, int32_t* face, float* u, float* v);
/* wp.mesh_query_point_sign_parity (mesh.h:309-498, 2362-2392): closest point as above; sign = -1 when at least half of
 * n_sample rays along (1,1,1) + U(-perturbation_scale, perturbation_scale)^3 (deterministic PCG stream, seed 42, the
 * three offsets drawn x, y, z -- the order of the reference's device builds) cross an odd number of faces, else +1;
 * the reference's defaults are n_sample = 1, perturbation_scale = 0.1 */
WP_B200_API int wp_b200_mesh_query_point_sign_parity(uint64_t id, const float* points, int64_t n, float max_dist,
                                                     int n_sample, float perturbation_scale, uint8_t* result, float* sign,
                                                     int32_t* face, float* u, float* v);
/* wp.mesh_query_point_sign_normal (mesh.h:860-1090): closest point on distances with a welding band of
 * average_edge_length * epsilon (reference default epsilon = 1e-3); sign = +1 when the angle-weighted normal
 * accumulated over the faces inside the band points towards the query, else -1.  The mesh's average edge length
 * (mesh.cu:38-60, 299-307) is recomputed from the current points by every call and stored in the descriptor field
 * wp::Mesh::average_edge_length. */
WP_B200_API int wp_b200_mesh_query_point_sign_normal(uint64_t id, const float* points, int64_t n, float max_dist,
                                                     float epsilon, uint8_t* result, float* sign, int32_t* face, float* u,
                                                     float* v);
/* Mesh.average_edge_length of the current points (synchronises the stream); returns 0 on an invalid id */
WP_B200_API int wp_b200_mesh_average_edge_length(uint64_t id, float* out);
/* wp.mesh_query_ray (mesh.h:1768-1891): normal is n x 3.  `roots` (optional, NULL = whole tree) restricts ray i to the
 * subtree of that node, as the reference's `root` argument does (e.g. a group root from wp_b200_bvh_get_group_root) */
WP_B200_API int wp_b200_mesh_query_ray(uint64_t id, const float* starts, const float* dirs, int64_t n, float max_t,
                                       uint8_t* result, float* sign, int32_t* face, float* t, float* u, float* v,
                                       float* normal, const int32_t* roots);
/* wp.mesh_query_ray_anyhit (mesh.h:1893-1974): result[i] = 1 when some triangle is hit with 0 <= t < max_t */
WP_B200_API int wp_b200_mesh_query_ray_anyhit(uint64_t id, const float* starts, const float* dirs, int64_t n,
                                              float max_t, uint8_t* result, const int32_t* roots);
/* wp.mesh_query_ray_count_intersections (mesh.h:1976-2032): number of triangles hit with t >= 0, unbounded ray */
WP_B200_API int wp_b200_mesh_query_ray_count_intersections(uint64_t id, const float* starts, const float* dirs,
                                                           int64_t n, int32_t* counts, const int32_t* roots);
/* wp.mesh_eval_position / wp.mesh_eval_velocity (mesh.h:2767-2805): out[i] = p*u + q*v + r*(1-u-v) of triangle
 * face[i], read from the mesh's CURRENT point / velocity array (zeros when the mesh has none); out is n x 3 */
WP_B200_API int wp_b200_mesh_eval_position(uint64_t id, const int32_t* face, const float* u, const float* v,
                                           int64_t n, float* out);
WP_B200_API int wp_b200_mesh_eval_velocity(uint64_t id, const int32_t* face, const float* u, const float* v,
                                           int64_t n, float* out);

/* wp.mesh_eval_face_normal (mesh.h:2870-2888): out[i] = normalize(cross(q - p, r - p)) of triangle face[i], from the
 * mesh's current points; out is n x 3 */
WP_B200_API int wp_b200_mesh_eval_face_normal(uint64_t id, const int32_t* face, int64_t n, float* out);
/* same with a per-entry mask: mask[i] == 0 gives the zero vector -- the `normal` field of mesh_query_ray recomputed from
 * (result, face) alone, bit for bit (used by the sharded ray query: 12 bytes per ray that need not cross NVLink) */
WP_B200_API int wp_b200_mesh_eval_face_normal_masked(uint64_t id, const int32_t* face, const uint8_t* mask, int64_t n,
                                                     float* out);
/* wp.mesh_query_furthest_point_no_sign (mesh.h:678-858): the farthest point of the mesh from each query (always a
 * vertex: (u, v) is (1,0), (0,1) or (0,0)), result = 1 when it lies strictly beyond min_dist */
WP_B200_API int wp_b200_mesh_query_furthest_point_no_sign(uint64_t id, const float* points, int64_t n, float min_dist,
                                                          uint8_t* result, int32_t* face, float* u, float* v);

/* the first three calls with HOST buffers: inputs are copied to the device, the query runs, results are
 * copied back, and the call returns after the results are valid (synchronous).  Work is chunked and
 * double-buffered through pinned staging so copies overlap traversal. */
WP_B200_API int wp_b200_mesh_query_point_no_sign_host(uint64_t id, const float* points, int64_t n, float max_dist,
                                                      uint8_t* result, int32_t* face, float* u, float* v);
WP_B200_API int wp_b200_mesh_query_point_host(uint64_t id, const float* points, int64_t n, float max_dist,
                                              uint8_t* result, float* sign, int32_t* face, float* u, float* v);
WP_B200_API int wp_b200_mesh_query_ray_host(uint64_t id, const float* starts, const float* dirs, int64_t n,
                                            float max_t, uint8_t* result, float* sign, int32_t* face, float* t,
                                            float* u, float* v, float* normal);

/* Morton resolution of trees created AFTER the call: 30 (default; the reference's 1024^3 grid, the only
 * mode with bit-exact parity) or 63 (2 097 152^3 grid, 64-bit keys: a tree-quality option for meshes with
 * far more than 2^20 occupied cells, where the 30-bit code degenerates into long runs of equal keys).
 * Grouped trees always use 64-bit keys (group << 32 | 30-bit code, bvh.cu:205-209). */
WP_B200_API void wp_b200_set_morton_bits(int bits);
WP_B200_API int wp_b200_get_morton_bits(void);

/* thread-to-query assignment of the point queries: 0 = input order, 1 = Morton order of the batch
 * (sorted on the device with the builder's radix sort; answers are unaffected), 2 = auto (default):
 * Morton order for batches of >= 32768 points. */
WP_B200_API void wp_b200_set_query_order(int mode);
WP_B200_API int wp_b200_get_query_order(void);
/* thread-to-ray assignment of wp_b200_mesh_query_ray: 0 = input order (default; primary rays are coherent as
 * given), 1 = sort the batch by origin cell, then direction cell, before tracing (incoherent batches). */
WP_B200_API void wp_b200_set_ray_order(int mode);
WP_B200_API int wp_b200_get_ray_order(void);

/* traversal counters of the NEXT query call on this thread: when `enable` is non-zero the next
 * query also counts 64-byte sibling-pair fetches and 48-byte triangle fetches (slower, used for the
 * bytes-fetched / nodes-per-second report); read them back with wp_b200_query_stats_read. */
WP_B200_API void wp_b200_query_stats_enable(int enable);
/* live timing of the point / ray traversal kernel alone (CUDA events on the launch stream around the kernel, not the
 * batch ordering before it): enable, run, read the summed launch durations and the launch count since the last read */
WP_B200_API void wp_b200_kernel_timing_enable(int enable);
WP_B200_API void wp_b200_kernel_timing_read(float* total_ms, int* launches);
WP_B200_API void wp_b200_query_stats_read(unsigned long long* pair_fetches, unsigned long long* tri_fetches);

/* wp.mesh_query_aabb (mesh.h:2476-2712): faces whose AABB (as of the last build / refit) overlaps the query box, in
 * the reference iterator's order; same count -> scan -> fill protocol as the wp.Bvh queries above */
WP_B200_API int wp_b200_mesh_query_aabb_count(uint64_t id, const float* lowers, const float* uppers, int64_t n,
                                              int32_t* counts);
WP_B200_API int wp_b200_mesh_query_aabb_fill(uint64_t id, const float* lowers, const float* uppers, int64_t n,
                                             const int32_t* offsets, int32_t* indices);
/* ---------------------------------------------------------------------------------------------
 * Generic wp.Bvh queries (batched forms of wp.bvh_query_aabb / wp.bvh_query_ray + bvh_query_next,
 * warp/native/bvh.h:494-600): every item whose AABB overlaps the query box / is entered by the query ray
 * before max_dist.  Hits are produced in the order the reference iterator yields them, in CSR form:
 *   1. *_count  -> counts[n]          2. wp_b200_exclusive_scan_i32(counts, offsets, n) -> offsets[n+1]
 *   3. *_fill   -> indices[offsets[n]]        (all device pointers, current stream; 1 ok / 0 error)
 * ------------------------------------------------------------------------------------------- */
/* `roots` (optional, NULL = whole tree): per-query start node, a reference node index such as the ones
 * wp_b200_bvh_get_group_root returns (bvh.h:504: root == -1 means the tree root) */
WP_B200_API int wp_b200_bvh_query_aabb_count(uint64_t id, const float* lowers, const float* uppers, const int32_t* roots,
                                             int64_t n, int32_t* counts);
WP_B200_API int wp_b200_bvh_query_aabb_fill(uint64_t id, const float* lowers, const float* uppers, const int32_t* roots,
                                            int64_t n, const int32_t* offsets, int32_t* indices);
WP_B200_API int wp_b200_bvh_query_ray_count(uint64_t id, const float* starts, const float* dirs, const int32_t* roots,
                                            int64_t n, float max_dist, int32_t* counts);
WP_B200_API int wp_b200_bvh_query_ray_fill(uint64_t id, const float* starts, const float* dirs, const int32_t* roots,
                                           int64_t n, float max_dist, const int32_t* offsets, int32_t* indices);
/* wp.bvh_query_sphere (bvh.h:542-551, node test intersect.h:197-205): items whose AABB lies within radii[i] of
 * centers[i]; wp.bvh_query_capsule (bvh.h:529-540, node test bvh.h:472-482): items whose AABB, inflated by radii[i],
 * is entered by the ray starts[i] + t * dirs[i] at t <= max_dist (closed).  Negative radii count as 0.  Same count /
 * scan / fill protocol and hit order as the AABB and ray lists. */
WP_B200_API int wp_b200_bvh_query_sphere_count(uint64_t id, const float* centers, const float* radii, const int32_t* roots,
                                               int64_t n, int32_t* counts);
WP_B200_API int wp_b200_bvh_query_sphere_fill(uint64_t id, const float* centers, const float* radii, const int32_t* roots,
                                              int64_t n, const int32_t* offsets, int32_t* indices);
WP_B200_API int wp_b200_bvh_query_capsule_count(uint64_t id, const float* starts, const float* dirs, const float* radii,
                                                const int32_t* roots, int64_t n, float max_dist, int32_t* counts);
WP_B200_API int wp_b200_bvh_query_capsule_fill(uint64_t id, const float* starts, const float* dirs, const float* radii,
                                               const int32_t* roots, int64_t n, float max_dist, const int32_t* offsets,
                                               int32_t* indices);
/* wp.mesh_query_sphere + the mesh_query_sphere_next loop (mesh.h:2457-2737): the faces that intersect the sphere
 * (centers[i], radii[i]) -- exact sphere / box test on nodes and face boxes, then the closest point of the triangle
 * within the radius -- in the iterator's order; faces are taken as of the last build / refit */
WP_B200_API int wp_b200_mesh_query_sphere_count(uint64_t id, const float* centers, const float* radii, int64_t n,
                                                int32_t* counts);
WP_B200_API int wp_b200_mesh_query_sphere_fill(uint64_t id, const float* centers, const float* radii, int64_t n,
                                               const int32_t* offsets, int32_t* indices);
/* wp.bvh_get_group_root (bvh.h:376-390) for a batch of group ids: roots[i] = reference index of the node that holds
 * exactly the items of group_ids[i] (a leaf for a one-item group), -1 when the group does not occur.  On a tree
 * built without groups every item is in group 0. */
WP_B200_API int wp_b200_bvh_get_group_root(uint64_t id, const int32_t* group_ids, int64_t n, int32_t* roots);
WP_B200_API int wp_b200_exclusive_scan_i32(const int32_t* counts, int32_t* offsets, int64_t n);

/* how refit() walks the tree: 0 = auto (default: wavefront from 2^21 items up), 1 = atomic arrival counters
 * (the reference's scheme, bvh.cu:42-144), 2 = wavefront: a visiting order planned once per build, levels inside
 * blocks of 1024 sorted positions in shared memory, counters only above the blocks.  Same boxes bit for bit. */
WP_B200_API void wp_b200_set_refit_mode(int mode);
WP_B200_API int wp_b200_get_refit_mode(void);

/* Creators with the one build-time extension argument: morton_bits = 30 (the reference's 1024^3 code, bit-exact
 * parity, bvh.cu:184-214), 63 (21 bits per axis, a quality option for meshes with far more than 2^20 occupied cells;
 * not combinable with groups) or 0 = the process default (wp_b200_set_morton_bits).  The reference-named creators
 * above are these with morton_bits = 0.  Everything else as wp_bvh_create_device / wp_mesh_create_device. */
WP_B200_API uint64_t wp_b200_bvh_create_device_ex(void* context, wp_vec3* lowers, wp_vec3* uppers, int num_items,
                                                  int constructor_type, int* groups, int leaf_size, int morton_bits);
WP_B200_API uint64_t wp_b200_mesh_create_device_ex(void* context, wp_array_t points, wp_array_t velocities, wp_array_t tris,
                                                   int num_points, int num_tris, int support_winding_number,
                                                   int constructor_type, int* groups, int bvh_leaf_size, int morton_bits);

/* Per-object options (thread-safe alternative to the process-wide setters, which only provide defaults):
 * "refit_mode" 0 / 1 / 2, "query_order" 0 input / 1 curve / 2 auto, "ray_order" 0 / 1, "auto_reference_layout" 0 / 1;
 * -1 = follow the process default.  "morton_bits" can be read, not set.  1 ok / 0 error. */
WP_B200_API int wp_b200_bvh_set_option(uint64_t id, const char* name, int value);
WP_B200_API int wp_b200_bvh_get_option(uint64_t id, const char* name, int* value);

/* Drop-in use under the reference's Python layer: when enabled, every tree created
 * afterwards keeps the reference-layout mirror of its descriptor (node_lowers / node_uppers / node_parents / root,
 * Mesh::lowers / uppers, average_edge_length) current after create / refit / rebuild / set_points, so unmodified Warp
 * kernels can traverse through `id`.  Off by default (the mirror costs one extra pass per refit). */
/* host-only (needs no device): item order (primitive_indices) and leaf starts of the sah (0) / median (1) tree the library
 * builds for these boxes (csrc/host_build.cu, restating warp/native/bvh.cpp:216-572); returns the tree depth, -1 on bad
 * arguments.  For CPU-side parity checks against the reference's host builder. */
WP_B200_API int wp_b200_host_build_order(const float* lowers, const float* uppers, int n, int leaf_size, int constructor_type,
                                         int* order_out, unsigned char* leaf_start_out);

/* switches of measured-and-rejected alternatives kept as tested code paths: "small_nodes" = 1 / 0 / -1 (environment
 * default) -- the builder's Karras-style pass over small distinct-key nodes.  1 ok / 0 unknown */
WP_B200_API int wp_b200_set_experiment(const char* name, int value);
WP_B200_API void wp_b200_set_auto_reference_layout(int enable);
WP_B200_API int wp_b200_get_auto_reference_layout(void);

/* in-place LBVH rebuild of a mesh's tree from the current vertices (the reference only offers this
 * for wp.Bvh, bvh.cu:819-843; here a Mesh gets it too): no allocation, same buffers. 1 ok / 0 error */
WP_B200_API int wp_b200_mesh_rebuild_device(uint64_t id);

/* ---- synthetic-workload generators (bench.py / tests).  Device-side, counter-based RNG (splitmix64 of
 * the GLOBAL index, so a shard [first_index, first_index + n) of a larger batch is reproducible on any rank), on the
 * calling thread's current device and its current stream, capture-safe.  1 ok / 0 error.
 * box queries: out[n] vec3 uniform in [lower, upper] (host pointers to 3 floats each).
 * cloth: vertices of an n_side^2 grid on [0,1]^2 at frame (*frame_dev + frame_offset), z = 0.05 sin(12x + 0.05f)
 * cos(9y + 0.03f); queries of frame f = *frame_dev: random cloth vertices at frame f - 1 + N(0, sigma) per axis. */
WP_B200_API int wp_b200_gen_box_queries(float* out, int64_t n, int64_t first_index, uint64_t seed, const float* lower,
                                        const float* upper);
WP_B200_API int wp_b200_gen_cloth_points(float* points, int n_side, const int* frame_dev, int frame_offset);
WP_B200_API int wp_b200_gen_cloth_queries(float* out, int64_t nq, int n_side, const int* frame_dev, float sigma);
WP_B200_API int wp_b200_counter_add(int* counter_dev, int value);

/* introspection for parity checks / drop-in kernels */
typedef struct {
    int num_items, leaf_size, max_nodes, root, height, deep, key_bits;
    float total_lower[3], total_upper[3], inv_edges[3];
} wp_b200_bvh_info_t;
WP_B200_API int wp_b200_bvh_info(uint64_t id, wp_b200_bvh_info_t* info); /* synchronises the stream */
/* (re)materialise node_lowers / node_uppers / node_parents / root of the descriptor at `id` in the
 * reference's layout (bvh.h:161-207) from the native pair layout.  Works for Bvh and Mesh ids. */
WP_B200_API int wp_b200_bvh_sync_reference_layout(uint64_t id);
/* host copies of the tree products (any pointer may be NULL): sorted keys [n] (key_bits / 8 bytes each),
 * primitive_indices [n], node_lowers / node_uppers [2n-1] x 16 bytes, node_parents [2n-1], root [1].
 * Calls wp_b200_bvh_sync_reference_layout first and synchronises. */
/* EXPERIMENT, not part of the drop-in surface: parents of the n - 1 internal nodes recomputed from
 * the sorted keys alone by a dependency-free kernel (the reference LBVH is the Cartesian tree of the key-delta array,
 * bvh.cu:218-226, 300-334); parents_out is a device array of n - 1 int32 (reference node indices, -1 = root).  Returns
 * the average kernel time in microseconds over `reps` launches, -1 on error, -2 when a run of equal keys exceeds what
 * the prototype replays */
WP_B200_API float wp_b200_experiment_parallel_topology(uint64_t id, int32_t* parents_out, int reps);
WP_B200_API int wp_b200_bvh_download(uint64_t id, void* keys, int32_t* primitive_indices, void* node_lowers,
                                     void* node_uppers, int32_t* node_parents, int32_t* root);

/* ---------------------------------------------------------------------------------------------
 * multi-GPU: query batches are sharded by the host, results gathered with NCCL over NVLink.
 * One communicator per process (one process per GPU).  libnccl.so.2 is dlopen()ed on first use.
 * ------------------------------------------------------------------------------------------- */
WP_B200_API int wp_b200_nccl_load(const char* libnccl_path);       /* NULL = default search */
WP_B200_API int wp_b200_nccl_unique_id(void* id128);               /* rank 0: fills 128 bytes */
WP_B200_API int wp_b200_nccl_init(const void* id128, int world_size, int rank);
WP_B200_API int wp_b200_nccl_allgather(const void* send, void* recv, size_t bytes_per_rank); /* device ptrs */
/* pipelined gather of one part of every shard into the rank-major result, on the library's communication stream:
 * recv + r * shard_stride_bytes + offset_bytes <- rank r's `send` (part_bytes); fork = the communication stream waits
 * for the current stream (the part's query), join = the current stream waits for the communication stream */
WP_B200_API int wp_b200_nccl_allgather_part(const void* send, void* recv, size_t part_bytes, size_t shard_stride_bytes,
                                            size_t offset_bytes);
/* several fields in one NCCL launch (group of all-gathers), on the current stream or (on_comm_stream = 1) on the
 * library's communication stream; mark / wait_mark: per-buffer completion points on the communication stream */
WP_B200_API int wp_b200_nccl_allgather_multi(const void* const* send, void* const* recv, const size_t* bytes_per_rank,
                                             int count, int on_comm_stream);
/* peer-memory gather over NVLink: CUDA IPC mapping of a peer's result buffers, then every rank pushes its shard into
 * every peer's buffer with copy-engine memcpys on the communication stream, fenced by two 4-byte NCCL all-reduces */
WP_B200_API int wp_b200_ipc_get_handle(void* device_ptr, void* handle72);   /* 64-byte IPC handle + 8-byte offset */
WP_B200_API void* wp_b200_ipc_open_handle(const void* handle72);
WP_B200_API void wp_b200_ipc_close_handle(void* peer_ptr, const void* handle72);
WP_B200_API int wp_b200_p2p_allgather_multi(const void* const* send, void* const* own_recv, void* const* peer_recv,
                                            const size_t* bytes_per_rank, int count, int rank);
WP_B200_API void* wp_b200_nccl_comm_stream(void);
WP_B200_API int wp_b200_nccl_mark(int k);
WP_B200_API int wp_b200_nccl_wait_mark(int k);
WP_B200_API int wp_b200_nccl_fork(void);
WP_B200_API int wp_b200_nccl_join(void);
WP_B200_API int wp_b200_nccl_allreduce_max_f32(float* inout_device, size_t count);
WP_B200_API int wp_b200_nccl_barrier(void);
WP_B200_API void wp_b200_nccl_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* WARP_B200_H */
