// Host-side bookkeeping for one BVH / mesh object.  The public handle ("id") is the address of a
// device-resident descriptor laid out like the reference's wp::BVH / wp::Mesh (see
// include/warp_b200.h); this struct is what the library itself works from.
#pragma once

#include "common.cuh"

struct BvhState {
    int n = 0;
    int leaf_size = 1;
    int constructor_type = 2;
    bool is_mesh = false;
    bool host_built = false;  // topology from a host constructor (host_build.cu): no keys, no reference numbering, no refit plan
    int device = 0;
    void* context = nullptr;  // the CUcontext (or ordinal token) the creator passed; echoed in the descriptor

    // per-object options (wp_b200_bvh_set_option); the process-wide setters only provide the defaults at creation
    int morton_bits = 30;   // 30 = the reference's code (parity), 63 = 21 bits per axis (fixed at creation)
    int refit_mode = -1;    // -1 process default, 0 auto, 1 atomic counters, 2 wavefront
    int query_order = -1;   // point batches: -1 process default, 0 input order, 1 curve order, 2 auto (>= 32768 points)
    int ray_order = -1;     // ray batches: -1 process default, 0 input order, 1 origin / direction order
    int auto_reference_layout = 0;  // 1: create / refit / rebuild also refresh the reference-layout mirror (drop-in use)

    // borrowed inputs (owned by the caller, like the reference: warp.h:105-106)
    const float* item_lowers = nullptr;
    const float* item_uppers = nullptr;
    const int* groups = nullptr;
    const float* points = nullptr;
    const int* indices = nullptr;
    int num_points = 0;

    // owned tree storage: every pointer below up to `partials` is carved out of one arena
    void* arena = nullptr;
    size_t arena_bytes = 0;
    bool arena_async = false;
    void* keys = nullptr;         // n sorted keys: uint32 (30-bit Morton) or uint64 (group<<32|code, or 63-bit Morton)
    int key_bytes = 4;
    int* prim = nullptr;          // n primitive_indices
    NodeRec* pairs = nullptr;     // 2*(n-1) node records
    int* parent_int = nullptr;    // n-1: parent (reference index) of internal node n+s, -1 for the root
    int* pos_parent = nullptr;    // n: parent of the visible leaf starting at sorted position i, else -1
    unsigned* counters = nullptr; // n-1 arrival counters (build: count | height<<8; refit: parity)
    float4* tris = nullptr;       // 3n packed triangles in sorted order (mesh only)
    TreeHeader* header = nullptr;
    uint16_t* heights = nullptr;  // n-1: height of internal node n+s (original leaves = 0), capped at 0xffff

    // refit plan (bvh_refit.cu), built lazily by the first refit after a build: the visible internal nodes whose
    // range lies inside one block of WB_WAVE_BP sorted positions, grouped by block and ordered by height
    bool plan_valid = false;
    uint32_t* plan_keys = nullptr;   // n-1 sorted keys: block << 12 | height; 0xFFFFFFFE spanning, 0xFFFFFFFF leaf / muted
    int* plan_nodes = nullptr;       // n-1: internal slot s of each entry
    uint32_t* plan_dst = nullptr;    // n-1: record the entry's box goes to (2 * parent slot + side; bit 31: parent spans blocks)
    int* plan_begin = nullptr;       // per block: its entries are [plan_begin[b], plan_end[b])
    int* plan_end = nullptr;
    uint8_t* unit_flags = nullptr;   // n: 1 where the visible leaf starting at the position has a block-spanning parent
    uint32_t* plan_top = nullptr;    // <= n entries (2 * parent slot + side): the nodes / leaves that announce themselves
    int* plan_ntop = nullptr;        //   on the global counters, and their number (device side)

    // build workspace, kept so rebuild() allocates nothing (bvh.cu:790-803 semantics)
    void* keys_alt = nullptr;
    int* prim_alt = nullptr;
    uint32_t* ghist = nullptr;        // 4 x 256 digit histograms
    uint32_t* tile_status = nullptr;  // 4 passes x tiles x 256 look-back words
    unsigned* tickets = nullptr;      // small block of counters (tile tickets, last-block tickets)
    float* partials = nullptr;        // per-block scene-bounds partials
    int* bfs_counts = nullptr;        // 64 ints: frontier sizes of the top-down depth sweep (bvh_build.cu k_bfs_level)
    double* edge_partials = nullptr;  // 296 doubles: per-block partial sums of the average-edge-length reduction (query.cu)
    void* cub_temp = nullptr;         // only used by the WARP_B200_SORT=cub cross-check path
    size_t cub_temp_bytes = 0;
    int num_tiles = 0;
    int bounds_blocks = 0;

    // lazily materialised mirror in the reference's own node layout (bvh.h:161-207)
    void* ref_lowers = nullptr;
    void* ref_uppers = nullptr;
    int* ref_parents = nullptr;
    int* ref_root = nullptr;
    int* ref_counts = nullptr;

    // per-item bounds in the reference's layout: wp::Mesh::lowers / uppers and BVH::item_lowers / item_uppers of a mesh
    // (mesh.cu:279-280, 310-313; read by mesh_query_aabb / bvh_query_* kernels through the id).  Materialised together
    // with the node arrays by wb_export_reference_layout.
    float* tri_lowers = nullptr;
    float* tri_uppers = nullptr;

    void* dev_desc = nullptr;  // device copy of the reference-compatible descriptor; its address is the id
    void* host_desc = nullptr; // pinned staging ring of the descriptor (stable addresses: uploads are async and capture-safe)
    unsigned desc_slot = 0;
};

struct MeshState {
    BvhState bvh;
    uint64_t points_data = 0, velocities_data = 0, indices_data = 0;
    int num_points = 0, num_tris = 0;
    int points_shape0 = 0, velocities_shape0 = 0;
    void* dev_desc = nullptr;
    void* host_desc = nullptr;  // pinned staging ring (see BvhState::host_desc)
    unsigned desc_slot = 0;
    bool desc_initialised = false;
};

// build / refit / export drivers (bvh_build.cu, bvh_refit.cu); all enqueue on `stream`
const char* wb_build(BvhState& s, cudaStream_t stream);
const char* wb_build_host(BvhState& s, cudaStream_t stream);  // constructor_type 0 (sah) / 1 (median), host_build.cu
const char* wb_refit(BvhState& s, cudaStream_t stream);
extern int g_wb_small_nodes;  // experiment switch of the builder's Karras-style small-node pass (bvh_build.cu)
extern int g_wb_refit_mode;  // default refit mode of new trees: 0 auto, 1 atomic counters, 2 wavefront
const char* wb_refit_plan(BvhState& s, cudaStream_t stream);  // (bvh_build.cu: shares the radix sort)
const char* wb_refit_merge(BvhState& s, cudaStream_t stream);  // bottom-up pass of the refit (bvh_build.cu)
// experiment: parents of all internal nodes from the sorted keys alone (bvh_build.cu, k_topology)
const char* wb_experiment_topology(BvhState& s, int* parent_out, int* fail, cudaStream_t stream);
const char* wb_export_reference_layout(BvhState& s, cudaStream_t stream);
const char* wb_alloc_tree(BvhState& s, cudaStream_t stream);
void wb_free_tree(BvhState& s, cudaStream_t stream);
