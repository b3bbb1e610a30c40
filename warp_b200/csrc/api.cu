// C ABI of libwarp_b200.so (declared in include/warp_b200.h).
#include "../../include/warp_b200.h"

#include "order.h"
#include "query.h"
#include "state.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

static_assert(sizeof(wp_array_t) == 56, "wp::array_t layout (warp/native/array.h:173-277)");
static_assert(sizeof(wp_b200_bvh_desc) == 112, "wp::BVH layout (warp/native/bvh.h:176-207)");
static_assert(sizeof(wp_b200_mesh_desc) == 328, "wp::Mesh layout (warp/native/mesh.h:18-35)");
static_assert(sizeof(NodeRec) == 32, "node record");

namespace {

char g_error[4096] = "";
std::mutex g_lock;
std::map<uint64_t, BvhState*> g_bvhs;
std::map<uint64_t, MeshState*> g_meshes;
cudaStream_t g_stream[64] = {};  // current stream per device (0 = legacy default stream)

#ifndef WB_L2_PERSIST_DEFAULT
#define WB_L2_PERSIST_DEFAULT 0
#endif
thread_local bool t_stats_enabled = false;
// Process-wide DEFAULTS, copied into every tree when it is created (BvhState::morton_bits / query_order / ray_order /
// refit_mode); changing them later never touches an existing object -- wp_b200_bvh_set_option does that, per object.
int g_morton_bits = 30;         // Morton resolution: 30 (reference parity) or 63
int g_query_order = 2;          // 0 input order, 1 curve order, 2 auto (curve order for batches >= 32768 points)
int g_ray_order = 0;            // 0 input order (default), 1 origin/direction order
int g_auto_reference_layout = 0;  // 1: new trees keep the reference-layout mirror current (set by the drop-in stub)
// query-ordering scratch, one per (device, stream): two batches on different streams never share a permutation buffer,
// and batches on the same stream are ordered by the stream itself
std::map<std::pair<int, cudaStream_t>, OrderScratch*> g_order;
unsigned long long* g_stats_dev = nullptr;

OrderScratch& order_scratch(int device, cudaStream_t stream)
{
    std::lock_guard<std::mutex> g(g_lock);
    OrderScratch*& p = g_order[std::make_pair(device, stream)];
    if (!p)
        p = new OrderScratch();
    return *p;
}

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int current_device()
{
    int d = 0;
    cudaGetDevice(&d);
    return d;
}

// ---- CUcontext handling.  The reference passes a real CUcontext as `context` (warp/_src/types.py:5947-5950,
// 6204-6205: `self.device.context`), so the handle is resolved through the driver API: libcuda.so.1 is dlopen()ed
// lazily (the library must still load, and export its symbols, on a box without a driver).  NULL = the calling
// thread's current device.  Values 1..64 are the ordinal + 1 tokens handed out when no driver library is present.
typedef int (*cu_ctx_get_current_t)(void**);
typedef int (*cu_ctx_push_t)(void*);
typedef int (*cu_ctx_pop_t)(void**);
typedef int (*cu_ctx_get_device_t)(int*);
struct DriverApi {
    bool tried = false, ok = false;
    cu_ctx_get_current_t get_current = nullptr;
    cu_ctx_push_t push = nullptr;
    cu_ctx_pop_t pop = nullptr;
    cu_ctx_get_device_t get_device = nullptr;
};
DriverApi g_driver;
std::map<void*, int> g_ctx_device;  // CUcontext -> device ordinal (guarded by g_lock)

const DriverApi& driver_api()
{
    std::lock_guard<std::mutex> g(g_lock);
    if (!g_driver.tried) {
        g_driver.tried = true;
        void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            g_driver.get_current = (cu_ctx_get_current_t)dlsym(h, "cuCtxGetCurrent");
            g_driver.push = (cu_ctx_push_t)dlsym(h, "cuCtxPushCurrent_v2");
            g_driver.pop = (cu_ctx_pop_t)dlsym(h, "cuCtxPopCurrent_v2");
            g_driver.get_device = (cu_ctx_get_device_t)dlsym(h, "cuCtxGetDevice");
            g_driver.ok = g_driver.get_current && g_driver.push && g_driver.pop && g_driver.get_device;
        }
    }
    return g_driver;
}

int context_device(void* context)
{
    if (!context)
        return current_device();
    const uintptr_t v = (uintptr_t)context;
    if (v <= 64)
        return (int)v - 1;
    {
        std::lock_guard<std::mutex> g(g_lock);
        auto it = g_ctx_device.find(context);
        if (it != g_ctx_device.end())
            return it->second;
    }
    const DriverApi& d = driver_api();
    int dev = current_device();
    if (d.ok && d.push(context) == 0) {
        int got = -1;
        if (d.get_device(&got) == 0 && got >= 0)
            dev = got;
        void* popped = nullptr;
        d.pop(&popped);
    }
    std::lock_guard<std::mutex> g(g_lock);
    g_ctx_device[context] = dev;
    return dev;
}

// the primary context of `ordinal` as the driver knows it (what Warp's Device.context holds), or the ordinal token
void* device_primary_context(int ordinal)
{
    const DriverApi& d = driver_api();
    if (!d.ok)
        return (void*)(intptr_t)(ordinal + 1);
    int prev = 0;
    cudaGetDevice(&prev);
    void* ctx = nullptr;
    if (cudaSetDevice(ordinal) == cudaSuccess && cudaFree(nullptr) == cudaSuccess)  // cudaFree(0) creates the primary context
        d.get_current(&ctx);
    cudaSetDevice(prev);
    if (!ctx)
        return (void*)(intptr_t)(ordinal + 1);
    std::lock_guard<std::mutex> g(g_lock);
    g_ctx_device[ctx] = ordinal;
    return ctx;
}

cudaStream_t current_stream(int device) { return (device >= 0 && device < 64) ? g_stream[device] : 0; }

struct DeviceGuard {
    int prev;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (dev != prev)
            cudaSetDevice(dev);
    }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

bool check(cudaError_t e, const char* what)
{
    if (e == cudaSuccess)
        return true;
    set_error("Warp B200 CUDA error in %s: %s", what, cudaGetErrorString(e));
    return false;
}

void fill_bvh_desc(const BvhState& s, wp_b200_bvh_desc& d, void* context)
{
    memset(&d, 0, sizeof(d));
    d.node_lowers = s.ref_lowers;
    d.node_uppers = s.ref_uppers;
    d.node_parents = s.ref_parents;
    d.node_counts = s.ref_counts;
    d.primitive_indices = s.prim;
    d.max_depth = 0;
    d.max_nodes = s.n > 0 ? 2 * s.n - 1 : 0;
    d.num_nodes = d.max_nodes;
    d.num_leaf_nodes = s.n;
    d.root = s.ref_root;
    // a mesh's BVH items are its per-triangle bounds (mesh.cu:310-313), materialised by wb_export_reference_layout
    d.item_lowers = (wp_vec3*)(s.is_mesh ? s.tri_lowers : s.item_lowers);
    d.item_uppers = (wp_vec3*)(s.is_mesh ? s.tri_uppers : s.item_uppers);
    d.item_groups = (int*)s.groups;
    d.num_items = s.n;
    d.leaf_size = s.leaf_size;
    d.constructor_type = s.constructor_type;
    d.context = context;
}

BvhState* find_tree(uint64_t id, MeshState** mesh_out = nullptr)
{
    std::lock_guard<std::mutex> g(g_lock);
    auto m = g_meshes.find(id);
    if (m != g_meshes.end()) {
        if (mesh_out)
            *mesh_out = m->second;
        return &m->second->bvh;
    }
    auto b = g_bvhs.find(id);
    if (b != g_bvhs.end())
        return b->second;
    return nullptr;
}

// The descriptor is rewritten in stream order from a pinned staging copy owned by the object (stable address, so the
// upload is truly asynchronous and legal under stream capture); everything up to average_edge_length is uploaded --
// that last field is device-computed (wb_query_point_sign_normal / wp_b200_bvh_sync_reference_layout) and must
// survive a points / velocities swap.  The staging copy is a ring of DESC_RING slots so that back-to-back updates
// (points = A; points = B) each upload their own bytes.
constexpr unsigned DESC_RING = 8;
bool upload_desc(MeshState* ms, BvhState* bs)
{
    if (ms) {
        if (!ms->host_desc && !check(cudaMallocHost(&ms->host_desc, DESC_RING * sizeof(wp_b200_mesh_desc)), "descriptor staging"))
            return false;
        wp_b200_mesh_desc& d = ((wp_b200_mesh_desc*)ms->host_desc)[ms->desc_slot++ % DESC_RING];
        memset(&d, 0, sizeof(d));
        d.points.data = ms->points_data, d.points.shape[0] = ms->points_shape0, d.points.strides[0] = 12, d.points.ndim = 1;
        d.velocities.data = ms->velocities_data, d.velocities.shape[0] = ms->velocities_shape0;
        d.velocities.strides[0] = 12, d.velocities.ndim = ms->velocities_data ? 1 : 0;
        d.indices.data = ms->indices_data, d.indices.shape[0] = ms->num_tris * 3, d.indices.strides[0] = 4, d.indices.ndim = 1;
        d.lowers = (wp_vec3*)ms->bvh.tri_lowers, d.uppers = (wp_vec3*)ms->bvh.tri_uppers;
        d.num_points = ms->num_points, d.num_tris = ms->num_tris;
        fill_bvh_desc(ms->bvh, d.bvh, ms->bvh.context);
        d.context = d.bvh.context;
        const size_t bytes = ms->desc_initialised ? offsetof(wp_b200_mesh_desc, average_edge_length) : sizeof(d);
        ms->desc_initialised = true;
        return check(cudaMemcpyAsync(ms->dev_desc, &d, bytes, cudaMemcpyHostToDevice, current_stream(ms->bvh.device)),
                     "descriptor upload");
    }
    if (!bs->host_desc && !check(cudaMallocHost(&bs->host_desc, DESC_RING * sizeof(wp_b200_bvh_desc)), "descriptor staging"))
        return false;
    wp_b200_bvh_desc& d = ((wp_b200_bvh_desc*)bs->host_desc)[bs->desc_slot++ % DESC_RING];
    fill_bvh_desc(*bs, d, bs->context);
    return check(cudaMemcpyAsync(bs->dev_desc, &d, sizeof(d), cudaMemcpyHostToDevice, current_stream(bs->device)),
                 "descriptor upload");
}

bool constructor_supported(int constructor_type, const int* groups)
{
    if (constructor_type == WP_BVH_CONSTRUCTOR_LBVH)
        return true;
    if (constructor_type == 0 || constructor_type == 1) {  // sah / median: built on the host, uploaded (host_build.cu)
        if (!groups)
            return true;
        set_error("Warp error: grouped trees are built with constructor 'lbvh' (%d) in the B200 library; the host "
                  "constructors (sah / median) take ungrouped items only", WP_BVH_CONSTRUCTOR_LBVH);
        return false;
    }
    set_error("Warp error: BVH constructor %d is not available in the B200 library (lbvh = %d, sah = 0, median = 1; "
              "cuBQL is out of scope)", constructor_type, WP_BVH_CONSTRUCTOR_LBVH);
    return false;
}

const char* build_tree(BvhState& s, cudaStream_t stream)
{
    return s.constructor_type == WP_BVH_CONSTRUCTOR_LBVH ? wb_build(s, stream) : wb_build_host(s, stream);
}

TreeView make_view(const BvhState& s)
{
    TreeView tv;
    tv.pairs = s.pairs;
    tv.header = s.header;
    tv.tris = s.tris;
    tv.prim = s.prim;
    tv.parent_int = s.parent_int;
    tv.pos_parent = s.pos_parent;
    tv.n = s.n;
    return tv;
}

unsigned long long* stats_buffer()
{
    if (!t_stats_enabled)
        return nullptr;
    if (!g_stats_dev) {
        if (cudaMalloc(&g_stats_dev, 2 * sizeof(unsigned long long)) != cudaSuccess)
            return nullptr;
    }
    cudaMemsetAsync(g_stats_dev, 0, 2 * sizeof(unsigned long long), current_stream(current_device()));
    return g_stats_dev;
}

// ---- live kernel timing (bench.py roofline): CUDA events around the traversal kernel itself, on the stream it is
// launched on, accumulated until read.  Off by default; skipped while the stream is being captured.
struct KernelTimer {
    bool enabled = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    std::vector<cudaEvent_t> pool;
};
KernelTimer g_ktimer;

cudaEvent_t ktimer_event()
{
    if (!g_ktimer.pool.empty()) {
        cudaEvent_t e = g_ktimer.pool.back();
        g_ktimer.pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct KernelTimerScope {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    explicit KernelTimerScope(cudaStream_t stream) : st(stream)
    {
        if (!g_ktimer.enabled)
            return;
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (stream && cudaStreamIsCapturing(stream, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone)
            return;
        std::lock_guard<std::mutex> g(g_lock);
        a = ktimer_event(), b = ktimer_event();
        cudaEventRecord(a, st);
    }
    ~KernelTimerScope()
    {
        if (!a)
            return;
        cudaEventRecord(b, st);
        std::lock_guard<std::mutex> g(g_lock);
        g_ktimer.pending.emplace_back(a, b);
    }
};

// memory-side variant of the unsigned closest-point kernel (query.cu QM_* bits); WARP_B200_QMODE overrides the default
#ifndef WB_QMODE_DEFAULT
#define WB_QMODE_DEFAULT 22  // QM_STREAM | QM_PACKED | QM_STACK8: measured best on C2 / C4 (scripts/qmode_ab.py)
#endif
int query_mode()
{
    static const int mode = [] {
        const char* e = getenv("WARP_B200_QMODE");
        return e ? atoi(e) : WB_QMODE_DEFAULT;
    }();
    return mode;
}

// L2 residency of the tree (BASELINE north_star: "node layout kept L2-resident"): while a traversal kernel runs, the
// sibling-pair array is covered by a PERSISTING access-policy window on the launch stream, sized to the device's
// persisting-L2 carve-out (hit ratio = carve-out / array size when the array is larger), everything else on the stream
// being streaming-class traffic.  WARP_B200_L2_PERSIST=0 turns it off (A/B).
struct L2PersistScope {
    cudaStream_t st;
    bool active = false;
    static int enabled()
    {
        static const int on = [] {
            const char* e = getenv("WARP_B200_L2_PERSIST");
            return e ? atoi(e) : WB_L2_PERSIST_DEFAULT;
        }();
        return on;
    }
    L2PersistScope(const BvhState& s, cudaStream_t stream) : st(stream)
    {
        if (!enabled() || s.n < 2)
            return;
        static size_t carve[64] = {};
        static int max_window[64] = {};
        const int d = s.device;
        if (d < 0 || d >= 64)
            return;
        if (!max_window[d]) {
            int persist_max = 0, win = 0;
            cudaDeviceGetAttribute(&persist_max, cudaDevAttrMaxPersistingL2CacheSize, d);
            cudaDeviceGetAttribute(&win, cudaDevAttrMaxAccessPolicyWindowSize, d);
            if (persist_max <= 0 || win <= 0) {
                max_window[d] = -1;
                return;
            }
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)persist_max);
            carve[d] = (size_t)persist_max, max_window[d] = win;
        }
        if (max_window[d] < 0)
            return;
        const size_t bytes = sizeof(NodeRec) * 2 * (size_t)(s.n - 1);
        cudaStreamAttrValue v;
        memset(&v, 0, sizeof(v));
        v.accessPolicyWindow.base_ptr = (void*)s.pairs;
        v.accessPolicyWindow.num_bytes = bytes < (size_t)max_window[d] ? bytes : (size_t)max_window[d];
        const double ratio = (double)carve[d] / (double)v.accessPolicyWindow.num_bytes;
        v.accessPolicyWindow.hitRatio = ratio >= 1.0 ? 1.0f : (float)ratio;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        active = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v) == cudaSuccess;
        if (!active)
            cudaGetLastError();
    }
    ~L2PersistScope()
    {
        if (!active)
            return;
        cudaStreamAttrValue v;
        memset(&v, 0, sizeof(v));
        v.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v);
    }
};

MeshState* query_mesh(uint64_t id)
{
    MeshState* ms = nullptr;
    find_tree(id, &ms);
    if (!ms)
        set_error("Warp error: invalid mesh id");
    return ms;
}

// grow-only device scratch for the *_host entry points: two lanes, each with its own stream
struct HostLane {
    cudaStream_t stream = nullptr;
    void* buf = nullptr;
    size_t bytes = 0;
};
HostLane g_lanes[64][2];

bool lane_reserve(HostLane& l, size_t bytes)
{
    if (!l.stream && !check(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking), "stream create"))
        return false;
    if (bytes > l.bytes) {
        if (l.buf)
            cudaFree(l.buf);
        l.buf = nullptr, l.bytes = 0;
        if (!check(cudaMalloc(&l.buf, bytes), "scratch alloc"))
            return false;
        l.bytes = bytes;
    }
    return true;
}

// Queries per staged chunk of the *_host entry points (env WARP_B200_HOST_CHUNK forces a fixed size).  Measured on
// B200 / C2 / 16.8 M pinned queries, fixed chunk sizes: 2 M 516, 3 M 538, 4 M 547, 6 M 555, 8 M 552 M queries/s -- small
// chunks lose more in traversal coherence (a Morton-sorted chunk is sparser than the sorted batch) than they gain
// in copy overlap; a three-stage copy/compute/copy pipeline was slower still.  Default: batches up to 4 M go in one
// piece, larger ones in ceil(n / 6 M) >= 2 equal chunks; above 8 M the closest-point calls additionally start and
// end with a 1 M chunk (563 vs 557 M queries/s: the first upload and the last download are the exposed copies).
int64_t host_chunk_forced()
{
    static int64_t forced = -1;
    if (forced < 0) {
        const char* e = getenv("WARP_B200_HOST_CHUNK");
        forced = e ? atoll(e) : 0;
        if (forced && forced < 1024)
            forced = 1024;
    }
    return forced;
}

int64_t host_chunk(int64_t n)
{
    if (const int64_t forced = host_chunk_forced())
        return forced;
    if (n <= (1ll << 22))
        return n > 0 ? n : 1;
    const int64_t k = (n + (6ll << 20) - 1) / (6ll << 20);
    const int64_t parts = k < 2 ? 2 : k;
    return (((n + parts - 1) / parts) + 1023) & ~(int64_t)1023;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

const char* wp_get_error_string(void) { return g_error; }

int wp_init(const char*) { return 0; }
int wp_is_cuda_enabled(void) { return 1; }

int wp_cuda_device_get_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void* wp_cuda_device_get_primary_context(int ordinal) { return device_primary_context(ordinal); }
void* wp_cuda_context_get_current(void) { return device_primary_context(current_device()); }
void wp_cuda_context_set_current(void* context)
{
    if (context)
        cudaSetDevice(context_device(context));
}
void wp_cuda_context_synchronize(void* context)
{
    DeviceGuard g(context_device(context));
    check(cudaDeviceSynchronize(), "device synchronize");
}
void* wp_cuda_context_get_stream(void* context) { return current_stream(context_device(context)); }
void wp_cuda_context_set_stream(void* context, void* stream, int sync)
{
    const int d = context_device(context);
    if (d < 0 || d >= 64)
        return;
    if (sync)
        cudaStreamSynchronize(g_stream[d]);
    g_stream[d] = (cudaStream_t)stream;
}
void* wp_cuda_stream_create(void* context, int priority)
{
    DeviceGuard g(context_device(context));
    cudaStream_t s = nullptr;
    if (!check(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, priority), "stream create"))
        return nullptr;
    return s;
}
void wp_cuda_stream_destroy(void*, void* stream)
{
    // the stream's query-ordering scratch (per (device, stream), up to ~45 bytes per query of the largest batch) goes with
    // it: a program that creates a stream per captured graph would otherwise keep every scratch forever
    std::vector<OrderScratch*> gone;
    {
        std::lock_guard<std::mutex> g(g_lock);
        for (auto it = g_order.begin(); it != g_order.end();) {
            if (it->first.second == (cudaStream_t)stream) {
                gone.push_back(it->second);
                it = g_order.erase(it);
            } else {
                ++it;
            }
        }
    }
    if (!gone.empty())
        cudaStreamSynchronize((cudaStream_t)stream);  // nothing on the stream may still read the scratch
    for (OrderScratch* ws : gone) {
        if (ws) {
            wb_order_free(*ws);
            delete ws;
        }
    }
    cudaStreamDestroy((cudaStream_t)stream);
}
void wp_cuda_stream_synchronize(void* stream) { check(cudaStreamSynchronize((cudaStream_t)stream), "stream synchronize"); }
void* wp_cuda_event_create(void* context, unsigned flags)
{
    DeviceGuard g(context_device(context));
    cudaEvent_t e = nullptr;
    if (!check(cudaEventCreateWithFlags(&e, (flags & 1u) ? cudaEventDisableTiming : cudaEventDefault), "event create"))
        return nullptr;
    return e;
}
void wp_cuda_event_destroy(void* event) { cudaEventDestroy((cudaEvent_t)event); }
void wp_cuda_event_record(void* event, void* stream, int) { check(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream), "event record"); }
void wp_cuda_event_synchronize(void* event) { check(cudaEventSynchronize((cudaEvent_t)event), "event synchronize"); }
float wp_cuda_event_elapsed_time(void* a, void* b)
{
    float ms = 0.f;
    check(cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b), "event elapsed");
    return ms;
}

// CUDA graph capture of the path's stream work (refit / in-place rebuild / device-buffer queries), names and
// argument meaning of warp/native/warp.h:766-771.  The stream must be a created stream (not the legacy default).
int wp_cuda_graph_begin_capture(void* context, void* stream, int external, int mode)
{
    if (external)
        return 1;  // the caller's own capture is already active on `stream`
    const int dev = context_device(context);
    DeviceGuard g(dev);
    {
        // query-ordering scratch is per (device, stream) and cannot grow inside a capture: give this stream one as large
        // as the largest in use on the device (i.e. what the warm-up run of the loop body needed)
        long long cap = 0;
        {
            std::lock_guard<std::mutex> l(g_lock);
            for (auto& kv : g_order)
                if (kv.first.first == dev && kv.second && kv.second->capacity > cap)
                    cap = kv.second->capacity;
        }
        if (cap > 0) {
            const char* e = wb_order_reserve(order_scratch(dev, (cudaStream_t)stream), cap, nullptr);
            if (e) {
                set_error("Warp error: %s", e);
                return 0;
            }
        }
    }
    return check(cudaStreamBeginCapture((cudaStream_t)stream, (cudaStreamCaptureMode)mode), "graph begin capture") ? 1 : 0;
}
int wp_cuda_graph_end_capture(void* context, void* stream, void** graph_ret)
{
    DeviceGuard g(context_device(context));
    cudaGraph_t graph = nullptr;
    const bool ok = check(cudaStreamEndCapture((cudaStream_t)stream, &graph), "graph end capture");
    if (graph_ret)
        *graph_ret = graph;
    else if (graph)
        cudaGraphDestroy(graph);
    return ok && graph ? 1 : 0;
}
int wp_cuda_graph_create_exec(void* context, void*, void* graph, void** graph_exec_ret)
{
    DeviceGuard g(context_device(context));
    cudaGraphExec_t exec = nullptr;
    if (!check(cudaGraphInstantiateWithFlags(&exec, (cudaGraph_t)graph, 0), "graph instantiate"))
        return 0;
    *graph_exec_ret = exec;
    return 1;
}
int wp_cuda_graph_launch(void* graph_exec, void* stream)
{
    return check(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream), "graph launch") ? 1 : 0;
}
int wp_cuda_graph_destroy(void*, void* graph) { return check(cudaGraphDestroy((cudaGraph_t)graph), "graph destroy") ? 1 : 0; }
int wp_cuda_graph_exec_destroy(void*, void* graph_exec)
{
    return check(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec), "graph exec destroy") ? 1 : 0;
}

void* wp_alloc_device(void* context, size_t s, const char*)
{
    DeviceGuard g(context_device(context));
    void* p = nullptr;
    if (s == 0)
        return nullptr;
    if (!check(cudaMalloc(&p, s), "wp_alloc_device"))
        return nullptr;
    return p;
}
void wp_free_device(void* context, void* ptr)
{
    if (!ptr)
        return;
    DeviceGuard g(context_device(context));
    cudaFree(ptr);
}
void* wp_alloc_pinned(size_t s, const char*)
{
    void* p = nullptr;
    if (s == 0)
        return nullptr;
    if (!check(cudaMallocHost(&p, s), "wp_alloc_pinned"))
        return nullptr;
    return p;
}
void wp_free_pinned(void* ptr)
{
    if (ptr)
        cudaFreeHost(ptr);
}
int wp_memcpy_h2d(void* context, void* dest, void* src, size_t n, void* stream)
{
    DeviceGuard g(context_device(context));
    return check(cudaMemcpyAsync(dest, src, n, cudaMemcpyHostToDevice, (cudaStream_t)stream), "memcpy h2d");
}
int wp_memcpy_d2h(void* context, void* dest, void* src, size_t n, void* stream)
{
    DeviceGuard g(context_device(context));
    return check(cudaMemcpyAsync(dest, src, n, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "memcpy d2h");
}
int wp_memcpy_d2d(void* context, void* dest, void* src, size_t n, void* stream)
{
    DeviceGuard g(context_device(context));
    return check(cudaMemcpyAsync(dest, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream), "memcpy d2d");
}
int wp_memset_device(void* context, void* dest, int value, size_t n, void* stream)
{
    DeviceGuard g(context_device(context));
    return check(cudaMemsetAsync(dest, value, n, (cudaStream_t)stream), "memset");
}

int wp_b200_device_attr(int ordinal, const char* name, long long* value)
{
    int v = 0;
    cudaError_t e = cudaSuccess;
    if (!strcmp(name, "sm_count"))
        e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, ordinal);
    else if (!strcmp(name, "l2_bytes"))
        e = cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, ordinal);
    else if (!strcmp(name, "clock_khz"))
        e = cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, ordinal);
    else if (!strcmp(name, "cc_major"))
        e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, ordinal);
    else if (!strcmp(name, "cc_minor"))
        e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, ordinal);
    else if (!strcmp(name, "mem_free") || !strcmp(name, "mem_total")) {
        DeviceGuard g(ordinal);
        size_t f = 0, t = 0;
        e = cudaMemGetInfo(&f, &t);
        *value = (long long)(!strcmp(name, "mem_free") ? f : t);
        return check(e, "mem info");
    } else {
        set_error("unknown device attribute %s", name);
        return 0;
    }
    *value = v;
    return check(e, "device attribute");
}

// device ordinal that owns a device pointer (cudaPointerGetAttributes); 0 when the pointer is not device memory
int wp_b200_pointer_device(const void* ptr, long long* ordinal)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess || (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        return 0;
    }
    *ordinal = attr.device;
    return 1;
}

int wp_b200_device_name(int ordinal, char* buf, int len)
{
    cudaDeviceProp p;
    if (!check(cudaGetDeviceProperties(&p, ordinal), "device properties"))
        return 0;
    snprintf(buf, len, "%s", p.name);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// Bvh
// ------------------------------------------------------------------------------------------------
static int sync_reference_layout(BvhState* s, MeshState* m);

uint64_t wp_bvh_create_device(void* context, wp_vec3* lowers, wp_vec3* uppers, int num_items, int constructor_type,
                              int* groups, int leaf_size)
{
    return wp_b200_bvh_create_device_ex(context, lowers, uppers, num_items, constructor_type, groups, leaf_size, 0);
}

uint64_t wp_b200_bvh_create_device_ex(void* context, wp_vec3* lowers, wp_vec3* uppers, int num_items, int constructor_type,
                                      int* groups, int leaf_size, int morton_bits)
{
    g_error[0] = 0;  // the drop-in stub reads this library's message only when the last routed call failed
    if (!constructor_supported(constructor_type, groups))
        return 0;
    if (morton_bits == 0)
        morton_bits = g_morton_bits;
    if (morton_bits != 30 && morton_bits != 63) {
        set_error("Warp error: morton_bits must be 30 or 63 (got %d)", morton_bits);
        return 0;
    }
    if (morton_bits == 63 && groups) {
        set_error("Warp error: morton_bits=63 cannot be combined with groups (the key holds group << 32 | 30-bit code)");
        return 0;
    }
    if (num_items < 0 || leaf_size < 1) {
        set_error("Warp error: invalid BVH arguments (num_items=%d, leaf_size=%d)", num_items, leaf_size);
        return 0;
    }
    const int dev = context_device(context);
    DeviceGuard g(dev);
    BvhState* s = new BvhState();
    s->n = num_items, s->leaf_size = leaf_size, s->constructor_type = constructor_type, s->device = dev;
    s->context = context ? context : device_primary_context(dev);
    s->morton_bits = morton_bits;
    s->auto_reference_layout = constructor_type == WP_BVH_CONSTRUCTOR_LBVH ? g_auto_reference_layout : 0;  // no mirror of host-built trees
    s->key_bytes = (groups || morton_bits == 63) ? 8 : 4;
    s->item_lowers = (const float*)lowers, s->item_uppers = (const float*)uppers, s->groups = groups;
    const char* err = num_items > 0 ? wb_alloc_tree(*s, current_stream(dev)) : nullptr;
    if (!err)
        err = build_tree(*s, current_stream(dev));
    if (err || !check(cudaMalloc(&s->dev_desc, sizeof(wp_b200_bvh_desc)), "descriptor alloc") || !upload_desc(nullptr, s)
        || (s->auto_reference_layout && !sync_reference_layout(s, nullptr))) {
        if (err)
            set_error("Warp error: BVH build failed: %s", err);
        if (s->dev_desc)
            cudaFree(s->dev_desc);
        if (s->host_desc)
            cudaFreeHost(s->host_desc);
        wb_free_tree(*s, current_stream(dev));
        delete s;
        return 0;
    }
    const uint64_t id = (uint64_t)s->dev_desc;
    std::lock_guard<std::mutex> l(g_lock);
    g_bvhs[id] = s;
    return id;
}

void wp_bvh_destroy_device(uint64_t id)
{
    g_error[0] = 0;  // the drop-in stub reads this library's message only when the last routed call failed
    BvhState* s = nullptr;
    {
        std::lock_guard<std::mutex> l(g_lock);
        auto it = g_bvhs.find(id);
        if (it == g_bvhs.end())
            return;
        s = it->second;
        g_bvhs.erase(it);
    }
    DeviceGuard g(s->device);
    cudaStreamSynchronize(current_stream(s->device));
    cudaFree(s->dev_desc);
    if (s->host_desc)
        cudaFreeHost(s->host_desc);
    wb_free_tree(*s, current_stream(s->device));
    delete s;
}

void wp_bvh_refit_device(uint64_t id)
{
    g_error[0] = 0;  // the drop-in stub reads this library's message only when the last routed call failed
    BvhState* s = find_tree(id);
    if (!s)
        return;
    DeviceGuard g(s->device);
    const char* err = wb_refit(*s, current_stream(s->device));
    if (err)
        set_error("Warp error: BVH refit failed: %s", err);
    else if (s->auto_reference_layout)
        sync_reference_layout(s, nullptr);
}

void wp_bvh_rebuild_device(uint64_t id)
{
    g_error[0] = 0;  // the drop-in stub reads this library's message only when the last routed call failed
    BvhState* s = find_tree(id);
    if (!s)
        return;
    DeviceGuard g(s->device);
    const char* err = wb_build(*s, current_stream(s->device));
    if (err)
        set_error("Warp error: BVH rebuild failed: %s", err);
    else if (s->auto_reference_layout)
        sync_reference_layout(s, nullptr);
}

// ------------------------------------------------------------------------------------------------
// Mesh
// ------------------------------------------------------------------------------------------------
uint64_t wp_mesh_create_device(void* context, wp_array_t points, wp_array_t velocities, wp_array_t tris, int num_points,
                               int num_tris, int support_winding_number, int constructor_type, int* groups,
                               int bvh_leaf_size)
{
    return wp_b200_mesh_create_device_ex(context, points, velocities, tris, num_points, num_tris, support_winding_number,
                                         constructor_type, groups, bvh_leaf_size, 0);
}

uint64_t wp_b200_mesh_create_device_ex(void* context, wp_array_t points, wp_array_t velocities, wp_array_t tris,
                                       int num_points, int num_tris, int support_winding_number, int constructor_type,
                                       int* groups, int bvh_leaf_size, int morton_bits)
{
    g_error[0] = 0;  // the drop-in stub reads this library's message only when the last routed call failed
    if (!constructor_supported(constructor_type, groups))
        return 0;
    if (morton_bits == 0)
        morton_bits = g_morton_bits;
    if (morton_bits != 30 && morton_bits != 63) {
        set_error("Warp error: morton_bits must be 30 or 63 (got %d)", morton_bits);
        return 0;
    }
    if (morton_bits == 63 && groups) {
        set_error("Warp error: morton_bits=63 cannot be combined with groups (the key holds group << 32 | 30-bit code)");
        return 0;
    }
    if (support_winding_number) {
        set_error("Warp error: support_winding_number=True is out of scope for the B200 mesh path");
        return 0;
    }
    if (num_points < 0 || num_tris < 0 || bvh_leaf_size < 1 || (num_tris > 0 && (!points.data || !tris.data))) {
        set_error("Warp error: invalid mesh arguments (num_points=%d, num_tris=%d, leaf_size=%d)", num_points, num_tris,
                  bvh_leaf_size);
        return 0;
    }
    const int dev = context_device(context);
    DeviceGuard g(dev);
    MeshState* m = new MeshState();
    m->points_data = points.data, m->velocities_data = velocities.data, m->indices_data = tris.data;
    m->num_points = num_points, m->num_tris = num_tris;
    m->points_shape0 = points.shape[0], m->velocities_shape0 = velocities.shape[0];
    BvhState& s = m->bvh;
    s.n = num_tris, s.leaf_size = bvh_leaf_size, s.constructor_type = constructor_type, s.device = dev;
    s.is_mesh = true;
    s.groups = groups;
    s.context = context ? context : device_primary_context(dev);
    s.morton_bits = morton_bits;
    s.auto_reference_layout = constructor_type == WP_BVH_CONSTRUCTOR_LBVH ? g_auto_reference_layout : 0;  // no mirror of host-built trees
    s.key_bytes = (groups || morton_bits == 63) ? 8 : 4;
    s.points = (const float*)points.data, s.indices = (const int*)tris.data, s.num_points = num_points;
    const char* err = num_tris > 0 ? wb_alloc_tree(s, current_stream(dev)) : nullptr;
    if (!err)
        err = build_tree(s, current_stream(dev));
    if (err || !check(cudaMalloc(&m->dev_desc, sizeof(wp_b200_mesh_desc)), "descriptor alloc") || !upload_desc(m, nullptr)
        || (s.auto_reference_layout && !sync_reference_layout(&s, m))) {
        if (err)
            set_error("Warp error: mesh build failed: %s", err);
        if (m->dev_desc)
            cudaFree(m->dev_desc);
        if (m->host_desc)
            cudaFreeHost(m->host_desc);
        wb_free_tree(s, current_stream(dev));
        delete m;
        return 0;
    }
    const uint64_t id = (uint64_t)m->dev_desc;
    std::lock_guard<std::mutex> l(g_lock);
    g_meshes[id] = m;
    return id;
}

void wp_mesh_destroy_device(uint64_t id)
{
    g_error[0] = 0;  // the drop-in stub reads this library's message only when the last routed call failed
    MeshState* m = nullptr;
    {
        std::lock_guard<std::mutex> l(g_lock);
        auto it = g_meshes.find(id);
        if (it == g_meshes.end())
            return;
        m = it->second;
        g_meshes.erase(it);
    }
    DeviceGuard g(m->bvh.device);
    cudaStreamSynchronize(current_stream(m->bvh.device));
    cudaFree(m->dev_desc);
    if (m->host_desc)
        cudaFreeHost(m->host_desc);
    wb_free_tree(m->bvh, current_stream(m->bvh.device));
    delete m;
}

int wp_mesh_refit_device(uint64_t id)
{
    g_error[0] = 0;  // the drop-in stub reads this library's message only when the last routed call failed
    MeshState* m = nullptr;
    find_tree(id, &m);
    if (!m) {
        set_error("Warp error: invalid mesh id");
        return 0;
    }
    DeviceGuard g(m->bvh.device);
    const char* err = wb_refit(m->bvh, current_stream(m->bvh.device));
    if (err) {
        set_error("Warp error: mesh refit failed: %s", err);
        return 0;
    }
    if (m->bvh.auto_reference_layout)
        return sync_reference_layout(&m->bvh, m);
    return 1;
}

void wp_b200_set_refit_mode(int mode) { g_wb_refit_mode = (mode == 1 || mode == 2) ? mode : 0; }
int wp_b200_get_refit_mode(void) { return g_wb_refit_mode; }
// process-wide switches of measured-and-rejected alternatives that stay in the library as tested code paths:
// "small_nodes" (builder: Karras-style pass for small distinct-key nodes before the merge; -1 = environment default)
int wp_b200_set_experiment(const char* name, int value)
{
    if (name && !strcmp(name, "small_nodes")) {
        g_wb_small_nodes = value < 0 ? -1 : (value ? 1 : 0);
        return 1;
    }
    set_error("Warp error: unknown experiment %s", name ? name : "(null)");
    return 0;
}

void wp_b200_set_auto_reference_layout(int enable) { g_auto_reference_layout = enable ? 1 : 0; }
int wp_b200_get_auto_reference_layout(void) { return g_auto_reference_layout; }

// per-object options: "refit_mode" (0 auto / 1 atomic / 2 wavefront), "query_order" (0 input / 1 curve / 2 auto),
// "ray_order" (0 / 1), "auto_reference_layout" (0 / 1); value -1 = follow the process-wide default again.
// "morton_bits" is read-only (fixed at creation).
int wp_b200_bvh_set_option(uint64_t id, const char* name, int value)
{
    BvhState* s = find_tree(id);
    if (!s || !name) {
        set_error("Warp error: invalid id");
        return 0;
    }
    if (!strcmp(name, "refit_mode") && value >= -1 && value <= 2)
        s->refit_mode = value;
    else if (!strcmp(name, "query_order") && value >= -1 && value <= 2)
        s->query_order = value;
    else if (!strcmp(name, "ray_order") && value >= -1 && value <= 1)
        s->ray_order = value;
    else if (!strcmp(name, "auto_reference_layout") && value >= 0 && value <= 1)
        s->auto_reference_layout = value;
    else {
        set_error("Warp error: unknown option or value out of range: %s = %d", name, value);
        return 0;
    }
    return 1;
}

int wp_b200_bvh_get_option(uint64_t id, const char* name, int* value)
{
    BvhState* s = find_tree(id);
    if (!s || !name || !value) {
        set_error("Warp error: invalid id");
        return 0;
    }
    if (!strcmp(name, "refit_mode"))
        *value = s->refit_mode;
    else if (!strcmp(name, "query_order"))
        *value = s->query_order;
    else if (!strcmp(name, "ray_order"))
        *value = s->ray_order;
    else if (!strcmp(name, "auto_reference_layout"))
        *value = s->auto_reference_layout;
    else if (!strcmp(name, "morton_bits"))
        *value = s->morton_bits;
    else {
        set_error("Warp error: unknown option %s", name);
        return 0;
    }
    return 1;
}

int wp_b200_mesh_rebuild_device(uint64_t id)
{
    MeshState* m = nullptr;
    find_tree(id, &m);
    if (!m) {
        set_error("Warp error: invalid mesh id");
        return 0;
    }
    DeviceGuard g(m->bvh.device);
    const char* err = wb_build(m->bvh, current_stream(m->bvh.device));
    if (err) {
        set_error("Warp error: mesh rebuild failed: %s", err);
        return 0;
    }
    if (m->bvh.auto_reference_layout)
        return sync_reference_layout(&m->bvh, m);
    return 1;
}

int wp_mesh_set_points_device(uint64_t id, wp_array_t points)
{
    g_error[0] = 0;  // the drop-in stub reads this library's message only when the last routed call failed
    MeshState* m = nullptr;
    find_tree(id, &m);
    if (!m) {
        set_error("Warp error: invalid mesh id");
        return 0;
    }
    if (points.ndim != 1 || points.shape[0] != m->points_shape0) {
        set_error("Warp error: new points input for wp_mesh_set_points_device does not match the original points shape");
        return 0;
    }
    DeviceGuard g(m->bvh.device);
    m->points_data = points.data;
    m->bvh.points = (const float*)points.data;
    if (!upload_desc(m, nullptr))
        return 0;
    return wp_mesh_refit_device(id);
}

void wp_mesh_set_velocities_device(uint64_t id, wp_array_t velocities)
{
    g_error[0] = 0;  // the drop-in stub reads this library's message only when the last routed call failed
    MeshState* m = nullptr;
    find_tree(id, &m);
    if (!m) {
        set_error("Warp error: invalid mesh id");
        return;
    }
    if (velocities.ndim != 1 || velocities.shape[0] != m->velocities_shape0) {
        set_error("Warp error: new velocities input for wp_mesh_set_velocities_device does not match the original "
                  "velocities shape");
        return;
    }
    DeviceGuard g(m->bvh.device);
    m->velocities_data = velocities.data;
    upload_desc(m, nullptr);
}

// ------------------------------------------------------------------------------------------------
// queries (device pointers)
// ------------------------------------------------------------------------------------------------
void wp_b200_query_stats_enable(int enable) { t_stats_enabled = enable != 0; }

void wp_b200_set_morton_bits(int bits) { g_morton_bits = (bits == 63) ? 63 : 30; }
int wp_b200_get_morton_bits(void) { return g_morton_bits; }
void wp_b200_set_query_order(int mode) { g_query_order = mode; }
void wp_b200_set_ray_order(int mode) { g_ray_order = mode ? 1 : 0; }
int wp_b200_get_ray_order(void) { return g_ray_order; }
int wp_b200_get_query_order(void) { return g_query_order; }

// kernel timing of the point / ray traversal kernels (see KernelTimerScope): enable, run, read = (sum of the launch
// durations in ms, number of launches) since the last read; reading waits for the timed launches to finish
void wp_b200_kernel_timing_enable(int enable)
{
    std::lock_guard<std::mutex> g(g_lock);
    g_ktimer.enabled = enable != 0;
}

void wp_b200_kernel_timing_read(float* total_ms, int* launches)
{
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> todo;
    {
        std::lock_guard<std::mutex> g(g_lock);
        todo.swap(g_ktimer.pending);
    }
    float sum = 0.f;
    int count = 0;
    for (auto& ev : todo) {
        float ms = 0.f;
        if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess)
            sum += ms, ++count;
        else
            cudaGetLastError();
    }
    {
        std::lock_guard<std::mutex> g(g_lock);
        for (auto& ev : todo)
            g_ktimer.pool.push_back(ev.first), g_ktimer.pool.push_back(ev.second);
    }
    if (total_ms)
        *total_ms = sum;
    if (launches)
        *launches = count;
}

void wp_b200_query_stats_read(unsigned long long* pair_fetches, unsigned long long* tri_fetches)
{
    unsigned long long h[2] = { 0, 0 };
    if (g_stats_dev) {
        cudaStreamSynchronize(current_stream(current_device()));
        cudaMemcpy(h, g_stats_dev, sizeof(h), cudaMemcpyDeviceToHost);
    }
    if (pair_fetches)
        *pair_fetches = h[0];
    if (tri_fetches)
        *tri_fetches = h[1];
}

static int zero_point_outputs(int64_t n, uint8_t* result, float* sign, int32_t* face, float* u, float* v, cudaStream_t st)
{
    bool ok = check(cudaMemsetAsync(result, 0, (size_t)n, st), "memset");
    if (sign)
        ok = ok && check(cudaMemsetAsync(sign, 0, 4 * (size_t)n, st), "memset");
    ok = ok && check(cudaMemsetAsync(face, 0, 4 * (size_t)n, st), "memset");
    ok = ok && check(cudaMemsetAsync(u, 0, 4 * (size_t)n, st), "memset");
    ok = ok && check(cudaMemsetAsync(v, 0, 4 * (size_t)n, st), "memset");
    return ok ? 1 : 0;
}

static int query_point_on(MeshState* m, const float* points, int64_t n, float max_dist, int with_sign, uint8_t* result,
                          float* sign, int32_t* face, float* u, float* v, cudaStream_t st, int lane = 0)
{
    if (n <= 0)
        return 1;
    if (m->bvh.n == 0)
        return zero_point_outputs(n, result, sign, face, u, v, st);
    const int* perm = nullptr;
    uint4* packed = nullptr;
    float* sorted_pts = nullptr;
    const int order = m->bvh.query_order >= 0 ? m->bvh.query_order : g_query_order;
    if ((order == 1 || (order == 2 && n >= 32768)) && n < (1ll << 30)) {
        OrderScratch& ws = order_scratch(m->bvh.device, st);
#ifndef WB_SIGN_HILBERT
#define WB_SIGN_HILBERT 0
#endif
        const char* oerr = wb_morton_order(ws, points, n, st, WB_SIGN_HILBERT || !with_sign);
        if (oerr) {
            set_error("Warp error: query ordering failed: %s", oerr);
            return 0;
        }
        perm = ws.idx, packed = ws.packed, sorted_pts = ws.sorted_pts;
    }
    const char* err;
    {
        KernelTimerScope timed(st);
        L2PersistScope l2(m->bvh, st);
        // the host-buffer lanes keep the direct SoA stores: with packed records + unpack pass the pinned-buffer pipeline
        // measured 533 instead of 572 M queries/s end to end (the device-resident batch gains 1.5 % from them)
        const int mode = lane ? (query_mode() & ~4) : query_mode();
        err = wb_query_point(make_view(m->bvh), points, perm, n, max_dist, with_sign, result, sign, face, u, v,
                             stats_buffer(), st, mode, packed, sorted_pts);
    }
    if (err) {
        set_error("Warp error: mesh point query failed: %s", err);
        return 0;
    }
    return 1;
}

static int query_ray_on(MeshState* m, const float* starts, const float* dirs, int64_t n, float max_t, uint8_t* result,
                        float* sign, int32_t* face, float* t, float* u, float* v, float* normal, cudaStream_t st, int lane = 0,
                        const int32_t* roots = nullptr)
{
    if (n <= 0)
        return 1;
    if (m->bvh.n == 0) {
        int ok = zero_point_outputs(n, result, sign, face, u, v, st);
        ok = ok && check(cudaMemsetAsync(t, 0, 4 * (size_t)n, st), "memset");
        ok = ok && check(cudaMemsetAsync(normal, 0, 12 * (size_t)n, st), "memset");
        return ok;
    }
    const int* perm = nullptr;
    const int order = m->bvh.ray_order >= 0 ? m->bvh.ray_order : g_ray_order;
    if (order == 1 && n < (1ll << 30)) {  // opt-in: coherent batches (primary rays) gain nothing from it
        OrderScratch& ws = order_scratch(m->bvh.device, st);
        const char* oerr = wb_ray_order(ws, starts, dirs, n, st);
        if (oerr) {
            set_error("Warp error: ray ordering failed: %s", oerr);
            return 0;
        }
        perm = ws.idx;
    }
    const char* err;
    {
        KernelTimerScope timed(st);
        L2PersistScope l2(m->bvh, st);
        err = wb_query_ray(make_view(m->bvh), starts, dirs, perm, roots, n, max_t, result, sign, face, t, u, v, normal,
                           stats_buffer(), st);
    }
    if (err) {
        set_error("Warp error: mesh ray query failed: %s", err);
        return 0;
    }
    return 1;
}

int wp_b200_mesh_query_point_no_sign(uint64_t id, const float* points, int64_t n, float max_dist, uint8_t* result,
                                     int32_t* face, float* u, float* v)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    DeviceGuard g(m->bvh.device);
    return query_point_on(m, points, n, max_dist, 0, result, nullptr, face, u, v, current_stream(m->bvh.device));
}

int wp_b200_mesh_query_point(uint64_t id, const float* points, int64_t n, float max_dist, uint8_t* result, float* sign,
                             int32_t* face, float* u, float* v)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    DeviceGuard g(m->bvh.device);
    return query_point_on(m, points, n, max_dist, 1, result, sign, face, u, v, current_stream(m->bvh.device));
}

int wp_b200_mesh_query_point_sign_parity(uint64_t id, const float* points, int64_t n, float max_dist, int n_sample,
                                         float perturbation_scale, uint8_t* result, float* sign, int32_t* face, float* u, float* v)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    DeviceGuard g(m->bvh.device);
    cudaStream_t st = current_stream(m->bvh.device);
    if (n <= 0)
        return 1;
    // the closest point is mesh_query_point_no_sign's (same traversal, mesh.h:324-477); then the parity vote
    if (!query_point_on(m, points, n, max_dist, 0, result, nullptr, face, u, v, st))
        return 0;
    if (m->bvh.n == 0)
        return check(cudaMemsetAsync(sign, 0, 4 * (size_t)n, st), "memset");
    const char* err = wb_sign_parity(make_view(m->bvh), points, n, n_sample, perturbation_scale, result, sign, st);
    if (err) {
        set_error("Warp error: mesh parity sign failed: %s", err);
        return 0;
    }
    return 1;
}

int wp_b200_mesh_query_point_sign_normal(uint64_t id, const float* points, int64_t n, float max_dist, float epsilon,
                                         uint8_t* result, float* sign, int32_t* face, float* u, float* v)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    DeviceGuard g(m->bvh.device);
    cudaStream_t st = current_stream(m->bvh.device);
    if (n <= 0)
        return 1;
    if (m->bvh.n == 0)
        return zero_point_outputs(n, result, sign, face, u, v, st);
    const int* perm = nullptr;
    const int order = m->bvh.query_order >= 0 ? m->bvh.query_order : g_query_order;
    if ((order == 1 || (order == 2 && n >= 32768)) && n < (1ll << 30)) {
        OrderScratch& ws = order_scratch(m->bvh.device, st);
        const char* oerr = wb_morton_order(ws, points, n, st, true);
        if (oerr) {
            set_error("Warp error: query ordering failed: %s", oerr);
            return 0;
        }
        perm = ws.idx;
    }
    float* avg = &((wp_b200_mesh_desc*)m->dev_desc)->average_edge_length;
    const char* err = wb_query_point_sign_normal(make_view(m->bvh), m->bvh.points, m->bvh.indices, points, perm, n, max_dist,
                                                 epsilon, m->bvh.edge_partials, avg, result, sign, face, u, v, st);
    if (err) {
        set_error("Warp error: mesh normal-sign query failed: %s", err);
        return 0;
    }
    return 1;
}

int wp_b200_mesh_average_edge_length(uint64_t id, float* out)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    DeviceGuard g(m->bvh.device);
    cudaStream_t st = current_stream(m->bvh.device);
    float* avg = &((wp_b200_mesh_desc*)m->dev_desc)->average_edge_length;
    *out = 0.0f;
    if (m->bvh.n == 0)
        return 1;
    const char* err = wb_query_point_sign_normal(make_view(m->bvh), m->bvh.points, m->bvh.indices, nullptr, nullptr, 0, 0.f,
                                                 0.f, m->bvh.edge_partials, avg, nullptr, nullptr, nullptr, nullptr,
                                                 nullptr, st);
    if (err) {
        set_error("Warp error: average edge length failed: %s", err);
        return 0;
    }
    return check(cudaMemcpyAsync(out, avg, sizeof(float), cudaMemcpyDeviceToHost, st), "memcpy")
        && check(cudaStreamSynchronize(st), "synchronize");
}

int wp_b200_mesh_query_ray(uint64_t id, const float* starts, const float* dirs, int64_t n, float max_t, uint8_t* result,
                           float* sign, int32_t* face, float* t, float* u, float* v, float* normal, const int32_t* roots)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    DeviceGuard g(m->bvh.device);
    return query_ray_on(m, starts, dirs, n, max_t, result, sign, face, t, u, v, normal, current_stream(m->bvh.device), 0, roots);
}

int wp_b200_mesh_query_ray_anyhit(uint64_t id, const float* starts, const float* dirs, int64_t n, float max_t,
                                  uint8_t* result, const int32_t* roots)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    DeviceGuard g(m->bvh.device);
    cudaStream_t st = current_stream(m->bvh.device);
    if (n <= 0)
        return 1;
    if (m->bvh.n == 0)
        return check(cudaMemsetAsync(result, 0, (size_t)n, st), "memset");
    const char* err = wb_query_ray_anyhit(make_view(m->bvh), starts, dirs, roots, n, max_t, result, st);
    if (err) {
        set_error("Warp error: mesh any-hit query failed: %s", err);
        return 0;
    }
    return 1;
}

int wp_b200_mesh_query_ray_count_intersections(uint64_t id, const float* starts, const float* dirs, int64_t n,
                                               int32_t* counts, const int32_t* roots)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    DeviceGuard g(m->bvh.device);
    cudaStream_t st = current_stream(m->bvh.device);
    if (n <= 0)
        return 1;
    if (m->bvh.n == 0)
        return check(cudaMemsetAsync(counts, 0, 4 * (size_t)n, st), "memset");
    const char* err = wb_query_ray_count(make_view(m->bvh), starts, dirs, roots, n, counts, st);
    if (err) {
        set_error("Warp error: mesh intersection count failed: %s", err);
        return 0;
    }
    return 1;
}

static int mesh_eval(uint64_t id, int velocity, const int32_t* face, const float* u, const float* v, int64_t n, float* out)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    DeviceGuard g(m->bvh.device);
    cudaStream_t st = current_stream(m->bvh.device);
    if (n <= 0)
        return 1;
    const uint64_t attr = velocity ? m->velocities_data : m->points_data;
    if (!attr || !m->indices_data)  // mesh.h:2771-2772, 2791-2792: vec3()
        return check(cudaMemsetAsync(out, 0, 12 * (size_t)n, st), "memset");
    const char* err = wb_mesh_eval((const float*)attr, (const int*)m->indices_data, face, u, v, n, out, st);
    if (err) {
        set_error("Warp error: mesh evaluation failed: %s", err);
        return 0;
    }
    return 1;
}

int wp_b200_mesh_eval_position(uint64_t id, const int32_t* face, const float* u, const float* v, int64_t n, float* out)
{
    return mesh_eval(id, 0, face, u, v, n, out);
}

int wp_b200_mesh_eval_velocity(uint64_t id, const int32_t* face, const float* u, const float* v, int64_t n, float* out)
{
    return mesh_eval(id, 1, face, u, v, n, out);
}

int wp_b200_mesh_eval_face_normal(uint64_t id, const int32_t* face, int64_t n, float* out)
{
    return wp_b200_mesh_eval_face_normal_masked(id, face, nullptr, n, out);
}

int wp_b200_mesh_eval_face_normal_masked(uint64_t id, const int32_t* face, const uint8_t* mask, int64_t n, float* out)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    DeviceGuard g(m->bvh.device);
    cudaStream_t st = current_stream(m->bvh.device);
    if (n <= 0)
        return 1;
    if (!m->points_data || !m->indices_data)  // mesh.h:2874-2875: vec3()
        return check(cudaMemsetAsync(out, 0, 12 * (size_t)n, st), "memset");
    const char* err = wb_mesh_face_normal((const float*)m->points_data, (const int*)m->indices_data, face, mask, n, out, st);
    if (err) {
        set_error("Warp error: mesh face normal failed: %s", err);
        return 0;
    }
    return 1;
}

int wp_b200_mesh_query_furthest_point_no_sign(uint64_t id, const float* points, int64_t n, float min_dist, uint8_t* result,
                                              int32_t* face, float* u, float* v)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    DeviceGuard g(m->bvh.device);
    cudaStream_t st = current_stream(m->bvh.device);
    if (n <= 0)
        return 1;
    if (m->bvh.n == 0)
        return zero_point_outputs(n, result, nullptr, face, u, v, st);
    const char* err = wb_query_furthest(make_view(m->bvh), points, n, min_dist, result, face, u, v, st);
    if (err) {
        set_error("Warp error: mesh furthest-point query failed: %s", err);
        return 0;
    }
    return 1;
}

// ------------------------------------------------------------------------------------------------
// queries (host buffers): chunked, two lanes so the copies of one chunk overlap the traversal of
// the other.  Fully asynchronous only when the caller's buffers are pinned (wp_alloc_pinned).
// ------------------------------------------------------------------------------------------------
static int point_host(uint64_t id, const float* points, int64_t n, float max_dist, int with_sign, uint8_t* result,
                      float* sign, int32_t* face, float* u, float* v)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    const int dev = m->bvh.device;
    DeviceGuard g(dev);
    cudaStreamSynchronize(current_stream(dev));  // the tree must be complete before the lanes read it
    const int64_t chunk = n < host_chunk(n) ? (n > 0 ? n : 1) : host_chunk(n);
    // a short first and last chunk (1 M) keep the copies nobody can overlap -- the first upload, the last download -- small
    const int64_t edge = (!host_chunk_forced() && n > (1ll << 23)) ? (1ll << 20) : 0;
    const int64_t middle = n - 2 * edge;
    const int64_t mid_parts = (middle + (6ll << 20) - 1) / (6ll << 20);  // equal middle chunks of at most 6 M
    const int64_t mid_chunk = edge ? (((middle + mid_parts - 1) / mid_parts + 1023) & ~(int64_t)1023) : chunk;
    const int64_t cap = edge ? (mid_chunk > edge ? mid_chunk : edge) : chunk;
    const size_t o_pts = 0, o_res = o_pts + align256(12 * cap), o_sign = o_res + align256(cap),
                 o_face = o_sign + align256(4 * cap), o_u = o_face + align256(4 * cap),
                 o_v = o_u + align256(4 * cap), total = o_v + align256(4 * cap);
    int ok = 1;
    for (int64_t base = 0, k = 0; base < n && ok; ++k) {
        HostLane& l = g_lanes[dev][k & 1];
        if (!lane_reserve(l, total))
            return 0;
        int64_t c = edge ? ((base == 0 || n - base <= edge) ? edge : mid_chunk) : chunk;
        if (edge && base > 0 && n - base > edge && n - base - c < edge)
            c = n - base - edge;  // the last middle chunk stops where the tail chunk starts
        if (c > n - base)
            c = n - base;
        char* b = (char*)l.buf;
        ok = ok && check(cudaMemcpyAsync(b + o_pts, points + 3 * base, 12 * c, cudaMemcpyHostToDevice, l.stream), "h2d");
        ok = ok && query_point_on(m, (const float*)(b + o_pts), c, max_dist, with_sign, (uint8_t*)(b + o_res),
                                  with_sign ? (float*)(b + o_sign) : nullptr, (int32_t*)(b + o_face), (float*)(b + o_u),
                                  (float*)(b + o_v), l.stream, 1 + (int)(k & 1));
        ok = ok && check(cudaMemcpyAsync(result + base, b + o_res, c, cudaMemcpyDeviceToHost, l.stream), "d2h");
        if (with_sign)
            ok = ok && check(cudaMemcpyAsync(sign + base, b + o_sign, 4 * c, cudaMemcpyDeviceToHost, l.stream), "d2h");
        ok = ok && check(cudaMemcpyAsync(face + base, b + o_face, 4 * c, cudaMemcpyDeviceToHost, l.stream), "d2h");
        ok = ok && check(cudaMemcpyAsync(u + base, b + o_u, 4 * c, cudaMemcpyDeviceToHost, l.stream), "d2h");
        ok = ok && check(cudaMemcpyAsync(v + base, b + o_v, 4 * c, cudaMemcpyDeviceToHost, l.stream), "d2h");
        base += c;
    }
    for (int k = 0; k < 2; ++k)
        if (g_lanes[dev][k].stream)
            ok = check(cudaStreamSynchronize(g_lanes[dev][k].stream), "lane synchronize") && ok;
    return ok;
}

int wp_b200_mesh_query_point_no_sign_host(uint64_t id, const float* points, int64_t n, float max_dist, uint8_t* result,
                                          int32_t* face, float* u, float* v)
{
    return point_host(id, points, n, max_dist, 0, result, nullptr, face, u, v);
}

int wp_b200_mesh_query_point_host(uint64_t id, const float* points, int64_t n, float max_dist, uint8_t* result,
                                  float* sign, int32_t* face, float* u, float* v)
{
    return point_host(id, points, n, max_dist, 1, result, sign, face, u, v);
}

int wp_b200_mesh_query_ray_host(uint64_t id, const float* starts, const float* dirs, int64_t n, float max_t,
                                uint8_t* result, float* sign, int32_t* face, float* t, float* u, float* v, float* normal)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    const int dev = m->bvh.device;
    DeviceGuard g(dev);
    cudaStreamSynchronize(current_stream(dev));
    const int64_t chunk = n < host_chunk(n) ? (n > 0 ? n : 1) : host_chunk(n);
    const size_t o_s = 0, o_d = o_s + align256(12 * chunk), o_res = o_d + align256(12 * chunk),
                 o_sign = o_res + align256(chunk), o_face = o_sign + align256(4 * chunk),
                 o_t = o_face + align256(4 * chunk), o_u = o_t + align256(4 * chunk), o_v = o_u + align256(4 * chunk),
                 o_n = o_v + align256(4 * chunk), total = o_n + align256(12 * chunk);
    int ok = 1;
    for (int64_t base = 0, k = 0; base < n && ok; base += chunk, ++k) {
        HostLane& l = g_lanes[dev][k & 1];
        if (!lane_reserve(l, total))
            return 0;
        const int64_t c = (n - base < chunk) ? n - base : chunk;
        char* b = (char*)l.buf;
        ok = ok && check(cudaMemcpyAsync(b + o_s, starts + 3 * base, 12 * c, cudaMemcpyHostToDevice, l.stream), "h2d");
        ok = ok && check(cudaMemcpyAsync(b + o_d, dirs + 3 * base, 12 * c, cudaMemcpyHostToDevice, l.stream), "h2d");
        ok = ok && query_ray_on(m, (const float*)(b + o_s), (const float*)(b + o_d), c, max_t, (uint8_t*)(b + o_res),
                                (float*)(b + o_sign), (int32_t*)(b + o_face), (float*)(b + o_t), (float*)(b + o_u),
                                (float*)(b + o_v), (float*)(b + o_n), l.stream, 1 + (int)(k & 1));
        ok = ok && check(cudaMemcpyAsync(result + base, b + o_res, c, cudaMemcpyDeviceToHost, l.stream), "d2h");
        ok = ok && check(cudaMemcpyAsync(sign + base, b + o_sign, 4 * c, cudaMemcpyDeviceToHost, l.stream), "d2h");
        ok = ok && check(cudaMemcpyAsync(face + base, b + o_face, 4 * c, cudaMemcpyDeviceToHost, l.stream), "d2h");
        ok = ok && check(cudaMemcpyAsync(t + base, b + o_t, 4 * c, cudaMemcpyDeviceToHost, l.stream), "d2h");
        ok = ok && check(cudaMemcpyAsync(u + base, b + o_u, 4 * c, cudaMemcpyDeviceToHost, l.stream), "d2h");
        ok = ok && check(cudaMemcpyAsync(v + base, b + o_v, 4 * c, cudaMemcpyDeviceToHost, l.stream), "d2h");
        ok = ok && check(cudaMemcpyAsync(normal + 3 * base, b + o_n, 12 * c, cudaMemcpyDeviceToHost, l.stream), "d2h");
    }
    for (int k = 0; k < 2; ++k)
        if (g_lanes[dev][k].stream)
            ok = check(cudaStreamSynchronize(g_lanes[dev][k].stream), "lane synchronize") && ok;
    return ok;
}

// ------------------------------------------------------------------------------------------------
// generic BVH queries (wp.Bvh): count pass, device scan, fill pass
// ------------------------------------------------------------------------------------------------
static long long* g_scan_scratch[64] = {};
static size_t g_scan_scratch_words[64] = {};
static std::mutex g_scan_lock;

static int bvh_query_common(uint64_t id, int ray, const float* qa, const float* qb, const int32_t* roots, int64_t n,
                            float max_dist, int32_t* counts, const int32_t* offsets, int32_t* indices, bool want_mesh = false,
                            const float* radii = nullptr)
{
    MeshState* ms = nullptr;
    BvhState* s = find_tree(id, &ms);
    if (!s || (ms != nullptr) != want_mesh) {
        set_error(want_mesh ? "Warp error: invalid mesh id" : "Warp error: invalid BVH id (generic queries take a wp.Bvh)");
        return 0;
    }
    DeviceGuard g(s->device);
    cudaStream_t st = current_stream(s->device);
    if (n <= 0)
        return 1;
    if (s->n == 0) {
        if (!offsets)
            return check(cudaMemsetAsync(counts, 0, 4 * (size_t)n, st), "memset") ? 1 : 0;
        return 1;
    }
    const char* err = wb_bvh_query(make_view(*s), want_mesh ? nullptr : s->item_lowers, want_mesh ? nullptr : s->item_uppers,
                                   ray, qa, qb, radii, roots, n, max_dist, counts, offsets, indices, st);
    if (err) {
        set_error("Warp error: BVH query failed: %s", err);
        return 0;
    }
    return 1;
}

int wp_b200_bvh_query_aabb_count(uint64_t id, const float* lowers, const float* uppers, const int32_t* roots, int64_t n,
                                 int32_t* counts)
{
    return bvh_query_common(id, 0, lowers, uppers, roots, n, 0.f, counts, nullptr, nullptr);
}
int wp_b200_bvh_query_aabb_fill(uint64_t id, const float* lowers, const float* uppers, const int32_t* roots, int64_t n,
                                const int32_t* offsets, int32_t* indices)
{
    return bvh_query_common(id, 0, lowers, uppers, roots, n, 0.f, nullptr, offsets, indices);
}
int wp_b200_bvh_query_ray_count(uint64_t id, const float* starts, const float* dirs, const int32_t* roots, int64_t n,
                                float max_dist, int32_t* counts)
{
    return bvh_query_common(id, 1, starts, dirs, roots, n, max_dist, counts, nullptr, nullptr);
}
int wp_b200_bvh_query_ray_fill(uint64_t id, const float* starts, const float* dirs, const int32_t* roots, int64_t n,
                               float max_dist, const int32_t* offsets, int32_t* indices)
{
    return bvh_query_common(id, 1, starts, dirs, roots, n, max_dist, nullptr, offsets, indices);
}

int wp_b200_bvh_query_sphere_count(uint64_t id, const float* centers, const float* radii, const int32_t* roots, int64_t n,
                                   int32_t* counts)
{
    return bvh_query_common(id, 2, centers, centers, roots, n, 0.f, counts, nullptr, nullptr, false, radii);
}
int wp_b200_bvh_query_sphere_fill(uint64_t id, const float* centers, const float* radii, const int32_t* roots, int64_t n,
                                  const int32_t* offsets, int32_t* indices)
{
    return bvh_query_common(id, 2, centers, centers, roots, n, 0.f, nullptr, offsets, indices, false, radii);
}
int wp_b200_bvh_query_capsule_count(uint64_t id, const float* starts, const float* dirs, const float* radii,
                                    const int32_t* roots, int64_t n, float max_dist, int32_t* counts)
{
    return bvh_query_common(id, 3, starts, dirs, roots, n, max_dist, counts, nullptr, nullptr, false, radii);
}
int wp_b200_bvh_query_capsule_fill(uint64_t id, const float* starts, const float* dirs, const float* radii,
                                   const int32_t* roots, int64_t n, float max_dist, const int32_t* offsets, int32_t* indices)
{
    return bvh_query_common(id, 3, starts, dirs, roots, n, max_dist, nullptr, offsets, indices, false, radii);
}

int wp_b200_mesh_query_aabb_count(uint64_t id, const float* lowers, const float* uppers, int64_t n, int32_t* counts)
{
    return bvh_query_common(id, 0, lowers, uppers, nullptr, n, 0.f, counts, nullptr, nullptr, true);
}
int wp_b200_mesh_query_aabb_fill(uint64_t id, const float* lowers, const float* uppers, int64_t n, const int32_t* offsets,
                                 int32_t* indices)
{
    return bvh_query_common(id, 0, lowers, uppers, nullptr, n, 0.f, nullptr, offsets, indices, true);
}

static int mesh_query_sphere_common(uint64_t id, const float* centers, const float* radii, int64_t n, int32_t* counts,
                                    const int32_t* offsets, int32_t* indices)
{
    MeshState* m = query_mesh(id);
    if (!m)
        return 0;
    DeviceGuard g(m->bvh.device);
    cudaStream_t st = current_stream(m->bvh.device);
    if (n <= 0)
        return 1;
    if (m->bvh.n == 0)
        return offsets ? 1 : (check(cudaMemsetAsync(counts, 0, 4 * (size_t)n, st), "memset") ? 1 : 0);
    const char* err = wb_mesh_query_sphere(make_view(m->bvh), centers, radii, n, counts, offsets, indices, st);
    if (err) {
        set_error("Warp error: mesh sphere query failed: %s", err);
        return 0;
    }
    return 1;
}

int wp_b200_mesh_query_sphere_count(uint64_t id, const float* centers, const float* radii, int64_t n, int32_t* counts)
{
    return mesh_query_sphere_common(id, centers, radii, n, counts, nullptr, nullptr);
}
int wp_b200_mesh_query_sphere_fill(uint64_t id, const float* centers, const float* radii, int64_t n, const int32_t* offsets,
                                   int32_t* indices)
{
    return mesh_query_sphere_common(id, centers, radii, n, nullptr, offsets, indices);
}

// bvh_get_group_root (bvh.h:376-390) for a batch of group ids: reference node index of the subtree that holds
// exactly the items of the group, -1 when the group does not occur
int wp_b200_bvh_get_group_root(uint64_t id, const int32_t* group_ids, int64_t n, int32_t* roots)
{
    MeshState* ms = nullptr;
    BvhState* s = find_tree(id, &ms);
    if (!s) {
        set_error("Warp error: invalid id");
        return 0;
    }
    DeviceGuard g(s->device);
    cudaStream_t st = current_stream(s->device);
    if (n <= 0)
        return 1;
    if (s->n == 0)
        return check(cudaMemsetAsync(roots, 0xff, 4 * (size_t)n, st), "memset") ? 1 : 0;
    const char* err = wb_group_roots(make_view(*s), s->groups ? s->keys : nullptr, group_ids, n, roots, st);
    if (err) {
        set_error("Warp error: group root lookup failed: %s", err);
        return 0;
    }
    return 1;
}

int wp_b200_exclusive_scan_i32(const int32_t* counts, int32_t* offsets, int64_t n)
{
    // the scan runs where its buffers live (a Bvh on cuda:1 hands in cuda:1 pointers while the calling thread's current
    // device may be cuda:0), on that device's current stream -- the stream the count pass was enqueued on
    int dev = current_device();
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, counts) == cudaSuccess && attr.type == cudaMemoryTypeDevice)
        dev = attr.device;
    else
        cudaGetLastError();
    if (dev < 0 || dev >= 64) {
        set_error("Warp error: scan buffers live on an unsupported device ordinal %d", dev);
        return 0;
    }
    DeviceGuard g(dev);
    std::lock_guard<std::mutex> scan_lock(g_scan_lock);
    const size_t words = (size_t)(n > 0 ? (n + 2047) / 2048 : 0) + 2;
    if (words > g_scan_scratch_words[dev]) {
        if (g_scan_scratch[dev])
            cudaFree(g_scan_scratch[dev]);
        g_scan_scratch[dev] = nullptr, g_scan_scratch_words[dev] = 0;
        if (!check(cudaMalloc(&g_scan_scratch[dev], 8 * words), "scan scratch"))
            return 0;
        g_scan_scratch_words[dev] = words;
    }
    const char* err = wb_exclusive_scan(counts, offsets, n, g_scan_scratch[dev], current_stream(dev));
    if (err) {
        set_error("Warp error: scan failed: %s", err);
        return 0;
    }
    return 1;
}

// ------------------------------------------------------------------------------------------------
// reference-layout mirror: node_lowers / node_uppers / node_parents / root (bvh.h:161-207), and for meshes the
// per-triangle lowers / uppers (mesh.cu:279-280, 310-313) and average_edge_length -- everything an unmodified Warp
// kernel reads through the id.  The caller holds a DeviceGuard for the tree's device.
// ------------------------------------------------------------------------------------------------
static int sync_reference_layout(BvhState* s, MeshState* m)
{
    if (s->host_built) {
        set_error("Warp error: the reference-layout mirror is not available for trees of the host constructors (sah / median): "
                  "their nodes are not numbered like the reference's (host_build.cu)");
        return 0;
    }
    const bool first = s->ref_lowers == nullptr;
    const char* err = wb_export_reference_layout(*s, current_stream(s->device));
    if (err) {
        set_error("Warp error: reference-layout export failed: %s", err);
        return 0;
    }
    if (first && s->n > 0 && !upload_desc(m, m ? nullptr : s))
        return 0;
    if (m && s->n > 0) {
        // a Warp kernel that calls mesh_query_point_sign_normal reads wp::Mesh::average_edge_length through the id
        // (mesh.h:889): refresh it together with the node arrays
        float* avg = &((wp_b200_mesh_desc*)m->dev_desc)->average_edge_length;
        err = wb_query_point_sign_normal(make_view(*s), s->points, s->indices, nullptr, nullptr, 0, 0.f, 0.f, s->edge_partials,
                                         avg, nullptr, nullptr, nullptr, nullptr, nullptr, current_stream(s->device));
        if (err) {
            set_error("Warp error: average edge length failed: %s", err);
            return 0;
        }
    }
    return 1;
}

// ------------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------------
int wp_b200_bvh_info(uint64_t id, wp_b200_bvh_info_t* info)
{
    MeshState* m = nullptr;
    BvhState* s = find_tree(id, &m);
    if (!s) {
        set_error("Warp error: invalid id");
        return 0;
    }
    DeviceGuard g(s->device);
    memset(info, 0, sizeof(*info));
    info->num_items = s->n, info->leaf_size = s->leaf_size, info->max_nodes = s->n > 0 ? 2 * s->n - 1 : 0;
    info->root = -1;
    if (s->n == 0)
        return 1;
    TreeHeader h;
    cudaStreamSynchronize(current_stream(s->device));
    if (!check(cudaMemcpy(&h, s->header, sizeof(h), cudaMemcpyDeviceToHost), "header download"))
        return 0;
    info->root = (int)(h.root_ref & WB_IDX_MASK);
    info->height = h.height, info->deep = h.deep;
    info->key_bits = 8 * s->key_bytes;
    for (int k = 0; k < 3; ++k)
        info->total_lower[k] = h.total_lo[k], info->total_upper[k] = h.total_hi[k], info->inv_edges[k] = h.inv_edges[k];
    return 1;
}

int wp_b200_bvh_sync_reference_layout(uint64_t id)
{
    MeshState* m = nullptr;
    BvhState* s = find_tree(id, &m);
    if (!s) {
        set_error("Warp error: invalid id");
        return 0;
    }
    DeviceGuard g(s->device);
    return sync_reference_layout(s, m);
}

// EXPERIMENT (not used by any build): parents (reference node indices, -1 for the root) of the n - 1
// internal nodes, recomputed from the tree's sorted keys by the dependency-free k_topology kernel, `reps` times; returns
// the average kernel time in microseconds, -1 on error, -2 when a run of equal keys was too long for the prototype
float wp_b200_experiment_parallel_topology(uint64_t id, int32_t* parents_out, int reps)
{
    BvhState* s = find_tree(id);
    if (!s || s->n < 2 || reps < 1)
        return -1.0f;
    DeviceGuard g(s->device);
    cudaStream_t st = current_stream(s->device);
    int* fail = nullptr;
    if (cudaMalloc(&fail, sizeof(int)) != cudaSuccess)
        return -1.0f;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const char* err = wb_experiment_topology(*s, parents_out, fail, st);  // warm-up
    cudaEventRecord(e0, st);
    for (int k = 0; k < reps && !err; ++k)
        err = wb_experiment_topology(*s, parents_out, fail, st);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    int h_fail = 0;
    cudaMemcpy(&h_fail, fail, sizeof(int), cudaMemcpyDeviceToHost);
    cudaEventDestroy(e0), cudaEventDestroy(e1), cudaFree(fail);
    if (err)
        return -1.0f;
    return h_fail ? -2.0f : 1000.0f * ms / (float)reps;
}

int wp_b200_bvh_download(uint64_t id, void* keys, int32_t* primitive_indices, void* node_lowers, void* node_uppers,
                         int32_t* node_parents, int32_t* root)
{
    const bool want_nodes = node_lowers || node_uppers || node_parents || root;
    if (want_nodes && !wp_b200_bvh_sync_reference_layout(id))  // (keys / primitive_indices alone need no mirror)
        return 0;
    BvhState* s = find_tree(id);
    if (!s) {
        set_error("Warp error: invalid id");
        return 0;
    }
    if (s->n == 0)
        return 1;
    DeviceGuard g(s->device);
    cudaStream_t st = current_stream(s->device);
    const size_t n = (size_t)s->n, mx = 2 * n - 1;
    bool ok = check(cudaStreamSynchronize(st), "synchronize");
    if (keys)
        ok = ok && check(cudaMemcpy(keys, s->keys, (size_t)s->key_bytes * n, cudaMemcpyDeviceToHost), "download");
    if (primitive_indices)
        ok = ok && check(cudaMemcpy(primitive_indices, s->prim, 4 * n, cudaMemcpyDeviceToHost), "download");
    if (node_lowers)
        ok = ok && check(cudaMemcpy(node_lowers, s->ref_lowers, 16 * mx, cudaMemcpyDeviceToHost), "download");
    if (node_uppers)
        ok = ok && check(cudaMemcpy(node_uppers, s->ref_uppers, 16 * mx, cudaMemcpyDeviceToHost), "download");
    if (node_parents)
        ok = ok && check(cudaMemcpy(node_parents, s->ref_parents, 4 * mx, cudaMemcpyDeviceToHost), "download");
    if (root)
        ok = ok && check(cudaMemcpy(root, s->ref_root, 4, cudaMemcpyDeviceToHost), "download");
    return ok ? 1 : 0;
}

}  // extern "C"
