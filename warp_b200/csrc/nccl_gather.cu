// NCCL plumbing for the sharded-query path: one communicator per process (one process per GPU),
// all-gather of SoA result arrays over NVLink / NVSwitch.  libnccl.so.2 is dlopen()ed so the
// library loads (and every single-GPU entry point works) on machines without NCCL.
#include "../../include/warp_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

struct NcclUniqueId {
    char internal[128];
};
typedef void* NcclComm;
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(NcclComm);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef const char* (*fn_error_string)(int);
typedef int (*fn_send)(const void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_recv)(void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_group)(void);

constexpr int kNcclInt8 = 0, kNcclFloat32 = 7, kNcclMax = 2;

void* g_lib = nullptr;
fn_get_unique_id p_get_unique_id = nullptr;
fn_comm_init_rank p_comm_init_rank = nullptr;
fn_comm_destroy p_comm_destroy = nullptr;
fn_all_gather p_all_gather = nullptr;
fn_all_reduce p_all_reduce = nullptr;
fn_error_string p_error_string = nullptr;
fn_send p_send = nullptr;
fn_recv p_recv = nullptr;
fn_group p_group_start = nullptr, p_group_end = nullptr;
NcclComm g_comm = nullptr;
int g_world = 0;
cudaEvent_t g_fork_event = nullptr, g_join_event = nullptr;
constexpr int kMarks = 8;
cudaEvent_t g_mark[kMarks] = {};
cudaStream_t g_comm_stream = nullptr;
char g_nccl_error[512] = "";

int fail(const char* what, int code)
{
    snprintf(g_nccl_error, sizeof(g_nccl_error), "NCCL error in %s: %s", what,
             (p_error_string && code) ? p_error_string(code) : "library not loaded");
    fprintf(stderr, "%s\n", g_nccl_error);
    return 0;
}

}  // namespace

extern "C" {

int wp_b200_nccl_load(const char* path)
{
    if (g_lib)
        return 1;
    g_lib = dlopen(path ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!g_lib) {
        snprintf(g_nccl_error, sizeof(g_nccl_error), "cannot load NCCL: %s", dlerror());
        return 0;
    }
    p_get_unique_id = (fn_get_unique_id)dlsym(g_lib, "ncclGetUniqueId");
    p_comm_init_rank = (fn_comm_init_rank)dlsym(g_lib, "ncclCommInitRank");
    p_comm_destroy = (fn_comm_destroy)dlsym(g_lib, "ncclCommDestroy");
    p_all_gather = (fn_all_gather)dlsym(g_lib, "ncclAllGather");
    p_all_reduce = (fn_all_reduce)dlsym(g_lib, "ncclAllReduce");
    p_error_string = (fn_error_string)dlsym(g_lib, "ncclGetErrorString");
    p_send = (fn_send)dlsym(g_lib, "ncclSend");
    p_recv = (fn_recv)dlsym(g_lib, "ncclRecv");
    p_group_start = (fn_group)dlsym(g_lib, "ncclGroupStart");
    p_group_end = (fn_group)dlsym(g_lib, "ncclGroupEnd");
    if (!p_get_unique_id || !p_comm_init_rank || !p_comm_destroy || !p_all_gather || !p_all_reduce) {
        snprintf(g_nccl_error, sizeof(g_nccl_error), "NCCL library lacks required symbols");
        return 0;
    }
    return 1;
}

int wp_b200_nccl_unique_id(void* id128)
{
    if (!g_lib && !wp_b200_nccl_load(nullptr))
        return 0;
    NcclUniqueId id;
    const int rc = p_get_unique_id(&id);
    if (rc)
        return fail("ncclGetUniqueId", rc);
    memcpy(id128, &id, sizeof(id));
    return 1;
}

int wp_b200_nccl_init(const void* id128, int world_size, int rank)
{
    if (!g_lib && !wp_b200_nccl_load(nullptr))
        return 0;
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    const int rc = p_comm_init_rank(&g_comm, world_size, id, rank);
    if (rc)
        return fail("ncclCommInitRank", rc);
    g_world = world_size;
    if (!g_comm_stream) {
        // HIGHEST priority: a kernel on this stream (NCCL's, the 4-byte fences of the peer-memory gather, the ray normals)
        // must get SM slots while a traversal grid of 131 072 blocks is still being dispatched on the compute stream --
        // at equal priority the block scheduler serves the earlier launch first and the "overlapped" gather of batch k
        // only starts in the tail of traversal k + 1 (measured: step = traversal + gather instead of their maximum)
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        const char* e = getenv("WARP_B200_COMM_PRIORITY");  // 0 = default priority (A/B)
        cudaStreamCreateWithPriority(&g_comm_stream, cudaStreamNonBlocking, (e && atoi(e) == 0) ? lo : hi);
    }
    if (!g_fork_event) {
        cudaEventCreateWithFlags(&g_fork_event, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&g_join_event, cudaEventDisableTiming);
    }
    return 1;
}

// gathers on the library's current stream of the current device (same stream the queries ran on)
int wp_b200_nccl_allgather(const void* send, void* recv, size_t bytes_per_rank)
{
    if (!g_comm)
        return fail("allgather (communicator not initialised)", 0);
    cudaStream_t st = (cudaStream_t)wp_cuda_context_get_stream(nullptr);
    const int rc = p_all_gather(send, recv, bytes_per_rank, kNcclInt8, g_comm, st);
    return rc ? fail("ncclAllGather", rc) : 1;
}

// Pipelined gather: part [offset, offset + part_bytes) of every rank's shard goes to recv + r * shard_stride + offset
// on every rank (the final rank-major layout, no staging), as one group of ncclSend / ncclRecv on the library's
// COMMUNICATION stream, so that it runs under the traversal of the next part.  wp_b200_nccl_fork() makes the
// communication stream wait for what has been enqueued on the current stream (the part's query), and
// wp_b200_nccl_join() makes the current stream wait for the communication stream (end of the step).
int wp_b200_nccl_allgather_part(const void* send, void* recv, size_t part_bytes, size_t shard_stride_bytes, size_t offset_bytes)
{
    if (!g_comm)
        return fail("allgather_part (communicator not initialised)", 0);
    if (!p_send || !p_recv || !p_group_start || !p_group_end)
        return fail("allgather_part (ncclSend / ncclRecv not available)", 0);
    if (part_bytes == 0)
        return 1;
    int rc = p_group_start();
    for (int r = 0; r < g_world && !rc; ++r) {
        rc = p_send(send, part_bytes, kNcclInt8, r, g_comm, g_comm_stream);
        if (!rc)
            rc = p_recv((char*)recv + (size_t)r * shard_stride_bytes + offset_bytes, part_bytes, kNcclInt8, r, g_comm,
                        g_comm_stream);
    }
    const int rc2 = p_group_end();
    return (rc || rc2) ? fail("ncclSend/ncclRecv group", rc ? rc : rc2) : 1;
}

// Several fields of one result set in ONE NCCL launch (ncclGroupStart / End around the per-field all-gathers): what the
// sharded queries gather after a traversal.  on_comm_stream = 1 enqueues on the library's communication stream (after a
// wp_b200_nccl_fork()) so that the gather of one batch runs under the traversal of the next.
int wp_b200_nccl_allgather_multi(const void* const* send, void* const* recv, const size_t* bytes_per_rank, int count,
                                 int on_comm_stream)
{
    if (!g_comm)
        return fail("allgather_multi (communicator not initialised)", 0);
    if (!p_group_start || !p_group_end)
        return fail("allgather_multi (ncclGroupStart / End not available)", 0);
    cudaStream_t st = on_comm_stream ? g_comm_stream : (cudaStream_t)wp_cuda_context_get_stream(nullptr);
    int rc = p_group_start();
    for (int k = 0; k < count && !rc; ++k)
        if (bytes_per_rank[k])
            rc = p_all_gather(send[k], recv[k], bytes_per_rank[k], kNcclInt8, g_comm, st);
    const int rc2 = p_group_end();
    return (rc || rc2) ? fail("grouped ncclAllGather", rc ? rc : rc2) : 1;
}

// mark(k): remember the current tail of the communication stream; wait_mark(k): the CURRENT stream waits for that point
// (a buffer that gather k read may be overwritten afterwards) -- finer than wp_b200_nccl_join(), which waits for everything
int wp_b200_nccl_mark(int k)
{
    if (!g_comm || k < 0 || k >= kMarks)
        return fail("mark (communicator not initialised or bad index)", 0);
    if (!g_mark[k] && cudaEventCreateWithFlags(&g_mark[k], cudaEventDisableTiming) != cudaSuccess)
        return 0;
    return cudaEventRecord(g_mark[k], g_comm_stream) == cudaSuccess;
}

int wp_b200_nccl_wait_mark(int k)
{
    if (!g_comm || k < 0 || k >= kMarks)
        return fail("wait_mark (communicator not initialised or bad index)", 0);
    if (!g_mark[k])
        return 1;  // never recorded: nothing to wait for
    cudaStream_t st = (cudaStream_t)wp_cuda_context_get_stream(nullptr);
    return cudaStreamWaitEvent(st, g_mark[k], 0) == cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
// Peer-memory all-gather: every rank PUSHES its shard straight into every peer's result buffer with
// cudaMemcpyAsync over NVLink (copy engines, no SM), instead of an SM-driven NCCL kernel that competes with the
// traversal it is supposed to hide under.  One process per GPU, so peer buffers are reached through CUDA IPC handles
// (exchanged once per buffer set by the caller, with NCCL).  Ordering across ranks uses two 4-byte NCCL all-reduces on
// the communication stream: the first ("everyone has forked") keeps a fast rank from overwriting a buffer a slow rank
// is still reading, the second ("everyone has pushed") tells each rank that all its incoming shards have landed.
// ------------------------------------------------------------------------------------------------
// handle72 = the 64-byte cudaIpcMemHandle_t of the allocation that holds device_ptr + the 8-byte offset of device_ptr
// inside it (the runtime may carve small buffers out of a larger block; an IPC handle always maps the whole block)
int wp_b200_ipc_get_handle(void* device_ptr, void* handle72)
{
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, device_ptr) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t");
    unsigned long long offset = 0;
    typedef int (*fn_range)(unsigned long long*, size_t*, unsigned long long);
    static fn_range get_range = [] {
        void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        return lib ? (fn_range)dlsym(lib, "cuMemGetAddressRange_v2") : (fn_range) nullptr;
    }();
    if (get_range) {
        unsigned long long base = 0;
        size_t size = 0;
        if (get_range(&base, &size, (unsigned long long)(uintptr_t)device_ptr) == 0 && base)
            offset = (unsigned long long)(uintptr_t)device_ptr - base;
    }
    memcpy(handle72, &h, sizeof(h));
    memcpy((char*)handle72 + 64, &offset, 8);
    return 1;
}

// maps a peer's allocation into this process; returns the peer buffer's address here (base + offset), NULL on failure
void* wp_b200_ipc_open_handle(const void* handle72)
{
    cudaIpcMemHandle_t h;
    unsigned long long offset = 0;
    memcpy(&h, handle72, sizeof(h));
    memcpy(&offset, (const char*)handle72 + 64, 8);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return (char*)p + offset;
}

// takes what wp_b200_ipc_open_handle returned and the same handle (for the offset)
void wp_b200_ipc_close_handle(void* peer_ptr, const void* handle72)
{
    if (!peer_ptr)
        return;
    unsigned long long offset = 0;
    memcpy(&offset, (const char*)handle72 + 64, 8);
    cudaIpcCloseMemHandle((char*)peer_ptr - offset);
}

// peer_recv[r * count + k] = field k of rank r's gathered buffer as mapped into this process (NULL for r == own rank:
// the own shard is copied locally into own_recv[k]); send[k] = this rank's shard of field k, bytes_per_rank[k] bytes.
// Enqueued on the communication stream (after a wp_b200_nccl_fork()).
int wp_b200_p2p_allgather_multi(const void* const* send, void* const* own_recv, void* const* peer_recv,
                                const size_t* bytes_per_rank, int count, int rank)
{
    if (!g_comm)
        return fail("p2p_allgather (communicator not initialised)", 0);
    static float* token = nullptr;
    if (!token) {
        cudaMalloc(&token, 2 * sizeof(float));
        cudaMemset(token, 0, 2 * sizeof(float));
    }
    int rc = p_all_reduce(token, token, 1, kNcclFloat32, kNcclMax, g_comm, g_comm_stream);  // everyone has forked
    if (rc)
        return fail("ncclAllReduce (p2p pre-sync)", rc);
    // one stream per peer, so that the pushes to different peers run on different copy engines at the same time
    // (a single stream serialises them: measured 270 GB/s of NVLink egress per rank instead of what the links allow)
    constexpr int kPeerStreams = 16;
    static cudaStream_t peer_stream[kPeerStreams] = {};
    static cudaEvent_t peer_done[kPeerStreams] = {};
    static cudaEvent_t go = nullptr;
    if (!go)
        cudaEventCreateWithFlags(&go, cudaEventDisableTiming);
    bool ok = cudaEventRecord(go, g_comm_stream) == cudaSuccess;
    for (int k = 0; k < count && ok; ++k) {
        if (!bytes_per_rank[k])
            continue;
        ok = cudaMemcpyAsync((char*)own_recv[k] + (size_t)rank * bytes_per_rank[k], send[k], bytes_per_rank[k],
                             cudaMemcpyDeviceToDevice, g_comm_stream) == cudaSuccess;
    }
    // peers in a rotated order, so that at any moment every rank is mostly receiving from a different sender
    for (int step = 1; step < g_world && ok; ++step) {
        const int r = (rank + step) % g_world;
        const int lane = (step - 1) % kPeerStreams;
        if (!peer_stream[lane]) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            const char* e = getenv("WARP_B200_COMM_PRIORITY");
            cudaStreamCreateWithPriority(&peer_stream[lane], cudaStreamNonBlocking, (e && atoi(e) == 0) ? lo : hi);
            cudaEventCreateWithFlags(&peer_done[lane], cudaEventDisableTiming);
        }
        ok = cudaStreamWaitEvent(peer_stream[lane], go, 0) == cudaSuccess;
        for (int k = 0; k < count && ok; ++k) {
            if (!bytes_per_rank[k])
                continue;
            char* dst = (char*)peer_recv[(size_t)r * count + k];
            if (!dst)
                return fail("p2p_allgather (peer buffer not mapped)", 0);
            ok = cudaMemcpyAsync(dst + (size_t)rank * bytes_per_rank[k], send[k], bytes_per_rank[k], cudaMemcpyDeviceToDevice,
                                 peer_stream[lane]) == cudaSuccess;
        }
        ok = ok && cudaEventRecord(peer_done[lane], peer_stream[lane]) == cudaSuccess
            && cudaStreamWaitEvent(g_comm_stream, peer_done[lane], 0) == cudaSuccess;
    }
    if (!ok) {
        cudaGetLastError();
        return fail("p2p_allgather (cudaMemcpyAsync to a peer)", 0);
    }
    rc = p_all_reduce(token + 1, token + 1, 1, kNcclFloat32, kNcclMax, g_comm, g_comm_stream);  // everyone has pushed
    return rc ? fail("ncclAllReduce (p2p post-sync)", rc) : 1;
}

// the library's communication stream (created by wp_b200_nccl_init): callers that want a kernel of theirs to run in
// communication order (e.g. the ray normals recomputed from gathered faces) make it current around the launch
void* wp_b200_nccl_comm_stream(void) { return g_comm_stream; }

int wp_b200_nccl_fork(void)
{
    if (!g_comm)
        return fail("fork (communicator not initialised)", 0);
    cudaStream_t st = (cudaStream_t)wp_cuda_context_get_stream(nullptr);
    return cudaEventRecord(g_fork_event, st) == cudaSuccess && cudaStreamWaitEvent(g_comm_stream, g_fork_event, 0) == cudaSuccess;
}

int wp_b200_nccl_join(void)
{
    if (!g_comm)
        return fail("join (communicator not initialised)", 0);
    cudaStream_t st = (cudaStream_t)wp_cuda_context_get_stream(nullptr);
    return cudaEventRecord(g_join_event, g_comm_stream) == cudaSuccess && cudaStreamWaitEvent(st, g_join_event, 0) == cudaSuccess;
}

int wp_b200_nccl_allreduce_max_f32(float* inout_device, size_t count)
{
    if (!g_comm)
        return fail("allreduce (communicator not initialised)", 0);
    cudaStream_t st = (cudaStream_t)wp_cuda_context_get_stream(nullptr);
    const int rc = p_all_reduce(inout_device, inout_device, count, kNcclFloat32, kNcclMax, g_comm, st);
    return rc ? fail("ncclAllReduce", rc) : 1;
}

int wp_b200_nccl_barrier(void)
{
    static float* token = nullptr;
    if (!token) {
        cudaMalloc(&token, sizeof(float));
        cudaMemset(token, 0, sizeof(float));
    }
    if (!wp_b200_nccl_allreduce_max_f32(token, 1))
        return 0;
    return cudaStreamSynchronize((cudaStream_t)wp_cuda_context_get_stream(nullptr)) == cudaSuccess;
}

void wp_b200_nccl_destroy(void)
{
    if (g_comm && p_comm_destroy)
        p_comm_destroy(g_comm);
    g_comm = nullptr;
}

}  // extern "C"
