// NCCL plumbing for the sharded-query path: one communicator per process (one process per GPU),
// all-gather of SoA result arrays over NVLink / NVSwitch.  libnccl.so.2 is dlopen()ed so the
// library loads (and every single-GPU entry point works) on machines without NCCL.
#include "../../include/warp_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
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

constexpr int kNcclInt8 = 0, kNcclFloat32 = 7, kNcclMax = 2;

void* g_lib = nullptr;
fn_get_unique_id p_get_unique_id = nullptr;
fn_comm_init_rank p_comm_init_rank = nullptr;
fn_comm_destroy p_comm_destroy = nullptr;
fn_all_gather p_all_gather = nullptr;
fn_all_reduce p_all_reduce = nullptr;
fn_error_string p_error_string = nullptr;
NcclComm g_comm = nullptr;
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
    if (!g_comm_stream)
        cudaStreamCreateWithFlags(&g_comm_stream, cudaStreamNonBlocking);
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
