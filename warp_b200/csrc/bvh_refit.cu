// Bottom-up AABB refit over the sibling-pair layout.
//
// Behavioural contract = warp/native/bvh.cu:42-144 + mesh.cu:368-407: every visible node's box
// becomes the exact union (fminf/fmaxf) of the boxes of the items below it.  The reference starts
// one thread per ORIGINAL leaf and climbs through the muted nodes under each packed leaf; here one
// thread per sorted position gathers its triangle straight from the vertex array (no lowers/uppers
// round trip, no edge-length / scan passes) and refreshes the packed-triangle cache the queries read;
// the first thread of each VISIBLE packed leaf (found through pos_parent[], written by the builder)
// unions the leaf's <= leaf_size boxes and the merge pass climbs with one atomic arrival counter per
// internal node.  The counters are never cleared: they are
// even after a build and every refit adds exactly 2, so "second to arrive" == odd old value.
#include "state.h"
#include "merge.cuh"  // wb_store_box, wb_climb

#include <cstdlib>
#include <cstring>

namespace {

constexpr int BT = 256;

// item (triangle / box) at sorted position k: bounds, and for meshes the refreshed packed-triangle record
template <class Src, bool WRITE>
__device__ __forceinline__ void refit_item(const Src& src, const int* __restrict__ prim, float4* __restrict__ tris, int k,
                                           float3& lo, float3& hi)
{
    const int item = __ldg(prim + k);
    if constexpr (Src::kIsMesh) {
        float3 p, q, r;
        src.tri(item, p, q, r);
        lo = wb_min3(wb_min3(p, q), r);
        hi = wb_max3(wb_max3(p, q), r);
        if (WRITE) {
            const float3 e0 = wb_sub(q, p), e1 = wb_sub(r, p), e2 = wb_sub(r, q);
            const float3 nrm = wb_cross(e0, e1);
            const float area2 = sqrtf(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z);
            const bool sliver = area2 / (wb_dot(e0, e0) + wb_dot(e1, e1) + wb_dot(e2, e2)) < 1.e-6f;
            float4* t = tris + 3 * (size_t)k;
            t[0] = make_float4(p.x, p.y, p.z, q.x);
            t[1] = make_float4(q.y, q.z, r.x, r.y);
            t[2] = make_float4(r.z, __int_as_float(item), __uint_as_float(sliver ? WB_TRI_SLIVER : 0u), 0.f);
        }
    } else {
        src.bounds(item, lo, hi);
    }
}

// One thread per sorted POSITION gathers its item and refreshes its triangle record (coalesced, like the
// builder's k_leaves); the item boxes are staged in shared memory, and the thread at the first position of each
// visible leaf unions its leaf's boxes from there.  Items of a leaf that continue past the block's last
// position are gathered again by the leaf's thread (their records are written by the block that owns them).
template <class Src>
__global__ void __launch_bounds__(BT)
k_refit_leaves(Src src, int n, const int* __restrict__ prim, const int* __restrict__ pos_parent, NodeRec* pairs,
               float4* __restrict__ tris, TreeHeader* hdr)
{
    __shared__ float sbox[6][BT];
    const int tid = threadIdx.x;
    const int i = blockIdx.x * BT + tid;
    float3 lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    int parent = WB_NO_PARENT;
    if (i < n) {
        parent = __ldg(pos_parent + i);
        refit_item<Src, true>(src, prim, tris, i, lo, hi);
        sbox[0][tid] = lo.x, sbox[1][tid] = lo.y, sbox[2][tid] = lo.z;
        sbox[3][tid] = hi.x, sbox[4][tid] = hi.y, sbox[5][tid] = hi.z;
    }
    __syncthreads();
    if (parent == WB_NO_PARENT)
        return;

    const int s = parent - n;  // internal slot of the parent
    int side = 0;
    int count;
    if (parent == WB_ROOT_PARENT) {
        count = n;
    } else if (i <= s) {
        side = 0;
        count = s - i + 1;
    } else {
        side = 1;
        count = (int)pairs[2 * (size_t)s + 1].aux - s;
    }

    int k = 1;
    for (; k < count && tid + k < BT; ++k) {
        lo = wb_min3(lo, make_float3(sbox[0][tid + k], sbox[1][tid + k], sbox[2][tid + k]));
        hi = wb_max3(hi, make_float3(sbox[3][tid + k], sbox[4][tid + k], sbox[5][tid + k]));
    }
    for (; k < count; ++k) {
        float3 a, b;
        refit_item<Src, false>(src, prim, tris, i + k, a, b);
        lo = wb_min3(lo, a);
        hi = wb_max3(hi, b);
    }

    if (parent == WB_ROOT_PARENT) {
        hdr->lx = lo.x, hdr->ly = lo.y, hdr->lz = lo.z;
        hdr->hx = hi.x, hdr->hy = hi.y, hdr->hz = hi.z;
        return;
    }

    // the merge pass (k_merge<true>, merge.cuh) unions these leaf boxes bottom-up
    wb_store_box(pairs + 2 * (size_t)s + side, lo, hi);
}

// ------------------------------------------------------------------------------------------------
// Wavefront refit (large trees): the tree is static between builds, so the visiting order is planned once
// (wb_refit_plan, bvh_build.cu) and a refit is three streaming-friendly kernels with no arrival counter below the
// block level:
//  K1 k_refit_leaves  gather, refresh the triangle records, union each visible leaf's boxes into its record;
//  K2 k_refit_levels  a block owns WB_WAVE_BP sorted positions and walks ITS internal nodes (sorted by height)
//                     level by level with boxes in shared memory -- one slot per position: a left child sits in
//                     the slot of its last position, a right child in the slot of its first, and the next user of
//                     a slot is always an ancestor -- writing each union into its record in the parent's pair;
//  K3 k_refit_climb   nodes / leaves whose parent spans blocks (a static list) announce themselves on the GLOBAL
//                     counters and climb the spine above the blocks (wb_climb, merge.cuh).
// Boxes are exact min / max unions, so the result is bit-identical to any other visiting order.
// Measured (B200, refit ms, wavefront / atomic): 1.3 M tris 0.134 / 0.129, 4 M 0.262 / 0.287, 40 M 2.17 / 2.36
// (63-bit keys 1.87 / 2.05), 100 M 63-bit 4.50 / 4.94 (0.64 / 0.58 of the HBM roofline); a single kernel doing all
// three stages was slower (7.5 ms at 100 M: blocks sit on their SM while a few threads climb).
// ------------------------------------------------------------------------------------------------
constexpr int WT = 256;                  // threads per block
constexpr int WITEMS = WB_WAVE_BP / WT;  // positions (and at most plan entries) per thread

// K2: levels inside the blocks (leaf boxes come from the records k_refit_leaves just wrote)
__global__ void __launch_bounds__(WT)
k_refit_levels(MergeArgs<uint32_t> a, const uint8_t* __restrict__ unit_flags, const uint32_t* __restrict__ plan_keys,
               const int* __restrict__ plan_nodes, const uint32_t* __restrict__ plan_dst, const int* __restrict__ plan_begin,
               const int* __restrict__ plan_end)
{
    __shared__ float sbox[6][WB_WAVE_BP];
    __shared__ unsigned snext[2];
    const int begin = plan_begin[blockIdx.x], end = plan_end[blockIdx.x];
    if (begin >= end)
        return;  // no internal node lies inside this block
    const int n = a.n;
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * WB_WAVE_BP;
    const int b1 = min(b0 + WB_WAVE_BP - 1, n - 1);
#pragma unroll
    for (int j = 0; j < WITEMS; ++j) {
        const int k = j * WT + tid, i = b0 + k;
        if (i > b1)
            continue;
        const int parent = __ldg(a.pos_parent + i);
        if (parent < 0 || unit_flags[i])
            continue;
        const int s = parent - n;
        const float4* r4 = reinterpret_cast<const float4*>(a.pairs + 2 * (size_t)s + (i <= s ? 0 : 1));
        const float4 r0 = r4[0], r1 = r4[1];
        const int q = (i <= s) ? s - b0 : k;  // slot: last position of a left child, first of a right child
        sbox[0][q] = r0.x, sbox[1][q] = r0.y, sbox[2][q] = r0.z;
        sbox[3][q] = r1.x, sbox[4][q] = r1.y, sbox[5][q] = r1.z;
    }
    if (tid < 2)
        snext[tid] = 0xffffffffu;
    uint32_t ekey[WITEMS], edst[WITEMS];
    int enode[WITEMS];
#pragma unroll
    for (int j = 0; j < WITEMS; ++j) {
        const int e = begin + j * WT + tid;
        ekey[j] = 0xffffffffu;
        if (e < end) {
            ekey[j] = plan_keys[e] & ((1u << WB_PLAN_HEIGHT_BITS) - 1u);
            enode[j] = plan_nodes[e];
            edst[j] = plan_dst[e];
        }
    }
    __syncthreads();
    for (int it = 0;; ++it) {
        unsigned mine = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < WITEMS; ++j)
            mine = min(mine, ekey[j]);
        mine = __reduce_min_sync(0xffffffffu, mine);
        if ((tid & 31) == 0 && mine != 0xffffffffu)
            atomicMin(&snext[it & 1], mine);
        __syncthreads();
        const unsigned level = snext[it & 1];
        if (tid == 0)
            snext[(it + 1) & 1] = 0xffffffffu;
        if (level == 0xffffffffu)
            break;
#pragma unroll
        for (int j = 0; j < WITEMS; ++j) {
            if (ekey[j] != level)
                continue;
            ekey[j] = 0xffffffffu;
            const int ql = enode[j] - b0, qr = ql + 1;
            const float3 lo = wb_min3(make_float3(sbox[0][ql], sbox[1][ql], sbox[2][ql]), make_float3(sbox[0][qr], sbox[1][qr], sbox[2][qr]));
            const float3 hi = wb_max3(make_float3(sbox[3][ql], sbox[4][ql], sbox[5][ql]), make_float3(sbox[3][qr], sbox[4][qr], sbox[5][qr]));
            if (edst[j] == WB_PLAN_DST_ROOT) {
                a.hdr->lx = lo.x, a.hdr->ly = lo.y, a.hdr->lz = lo.z;
                a.hdr->hx = hi.x, a.hdr->hy = hi.y, a.hdr->hz = hi.z;
            } else {
                const uint32_t d = edst[j] & 0x7fffffffu;
                wb_store_box(a.pairs + d, lo, hi);
                if (!(edst[j] & 0x80000000u)) {
                    const int q = (int)(d >> 1) + (int)(d & 1u) - b0;
                    sbox[0][q] = lo.x, sbox[1][q] = lo.y, sbox[2][q] = lo.z;
                    sbox[3][q] = hi.x, sbox[4][q] = hi.y, sbox[5][q] = hi.z;
                }
            }
        }
        __syncthreads();
    }
}

// K3: the spine above the blocks -- one thread per node / leaf whose parent spans blocks (a static list)
__global__ void __launch_bounds__(BT)
k_refit_climb(MergeArgs<uint32_t> a, const uint32_t* __restrict__ top, const int* __restrict__ ntop)
{
    const int m = *ntop;
    for (int i = blockIdx.x * BT + threadIdx.x; i < m; i += gridDim.x * BT) {
        const uint32_t d = top[i];
        wb_climb<true, uint32_t, false>(a, (int)(d >> 1), (int)(d & 1u), 0u);
    }
}

}  // namespace

#define WB_CUDA_TRY(expr)                  \
    do {                                   \
        cudaError_t _e = (expr);           \
        if (_e != cudaSuccess)             \
            return cudaGetErrorString(_e); \
    } while (0)

int g_wb_refit_mode = 0;  // default of new trees: 0 auto, 1 atomic counters, 2 wavefront (wp_b200_set_refit_mode)

const char* wb_refit(BvhState& s, cudaStream_t stream)
{
    if (s.n <= 0)
        return nullptr;
    // auto: wavefront from 2 M items up (see the table above); wp_b200_bvh_set_option(id, "refit_mode") / WARP_B200_REFIT force one
    static const char* mode_env = getenv("WARP_B200_REFIT");
    static const int env_mode = !mode_env ? 0 : (strcmp(mode_env, "atomic") == 0 ? 1 : 2);
    const int base_mode = s.refit_mode >= 0 ? s.refit_mode : g_wb_refit_mode;
    const int mode = base_mode ? base_mode : env_mode;
    // (a host-built tree has no heights / key ranges for the plan: it always takes the counter climb)
    const bool wave = !s.host_built && s.n >= 2 && s.n < (1 << 30) && (mode == 0 ? s.n >= (1 << 21) : mode == 2);
    if (wave && !s.plan_valid)
        if (const char* e = wb_refit_plan(s, stream))
            return e;
    const int grid = wb_div_up(s.n, BT);
    if (s.is_mesh)
        k_refit_leaves<<<grid, BT, 0, stream>>>(MeshSource { s.points, s.indices }, s.n, s.prim, s.pos_parent, s.pairs,
                                                s.tris, s.header);
    else
        k_refit_leaves<<<grid, BT, 0, stream>>>(BoxSource { s.item_lowers, s.item_uppers }, s.n, s.prim, s.pos_parent,
                                                s.pairs, s.tris, s.header);
    WB_CUDA_TRY(cudaGetLastError());
    if (!wave)
        return wb_refit_merge(s, stream);
    const MergeArgs<uint32_t> ma { s.n, s.leaf_size, nullptr, s.prim, s.pairs, s.parent_int, s.pos_parent, s.counters, s.header,
                                   nullptr, nullptr };
    k_refit_levels<<<wb_div_up(s.n, WB_WAVE_BP), WT, 0, stream>>>(ma, s.unit_flags, s.plan_keys, s.plan_nodes, s.plan_dst,
                                                                  s.plan_begin, s.plan_end);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.device);
    k_refit_climb<<<min(sms * 8, wb_div_up(s.n, BT)), BT, 0, stream>>>(ma, s.plan_top, s.plan_ntop);
    WB_CUDA_TRY(cudaGetLastError());
    return nullptr;
}
