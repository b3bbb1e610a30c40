// LBVH construction for sm_100a: scene bounds -> Morton keys + digit histograms -> onesweep LSD
// radix sort of (key, item) -> leaf records -> chunked bottom-up merge directly into the
// sibling-pair node layout (common.cuh) -> depth-rule fix-up -> (optional) export to the
// reference's two-array layout.
//
// Behavioural contract = warp/native/bvh.cu:184-613 (reference LBVH): identical Morton keys,
// identical sorted order (any stable sort), identical parent choice / tie break / packed-leaf
// marking.  What differs is how the bytes move: 9 launches instead of ~20, 32-bit keys and 4 digit
// passes instead of 64-bit keys and 8 when there are no groups, no lowers/uppers round trip for
// meshes, no delta / range passes (deltas are recomputed from the sorted keys, ranges travel in
// the node records), height tracked in the arrival counters so the per-node depth walk of
// mark_packed_leaf_nodes (bvh.cu:402-443) only runs for trees that are actually >= 32 deep.
//
// Key flavours (template KeyT / GROUPED, see merge.cuh): uint32 30-bit code (parity mode, default),
// uint64 group<<32|code (grouped trees, bvh.cu:205-209), uint64 63-bit code (quality option for very
// large meshes; NOT a parity mode -- the reference only ever produces the 30-bit code).
#include "state.h"
#include "merge.cuh"
#include "order.h"

#include <cub/device/device_radix_sort.cuh>  // WARP_B200_SORT=cub cross-check path only

#include <cstdlib>
#include <cstring>

namespace {

constexpr int BT = 256;  // threads per block for the streaming kernels

#define WB_CUDA_TRY(expr)                  \
    do {                                   \
        cudaError_t _e = (expr);           \
        if (_e != cudaSuccess)             \
            return cudaGetErrorString(_e); \
    } while (0)

__global__ void k_iota(int* out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = i;
}

// ---------------------------------------------------------------------------------------------
// block reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// reduce (lo, hi) over the block; result valid in thread 0
__device__ __forceinline__ void block_minmax(float3& lo, float3& hi, float (*sm)[6])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    lo.x = warp_min(lo.x), lo.y = warp_min(lo.y), lo.z = warp_min(lo.z);
    hi.x = warp_max(hi.x), hi.y = warp_max(hi.y), hi.z = warp_max(hi.z);
    __syncthreads();
    if (lane == 0) {
        sm[warp][0] = lo.x, sm[warp][1] = lo.y, sm[warp][2] = lo.z;
        sm[warp][3] = hi.x, sm[warp][4] = hi.y, sm[warp][5] = hi.z;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < BT / 32; ++w) {
            lo.x = fminf(lo.x, sm[w][0]), lo.y = fminf(lo.y, sm[w][1]), lo.z = fminf(lo.z, sm[w][2]);
            hi.x = fmaxf(hi.x, sm[w][3]), hi.y = fmaxf(hi.y, sm[w][4]), hi.z = fmaxf(hi.z, sm[w][5]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1: per-block partial scene bounds (bvh.cu:449-479).  No cross-block handshake here: the few
// hundred partials are reduced redundantly by every block of K2 (a 14 KB L2 read), which removes
// the same-address ticket atomics and the extra kernel.  Also clears the digit histograms and
// tile tickets used by the following kernels.
// ---------------------------------------------------------------------------------------------
template <class Src>
__global__ void __launch_bounds__(BT)
k_scene_bounds(Src src, int n, float* __restrict__ partials, unsigned* __restrict__ tickets,
               uint32_t* __restrict__ ghist, uint4* __restrict__ clear = nullptr, size_t clear_words4 = 0)
{
    __shared__ float sm[BT / 32][6];

    // the sort's look-back words start from zero: cleared here by all blocks (16 bytes per store) instead of a memset launch
    for (size_t k = (size_t)blockIdx.x * BT + threadIdx.x; k < clear_words4; k += (size_t)gridDim.x * BT)
        clear[k] = make_uint4(0u, 0u, 0u, 0u);

    if (blockIdx.x == 0) {
        for (int k = threadIdx.x; k < 8 * 256; k += BT)
            ghist[k] = 0;
        if (threadIdx.x < 16)
            tickets[threadIdx.x] = 0;  // per-pass tile tickets
    }

    float3 lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    const int stride = gridDim.x * BT;
    int i = blockIdx.x * BT + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {  // four independent gathers in flight
        float3 a0, b0, a1, b1, a2, b2, a3, b3;
        src.bounds(i, a0, b0);
        src.bounds(i + stride, a1, b1);
        src.bounds(i + 2 * stride, a2, b2);
        src.bounds(i + 3 * stride, a3, b3);
        lo = wb_min3(wb_min3(lo, wb_min3(a0, a1)), wb_min3(a2, a3));
        hi = wb_max3(wb_max3(hi, wb_max3(b0, b1)), wb_max3(b2, b3));
    }
    for (; i < n; i += stride) {
        float3 a, b;
        src.bounds(i, a, b);
        lo = wb_min3(lo, a);
        hi = wb_max3(hi, b);
    }
    block_minmax(lo, hi, sm);
    if (threadIdx.x == 0) {
        float* p = partials + 6 * blockIdx.x;
        p[0] = lo.x, p[1] = lo.y, p[2] = lo.z, p[3] = hi.x, p[4] = hi.y, p[5] = hi.z;
    }
}

// reduce the K1 partials inside a block; every thread returns the scene bounds
__device__ __forceinline__ void reduce_partials(const float* __restrict__ partials, int num_partials, float (*sm)[6],
                                                float3& lo, float3& hi)
{
    __shared__ float total[6];
    lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int b = threadIdx.x; b < num_partials; b += BT) {
        const float* p = partials + 6 * b;
        lo = wb_min3(lo, make_float3(p[0], p[1], p[2]));
        hi = wb_max3(hi, make_float3(p[3], p[4], p[5]));
    }
    block_minmax(lo, hi, sm);
    if (threadIdx.x == 0)
        total[0] = lo.x, total[1] = lo.y, total[2] = lo.z, total[3] = hi.x, total[4] = hi.y, total[5] = hi.z;
    __syncthreads();
    lo = make_float3(total[0], total[1], total[2]);
    hi = make_float3(total[3], total[4], total[5]);
}

// ---------------------------------------------------------------------------------------------
// K2: Morton keys (bvh.h:257-275, bvh.cu:184-214) + the 8-bit digit histograms of the sort
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread3(uint32_t v)
{
    v = (v ^ (v << 16)) & 0xff0000ffu;
    v = (v ^ (v << 8)) & 0x0300f00fu;
    v = (v ^ (v << 4)) & 0x030c30c3u;
    v = (v ^ (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t quant1024(float x)
{
    const int q = (int)(x * 1024.0f);  // truncation toward zero
    return (uint32_t)min(max(q, 0), 1023);
}
__device__ __forceinline__ uint32_t morton30(float x, float y, float z)
{
    return (spread3(quant1024(z)) << 2) | (spread3(quant1024(y)) << 1) | spread3(quant1024(x));
}
// 21 bits per axis (2 097 152^3 grid); same centroid / scale arithmetic as the 30-bit code
__device__ __forceinline__ uint64_t spread3_21(uint64_t v)
{
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
__device__ __forceinline__ uint64_t quant2m(float x)
{
    const int q = (int)(x * 2097152.0f);
    return (uint64_t)min(max(q, 0), 2097151);
}
__device__ __forceinline__ uint64_t morton63(float x, float y, float z)
{
    return (spread3_21(quant2m(z)) << 2) | (spread3_21(quant2m(y)) << 1) | spread3_21(quant2m(x));
}

// 24-bit Hilbert index of a point of the unit cube (256^3 grid; Skilling's transpose form): consecutive indices are
// always face-adjacent cells, without the jumps a Morton curve makes at every power-of-two boundary.  Used only to
// ORDER query batches (wb_morton_order) -- never for tree keys, which follow the reference's Morton code.
__device__ __forceinline__ uint32_t hilbert24(float x, float y, float z)
{
    uint32_t X[3] = { quant1024(x) >> 2, quant1024(y) >> 2, quant1024(z) >> 2 };
    constexpr uint32_t M = 1u << 7;
#pragma unroll
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) {
                X[0] ^= P;
            } else {
                const uint32_t t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    uint32_t t = 0;
#pragma unroll
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q)
            t ^= Q - 1;
    X[0] ^= t, X[1] ^= t, X[2] ^= t;
    return (spread3(X[0]) << 2) | (spread3(X[1]) << 1) | spread3(X[2]);
}

template <class KeyT, bool GROUPED>
__device__ __forceinline__ KeyT make_key(float x, float y, float z, const int* __restrict__ groups, int item)
{
    if constexpr (sizeof(KeyT) == 4) {
        return morton30(x, y, z);
    } else if constexpr (GROUPED) {
        return ((uint64_t)(uint32_t)__ldg(groups + item) << 32) | (uint64_t)morton30(x, y, z);
    } else {
        return morton63(x, y, z);
    }
}

// histogram increment for digits that neighbouring items mostly share: if every valid lane of the
// warp holds the same digit, one lane adds the population count; otherwise fall back to per-lane atomics
__device__ __forceinline__ void hist_add_coherent(uint32_t* h, uint32_t d, bool valid)
{
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    if (vmask == 0u)
        return;
    const int leader = __ffs(vmask) - 1;
    const uint32_t d0 = __shfl_sync(0xffffffffu, d, leader);
    if (__all_sync(0xffffffffu, !valid || d == d0)) {
        if ((int)(threadIdx.x & 31) == leader)
            atomicAdd(&h[d0], (uint32_t)__popc(vmask));
    } else if (valid) {
        atomicAdd(&h[d], 1u);
    }
}

template <class Src, class KeyT, bool GROUPED>
__global__ void __launch_bounds__(BT)
k_morton_hist(Src src, int n, const float* __restrict__ partials, int num_partials, TreeHeader* __restrict__ hdr,
              const int* __restrict__ groups, KeyT* __restrict__ keys, uint32_t* __restrict__ ghist, int key_shift = 0)
{
    constexpr int PASSES = sizeof(KeyT);
    __shared__ uint32_t h[PASSES * 256];
    __shared__ float sm[BT / 32][6];
    for (int k = threadIdx.x; k < PASSES * 256; k += BT)
        h[k] = 0;

    // scene bounds and 1 / (extent + 1e-4) (IEEE division, bvh.cu:482-488), recomputed by every block
    float3 glo, ghi;
    reduce_partials(partials, num_partials, sm, glo, ghi);
    const float glx = glo.x, gly = glo.y, glz = glo.z;
    const float ivx = 1.0f / ((ghi.x - glo.x) + 0.0001f);
    const float ivy = 1.0f / ((ghi.y - glo.y) + 0.0001f);
    const float ivz = 1.0f / ((ghi.z - glo.z) + 0.0001f);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        hdr->total_lo[0] = glo.x, hdr->total_lo[1] = glo.y, hdr->total_lo[2] = glo.z;
        hdr->total_hi[0] = ghi.x, hdr->total_hi[1] = ghi.y, hdr->total_hi[2] = ghi.z;
        hdr->inv_edges[0] = ivx, hdr->inv_edges[1] = ivy, hdr->inv_edges[2] = ivz;
    }

    const int stride = gridDim.x * BT;
    const int iters = (n + stride - 1) / stride;
    for (int it = 0; it < iters; ++it) {
        const int i = it * stride + blockIdx.x * BT + threadIdx.x;
        const bool valid = i < n;
        KeyT code = 0;
        if (valid) {
            float3 lo, hi;
            src.bounds(i, lo, hi);
            const float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
            if (sizeof(KeyT) == 4 && key_shift < 0)  // query ordering only: Hilbert index instead of the Morton code
                code = (KeyT)hilbert24((cx - glx) * ivx, (cy - gly) * ivy, (cz - glz) * ivz);
            else
                code = make_key<KeyT, GROUPED>((cx - glx) * ivx, (cy - gly) * ivy, (cz - glz) * ivz, groups, i) >> key_shift;
            keys[i] = code;
        }
        // low digit: essentially random across a warp -> plain shared atomics; the higher digits are
        // usually identical across a warp of neighbouring items -> one add per warp when they are
        if (valid)
            atomicAdd(&h[(uint32_t)code & 255u], 1u);
#pragma unroll
        for (int p = 1; p < PASSES; ++p)
            hist_add_coherent(h + 256 * p, (uint32_t)(code >> (8 * p)) & 255u, valid);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < PASSES * 256; k += BT)
        if (h[k])
            atomicAdd(&ghist[k], h[k]);
}

// ---------------------------------------------------------------------------------------------
// K3: one onesweep pass (8-bit digit): per-tile ranking with warp match, decoupled look-back
// across tiles, shared-memory reorder so global writes are digit-contiguous.  Stable.
// ---------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS_LARGE = 16;  // 4096-key tiles: fewer look-back words for big inputs
constexpr int RS_ITEMS_SMALL = 8;   // 2048-key tiles: more resident warps while the input is small enough to be latency bound
#ifndef WB_RS_SMALL_LIMIT
#define WB_RS_SMALL_LIMIT (1ll << 24)
#endif
constexpr long long RS_SMALL_LIMIT = WB_RS_SMALL_LIMIT;
// 64-bit keys always use the small tile (a 4096-key tile of 8-byte keys would not fit 48 KB of static shared memory)
__host__ __device__ inline int rs_items_for(long long n, int key_bytes = 4)
{
    return (n < RS_SMALL_LIMIT || key_bytes == 8) ? RS_ITEMS_SMALL : RS_ITEMS_LARGE;
}
__host__ __device__ inline int rs_tile_for(long long n, int key_bytes = 4) { return RS_THREADS * rs_items_for(n, key_bytes); }
#define RS_FLAG_AGG (1u << 30)
#define RS_FLAG_INC (2u << 30)
#define RS_VAL_MASK ((1u << 30) - 1u)

template <bool IMPLICIT_VALS, int ITEMS, class KeyT>
__global__ void __launch_bounds__(RS_THREADS)
k_onesweep_pass(const KeyT* __restrict__ keys_in, const int* __restrict__ vals_in, KeyT* __restrict__ keys_out,
                int* __restrict__ vals_out, int n, int shift, const uint32_t* __restrict__ ghist_pass,
                volatile uint32_t* __restrict__ tile_status, unsigned* __restrict__ tile_ticket)
{
    constexpr int TILE = RS_THREADS * ITEMS;
    __shared__ uint32_t warp_hist[RS_WARPS][257];  // bin 256 collects out-of-range lanes
    __shared__ KeyT s_keys[TILE];
    __shared__ int s_vals[TILE];
    __shared__ int s_delta[256];  // global position = s_delta[digit] + position in tile
    __shared__ uint32_t s_scan[RS_WARPS];
    __shared__ uint32_t s_scan2[RS_WARPS];
    __shared__ int s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
        s_tile = (int)atomicAdd(tile_ticket, 1u);  // tiles start in ticket order => look-back never waits on an unscheduled tile
    for (int k = tid; k < RS_WARPS * 257; k += RS_THREADS)
        (&warp_hist[0][0])[k] = 0;
    __syncthreads();
    const int tile = s_tile;
    const int tile_base = tile * TILE;
    const int tile_count = min(TILE, n - tile_base);

    // -- load (warp-striped: item k of lane l sits at warp_base + 32k + l, so rank order = memory order)
    KeyT key[ITEMS];
    int val[ITEMS];
    uint32_t rank[ITEMS];
    const int warp_base = tile_base + warp * (32 * ITEMS);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int idx = warp_base + 32 * k + lane;
        const bool valid = idx < n;
        key[k] = valid ? keys_in[idx] : (KeyT)~(KeyT)0;
        val[k] = IMPLICIT_VALS ? idx : (valid ? vals_in[idx] : 0);
    }

    // -- rank within the warp, item by item
    uint32_t* wh = warp_hist[warp];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int idx = warp_base + 32 * k + lane;
        const uint32_t d = (idx < n) ? ((uint32_t)(key[k] >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t before = wh[d];
        __syncwarp();
        if ((peers & lt_mask) == 0)
            wh[d] = before + __popc(peers);
        __syncwarp();
        rank[k] = before + __popc(peers & lt_mask);
    }
    __syncthreads();

    // -- per digit (thread == digit): exclusive offsets across warps, tile total
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
        const uint32_t c = warp_hist[w][tid];
        warp_hist[w][tid] = count;
        count += c;
    }

    // -- publish the tile aggregate, then look back for the exclusive prefix of this digit
    volatile uint32_t* my_status = tile_status + (size_t)tile * 256 + tid;
    uint32_t excl = 0;
    if (tile == 0) {
        *my_status = RS_FLAG_INC | count;
    } else {
        *my_status = RS_FLAG_AGG | count;
        // walk back over the predecessors' words LB at a time: the loads of one batch are issued
        // back to back (one L2 round trip for the batch instead of one per tile), then consumed in
        // order; a word that is not published yet is re-polled in place
#ifndef WB_RS_LB
#define WB_RS_LB 4  // measured (1.3 M / 4 M / 10 M items, rebuild ms): 16: 0.367 / 0.88 / 2.05, 8: 0.356 / 0.85 / 1.97, 4: 0.350 / 0.83 / 1.95
#endif
        constexpr int LB = WB_RS_LB;
        bool found = false;
        for (int j = tile - 1; !found; j -= LB) {
            uint32_t v[LB];
#pragma unroll
            for (int b = 0; b < LB; ++b)
                v[b] = (j - b >= 0) ? tile_status[(size_t)(j - b) * 256 + tid] : RS_FLAG_INC;
#pragma unroll
            for (int b = 0; b < LB; ++b) {
                if (!found) {
                    uint32_t w = v[b];
                    while ((w & ~RS_VAL_MASK) == 0)
                        w = tile_status[(size_t)(j - b) * 256 + tid];
                    excl += w & RS_VAL_MASK;
                    found = (w & RS_FLAG_INC) != 0;
                }
            }
        }
        *my_status = RS_FLAG_INC | (excl + count);
    }

    // -- exclusive scans over the 256 digits: tile-local starts and global digit starts
    const uint32_t gcount = ghist_pass[tid];
    uint32_t inc_t = count, inc_g = gcount;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, inc_t, o);
        const uint32_t b = __shfl_up_sync(0xffffffffu, inc_g, o);
        if (lane >= o)
            inc_t += a, inc_g += b;
    }
    if (lane == 31)
        s_scan[warp] = inc_t, s_scan2[warp] = inc_g;
    __syncthreads();
    uint32_t off_t = 0, off_g = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w)
        if (w < warp)
            off_t += s_scan[w], off_g += s_scan2[w];
    const uint32_t tile_start = off_t + inc_t - count;   // first slot of this digit inside the tile
    const uint32_t glob_start = off_g + inc_g - gcount;  // first slot of this digit in the output
    s_delta[tid] = (int)(glob_start + excl) - (int)tile_start;
    // fold the tile-local digit start into the per-warp offsets
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w)
        warp_hist[w][tid] += tile_start;
    __syncthreads();

    // -- reorder through shared memory
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int idx = warp_base + 32 * k + lane;
        if (idx < n) {
            const uint32_t d = (uint32_t)(key[k] >> shift) & 255u;
            const uint32_t pos = wh[d] + rank[k];
            s_keys[pos] = key[k];
            s_vals[pos] = val[k];
        }
    }
    __syncthreads();

    // -- digit-contiguous global writes
    for (int j = tid; j < tile_count; j += RS_THREADS) {
        const KeyT kk = s_keys[j];
        const int g = s_delta[(uint32_t)(kk >> shift) & 255u] + j;
        keys_out[g] = kk;
        vals_out[g] = s_vals[j];
    }
}

// `passes` (default sizeof(KeyT)) 8-bit passes over (keys, vals) <-> (keys_alt, vals_alt); an even pass count ends
// in (keys, vals), an odd one in (keys_alt, vals_alt); and the buffers the descriptor points at never change (rebuild stays
// capture safe).  Pass 0 reads `keys` only and uses the element index as the value.
template <class KeyT>
void onesweep_sort(KeyT* keys, KeyT* keys_alt, int* vals, int* vals_alt, int n, const uint32_t* ghist,
                   uint32_t* tile_status, unsigned* tickets, cudaStream_t stream, int passes = (int)sizeof(KeyT))
{
    const int items = rs_items_for(n, (int)sizeof(KeyT));
    const int tiles = wb_div_up(n, RS_THREADS * items);
    for (int pass = 0; pass < passes; ++pass) {
        volatile uint32_t* status = tile_status + (size_t)pass * 256 * tiles;
        const bool fwd = (pass % 2) == 0;
        const KeyT* kin = fwd ? keys : keys_alt;
        const int* vin = fwd ? vals : vals_alt;
        KeyT* kout = fwd ? keys_alt : keys;
        int* vout = fwd ? vals_alt : vals;
        const uint32_t* gh = ghist + 256 * pass;
        unsigned* ticket = tickets + pass;
        if (items == RS_ITEMS_SMALL) {
            if (pass == 0)
                k_onesweep_pass<true, RS_ITEMS_SMALL, KeyT><<<tiles, RS_THREADS, 0, stream>>>(kin, nullptr, kout, vout, n, 8 * pass, gh, status, ticket);
            else
                k_onesweep_pass<false, RS_ITEMS_SMALL, KeyT><<<tiles, RS_THREADS, 0, stream>>>(kin, vin, kout, vout, n, 8 * pass, gh, status, ticket);
        } else if constexpr (sizeof(KeyT) == 4) {
            if (pass == 0)
                k_onesweep_pass<true, RS_ITEMS_LARGE, KeyT><<<tiles, RS_THREADS, 0, stream>>>(kin, nullptr, kout, vout, n, 8 * pass, gh, status, ticket);
            else
                k_onesweep_pass<false, RS_ITEMS_LARGE, KeyT><<<tiles, RS_THREADS, 0, stream>>>(kin, vin, kout, vout, n, 8 * pass, gh, status, ticket);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K4a: leaves (bvh.cu:228-255) -- one thread per sorted position: gather the triangle, refresh the
// packed-triangle cache and write the leaf's node record straight into its parent's pair (which
// pair is a function of the neighbouring keys only).  K4b (merge.cuh) then merges bottom-up.
// ---------------------------------------------------------------------------------------------
template <class Src, class KeyT, bool GROUPED>
__global__ void __launch_bounds__(BT)
k_leaves(Src src, int n, const KeyT* __restrict__ keys, const int* __restrict__ prim, NodeRec* __restrict__ pairs,
         int* __restrict__ pos_parent, float4* __restrict__ tris, unsigned* __restrict__ counters)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= n)
        return;
    if (i < n - 1)
        counters[i] = 0u;  // arrival counter of internal node n + i (merge.cuh)
    const int item = __ldg(prim + i);
    float3 lo, hi;
    if constexpr (Src::kIsMesh) {
        float3 p, q, r;
        src.tri(item, p, q, r);
        lo = wb_min3(wb_min3(p, q), r);
        hi = wb_max3(wb_max3(p, q), r);
        // sliver flag of the closest-point query (mesh.h:557-564), a per-triangle constant
        const float3 e0 = wb_sub(q, p), e1 = wb_sub(r, p), e2 = wb_sub(r, q);
        const float3 nrm = wb_cross(e0, e1);
        const float area2 = sqrtf(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z);
        const bool sliver = area2 / (wb_dot(e0, e0) + wb_dot(e1, e1) + wb_dot(e2, e2)) < 1.e-6f;
        float4* t = tris + 3 * (size_t)i;
        t[0] = make_float4(p.x, p.y, p.z, q.x);
        t[1] = make_float4(q.y, q.z, r.x, r.y);
        t[2] = make_float4(r.z, __int_as_float(item), __uint_as_float(sliver ? WB_TRI_SLIVER : 0u), 0.f);
    } else {
        src.bounds(item, lo, hi);
    }
    pos_parent[i] = WB_NO_PARENT;
    const bool go_right = wb_goes_right<KeyT, GROUPED>(keys, prim, n, i, i);
    NodeRec* rec = pairs + 2 * (size_t)(go_right ? i : i - 1) + (go_right ? 0 : 1);
    wb_store_rec(rec, lo, hi, (uint32_t)i | WB_LEAF, (uint32_t)i);  // a single item always fits leaf_size >= 1
}

// ---------------------------------------------------------------------------------------------
// K4b: small nodes, Karras style -- one thread per split position, no atomics, no dependency chain.
// The reference LBVH is the Cartesian tree of the key deltas (common-prefix lengths of neighbouring sorted keys,
// bvh.cu:218-226) wherever keys are distinct: the node split after position s covers [L, R], L / R being the nearest
// split on either side whose delta is smaller (a smaller delta = a shorter common prefix = a higher node).  A node
// whose range holds at most WB_SMALL_MAX positions, all with distinct keys, is therefore found by looking at no more
// than WB_SMALL_MAX - 1 deltas around s, and everything about it follows locally: its box (union of its <= 8 leaf
// records), its parent (the reference's own rule, wb_goes_right_k, on [L, R]), the packed-leaf flags of its children,
// its height (longest chain of successive delta minima).  About 80 % of the internal nodes of a mesh are of this kind.
// What is left -- nodes over more than 8 positions, and everything inside runs of EQUAL keys, where the reference
// breaks ties by primitive parity (bvh.cu:325-329) and the tree is not a Cartesian tree -- is merged bottom-up by
// k_merge<build>, which starts from the tops of the finished subtrees (unit[], merge.cuh) instead of the leaves.
// ---------------------------------------------------------------------------------------------
constexpr int SN_BP = 1024;   // splits per block
constexpr int SN_T = 256;
constexpr int SN_HALO = 16;   // keys staged on either side of the block's positions

// range [L, R] of the node split after s if it is a "small distinct" node; false otherwise (equal keys at s or inside
// the range, more than WB_SMALL_MAX positions, or the tree root)
template <bool GROUPED, class K> __device__ __forceinline__ bool small_range(const K& k, int n, int s, int& L, int& R)
{
    constexpr int W = 8 * (int)sizeof(decltype(k.key(0)));
    const int d = wb_key_delta_k(k, s);
    if (d == W)
        return false;
    int a = 0, b = 0;
    for (;; ++a) {
        const int j = s - 1 - a;
        if (j < 0)
            break;
        const int dj = wb_key_delta_k(k, j);
        if (dj < d)
            break;
        if (dj == W || dj == d || a + 2 > WB_SMALL_MAX - 1)
            return false;
    }
    for (;; ++b) {
        const int j = s + 1 + b;
        if (j > n - 2)
            break;
        const int dj = wb_key_delta_k(k, j);
        if (dj < d)
            break;
        if (dj == W || dj == d || a + b + 2 > WB_SMALL_MAX - 1)
            return false;
    }
    L = s - a, R = s + 1 + b;
    return !(L == 0 && R == n - 1);
}

template <class KeyT, bool GROUPED>
__global__ void __launch_bounds__(SN_T)
k_small_nodes(int n, int leaf_size, const KeyT* __restrict__ keys, const int* __restrict__ prim, NodeRec* pairs,
              int* __restrict__ parent_int, int* __restrict__ pos_parent, uint16_t* __restrict__ heights, int* __restrict__ unit)
{
    constexpr int WIN = SN_BP + 2 * SN_HALO + 2;
    __shared__ KeyT skeys[WIN];
    __shared__ unsigned char spar[WIN];
    __shared__ unsigned char sinfo[SN_BP + 1];  // split b0 - 1 + k: 0x80 | a | b << 3 when it is a small node
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * SN_BP;
    const int b1 = min(b0 + SN_BP - 1, n - 2);  // last split of the block
    const int base = b0 - SN_HALO - 1;
    for (int k = tid; k < WIN; k += SN_T) {
        const long long g = (long long)base + k;
        if (g >= 0 && g < n) {
            skeys[k] = __ldg(keys + g);
            spar[k] = (unsigned char)(__ldg(prim + g) % 2);
        }
    }
    __syncthreads();
    const BlockKeys<KeyT> bk { skeys, spar, base };
    for (int k = tid; k < SN_BP + 1; k += SN_T) {
        const int s = b0 - 1 + k;
        unsigned char info = 0;
        int L, R;
        if (s >= 0 && s <= b1 && small_range<GROUPED>(bk, n, s, L, R))
            info = (unsigned char)(0x80 | (s - L) | ((R - s - 1) << 3));
        sinfo[k] = info;
    }
    __syncthreads();
    for (int k = tid; k < SN_BP; k += SN_T) {
        const int s = b0 + k;
        if (s > b1)
            break;
        const unsigned char info = sinfo[k + 1];
        const bool left_done = (sinfo[k] & 0x80) != 0;
        if (!(info & 0x80)) {
            // position s is a plain leaf unless the split on its left is a small node (which then covers it)
            if (!left_done)
                unit[s] = WB_UNIT_LEAF;
            if (s == n - 2)
                unit[n - 1] = WB_UNIT_LEAF;  // nothing to the right of the last position
            continue;
        }
        const int L = s - (info & 7), R = s + 1 + ((info >> 3) & 7);
        // deltas of the range, 7 bits each (at most 7 of them)
        unsigned long long dd = 0;
        for (int j = L; j < R; ++j)
            dd |= (unsigned long long)wb_key_delta_k(bk, j) << (7 * (j - L));
        // box: union of the leaf records the leaf pass wrote (each into its own parent's pair)
        float3 lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
        int height = 0;
        for (int i = L; i <= R; ++i) {
            const bool gr = wb_goes_right_k<GROUPED>(bk, n, i, i);
            const float4* r4 = reinterpret_cast<const float4*>(pairs + 2 * (size_t)(gr ? i : i - 1) + (gr ? 0 : 1));
            const float4 r0 = r4[0], r1 = r4[1];
            lo = wb_min3(lo, make_float3(r0.x, r0.y, r0.z));
            hi = wb_max3(hi, make_float3(r1.x, r1.y, r1.z));
            // depth of leaf i below this node = number of successive delta minima walking away from it, both ways
            int depth = 0, run = 128;
            for (int j = i - 1; j >= L; --j) {
                const int dj = (int)((dd >> (7 * (j - L))) & 127u);
                if (dj < run)
                    run = dj, ++depth;
            }
            run = 128;
            for (int j = i; j < R; ++j) {
                const int dj = (int)((dd >> (7 * (j - L))) & 127u);
                if (dj < run)
                    run = dj, ++depth;
            }
            height = max(height, depth);
        }
        // packed leaves among the children (wb_absorb's rule)
        if (!wb_size_leaf_k<GROUPED>(bk, leaf_size, L, R)) {
            if (wb_size_leaf_k<GROUPED>(bk, leaf_size, L, s))
                pos_parent[L] = n + s;
            if (wb_size_leaf_k<GROUPED>(bk, leaf_size, s + 1, R))
                pos_parent[s + 1] = n + s;
        }
        heights[s] = wb_pack_height((unsigned)height, R - L + 1);
        // parent choice and record, exactly as the merge does for a node in hand
        const bool go_right = wb_goes_right_k<GROUPED>(bk, n, L, R);
        const int ps = go_right ? R : L - 1;
        parent_int[s] = n + ps;
        wb_store_rec(pairs + 2 * (size_t)ps + (go_right ? 0 : 1), lo, hi,
                     (uint32_t)(n + s) | (wb_size_leaf_k<GROUPED>(bk, leaf_size, L, R) ? WB_LEAF : 0u),
                     (uint32_t)(go_right ? L : R));
        // top of a finished subtree (its parent is not a small node): the merge pass starts from here
        int pL, pR;
        if (!small_range<GROUPED>(bk, n, ps, pL, pR)) {
            unit[L] = WB_UNIT_TOP | (s - L) | ((R - L) << 3) | (height << 6);
            for (int i = L + 1; i <= R; ++i)
                unit[i] = WB_UNIT_COVERED;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K5/K6: depth rule (bvh.cu:419-441): a single-group node at depth >= 32 (root = 1) becomes a
// packed leaf whatever its size.  Only trees taller than 30 edges can contain such a node;
// everything else returns after one header read.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int depth_capped(const int* __restrict__ parent_int, int n, int internal_node)
{
    // depth of an internal node (root = 1), capped at WB_MAX_DEPTH + 1
    int depth = 1;
    int p = parent_int[internal_node - n];
    while (p != WB_NO_PARENT && depth <= WB_MAX_DEPTH) {
        p = parent_int[p - n];
        depth++;
    }
    return depth;
}

template <class KeyT, bool GROUPED>
__device__ __forceinline__ bool node_single_group(const KeyT* __restrict__ keys, const NodeRec* __restrict__ pairs, int s)
{
    if (!GROUPED)
        return true;
    return wb_group_of(keys, (int)pairs[2 * (size_t)s].aux) == wb_group_of(keys, (int)pairs[2 * (size_t)s + 1].aux);
}

// marked by the depth rule: internal node n+s that is not a size leaf, single group, depth >= 32
template <class KeyT, bool GROUPED>
__device__ __forceinline__ bool depth_marked(const KeyT* __restrict__ keys, const NodeRec* __restrict__ pairs,
                                             const int* __restrict__ parent_int, int n, int s)
{
    return node_single_group<KeyT, GROUPED>(keys, pairs, s) && depth_capped(parent_int, n, n + s) >= WB_MAX_DEPTH;
}

// The reference walks the whole parent chain of every node (bvh.cu:402-443): ~32 dependent loads per node, which at
// 10 M items costs as much as the sort.  Here depths are memoised in a byte per internal node (0 = unknown, else the
// depth capped at WB_MAX_DEPTH + 1): k_deep_top initialises the table and walks the chains of the few nodes whose
// subtree is at least DEEP_TOP_HEIGHT tall; every other walk (k_deep_fix) stops at the first ancestor whose depth is
// known -- a tall node is at most DEEP_TOP_HEIGHT hops away -- and records its own result for the walks below it.
// Racing readers see either 0 or the final value, so the outcome does not depend on timing.
constexpr int DEEP_TOP_HEIGHT = 6;
#ifndef WB_DEEP_BFS_MIN
#define WB_DEEP_BFS_MIN (1 << 24)
#endif

__device__ __forceinline__ int depth_memo(const int* __restrict__ parent_int, const uint8_t* memo, int n, int slot)
{
    int hops = 0;
    int a = slot;
    for (;;) {
        const int p = parent_int[a];
        if (p == WB_NO_PARENT)
            return hops + 1;  // `a` is the root (depth 1)
        a = p - n;
        ++hops;
        if (memo) {
            const int d = (int)__ldcg(memo + a);
            if (d)
                return min(d + hops, WB_MAX_DEPTH + 1);
        }
        if (hops >= WB_MAX_DEPTH)
            return WB_MAX_DEPTH + 1;
    }
}

__global__ void __launch_bounds__(BT)
k_deep_top(int n, const TreeHeader* __restrict__ hdr, const int* __restrict__ parent_int,
           const uint16_t* __restrict__ heights, uint8_t* __restrict__ memo)
{
    if (hdr->height + 1 < WB_MAX_DEPTH)
        return;
    for (int s = blockIdx.x * BT + threadIdx.x; s < n - 1; s += gridDim.x * BT)
        memo[s] = (heights[s] & WB_HEIGHT_MASK) >= DEEP_TOP_HEIGHT ? (uint8_t)depth_memo(parent_int, nullptr, n, s) : (uint8_t)0;
}

// one thread per internal node.  A node that holds at most leaf_size positions is a size leaf or lies below one: settled
// (decided from the range length packed into heights[], two coalesced bytes, for the usual leaf sizes; from the pair
// record otherwise).  Every other node gets its depth -- memoised, or by walking to the first ancestor whose depth is
// known -- and, when it is the TOPMOST node at depth >= 32 of its group, becomes a visible leaf: flag in its record,
// pos_parent[] of its first position set, the entries of the size leaves it swallows cleared.
template <class KeyT, bool GROUPED>
__global__ void __launch_bounds__(BT)
k_deep_fix(int n, int leaf_size, TreeHeader* hdr, const KeyT* __restrict__ keys, const int* __restrict__ parent_int,
           const uint16_t* __restrict__ heights, uint8_t* memo, NodeRec* pairs, int* pos_parent)
{
    if (hdr->height + 1 < WB_MAX_DEPTH)
        return;
    for (int base = blockIdx.x * BT; base < n - 1; base += gridDim.x * BT) {
    const int s = base + (int)threadIdx.x;
    bool mark = false;  // this node is the topmost one at depth >= 32 of its group
    int left = 0, right = 0;
    if (s < n - 1) {
        bool settled;
        if (!GROUPED && leaf_size <= 16) {
            settled = (int)(__ldg(heights + s) >> 12) + 2 <= leaf_size;
        } else {
            settled = wb_size_leaf<KeyT, GROUPED>(keys, leaf_size, (int)pairs[2 * (size_t)s].aux, (int)pairs[2 * (size_t)s + 1].aux);
        }
        if (!settled) {
            int depth = (int)__ldcg(memo + s);
            if (depth == 0) {
                depth = depth_memo(parent_int, memo, n, s);
                memo[s] = (uint8_t)depth;
            }
            if (depth >= WB_MAX_DEPTH && node_single_group<KeyT, GROUPED>(keys, pairs, s)) {
                const int parent = parent_int[s];  // depth >= 32 => it has a parent
                const int ps = parent - n;
                // the parent sits one level up: when it is marked too, this node is muted below it
                if (!(depth - 1 >= WB_MAX_DEPTH && node_single_group<KeyT, GROUPED>(keys, pairs, ps))) {
                    left = (int)pairs[2 * (size_t)s].aux, right = (int)pairs[2 * (size_t)s + 1].aux;
                    NodeRec* rec = pairs + 2 * (size_t)ps + (s < ps ? 0 : 1);
                    rec->ref |= WB_LEAF;
                    pos_parent[left] = parent;
                    hdr->deep = 1;
                    mark = true;
                }
            }
        }
    }
    // the size leaves a marked node swallows are no longer visible: their pos_parent[] entries are cleared by the whole
    // warp, 32 consecutive positions per step (a depth-rule leaf of a 100 M-triangle mesh holds ~100 positions)
    unsigned todo = __ballot_sync(0xffffffffu, mark);
    const int lane = (int)(threadIdx.x & 31);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int l = __shfl_sync(0xffffffffu, left, src), r = __shfl_sync(0xffffffffu, right, src);
        for (int p = l + 1 + lane; p <= r; p += 32)
            pos_parent[p] = WB_NO_PARENT;
    }
    }  // grid-stride
}

// ---------------------------------------------------------------------------------------------
// K5': the same rule for LARGE ungrouped trees, top-down.  The memoised walks above touch every node; on a 100 M-triangle
// mesh with 30-bit keys (1 M runs of ~100 equal keys, each a comb ~50 deep) that is 10 ms of a 27 ms build.  But only nodes
// at depth <= 32 matter: the node that becomes a depth-rule leaf is exactly a non-leaf node AT depth 32 (its parent, at
// depth 31, is unmarked), and nothing below it is ever looked at.  So: a breadth-first sweep from the root, one launch
// per level, 31 levels, each reading the 64-byte pair record of the frontier's nodes and appending their non-leaf
// children; at level 32 the children are marked instead.  Nodes deeper than 32 are never touched.  Queues: the sort's
// two spare buffers.  (Grouped trees keep the walks: a node that spans groups is not marked whatever its depth, so the
// topmost marked node can sit below depth 32.)
// ---------------------------------------------------------------------------------------------
__global__ void k_bfs_start(int n, const TreeHeader* __restrict__ hdr, int* __restrict__ queue, int* __restrict__ counts)
{
    // counts[d] = number of nodes at depth d in the frontier (depth 1 = the root)
    for (int k = threadIdx.x; k < 40; k += blockDim.x)
        counts[k] = 0;
    __syncthreads();
    if (threadIdx.x == 0 && hdr->height + 1 >= WB_MAX_DEPTH && !(hdr->root_ref & WB_LEAF)) {
        queue[0] = (int)(hdr->root_ref & WB_IDX_MASK) - n;
        counts[1] = 1;
    }
}

// frontier of depth `depth` (slots in `in`) -> its non-leaf children at depth + 1 (slots in `out`), or, when the
// children sit at depth 32, the depth-rule leaves
__global__ void __launch_bounds__(BT)
k_bfs_level(int n, int depth, TreeHeader* hdr, const int* __restrict__ in, int* __restrict__ out, int* __restrict__ counts,
            NodeRec* pairs, int* pos_parent)
{
    const int m = counts[depth];
    const int lane = (int)(threadIdx.x & 31);
    const bool last = depth + 1 >= WB_MAX_DEPTH;
    for (int base = blockIdx.x * BT; base < m; base += gridDim.x * BT) {
        const int i = base + (int)threadIdx.x;
        int clear_from[2] = { 0, 0 }, clear_to[2] = { -1, -1 };
        if (i < m) {
            const int s = in[i];
            const float4* p4 = reinterpret_cast<const float4*>(pairs + 2 * (size_t)s);
            const uint32_t lref = __float_as_uint(p4[0].w), laux = __float_as_uint(p4[1].w);
            const uint32_t rref = __float_as_uint(p4[2].w), raux = __float_as_uint(p4[3].w);
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const uint32_t ref = side ? rref : lref;
                if ((ref & WB_LEAF) || (int)ref < n)
                    continue;  // a packed leaf (by size) or an original leaf: settled
                if (!last) {
                    out[atomicAdd(&counts[depth + 1], 1)] = (int)ref - n;
                } else {
                    // depth 32: this child becomes a visible leaf over its whole range
                    const int left = side ? s + 1 : (int)laux, right = side ? (int)raux : s;
                    pairs[2 * (size_t)s + side].ref = ref | WB_LEAF;
                    pos_parent[left] = n + s;
                    clear_from[side] = left + 1, clear_to[side] = right;
                    hdr->deep = 1;
                }
            }
        }
        if (last) {  // the size leaves a marked node swallows lose their entries, 32 positions per step
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                unsigned todo = __ballot_sync(0xffffffffu, clear_to[side] >= clear_from[side] && clear_to[side] >= 0);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int a = __shfl_sync(0xffffffffu, clear_from[side], src), b = __shfl_sync(0xffffffffu, clear_to[side], src);
                    for (int p = a + lane; p <= b; p += 32)
                        pos_parent[p] = WB_NO_PARENT;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// export: rebuild the reference's node_lowers / node_uppers / node_parents / root arrays
// (bvh.h:161-207) from the pair layout, for parity checks and for Warp kernels that walk `id`.
// ---------------------------------------------------------------------------------------------
struct RefHalf {
    float x, y, z;
    uint32_t ib;
};

template <class KeyT, bool GROUPED>
__global__ void __launch_bounds__(BT)
k_export_reference_layout(int n, int leaf_size, const TreeHeader* __restrict__ hdr, const KeyT* __restrict__ keys,
                          const NodeRec* __restrict__ pairs, const int* __restrict__ parent_int, RefHalf* lowers,
                          RefHalf* uppers, int* parents, int* root)
{
    const int c = blockIdx.x * BT + threadIdx.x;
    if (c >= 2 * n - 1)
        return;
    const int root_node = (int)(hdr->root_ref & WB_IDX_MASK);
    if (c == root_node) {
        lowers[c].x = hdr->lx, lowers[c].y = hdr->ly, lowers[c].z = hdr->lz;
        uppers[c].x = hdr->hx, uppers[c].y = hdr->hy, uppers[c].z = hdr->hz;
        parents[c] = -1;
        *root = c;
    }
    if (c < n) {  // original leaf: packed-leaf marking turns [c, c] into [c, c+1)
        lowers[c].ib = WB_LEAF | (uint32_t)c;
        uppers[c].ib = (uint32_t)(c + 1);
        return;
    }
    const int s = c - n;
    const NodeRec L = pairs[2 * (size_t)s], R = pairs[2 * (size_t)s + 1];
    const int cl = (int)(L.ref & WB_IDX_MASK), cr = (int)(R.ref & WB_IDX_MASK);
    lowers[cl].x = L.lx, lowers[cl].y = L.ly, lowers[cl].z = L.lz;
    uppers[cl].x = L.hx, uppers[cl].y = L.hy, uppers[cl].z = L.hz;
    lowers[cr].x = R.lx, lowers[cr].y = R.ly, lowers[cr].z = R.lz;
    uppers[cr].x = R.hx, uppers[cr].y = R.hy, uppers[cr].z = R.hz;
    parents[cl] = c;
    parents[cr] = c;
    const int left = (int)L.aux, right = (int)R.aux;
    bool leaf = wb_size_leaf<KeyT, GROUPED>(keys, leaf_size, left, right);
    if (!leaf && hdr->height + 1 >= WB_MAX_DEPTH)
        leaf = depth_marked<KeyT, GROUPED>(keys, pairs, parent_int, n, s);
    if (leaf) {
        lowers[c].ib = WB_LEAF | (uint32_t)left;
        uppers[c].ib = (uint32_t)(right + 1);
    } else {
        lowers[c].ib = (uint32_t)cl;
        uppers[c].ib = (uint32_t)cr;
    }
}

// per-triangle bounds in item order: wp::Mesh::lowers / uppers == BVH::item_lowers / item_uppers of a mesh
// (compute_triangle_bounds, mesh.cu:16-36), for Warp kernels that read them through the id (mesh.h:2574, bvh.h:602)
__global__ void __launch_bounds__(BT)
k_export_item_bounds(MeshSource src, int n, float* __restrict__ lowers, float* __restrict__ uppers)
{
    const int t = blockIdx.x * BT + threadIdx.x;
    if (t >= n)
        return;
    float3 lo, hi;
    src.bounds(t, lo, hi);
    lowers[3 * (size_t)t] = lo.x, lowers[3 * (size_t)t + 1] = lo.y, lowers[3 * (size_t)t + 2] = lo.z;
    uppers[3 * (size_t)t] = hi.x, uppers[3 * (size_t)t + 1] = hi.y, uppers[3 * (size_t)t + 2] = hi.z;
}

// single item: the root is leaf 0 (bvh.cu:285-290 with n == 1)
template <class Src, class KeyT, bool GROUPED>
__global__ void k_single_item(Src src, int leaf_size, const int* groups, int* prim, KeyT* keys, int* pos_parent,
                              float4* tris, TreeHeader* hdr)
{
    float3 lo, hi;
    src.bounds(0, lo, hi);
    if constexpr (Src::kIsMesh) {
        float3 p, q, r;
        src.tri(0, p, q, r);
        const float3 e0 = wb_sub(q, p), e1 = wb_sub(r, p), e2 = wb_sub(r, q);
        const float3 nrm = wb_cross(e0, e1);
        const float area2 = sqrtf(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z);
        const bool sliver = area2 / (wb_dot(e0, e0) + wb_dot(e1, e1) + wb_dot(e2, e2)) < 1.e-6f;
        tris[0] = make_float4(p.x, p.y, p.z, q.x);
        tris[1] = make_float4(q.y, q.z, r.x, r.y);
        tris[2] = make_float4(r.z, __int_as_float(0), __uint_as_float(sliver ? WB_TRI_SLIVER : 0u), 0.f);
    }
    prim[0] = 0;
    pos_parent[0] = WB_ROOT_PARENT;
    hdr->lx = lo.x, hdr->ly = lo.y, hdr->lz = lo.z;
    hdr->hx = hi.x, hdr->hy = hi.y, hdr->hz = hi.z;
    hdr->root_ref = WB_LEAF | 0u;
    hdr->root_count = 1;
    hdr->height = 0;
    hdr->deep = 0;
    hdr->n = 1;
    hdr->leaf_size = leaf_size;
    hdr->total_lo[0] = lo.x, hdr->total_lo[1] = lo.y, hdr->total_lo[2] = lo.z;
    hdr->total_hi[0] = hi.x, hdr->total_hi[1] = hi.y, hdr->total_hi[2] = hi.z;
    hdr->inv_edges[0] = 1.0f / ((hi.x - lo.x) + 0.0001f);
    hdr->inv_edges[1] = 1.0f / ((hi.y - lo.y) + 0.0001f);
    hdr->inv_edges[2] = 1.0f / ((hi.z - lo.z) + 0.0001f);
    const float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
    keys[0] = make_key<KeyT, GROUPED>((cx - lo.x) * hdr->inv_edges[0], (cy - lo.y) * hdr->inv_edges[1],
                                      (cz - lo.z) * hdr->inv_edges[2], groups, 0);
}

// WARP_B200_SMALL_NODES=1 turns the Karras-style small-node pass on.  OFF by default: measured on B200 it costs more than it
// saves (rebuild ms, off / on: C2 sphere 0.370 / 0.451, 10 M heightfield 30-bit 1.95 / 2.07 and 63-bit 2.49 / 2.87, 4 M cloth
// 0.83 / 0.91).  k_small_nodes takes 97 us at C2 while k_merge only drops from 144 to 129 us although ~80 % of its node
// merges are gone: the merge is bound by the latency of its dependent chain (chunk -> block counters -> global counters
// up the spine), not by the bulk of small merges.  Kept as a tested alternative (tests/test_gpu_build.py runs both).
bool small_nodes_enabled()
{
    static const bool env_on = [] {
        const char* v = getenv("WARP_B200_SMALL_NODES");
        return v && atoi(v) != 0;
    }();
    return g_wb_small_nodes >= 0 ? g_wb_small_nodes != 0 : env_on;
}

// item count from which the depth rule is applied by the top-down sweep (k_bfs_level) instead of the memoised walks;
// WARP_B200_DEEP_BFS_MIN overrides (0 = always, a huge value = never)
int deep_bfs_threshold()
{
    static const long long t = [] {
        const char* v = getenv("WARP_B200_DEEP_BFS_MIN");
        return v ? atoll(v) : (long long)WB_DEEP_BFS_MIN;
    }();
    return t > 0x7fffffffll ? 0x7fffffff : (int)t;
}

bool use_cub_sort()
{
    const char* v = getenv("WARP_B200_SORT");
    return v && strcmp(v, "cub") == 0;
}

template <class Src, class KeyT, bool GROUPED> const char* build_impl(BvhState& s, Src src, cudaStream_t stream)
{
    const int n = s.n;
    KeyT* keys = (KeyT*)s.keys;
    KeyT* keys_alt = (KeyT*)s.keys_alt;
    constexpr int PASSES = sizeof(KeyT);
    if (n == 1) {
        k_single_item<Src, KeyT, GROUPED><<<1, 1, 0, stream>>>(src, s.leaf_size, s.groups, s.prim, keys, s.pos_parent,
                                                               s.tris, s.header);
        WB_CUDA_TRY(cudaGetLastError());
        return nullptr;
    }

    // K1 scene bounds (+ clears histograms / tickets and the sort's look-back words; the arrival counters are cleared by
    // the leaf pass -- two memset launches less)
    const int sort_tiles = wb_div_up(n, rs_tile_for(n, (int)sizeof(KeyT)));  // what onesweep_sort will use (<= s.num_tiles)
    k_scene_bounds<<<s.bounds_blocks, BT, 0, stream>>>(src, n, s.partials, s.tickets, s.ghist, (uint4*)s.tile_status,
                                                       (size_t)64 * PASSES * (size_t)sort_tiles);
    // K2 Morton keys + histograms
    k_morton_hist<Src, KeyT, GROUPED><<<s.bounds_blocks, BT, 0, stream>>>(src, n, s.partials, s.bounds_blocks, s.header,
                                                                         s.groups, keys, s.ghist);

    if (use_cub_sort()) {
        // library cross-check path (tests only): stable LSD sort over every key bit
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, keys, keys_alt, s.prim_alt, s.prim, n, 0, 8 * PASSES, stream);
        if (need > s.cub_temp_bytes) {
            if (s.cub_temp)
                cudaFree(s.cub_temp);
            WB_CUDA_TRY(cudaMalloc(&s.cub_temp, need));
            s.cub_temp_bytes = need;
        }
        k_iota<<<wb_div_up(n, BT), BT, 0, stream>>>(s.prim_alt, n);
        WB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(s.cub_temp, need, keys, keys_alt, s.prim_alt, s.prim, n, 0,
                                                    8 * PASSES, stream));
        WB_CUDA_TRY(cudaMemcpyAsync(keys, keys_alt, sizeof(KeyT) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
    } else {
        // K3 x PASSES
        onesweep_sort<KeyT>(keys, keys_alt, s.prim, s.prim_alt, n, s.ghist, s.tile_status, s.tickets, stream);
    }

    // K4a leaves, K4b chunked bottom-up merge
    k_leaves<Src, KeyT, GROUPED><<<wb_div_up(n, BT), BT, 0, stream>>>(src, n, keys, s.prim, s.pairs, s.pos_parent, s.tris, s.counters);
    {
        // K4b small distinct-key nodes, one thread per split (unit[] lives in the sort's spare value buffer, free until the
        // depth pass reuses it); K4c merges what is left, starting from the tops of the finished subtrees
        int* unit = nullptr;
        if (small_nodes_enabled() && n > 64) {
            unit = s.prim_alt;
            k_small_nodes<KeyT, GROUPED><<<wb_div_up(n - 1, SN_BP), SN_T, 0, stream>>>(n, s.leaf_size, keys, s.prim, s.pairs, s.parent_int,
                                                                                     s.pos_parent, s.heights, unit);
        }
        const MergeArgs<KeyT> ma { n, s.leaf_size, keys, s.prim, s.pairs, s.parent_int, s.pos_parent, s.counters, s.header, s.heights, unit };
        s.plan_valid = false;
        k_merge<false, KeyT, GROUPED><<<wb_div_up(n, BP), TBM, 0, stream>>>(ma);
    }
    // K5/K6 depth rule (early-out unless the tree is at least 32 levels deep); the depth table lives in the sort's
    // spare value buffer, free until the next sort
    if (!GROUPED && n >= deep_bfs_threshold()) {
        // large ungrouped trees: top-down sweep over the nodes of depth <= 32 only (see k_bfs_level)
        int* q0 = (int*)keys_alt;
        int* q1 = s.prim_alt;
        int* counts = (int*)s.bfs_counts;
        k_bfs_start<<<1, 64, 0, stream>>>(n, s.header, q0, counts);
        const int grid = min(148 * 8, wb_div_up(n, BT));  // grid-stride over the frontier
        for (int depth = 1; depth < WB_MAX_DEPTH; ++depth)
            k_bfs_level<<<grid, BT, 0, stream>>>(n, depth, s.header, (depth & 1) ? q0 : q1, (depth & 1) ? q1 : q0, counts, s.pairs,
                                                 s.pos_parent);
        WB_CUDA_TRY(cudaGetLastError());
        return nullptr;
    }
    const int deep_grid = min(148 * 8, wb_div_up(n - 1, BT));  // grid-stride: the usual early-out costs ~1000 blocks, not n / 256
    k_deep_top<<<deep_grid, BT, 0, stream>>>(n, s.header, s.parent_int, s.heights, (uint8_t*)s.prim_alt);
    k_deep_fix<KeyT, GROUPED><<<deep_grid, BT, 0, stream>>>(n, s.leaf_size, s.header, keys, s.parent_int, s.heights,
                                                            (uint8_t*)s.prim_alt, s.pairs, s.pos_parent);
    WB_CUDA_TRY(cudaGetLastError());
    return nullptr;
}

template <class Src> const char* build_dispatch(BvhState& s, Src src, cudaStream_t stream)
{
    if (s.key_bytes == 4)
        return build_impl<Src, uint32_t, false>(s, src, stream);
    if (s.groups)
        return build_impl<Src, uint64_t, true>(s, src, stream);
    return build_impl<Src, uint64_t, false>(s, src, stream);
}

// one stream-ordered arena per tree: a single cudaMallocAsync from the device's default pool (kept
// warm by an unlimited release threshold), carved into 256-byte aligned sub-buffers
bool pool_ready(int device)
{
    static bool done[64] = {};
    if (device < 0 || device >= 64)
        return false;
    if (!done[device]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) != cudaSuccess)
            return false;
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        done[device] = true;
    }
    return true;
}

// streaming kernels K1/K2: four blocks per SM, each thread strides over ~n / (148*4*256) items
int bounds_grid(long long n)
{
    static int sm_count[64] = {};  // per device (the calling thread's current device: every caller holds a DeviceGuard)
    int dev = 0;
    cudaGetDevice(&dev);
    int sms = (dev >= 0 && dev < 64) ? sm_count[dev] : 0;
    if (!sms) {
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
        if (dev >= 0 && dev < 64)
            sm_count[dev] = sms;
    }
#ifndef WB_BOUNDS_BLOCKS_PER_SM
#define WB_BOUNDS_BLOCKS_PER_SM 4  // 2 / 4 / 8 / 16 measured within 1.5 % of each other at 1.3 M - 100 M triangles
#endif
    return (int)min((long long)sms * WB_BOUNDS_BLOCKS_PER_SM, max(1ll, (n + BT - 1) / BT));
}

}  // namespace

int g_wb_small_nodes = -1;  // -1: follow WARP_B200_SMALL_NODES (default off); 0 / 1: wp_b200_set_experiment("small_nodes", v)

const char* wb_alloc_tree(BvhState& s, cudaStream_t stream)
{
    const size_t n = (size_t)s.n;
    const size_t ni = n > 1 ? n - 1 : 1;
    const size_t kb = (size_t)s.key_bytes;
    // look-back words are sized for the SMALL tile whatever tile the build sort picks: the refit plan sorts n - 1 keys
    // (and may therefore fall below RS_SMALL_LIMIT when the build did not, e.g. n = 2^24) out of the same buffer
    s.num_tiles = wb_div_up((long long)n, RS_THREADS * RS_ITEMS_SMALL);
    s.bounds_blocks = bounds_grid((long long)n);
    size_t off = 0;
    auto take = [&off](size_t bytes) {
        const size_t at = off;
        off += (bytes + 255) & ~(size_t)255;
        return at;
    };
    const size_t o_header = take(sizeof(TreeHeader)), o_tickets = take(sizeof(unsigned) * 16),
                 o_ghist = take(sizeof(uint32_t) * 8 * 256), o_partials = take(sizeof(float) * 6 * (size_t)s.bounds_blocks),
                 o_edge = take(sizeof(double) * 296), o_bfs = take(sizeof(int) * 64),
                 o_keys = take(kb * n), o_keys_alt = take(kb * n), o_prim = take(4 * n), o_prim_alt = take(4 * n),
                 o_pairs = take(sizeof(NodeRec) * 2 * ni), o_parent = take(4 * ni), o_pos = take(4 * n),
                 o_counters = take(4 * ni), o_tris = take(s.is_mesh ? sizeof(float4) * 3 * n : 0),
                 o_status = take(sizeof(uint32_t) * 256 * kb * (size_t)s.num_tiles), o_heights = take(2 * ni),
                 o_pkeys = take(4 * ni), o_pnodes = take(4 * ni), o_pdst = take(4 * ni),
                 o_pbegin = take(4 * (size_t)wb_div_up((long long)n, WB_WAVE_BP)),
                 o_pend = take(4 * (size_t)wb_div_up((long long)n, WB_WAVE_BP)), o_uflags = take(n), o_ptop = take(4 * n),
                 o_pntop = take(4);
    s.arena_bytes = off;
    s.arena_async = pool_ready(s.device);
    if (s.arena_async)
        WB_CUDA_TRY(cudaMallocAsync(&s.arena, off, stream));
    else
        WB_CUDA_TRY(cudaMalloc(&s.arena, off));
    char* b = (char*)s.arena;
    s.header = (TreeHeader*)(b + o_header), s.tickets = (unsigned*)(b + o_tickets), s.ghist = (uint32_t*)(b + o_ghist);
    s.partials = (float*)(b + o_partials), s.edge_partials = (double*)(b + o_edge), s.bfs_counts = (int*)(b + o_bfs), s.keys = b + o_keys, s.keys_alt = b + o_keys_alt;
    s.prim = (int*)(b + o_prim), s.prim_alt = (int*)(b + o_prim_alt), s.pairs = (NodeRec*)(b + o_pairs);
    s.parent_int = (int*)(b + o_parent), s.pos_parent = (int*)(b + o_pos), s.counters = (unsigned*)(b + o_counters);
    s.tris = s.is_mesh ? (float4*)(b + o_tris) : nullptr, s.tile_status = (uint32_t*)(b + o_status);
    s.heights = (uint16_t*)(b + o_heights), s.plan_keys = (uint32_t*)(b + o_pkeys), s.plan_nodes = (int*)(b + o_pnodes);
    s.plan_dst = (uint32_t*)(b + o_pdst), s.plan_begin = (int*)(b + o_pbegin), s.plan_end = (int*)(b + o_pend);
    s.unit_flags = (uint8_t*)(b + o_uflags), s.plan_top = (uint32_t*)(b + o_ptop), s.plan_ntop = (int*)(b + o_pntop);
    s.plan_valid = false;
    // header, tickets, histograms start from zero (one small memset: they are adjacent)
    WB_CUDA_TRY(cudaMemsetAsync(s.arena, 0, o_partials, stream));
    return nullptr;
}

void wb_free_tree(BvhState& s, cudaStream_t stream)
{
    if (s.arena) {
        if (s.arena_async)
            cudaFreeAsync(s.arena, stream);
        else
            cudaFree(s.arena);
    }
    void* ptrs[] = { s.cub_temp, s.ref_lowers, s.ref_uppers, s.ref_parents, s.ref_root, s.ref_counts, s.tri_lowers, s.tri_uppers };
    for (void* p : ptrs)
        if (p)
            cudaFree(p);
    s = BvhState();
}

const char* wb_build(BvhState& s, cudaStream_t stream)
{
    if (s.n <= 0)
        return nullptr;
    s.host_built = false;  // an in-place rebuild is always an LBVH (bvh.cu:819-843), whatever built the tree first
    if (s.is_mesh)
        return build_dispatch(s, MeshSource { s.points, s.indices }, stream);
    return build_dispatch(s, BoxSource { s.item_lowers, s.item_uppers }, stream);
}

// ------------------------------------------------------------------------------------------------
// EXPERIMENT (not on the build path; wp_b200_experiment_parallel_topology): the parent of every
// internal node computed INDEPENDENTLY from the sorted keys -- the reference LBVH is the Cartesian tree of the key-delta
// array, so the node split after position s covers (L, R] with L / R the nearest split on either side whose delta is
// smaller: two galloping searches on clz(key[j] ^ key[s]), no atomics, no dependency chain.  Runs of equal keys, where
// the parity tie-break of bvh.cu:325-329 decides, are replayed sequentially by the thread of the run's first split
// (runs longer than TOPO_RUN_MAX raise `fail`: a production kernel would hand those to the merge kernel).
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int TOPO_RUN_MAX = 64;

template <class KeyT>
__device__ __forceinline__ bool topo_goes_right(const KeyT* __restrict__ keys, const int* __restrict__ prim, int n, int l, int r)
{
    if (l == 0)
        return true;
    if (r == n - 1)
        return false;
    const int dr = wb_clz_key((KeyT)(keys[r] ^ keys[r + 1])), dl = wb_clz_key((KeyT)(keys[l - 1] ^ keys[l]));
    if (dr != dl)
        return dr > dl;
    return ((prim[l - 1] % 2) ^ (prim[r] % 2)) != 0;
}

template <class KeyT>
__global__ void __launch_bounds__(BT)
k_topology(int n, const KeyT* __restrict__ keys, const int* __restrict__ prim, int* __restrict__ parent_out, int* __restrict__ fail)
{
    const int s = blockIdx.x * BT + threadIdx.x;
    if (s >= n - 1)
        return;
    const KeyT ks = keys[s], ks1 = keys[s + 1];
    if (ks == ks1) {
        if (s > 0 && keys[s - 1] == ks)
            return;  // not the first split of its run
        const int p = s;
        int q = s + 1;
        while (q + 1 < n && keys[q + 1] == ks)
            ++q;
        if (q - p + 1 > TOPO_RUN_MAX) {
            *fail = 1;
            return;
        }
        int parked[TOPO_RUN_MAX];
        int depth = 0;
        for (int i = p; i <= q; ++i) {
            int l = i, split = -1;
            const int r = i;
            for (;;) {
                const bool run_root = (l == p && r == q);
                const bool whole = (l == 0 && r == n - 1);
                const bool gr = topo_goes_right(keys, prim, n, l, r);
                if (split >= 0)
                    parent_out[split] = whole ? -1 : n + (gr ? r : l - 1);
                if (run_root || gr) {
                    if (!run_root)
                        parked[depth++] = l;
                    break;
                }
                split = l - 1;        // merges with the parked node that ends at l - 1
                l = parked[--depth];
            }
        }
        return;
    }
    const int d = wb_clz_key((KeyT)(ks ^ ks1));
    // left end: smallest j <= s with clz(key[j] ^ key[s]) > d (monotone in j on sorted keys)
    int l = s;
    {
        int step = 1;
        while (s - step >= 0 && wb_clz_key((KeyT)(keys[s - step] ^ ks)) > d)
            step <<= 1;
        int lo = max(s - step, -1), hi = s - (step >> 1);  // key[lo] fails (or lo == -1), key[hi] passes
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (wb_clz_key((KeyT)(keys[mid] ^ ks)) > d)
                hi = mid;
            else
                lo = mid;
        }
        l = hi;
    }
    int r = s + 1;
    {
        int step = 1;
        while (s + 1 + step < n && wb_clz_key((KeyT)(keys[s + 1 + step] ^ ks1)) > d)
            step <<= 1;
        int hi = min(s + 1 + step, n), lo = s + 1 + (step >> 1);  // key[lo] passes, key[hi] fails (or hi == n)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (wb_clz_key((KeyT)(keys[mid] ^ ks1)) > d)
                lo = mid;
            else
                hi = mid;
        }
        r = lo;
    }
    if (l == 0 && r == n - 1)
        parent_out[s] = -1;
    else
        parent_out[s] = n + (topo_goes_right(keys, prim, n, l, r) ? r : l - 1);
}
}  // namespace

const char* wb_experiment_topology(BvhState& s, int* parent_out, int* fail, cudaStream_t stream)
{
    if (s.n < 2)
        return nullptr;
    WB_CUDA_TRY(cudaMemsetAsync(fail, 0, sizeof(int), stream));
    const int grid = wb_div_up(s.n - 1, BT);
    if (s.key_bytes == 4)
        k_topology<uint32_t><<<grid, BT, 0, stream>>>(s.n, (const uint32_t*)s.keys, s.prim, parent_out, fail);
    else
        k_topology<uint64_t><<<grid, BT, 0, stream>>>(s.n, (const uint64_t*)s.keys, s.prim, parent_out, fail);
    WB_CUDA_TRY(cudaGetLastError());
    return nullptr;
}

const char* wb_export_reference_layout(BvhState& s, cudaStream_t stream)
{
    if (s.n <= 0)
        return nullptr;
    const size_t m = 2 * (size_t)s.n - 1;
    if (!s.ref_lowers) {
        WB_CUDA_TRY(cudaMalloc(&s.ref_lowers, sizeof(RefHalf) * m));
        WB_CUDA_TRY(cudaMalloc(&s.ref_uppers, sizeof(RefHalf) * m));
        WB_CUDA_TRY(cudaMalloc(&s.ref_parents, sizeof(int) * m));
        WB_CUDA_TRY(cudaMalloc(&s.ref_counts, sizeof(int) * m));
        WB_CUDA_TRY(cudaMalloc(&s.ref_root, sizeof(int)));
        WB_CUDA_TRY(cudaMemset(s.ref_counts, 0, sizeof(int) * m));
        if (s.is_mesh) {
            WB_CUDA_TRY(cudaMalloc((void**)&s.tri_lowers, 12 * (size_t)s.n));
            WB_CUDA_TRY(cudaMalloc((void**)&s.tri_uppers, 12 * (size_t)s.n));
        }
    }
    if (s.is_mesh)
        k_export_item_bounds<<<wb_div_up(s.n, BT), BT, 0, stream>>>(MeshSource { s.points, s.indices }, s.n, s.tri_lowers,
                                                                    s.tri_uppers);
    WB_CUDA_TRY(cudaMemsetAsync(s.ref_lowers, 0, sizeof(RefHalf) * m, stream));
    WB_CUDA_TRY(cudaMemsetAsync(s.ref_uppers, 0, sizeof(RefHalf) * m, stream));
    const int grid = wb_div_up((long long)m, BT);
    RefHalf* lo = (RefHalf*)s.ref_lowers;
    RefHalf* hi = (RefHalf*)s.ref_uppers;
    if (s.key_bytes == 4)
        k_export_reference_layout<uint32_t, false><<<grid, BT, 0, stream>>>(s.n, s.leaf_size, s.header, (const uint32_t*)s.keys,
                                                                            s.pairs, s.parent_int, lo, hi, s.ref_parents, s.ref_root);
    else if (s.groups)
        k_export_reference_layout<uint64_t, true><<<grid, BT, 0, stream>>>(s.n, s.leaf_size, s.header, (const uint64_t*)s.keys,
                                                                           s.pairs, s.parent_int, lo, hi, s.ref_parents, s.ref_root);
    else
        k_export_reference_layout<uint64_t, false><<<grid, BT, 0, stream>>>(s.n, s.leaf_size, s.header, (const uint64_t*)s.keys,
                                                                            s.pairs, s.parent_int, lo, hi, s.ref_parents, s.ref_root);
    WB_CUDA_TRY(cudaGetLastError());
    return nullptr;
}

// the refit's merge pass lives here because it shares the key-typed instantiations
const char* wb_refit_merge(BvhState& s, cudaStream_t stream)
{
    if (s.n <= 1)
        return nullptr;
    const int grid = wb_div_up(s.n, BP);
    if (s.key_bytes == 4) {
        const MergeArgs<uint32_t> ma { s.n, s.leaf_size, (const uint32_t*)s.keys, s.prim, s.pairs, s.parent_int, s.pos_parent, s.counters, s.header, nullptr, nullptr };
        k_merge<true, uint32_t, false><<<grid, TBM, 0, stream>>>(ma);
    } else {
        const MergeArgs<uint64_t> ma { s.n, s.leaf_size, (const uint64_t*)s.keys, s.prim, s.pairs, s.parent_int, s.pos_parent, s.counters, s.header, nullptr, nullptr };
        k_merge<true, uint64_t, false><<<grid, TBM, 0, stream>>>(ma);  // the static-tree replay never consults groups
    }
    WB_CUDA_TRY(cudaGetLastError());
    return nullptr;
}

// ------------------------------------------------------------------------------------------------
// Refit plan (consumed by k_refit_wave, bvh_refit.cu).  The tree is static between builds, so the order in which
// a refit may visit the nodes can be fixed once: every visible internal node whose range lies inside one block of
// WB_WAVE_BP sorted positions gets the key  block << 12 | height  and the nodes are sorted by it (same onesweep) --
// a block then walks its own nodes level by level with __syncthreads() instead of atomic arrival counters.  Nodes
// that span blocks (key TOP) are left to the global counters; packed leaves and muted nodes (key SKIP) need nothing.
// ------------------------------------------------------------------------------------------------
namespace {

// the leaf flag of internal node n+s lives in its record inside the parent's pair (the root's in the header)
__device__ __forceinline__ bool plan_flagged_leaf(const NodeRec* __restrict__ pairs, const int* __restrict__ parent_int,
                                                  const TreeHeader* __restrict__ hdr, int n, int s)
{
    const int p = __ldg(parent_int + s);
    if (p == WB_NO_PARENT)
        return (hdr->root_ref & WB_LEAF) != 0u;
    const int ps = p - n;
    return (pairs[2 * (size_t)ps + ((int)pairs[2 * (size_t)s + 1].aux == ps ? 0 : 1)].ref & WB_LEAF) != 0u;
}

__global__ void __launch_bounds__(BT)
k_plan_keys(int n, const NodeRec* __restrict__ pairs, const int* __restrict__ parent_int, const TreeHeader* __restrict__ hdr,
            const uint16_t* __restrict__ heights, uint32_t* __restrict__ keys, uint32_t* __restrict__ ghist)
{
    __shared__ uint32_t h[4 * 256];
    for (int k = threadIdx.x; k < 4 * 256; k += BT)
        h[k] = 0;
    __syncthreads();
    const int m = n - 1;
    const int stride = gridDim.x * BT;
    const int iters = (m + stride - 1) / stride;
    for (int it = 0; it < iters; ++it) {
        const int s = it * stride + blockIdx.x * BT + threadIdx.x;
        const bool valid = s < m;
        uint32_t key = 0;
        if (valid) {
            const int l = (int)pairs[2 * (size_t)s].aux, r = (int)pairs[2 * (size_t)s + 1].aux;
            // packed leaves and everything below a size leaf carry the flag themselves; below a DEPTH-rule leaf
            // (hdr->deep, bvh.cu:419-441) only the topmost node is flagged, so look for a flagged ancestor
            bool leaf = plan_flagged_leaf(pairs, parent_int, hdr, n, s);
            if (!leaf && hdr->deep) {
                for (int q = __ldg(parent_int + s); q != WB_NO_PARENT; q = __ldg(parent_int + (q - n)))
                    if (plan_flagged_leaf(pairs, parent_int, hdr, n, q - n)) {
                        leaf = true;
                        break;
                    }
            }
            if (leaf)
                key = WB_PLAN_SKIP;
            else if (l / WB_WAVE_BP != r / WB_WAVE_BP)
                key = WB_PLAN_TOP;
            else
                key = ((uint32_t)(l / WB_WAVE_BP) << WB_PLAN_HEIGHT_BITS) | min((uint32_t)__ldg(heights + s) & WB_HEIGHT_MASK, (1u << WB_PLAN_HEIGHT_BITS) - 1u);
            keys[s] = key;
        }
#pragma unroll
        for (int pass = 0; pass < 4; ++pass)
            hist_add_coherent(h + 256 * pass, (key >> (8 * pass)) & 255u, valid);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 4 * 256; k += BT)
        if (h[k])
            atomicAdd(&ghist[k], h[k]);
}

// after the sort: where each entry's box goes, and the [begin, end) run of every block
__global__ void __launch_bounds__(BT)
k_plan_finish(int n, const uint32_t* __restrict__ keys, const int* __restrict__ nodes, const NodeRec* __restrict__ pairs,
              const int* __restrict__ parent_int, uint32_t* __restrict__ dst, int* __restrict__ begin, int* __restrict__ end,
              uint32_t* __restrict__ top, int* __restrict__ ntop)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    const int m = n - 1;
    if (i >= m)
        return;
    const uint32_t key = keys[i];
    if (key >= WB_PLAN_TOP)
        return;
    const int s = nodes[i];
    const int p = __ldg(parent_int + s);
    if (p == WB_NO_PARENT) {
        dst[i] = WB_PLAN_DST_ROOT;
    } else {
        const int ps = p - n;
        const int side = ((int)pairs[2 * (size_t)s + 1].aux == ps) ? 0 : 1;
        const int pl = (int)pairs[2 * (size_t)ps].aux, pr = (int)pairs[2 * (size_t)ps + 1].aux;
        const bool spans = pl / WB_WAVE_BP != pr / WB_WAVE_BP;
        dst[i] = (uint32_t)(2 * ps + side) | (spans ? 0x80000000u : 0u);
        if (spans)
            top[atomicAdd(ntop, 1)] = (uint32_t)(2 * ps + side);
    }
    const uint32_t b = key >> WB_PLAN_HEIGHT_BITS;
    if (i == 0 || (keys[i - 1] >> WB_PLAN_HEIGHT_BITS) != b)
        begin[b] = i;
    const uint32_t next = (i + 1 < m) ? keys[i + 1] : WB_PLAN_SKIP;
    if (next >= WB_PLAN_TOP || (next >> WB_PLAN_HEIGHT_BITS) != b)
        end[b] = i + 1;
}

// visible leaves whose parent spans blocks announce themselves on the global counters
__global__ void __launch_bounds__(BT)
k_plan_units(int n, const int* __restrict__ pos_parent, const NodeRec* __restrict__ pairs, uint8_t* __restrict__ flags,
             uint32_t* __restrict__ top, int* __restrict__ ntop)
{
    const int i = blockIdx.x * BT + threadIdx.x;
    if (i >= n)
        return;
    const int p = __ldg(pos_parent + i);
    uint8_t f = 0;
    if (p >= 0) {
        const int ps = p - n;
        f = ((int)pairs[2 * (size_t)ps].aux / WB_WAVE_BP != (int)pairs[2 * (size_t)ps + 1].aux / WB_WAVE_BP) ? 1 : 0;
        if (f)
            top[atomicAdd(ntop, 1)] = (uint32_t)(2 * ps + (i <= ps ? 0 : 1));
    }
    flags[i] = f;
}

}  // namespace

const char* wb_refit_plan(BvhState& s, cudaStream_t stream)
{
    const int n = s.n, m = n - 1;
    if (m < 1)
        return nullptr;
    const int tiles = wb_div_up(m, rs_tile_for(m));
    const int nblocks = wb_div_up(n, WB_WAVE_BP);
    WB_CUDA_TRY(cudaMemsetAsync(s.ghist, 0, sizeof(uint32_t) * 8 * 256, stream));
    WB_CUDA_TRY(cudaMemsetAsync(s.tickets, 0, sizeof(unsigned) * 16, stream));
    WB_CUDA_TRY(cudaMemsetAsync(s.tile_status, 0, sizeof(uint32_t) * 256 * 4 * (size_t)tiles, stream));
    WB_CUDA_TRY(cudaMemsetAsync(s.plan_begin, 0, sizeof(int) * (size_t)nblocks, stream));
    WB_CUDA_TRY(cudaMemsetAsync(s.plan_end, 0, sizeof(int) * (size_t)nblocks, stream));
    WB_CUDA_TRY(cudaMemsetAsync(s.plan_ntop, 0, sizeof(int), stream));
    k_plan_keys<<<bounds_grid(m), BT, 0, stream>>>(n, s.pairs, s.parent_int, s.header, s.heights, s.plan_keys, s.ghist);
    // keys_alt / prim_alt are build scratch (>= 4 bytes per item), free between builds
    onesweep_sort<uint32_t>(s.plan_keys, (uint32_t*)s.keys_alt, s.plan_nodes, s.prim_alt, m, s.ghist, s.tile_status, s.tickets,
                            stream);
    k_plan_finish<<<wb_div_up(m, BT), BT, 0, stream>>>(n, s.plan_keys, s.plan_nodes, s.pairs, s.parent_int, s.plan_dst,
                                                       s.plan_begin, s.plan_end, s.plan_top, s.plan_ntop);
    k_plan_units<<<wb_div_up(n, BT), BT, 0, stream>>>(n, s.pos_parent, s.pairs, s.unit_flags, s.plan_top, s.plan_ntop);
    WB_CUDA_TRY(cudaGetLastError());
    s.plan_valid = true;
    return nullptr;
}

// ------------------------------------------------------------------------------------------------
// Morton ordering of query points (order.h)
// ------------------------------------------------------------------------------------------------
void wb_order_free(OrderScratch& ws)
{
    void* ptrs[] = { ws.keys, ws.keys_alt, ws.idx, ws.idx_alt, ws.ghist, ws.tile_status, ws.tickets, ws.partials, ws.hdr,
                     ws.packed, ws.sorted_pts };
    for (void* p : ptrs)
        if (p)
            cudaFree(p);
    ws = OrderScratch();
}

const char* wb_morton_order(OrderScratch& ws, const float* pts, long long n, cudaStream_t stream, bool hilbert)
{
    if (n <= 0)
        return nullptr;
    if (n >= (1ll << 30))
        return "query batches are ordered in chunks of fewer than 2^30 points";
    if (const char* e = wb_order_reserve(ws, n, stream))
        return e;
    const int ni = (int)n;
    const int tiles = wb_div_up(n, rs_tile_for(n));
    const int blocks = bounds_grid(n);
    const BoxSource src { pts, pts };  // a point is its own (degenerate) box; its centroid is the point itself
    k_scene_bounds<<<blocks, BT, 0, stream>>>(src, ni, ws.partials, ws.tickets, ws.ghist);
    WB_CUDA_TRY(cudaMemsetAsync(ws.tile_status, 0, sizeof(uint32_t) * 256 * 4 * (size_t)tiles, stream));
    // ordering needs no parity with anything: 24 key bits (a 256^3 grid, about one query per cell at 16 M queries)
    // order a batch as well as all 30 (16 / 18 / 21 / 24 / 30 bits: 598 / 655 / 695 / 700 / 696 M queries/s on C2), in
    // three digit passes instead of four.  Three passes end in the
    // second buffer pair, so the keys start in keys_alt and the permutation lands in ws.idx.
#ifndef WB_ORDER_SHIFT
#define WB_ORDER_SHIFT 6   // 30 - 6 = 24 key bits ...
#define WB_ORDER_PASSES 3  // ... in three 8-bit passes
#endif
    constexpr bool odd = (WB_ORDER_PASSES & 1) != 0;
    uint32_t* k0 = odd ? ws.keys_alt : ws.keys;
    uint32_t* k1 = odd ? ws.keys : ws.keys_alt;
    int* i0 = odd ? ws.idx_alt : ws.idx;
    int* i1 = odd ? ws.idx : ws.idx_alt;
    // curve: Hilbert (consecutive cells always adjacent: 715 vs 701 M closest-point queries/s on C2), or Morton for the
    // signed query, whose +x probes like the x-fastest Morton order better (118 vs 109 M queries/s)
    k_morton_hist<BoxSource, uint32_t, false><<<blocks, BT, 0, stream>>>(src, ni, ws.partials, blocks, ws.hdr, nullptr, k0,
                                                                         ws.ghist, hilbert ? -1 : WB_ORDER_SHIFT);
    onesweep_sort<uint32_t>(k0, k1, i0, i1, ni, ws.ghist, ws.tile_status, ws.tickets, stream, WB_ORDER_PASSES);
    WB_CUDA_TRY(cudaGetLastError());
    return nullptr;
}

// ------------------------------------------------------------------------------------------------
// ordering of rays: key = Morton code of the origin on a 64^3 grid (18 bits) above an octahedral 64 x 64
// direction cell (12 bits).  Like the point ordering this only decides which thread traces which ray.
// ------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ uint32_t spread3_6(uint32_t v)  // 6 bits -> every third bit
{
    v &= 63u;
    v = (v | (v << 8)) & 0x0000300fu;
    v = (v | (v << 4)) & 0x000030c3u;
    v = (v | (v << 2)) & 0x00009249u;
    return v;
}

__global__ void __launch_bounds__(BT)
k_ray_key_hist(const float* __restrict__ starts, const float* __restrict__ dirs, int n, const float* __restrict__ partials,
               int num_partials, uint32_t* __restrict__ keys, uint32_t* __restrict__ ghist)
{
    __shared__ uint32_t h[4 * 256];
    __shared__ float sm[BT / 32][6];
    for (int k = threadIdx.x; k < 4 * 256; k += BT)
        h[k] = 0;
    float3 glo, ghi;
    reduce_partials(partials, num_partials, sm, glo, ghi);
    const float ivx = 64.0f / ((ghi.x - glo.x) + 0.0001f), ivy = 64.0f / ((ghi.y - glo.y) + 0.0001f),
                ivz = 64.0f / ((ghi.z - glo.z) + 0.0001f);
    const int stride = gridDim.x * BT;
    const int iters = (n + stride - 1) / stride;
    for (int it = 0; it < iters; ++it) {
        const int i = it * stride + blockIdx.x * BT + threadIdx.x;
        const bool valid = i < n;
        uint32_t code = 0;
        if (valid) {
            const float ox = __ldg(starts + 3 * (size_t)i), oy = __ldg(starts + 3 * (size_t)i + 1), oz = __ldg(starts + 3 * (size_t)i + 2);
            const float dx = __ldg(dirs + 3 * (size_t)i), dy = __ldg(dirs + 3 * (size_t)i + 1), dz = __ldg(dirs + 3 * (size_t)i + 2);
            const uint32_t qx = (uint32_t)min(max((int)((ox - glo.x) * ivx), 0), 63);
            const uint32_t qy = (uint32_t)min(max((int)((oy - glo.y) * ivy), 0), 63);
            const uint32_t qz = (uint32_t)min(max((int)((oz - glo.z) * ivz), 0), 63);
            // octahedral map of the direction to [0,1]^2
            const float l1 = fabsf(dx) + fabsf(dy) + fabsf(dz);
            const float inv = l1 > 0.f ? 1.0f / l1 : 0.f;
            float u = dx * inv, v = dy * inv;
            if (dz < 0.f) {
                const float uu = (1.0f - fabsf(v)) * (u >= 0.f ? 1.f : -1.f), vv = (1.0f - fabsf(u)) * (v >= 0.f ? 1.f : -1.f);
                u = uu, v = vv;
            }
            const uint32_t du = (uint32_t)min(max((int)((u * 0.5f + 0.5f) * 64.0f), 0), 63);
            const uint32_t dv = (uint32_t)min(max((int)((v * 0.5f + 0.5f) * 64.0f), 0), 63);
            const uint32_t omort = (spread3_6(qz) << 2) | (spread3_6(qy) << 1) | spread3_6(qx);  // 18 bits
            uint32_t dmort = 0;                                                                   // 12 bits, 2-D interleave
#pragma unroll
            for (int b = 0; b < 6; ++b)
                dmort |= (((du >> b) & 1u) << (2 * b)) | (((dv >> b) & 1u) << (2 * b + 1));
            code = (omort << 12) | dmort;
            keys[i] = code;
        }
        if (valid)
            atomicAdd(&h[code & 255u], 1u);
#pragma unroll
        for (int p = 1; p < 4; ++p)
            hist_add_coherent(h + 256 * p, (code >> (8 * p)) & 255u, valid);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 4 * 256; k += BT)
        if (h[k])
            atomicAdd(&ghist[k], h[k]);
}

}  // namespace

const char* wb_order_reserve(OrderScratch& ws, long long n, cudaStream_t stream)
{
    if (n <= ws.capacity)
        return nullptr;
    if (stream) {  // growing allocates: not possible while the stream is being captured into a graph
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone)
            return "the query-ordering scratch of this stream is too small and cannot grow during graph capture: run the "
                   "query once before capturing (wp_cuda_graph_begin_capture sizes a new stream's scratch like the "
                   "largest one in use on the device)";
    }
    wb_order_free(ws);
    const size_t cap = (size_t)n;
    const size_t tiles = (size_t)wb_div_up(n, RS_THREADS * RS_ITEMS_SMALL);
    WB_CUDA_TRY(cudaMalloc(&ws.keys, 4 * cap));
    WB_CUDA_TRY(cudaMalloc(&ws.keys_alt, 4 * cap));
    WB_CUDA_TRY(cudaMalloc(&ws.idx, 4 * cap));
    WB_CUDA_TRY(cudaMalloc(&ws.idx_alt, 4 * cap));
    WB_CUDA_TRY(cudaMalloc(&ws.ghist, 4 * 8 * 256));
    WB_CUDA_TRY(cudaMalloc(&ws.tile_status, 4 * 256 * 4 * tiles));
    WB_CUDA_TRY(cudaMalloc(&ws.tickets, 4 * 16));
    WB_CUDA_TRY(cudaMemset(ws.tickets, 0, 4 * 16));
    WB_CUDA_TRY(cudaMalloc(&ws.partials, 4 * 6 * 4096));
    WB_CUDA_TRY(cudaMalloc(&ws.hdr, sizeof(TreeHeader)));
    WB_CUDA_TRY(cudaMalloc((void**)&ws.packed, 16 * cap));
    WB_CUDA_TRY(cudaMalloc((void**)&ws.sorted_pts, 12 * cap + 16));
    ws.capacity = n;
    return nullptr;
}

const char* wb_ray_order(OrderScratch& ws, const float* starts, const float* dirs, long long n, cudaStream_t stream)
{
    if (n <= 0)
        return nullptr;
    if (n >= (1ll << 30))
        return "ray batches are ordered in chunks of fewer than 2^30 rays";
    if (const char* e = wb_order_reserve(ws, n, stream))
        return e;
    const int ni = (int)n;
    const int tiles = wb_div_up(n, rs_tile_for(n));
    const int blocks = bounds_grid(n);
    const BoxSource src { starts, starts };
    k_scene_bounds<<<blocks, BT, 0, stream>>>(src, ni, ws.partials, ws.tickets, ws.ghist);
    WB_CUDA_TRY(cudaMemsetAsync(ws.tile_status, 0, sizeof(uint32_t) * 256 * 4 * (size_t)tiles, stream));
    k_ray_key_hist<<<blocks, BT, 0, stream>>>(starts, dirs, ni, ws.partials, blocks, ws.keys, ws.ghist);
    onesweep_sort<uint32_t>(ws.keys, ws.keys_alt, ws.idx, ws.idx_alt, ni, ws.ghist, ws.tile_status, ws.tickets, stream);
    WB_CUDA_TRY(cudaGetLastError());
    return nullptr;
}
