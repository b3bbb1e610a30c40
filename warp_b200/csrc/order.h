// Morton ordering of a batch of query points (bvh_build.cu): reuses the builder's scene-bounds,
// key and onesweep kernels to produce a permutation that makes neighbouring threads traverse
// neighbouring parts of the tree.  The permutation only changes WHICH thread answers a query, never
// the answer.
#pragma once
#include "common.cuh"

struct OrderScratch {
    uint32_t* keys = nullptr;
    uint32_t* keys_alt = nullptr;
    int* idx = nullptr;
    int* idx_alt = nullptr;
    uint32_t* ghist = nullptr;
    uint32_t* tile_status = nullptr;
    unsigned* tickets = nullptr;
    float* partials = nullptr;
    TreeHeader* hdr = nullptr;
    uint4* packed = nullptr;      // capacity records: results of an ordered batch before the unpack pass (query.cu QM_PACKED)
    float* sorted_pts = nullptr;  // 12 * capacity + 16 bytes: the batch in curve order (query.cu QM_STAGED)
    long long capacity = 0;
};

// after the call (stream-ordered) ws.idx[0..n) holds the query indices along a space-filling curve over the batch's
// bounds: the 24-bit Hilbert curve (hilbert = true) or the top 24 bits of the Morton code
const char* wb_morton_order(OrderScratch& ws, const float* pts, long long n, cudaStream_t stream, bool hilbert = false);
// same for rays: 18 bits of origin cell (64^3 grid over the origins' bounds) above 12 bits of direction cell
// (octahedral 64 x 64), so rays that start together and point the same way share a warp
const char* wb_ray_order(OrderScratch& ws, const float* starts, const float* dirs, long long n, cudaStream_t stream);
void wb_order_free(OrderScratch& ws);
// grow-only reservation for batches of up to n entries (allocates; fails with a message while `stream` is being captured)
const char* wb_order_reserve(OrderScratch& ws, long long n, cudaStream_t stream);
