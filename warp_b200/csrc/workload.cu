// Synthetic-workload generators for bench.py / tests: query sets and per-frame cloth vertices made ON
// THE DEVICE from a counter-based RNG, so the 1 B-query config (C5) needs no host generation or upload and the
// 1000-frame cloth loop (C4) can be one CUDA graph per frame.  Not part of the query path; everything is enqueued
// on the device's current stream and is capture-safe (no allocation, no synchronisation).
#include "../../include/warp_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace {

__host__ __device__ inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// uniform in [0, 1): top 24 bits of the hash
__host__ __device__ inline float u01(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }

__global__ void k_gen_box_queries(float* __restrict__ out, long long n, long long first, uint64_t seed, float lx, float ly,
                                  float lz, float hx, float hy, float hz)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const uint64_t g = (uint64_t)(first + i);
    const uint64_t base = splitmix64(seed ^ (g * 0xD1342543DE82EF95ull));
    const float a = u01(splitmix64(base + 1)), b = u01(splitmix64(base + 2)), c = u01(splitmix64(base + 3));
    out[3 * i + 0] = lx + a * (hx - lx);
    out[3 * i + 1] = ly + b * (hy - ly);
    out[3 * i + 2] = lz + c * (hz - lz);
}

__device__ inline float cloth_z(float x, float y, int frame)
{
    return 0.05f * sinf(12.0f * x + 0.05f * (float)frame) * cosf(9.0f * y + 0.03f * (float)frame);
}

// vertices of an n x n cloth grid on [0,1]^2 at frame *frame (+ offset): z = 0.05 sin(12x + 0.05f) cos(9y + 0.03f)
__global__ void k_gen_cloth_points(float* __restrict__ points, int n, const int* __restrict__ frame, int offset)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * n)
        return;
    const int f = *frame + offset;
    const int ix = (int)(i / n), iy = (int)(i % n);
    const float x = (float)ix / (float)(n - 1), y = (float)iy / (float)(n - 1);
    points[3 * i + 0] = x;
    points[3 * i + 1] = y;
    points[3 * i + 2] = cloth_z(x, y, f);
}

// queries of frame f = *frame: a random cloth vertex AT FRAME f - 1, jittered by N(0, sigma) per axis (seed 5 + f)
__global__ void k_gen_cloth_queries(float* __restrict__ out, long long nq, int n, const int* __restrict__ frame, float sigma)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq)
        return;
    const int f = *frame;
    const uint64_t base = splitmix64((uint64_t)(5 + f) * 0x9E3779B97F4A7C15ull ^ ((uint64_t)i * 0xD1342543DE82EF95ull));
    const uint64_t pick = splitmix64(base + 1) % ((uint64_t)n * (uint64_t)n);
    const int ix = (int)(pick / (uint64_t)n), iy = (int)(pick % (uint64_t)n);
    const float x = (float)ix / (float)(n - 1), y = (float)iy / (float)(n - 1);
    // Box-Muller, two uniforms per pair of normals
    const float u1 = fmaxf(u01(splitmix64(base + 2)), 1.0e-7f), u2 = u01(splitmix64(base + 3));
    const float u3 = fmaxf(u01(splitmix64(base + 4)), 1.0e-7f), u4 = u01(splitmix64(base + 5));
    const float r1 = sqrtf(-2.0f * logf(u1)), r2 = sqrtf(-2.0f * logf(u3));
    out[3 * i + 0] = x + sigma * r1 * cosf(6.2831853f * u2);
    out[3 * i + 1] = y + sigma * r1 * sinf(6.2831853f * u2);
    out[3 * i + 2] = cloth_z(x, y, f - 1) + sigma * r2 * cosf(6.2831853f * u4);
}

__global__ void k_counter_add(int* counter, int value) { *counter += value; }

}  // namespace

extern "C" {
void* wp_cuda_context_get_stream(void* context);

int wp_b200_gen_box_queries(float* out, int64_t n, int64_t first_index, uint64_t seed, const float* lower, const float* upper)
{
    if (n <= 0)
        return 1;
    cudaStream_t st = (cudaStream_t)wp_cuda_context_get_stream(nullptr);
    k_gen_box_queries<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, n, first_index, seed, lower[0], lower[1], lower[2],
                                                                  upper[0], upper[1], upper[2]);
    return cudaGetLastError() == cudaSuccess;
}

int wp_b200_gen_cloth_points(float* points, int n_side, const int* frame_dev, int frame_offset)
{
    const long long nv = (long long)n_side * n_side;
    if (nv <= 0)
        return 1;
    cudaStream_t st = (cudaStream_t)wp_cuda_context_get_stream(nullptr);
    k_gen_cloth_points<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(points, n_side, frame_dev, frame_offset);
    return cudaGetLastError() == cudaSuccess;
}

int wp_b200_gen_cloth_queries(float* out, int64_t nq, int n_side, const int* frame_dev, float sigma)
{
    if (nq <= 0)
        return 1;
    cudaStream_t st = (cudaStream_t)wp_cuda_context_get_stream(nullptr);
    k_gen_cloth_queries<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(out, nq, n_side, frame_dev, sigma);
    return cudaGetLastError() == cudaSuccess;
}

int wp_b200_counter_add(int* counter_dev, int value)
{
    cudaStream_t st = (cudaStream_t)wp_cuda_context_get_stream(nullptr);
    k_counter_add<<<1, 1, 0, st>>>(counter_dev, value);
    return cudaGetLastError() == cudaSuccess;
}
}
