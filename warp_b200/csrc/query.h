// batched query drivers (query.cu); pointers are device pointers, work is enqueued on `stream`
#pragma once
#include "common.cuh"

// perm: optional permutation (thread slot -> query index), e.g. from wb_morton_order
const char* wb_query_point(const TreeView& tv, const float* pts, const int* perm, long long nq, float max_dist,
                           int with_sign, uint8_t* result, float* sign, int* face, float* u, float* v,
                           unsigned long long* stats, cudaStream_t stream);
const char* wb_query_ray(const TreeView& tv, const float* starts, const float* dirs, const int* perm, long long nq,
                         float max_t, uint8_t* result, float* sign, int* face, float* t, float* u, float* v,
                         float* normal, unsigned long long* stats, cudaStream_t stream);
