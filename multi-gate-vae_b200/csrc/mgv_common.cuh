// Shared device/host helpers for the mgv_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/mgv_b200.h"

#define MGV_WARP 32

void mgv_set_error(const char* fmt, ...);
int mgv_check_cuda(cudaError_t e, const char* what);
void mgv_count_launches(int n);

#define MGV_CUDA(call)                                            \
    do {                                                          \
        int _rc = mgv_check_cuda((call), #call);                  \
        if (_rc != MGV_OK) return _rc;                            \
    } while (0)

#define MGV_REQUIRE(cond, ...)                                    \
    do {                                                          \
        if (!(cond)) {                                            \
            mgv_set_error(__VA_ARGS__);                           \
            return MGV_ERR_ARG;                                   \
        }                                                         \
    } while (0)

static inline size_t mgv_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct MgvArena {
    char* base;
    size_t cap, off;
    MgvArena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0) {}
    template <typename T>
    T* take(size_t count) {
        off = mgv_align_up(off, 256);
        T* r = (T*)(base + off);
        off += count * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Grid-wide barrier for persistent kernels launched with cudaLaunchCooperativeKernel (all CTAs
// co-resident).  `counter` is a monotonically increasing ticket counter, zeroed by the caller
// before the launch.  Thread 0 does release/acquire at gpu scope; the acquire side invalidates
// this SM's L1 so plain loads after the barrier observe other CTAs' stores.
__device__ __forceinline__ unsigned mgv_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mgv_grid_sync(unsigned* counter, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned ticket = atomicAdd(counter, 1u);
        unsigned target = (ticket / nblocks + 1u) * nblocks;
        while (mgv_ld_acquire(counter) < target) {
            __nanosleep(20);
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ float mgv_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float mgv_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float4 mgv_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void mgv_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 mgv_ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float mgv_dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ void mgv_fma4(float4& acc, float s, float4 v) {
    acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y);
    acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}

// out[r] += sum_k Xs[r*ldx + k] * Wt[k*ldw + col]   (r < ROWS, k < K, K % 4 == 0)
// Xs in shared memory (broadcast reads), Wt k-major in global memory (coalesced over col, read-only path).
template <int ROWS, int K>
__device__ __forceinline__ void mgv_gemm_col(const float* Xs, int ldx, const float* __restrict__ Wt, int ldw,
                                             int col, float (&acc)[ROWS]) {
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float w0 = __ldg(Wt + (size_t)(k + 0) * ldw + col);
        float w1 = __ldg(Wt + (size_t)(k + 1) * ldw + col);
        float w2 = __ldg(Wt + (size_t)(k + 2) * ldw + col);
        float w3 = __ldg(Wt + (size_t)(k + 3) * ldw + col);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            float4 x = mgv_ld4(Xs + r * ldx + k);
            acc[r] = fmaf(x.x, w0, acc[r]);
            acc[r] = fmaf(x.y, w1, acc[r]);
            acc[r] = fmaf(x.z, w2, acc[r]);
            acc[r] = fmaf(x.w, w3, acc[r]);
        }
    }
}
// Three output columns (col, col+64, col+128) of a [K][192] k-major matrix at once (GRU r,z,n of one unit).
template <int ROWS, int K>
__device__ __forceinline__ void mgv_gemm_col3(const float* Xs, int ldx, const float* __restrict__ Wt, int col,
                                              float (&ar)[ROWS], float (&az)[ROWS], float (&an)[ROWS]) {
#pragma unroll 1
    for (int k = 0; k < K; k += 4) {
        float wr[4], wz[4], wn[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float* p = Wt + (size_t)(k + i) * 192 + col;
            wr[i] = __ldg(p); wz[i] = __ldg(p + 64); wn[i] = __ldg(p + 128);
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            float4 x = mgv_ld4(Xs + r * ldx + k);
            ar[r] = fmaf(x.x, wr[0], ar[r]); ar[r] = fmaf(x.y, wr[1], ar[r]);
            ar[r] = fmaf(x.z, wr[2], ar[r]); ar[r] = fmaf(x.w, wr[3], ar[r]);
            az[r] = fmaf(x.x, wz[0], az[r]); az[r] = fmaf(x.y, wz[1], az[r]);
            az[r] = fmaf(x.z, wz[2], az[r]); az[r] = fmaf(x.w, wz[3], az[r]);
            an[r] = fmaf(x.x, wn[0], an[r]); an[r] = fmaf(x.y, wn[1], an[r]);
            an[r] = fmaf(x.z, wn[2], an[r]); an[r] = fmaf(x.w, wn[3], an[r]);
        }
    }
}
#endif  // __CUDACC__
