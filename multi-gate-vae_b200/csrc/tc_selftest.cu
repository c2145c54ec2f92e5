// Diagnostic entry point: one 128-row tcgen05 tile product through exactly the operand layouts,
// descriptors and fp16 hi/lo split the production kernels use (mgv_tc.cuh), in every operand
// orientation they need.  tests/test_gpu_tc.py checks each mode against a float64 matmul, so a wrong
// descriptor bit or swizzle shows up here and not as a parity failure deep inside a fused kernel.
#include "mgv_tc.cuh"

namespace {

constexpr int TM = 128;

// [rows x cols] fp32 row-major (leading dim ld) -> SW128 hi/lo planes, one 64-column block after the other
// (block stride = rows * 128 bytes).
__device__ void store_sw128(uint32_t hi_base, uint32_t lo_base, const float* src, int ld, int rows, int cols, int tid, int nthr) {
    const int chunks = cols / 8;
    for (int i = tid; i < rows * chunks; i += nthr) {
        const int r = i / chunks, c = i % chunks;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = src[(size_t)r * ld + c * 8 + e];
        uint4 hi, lo;
        tc::split8(v, hi, lo);
        const uint32_t off = (uint32_t)(c / 8) * rows * 128 + tc::sw128_off(r, c % 8);
        tc::st_shared_v4(hi_base + off, hi);
        tc::st_shared_v4(lo_base + off, lo);
    }
}
// [rows x 16] -> plain (no swizzle) hi/lo planes
__device__ void store_plain16(uint32_t hi_base, uint32_t lo_base, const float* src, int ld, int rows, int tid, int nthr) {
    for (int i = tid; i < rows * 2; i += nthr) {
        const int r = i / 2, c = i % 2;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = src[(size_t)r * ld + c * 8 + e];
        uint4 hi, lo;
        tc::split8(v, hi, lo);
        const uint32_t off = tc::plain16_off(r, c);
        tc::st_shared_v4(hi_base + off, hi);
        tc::st_shared_v4(lo_base + off, lo);
    }
}

__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(int mode, const float* __restrict__ A, const float* __restrict__ B,
                                                             float* __restrict__ D, int K, int N) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (tc::smem_u32(smem) + 1023u) & ~1023u;
    // plane sizes (bytes) per mode
    uint32_t a_bytes = 0, b_bytes = 0;
    if (mode == 0) { a_bytes = (uint32_t)(K / 64) * TM * 128; b_bytes = (uint32_t)(K / 64) * N * 128; }
    if (mode == 1) { a_bytes = 2u * TM * 128; b_bytes = (uint32_t)(N / 64) * TM * 128; }
    if (mode == 2) { a_bytes = TM * 32; b_bytes = (uint32_t)N * 32; }
    if (mode == 3) { a_bytes = 2u * TM * 128; b_bytes = TM * 32; }
    if (mode == 4) { a_bytes = 3u * TM * 128; b_bytes = 192u * 128; }
    if (mode == 5) { a_bytes = 0; b_bytes = 2u * 64 * 128; }
    a_bytes = (a_bytes + 1023u) & ~1023u;
    b_bytes = (b_bytes + 1023u) & ~1023u;
    const uint32_t a_hi = base, a_lo = base + a_bytes, b_hi = base + 2 * a_bytes, b_lo = b_hi + b_bytes;

    if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
    if (tid == 0) { tc::mbar_init(tc::smem_u32(&bar), 1); tc::fence_barrier_init(); }

    if (mode == 0) { store_sw128(a_hi, a_lo, A, K, TM, K, tid, 128); store_sw128(b_hi, b_lo, B, K, N, K, tid, 128); }
    if (mode == 1) { store_sw128(a_hi, a_lo, A, 128, TM, 128, tid, 128); store_sw128(b_hi, b_lo, B, N, TM, N, tid, 128); }
    if (mode == 2) { store_plain16(a_hi, a_lo, A, 16, TM, tid, 128); store_plain16(b_hi, b_lo, B, 16, N, tid, 128); }
    if (mode == 3) { store_sw128(a_hi, a_lo, A, 128, TM, 128, tid, 128); store_plain16(b_hi, b_lo, B, 16, TM, tid, 128); }
    if (mode == 4) { store_sw128(a_hi, a_lo, A, 192, TM, 192, tid, 128); store_sw128(b_hi, b_lo, B, 64, 192, 64, tid, 128); }
    // mode 5: A [128 x 64] lives in TENSOR MEMORY (fp16 hi/lo, 2 values per 32-bit column: K step j = columns
    // 16 j .. 16 j + 7 hi, 16 j + 8 .. 16 j + 15 lo, starting at column 128); B = W [64 x 128] as two SW128 blocks
    // (columns 0..63, 64..127) read MN-major with N = 128.
    if (mode == 5) { store_sw128(b_hi, b_lo, B, 128, 64, 64, tid, 128); store_sw128(b_hi + 8192, b_lo + 8192, B + 64, 128, 64, 64, tid, 128); }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (mode == 5) {
        const int row = warp * 32 + lane;
        const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + 128u;
        for (int j = 0; j < 4; ++j) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 8; ++e) tc::split2(A[row * 64 + 16 * j + 2 * e], A[row * 64 + 16 * j + 2 * e + 1], pk[e], pk[8 + e]);
            tc::tmem_st16(ta + 16u * j, pk);
        }
        tc::tmem_st_wait();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }

    if (tid == 0) {
        uint32_t acc = 0;
        if (mode == 0) {
            const uint32_t idesc = tc::make_idesc(128, N, false, false);
            for (int kb = 0; kb < K / 64; ++kb)
                for (int j = 0; j < 4; ++j) {
                    const uint32_t ao = (uint32_t)kb * TM * 128 + 32 * j, bo = (uint32_t)kb * N * 128 + 32 * j;
                    tc::mma3(tmem, tc::desc_k_sw128(a_hi + ao), tc::desc_k_sw128(a_lo + ao), tc::desc_k_sw128(b_hi + bo),
                             tc::desc_k_sw128(b_lo + bo), idesc, acc);
                    acc = 1;
                }
        } else if (mode == 1) {
            const uint32_t idesc = tc::make_idesc(128, N, true, true);
            for (int j = 0; j < TM / 16; ++j) {
                const uint32_t o = 2048u * j;
                tc::mma3(tmem, tc::desc_mn_sw128(a_hi + o, TM * 128), tc::desc_mn_sw128(a_lo + o, TM * 128),
                         tc::desc_mn_sw128(b_hi + o, TM * 128), tc::desc_mn_sw128(b_lo + o, TM * 128), idesc, acc);
                acc = 1;
            }
        } else if (mode == 2) {
            const uint32_t idesc = tc::make_idesc(128, N, false, false);
            tc::mma3(tmem, tc::desc_k_plain16(a_hi), tc::desc_k_plain16(a_lo), tc::desc_k_plain16(b_hi), tc::desc_k_plain16(b_lo), idesc, 0);
        } else if (mode == 3) {
            const uint32_t idesc = tc::make_idesc(128, 16, true, true);
            for (int j = 0; j < TM / 16; ++j) {
                const uint32_t ao = 2048u * j, bo = 512u * j;
                tc::mma3(tmem, tc::desc_mn_sw128(a_hi + ao, TM * 128), tc::desc_mn_sw128(a_lo + ao, TM * 128),
                         tc::desc_mn_plain16(b_hi + bo), tc::desc_mn_plain16(b_lo + bo), idesc, acc);
                acc = 1;
            }
        } else if (mode == 4) {
            const uint32_t idesc = tc::make_idesc(128, 64, false, true);
            for (int kk = 0; kk < 192; kk += 16) {
                const uint32_t ao = (uint32_t)(kk / 64) * TM * 128 + 2 * (kk % 64), bo = 1024u * (kk / 8);
                tc::mma3(tmem, tc::desc_k_sw128(a_hi + ao), tc::desc_k_sw128(a_lo + ao), tc::desc_mn_sw128(b_hi + bo, 0),
                         tc::desc_mn_sw128(b_lo + bo, 0), idesc, acc);
                acc = 1;
            }
        }
        else if (mode == 5) {
            const uint32_t idesc = tc::make_idesc(128, 128, false, true);
            for (int j = 0; j < 4; ++j) {
                const uint32_t bo = 2048u * j;
                tc::mma3p_ts<false>(tmem, tmem + 128u + 16u * j, tmem + 128u + 16u * j + 8u,
                                    tc::desc_mn_sw128(b_hi + bo, 8192), tc::desc_mn_sw128(b_lo + bo, 8192), idesc, acc);
                acc = 1;
            }
        }
        tc::mma_commit(tc::smem_u32(&bar));
    }
    tc::mbar_wait(tc::smem_u32(&bar), 0);
    tc::fence_after_sync();
    const int nout = (mode == 3) ? 16 : (mode == 4 ? 64 : (mode == 5 ? 128 : N));
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < nout; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e)
            if (c0 + e < nout) D[(size_t)row * nout + c0 + e] = v[e];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace

extern "C" int mgv_tc_selftest(int32_t mode, const float* A, const float* B, float* D, int32_t K, int32_t N, mgv_stream_t stream) {
    MGV_REQUIRE(mode >= 0 && mode <= 5, "tc selftest: mode 0..5");
    MGV_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "tc selftest: N must be a multiple of 16 in [16, 256]");
    MGV_REQUIRE(mode != 0 || (K >= 64 && K <= 128 && K % 64 == 0), "tc selftest: mode 0 needs K in {64, 128}");
    MGV_REQUIRE(mode != 1 || N % 64 == 0, "tc selftest: mode 1 needs N % 64 == 0");
    const size_t smem = 200 * 1024;
    MGV_CUDA(cudaFuncSetAttribute((const void*)tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mode, A, B, D, K, N);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_tc_selftest");
}
