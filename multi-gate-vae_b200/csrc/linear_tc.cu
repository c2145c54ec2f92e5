// Forward and data gradient of the small nn.Linear layers around the path (hs_linear 128 -> 64, hs_decompose 64 -> 128,
// dg_ae_model_mig.py:46-47; the four 64 -> 64 layers of DirectedGVAE.sample, digvae_model.py:105-142) on the 5th-generation
// tensor cores:
//     out[N][P] = in[N][Q] . M[P][Q]^T (+ bias[P])      M = W (forward: P = out features, Q = in features)
//                                                       M = W^T (data gradient: P = in features, Q = out features, no bias)
// N = nodes of the batch (10^4 .. 10^7), P, Q in {64, 128}.  The product is HBM-bound (4 (P + Q) bytes per node); a library runs it as
// an fp32 SIMT GEMM at ~25 TFLOP/s (37 us at N = 65 818 for 128 -> 64).  Here: persistent CTAs, 128-node tiles, loader warps
// split the rows into fp16 hi/lo planes in the UMMA layout (mgv_tc.cuh: three products per K step, fp32-accurate), one thread
// issues tcgen05.mma into a double-buffered TMEM accumulator, four epilogue warps add the bias and store.  Every input ROW is
// scaled by its own power of two (row maximum -> [2^8, 2^9)) before the split and un-scaled in the epilogue (row = TMEM lane =
// thread): gradient rows of a mean loss over 10^5 nodes sit at 1e-7, deep in fp16's subnormals.
#include "mgv_tc.cuh"

namespace {

constexpr int TM = 128;                       // rows per tile = UMMA M
constexpr int LOAD_WARPS = 16, EPI_WARPS = 4;
constexpr int LT_THREADS = (LOAD_WARPS + 1 + EPI_WARPS) * 32;
constexpr uint32_t A_PLANE = 2 * TM * 128;    // one fp16 plane of a tile at Q = 128: 2 K blocks of [128 rows][64 columns]
constexpr uint32_t W_PLANE = 2 * 128 * 128;   // one fp16 plane of the weight image at P = Q = 128
constexpr uint32_t S_W_HI = 0, S_W_LO = W_PLANE, S_A = 2 * W_PLANE;              // then 2 buffers x (hi, lo)
constexpr uint32_t S_BAR = S_A + 4 * A_PLANE, S_TMEM = S_BAR + 64, S_BIAS = S_TMEM + 16, S_SCALE = S_BIAS + 128 * 4;      // row scales [4][128]
constexpr uint32_t S_TOTAL = S_SCALE + 4 * TM * 4 + 1024;
static_assert(S_TOTAL <= 227 * 1024, "linear_tc: shared memory");

struct LinTC {
    const float* in; const float* W; const float* bias; float* out;
    long long N; int P, Q, transposed;
};

__global__ void __launch_bounds__(LT_THREADS, 1) linear_tc_kernel(const LinTC p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - tc::smem_u32(smem_raw));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_a_full = sbase + S_BAR, bar_a_empty = bar_a_full + 16, bar_acc_full = bar_a_full + 32, bar_acc_empty = bar_a_full + 48;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + S_TMEM);
    float* s_bias = reinterpret_cast<float*>(sgen + S_BIAS);
    float* s_scale = reinterpret_cast<float*>(sgen + S_SCALE);
    const int P = p.P, Q = p.Q, kbs = Q / 64;
    const uint32_t w_kb = (uint32_t)P * 128u;            // bytes of one K block of a weight plane
    const long long ntiles = (p.N + TM - 1) / TM;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(bar_a_full + 8 * b, LOAD_WARPS * 32);
            tc::mbar_init(bar_a_empty + 8 * b, 1);
            tc::mbar_init(bar_acc_full + 8 * b, 1);
            tc::mbar_init(bar_acc_empty + 8 * b, EPI_WARPS * 32);
        }
        tc::fence_barrier_init();
    }
    if (warp == LOAD_WARPS) tc::tmem_alloc(tmem_slot, 256);
    // weight image: M[p][q] as fp16 hi / lo planes, K-major SW128, K blocks of 64 columns (all threads)
    for (int idx = tid; idx < P * Q / 8; idx += LT_THREADS) {
        const int r = idx / (Q / 8), c8 = idx % (Q / 8), q0 = 8 * c8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = p.transposed ? __ldg(p.W + (size_t)(q0 + e) * P + r) : __ldg(p.W + (size_t)r * Q + q0 + e);
        uint4 hi, lo;
        tc::split8(v, hi, lo);
        const uint32_t off = (uint32_t)(c8 >> 3) * w_kb + tc::sw128_off(r, c8 & 7);
        tc::st_shared_v4(sbase + S_W_HI + off, hi);
        tc::st_shared_v4(sbase + S_W_LO + off, lo);
    }
    for (int i = tid; i < 128; i += LT_THREADS) s_bias[i] = (p.bias != nullptr && i < P) ? __ldg(p.bias + i) : 0.f;
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp < LOAD_WARPS) {
        // ===================================================================== loaders: rows -> hi / lo planes of the tile buffer
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int b = it & 1;
            tc::mbar_wait_warp(bar_a_empty + 8 * b, (uint32_t)(((it >> 1) & 1) ^ 1), lane, 0);     // the MMAs have read this buffer
            const uint32_t a_hi = sbase + S_A + (uint32_t)b * 2 * A_PLANE, a_lo = a_hi + A_PLANE;
            const long long r0 = tile * TM;
            const int chunks = TM * Q / 8;
            float4 va[4], vb[4];
            // four 32-byte chunks per thread in flight (chunks of a tile are contiguous in memory: fully coalesced)
            for (int c0 = tid; c0 < chunks; c0 += 4 * LOAD_WARPS * 32) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = c0 + j * LOAD_WARPS * 32;
                    const int r = c / (Q / 8);
                    va[j] = make_float4(0.f, 0.f, 0.f, 0.f); vb[j] = va[j];
                    if (c < chunks && r0 + r < p.N) {
                        const float4* src = reinterpret_cast<const float4*>(p.in + (size_t)(r0 + r) * Q) + 2 * (c % (Q / 8));
                        va[j] = __ldg(src); vb[j] = __ldg(src + 1);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = c0 + j * LOAD_WARPS * 32;
                    // the Q / 8 chunks of a row sit in consecutive lanes: row maximum by shuffles inside that lane group
                    float amax = fmaxf(fmaxf(fmaxf(fabsf(va[j].x), fabsf(va[j].y)), fmaxf(fabsf(va[j].z), fabsf(va[j].w))),
                                       fmaxf(fmaxf(fabsf(vb[j].x), fabsf(vb[j].y)), fmaxf(fabsf(vb[j].z), fabsf(vb[j].w))));
                    for (int o = 1; o < Q / 8; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
                    float sc = 1.0f;
                    if (amax > 0.f && isfinite(amax)) {
                        int k = 8 - ((int)((__float_as_uint(amax) >> 23) & 0xff) - 127);
                        k = k < -100 ? -100 : (k > 100 ? 100 : k);
                        sc = __uint_as_float((uint32_t)(k + 127) << 23);
                    }
                    if (c < chunks) {
                        const int r = c / (Q / 8), c8 = c % (Q / 8);
                        if (c8 == 0) s_scale[(it & 3) * TM + r] = sc;
                        const float v[8] = {va[j].x * sc, va[j].y * sc, va[j].z * sc, va[j].w * sc, vb[j].x * sc, vb[j].y * sc, vb[j].z * sc, vb[j].w * sc};
                        uint4 hi, lo;
                        tc::split8(v, hi, lo);
                        const uint32_t off = (uint32_t)(c8 >> 3) * (TM * 128u) + tc::sw128_off(r, c8 & 7);
                        tc::st_shared_v4(a_hi + off, hi);
                        tc::st_shared_v4(a_lo + off, lo);
                    }
                }
            }
            tc::fence_async_smem();
            tc::mbar_arrive(bar_a_full + 8 * b);
        }
    } else if (warp == LOAD_WARPS) {
        // ===================================================================== MMA issue (one thread)
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc(TM, P, false, false);
            int it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int b = it & 1;
                tc::mbar_wait(bar_acc_empty + 8 * b, (uint32_t)(((it >> 1) & 1) ^ 1));     // the epilogue drained this accumulator
                tc::mbar_wait(bar_a_full + 8 * b, (uint32_t)((it >> 1) & 1));
                tc::fence_after_sync();
                const uint32_t a_hi = sbase + S_A + (uint32_t)b * 2 * A_PLANE, a_lo = a_hi + A_PLANE;
                const uint32_t d = tmem + (uint32_t)b * 128u;
                for (int kb = 0; kb < kbs; ++kb)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        tc::mma3(d, tc::desc_k_sw128(a_hi + kb * (TM * 128u) + 32 * j), tc::desc_k_sw128(a_lo + kb * (TM * 128u) + 32 * j),
                                 tc::desc_k_sw128(sbase + S_W_HI + kb * w_kb + 32 * j), tc::desc_k_sw128(sbase + S_W_LO + kb * w_kb + 32 * j),
                                 idesc, (kb | j) ? 1u : 0u);
                tc::mma_commit(bar_a_empty + 8 * b);
                tc::mma_commit(bar_acc_full + 8 * b);
            }
        }
    } else {
        // ===================================================================== epilogue: thread = tile row = TMEM lane
        const int q = warp & 3;
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int b = it & 1;
            tc::mbar_wait_warp(bar_acc_full + 8 * b, (uint32_t)((it >> 1) & 1), lane, 0);
            tc::fence_after_sync();
            const long long row = tile * TM + q * 32 + lane;
            float* dst = p.out + (size_t)row * P;
            const float un = 1.0f / s_scale[(it & 3) * TM + q * 32 + lane];
#pragma unroll 1
            for (int c = 0; c < P; c += 16) {
                float v[16];
                tc::tmem_ld16(tl + (uint32_t)b * 128u + c, v);
                tc::tmem_ld_wait();
                if (row < p.N) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4)
                        *reinterpret_cast<float4*>(dst + c + e) = make_float4(fmaf(v[e], un, s_bias[c + e]), fmaf(v[e + 1], un, s_bias[c + e + 1]),
                                                                              fmaf(v[e + 2], un, s_bias[c + e + 2]), fmaf(v[e + 3], un, s_bias[c + e + 3]));
                }
            }
            tc::fence_before_sync();
            tc::mbar_arrive(bar_acc_empty + 8 * b);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == LOAD_WARPS) tc::tmem_dealloc(tmem, 256);
}

}  // namespace

// out[N][P] = in[N][Q] . M[P][Q]^T (+ bias);  transposed = 0: M = W given as [P][Q] (forward of y = x W^T + b),
// transposed = 1: M = W^T with W given as [Q][P] (data gradient gx = gy W).  P, Q in {64, 128}; bias may be NULL.
extern "C" int mgv_linear_tc(const float* in, int64_t N, const float* W, const float* bias, int32_t P, int32_t Q, int32_t transposed,
                             float* out, mgv_stream_t stream) {
    MGV_REQUIRE(N >= 0 && W && out && (N == 0 || in), "mgv_linear_tc: bad argument");
    MGV_REQUIRE((P == 64 || P == 128) && (Q == 64 || Q == 128), "mgv_linear_tc: P and Q must be 64 or 128 (got %d, %d)", P, Q);
    MGV_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "mgv_linear_tc: in / out must be 16-byte aligned");
    if (N == 0) return MGV_OK;
    int dev = 0, sms = 0;
    MGV_CUDA(cudaGetDevice(&dev));
    MGV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    LinTC p{in, W, bias, out, (long long)N, P, Q, transposed ? 1 : 0};
    const long long ntiles = (N + TM - 1) / TM;
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    MGV_CUDA(cudaFuncSetAttribute((const void*)linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_TOTAL));
    linear_tc_kernel<<<grid, LT_THREADS, S_TOTAL, (cudaStream_t)stream>>>(p);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_linear_tc");
}
