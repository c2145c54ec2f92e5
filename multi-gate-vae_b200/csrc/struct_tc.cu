// Struct encoder on the 5th-generation tensor cores: MultiGCNEncoder.forward (digae_layer.py:257-277)
// with AggConv (arch/gcn_conv.py:30-42), one fused kernel per half-round step:
//   gather-sum of neighbour states -> GRU_{70->64}([W agg + deg b || x], state) -> LayerNorm
// Step k = 1..2R uses in-neighbours when k is odd (aggr/update) and out-neighbours when k is even
// (aggr_r/update_r); LayerNorm parameters are shared by both directions (digae_layer.py:270,275).
// source_conv and target_conv (digae_layer.py:294-297) run batched: blockIdx.y = encoder.
//
// A step is ONE tile product per 128 nodes (the AggConv linear is pre-composed into the GRU input weights
// on the host, Wc = W_ih[:, :64] W, bc = W_ih[:, :64] b):
//     [agg(64) | h(64) | x(8) deg 1 (16)]  x  [Wc | W_hh | W_ih_x bc b_ih b_hh]^T  ->  r, z, gi_n, gh_n   (256 fp32 columns)
// issued as tcgen05.mma kind::f16 on fp16 hi/lo planes (csrc/mgv_tc.cuh) with the accumulator in tensor
// memory; biases and the degree term ride along as two extra K columns, so the epilogue is gates + LayerNorm.
//
// Persistent CTAs (one per SM, weights resident in shared memory for the whole launch), warp-specialised:
//   warps 8-15  gather: neighbour sums, own state and [x deg 1] -> fp16 hi/lo -> operand tile (UMMA layout)
//   warp  8/l0  issues the MMAs of a tile once the tile is full; completion frees the tile (tcgen05.commit)
//   warps 0-7   epilogue: tensor memory -> GRU gates -> LayerNorm -> state_k   (two accumulator buffers, so the
//               epilogue of tile t overlaps the gather and MMAs of tile t+1)
#include "mgv_tc.cuh"

namespace {

constexpr int D = MGV_D;              // 64
constexpr int G3 = 3 * D;             // 192
constexpr int TM = 128;               // nodes per tile
constexpr int SPACK = MGV_STRUCT_PACK_FLOATS;
// natural fp32 weight block (include/mgv_b200.h)
constexpr int O_WCX = 0, O_WHH = 14592, O_BC = 27648, O_BIH = 27840, O_BHH = 28032, O_LNW = 28224, O_LNB = 28288;
constexpr int LDC = 76, LDM = 68;
constexpr int NODE_MASK = (1 << MGV_CODE_SHIFT) - 1;
constexpr float LN_EPS = 1e-5f;

// ---- shared-memory / weight-image layout (bytes).  The first IMG_W bytes are the per-(encoder, direction)
// weight image prepared by struct_image_kernel; it is copied verbatim.
constexpr uint32_t WC_HI = 0, WC_LO = 24576, WHH_HI = 49152, WHH_LO = 73728, WX_HI = 98304, WX_LO = 106496;
constexpr uint32_t IMG_W = 114688;                  // weight planes
constexpr uint32_t IMG_BYTES = IMG_W + 512;         // + ln_w[64], ln_b[64] fp32
constexpr uint32_t A_AGG_HI = IMG_W, A_AGG_LO = A_AGG_HI + 16384, A_H_HI = A_AGG_LO + 16384, A_H_LO = A_H_HI + 16384;
constexpr uint32_t A_X_HI = A_H_LO + 16384, A_X_LO = A_X_HI + 4096;
constexpr uint32_t S_LN = A_X_LO + 4096;            // 128 floats
constexpr uint32_t S_EX = S_LN + 512;               // LayerNorm partial sums [2][128][2] floats
constexpr uint32_t S_BAR = S_EX + 2048;             // 7 mbarriers
constexpr uint32_t S_TMEM = S_BAR + 64;
constexpr uint32_t F_SMEM = S_TMEM + 64 + 1024;     // + alignment slack

constexpr int THREADS = 512;
constexpr int EPI_WARPS = 8, GATHER_WARPS = 8;

struct StepTC {
    int N, feat, layernorm;
    const int* ptr;            // neighbour CSR of this step's direction
    const int* idx;
    const int* order;          // degree order of this direction (tile row -> node id)
    const unsigned* tile_cost; // [ntiles + 1] exclusive prefix of the tile cost model
    const float* x;            // [N][feat]
    const uint8_t* image;      // weight image of (enc 0, this dir); encoder stride 2 * IMG_BYTES
    const float* prev;         // state_{k-1}, enc 0
    float* next;               // state_k, enc 0
    size_t enc_stride;         // floats between encoders in the states buffer
};

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

__device__ __forceinline__ void ldg8(const float* p, float (&v)[8]) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void stg8(float* p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void split_store_sw128(uint32_t hi_base, uint32_t lo_base, int row, int c, const float (&v)[8]) {
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    const uint32_t off = tc::sw128_off(row, c);
    tc::st_shared_v4(hi_base + off, hi);
    tc::st_shared_v4(lo_base + off, lo);
}

// ======================================================================================= weight image
// natural fp32 block -> fp16 hi/lo planes in the UMMA layouts the step kernels consume.
__global__ void struct_image_kernel(const float* __restrict__ pack, uint8_t* __restrict__ image, int blocks) {
    const int blk = blockIdx.y;
    if (blk >= blocks) return;
    const float* W = pack + (size_t)blk * SPACK;
    uint8_t* img = image + (size_t)blk * IMG_BYTES;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float v[8];
    uint4 hi, lo;
    if (i < G3 * 8) {                                  // Wc: 192 rows x 8 chunks
        const int o = i >> 3, c = i & 7;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = W[O_WCX + o * LDC + c * 8 + e];
        tc::split8(v, hi, lo);
        const uint32_t off = tc::sw128_off(o, c);
        *reinterpret_cast<uint4*>(img + WC_HI + off) = hi;
        *reinterpret_cast<uint4*>(img + WC_LO + off) = lo;
    } else if (i < 2 * G3 * 8) {                       // Whh
        const int j = i - G3 * 8, o = j >> 3, c = j & 7;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = W[O_WHH + o * LDM + c * 8 + e];
        tc::split8(v, hi, lo);
        const uint32_t off = tc::sw128_off(o, c);
        *reinterpret_cast<uint4*>(img + WHH_HI + off) = hi;
        *reinterpret_cast<uint4*>(img + WHH_LO + off) = lo;
    } else if (i < 2 * G3 * 8 + 256 * 2) {             // [W_ih_x | bc | bias] : 256 accumulator columns x 16
        const int j = i - 2 * G3 * 8, n = j >> 1, c = j & 1;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (c == 0) {
            if (n < G3) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = W[O_WCX + n * LDC + D + e];
            }
        } else {
            if (n < 2 * D) { v[0] = W[O_BC + n]; v[1] = W[O_BIH + n] + W[O_BHH + n]; }     // r, z: merged biases
            else if (n < G3) { v[0] = W[O_BC + n]; v[1] = W[O_BIH + n]; }                   // gi_n
            else { v[1] = W[O_BHH + n - D]; }                                               // gh_n
        }
        tc::split8(v, hi, lo);
        const uint32_t off = tc::plain16_off(n, c);
        *reinterpret_cast<uint4*>(img + WX_HI + off) = hi;
        *reinterpret_cast<uint4*>(img + WX_LO + off) = lo;
    } else if (i < 2 * G3 * 8 + 256 * 2 + 2 * D) {
        const int j = i - (2 * G3 * 8 + 256 * 2);
        reinterpret_cast<float*>(img + IMG_W)[j] = W[O_LNW + j];      // ln_w then ln_b (contiguous in the block)
    }
}

// ======================================================================================= forward step
__global__ void __launch_bounds__(THREADS, 1) struct_fwd_tc_kernel(const StepTC p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - tc::smem_u32(smem_raw));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int enc = blockIdx.y;
    const float* prev = p.prev + (size_t)enc * p.enc_stride;
    float* next = p.next + (size_t)enc * p.enc_stride;
    const uint8_t* image = p.image + (size_t)enc * 2 * IMG_BYTES;

    const uint32_t bar_a_full = sbase + S_BAR, bar_a_empty = bar_a_full + 8;
    const uint32_t bar_acc_full = bar_a_full + 16, bar_acc_empty = bar_a_full + 32;      // [2] each
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + S_TMEM);
    float* s_ln = reinterpret_cast<float*>(sgen + S_LN);
    float* s_ex = reinterpret_cast<float*>(sgen + S_EX);

    // ---- one-time setup: barriers, weights -> shared memory (bulk async copy, overlaps the first gather), tensor memory
    const uint32_t bar_w = bar_a_full + 48;
    if (tid == 0) {
        tc::mbar_init(bar_a_full, GATHER_WARPS * 32);
        tc::mbar_init(bar_a_empty, 1);
        tc::mbar_init(bar_acc_full, 1);
        tc::mbar_init(bar_acc_full + 8, 1);
        tc::mbar_init(bar_acc_empty, EPI_WARPS * 32);
        tc::mbar_init(bar_acc_empty + 8, EPI_WARPS * 32);
        tc::mbar_init(bar_w, 1);
        tc::fence_barrier_init();
        tc::mbar_expect_tx(bar_w, IMG_W);
#pragma unroll 1
        for (uint32_t o = 0; o < IMG_W; o += 16384u) tc::bulk_g2s(sbase + o, image + o, 16384u, bar_w);
    }
    if (tid < 2 * D) s_ln[tid] = __ldg(reinterpret_cast<const float*>(image + IMG_W) + tid);
    if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
    // contiguous tile range of equal modelled cost for this CTA (tiles are in descending-degree order)
    const int ntiles = (p.N + TM - 1) / TM;
    int tile_beg, tile_end;
    {
        const unsigned long long total = p.tile_cost[ntiles];
        const unsigned lo = (unsigned)(total * blockIdx.x / gridDim.x), hi = (unsigned)(total * (blockIdx.x + 1) / gridDim.x);
        tile_beg = tc::warp_lower_bound(p.tile_cost, ntiles, lo, lane);
        tile_end = tc::warp_lower_bound(p.tile_cost, ntiles, hi, lane);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp >= EPI_WARPS) {
        // ===================================================================== gather (+ MMA issue)
        const int gw = warp - EPI_WARPS, rg = lane >> 3, c = lane & 7;
        int it = 0;
        for (int tile = tile_beg; tile < tile_end; ++tile, ++it) {
            const int t0 = tile * TM;
            // ---- prologue in registers (overlaps the previous tile's MMAs): node ids, CSR ranges, first neighbour,
            //      then the own-state rows, the lane's feature element and the first neighbour rows
            int node[4], beg[4], cnt[4], jn[4];
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                const int r = t0 + gw * 16 + ps * 4 + rg;
                node[ps] = r < p.N ? p.order[r] : -1;
            }
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                beg[ps] = 0; cnt[ps] = 0;
                if (node[ps] >= 0) { beg[ps] = p.ptr[node[ps]]; cnt[ps] = p.ptr[node[ps] + 1] - beg[ps]; }
            }
            int maxc = 0;
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                jn[ps] = cnt[ps] > 0 ? (p.idx[beg[ps]] & NODE_MASK) : 0;
                maxc = max(maxc, cnt[ps]);
            }
            float4 ha[4], hb[4], va[4], vb[4];
            float xe[4];
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                ha[ps] = make_float4(0.f, 0.f, 0.f, 0.f); hb[ps] = ha[ps];
                xe[ps] = 0.f;
                if (node[ps] >= 0) {
                    ha[ps] = mgv_ld4(prev + (size_t)node[ps] * D + c * 8);
                    hb[ps] = mgv_ld4(prev + (size_t)node[ps] * D + c * 8 + 4);
                    if (c < p.feat) xe[ps] = p.x[(size_t)node[ps] * p.feat + c];
                }
            }
            int j[4];
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                j[ps] = jn[ps];
                if (1 < cnt[ps]) jn[ps] = p.idx[beg[ps] + 1] & NODE_MASK;
            }
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                va[ps] = make_float4(0.f, 0.f, 0.f, 0.f); vb[ps] = va[ps];
                if (0 < cnt[ps]) {
                    va[ps] = mgv_ld4(prev + (size_t)j[ps] * D + c * 8);
                    vb[ps] = mgv_ld4(prev + (size_t)j[ps] * D + c * 8 + 4);
                }
            }
            tc::mbar_wait(bar_a_empty, (uint32_t)((it & 1) ^ 1));            // previous tile's MMAs have read the stage
            // ---- own state rows and the [x deg 1] block
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                const int row = gw * 16 + ps * 4 + rg;
                const float h8[8] = {ha[ps].x, ha[ps].y, ha[ps].z, ha[ps].w, hb[ps].x, hb[ps].y, hb[ps].z, hb[ps].w};
                split_store_sw128(sbase + A_H_HI, sbase + A_H_LO, row, c, h8);
                float xv[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) xv[e] = __shfl_sync(0xffffffffu, xe[ps], (lane & 24) + e);   // features of this row
                if (c == 1) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) xv[e] = 0.f;
                    if (node[ps] >= 0) { xv[0] = (float)cnt[ps]; xv[1] = 1.0f; }
                }
                if (c < 2) {
                    uint4 hi, lo;
                    tc::split8(xv, hi, lo);
                    const uint32_t off = tc::plain16_off(row, c);
                    tc::st_shared_v4(sbase + A_X_HI + off, hi);
                    tc::st_shared_v4(sbase + A_X_LO + off, lo);
                }
            }
            // ---- neighbour sums: one neighbour of each of the lane's 4 rows per trip (8 x 16-byte loads in flight),
            //      next trip's neighbour ids prefetched
            float acc[4][8];
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                acc[ps][0] = va[ps].x; acc[ps][1] = va[ps].y; acc[ps][2] = va[ps].z; acc[ps][3] = va[ps].w;
                acc[ps][4] = vb[ps].x; acc[ps][5] = vb[ps].y; acc[ps][6] = vb[ps].z; acc[ps][7] = vb[ps].w;
            }
            for (int sl = 1; sl < maxc; ++sl) {
#pragma unroll
                for (int ps = 0; ps < 4; ++ps) {
                    j[ps] = jn[ps];
                    if (sl + 1 < cnt[ps]) jn[ps] = p.idx[beg[ps] + sl + 1] & NODE_MASK;
                }
#pragma unroll
                for (int ps = 0; ps < 4; ++ps) {
                    va[ps] = make_float4(0.f, 0.f, 0.f, 0.f); vb[ps] = va[ps];
                    if (sl < cnt[ps]) {
                        va[ps] = mgv_ld4(prev + (size_t)j[ps] * D + c * 8);
                        vb[ps] = mgv_ld4(prev + (size_t)j[ps] * D + c * 8 + 4);
                    }
                }
#pragma unroll
                for (int ps = 0; ps < 4; ++ps) {
                    acc[ps][0] += va[ps].x; acc[ps][1] += va[ps].y; acc[ps][2] += va[ps].z; acc[ps][3] += va[ps].w;
                    acc[ps][4] += vb[ps].x; acc[ps][5] += vb[ps].y; acc[ps][6] += vb[ps].z; acc[ps][7] += vb[ps].w;
                }
            }
#pragma unroll
            for (int ps = 0; ps < 4; ++ps)
                split_store_sw128(sbase + A_AGG_HI, sbase + A_AGG_LO, gw * 16 + ps * 4 + rg, c, acc[ps]);
            tc::fence_async_smem();
            tc::mbar_arrive(bar_a_full);
            if (warp == EPI_WARPS && lane == 0) {
                const int b = it & 1;
                if (it == 0) tc::mbar_wait(bar_w, 0u);                              // weight image has landed
                tc::mbar_wait(bar_a_full, (uint32_t)(it & 1));
                tc::mbar_wait(bar_acc_empty + 8 * b, (uint32_t)(((it >> 1) & 1) ^ 1));   // epilogue drained this buffer
                tc::fence_after_sync();
                const uint32_t d = tmem + (uint32_t)b * 256u;
                // [x deg 1] block first: initialises all 256 columns (biases, degree term, feature term)
                tc::mma3(d, tc::desc_k_plain16(sbase + A_X_HI), tc::desc_k_plain16(sbase + A_X_LO),
                         tc::desc_k_plain16(sbase + WX_HI), tc::desc_k_plain16(sbase + WX_LO), tc::make_idesc(128, 256, false, false), 0u);
#pragma unroll
                for (int j = 0; j < 4; ++j)      // agg . Wc^T -> r, z, gi_n
                    tc::mma3(d, tc::desc_k_sw128(sbase + A_AGG_HI + 32 * j), tc::desc_k_sw128(sbase + A_AGG_LO + 32 * j),
                             tc::desc_k_sw128(sbase + WC_HI + 32 * j), tc::desc_k_sw128(sbase + WC_LO + 32 * j),
                             tc::make_idesc(128, 192, false, false), 1u);
#pragma unroll
                for (int j = 0; j < 4; ++j)      // h . Whh[r, z]^T -> r, z
                    tc::mma3(d, tc::desc_k_sw128(sbase + A_H_HI + 32 * j), tc::desc_k_sw128(sbase + A_H_LO + 32 * j),
                             tc::desc_k_sw128(sbase + WHH_HI + 32 * j), tc::desc_k_sw128(sbase + WHH_LO + 32 * j),
                             tc::make_idesc(128, 128, false, false), 1u);
#pragma unroll
                for (int j = 0; j < 4; ++j)      // h . Whh[n]^T -> gh_n
                    tc::mma3(d + 192u, tc::desc_k_sw128(sbase + A_H_HI + 32 * j), tc::desc_k_sw128(sbase + A_H_LO + 32 * j),
                             tc::desc_k_sw128(sbase + WHH_HI + 16384 + 32 * j), tc::desc_k_sw128(sbase + WHH_LO + 16384 + 32 * j),
                             tc::make_idesc(128, 64, false, false), 1u);
                tc::mma_commit(bar_a_empty);
                tc::mma_commit(bar_acc_full + 8 * b);
            }
            __syncwarp();
        }
    } else {
        // ===================================================================== epilogue
        const int q = warp & 3, half = warp >> 2;
        const int row = q * 32 + lane, ubase = half * 32;
        int it = 0;
        for (int tile = tile_beg; tile < tile_end; ++tile, ++it) {
            const int b = it & 1;
            const bool valid = tile * TM + row < p.N;
            const int node = valid ? p.order[tile * TM + row] : 0;
            float h[32];
            if (valid) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) ldg8(prev + (size_t)node * D + ubase + 8 * ch, *reinterpret_cast<float(*)[8]>(&h[8 * ch]));
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) h[e] = 0.f;
            }
            tc::mbar_wait(bar_acc_full + 8 * b, (uint32_t)((it >> 1) & 1));
            tc::fence_after_sync();
            const uint32_t ta = tmem + (uint32_t)b * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)ubase;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float gr[8], gz[8], gi[8], gh[8];
                tmem_ld8(ta + 8 * ch, gr);
                tmem_ld8(ta + 64 + 8 * ch, gz);
                tmem_ld8(ta + 128 + 8 * ch, gi);
                tmem_ld8(ta + 192 + 8 * ch, gh);
                tc::tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float r = fast_sigmoid(gr[e]), z = fast_sigmoid(gz[e]);
                    const float n = fast_tanh(fmaf(r, gh[e], gi[e]));
                    h[8 * ch + e] = fmaf(z, h[8 * ch + e] - n, n);          // (1 - z) n + z h
                }
            }
            tc::fence_before_sync();
            tc::mbar_arrive(bar_acc_empty + 8 * b);
            if (p.layernorm) {
                float s = 0.f;
#pragma unroll
                for (int e = 0; e < 32; ++e) s += h[e];
                s_ex[row * 2 + half] = s;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float mean = (s_ex[row * 2] + s_ex[row * 2 + 1]) * (1.0f / D);
                float v = 0.f;
#pragma unroll
                for (int e = 0; e < 32; ++e) { h[e] -= mean; v = fmaf(h[e], h[e], v); }
                s_ex[256 + row * 2 + half] = v;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float rstd = rsqrtf((s_ex[256 + row * 2] + s_ex[256 + row * 2 + 1]) * (1.0f / D) + LN_EPS);
#pragma unroll
                for (int e = 0; e < 32; ++e) h[e] = fmaf(h[e] * rstd, s_ln[ubase + e], s_ln[D + ubase + e]);
            }
            if (valid) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) stg8(next + (size_t)node * D + ubase + 8 * ch, *reinterpret_cast<float(*)[8]>(&h[8 * ch]));
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

__global__ void fill_ones_kernel(float* p, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 1.0f;
}

}  // namespace

size_t mgv_struct_image_bytes(int num_enc) { return mgv_align_up((size_t)num_enc * 2 * IMG_BYTES, 256); }

int mgv_struct_build_image(const float* weights, int num_enc, uint8_t* image, cudaStream_t st) {
    const int items = 2 * G3 * 8 + 256 * 2 + 2 * D;
    struct_image_kernel<<<dim3((items + 255) / 256, num_enc * 2), 256, 0, st>>>(weights, image, num_enc * 2);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "struct_image_kernel");
}

extern "C" size_t mgv_struct_fwd_workspace_bytes(int64_t N, int32_t num_enc) {
    (void)N;
    return mgv_struct_image_bytes(num_enc) + 1024;
}

extern "C" int mgv_struct_encoder_fwd(const mgv_schedule* sch, int32_t num_enc, int32_t rounds, int32_t layernorm,
                                      int32_t feat, const float* x, const float* weights, float* states,
                                      void* ws, size_t ws_bytes, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MGV_REQUIRE(sch != nullptr, "struct encoder: null schedule");
    MGV_REQUIRE(num_enc >= 1 && num_enc <= 2, "struct encoder: num_enc must be 1 or 2");
    MGV_REQUIRE(rounds >= 1, "struct encoder: rounds must be >= 1");
    MGV_REQUIRE(feat >= 0 && feat <= MGV_MAX_FEAT, "struct encoder: dim_feature %d > %d", feat, MGV_MAX_FEAT);
    const int N = sch->N;
    if (N == 0) return MGV_OK;
    MGV_REQUIRE(sch->deg_order_in && sch->deg_order_out && sch->tile_cost_in && sch->tile_cost_out,
                "struct encoder: the schedule carries no degree order (mgv_build_degree_order)");
    if (ws_bytes < mgv_struct_fwd_workspace_bytes(N, num_enc)) {
        mgv_set_error("mgv_struct_encoder_fwd: workspace %zu < %zu bytes", ws_bytes, mgv_struct_fwd_workspace_bytes(N, num_enc));
        return MGV_ERR_WORKSPACE;
    }
    MgvArena a(ws, ws_bytes);
    uint8_t* image = a.take<uint8_t>((size_t)num_enc * 2 * IMG_BYTES);
    int rc = mgv_struct_build_image(weights, num_enc, image, st);
    if (rc != MGV_OK) return rc;
    const int steps = 2 * rounds;
    const size_t slot = (size_t)N * D;
    const size_t enc_stride = (size_t)(steps + 1) * slot;
    MGV_CUDA(cudaFuncSetAttribute((const void*)struct_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM));
    int dev = 0, sms = 0;
    MGV_CUDA(cudaGetDevice(&dev));
    MGV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int ntiles = (N + TM - 1) / TM;
    int gx = sms / num_enc;
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    for (int e = 0; e < num_enc; ++e) {
        fill_ones_kernel<<<(unsigned)((slot + 255) / 256), 256, 0, st>>>(states + e * enc_stride, slot);
        mgv_count_launches(1);
    }
    for (int k = 1; k <= steps; ++k) {
        const int dir = (k & 1) ? 0 : 1;
        StepTC p{};
        p.N = N; p.feat = feat; p.layernorm = layernorm;
        p.ptr = dir == 0 ? sch->in_ptr : sch->out_ptr;
        p.idx = dir == 0 ? sch->in_src : sch->out_pack;
        p.order = dir == 0 ? sch->deg_order_in : sch->deg_order_out;
        p.tile_cost = dir == 0 ? sch->tile_cost_in : sch->tile_cost_out;
        p.x = x;
        p.image = image + (size_t)dir * IMG_BYTES;
        p.prev = states + (size_t)(k - 1) * slot;
        p.next = states + (size_t)k * slot;
        p.enc_stride = enc_stride;
        struct_fwd_tc_kernel<<<dim3(gx, num_enc), THREADS, F_SMEM, st>>>(p);
        mgv_count_launches(1);
    }
    return mgv_check_cuda(cudaGetLastError(), "mgv_struct_encoder_fwd");
}
