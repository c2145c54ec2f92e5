// Struct encoder BACKWARD (the forward lives in struct_tc.cu on tcgen05): MultiGCNEncoder.forward
// (digae_layer.py:257-277) with AggConv (arch/gcn_conv.py:30-42) differentiated step by step in reverse.
// One fused kernel per half-round step k = 2R .. 1 (blockIdx.y = encoder):
//   gather   neighbour sum of state_{k-1} (recompute) and of the neighbours' d agg_{k+1} (+ d part) = d state_k
//   GEMM     recompute the GRU pre-activations  [agg | x | deg 1] Wcx^T,  h Whh^T
//   pointwise LayerNorm backward, GRU backward -> d gi / d gh  (scaled per tile by a power of two into fp16 range)
//   GEMM     d agg = d gi Wc,  d part = g z + d gh Whh         (consumed by step k-1)
//   GEMM     d Wcx += d gi^T [agg | x | deg 1],  d Whh += d gh^T h   (persistent register fragments, flushed once)
// All products are mma.sync m16n8k16 on fp16 hi/lo planes (mgv_mma16.cuh): tiles and weights are split once when
// written to shared memory; biases ride along as the "1" column of the [agg | x | deg 1] tile, so d bc / d b_ih /
// d b_hh fall out of the weight-gradient product.
#include "mgv_mma16.cuh"

long long* mgv_debug_trace();

namespace {

constexpr int D = MGV_D;              // 64
constexpr int G3 = 3 * D;             // 192
constexpr int SPACK = MGV_STRUCT_PACK_FLOATS;
constexpr int SGRAD = MGV_STRUCT_GRAD_FLOATS;
// natural fp32 weight / gradient block (include/mgv_b200.h)
constexpr int O_WCX = 0, O_WHH = 14592, O_BC = 27648, O_BIH = 27840, O_BHH = 28032, O_LNW = 28224, O_LNB = 28288;
constexpr int NLDC = 76, NLDM = 68;
constexpr int NODE_MASK = (1 << MGV_CODE_SHIFT) - 1;
constexpr float LN_EPS = 1e-5f;

constexpr int THREADS = 512;
constexpr int WARPS = THREADS / 32;   // 16
constexpr int TM = 32;                // nodes per tile
constexpr int KX = 80;                // [agg 64 | x 8 | deg | 1 | 0 x 6]
constexpr int LDC = 88, LDW = 72, LDG = 264, LDF = 68;   // row strides: halves (planes) / floats (LDF)
// shared memory (bytes)
constexpr uint32_t WCX_HI = 0, WCX_LO = WCX_HI + G3 * LDC * 2, WHH_HI = WCX_LO + G3 * LDC * 2, WHH_LO = WHH_HI + G3 * LDW * 2;
constexpr uint32_t S_BHH = WHH_LO + G3 * LDW * 2, S_LN = S_BHH + G3 * 4;
constexpr uint32_t AS_HI = S_LN + 2 * D * 4, AS_LO = AS_HI + TM * LDC * 2, HS_HI = AS_LO + TM * LDC * 2, HS_LO = HS_HI + TM * LDW * 2;
constexpr uint32_t S_H32 = HS_LO + TM * LDW * 2, S_GS = S_H32 + TM * LDF * 4, S_XH = S_GS + TM * LDF * 4;
constexpr uint32_t DG_HI = S_XH + TM * LDF * 4, DG_LO = DG_HI + TM * LDG * 2;
constexpr uint32_t S_LNA = DG_LO + TM * LDG * 2, S_MAX = S_LNA + 2 * D * 4;
constexpr uint32_t B_SMEM = S_MAX + 16;
static_assert(WCX_LO % 16 == 0 && WHH_HI % 16 == 0 && AS_HI % 16 == 0 && HS_HI % 16 == 0 && DG_HI % 16 == 0 && DG_LO % 16 == 0, "plane alignment");
static_assert(B_SMEM <= 227 * 1024, "struct backward: shared memory");

struct StepDev {
    int N, feat, layernorm, first, last, dir;
    const int* ptr;            // neighbour CSR of this step's direction
    const int* idx;
    const float* x;            // [N][feat]
    const float* weights;      // natural block of (enc 0, this dir); encoder stride 2*SPACK
    const float* prev;         // state_{k-1}, enc 0
    size_t enc_stride;         // floats between encoders in the states buffer
    const float* gout;         // [enc][N][64]
    const float* in_part; const float* in_agg;
    float* out_part; float* out_agg;     // [enc][N][64]
    float* partial;            // [enc][gx][2][SGRAD]
    long long* trace;          // optional [CTA][16 tiles][16] clock64 samples (dev tool)
};
#define BTRACE(slot) do { if (p.trace && tid == 0 && it < 16) p.trace[(((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + it) * 16 + (slot)] = clock64(); } while (0)

__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// One half-warp per node: sum of the neighbours' rows of `a` (and of `b` over the same neighbours).
template <bool TWO>
__device__ __forceinline__ void gather_sum(const StepDev& p, const float* a, const float* b, int node, int l16,
                                           float4& sa, float4& sb, int& deg) {
    const int beg = p.ptr[node], end = p.ptr[node + 1];
    deg = end - beg;
    sa = make_float4(0.f, 0.f, 0.f, 0.f);
    sb = sa;
    for (int q0 = beg; q0 < end; q0 += 4) {
        float4 va[4], vb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            vb[i] = va[i];
            if (q0 + i < end) {
                const int j = p.idx[q0 + i] & NODE_MASK;
                va[i] = mgv_ld4(a + (size_t)j * D + 4 * l16);
                if (TWO) vb[i] = mgv_ld4(b + (size_t)j * D + 4 * l16);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            add4(sa, va[i]);
            if (TWO) add4(sb, vb[i]);
        }
    }
}

// GRU gates with 5 MUFU ops per unit (same formulation as the forward kernel, struct_tc.cu).
__device__ __forceinline__ void gru_gates(float gr, float gz, float gi, float gh, float& r, float& z, float& n) {
    const float a = __expf(-fminf(fmaxf(gr, -28.f), 28.f));
    const float b = __expf(-fminf(fmaxf(gz, -28.f), 28.f));
    const float inv = __fdividef(1.0f, (1.0f + a) * (1.0f + b));
    r = (1.0f + b) * inv;
    z = (1.0f + a) * inv;
    const float y = fminf(fmaxf(fmaf(r, gh, gi), -14.f), 14.f);
    n = 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * y));
}

template <bool LOWP>
__device__ __forceinline__ void st_plane4(uint8_t* hi_plane, uint8_t* lo_plane, uint32_t half_off, float a, float b, float c, float d) {
    uint2 hi, lo;
    m16::split2p<LOWP>(a, b, hi.x, lo.x);
    m16::split2p<LOWP>(c, d, hi.y, lo.y);
    *reinterpret_cast<uint2*>(hi_plane + half_off * 2) = hi;
    if (!LOWP) *reinterpret_cast<uint2*>(lo_plane + half_off * 2) = lo;
}

// ======================================================================================= backward step
template <bool LOWP>
__global__ void __launch_bounds__(THREADS, 1) struct_bwd_kernel(const StepDev p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sb = m16::smem_u32(smem);
    float* Bhh = reinterpret_cast<float*>(smem + S_BHH);
    float* Ln = reinterpret_cast<float*>(smem + S_LN);
    float* H32 = reinterpret_cast<float*>(smem + S_H32);
    float* Gs = reinterpret_cast<float*>(smem + S_GS);          // d state_k, then d (pre-LN GRU output)
    float* Xh = reinterpret_cast<float*>(smem + S_XH);          // pre-LN GRU output
    float* LNA = reinterpret_cast<float*>(smem + S_LNA);        // d ln_w, d ln_b accumulators
    unsigned* smax = reinterpret_cast<unsigned*>(smem + S_MAX);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int enc = blockIdx.y;
    const float* prev = p.prev + (size_t)enc * p.enc_stride;
    const size_t eoff = (size_t)enc * p.N * D;
    const float* W = p.weights + (size_t)enc * 2 * SPACK;

    // ---- weights -> fp16 hi/lo planes (split once per launch)
    for (int i = tid; i < G3 * 10; i += THREADS) {
        const int o = i / 10, c = i % 10;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (c < 9) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __ldg(W + O_WCX + o * NLDC + c * 8 + e);       // c == 8: feature columns 64..71
        } else {
            v[0] = __ldg(W + O_BC + o);
            v[1] = __ldg(W + O_BIH + o);
        }
        const uint32_t off = (uint32_t)(o * LDC + c * 8);
        st_plane4<LOWP>(smem + WCX_HI, smem + WCX_LO, off, v[0], v[1], v[2], v[3]);
        st_plane4<LOWP>(smem + WCX_HI, smem + WCX_LO, off + 4, v[4], v[5], v[6], v[7]);
    }
    for (int i = tid; i < G3 * 8; i += THREADS) {
        const int o = i >> 3, c = i & 7;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __ldg(W + O_WHH + o * NLDM + c * 8 + e);
        const uint32_t off = (uint32_t)(o * LDW + c * 8);
        st_plane4<LOWP>(smem + WHH_HI, smem + WHH_LO, off, v[0], v[1], v[2], v[3]);
        st_plane4<LOWP>(smem + WHH_HI, smem + WHH_LO, off + 4, v[4], v[5], v[6], v[7]);
    }
    if (tid < G3) Bhh[tid] = __ldg(W + O_BHH + tid);
    if (tid < 2 * D) { Ln[tid] = __ldg(W + O_LNW + tid); LNA[tid] = 0.f; }
    if (tid < 2) smax[tid] = 0u;
    const int ntiles = (p.N + TM - 1) / TM;
    const int half = lane >> 4, l16 = lane & 15;
    const int mt = warp & 1, u0 = (warp >> 1) * 8;
    const int g = lane >> 2, t = lane & 3;

    // persistent weight-gradient fragments (scaled by acc_scale):
    //   d Wcx[:, 0:64] : 6 m-tiles (96 gate rows per warp half) x 8 feature columns per warp
    //   d Whh          : 2 + 4 m-tiles (gate rows -> d gh columns: r, z contiguous, n at column 192)
    //   feature / bias : gate-gradient columns 16 warp .. +15  x  tile columns 64..71 (x) and 72..79 (deg, 1)
    const int wh = warp >> 3;
    const int wn0[1] = {8 * (warp & 7)};
    const int fn0[2] = {D, D + 8};
    float acc_cx[6][1][4], acc_ha[2][1][4], acc_hb[4][1][4], acc_fb[1][2][4];
    m16::zero_frag(acc_cx);
    m16::zero_frag(acc_ha);
    m16::zero_frag(acc_hb);
    m16::zero_frag(acc_fb);
    float acc_scale = 1.0f;
    float lnw_acc[2] = {0.f, 0.f}, lnb_acc[2] = {0.f, 0.f};     // d ln_w / d ln_b of columns lane, lane + 32 (rows of this warp)
    __syncthreads();

    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int t0 = tile * TM;
        BTRACE(0);
        {   // ---- gathers: neighbour sum of state_{k-1}; d state_k = part + sum of neighbours' d agg_{k+1}
            const int row = warp * 2 + half, node = t0 + row;
            float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sbv = sa, h4 = sa, g4 = sa, x4 = sa;
            int deg = 0;
            if (node < p.N) {
                if (p.last) {
                    gather_sum<false>(p, prev, nullptr, node, l16, sa, sbv, deg);
                    g4 = mgv_ld4(p.gout + eoff + (size_t)node * D + 4 * l16);
                } else {
                    gather_sum<true>(p, prev, p.in_agg + eoff, node, l16, sa, sbv, deg);
                    g4 = mgv_ld4(p.in_part + eoff + (size_t)node * D + 4 * l16);
                    add4(g4, sbv);
                }
                h4 = mgv_ld4(prev + (size_t)node * D + 4 * l16);
                if (l16 < 2) {
                    float xv[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) xv[e] = (4 * l16 + e < p.feat) ? p.x[(size_t)node * p.feat + 4 * l16 + e] : 0.f;
                    x4 = make_float4(xv[0], xv[1], xv[2], xv[3]);
                } else if (l16 == 2) {
                    x4 = make_float4((float)deg, 1.0f, 0.f, 0.f);
                }
            }
            st_plane4<LOWP>(smem + AS_HI, smem + AS_LO, (uint32_t)(row * LDC + 4 * l16), sa.x, sa.y, sa.z, sa.w);
            if (l16 < 4) st_plane4<LOWP>(smem + AS_HI, smem + AS_LO, (uint32_t)(row * LDC + D + 4 * l16), x4.x, x4.y, x4.z, x4.w);
            st_plane4<LOWP>(smem + HS_HI, smem + HS_LO, (uint32_t)(row * LDW + 4 * l16), h4.x, h4.y, h4.z, h4.w);
            mgv_st4(H32 + row * LDF + 4 * l16, h4);
            mgv_st4(Gs + row * LDF + 4 * l16, g4);
        }
        __syncthreads();
        BTRACE(1);
        // ---- recompute the step (gates stay in registers)
        float gr[4], gz[4], gn[4], hnb[4];
        {
            const int n0[3] = {u0, D + u0, 2 * D + u0};
            float ci[1][3][4], ch[1][3][4];
            m16::zero_frag(ci);
            m16::zero_frag(ch);
            m16::warp_gemm<1, 3, KX / 16, false, false, LOWP>(ci, sb + AS_HI, sb + AS_LO, LDC, mt * 16, 0, sb + WCX_HI, sb + WCX_LO, LDC, n0, 0, lane);
            m16::warp_gemm<1, 3, D / 16, false, false, LOWP>(ch, sb + HS_HI, sb + HS_LO, LDW, mt * 16, 0, sb + WHH_HI, sb + WHH_LO, LDW, n0, 0, lane);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int u = u0 + 2 * t + (e & 1);
                hnb[e] = ch[0][2][e] + Bhh[2 * D + u];
                gru_gates(ci[0][0][e] + ch[0][0][e] + Bhh[u], ci[0][1][e] + ch[0][1][e] + Bhh[D + u], ci[0][2][e], hnb[e],
                          gr[e], gz[e], gn[e]);
            }
        }
        BTRACE(2);
        if (p.layernorm) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int row = mt * 16 + g + ((e & 2) ? 8 : 0);
                const int u = u0 + 2 * t + (e & 1);
                Xh[row * LDF + u] = (1.0f - gz[e]) * gn[e] + gz[e] * H32[row * LDF + u];
            }
            __syncthreads();
            // LayerNorm backward, two rows per warp; d ln_w / d ln_b accumulate in shared memory
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int row = warp * 2 + rr;
                const float v0 = Xh[row * LDF + lane], v1 = Xh[row * LDF + 32 + lane];
                const float mean = mgv_warp_sum(v0 + v1) * (1.0f / D);
                const float d0 = v0 - mean, d1 = v1 - mean;
                const float var = mgv_warp_sum(d0 * d0 + d1 * d1) * (1.0f / D);
                const float rstd = rsqrtf(var + LN_EPS);
                const float x0 = d0 * rstd, x1 = d1 * rstd;
                const float gy0 = Gs[row * LDF + lane], gy1 = Gs[row * LDF + 32 + lane];
                lnw_acc[0] = fmaf(gy0, x0, lnw_acc[0]); lnw_acc[1] = fmaf(gy1, x1, lnw_acc[1]);
                lnb_acc[0] += gy0; lnb_acc[1] += gy1;
                const float dx0 = gy0 * Ln[lane], dx1 = gy1 * Ln[32 + lane];
                const float c1 = mgv_warp_sum(dx0 + dx1) * (1.0f / D);
                const float c2 = mgv_warp_sum(dx0 * x0 + dx1 * x1) * (1.0f / D);
                Gs[row * LDF + lane] = rstd * (dx0 - c1 - x0 * c2);
                Gs[row * LDF + 32 + lane] = rstd * (dx1 - c1 - x1 * c2);
            }
            __syncthreads();
        }
        BTRACE(3);
        // ---- GRU backward on the owned elements; tile-wide power-of-two scale for the fp16 planes
        float dr[4], dz[4], dni[4], dnh[4], dh_direct[4];
        {
            float amax = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int row = mt * 16 + g + ((e & 2) ? 8 : 0);
                const int u = u0 + 2 * t + (e & 1);
                const float gg = Gs[row * LDF + u];
                const float hp = H32[row * LDF + u];
                const float dn = gg * (1.0f - gz[e]);
                const float dzz = gg * (hp - gn[e]);
                dni[e] = dn * (1.0f - gn[e] * gn[e]);
                dr[e] = dni[e] * hnb[e] * gr[e] * (1.0f - gr[e]);
                dz[e] = dzz * gz[e] * (1.0f - gz[e]);
                dnh[e] = dni[e] * gr[e];
                dh_direct[e] = gg * gz[e];
                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(dr[e]), fabsf(dz[e])), fabsf(dni[e])));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            if (lane == 0) atomicMax(smax + (it & 1), __float_as_uint(amax));
            if (tid == 0) smax[(it + 1) & 1] = 0u;
        }
        __syncthreads();
        const float scale = m16::pow2_scale_keep(__uint_as_float(smax[it & 1]), acc_scale);
        const float inv_scale = 1.0f / scale;
        {
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int row = mt * 16 + g + 8 * hrow, e0 = 2 * hrow;
                const uint32_t off = (uint32_t)(row * LDG + u0 + 2 * t) * 2;
                uint32_t hi, lo;
                m16::split2p<LOWP>(dr[e0] * scale, dr[e0 + 1] * scale, hi, lo);
                *reinterpret_cast<uint32_t*>(smem + DG_HI + off) = hi; if (!LOWP) *reinterpret_cast<uint32_t*>(smem + DG_LO + off) = lo;
                m16::split2p<LOWP>(dz[e0] * scale, dz[e0 + 1] * scale, hi, lo);
                *reinterpret_cast<uint32_t*>(smem + DG_HI + off + 2 * D) = hi; if (!LOWP) *reinterpret_cast<uint32_t*>(smem + DG_LO + off + 2 * D) = lo;
                m16::split2p<LOWP>(dni[e0] * scale, dni[e0 + 1] * scale, hi, lo);
                *reinterpret_cast<uint32_t*>(smem + DG_HI + off + 4 * D) = hi; if (!LOWP) *reinterpret_cast<uint32_t*>(smem + DG_LO + off + 4 * D) = lo;
                m16::split2p<LOWP>(dnh[e0] * scale, dnh[e0 + 1] * scale, hi, lo);
                *reinterpret_cast<uint32_t*>(smem + DG_HI + off + 6 * D) = hi; if (!LOWP) *reinterpret_cast<uint32_t*>(smem + DG_LO + off + 6 * D) = lo;
            }
        }
        __syncthreads();
        BTRACE(4);
        // ---- data gradients: d agg = d gi . Wc ;  d part = g z + d gh . Whh   (both 32 x 64, K = 192)
        if (!p.first) {
            const int n0[1] = {u0};
            float ca[1][1][4], cp[1][1][4];
            m16::zero_frag(ca);
            m16::zero_frag(cp);
            m16::warp_gemm<1, 1, G3 / 16, false, true, LOWP>(ca, sb + DG_HI, sb + DG_LO, LDG, mt * 16, 0, sb + WCX_HI, sb + WCX_LO, LDC, n0, 0, lane);
            m16::warp_gemm<1, 1, 2 * D / 16, false, true, LOWP>(cp, sb + DG_HI, sb + DG_LO, LDG, mt * 16, 0, sb + WHH_HI, sb + WHH_LO, LDW, n0, 0, lane);
            m16::warp_gemm<1, 1, D / 16, false, true, LOWP>(cp, sb + DG_HI, sb + DG_LO, LDG, mt * 16, 3 * D, sb + WHH_HI, sb + WHH_LO, LDW, n0, 2 * D, lane);
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int node = t0 + mt * 16 + g + 8 * hrow;
                if (node < p.N) {
                    const size_t o = eoff + (size_t)node * D + u0 + 2 * t;
                    *reinterpret_cast<float2*>(p.out_agg + o) = make_float2(ca[0][0][2 * hrow] * inv_scale, ca[0][0][2 * hrow + 1] * inv_scale);
                    *reinterpret_cast<float2*>(p.out_part + o) =
                        make_float2(fmaf(cp[0][0][2 * hrow], inv_scale, dh_direct[2 * hrow]),
                                    fmaf(cp[0][0][2 * hrow + 1], inv_scale, dh_direct[2 * hrow + 1]));
                }
            }
        }
        BTRACE(5);
        // ---- weight gradients (tile buffers are read-only here)
        if (scale != acc_scale) {
            const float f = scale / acc_scale;
            m16::scale_frag(acc_cx, f);
            m16::scale_frag(acc_ha, f);
            m16::scale_frag(acc_hb, f);
            m16::scale_frag(acc_fb, f);
            acc_scale = scale;
        }
        m16::warp_gemm<6, 1, TM / 16, true, true, LOWP>(acc_cx, sb + DG_HI, sb + DG_LO, LDG, wh * 96, 0, sb + AS_HI, sb + AS_LO, LDC, wn0, 0, lane);
        m16::warp_gemm<2, 1, TM / 16, true, true, LOWP>(acc_ha, sb + DG_HI, sb + DG_LO, LDG, wh * 96, 0, sb + HS_HI, sb + HS_LO, LDW, wn0, 0, lane);
        m16::warp_gemm<4, 1, TM / 16, true, true, LOWP>(acc_hb, sb + DG_HI, sb + DG_LO, LDG, wh ? 3 * D : 32, 0, sb + HS_HI, sb + HS_LO, LDW, wn0, 0, lane);
        m16::warp_gemm<1, 2, TM / 16, true, true, LOWP>(acc_fb, sb + DG_HI, sb + DG_LO, LDG, 16 * warp, 0, sb + AS_HI, sb + AS_LO, LDC, fn0, 0, lane);
        BTRACE(6);
        __syncthreads();
        BTRACE(7);
    }
    // ---- flush this CTA's accumulators into its private partial block (summed over the launches of a call)
    float* part = p.partial + (((size_t)enc * gridDim.x + blockIdx.x) * 2 + p.dir) * SGRAD;
    const float un = 1.0f / acc_scale;
#pragma unroll
    for (int m = 0; m < 6; ++m) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int r8 = g + ((e & 2) ? 8 : 0), c = wn0[0] + 2 * t + (e & 1);
            part[O_WCX + (wh * 96 + 16 * m + r8) * NLDC + c] += acc_cx[m][0][e] * un;
            const int oh = wh * 96 + 16 * m + r8;                              // Whh gate row of fragment m
            part[O_WHH + oh * NLDM + c] += (m < 2 ? acc_ha[m][0][e] : acc_hb[m - 2][0][e]) * un;
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int oc = 16 * warp + g + ((e & 2) ? 8 : 0);                      // d gate column 0..255
        const int c = 2 * t + (e & 1);
        if (oc < G3) {
            part[O_WCX + oc * NLDC + D + c] += acc_fb[0][0][e] * un;           // feature columns
            if (c == 0) part[O_BC + oc] += acc_fb[0][1][e] * un;
            if (c == 1) {
                part[O_BIH + oc] += acc_fb[0][1][e] * un;
                if (oc < 2 * D) part[O_BHH + oc] += acc_fb[0][1][e] * un;      // r, z: d b_hh = d b_ih
            }
        } else if (c == 1) {
            part[O_BHH + oc - D] += acc_fb[0][1][e] * un;                      // n: sum of d gh_n
        }
    }
    atomicAdd(LNA + lane, lnw_acc[0]);
    atomicAdd(LNA + 32 + lane, lnw_acc[1]);
    atomicAdd(LNA + D + lane, lnb_acc[0]);
    atomicAdd(LNA + D + 32 + lane, lnb_acc[1]);
    __syncthreads();
    if (tid < 2 * D) part[O_LNW + tid] += LNA[tid];
}

__global__ void struct_reduce_kernel(const float* __restrict__ partial, int gx, float* __restrict__ grads, int num_enc) {
    const size_t total = (size_t)num_enc * 2 * SGRAD;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int enc = (int)(idx / (2 * SGRAD));
    const size_t rem = idx % (2 * (size_t)SGRAD);
    float s = 0.f;
    for (int b = 0; b < gx; ++b) s += partial[((size_t)enc * gx + b) * 2 * SGRAD + rem];
    grads[idx] = s;
}

int persistent_gx(int num_enc, int N, int* gx_out) {
    int dev = 0, sms = 0;
    MGV_CUDA(cudaGetDevice(&dev));
    MGV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int gx = sms / (num_enc > 0 ? num_enc : 1);
    const int ntiles = (N + TM - 1) / TM;
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    *gx_out = gx;
    return MGV_OK;
}

int check_args(const mgv_schedule* sch, int num_enc, int rounds, int feat) {
    MGV_REQUIRE(sch != nullptr, "struct encoder: null schedule");
    MGV_REQUIRE(num_enc >= 1 && num_enc <= 2, "struct encoder: num_enc must be 1 or 2");
    MGV_REQUIRE(rounds >= 1, "struct encoder: rounds must be >= 1");
    MGV_REQUIRE(feat >= 0 && feat <= MGV_MAX_FEAT, "struct encoder: dim_feature %d > %d", feat, MGV_MAX_FEAT);
    return MGV_OK;
}

void fill_step(StepDev& p, const mgv_schedule* sch, int k, int steps, int layernorm, int feat, const float* x,
               const float* weights, const float* states, size_t slot, size_t enc_stride) {
    const int dir = (k & 1) ? 0 : 1;
    p.N = sch->N; p.feat = feat; p.layernorm = layernorm; p.first = (k == 1); p.last = (k == steps); p.dir = dir;
    p.ptr = dir == 0 ? sch->in_ptr : sch->out_ptr;
    p.idx = dir == 0 ? sch->in_src : sch->out_pack;
    p.x = x;
    p.weights = weights + (size_t)dir * SPACK;
    p.prev = states + (size_t)(k - 1) * slot;
    p.enc_stride = enc_stride;
}

}  // namespace

int mgv_struct_bwd_legacy_grid(void) {
    int gx = 0;
    if (persistent_gx(1, 1 << 30, &gx) != MGV_OK) return -1;
    return gx;
}

size_t mgv_struct_bwd_legacy_workspace_bytes(int64_t N, int32_t num_enc) {
    int gx = mgv_struct_bwd_legacy_grid();
    if (gx < 1) gx = 256;
    size_t b = 0;
    b += 4 * mgv_align_up((size_t)num_enc * N * D * 4 + 256, 256);              // part/agg ping-pong
    b += mgv_align_up((size_t)gx * 2 * SGRAD * 4 + 256, 256);                    // partial: num_enc * (sms / num_enc) <= sms CTAs
    return b + 1024;
}

// mma.sync path: precision = 1 (bf16 planes), and precision = 0 when MGV_STRUCT_BWD=mma is set (A/B checks against the
// tcgen05 path in struct_bwd_tc.cu, which owns the public entry point).
int mgv_struct_encoder_bwd_legacy(const mgv_schedule* sch, int32_t num_enc, int32_t rounds, int32_t layernorm,
                                      int32_t feat, const float* x, const float* weights, const float* states,
                                      const float* gout, float* grads, void* ws, size_t ws_bytes,
                                      int32_t precision, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_args(sch, num_enc, rounds, feat);
    if (rc != MGV_OK) return rc;
    const int N = sch->N;
    MGV_CUDA(cudaMemsetAsync(grads, 0, (size_t)num_enc * 2 * SGRAD * sizeof(float), st));
    if (N == 0) return MGV_OK;
    if (ws_bytes < mgv_struct_bwd_legacy_workspace_bytes(N, num_enc)) {
        mgv_set_error("mgv_struct_encoder_bwd: workspace %zu < %zu bytes", ws_bytes,
                      mgv_struct_bwd_legacy_workspace_bytes(N, num_enc));
        return MGV_ERR_WORKSPACE;
    }
    int gx = 0;
    rc = persistent_gx(num_enc, N, &gx);
    if (rc != MGV_OK) return rc;
    const int steps = 2 * rounds;
    const size_t slot = (size_t)N * D;
    const size_t enc_stride = (size_t)(steps + 1) * slot;
    MgvArena a(ws, ws_bytes);
    float* part[2];
    float* agg[2];
    part[0] = a.take<float>((size_t)num_enc * slot); part[1] = a.take<float>((size_t)num_enc * slot);
    agg[0] = a.take<float>((size_t)num_enc * slot); agg[1] = a.take<float>((size_t)num_enc * slot);
    float* partial = a.take<float>((size_t)num_enc * gx * 2 * SGRAD);
    MGV_CUDA(cudaMemsetAsync(partial, 0, (size_t)num_enc * gx * 2 * SGRAD * sizeof(float), st));
    const size_t smem = (size_t)B_SMEM;
    MGV_REQUIRE(precision == 0 || precision == 1, "struct encoder: precision must be 0 (fp32-accurate) or 1 (bf16)");
    MGV_CUDA(cudaFuncSetAttribute((const void*)struct_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MGV_CUDA(cudaFuncSetAttribute((const void*)struct_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int k = steps; k >= 1; --k) {
        StepDev p{};
        fill_step(p, sch, k, steps, layernorm, feat, x, weights, states, slot, enc_stride);
        p.gout = gout;
        p.in_part = part[k & 1]; p.in_agg = agg[k & 1];
        p.out_part = part[(k - 1) & 1]; p.out_agg = agg[(k - 1) & 1];
        p.partial = partial;
        p.trace = (k == 2) ? mgv_debug_trace() : nullptr;
        if (precision == 1) struct_bwd_kernel<true><<<dim3(gx, num_enc), THREADS, smem, st>>>(p);
        else struct_bwd_kernel<false><<<dim3(gx, num_enc), THREADS, smem, st>>>(p);
        mgv_count_launches(1);
    }
    const size_t total = (size_t)num_enc * 2 * SGRAD;
    struct_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, gx, grads, num_enc);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_struct_encoder_bwd");
}
