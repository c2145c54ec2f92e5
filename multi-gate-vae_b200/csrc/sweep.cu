// Level sweep: the level-synchronous TFMlpAggr + GRU propagation of Model.forward
// (dg_ae_model_mig.py:84-129; aig :70-97, xmg :95-147, xag :88-121) and its backward, each as
// ONE persistent cooperative kernel over all rounds and levels (grid barrier between levels).
//
// Math per node i of code T at level >= 1 (arch/tfmlp.py:31-46):
//   x_j = [hs_j || hf_j]  for predecessors j (ascending edge id)
//   score_j = u_T . x_j               (u = msg_k.weight^T attn_lin.weight[64:]; the query part,
//                                       msg_k.bias and attn_lin.bias are constant per softmax group)
//   alpha_j = exp(score_j - max) / (sum exp + 1e-16)          (PyG softmax)
//   xbar = sum alpha_j x_j ; S = sum alpha_j ; m = Wv xbar + bv S
//   GRU(x = m, h = hf_i): r,z,n ; hf_i <- (1-z) n + z h        (torch.nn.GRU cell)
//
// CTAs are statically specialised per gate code (proportional to the code's node count) so a
// CTA streams one code's weights (L1/L2 resident) and, in backward, keeps that code's
// weight-gradient accumulators in shared memory for the whole sweep.
#include "mgv_mma16.cuh"

namespace {

constexpr int D = MGV_D;              // 64
constexpr int D2 = 2 * D;             // 128
constexpr int G3 = 3 * D;             // 192
constexpr int PACK = MGV_SWEEP_PACK_FLOATS;
constexpr int GRAD = MGV_SWEEP_GRAD_FLOATS;
// weight block offsets (floats) -- see include/mgv_b200.h
constexpr int O_U = 0, O_WVT = 128, O_BV = 8320, O_WIHT = 8384, O_WHHT = 20672, O_BIH = 32960, O_BHH = 33152;
constexpr int O_WV = 33344, O_WIH = 41536, O_WHH = 53824;
// gradient block offsets
constexpr int G_U = 0, G_WV = 128, G_BV = 8320, G_WIH = 8384, G_WHH = 20672, G_BIH = 32960, G_BHH = 33152;
constexpr int NODE_MASK = (1 << MGV_CODE_SHIFT) - 1;

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int LDX = D2 + 4;           // 132: padded row of a 128-wide tile
constexpr int LDM = D + 4;            // 68
constexpr int LDG = G3 + 4;           // 196

struct SweepDev {
    int N, L, R;
    unsigned handled;
    const int* order; const int* seg_ptr; const int* in_ptr; const int* in_src;
    const int* out_ptr; const int* out_pack; const int* out_slot;
    const float* weights;
    const float* hs;
    float* hf_all;
    int cta_start[MGV_NCODE + 1];
    unsigned* bar;
    // backward only
    float* ghs; float* ghf; float* dxb; float* alpha; float* dscore; float* partial; float* grads;
};

// Gather + additive attention of one node by one warp.  Lane l owns elements 4l..4l+3 of the
// 128-wide row [hs || hf].  Returns xbar chunk and S; optionally stores raw scores / alphas.
template <bool STORE_ALPHA>
__device__ __forceinline__ void gather_attend(const SweepDev& p, const float* hf_cur, const float* W, int node, int lane,
                                              float4& xbar, float& S) {
    const int beg = p.in_ptr[node], end = p.in_ptr[node + 1];
    const float4 u4 = mgv_ldg4(W + O_U + 4 * lane);
    const int off = (lane < 16) ? 4 * lane : 4 * (lane - 16);
    float mx = -INFINITY, sum = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q0 = beg; q0 < end; q0 += 4) {
        float4 x[4];
        float sc[4];
        const int cnt = min(4, end - q0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < cnt) {
                const int j = p.in_src[q0 + i];
                const float* row = (lane < 16) ? (p.hs + (size_t)j * D) : (hf_cur + (size_t)j * D);
                x[i] = mgv_ld4(row + off);
            } else {
                x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) sc[i] = mgv_warp_sum(mgv_dot4(x[i], u4));
        float nmx = mx;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < cnt) nmx = fmaxf(nmx, sc[i]);
        const float scale = (mx == -INFINITY) ? 0.f : expf(mx - nmx);
        sum *= scale;
        acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < cnt) {
                const float e = expf(sc[i] - nmx);
                sum += e;
                mgv_fma4(acc, e, x[i]);
                if (STORE_ALPHA && lane == 0) p.alpha[q0 + i] = sc[i];     // raw score for now
            }
        }
        mx = nmx;
    }
    const float inv = 1.0f / (sum + 1e-16f);
    xbar = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    S = sum * inv;
    if (STORE_ALPHA) {
        __syncwarp();
        for (int q = beg + lane; q < end; q += 32) p.alpha[q] = expf(p.alpha[q] - mx) * inv;
        __syncwarp();
    }
}

__device__ __forceinline__ void find_role(const SweepDev& p, int& code, int& rank, int& nct) {
    code = -1; rank = 0; nct = 0;
    const int b = blockIdx.x;
#pragma unroll
    for (int c = 0; c < MGV_NCODE; ++c) {
        if (b >= p.cta_start[c] && b < p.cta_start[c + 1]) {
            code = c; rank = b - p.cta_start[c]; nct = p.cta_start[c + 1] - p.cta_start[c];
        }
    }
}

// ======================================================================================= forward
// Tensor-core tiles (mma.sync m16n8k16 on fp16 hi/lo planes, mgv_mma16.cuh).  A CTA keeps ITS gate code's weights as
// planes in shared memory for the whole sweep; a tile is 32 nodes of one (level, code) segment:
//   gather (half-warp per node: [hs || hf] rows of the predecessors, online softmax, xbar)  ->  m = xbar Wv^T + bv S
//   ->  gi = m Wih^T, gh = h Whh^T  ->  GRU gates  ->  hf[node]
constexpr int FTM = 32;
constexpr int FTHREADS = 512;
constexpr int LDXH = 136, LDMH = 72, LDF32 = 68;                // plane row strides (halves) / fp32 tile stride
constexpr uint32_t FW_WV_HI = 0, FW_WV_LO = FW_WV_HI + D * LDXH * 2, FW_WIH_HI = FW_WV_LO + D * LDXH * 2,
                   FW_WIH_LO = FW_WIH_HI + G3 * LDMH * 2, FW_WHH_HI = FW_WIH_LO + G3 * LDMH * 2, FW_WHH_LO = FW_WHH_HI + G3 * LDMH * 2;
constexpr uint32_t FW_BIAS = FW_WHH_LO + G3 * LDMH * 2;        // u[128] bv[64] bih[192] bhh[192] fp32
constexpr uint32_t FW_XS_HI = FW_BIAS + 576 * 4, FW_XS_LO = FW_XS_HI + FTM * LDXH * 2;
constexpr uint32_t FW_MS_HI = FW_XS_LO + FTM * LDXH * 2, FW_MS_LO = FW_MS_HI + FTM * LDMH * 2;
constexpr uint32_t FW_HS_HI = FW_MS_LO + FTM * LDMH * 2, FW_HS_LO = FW_HS_HI + FTM * LDMH * 2;
constexpr uint32_t FW_H32 = FW_HS_LO + FTM * LDMH * 2, FW_SS = FW_H32 + FTM * LDF32 * 4, FW_IDS = FW_SS + FTM * 4;
constexpr uint32_t F_SMEM_BYTES = FW_IDS + FTM * 4;
static_assert(F_SMEM_BYTES <= 227 * 1024 && FW_XS_HI % 16 == 0 && FW_MS_HI % 16 == 0 && FW_HS_HI % 16 == 0, "sweep forward smem");

__device__ __forceinline__ void st_plane8(uint8_t* hi_plane, uint8_t* lo_plane, uint32_t half_off, const float (&v)[8]) {
    uint4 hi, lo;
    m16::split2(v[0], v[1], hi.x, lo.x);
    m16::split2(v[2], v[3], hi.y, lo.y);
    m16::split2(v[4], v[5], hi.z, lo.z);
    m16::split2(v[6], v[7], hi.w, lo.w);
    *reinterpret_cast<uint4*>(hi_plane + half_off * 2) = hi;
    *reinterpret_cast<uint4*>(lo_plane + half_off * 2) = lo;
}
// [rows][cols] fp32 row-major (global) -> hi/lo planes with row stride ld (halves)
__device__ __forceinline__ void load_planes(uint8_t* hi_plane, uint8_t* lo_plane, const float* __restrict__ W, int rows, int cols, int ld,
                                            int tid, int nthr) {
    const int chunks = cols / 8;
    for (int i = tid; i < rows * chunks; i += nthr) {
        const int r = i / chunks, c = i % chunks;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __ldg(W + (size_t)r * cols + c * 8 + e);
        st_plane8(hi_plane, lo_plane, (uint32_t)(r * ld + c * 8), v);
    }
}
__device__ __forceinline__ void gru_gates(float gr, float gz, float gi, float gh, float& r, float& z, float& n) {
    const float a = __expf(-fminf(fmaxf(gr, -28.f), 28.f));
    const float b = __expf(-fminf(fmaxf(gz, -28.f), 28.f));
    const float inv = __fdividef(1.0f, (1.0f + a) * (1.0f + b));
    r = (1.0f + b) * inv;
    z = (1.0f + a) * inv;
    const float y = fminf(fmaxf(fmaf(r, gh, gi), -14.f), 14.f);
    n = 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * y));
}

// Gather + additive attention of one node by one HALF-warp: lane l16 owns columns 8 l16 .. 8 l16 + 7 of the 128-wide
// row [hs || hf].  Returns the xbar chunk and S; optionally stores raw scores, then alphas.
template <bool STORE_ALPHA>
__device__ __forceinline__ void gather_attend16(const SweepDev& p, const float* hf_cur, const float* u8, int node, int l16,
                                                unsigned hmask, float (&xbar)[8], float& S) {
    const int beg = p.in_ptr[node], end = p.in_ptr[node + 1];
    float mx = -INFINITY, sum = 0.f;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int q0 = beg; q0 < end; q0 += 4) {
        float x[4][8], sc[4];
        const int cnt = min(4, end - q0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < cnt) {
                const int j = p.in_src[q0 + i];
                const float* row = (l16 < 8) ? (p.hs + (size_t)j * D + 8 * l16) : (hf_cur + (size_t)j * D + 8 * (l16 - 8));
                const float4 a = mgv_ld4(row), b = mgv_ld4(row + 4);
                x[i][0] = a.x; x[i][1] = a.y; x[i][2] = a.z; x[i][3] = a.w; x[i][4] = b.x; x[i][5] = b.y; x[i][6] = b.z; x[i][7] = b.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[i][e] = 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float d = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) d = fmaf(x[i][e], u8[e], d);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(hmask, d, o);
            sc[i] = d;
        }
        float nmx = mx;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < cnt) nmx = fmaxf(nmx, sc[i]);
        const float scale = (mx == -INFINITY) ? 0.f : expf(mx - nmx);
        sum *= scale;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] *= scale;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < cnt) {
                const float ex = expf(sc[i] - nmx);
                sum += ex;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(ex, x[i][e], acc[e]);
                if (STORE_ALPHA && l16 == 0) p.alpha[q0 + i] = sc[i];     // raw score for now
            }
        }
        mx = nmx;
    }
    const float inv = 1.0f / (sum + 1e-16f);
#pragma unroll
    for (int e = 0; e < 8; ++e) xbar[e] = acc[e] * inv;
    S = sum * inv;
    if (STORE_ALPHA) {
        __syncwarp(hmask);
        for (int q = beg + l16; q < end; q += 16) p.alpha[q] = expf(p.alpha[q] - mx) * inv;
        __syncwarp(hmask);
    }
}

__global__ void __launch_bounds__(FTHREADS, 1) sweep_fwd_kernel(const SweepDev p) {
    extern __shared__ __align__(128) uint8_t fsm[];
    const uint32_t sb = m16::smem_u32(fsm);
    float* Bias = reinterpret_cast<float*>(fsm + FW_BIAS);     // u | bv | bih | bhh
    float* H32 = reinterpret_cast<float*>(fsm + FW_H32);
    float* Ss = reinterpret_cast<float*>(fsm + FW_SS);
    int* Ids = reinterpret_cast<int*>(fsm + FW_IDS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = lane >> 4, l16 = lane & 15;
    const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
    int code, rank, nct;
    find_role(p, code, rank, nct);
    if (code >= 0) {
        const float* W = p.weights + (size_t)code * PACK;
        load_planes(fsm + FW_WV_HI, fsm + FW_WV_LO, W + O_WV, D, D2, LDXH, tid, FTHREADS);
        load_planes(fsm + FW_WIH_HI, fsm + FW_WIH_LO, W + O_WIH, G3, D, LDMH, tid, FTHREADS);
        load_planes(fsm + FW_WHH_HI, fsm + FW_WHH_LO, W + O_WHH, G3, D, LDMH, tid, FTHREADS);
        for (int i = tid; i < 576; i += FTHREADS)
            Bias[i] = __ldg(W + (i < 128 ? O_U + i : (i < 192 ? O_BV + i - 128 : (i < 384 ? O_BIH + i - 192 : O_BHH + i - 384))));
    }
    __syncthreads();
    const float* Bu = Bias; const float* Bv = Bias + 128; const float* Bih = Bias + 192; const float* Bhh = Bias + 384;
    float u8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) u8[e] = Bu[8 * l16 + e];
    const int mt = warp & 1, nt8 = (warp >> 1) * 8;
    const int g = lane >> 2, t = lane & 3;

    for (int r = 0; r < p.R; ++r) {
        const float* hf_prev = r > 0 ? p.hf_all + (size_t)(r - 1) * p.N * D : nullptr;
        float* hf_cur = p.hf_all + (size_t)r * p.N * D;
        for (int lvl = 1; lvl < p.L; ++lvl) {
            if (code >= 0) {
                const int sbeg = p.seg_ptr[lvl * MGV_NCODE + code], send = p.seg_ptr[lvl * MGV_NCODE + code + 1];
                for (int t0 = sbeg + rank * FTM; t0 < send; t0 += nct * FTM) {
                    const int rows = min(FTM, send - t0);
                    {   // ---- phase A: gather + attention, one half-warp per node
                        const int row = warp * 2 + half;
                        float xb[8], h8[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) { xb[e] = 0.f; h8[e] = 0.f; }
                        float S = 0.f;
                        int node = -1;
                        if (row < rows) {
                            node = p.order[t0 + row];
                            gather_attend16<false>(p, hf_cur, u8, node, l16, hmask, xb, S);
                            if (hf_prev != nullptr && l16 < 8) {
                                const float4 a = mgv_ld4(hf_prev + (size_t)node * D + 8 * l16), b = mgv_ld4(hf_prev + (size_t)node * D + 8 * l16 + 4);
                                h8[0] = a.x; h8[1] = a.y; h8[2] = a.z; h8[3] = a.w; h8[4] = b.x; h8[5] = b.y; h8[6] = b.z; h8[7] = b.w;
                            }
                        }
                        st_plane8(fsm + FW_XS_HI, fsm + FW_XS_LO, (uint32_t)(row * LDXH + 8 * l16), xb);
                        if (l16 < 8) {
                            st_plane8(fsm + FW_HS_HI, fsm + FW_HS_LO, (uint32_t)(row * LDMH + 8 * l16), h8);
                            mgv_st4(H32 + row * LDF32 + 8 * l16, make_float4(h8[0], h8[1], h8[2], h8[3]));
                            mgv_st4(H32 + row * LDF32 + 8 * l16 + 4, make_float4(h8[4], h8[5], h8[6], h8[7]));
                        }
                        if (l16 == 0) { Ss[row] = S; Ids[row] = node; }
                    }
                    __syncthreads();
                    {   // ---- phase B: m = xbar Wv^T + bv S   (warp: 16 rows x 8 columns)
                        const int n0[1] = {nt8};
                        float c[1][1][4];
                        m16::zero_frag(c);
                        m16::warp_gemm<1, 1, D2 / 16, false, false>(c, sb + FW_XS_HI, sb + FW_XS_LO, LDXH, mt * 16, 0, sb + FW_WV_HI, sb + FW_WV_LO,
                                                                    LDXH, n0, 0, lane);
#pragma unroll
                        for (int hrow = 0; hrow < 2; ++hrow) {
                            const int row = mt * 16 + g + 8 * hrow, col = nt8 + 2 * t;
                            const float s = Ss[row];
                            uint32_t hi, lo;
                            m16::split2(fmaf(Bv[col], s, c[0][0][2 * hrow]), fmaf(Bv[col + 1], s, c[0][0][2 * hrow + 1]), hi, lo);
                            *reinterpret_cast<uint32_t*>(fsm + FW_MS_HI + (row * LDMH + col) * 2) = hi;
                            *reinterpret_cast<uint32_t*>(fsm + FW_MS_LO + (row * LDMH + col) * 2) = lo;
                        }
                    }
                    __syncthreads();
                    {   // ---- phase C: GRU (warp: 16 rows x 8 units, all three gates)
                        const int n0[3] = {nt8, D + nt8, 2 * D + nt8};
                        float ci[1][3][4], ch[1][3][4];
                        m16::zero_frag(ci);
                        m16::zero_frag(ch);
                        m16::warp_gemm<1, 3, D / 16, false, false>(ci, sb + FW_MS_HI, sb + FW_MS_LO, LDMH, mt * 16, 0, sb + FW_WIH_HI, sb + FW_WIH_LO,
                                                                   LDMH, n0, 0, lane);
                        if (hf_prev != nullptr)
                            m16::warp_gemm<1, 3, D / 16, false, false>(ch, sb + FW_HS_HI, sb + FW_HS_LO, LDMH, mt * 16, 0, sb + FW_WHH_HI, sb + FW_WHH_LO,
                                                                       LDMH, n0, 0, lane);
#pragma unroll
                        for (int hrow = 0; hrow < 2; ++hrow) {
                            const int row = mt * 16 + g + 8 * hrow;
                            const int node = Ids[row];
                            float out[2];
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const int e = 2 * hrow + k, uu = nt8 + 2 * t + k;
                                float rr, zz, nn;
                                gru_gates(ci[0][0][e] + Bih[uu] + ch[0][0][e] + Bhh[uu], ci[0][1][e] + Bih[D + uu] + ch[0][1][e] + Bhh[D + uu],
                                          ci[0][2][e] + Bih[2 * D + uu], ch[0][2][e] + Bhh[2 * D + uu], rr, zz, nn);
                                const float hp = H32[row * LDF32 + uu];
                                out[k] = fmaf(zz, hp - nn, nn);
                            }
                            if (node >= 0) *reinterpret_cast<float2*>(hf_cur + (size_t)node * D + nt8 + 2 * t) = make_float2(out[0], out[1]);
                        }
                    }
                    __syncthreads();
                }
            }
            if (!(r == p.R - 1 && lvl == p.L - 1)) mgv_grid_sync(p.bar, gridDim.x);
        }
    }
}

// ======================================================================================= backward
constexpr int BTM = 16;                                        // nodes per tile
constexpr int B_TILE_FLOATS = 2 * BTM * LDX + 4 * BTM * LDM + 2 * BTM * LDG + 3 * BTM;
constexpr int B_SMEM_FLOATS = GRAD + B_TILE_FLOATS;

// Pull  sum over out-edges e=(v->k) of  alpha_e * dxbar_k + dscore_e * u_code(k)  (128-wide, lane chunk).
__device__ __forceinline__ float4 pull_out_edges(const SweepDev& p, int v, int lane) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int beg = p.out_ptr[v], end = p.out_ptr[v + 1];
    for (int q0 = beg; q0 < end; q0 += 4) {
        float4 dx[4], uu[4];
        float a[4], ds[4];
        const int cnt = min(4, end - q0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = 0.f; ds[i] = 0.f;
            dx[i] = make_float4(0.f, 0.f, 0.f, 0.f); uu[i] = dx[i];
            if (i < cnt) {
                const int pk = p.out_pack[q0 + i];
                const int c = (pk >> MGV_CODE_SHIFT) & 7;
                if ((p.handled >> c) & 1u) {
                    const int k = pk & NODE_MASK;
                    const int slot = p.out_slot[q0 + i];
                    a[i] = p.alpha[slot];
                    ds[i] = p.dscore[slot];
                    dx[i] = mgv_ld4(p.dxb + (size_t)k * D2 + 4 * lane);
                    uu[i] = mgv_ldg4(p.weights + (size_t)c * PACK + O_U + 4 * lane);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            mgv_fma4(acc, a[i], dx[i]);
            mgv_fma4(acc, ds[i], uu[i]);
        }
    }
    return acc;
}

__global__ void __launch_bounds__(THREADS, 1) sweep_bwd_kernel(const SweepDev p) {
    extern __shared__ __align__(16) float smem[];
    float* ACC = smem;                       // [GRAD] this CTA's weight-gradient accumulators
    float* Xs = ACC + GRAD;                  // [16][132] xbar
    float* DXs = Xs + BTM * LDX;             // [16][132] d xbar
    float* Ms = DXs + BTM * LDX;             // [16][68]  m
    float* Hs = Ms + BTM * LDM;              // [16][68]  h
    float* Gs = Hs + BTM * LDM;              // [16][68]  d hf (this round)
    float* DMs = Gs + BTM * LDM;             // [16][68]  d m
    float* DGI = DMs + BTM * LDM;            // [16][196] d gi (r,z,n)
    float* DGH = DGI + BTM * LDG;            // [16][196] d gh
    float* Ss = DGH + BTM * LDG;             // [16]
    float* dSs = Ss + BTM;                   // [16]
    int* Ids = reinterpret_cast<int*>(dSs + BTM);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int code, rank, nct;
    find_role(p, code, rank, nct);
    const float* W = p.weights + (size_t)(code < 0 ? 0 : code) * PACK;
    const int col = tid & 63, rg = tid >> 6;      // 64 columns x 4 row groups of 4 rows

    for (int i = tid; i < GRAD; i += THREADS) ACC[i] = 0.f;
    __syncthreads();

    for (int r = p.R - 1; r >= 0; --r) {
        const float* hf_prev = r > 0 ? p.hf_all + (size_t)(r - 1) * p.N * D : nullptr;
        const float* hf_cur = p.hf_all + (size_t)r * p.N * D;
        for (int lvl = p.L - 1; lvl >= 0; --lvl) {
            // ------------------------------------------------ full backward tiles of my code
            if (code >= 0 && lvl >= 1) {
                const int sb = p.seg_ptr[lvl * MGV_NCODE + code], se = p.seg_ptr[lvl * MGV_NCODE + code + 1];
                for (int t0 = sb + rank * BTM; t0 < se; t0 += nct * BTM) {
                    const int rows = min(BTM, se - t0);
                    // ---- phase P/A: pull d(hs,hf), recompute gather/attention
                    for (int row = warp; row < BTM; row += WARPS) {
                        float4 xb = make_float4(0.f, 0.f, 0.f, 0.f), h4 = xb, g4 = xb;
                        float S = 0.f;
                        int node = -1;
                        if (row < rows) {
                            node = p.order[t0 + row];
                            const float4 pl = pull_out_edges(p, node, lane);
                            if (lane < 16) {
                                float* gp = p.ghs + (size_t)node * D + 4 * lane;
                                float4 cur = mgv_ld4(gp);
                                cur.x += pl.x; cur.y += pl.y; cur.z += pl.z; cur.w += pl.w;
                                mgv_st4(gp, cur);
                            } else {
                                const float4 base = mgv_ld4(p.ghf + (size_t)node * D + 4 * (lane - 16));
                                g4 = make_float4(base.x + pl.x, base.y + pl.y, base.z + pl.z, base.w + pl.w);
                            }
                            gather_attend<true>(p, hf_cur, W, node, lane, xb, S);
                            if (hf_prev != nullptr && lane < 16) h4 = mgv_ld4(hf_prev + (size_t)node * D + 4 * lane);
                        }
                        mgv_st4(Xs + row * LDX + 4 * lane, xb);
                        if (lane < 16) mgv_st4(Hs + row * LDM + 4 * lane, h4);
                        else mgv_st4(Gs + row * LDM + 4 * (lane - 16), g4);
                        if (lane == 0) { Ss[row] = S; Ids[row] = node; }
                    }
                    __syncthreads();
                    // ---- phase B: m
                    {
                        float acc[4] = {0.f, 0.f, 0.f, 0.f};
                        mgv_gemm_col<4, D2>(Xs + rg * 4 * LDX, LDX, W + O_WVT, D, col, acc);
                        const float bv = __ldg(W + O_BV + col);
#pragma unroll
                        for (int i = 0; i < 4; ++i) Ms[(rg * 4 + i) * LDM + col] = fmaf(bv, Ss[rg * 4 + i], acc[i]);
                    }
                    __syncthreads();
                    // ---- phase C: GRU recompute + elementwise backward
                    float dh_direct[4];
                    {
                        float ar[4], az[4], an[4], hr[4], hz[4], hn[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) { ar[i] = az[i] = an[i] = 0.f; hr[i] = hz[i] = hn[i] = 0.f; }
                        mgv_gemm_col3<4, D>(Ms + rg * 4 * LDM, LDM, W + O_WIHT, col, ar, az, an);
                        if (hf_prev != nullptr) mgv_gemm_col3<4, D>(Hs + rg * 4 * LDM, LDM, W + O_WHHT, col, hr, hz, hn);
                        const float bir = __ldg(W + O_BIH + col), biz = __ldg(W + O_BIH + D + col), bin = __ldg(W + O_BIH + 2 * D + col);
                        const float bhr = __ldg(W + O_BHH + col), bhz = __ldg(W + O_BHH + D + col), bhn = __ldg(W + O_BHH + 2 * D + col);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int row = rg * 4 + i;
                            const float g = Gs[row * LDM + col];
                            const float hp = Hs[row * LDM + col];
                            const float rr = mgv_sigmoid(ar[i] + bir + hr[i] + bhr);
                            const float zz = mgv_sigmoid(az[i] + biz + hz[i] + bhz);
                            const float hnb = hn[i] + bhn;
                            const float nn = tanhf(an[i] + bin + rr * hnb);
                            const float dn = g * (1.0f - zz);
                            const float dz = g * (hp - nn);
                            const float dnpre = dn * (1.0f - nn * nn);
                            const float drpre = dnpre * hnb * rr * (1.0f - rr);
                            const float dzpre = dz * zz * (1.0f - zz);
                            dh_direct[i] = g * zz;
                            DGI[row * LDG + col] = drpre;
                            DGI[row * LDG + D + col] = dzpre;
                            DGI[row * LDG + 2 * D + col] = dnpre;
                            DGH[row * LDG + col] = drpre;
                            DGH[row * LDG + D + col] = dzpre;
                            DGH[row * LDG + 2 * D + col] = dnpre * rr;
                        }
                    }
                    __syncthreads();
                    // ---- phase D: d m = d gi . Wih ;  d h = g z + d gh . Whh
                    {
                        float acc[4] = {0.f, 0.f, 0.f, 0.f};
                        mgv_gemm_col<4, G3>(DGI + rg * 4 * LDG, LDG, W + O_WIH, D, col, acc);
#pragma unroll
                        for (int i = 0; i < 4; ++i) DMs[(rg * 4 + i) * LDM + col] = acc[i];
                        if (hf_prev != nullptr) {
                            float acch[4] = {0.f, 0.f, 0.f, 0.f};
                            mgv_gemm_col<4, G3>(DGH + rg * 4 * LDG, LDG, W + O_WHH, D, col, acch);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int node = Ids[rg * 4 + i];
                                if (node >= 0) p.ghf[(size_t)node * D + col] = dh_direct[i] + acch[i];
                            }
                        }
                    }
                    __syncthreads();
                    // ---- phase E: d xbar = d m . Wv   (128 columns x 2 row groups of 8)
                    {
                        const int k = tid & 127, rg2 = tid >> 7;
                        float acc[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
                        mgv_gemm_col<8, D>(DMs + rg2 * 8 * LDM, LDM, W + O_WV, D2, k, acc);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = rg2 * 8 + i;
                            DXs[row * LDX + k] = acc[i];
                            const int node = Ids[row];
                            if (node >= 0) p.dxb[(size_t)node * D2 + k] = acc[i];
                        }
                    }
                    __syncthreads();
                    // ---- phase F: attention backward per node (one warp per node)
                    for (int row = warp; row < rows; row += WARPS) {
                        const int node = Ids[row];
                        const float dS = mgv_warp_sum(__ldg(W + O_BV + lane) * DMs[row * LDM + lane] +
                                                      __ldg(W + O_BV + 32 + lane) * DMs[row * LDM + 32 + lane]);
                        const float4 dxb4 = mgv_ld4(DXs + row * LDX + 4 * lane);
                        const float4 xb4 = mgv_ld4(Xs + row * LDX + 4 * lane);
                        const int beg = p.in_ptr[node], end = p.in_ptr[node + 1];
                        const int off = (lane < 16) ? 4 * lane : 4 * (lane - 16);
                        float A = 0.f;
                        float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int q = beg; q < end; ++q) {
                            const int j = p.in_src[q];
                            const float* rowp = (lane < 16) ? (p.hs + (size_t)j * D) : (hf_cur + (size_t)j * D);
                            const float4 xj = mgv_ld4(rowp + off);
                            const float dal = mgv_warp_sum(mgv_dot4(dxb4, xj)) + dS;
                            const float a = p.alpha[q];
                            A = fmaf(a, dal, A);
                            mgv_fma4(v4, a * dal, xj);
                            if (lane == 0) p.dscore[q] = dal;
                        }
                        __syncwarp();
                        for (int q = beg + lane; q < end; q += 32) p.dscore[q] = p.alpha[q] * (p.dscore[q] - A);
                        // d u += sum_j dscore_j x_j = v - A xbar
                        atomicAdd(ACC + G_U + 4 * lane + 0, v4.x - A * xb4.x);
                        atomicAdd(ACC + G_U + 4 * lane + 1, v4.y - A * xb4.y);
                        atomicAdd(ACC + G_U + 4 * lane + 2, v4.z - A * xb4.z);
                        atomicAdd(ACC + G_U + 4 * lane + 3, v4.w - A * xb4.w);
                    }
                    // ---- phase G: weight-gradient accumulation (reads tile buffers only)
                    {
                        // dWih[o][c] += sum_row dgi[row][o] m[row][c]   (o in [48 og, 48 og + 48), c = col)
                        const int og = rg * 48;
                        float acc[48];
#pragma unroll
                        for (int i = 0; i < 48; ++i) acc[i] = 0.f;
                        for (int row = 0; row < BTM; ++row) {
                            const float mv = Ms[row * LDM + col];
#pragma unroll
                            for (int i = 0; i < 48; i += 4) {
                                const float4 d4 = mgv_ld4(DGI + row * LDG + og + i);
                                acc[i] = fmaf(d4.x, mv, acc[i]); acc[i + 1] = fmaf(d4.y, mv, acc[i + 1]);
                                acc[i + 2] = fmaf(d4.z, mv, acc[i + 2]); acc[i + 3] = fmaf(d4.w, mv, acc[i + 3]);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 48; ++i) ACC[G_WIH + (og + i) * D + col] += acc[i];
                        if (hf_prev != nullptr) {
#pragma unroll
                            for (int i = 0; i < 48; ++i) acc[i] = 0.f;
                            for (int row = 0; row < BTM; ++row) {
                                const float hv = Hs[row * LDM + col];
#pragma unroll
                                for (int i = 0; i < 48; i += 4) {
                                    const float4 d4 = mgv_ld4(DGH + row * LDG + og + i);
                                    acc[i] = fmaf(d4.x, hv, acc[i]); acc[i + 1] = fmaf(d4.y, hv, acc[i + 1]);
                                    acc[i + 2] = fmaf(d4.z, hv, acc[i + 2]); acc[i + 3] = fmaf(d4.w, hv, acc[i + 3]);
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 48; ++i) ACC[G_WHH + (og + i) * D + col] += acc[i];
                        }
                    }
                    {
                        // dWv[c][k] += sum_row dm[row][c] xbar[row][k]   (c in [32 cg, 32 cg + 32), k = tid & 127)
                        const int k = tid & 127, cg = (tid >> 7) * 32;
                        float acc[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                        for (int row = 0; row < BTM; ++row) {
                            const float xv = Xs[row * LDX + k];
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 d4 = mgv_ld4(DMs + row * LDM + cg + i);
                                acc[i] = fmaf(d4.x, xv, acc[i]); acc[i + 1] = fmaf(d4.y, xv, acc[i + 1]);
                                acc[i + 2] = fmaf(d4.z, xv, acc[i + 2]); acc[i + 3] = fmaf(d4.w, xv, acc[i + 3]);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) ACC[G_WV + (cg + i) * D2 + k] += acc[i];
                    }
                    if (tid < G3) {
                        float sgi = 0.f, sgh = 0.f;
                        for (int row = 0; row < BTM; ++row) { sgi += DGI[row * LDG + tid]; sgh += DGH[row * LDG + tid]; }
                        ACC[G_BIH + tid] += sgi;
                        ACC[G_BHH + tid] += sgh;
                    } else {
                        const int c = tid - G3;          // 64 threads: d bv
                        float s = 0.f;
                        for (int row = 0; row < BTM; ++row) s = fmaf(DMs[row * LDM + c], Ss[row], s);
                        ACC[G_BV + c] += s;
                    }
                    __syncthreads();
                }
            }
            // ------------------------------------------------ pull-only nodes (level 0 / codes without a module)
            {
                const int gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
                for (int c = 0; c < MGV_NCODE; ++c) {
                    if (lvl >= 1 && ((p.handled >> c) & 1u)) continue;
                    const int sb = p.seg_ptr[lvl * MGV_NCODE + c], se = p.seg_ptr[lvl * MGV_NCODE + c + 1];
                    for (int t = sb + gw; t < se; t += nw) {
                        const int node = p.order[t];
                        const float4 pl = pull_out_edges(p, node, lane);
                        if (lane < 16) {
                            float* gp = p.ghs + (size_t)node * D + 4 * lane;
                            float4 cur = mgv_ld4(gp);
                            cur.x += pl.x; cur.y += pl.y; cur.z += pl.z; cur.w += pl.w;
                            mgv_st4(gp, cur);
                        }
                    }
                }
            }
            mgv_grid_sync(p.bar, gridDim.x);
        }
    }
    // ---------------------------------------------------- flush accumulators, reduce over the CTAs of a code
    for (int i = tid; i < GRAD; i += THREADS) p.partial[(size_t)blockIdx.x * GRAD + i] = ACC[i];
    mgv_grid_sync(p.bar, gridDim.x);
    const size_t total = (size_t)MGV_NCODE * GRAD;
    for (size_t idx = (size_t)blockIdx.x * THREADS + tid; idx < total; idx += (size_t)gridDim.x * THREADS) {
        const int c = (int)(idx / GRAD), e = (int)(idx % GRAD);
        float s = 0.f;
        for (int b = p.cta_start[c]; b < p.cta_start[c + 1]; ++b) s += p.partial[(size_t)b * GRAD + e];
        p.grads[idx] = s;
    }
}

// Static CTA -> code assignment, proportional to the per-code node counts (at least one CTA per
// code that has nodes).
void assign_ctas(const int64_t* count, unsigned handled, int grid, int* start) {
    int n[MGV_NCODE];
    double total = 0;
    int active = 0;
    for (int c = 0; c < MGV_NCODE; ++c) {
        n[c] = 0;
        if (((handled >> c) & 1u) && count[c] > 0) { total += (double)count[c]; ++active; }
    }
    if (active > 0) {
        int used = 0;
        for (int c = 0; c < MGV_NCODE; ++c) {
            if (!(((handled >> c) & 1u) && count[c] > 0)) continue;
            int k = (int)((double)grid * (double)count[c] / total);
            if (k < 1) k = 1;
            n[c] = k;
            used += k;
        }
        while (used > grid) {                 // shave the largest
            int big = -1;
            for (int c = 0; c < MGV_NCODE; ++c) if (n[c] > 1 && (big < 0 || n[c] > n[big])) big = c;
            if (big < 0) break;
            --n[big]; --used;
        }
        while (used < grid) {                 // hand leftovers to the most loaded code
            int best = -1; double load = -1;
            for (int c = 0; c < MGV_NCODE; ++c) {
                if (n[c] == 0) continue;
                const double l = (double)count[c] / n[c];
                if (l > load) { load = l; best = c; }
            }
            ++n[best]; ++used;
        }
    }
    start[0] = 0;
    for (int c = 0; c < MGV_NCODE; ++c) start[c + 1] = start[c] + n[c];
}

int fill_common(SweepDev& d, const mgv_schedule* sch, int rounds, unsigned handled, const float* weights,
                const float* hs, int32_t* sync) {
    MGV_REQUIRE(sch != nullptr, "level sweep: null schedule");
    MGV_REQUIRE(rounds >= 1, "level sweep: rounds must be >= 1");
    MGV_REQUIRE(sch->N >= 0 && sch->L >= 1, "level sweep: bad schedule sizes");
    MGV_REQUIRE((handled & ~0x7Eu) == 0, "level sweep: handled_mask may only name codes 1..6");
    d.N = sch->N; d.L = sch->L; d.R = rounds; d.handled = handled;
    d.order = sch->order; d.seg_ptr = sch->seg_ptr; d.in_ptr = sch->in_ptr; d.in_src = sch->in_src;
    d.out_ptr = sch->out_ptr; d.out_pack = sch->out_pack; d.out_slot = sch->out_slot;
    d.weights = weights; d.hs = hs; d.bar = reinterpret_cast<unsigned*>(sync);
    d.ghs = d.ghf = d.dxb = d.alpha = d.dscore = d.partial = d.grads = nullptr;
    return MGV_OK;
}

int coop_grid(const void* kernel, size_t smem, int threads, int* grid_out) {
    int dev = 0, sms = 0, occ = 0;
    MGV_CUDA(cudaGetDevice(&dev));
    MGV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MGV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MGV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
    MGV_REQUIRE(occ >= 1, "level sweep: kernel does not fit on an SM (smem %zu)", smem);
    *grid_out = sms * occ;
    return MGV_OK;
}

}  // namespace

extern "C" int mgv_level_sweep_fwd(const mgv_schedule* sch, int32_t rounds, uint32_t handled_mask,
                                   const float* weights, const float* hs, float* hf_all,
                                   int32_t* sync, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SweepDev d;
    int rc = fill_common(d, sch, rounds, handled_mask, weights, hs, sync);
    if (rc != MGV_OK) return rc;
    d.hf_all = hf_all;
    if (sch->N == 0 || sch->L <= 1) return MGV_OK;          // no level >= 1: hf stays zero
    int grid = 0;
    const size_t smem = (size_t)F_SMEM_BYTES;
    rc = coop_grid((const void*)sweep_fwd_kernel, smem, FTHREADS, &grid);
    if (rc != MGV_OK) return rc;
    assign_ctas(sch->code_count, handled_mask, grid, d.cta_start);
    if (d.cta_start[MGV_NCODE] == 0) return MGV_OK;          // nothing to propagate
    MGV_CUDA(cudaMemsetAsync(sync, 0, 64 * sizeof(int32_t), st));
    void* args[] = {&d};
    MGV_CUDA(cudaLaunchCooperativeKernel((void*)sweep_fwd_kernel, dim3(grid), dim3(FTHREADS), args, smem, st));
    mgv_count_launches(1);
    return MGV_OK;
}

extern "C" int mgv_sweep_bwd_grid(void) {
    int grid = 0;
    if (coop_grid((const void*)sweep_bwd_kernel, (size_t)B_SMEM_FLOATS * sizeof(float), THREADS, &grid) != MGV_OK) return -1;
    return grid;
}

extern "C" size_t mgv_sweep_bwd_workspace_bytes(int64_t N, int64_t E) {
    int grid = mgv_sweep_bwd_grid();
    if (grid < 1) grid = 1;
    size_t b = 0;
    b += mgv_align_up((size_t)N * D2 * 4 + 256, 256);          // dxb
    b += 2 * mgv_align_up((size_t)E * 4 + 256, 256);           // alpha, dscore
    b += mgv_align_up((size_t)grid * GRAD * 4 + 256, 256);     // partial
    return b + 1024;
}

extern "C" int mgv_level_sweep_bwd(const mgv_schedule* sch, int32_t rounds, uint32_t handled_mask,
                                   const float* weights, const float* hs, const float* hf_all,
                                   float* ghs, float* ghf, float* grads,
                                   void* ws, size_t ws_bytes, int32_t* sync, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SweepDev d;
    int rc = fill_common(d, sch, rounds, handled_mask, weights, hs, sync);
    if (rc != MGV_OK) return rc;
    d.hf_all = const_cast<float*>(hf_all);
    MGV_CUDA(cudaMemsetAsync(grads, 0, (size_t)MGV_NCODE * GRAD * sizeof(float), st));
    if (sch->N == 0 || sch->L <= 1) return MGV_OK;
    int grid = 0;
    const size_t smem = (size_t)B_SMEM_FLOATS * sizeof(float);
    rc = coop_grid((const void*)sweep_bwd_kernel, smem, THREADS, &grid);
    if (rc != MGV_OK) return rc;
    assign_ctas(sch->code_count, handled_mask, grid, d.cta_start);
    if (d.cta_start[MGV_NCODE] == 0) return MGV_OK;
    if (ws_bytes < mgv_sweep_bwd_workspace_bytes(sch->N, sch->E)) {
        mgv_set_error("mgv_level_sweep_bwd: workspace %zu < %zu bytes", ws_bytes,
                      mgv_sweep_bwd_workspace_bytes(sch->N, sch->E));
        return MGV_ERR_WORKSPACE;
    }
    MgvArena a(ws, ws_bytes);
    d.dxb = a.take<float>((size_t)sch->N * D2);
    d.alpha = a.take<float>((size_t)sch->E + 1);
    d.dscore = a.take<float>((size_t)sch->E + 1);
    d.partial = a.take<float>((size_t)grid * GRAD);
    d.ghs = ghs; d.ghf = ghf; d.grads = grads;
    MGV_CUDA(cudaMemsetAsync(sync, 0, 64 * sizeof(int32_t), st));
    void* args[] = {&d};
    MGV_CUDA(cudaLaunchCooperativeKernel((void*)sweep_bwd_kernel, dim3(grid), dim3(THREADS), args, smem, st));
    mgv_count_launches(1);
    return MGV_OK;
}
