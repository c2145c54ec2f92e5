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
#include <stdlib.h>
#include <string.h>
#include "mgv_mma16.cuh"

long long* mgv_debug_trace();
// tensor-core kernels of the single-round sweep (sweep_tc.cu)
int mgv_sweep_tc_grid(int* grid_out);
int mgv_sweep_tc_fwd(const mgv_schedule* sch, unsigned handled, const int* cta_start, int grid, const float* weights,
                     const float* hs, float* hf, int32_t* sync, int precision, cudaStream_t st);
size_t mgv_sweep_tc_bwd_workspace_bytes(int64_t N, int64_t E);
bool mgv_sweep_tc_bwd_available();
int mgv_sweep_tc_bwd(const mgv_schedule* sch, unsigned handled, const int* cta_start, int grid, const float* weights,
                     const float* hs, const float* hf, float* ghs, float* ghf, float* grads, void* ws, size_t ws_bytes,
                     int32_t* sync, int precision, cudaStream_t st);
// num_rounds = 1 (the reference default) runs on the tcgen05 kernels; MGV_SWEEP=mma forces the mma.sync kernels (A/B checks)
static bool sweep_use_tc(int rounds) {
    const char* e = getenv("MGV_SWEEP");
    return rounds == 1 && !(e && !strcmp(e, "mma"));
}

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
    long long* trace;          // optional [CTA][16] accumulated clock64 cycles per phase (dev tool)
};
#define SWTRACE(slot) do { if (p.trace && tid == 0) { const long long now_ = clock64(); tr_acc[slot] += now_ - tr_last; tr_last = now_; } } while (0)

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

template <bool LOWP>
__device__ __forceinline__ void st_plane8(uint8_t* hi_plane, uint8_t* lo_plane, uint32_t half_off, const float (&v)[8]) {
    uint4 hi, lo;
    m16::split2p<LOWP>(v[0], v[1], hi.x, lo.x);
    m16::split2p<LOWP>(v[2], v[3], hi.y, lo.y);
    m16::split2p<LOWP>(v[4], v[5], hi.z, lo.z);
    m16::split2p<LOWP>(v[6], v[7], hi.w, lo.w);
    *reinterpret_cast<uint4*>(hi_plane + half_off * 2) = hi;
    if (!LOWP) *reinterpret_cast<uint4*>(lo_plane + half_off * 2) = lo;
}
// [rows][cols] fp32 row-major (global) -> hi/lo planes with row stride ld (halves)
template <bool LOWP>
__device__ __forceinline__ void load_planes(uint8_t* hi_plane, uint8_t* lo_plane, const float* __restrict__ W, int rows, int cols, int ld,
                                            int tid, int nthr) {
    const int chunks = cols / 8;
    for (int i = tid; i < rows * chunks; i += nthr) {
        const int r = i / chunks, c = i % chunks;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __ldg(W + (size_t)r * cols + c * 8 + e);
        st_plane8<LOWP>(hi_plane, lo_plane, (uint32_t)(r * ld + c * 8), v);
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

template <bool LOWP>
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
        load_planes<LOWP>(fsm + FW_WV_HI, fsm + FW_WV_LO, W + O_WV, D, D2, LDXH, tid, FTHREADS);
        load_planes<LOWP>(fsm + FW_WIH_HI, fsm + FW_WIH_LO, W + O_WIH, G3, D, LDMH, tid, FTHREADS);
        load_planes<LOWP>(fsm + FW_WHH_HI, fsm + FW_WHH_LO, W + O_WHH, G3, D, LDMH, tid, FTHREADS);
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
                        st_plane8<LOWP>(fsm + FW_XS_HI, fsm + FW_XS_LO, (uint32_t)(row * LDXH + 8 * l16), xb);
                        if (l16 < 8) {
                            st_plane8<LOWP>(fsm + FW_HS_HI, fsm + FW_HS_LO, (uint32_t)(row * LDMH + 8 * l16), h8);
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
                        m16::warp_gemm<1, 1, D2 / 16, false, false, LOWP>(c, sb + FW_XS_HI, sb + FW_XS_LO, LDXH, mt * 16, 0, sb + FW_WV_HI, sb + FW_WV_LO,
                                                                    LDXH, n0, 0, lane);
#pragma unroll
                        for (int hrow = 0; hrow < 2; ++hrow) {
                            const int row = mt * 16 + g + 8 * hrow, col = nt8 + 2 * t;
                            const float s = Ss[row];
                            uint32_t hi, lo;
                            m16::split2p<LOWP>(fmaf(Bv[col], s, c[0][0][2 * hrow]), fmaf(Bv[col + 1], s, c[0][0][2 * hrow + 1]), hi, lo);
                            *reinterpret_cast<uint32_t*>(fsm + FW_MS_HI + (row * LDMH + col) * 2) = hi;
                            if (!LOWP) *reinterpret_cast<uint32_t*>(fsm + FW_MS_LO + (row * LDMH + col) * 2) = lo;
                        }
                    }
                    __syncthreads();
                    {   // ---- phase C: GRU (warp: 16 rows x 8 units, all three gates)
                        const int n0[3] = {nt8, D + nt8, 2 * D + nt8};
                        float ci[1][3][4], ch[1][3][4];
                        m16::zero_frag(ci);
                        m16::zero_frag(ch);
                        m16::warp_gemm<1, 3, D / 16, false, false, LOWP>(ci, sb + FW_MS_HI, sb + FW_MS_LO, LDMH, mt * 16, 0, sb + FW_WIH_HI, sb + FW_WIH_LO,
                                                                   LDMH, n0, 0, lane);
                        if (hf_prev != nullptr)
                            m16::warp_gemm<1, 3, D / 16, false, false, LOWP>(ch, sb + FW_HS_HI, sb + FW_HS_LO, LDMH, mt * 16, 0, sb + FW_WHH_HI, sb + FW_WHH_LO,
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
// Reverse sweep on the same tensor-core tiles.  A tile is 16 nodes of one (level, code) segment; weight planes of the
// CTA's gate code stay in shared memory and are read in both orientations (ldmatrix / ldmatrix.trans); weight
// gradients are persistent register fragments, bias / attention-vector gradients shared-memory accumulators.
//   P/A  pull d(hs, hf) over out-edges, recompute gather / attention (alphas kept)        half-warp per node
//   B    m = xbar Wv^T + bv S                                                            C  GRU recompute + pointwise backward
//   D    d m = d gi Wih,  d h = g z + d gh Whh        E  d xbar = d m Wv        F  attention backward (warp per node)
//   G    d Wih += d gi^T m,  d Whh += d gh^T h,  d Wv += d m^T xbar
// Gate gradients are scaled per tile by a power of two into fp16 range before the hi/lo split (mgv_mma16.cuh).
constexpr int BTHREADS = 512;
constexpr int LDGH = 264;                                      // d gate planes: dr | dz | dn_i | dn_h
// Two instantiations: <16 nodes per tile, with the hidden-state products> for multi-round sweeps, and
// <32 nodes per tile, no W_hh / h tiles> for the common single-round sweep (h = 0: gh = b_hh), where the freed
// shared memory doubles the tile and every phase runs on all 16 warps.
template <int BTM, bool HAS_H>
struct BwdSmem {
    static constexpr uint32_t WV_HI = 0, WV_LO = WV_HI + D * LDXH * 2, WIH_HI = WV_LO + D * LDXH * 2, WIH_LO = WIH_HI + G3 * LDMH * 2;
    static constexpr uint32_t WHH_HI = WIH_LO + G3 * LDMH * 2, WHH_LO = WHH_HI + (HAS_H ? G3 * LDMH * 2 : 0);
    static constexpr uint32_t BIAS = WHH_LO + (HAS_H ? G3 * LDMH * 2 : 0);          // u[128] bv[64] bih[192] bhh[192] fp32
    static constexpr uint32_t ACC = BIAS + 576 * 4;                                  // d u[128] d bv[64] d bih[192] d bhh[192] fp32
    static constexpr uint32_t XS_HI = ACC + 576 * 4, XS_LO = XS_HI + BTM * LDXH * 2;
    static constexpr uint32_t X32 = XS_LO + BTM * LDXH * 2, DX32 = X32 + BTM * LDX * 4;
    static constexpr uint32_t MS_HI = DX32 + BTM * LDX * 4, MS_LO = MS_HI + BTM * LDMH * 2;
    static constexpr uint32_t HS_HI = MS_LO + BTM * LDMH * 2, HS_LO = HS_HI + (HAS_H ? BTM * LDMH * 2 : 0);
    static constexpr uint32_t H32 = HS_LO + (HAS_H ? BTM * LDMH * 2 : 0), GS = H32 + (HAS_H ? BTM * LDF32 * 4 : 0);
    static constexpr uint32_t DM_HI = GS + BTM * LDF32 * 4, DM_LO = DM_HI + BTM * LDMH * 2, DM32 = DM_LO + BTM * LDMH * 2;
    static constexpr uint32_t DG_HI = DM32 + BTM * LDF32 * 4, DG_LO = DG_HI + BTM * LDGH * 2;
    static constexpr uint32_t SS = DG_LO + BTM * LDGH * 2, IDS = SS + BTM * 4, MAX = IDS + BTM * 4;
    static constexpr uint32_t BYTES = MAX + 16;
    static_assert(BYTES <= 227 * 1024 && XS_HI % 16 == 0 && MS_HI % 16 == 0 && HS_HI % 16 == 0 && DM_HI % 16 == 0 && DG_HI % 16 == 0 &&
                  DG_LO % 16 == 0 && X32 % 16 == 0, "sweep backward smem");
};

// Pull  sum over out-edges e=(v->k) of  alpha_e * dxbar_k + dscore_e * u_code(k)  (128-wide, lane chunk of 4).
// Fan-out is heavy-tailed (primary inputs and shared sub-expressions feed tens of gates), and the per-edge chain
// out_pack -> out_slot -> alpha / dscore -> dxbar row is three dependent loads, so the edges are taken a WARP AT A TIME:
// every lane fetches the metadata of one edge (the chains of 32 edges run in parallel), the rows are then gathered
// 8 at a time with the ids broadcast by shuffles, and  sum_e dscore_e u_code(e)  is accumulated per code and applied
// once at the end (no u loads in the loop).
__device__ __forceinline__ float4 pull_out_edges(const SweepDev& p, int v, int lane) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int beg = p.out_ptr[v], end = p.out_ptr[v + 1];
    float sds[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    unsigned present = 0u;
    for (int q0 = beg; q0 < end; q0 += 32) {
        const int q = q0 + lane;
        int kk = -1, c = 0;
        float a = 0.f, ds = 0.f;
        if (q < end) {
            const int pk = p.out_pack[q];
            c = (pk >> MGV_CODE_SHIFT) & 7;
            if ((p.handled >> c) & 1u) {
                kk = pk & NODE_MASK;
                const int slot = p.out_slot[q];
                a = p.alpha[slot];
                ds = p.dscore[slot];
            }
        }
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) sds[cc] += (kk >= 0 && c == cc + 1) ? ds : 0.f;
        present |= (kk >= 0) ? (1u << c) : 0u;
        const int ne = min(32, end - q0);
        for (int i = 0; i < ne; i += 8) {
            float4 dx[8];
            float aa[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int kj = __shfl_sync(0xffffffffu, kk, (i + j) & 31);
                aa[j] = __shfl_sync(0xffffffffu, a, (i + j) & 31);
                dx[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i + j < ne && kj >= 0) dx[j] = mgv_ld4(p.dxb + (size_t)kj * D2 + 4 * lane);
                else aa[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) mgv_fma4(acc, aa[j], dx[j]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) present |= __shfl_xor_sync(0xffffffffu, present, o);
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) {
        if ((present >> (cc + 1)) & 1u) {
            float t = sds[cc];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            mgv_fma4(acc, t, mgv_ldg4(p.weights + (size_t)(cc + 1) * PACK + O_U + 4 * lane));
        }
    }
    return acc;
}
// The same for a half-warp: lane l16 owns columns 8 l16 .. 8 l16 + 7 (16 edges per metadata wave, rows 4 at a time).
__device__ __forceinline__ void pull_out_edges16(const SweepDev& p, int v, int l16, unsigned hmask, int lane, float (&acc)[8]) {
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    const int beg = p.out_ptr[v], end = p.out_ptr[v + 1];
    const int hb = lane & 16;
    float sds[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    unsigned present = 0u;
    for (int q0 = beg; q0 < end; q0 += 16) {
        const int q = q0 + l16;
        int kk = -1, c = 0;
        float a = 0.f, ds = 0.f;
        if (q < end) {
            const int pk = p.out_pack[q];
            c = (pk >> MGV_CODE_SHIFT) & 7;
            if ((p.handled >> c) & 1u) {
                kk = pk & NODE_MASK;
                const int slot = p.out_slot[q];
                a = p.alpha[slot];
                ds = p.dscore[slot];
            }
        }
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) sds[cc] += (kk >= 0 && c == cc + 1) ? ds : 0.f;
        present |= (kk >= 0) ? (1u << c) : 0u;
        const int ne = min(16, end - q0);
        for (int i = 0; i < ne; i += 4) {
            float4 dx[4][2];
            float aa[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kj = __shfl_sync(hmask, kk, hb + ((i + j) & 15));
                aa[j] = __shfl_sync(hmask, a, hb + ((i + j) & 15));
                dx[j][0] = make_float4(0.f, 0.f, 0.f, 0.f); dx[j][1] = dx[j][0];
                if (i + j < ne && kj >= 0) {
                    dx[j][0] = mgv_ld4(p.dxb + (size_t)kj * D2 + 8 * l16);
                    dx[j][1] = mgv_ld4(p.dxb + (size_t)kj * D2 + 8 * l16 + 4);
                } else aa[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[0] = fmaf(aa[j], dx[j][0].x, acc[0]); acc[1] = fmaf(aa[j], dx[j][0].y, acc[1]);
                acc[2] = fmaf(aa[j], dx[j][0].z, acc[2]); acc[3] = fmaf(aa[j], dx[j][0].w, acc[3]);
                acc[4] = fmaf(aa[j], dx[j][1].x, acc[4]); acc[5] = fmaf(aa[j], dx[j][1].y, acc[5]);
                acc[6] = fmaf(aa[j], dx[j][1].z, acc[6]); acc[7] = fmaf(aa[j], dx[j][1].w, acc[7]);
            }
        }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) present |= __shfl_xor_sync(hmask, present, o);
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) {
        if ((present >> (cc + 1)) & 1u) {
            float t = sds[cc];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(hmask, t, o);
            const float4 u0 = mgv_ldg4(p.weights + (size_t)(cc + 1) * PACK + O_U + 8 * l16), u1 = mgv_ldg4(p.weights + (size_t)(cc + 1) * PACK + O_U + 8 * l16 + 4);
            acc[0] = fmaf(t, u0.x, acc[0]); acc[1] = fmaf(t, u0.y, acc[1]); acc[2] = fmaf(t, u0.z, acc[2]); acc[3] = fmaf(t, u0.w, acc[3]);
            acc[4] = fmaf(t, u1.x, acc[4]); acc[5] = fmaf(t, u1.y, acc[5]); acc[6] = fmaf(t, u1.z, acc[6]); acc[7] = fmaf(t, u1.w, acc[7]);
        }
    }
}

template <int BTM, bool HAS_H, bool LOWP>
__global__ void __launch_bounds__(BTHREADS, 1) sweep_bwd_kernel(const SweepDev p) {
    using S = BwdSmem<BTM, HAS_H>;
    constexpr int MT = BTM / 16;
    static_assert(!HAS_H || MT == 1, "the multi-round variant uses 16-node tiles");
    extern __shared__ __align__(128) uint8_t bsm[];
    const uint32_t sb = m16::smem_u32(bsm);
    float* Bias = reinterpret_cast<float*>(bsm + S::BIAS);
    float* ACC = reinterpret_cast<float*>(bsm + S::ACC);        // d u | d bv | d bih | d bhh
    float* X32 = reinterpret_cast<float*>(bsm + S::X32);
    float* DX32 = reinterpret_cast<float*>(bsm + S::DX32);
    float* H32 = reinterpret_cast<float*>(bsm + S::H32);
    float* Gs = reinterpret_cast<float*>(bsm + S::GS);          // d hf of the tile, then g z (direct path to h)
    float* DM32 = reinterpret_cast<float*>(bsm + S::DM32);
    float* Ss = reinterpret_cast<float*>(bsm + S::SS);
    int* Ids = reinterpret_cast<int*>(bsm + S::IDS);
    unsigned* smax = reinterpret_cast<unsigned*>(bsm + S::MAX);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = lane >> 4, l16 = lane & 15;
    const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
    const int g = lane >> 2, t = lane & 3;
    int code, rank, nct;
    find_role(p, code, rank, nct);
    const float* W = p.weights + (size_t)(code < 0 ? 0 : code) * PACK;
    if (code >= 0) {
        load_planes<LOWP>(bsm + S::WV_HI, bsm + S::WV_LO, W + O_WV, D, D2, LDXH, tid, BTHREADS);
        load_planes<LOWP>(bsm + S::WIH_HI, bsm + S::WIH_LO, W + O_WIH, G3, D, LDMH, tid, BTHREADS);
        if (HAS_H) load_planes<LOWP>(bsm + S::WHH_HI, bsm + S::WHH_LO, W + O_WHH, G3, D, LDMH, tid, BTHREADS);
        for (int i = tid; i < 576; i += BTHREADS)
            Bias[i] = __ldg(W + (i < 128 ? O_U + i : (i < 192 ? O_BV + i - 128 : (i < 384 ? O_BIH + i - 192 : O_BHH + i - 384))));
    }
    for (int i = tid; i < 576; i += BTHREADS) ACC[i] = 0.f;
    if (tid < 2) smax[tid] = 0u;
    const float* Bu = Bias; const float* Bv = Bias + 128; const float* Bih = Bias + 192; const float* Bhh = Bias + 384;
    float* Au = ACC; float* Abv = ACC + 128; float* Abih = ACC + 192; float* Abhh = ACC + 384;
    float u8[8];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) u8[e] = Bu[8 * l16 + e];

    // persistent weight-gradient fragments (all scaled by acc_scale)
    const int wh = warp >> 3, wn8 = 8 * (warp & 7);
    const int mt = (MT == 2) ? wh : 0, mrow = mt * 16;
    const int wn0[1] = {wn8};
    const int vn0[1] = {8 * warp};
    float acc_ih[6][1][4], acc_ha[2][1][4], acc_hb[4][1][4], acc_v[4][1][4];
    m16::zero_frag(acc_ih);
    m16::zero_frag(acc_ha);
    m16::zero_frag(acc_hb);
    m16::zero_frag(acc_v);
    long long tr_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tr_last = clock64();
    float acc_scale = 1.0f;
    float du_acc[4] = {0.f, 0.f, 0.f, 0.f};                    // d u of columns 4 lane .. 4 lane + 3 (nodes of this warp)
    int it = 0;

    for (int r = p.R - 1; r >= 0; --r) {
        const float* hf_prev = (HAS_H && r > 0) ? p.hf_all + (size_t)(r - 1) * p.N * D : nullptr;
        const float* hf_cur = p.hf_all + (size_t)r * p.N * D;
        for (int lvl = p.L - 1; lvl >= 0; --lvl) {
            // ------------------------------------------------ full backward tiles of my code
            if (code >= 0 && lvl >= 1) {
                const int sbeg = p.seg_ptr[lvl * MGV_NCODE + code], send = p.seg_ptr[lvl * MGV_NCODE + code + 1];
                for (int t0 = sbeg + rank * BTM; t0 < send; t0 += nct * BTM, ++it) {
                    const int rows = min(BTM, send - t0);
                    SWTRACE(9);
                    // ---- phase P/A: pull d(hs, hf), recompute gather / attention  (warps 0-7, half-warp per node)
                    if (warp < BTM / 2) {
                        const int row = warp * 2 + half;
                        float xb[8], h8[8], g8[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) { xb[e] = 0.f; h8[e] = 0.f; g8[e] = 0.f; }
                        float S = 0.f;
                        int node = -1;
                        if (row < rows) {
                            node = p.order[t0 + row];
                            float pl[8];
                            pull_out_edges16(p, node, l16, hmask, lane, pl);
                            if (l16 < 8) {
                                float* gp = p.ghs + (size_t)node * D + 8 * l16;
                                float4 c0 = mgv_ld4(gp), c1 = mgv_ld4(gp + 4);
                                c0.x += pl[0]; c0.y += pl[1]; c0.z += pl[2]; c0.w += pl[3];
                                c1.x += pl[4]; c1.y += pl[5]; c1.z += pl[6]; c1.w += pl[7];
                                mgv_st4(gp, c0); mgv_st4(gp + 4, c1);
                                if (hf_prev != nullptr) {
                                    const float4 a = mgv_ld4(hf_prev + (size_t)node * D + 8 * l16), b2 = mgv_ld4(hf_prev + (size_t)node * D + 8 * l16 + 4);
                                    h8[0] = a.x; h8[1] = a.y; h8[2] = a.z; h8[3] = a.w; h8[4] = b2.x; h8[5] = b2.y; h8[6] = b2.z; h8[7] = b2.w;
                                }
                            } else {
                                const float* gq = p.ghf + (size_t)node * D + 8 * (l16 - 8);
                                const float4 c0 = mgv_ld4(gq), c1 = mgv_ld4(gq + 4);
                                g8[0] = c0.x + pl[0]; g8[1] = c0.y + pl[1]; g8[2] = c0.z + pl[2]; g8[3] = c0.w + pl[3];
                                g8[4] = c1.x + pl[4]; g8[5] = c1.y + pl[5]; g8[6] = c1.z + pl[6]; g8[7] = c1.w + pl[7];
                            }
                            gather_attend16<true>(p, hf_cur, u8, node, l16, hmask, xb, S);
                        }
                        st_plane8<LOWP>(bsm + S::XS_HI, bsm + S::XS_LO, (uint32_t)(row * LDXH + 8 * l16), xb);
                        mgv_st4(X32 + row * LDX + 8 * l16, make_float4(xb[0], xb[1], xb[2], xb[3]));
                        mgv_st4(X32 + row * LDX + 8 * l16 + 4, make_float4(xb[4], xb[5], xb[6], xb[7]));
                        if (l16 < 8) {
                            if (HAS_H) {
                                st_plane8<LOWP>(bsm + S::HS_HI, bsm + S::HS_LO, (uint32_t)(row * LDMH + 8 * l16), h8);
                                mgv_st4(H32 + row * LDF32 + 8 * l16, make_float4(h8[0], h8[1], h8[2], h8[3]));
                                mgv_st4(H32 + row * LDF32 + 8 * l16 + 4, make_float4(h8[4], h8[5], h8[6], h8[7]));
                            }
                        } else {
                            mgv_st4(Gs + row * LDF32 + 8 * (l16 - 8), make_float4(g8[0], g8[1], g8[2], g8[3]));
                            mgv_st4(Gs + row * LDF32 + 8 * (l16 - 8) + 4, make_float4(g8[4], g8[5], g8[6], g8[7]));
                        }
                        if (l16 == 0) { Ss[row] = S; Ids[row] = node; }
                    }
                    __syncthreads();
                    SWTRACE(0);
                    // ---- phase B: m = xbar Wv^T + bv S  (warps 0-7: 8 columns each)
                    if (warp < 8 * MT) {
                        float c[1][1][4];
                        m16::zero_frag(c);
                        m16::warp_gemm<1, 1, D2 / 16, false, false, LOWP>(c, sb + S::XS_HI, sb + S::XS_LO, LDXH, mrow, 0, sb + S::WV_HI, sb + S::WV_LO, LDXH,
                                                                    wn0, 0, lane);
#pragma unroll
                        for (int hrow = 0; hrow < 2; ++hrow) {
                            const int row = mrow + g + 8 * hrow, col = wn8 + 2 * t;
                            const float sv = Ss[row];
                            uint32_t hi, lo;
                            m16::split2p<LOWP>(fmaf(Bv[col], sv, c[0][0][2 * hrow]), fmaf(Bv[col + 1], sv, c[0][0][2 * hrow + 1]), hi, lo);
                            *reinterpret_cast<uint32_t*>(bsm + S::MS_HI + (row * LDMH + col) * 2) = hi;
                            if (!LOWP) *reinterpret_cast<uint32_t*>(bsm + S::MS_LO + (row * LDMH + col) * 2) = lo;
                        }
                    }
                    __syncthreads();
                    SWTRACE(1);
                    // ---- phase C: GRU recompute + pointwise backward  (warps 0-7: 8 units each)
                    float dr[4], dz[4], dni[4], dnh[4];
                    if (warp < 8 * MT) {
                        const int n0[3] = {wn8, D + wn8, 2 * D + wn8};
                        float ci[1][3][4], ch[1][3][4];
                        m16::zero_frag(ci);
                        m16::zero_frag(ch);
                        m16::warp_gemm<1, 3, D / 16, false, false, LOWP>(ci, sb + S::MS_HI, sb + S::MS_LO, LDMH, mrow, 0, sb + S::WIH_HI, sb + S::WIH_LO, LDMH,
                                                                   n0, 0, lane);
                        if (HAS_H && hf_prev != nullptr)
                            m16::warp_gemm<1, 3, D / 16, false, false, LOWP>(ch, sb + S::HS_HI, sb + S::HS_LO, LDMH, mrow, 0, sb + S::WHH_HI, sb + S::WHH_LO,
                                                                       LDMH, n0, 0, lane);
                        float amax = 0.f;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int row = mrow + g + ((e & 2) ? 8 : 0), uu = wn8 + 2 * t + (e & 1);
                            const float hnb = ch[0][2][e] + Bhh[2 * D + uu];
                            float rr, zz, nn;
                            gru_gates(ci[0][0][e] + Bih[uu] + ch[0][0][e] + Bhh[uu], ci[0][1][e] + Bih[D + uu] + ch[0][1][e] + Bhh[D + uu],
                                      ci[0][2][e] + Bih[2 * D + uu], hnb, rr, zz, nn);
                            const float gg = Gs[row * LDF32 + uu];
                            const float hp = HAS_H ? H32[row * LDF32 + uu] : 0.f;
                            dni[e] = gg * (1.0f - zz) * (1.0f - nn * nn);
                            dr[e] = dni[e] * hnb * rr * (1.0f - rr);
                            dz[e] = gg * (hp - nn) * zz * (1.0f - zz);
                            dnh[e] = dni[e] * rr;
                            Gs[row * LDF32 + uu] = gg * zz;                                  // direct path to h
                            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(dr[e]), fabsf(dz[e])), fabsf(dni[e])));
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
                        if (lane == 0) atomicMax(smax + (it & 1), __float_as_uint(amax));
                        // bias gradients: column sums over the 16 rows (lanes with equal t), then one atomic per column
                        float sums[8] = {dr[0] + dr[2], dr[1] + dr[3], dz[0] + dz[2], dz[1] + dz[3],
                                         dni[0] + dni[2], dni[1] + dni[3], dnh[0] + dnh[2], dnh[1] + dnh[3]};
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            sums[k] += __shfl_xor_sync(0xffffffffu, sums[k], 4);
                            sums[k] += __shfl_xor_sync(0xffffffffu, sums[k], 8);
                            sums[k] += __shfl_xor_sync(0xffffffffu, sums[k], 16);
                        }
                        if (g == 0) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const int uu = wn8 + 2 * t + k;
                                atomicAdd(Abih + uu, sums[k]); atomicAdd(Abih + D + uu, sums[2 + k]); atomicAdd(Abih + 2 * D + uu, sums[4 + k]);
                                atomicAdd(Abhh + uu, sums[k]); atomicAdd(Abhh + D + uu, sums[2 + k]); atomicAdd(Abhh + 2 * D + uu, sums[6 + k]);
                            }
                        }
                    }
                    if (tid == 0) smax[(it + 1) & 1] = 0u;
                    __syncthreads();
                    const float scale = m16::pow2_scale_band(__uint_as_float(smax[it & 1]), acc_scale, 1.0f, 512.0f);
                    const float inv_scale = 1.0f / scale;
                    if (warp < 8 * MT) {
#pragma unroll
                        for (int hrow = 0; hrow < 2; ++hrow) {
                            const int row = mrow + g + 8 * hrow, e0 = 2 * hrow;
                            const uint32_t off = (uint32_t)(row * LDGH + wn8 + 2 * t) * 2;
                            uint32_t hi, lo;
                            m16::split2p<LOWP>(dr[e0] * scale, dr[e0 + 1] * scale, hi, lo);
                            *reinterpret_cast<uint32_t*>(bsm + S::DG_HI + off) = hi; if (!LOWP) *reinterpret_cast<uint32_t*>(bsm + S::DG_LO + off) = lo;
                            m16::split2p<LOWP>(dz[e0] * scale, dz[e0 + 1] * scale, hi, lo);
                            *reinterpret_cast<uint32_t*>(bsm + S::DG_HI + off + 2 * D) = hi; if (!LOWP) *reinterpret_cast<uint32_t*>(bsm + S::DG_LO + off + 2 * D) = lo;
                            m16::split2p<LOWP>(dni[e0] * scale, dni[e0 + 1] * scale, hi, lo);
                            *reinterpret_cast<uint32_t*>(bsm + S::DG_HI + off + 4 * D) = hi; if (!LOWP) *reinterpret_cast<uint32_t*>(bsm + S::DG_LO + off + 4 * D) = lo;
                            m16::split2p<LOWP>(dnh[e0] * scale, dnh[e0 + 1] * scale, hi, lo);
                            *reinterpret_cast<uint32_t*>(bsm + S::DG_HI + off + 6 * D) = hi; if (!LOWP) *reinterpret_cast<uint32_t*>(bsm + S::DG_LO + off + 6 * D) = lo;
                        }
                    }
                    __syncthreads();
                    SWTRACE(2);
                    // ---- phase D: d m = d gi . Wih (warps 0-7);  d h = g z + d gh . Whh (warps 8-15)
                    if (warp < 8 * MT) {
                        float c[1][1][4];
                        m16::zero_frag(c);
                        m16::warp_gemm<1, 1, G3 / 16, false, true, LOWP>(c, sb + S::DG_HI, sb + S::DG_LO, LDGH, mrow, 0, sb + S::WIH_HI, sb + S::WIH_LO, LDMH,
                                                                   wn0, 0, lane);
                        float sbv[2] = {0.f, 0.f};
#pragma unroll
                        for (int hrow = 0; hrow < 2; ++hrow) {
                            const int row = mrow + g + 8 * hrow, col = wn8 + 2 * t;
                            uint32_t hi, lo;
                            m16::split2p<LOWP>(c[0][0][2 * hrow], c[0][0][2 * hrow + 1], hi, lo);              // stays scaled
                            *reinterpret_cast<uint32_t*>(bsm + S::DM_HI + (row * LDMH + col) * 2) = hi;
                            if (!LOWP) *reinterpret_cast<uint32_t*>(bsm + S::DM_LO + (row * LDMH + col) * 2) = lo;
                            const float d0 = c[0][0][2 * hrow] * inv_scale, d1 = c[0][0][2 * hrow + 1] * inv_scale;
                            DM32[row * LDF32 + col] = d0;
                            DM32[row * LDF32 + col + 1] = d1;
                            sbv[0] = fmaf(d0, Ss[row], sbv[0]);
                            sbv[1] = fmaf(d1, Ss[row], sbv[1]);
                        }
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            sbv[k] += __shfl_xor_sync(0xffffffffu, sbv[k], 4);
                            sbv[k] += __shfl_xor_sync(0xffffffffu, sbv[k], 8);
                            sbv[k] += __shfl_xor_sync(0xffffffffu, sbv[k], 16);
                        }
                        if (g == 0) { atomicAdd(Abv + wn8 + 2 * t, sbv[0]); atomicAdd(Abv + wn8 + 2 * t + 1, sbv[1]); }
                    } else if (HAS_H && hf_prev != nullptr) {
                        float c[1][1][4];
                        m16::zero_frag(c);
                        m16::warp_gemm<1, 1, 2 * D / 16, false, true, LOWP>(c, sb + S::DG_HI, sb + S::DG_LO, LDGH, 0, 0, sb + S::WHH_HI, sb + S::WHH_LO, LDMH,
                                                                      wn0, 0, lane);
                        m16::warp_gemm<1, 1, D / 16, false, true, LOWP>(c, sb + S::DG_HI, sb + S::DG_LO, LDGH, 0, 3 * D, sb + S::WHH_HI, sb + S::WHH_LO, LDMH,
                                                                  wn0, 2 * D, lane);
#pragma unroll
                        for (int hrow = 0; hrow < 2; ++hrow) {
                            const int row = g + 8 * hrow, col = wn8 + 2 * t;
                            const int node = Ids[row];
                            if (node >= 0)
                                *reinterpret_cast<float2*>(p.ghf + (size_t)node * D + col) =
                                    make_float2(fmaf(c[0][0][2 * hrow], inv_scale, Gs[row * LDF32 + col]),
                                                fmaf(c[0][0][2 * hrow + 1], inv_scale, Gs[row * LDF32 + col + 1]));
                        }
                    }
                    __syncthreads();
                    SWTRACE(3);
                    // ---- phase E: d xbar = d m . Wv   (16 warps x 8 of the 128 columns)
                    {
                        float c[MT][1][4];
                        m16::zero_frag(c);
                        m16::warp_gemm<MT, 1, D / 16, false, true, LOWP>(c, sb + S::DM_HI, sb + S::DM_LO, LDMH, 0, 0, sb + S::WV_HI, sb + S::WV_LO, LDXH,
                                                                   vn0, 0, lane);
#pragma unroll
                        for (int m = 0; m < MT; ++m)
#pragma unroll
                            for (int hrow = 0; hrow < 2; ++hrow) {
                                const int row = 16 * m + g + 8 * hrow, col = 8 * warp + 2 * t;
                                const float2 v = make_float2(c[m][0][2 * hrow] * inv_scale, c[m][0][2 * hrow + 1] * inv_scale);
                                *reinterpret_cast<float2*>(DX32 + row * LDX + col) = v;
                                const int node = Ids[row];
                                if (node >= 0) *reinterpret_cast<float2*>(p.dxb + (size_t)node * D2 + col) = v;
                            }
                    }
                    __syncthreads();
                    SWTRACE(4);
                    // ---- phase F: attention backward per node (one warp per node)
                    for (int row = warp; row < rows; row += BTHREADS / 32) {
                        const int node = Ids[row];
                        const float dS = mgv_warp_sum(Bv[lane] * DM32[row * LDF32 + lane] + Bv[32 + lane] * DM32[row * LDF32 + 32 + lane]);
                        const float4 dxb4 = mgv_ld4(DX32 + row * LDX + 4 * lane);
                        const float4 xb4 = mgv_ld4(X32 + row * LDX + 4 * lane);
                        const int beg = p.in_ptr[node], end = p.in_ptr[node + 1];
                        const int off = (lane < 16) ? 4 * lane : 4 * (lane - 16);
                        float A = 0.f;
                        float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int q0 = beg; q0 < end; q0 += 4) {           // 4 in-edges per trip: ids, then rows, in flight together
                            const int cnt = min(4, end - q0);
                            float a[4];
                            float4 xj[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                a[i] = 0.f;
                                xj[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (i < cnt) {
                                    const int j = p.in_src[q0 + i];
                                    a[i] = p.alpha[q0 + i];
                                    const float* rowp = (lane < 16) ? (p.hs + (size_t)j * D) : (hf_cur + (size_t)j * D);
                                    xj[i] = mgv_ld4(rowp + off);
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                if (i < cnt) {
                                    const float dal = mgv_warp_sum(mgv_dot4(dxb4, xj[i])) + dS;
                                    A = fmaf(a[i], dal, A);
                                    mgv_fma4(v4, a[i] * dal, xj[i]);
                                    if (lane == 0) p.dscore[q0 + i] = dal;
                                }
                            }
                        }
                        __syncwarp();
                        for (int q = beg + lane; q < end; q += 32) p.dscore[q] = p.alpha[q] * (p.dscore[q] - A);
                        // d u += sum_j dscore_j x_j = v - A xbar
                        du_acc[0] += v4.x - A * xb4.x; du_acc[1] += v4.y - A * xb4.y;
                        du_acc[2] += v4.z - A * xb4.z; du_acc[3] += v4.w - A * xb4.w;
                    }
                    SWTRACE(5);
                    // ---- phase G: weight gradients (tile buffers are read-only here)
                    if (scale != acc_scale) {
                        const float f = scale / acc_scale;
                        m16::scale_frag(acc_ih, f);
                        m16::scale_frag(acc_ha, f);
                        m16::scale_frag(acc_hb, f);
                        m16::scale_frag(acc_v, f);
                        acc_scale = scale;
                    }
                    m16::warp_gemm<6, 1, MT, true, true, LOWP>(acc_ih, sb + S::DG_HI, sb + S::DG_LO, LDGH, wh * 96, 0, sb + S::MS_HI, sb + S::MS_LO, LDMH, wn0, 0, lane);
                    if (HAS_H && hf_prev != nullptr) {
                        m16::warp_gemm<2, 1, MT, true, true, LOWP>(acc_ha, sb + S::DG_HI, sb + S::DG_LO, LDGH, wh * 96, 0, sb + S::HS_HI, sb + S::HS_LO, LDMH, wn0, 0, lane);
                        m16::warp_gemm<4, 1, MT, true, true, LOWP>(acc_hb, sb + S::DG_HI, sb + S::DG_LO, LDGH, wh ? 3 * D : 32, 0, sb + S::HS_HI, sb + S::HS_LO, LDMH, wn0, 0, lane);
                    }
                    m16::warp_gemm<4, 1, MT, true, true, LOWP>(acc_v, sb + S::DM_HI, sb + S::DM_LO, LDMH, 0, 0, sb + S::XS_HI, sb + S::XS_LO, LDXH, vn0, 0, lane);
                    __syncthreads();
                    SWTRACE(6);
                }
            }
            SWTRACE(9);
            // ------------------------------------------------ pull-only nodes (level 0 / codes without a module)
            {
                const int gw = blockIdx.x * (BTHREADS / 32) + warp, nw = gridDim.x * (BTHREADS / 32);
                for (int c = 0; c < MGV_NCODE; ++c) {
                    if (lvl >= 1 && ((p.handled >> c) & 1u)) continue;
                    const int sbeg = p.seg_ptr[lvl * MGV_NCODE + c], send = p.seg_ptr[lvl * MGV_NCODE + c + 1];
                    for (int tt = sbeg + gw; tt < send; tt += nw) {
                        const int node = p.order[tt];
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
            SWTRACE(7);
            mgv_grid_sync(p.bar, gridDim.x);
            SWTRACE(8);
        }
    }
    if (p.trace && tid == 0)
        for (int i = 0; i < 10; ++i) p.trace[(size_t)blockIdx.x * 16 + i] = tr_acc[i];
    // ---------------------------------------------------- flush accumulators, reduce over the CTAs of a code
    {
        float* part = p.partial + (size_t)blockIdx.x * GRAD;
        for (int i = tid; i < GRAD; i += BTHREADS) part[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(Au + 4 * lane + k, du_acc[k]);
        __syncthreads();
        const float un = 1.0f / acc_scale;
#pragma unroll
        for (int m = 0; m < 6; ++m) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int o = wh * 96 + 16 * m + g + ((e & 2) ? 8 : 0), c = wn8 + 2 * t + (e & 1);
                part[G_WIH + o * D + c] = acc_ih[m][0][e] * un;
                part[G_WHH + o * D + c] = (m < 2 ? acc_ha[m][0][e] : acc_hb[m - 2][0][e]) * un;
            }
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int o = 16 * m + g + ((e & 2) ? 8 : 0), k = 8 * warp + 2 * t + (e & 1);
                part[G_WV + o * D2 + k] = acc_v[m][0][e] * un;
            }
        }
        for (int i = tid; i < 576; i += BTHREADS)
            part[i < 128 ? G_U + i : (i < 192 ? G_BV + i - 128 : (i < 384 ? G_BIH + i - 192 : G_BHH + i - 384))] = ACC[i];
    }
    mgv_grid_sync(p.bar, gridDim.x);
    const size_t total = (size_t)MGV_NCODE * GRAD;
    for (size_t idx = (size_t)blockIdx.x * BTHREADS + tid; idx < total; idx += (size_t)gridDim.x * BTHREADS) {
        const int c = (int)(idx / GRAD), e = (int)(idx % GRAD);
        float sacc = 0.f;
        for (int bb = p.cta_start[c]; bb < p.cta_start[c + 1]; ++bb) sacc += p.partial[(size_t)bb * GRAD + e];
        p.grads[idx] = sacc;
    }
}

// Static CTA -> code assignment, proportional to the estimated work of each code's nodes at level >= 1: a fixed part per node
// (dense products, pointwise, store) plus a part per predecessor row gathered.  The mean fan-in per code is not known on the
// host; NOT (code 2 in every gate library of the reference) has one predecessor, the other codes share the remaining in-edges.
// At least one CTA per code that has nodes.
void assign_ctas(const int64_t* count, unsigned handled, int grid, int* start, int64_t E = -1) {
    int n[MGV_NCODE];
    double w[MGV_NCODE];
    double total = 0;
    int active = 0;
    double others = 0, nots = 0;
    for (int c = 0; c < MGV_NCODE; ++c) {
        if (!(((handled >> c) & 1u) && count[c] > 0)) continue;
        if (c == 2) nots += (double)count[c]; else others += (double)count[c];
    }
    double fan_other = 2.0;
    if (E >= 0 && others > 0) {
        fan_other = ((double)E - nots) / others;
        fan_other = fan_other < 1.0 ? 1.0 : (fan_other > 8.0 ? 8.0 : fan_other);
    }
    const char* env = getenv("MGV_SWEEP_FIXED_COST");
    const double fixed = env ? atof(env) : 24.0;          // measured: a row costs about the same whatever its fan-in
    for (int c = 0; c < MGV_NCODE; ++c) {
        n[c] = 0; w[c] = 0;
        if (((handled >> c) & 1u) && count[c] > 0) {
            w[c] = (double)count[c] * (fixed + (c == 2 ? 1.0 : fan_other));
            total += w[c]; ++active;
        }
    }
    if (active > 0) {
        int used = 0;
        for (int c = 0; c < MGV_NCODE; ++c) {
            if (w[c] <= 0) continue;
            int k = (int)((double)grid * w[c] / total);
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
                const double l = w[c] / n[c];
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
    MGV_REQUIRE(sch->streams <= 1 || sweep_use_tc(rounds),
                "level sweep: multi-round sweeps (and MGV_SWEEP=mma) need single-stream level lists (mgv_build_level_lists streams = 1)");
    d.ghs = d.ghf = d.dxb = d.alpha = d.dscore = d.partial = d.grads = nullptr;
    d.trace = nullptr;
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
                                   int32_t* sync, int32_t precision, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SweepDev d;
    int rc = fill_common(d, sch, rounds, handled_mask, weights, hs, sync);
    if (rc != MGV_OK) return rc;
    d.hf_all = hf_all;
    if (sch->N == 0 || sch->L <= 1) return MGV_OK;          // no level >= 1: hf stays zero
    int grid = 0;
    const size_t smem = (size_t)F_SMEM_BYTES;
    MGV_REQUIRE(precision == 0 || precision == 1, "level sweep: precision must be 0 (fp32-accurate) or 1 (bf16)");
    if (sweep_use_tc(rounds)) {
        rc = mgv_sweep_tc_grid(&grid);
        if (rc != MGV_OK) return rc;
        assign_ctas(sch->code_count, handled_mask, grid, d.cta_start, sch->E);
        if (d.cta_start[MGV_NCODE] == 0) return MGV_OK;
        MGV_CUDA(cudaMemsetAsync(sync, 0, 64 * sizeof(int32_t), st));
        return mgv_sweep_tc_fwd(sch, handled_mask, d.cta_start, grid, weights, hs, hf_all, sync, precision, st);
    }
    const void* fkern = precision == 1 ? (const void*)sweep_fwd_kernel<true> : (const void*)sweep_fwd_kernel<false>;
    rc = coop_grid(fkern, smem, FTHREADS, &grid);
    if (rc != MGV_OK) return rc;
    assign_ctas(sch->code_count, handled_mask, grid, d.cta_start);
    if (d.cta_start[MGV_NCODE] == 0) return MGV_OK;          // nothing to propagate
    MGV_CUDA(cudaMemsetAsync(sync, 0, 64 * sizeof(int32_t), st));
    void* args[] = {&d};
    MGV_CUDA(cudaLaunchCooperativeKernel((void*)fkern, dim3(grid), dim3(FTHREADS), args, smem, st));
    mgv_count_launches(1);
    return MGV_OK;
}

extern "C" int mgv_sweep_bwd_grid(void) {
    int g1 = 0, g2 = 0;
    if (coop_grid((const void*)sweep_bwd_kernel<16, true, false>, (size_t)BwdSmem<16, true>::BYTES, BTHREADS, &g1) != MGV_OK) return -1;
    if (coop_grid((const void*)sweep_bwd_kernel<32, false, false>, (size_t)BwdSmem<32, false>::BYTES, BTHREADS, &g2) != MGV_OK) return -1;
    return g1 > g2 ? g1 : g2;
}

extern "C" size_t mgv_sweep_bwd_workspace_bytes(int64_t N, int64_t E) {
    int grid = mgv_sweep_bwd_grid();
    if (grid < 1) grid = 1;
    size_t b = 0;
    b += mgv_align_up((size_t)N * D2 * 4 + 256, 256);          // dxb
    b += 2 * mgv_align_up((size_t)E * 4 + 256, 256);           // alpha, dscore
    b += mgv_align_up((size_t)grid * GRAD * 4 + 256, 256);     // partial
    const size_t tcb = mgv_sweep_tc_bwd_workspace_bytes(N, E);
    return (b > tcb ? b : tcb) + 1024;
}

extern "C" int mgv_level_sweep_bwd(const mgv_schedule* sch, int32_t rounds, uint32_t handled_mask,
                                   const float* weights, const float* hs, const float* hf_all,
                                   float* ghs, float* ghf, float* grads,
                                   void* ws, size_t ws_bytes, int32_t* sync, int32_t precision, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SweepDev d;
    int rc = fill_common(d, sch, rounds, handled_mask, weights, hs, sync);
    if (rc != MGV_OK) return rc;
    d.hf_all = const_cast<float*>(hf_all);
    MGV_CUDA(cudaMemsetAsync(grads, 0, (size_t)MGV_NCODE * GRAD * sizeof(float), st));
    if (sch->N == 0 || sch->L <= 1) return MGV_OK;
    int grid = 0;
    const bool single = rounds == 1;                          // h = 0 everywhere: the 32-node variant without W_hh
    MGV_REQUIRE(precision == 0 || precision == 1, "level sweep: precision must be 0 (fp32-accurate) or 1 (bf16)");
    if (sweep_use_tc(rounds) && mgv_sweep_tc_bwd_available()) {
        rc = mgv_sweep_tc_grid(&grid);
        if (rc != MGV_OK) return rc;
        assign_ctas(sch->code_count, handled_mask, grid, d.cta_start, sch->E);
        if (d.cta_start[MGV_NCODE] == 0) return MGV_OK;
        if (ws_bytes < mgv_sweep_bwd_workspace_bytes(sch->N, sch->E)) {
            mgv_set_error("mgv_level_sweep_bwd: workspace %zu < %zu bytes", ws_bytes, mgv_sweep_bwd_workspace_bytes(sch->N, sch->E));
            return MGV_ERR_WORKSPACE;
        }
        MGV_CUDA(cudaMemsetAsync(sync, 0, 64 * sizeof(int32_t), st));
        return mgv_sweep_tc_bwd(sch, handled_mask, d.cta_start, grid, weights, hs, hf_all, ghs, ghf, grads, ws, ws_bytes, sync, precision, st);
    }
    const void* kern = single ? (precision == 1 ? (const void*)sweep_bwd_kernel<32, false, true> : (const void*)sweep_bwd_kernel<32, false, false>)
                              : (precision == 1 ? (const void*)sweep_bwd_kernel<16, true, true> : (const void*)sweep_bwd_kernel<16, true, false>);
    const size_t smem = single ? (size_t)BwdSmem<32, false>::BYTES : (size_t)BwdSmem<16, true>::BYTES;
    rc = coop_grid(kern, smem, BTHREADS, &grid);
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
    d.trace = mgv_debug_trace();
    MGV_CUDA(cudaMemsetAsync(sync, 0, 64 * sizeof(int32_t), st));
    void* args[] = {&d};
    MGV_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(BTHREADS), args, smem, st));
    mgv_count_launches(1);
    return MGV_OK;
}
