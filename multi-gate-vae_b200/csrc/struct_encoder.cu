// Struct encoder BACKWARD (mma.sync 3xTF32 generation; the forward lives in struct_tc.cu on tcgen05).
// MultiGCNEncoder.forward (digae_layer.py:257-277) with AggConv
// (arch/gcn_conv.py:30-42), one fused kernel per half-round step:
//   gather-sum of neighbour states -> GRU_{70->64}([W agg + deg b || x], state) -> LayerNorm
// Step k = 1..2R uses in-neighbours when k is odd (aggr/update) and out-neighbours when k is
// even (aggr_r/update_r); LayerNorm parameters are shared by both directions (digae_layer.py:270,275).
// source_conv and target_conv (digae_layer.py:294-297) run batched: blockIdx.y = encoder.
//
// The AggConv linear is pre-composed into the GRU input weights on the host:
//   W_ih[:, :64] (W agg + deg b) = Wc agg + deg bc,   Wc = W_ih[:, :64] W,  bc = W_ih[:, :64] b
// so a step is ONE tile GEMM  [agg || x || h] (32 x 136)  x  [Wc | W_ih_x | W_hh]^T  on the tensor
// cores (3xTF32, fp32 accuracy), with one copy of the weights resident in shared memory for the
// whole launch (persistent CTAs loop over 32-node tiles).
// Backward recomputes the step from the saved post-LN states, runs LN/GRU backward in registers,
// the two data-gradient GEMMs from the same weight copy, and keeps the weight-gradient
// accumulators as persistent MMA fragments in registers (flushed once per launch).
#include "mgv_mma.cuh"

namespace {

constexpr int D = MGV_D;              // 64
constexpr int G3 = 3 * D;             // 192
constexpr int KX = D + MGV_MAX_FEAT;  // 72: [agg || x]
constexpr int SPACK = MGV_STRUCT_PACK_FLOATS;
constexpr int SGRAD = MGV_STRUCT_GRAD_FLOATS;
constexpr int LDC = KX + 4;           // 76: row stride of Wcx and of the [agg || x] tile
constexpr int LDM = D + 4;            // 68
constexpr int LDG = G3 + 8;           // 200: d gi / d gh tiles (also read transposed)
// weight / gradient block offsets (floats) -- see include/mgv_b200.h
constexpr int O_WCX = 0, O_WHH = 14592, O_BC = 27648, O_BIH = 27840, O_BHH = 28032, O_LNW = 28224, O_LNB = 28288;
constexpr int NODE_MASK = (1 << MGV_CODE_SHIFT) - 1;

constexpr int THREADS = 512;
constexpr int WARPS = THREADS / 32;   // 16
constexpr int TM = 32;                // nodes per tile
constexpr float LN_EPS = 1e-5f;

static_assert(O_LNB + D <= SPACK && SPACK % 4 == 0, "struct pack layout");

struct StepDev {
    int N, feat, layernorm, first, last, dir;
    const int* ptr;            // neighbour CSR of this step's direction
    const int* idx;
    const float* x;            // [N][feat]
    const float* weights;      // block of (enc 0, this dir); encoder stride 2*SPACK
    const float* prev;         // state_{k-1}, enc 0
    float* next;               // state_k, enc 0 (forward only)
    size_t enc_stride;         // floats between encoders in the states buffer
    // backward
    const float* gout;         // [enc][N][64]
    const float* in_part; const float* in_agg;
    float* out_part; float* out_agg;     // [enc][N][64]
    float* partial;            // [enc][gx][2][SGRAD]
};

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

__device__ __forceinline__ void load_weights(float* Ws, const float* Wg, int tid) {
    for (int i = tid * 4; i < SPACK; i += THREADS * 4) mgv_st4(Ws + i, mgv_ldg4(Wg + i));
}

// GRU pre-activations of the 4 fragment elements a thread owns (rows g / g+8, units u0+2t / +1).
struct Gates { float r[4], z[4], n[4], hnb[4]; };

__device__ __forceinline__ void step_gemm(const float* Ws, const float* As, const float* Hs, const float* Dg,
                                          int mt, int u0, int lane, Gates& o) {
    const int n0[3] = {u0, D + u0, 2 * D + u0};
    float ci[1][3][4], ch[1][3][4];
    mgv_zero_frag(ci);
    mgv_zero_frag(ch);
    mgv_warp_gemm<1, 3, KX / 8, false, true>(ci, As, LDC, mt * 16, Ws + O_WCX, LDC, n0, lane);
    mgv_warp_gemm<1, 3, D / 8, false, true>(ch, Hs, LDM, mt * 16, Ws + O_WHH, LDM, n0, lane);
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int row = mt * 16 + g + ((e & 2) ? 8 : 0);
        const int u = u0 + 2 * t + (e & 1);
        const float dg = Dg[row];
        const float gir = ci[0][0][e] + dg * Ws[O_BC + u] + Ws[O_BIH + u];
        const float giz = ci[0][1][e] + dg * Ws[O_BC + D + u] + Ws[O_BIH + D + u];
        const float gin = ci[0][2][e] + dg * Ws[O_BC + 2 * D + u] + Ws[O_BIH + 2 * D + u];
        o.r[e] = mgv_sigmoid(gir + ch[0][0][e] + Ws[O_BHH + u]);
        o.z[e] = mgv_sigmoid(giz + ch[0][1][e] + Ws[O_BHH + D + u]);
        o.hnb[e] = ch[0][2][e] + Ws[O_BHH + 2 * D + u];
        o.n[e] = tanhf(gin + o.r[e] * o.hnb[e]);
    }
}

// ======================================================================================= backward step
constexpr int B_SMEM_FLOATS = SPACK + TM * LDC + 3 * TM * LDM + 2 * TM * LDG + TM + 2 * D;

__global__ void __launch_bounds__(THREADS, 1) struct_bwd_kernel(const StepDev p) {
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;
    float* As = Ws + SPACK;              // [32][76] [neighbour sum of state_{k-1} || x]
    float* Hs = As + TM * LDC;           // [32][68] state_{k-1} of the node
    float* Gs = Hs + TM * LDM;           // [32][68] d state_k, then d (pre-LN GRU output)
    float* Xh = Gs + TM * LDM;           // [32][68] pre-LN GRU output
    float* DGI = Xh + TM * LDM;          // [32][200] d gi (r, z, n)
    float* DGH = DGI + TM * LDG;         // [32][200] d gh
    float* Dg = DGH + TM * LDG;          // [32]
    float* LNA = Dg + TM;                // [128] d ln_w, d ln_b accumulators
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int enc = blockIdx.y;
    const float* prev = p.prev + (size_t)enc * p.enc_stride;
    const size_t eoff = (size_t)enc * p.N * D;
    load_weights(Ws, p.weights + (size_t)enc * 2 * SPACK, tid);
    if (tid < 2 * D) LNA[tid] = 0.f;
    const int ntiles = (p.N + TM - 1) / TM;
    const int half = lane >> 4, l16 = lane & 15;
    const int mt = warp & 1, u0 = (warp >> 1) * 8;
    const int g = lane >> 2, t = lane & 3;

    // persistent weight-gradient fragments: d Wcx[:, 0:64] and d Whh as 6 m-tiles x 1 n-tile per warp,
    // the feature columns of d Wcx as one fragment on warps 0..11
    const int wn0[1] = {8 * (warp & 7)};
    const int wm0 = (warp >> 3) * 96;
    const int fn0[1] = {D};
    float acc_cx[6][1][4], acc_hh[6][1][4], acc_f[1][1][4];
    mgv_zero_frag(acc_cx);
    mgv_zero_frag(acc_hh);
    mgv_zero_frag(acc_f);
    float s_bc = 0.f, s_bih = 0.f, s_bhh = 0.f;      // column sums owned by threads 0..191

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t0 = tile * TM;
        {   // ---- gathers: neighbour sum of state_{k-1}; d state_k = part + sum of neighbours' d agg_{k+1}
            const int row = warp * 2 + half, node = t0 + row;
            float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa, h4 = sa, g4 = sa;
            int deg = 0;
            float xf = 0.f;
            if (node < p.N) {
                if (p.last) {
                    gather_sum<false>(p, prev, nullptr, node, l16, sa, sb, deg);
                    g4 = mgv_ld4(p.gout + eoff + (size_t)node * D + 4 * l16);
                } else {
                    gather_sum<true>(p, prev, p.in_agg + eoff, node, l16, sa, sb, deg);
                    g4 = mgv_ld4(p.in_part + eoff + (size_t)node * D + 4 * l16);
                    add4(g4, sb);
                }
                h4 = mgv_ld4(prev + (size_t)node * D + 4 * l16);
                if (l16 < p.feat) xf = p.x[(size_t)node * p.feat + l16];
            }
            mgv_st4(As + row * LDC + 4 * l16, sa);
            mgv_st4(Hs + row * LDM + 4 * l16, h4);
            mgv_st4(Gs + row * LDM + 4 * l16, g4);
            if (l16 < MGV_MAX_FEAT) As[row * LDC + D + l16] = xf;
            if (l16 == 0) Dg[row] = (float)deg;
        }
        __syncthreads();
        // ---- recompute the step (gates stay in registers)
        Gates G;
        step_gemm(Ws, As, Hs, Dg, mt, u0, lane, G);
        if (p.layernorm) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int row = mt * 16 + g + ((e & 2) ? 8 : 0);
                const int u = u0 + 2 * t + (e & 1);
                Xh[row * LDM + u] = (1.0f - G.z[e]) * G.n[e] + G.z[e] * Hs[row * LDM + u];
            }
            __syncthreads();
            // LayerNorm backward, two rows per warp; d ln_w / d ln_b accumulate in shared memory
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int row = warp * 2 + rr;
                const float v0 = Xh[row * LDM + lane], v1 = Xh[row * LDM + 32 + lane];
                const float mean = mgv_warp_sum(v0 + v1) * (1.0f / D);
                const float d0 = v0 - mean, d1 = v1 - mean;
                const float var = mgv_warp_sum(d0 * d0 + d1 * d1) * (1.0f / D);
                const float rstd = 1.0f / sqrtf(var + LN_EPS);
                const float x0 = d0 * rstd, x1 = d1 * rstd;
                const float gy0 = Gs[row * LDM + lane], gy1 = Gs[row * LDM + 32 + lane];
                atomicAdd(LNA + lane, gy0 * x0);
                atomicAdd(LNA + 32 + lane, gy1 * x1);
                atomicAdd(LNA + D + lane, gy0);
                atomicAdd(LNA + D + 32 + lane, gy1);
                const float dx0 = gy0 * Ws[O_LNW + lane], dx1 = gy1 * Ws[O_LNW + 32 + lane];
                const float c1 = mgv_warp_sum(dx0 + dx1) * (1.0f / D);
                const float c2 = mgv_warp_sum(dx0 * x0 + dx1 * x1) * (1.0f / D);
                Gs[row * LDM + lane] = rstd * (dx0 - c1 - x0 * c2);
                Gs[row * LDM + 32 + lane] = rstd * (dx1 - c1 - x1 * c2);
            }
            __syncthreads();
        }
        // ---- GRU backward on the owned elements
        float dh_direct[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int row = mt * 16 + g + ((e & 2) ? 8 : 0);
            const int u = u0 + 2 * t + (e & 1);
            const float gg = Gs[row * LDM + u];
            const float hp = Hs[row * LDM + u];
            const float dn = gg * (1.0f - G.z[e]);
            const float dz = gg * (hp - G.n[e]);
            const float dnpre = dn * (1.0f - G.n[e] * G.n[e]);
            const float drpre = dnpre * G.hnb[e] * G.r[e] * (1.0f - G.r[e]);
            const float dzpre = dz * G.z[e] * (1.0f - G.z[e]);
            dh_direct[e] = gg * G.z[e];
            DGI[row * LDG + u] = drpre;
            DGI[row * LDG + D + u] = dzpre;
            DGI[row * LDG + 2 * D + u] = dnpre;
            DGH[row * LDG + u] = drpre;
            DGH[row * LDG + D + u] = dzpre;
            DGH[row * LDG + 2 * D + u] = dnpre * G.r[e];
        }
        __syncthreads();
        // ---- data gradients: d agg = d gi . Wc ;  d part = g z + d gh . Whh   (both 32 x 64, K = 192)
        if (!p.first) {
            const int n0[1] = {u0};
            float ca[1][1][4], cp[1][1][4];
            mgv_zero_frag(ca);
            mgv_zero_frag(cp);
            mgv_warp_gemm<1, 1, G3 / 8, false, false>(ca, DGI, LDG, mt * 16, Ws + O_WCX, LDC, n0, lane);
            mgv_warp_gemm<1, 1, G3 / 8, false, false>(cp, DGH, LDG, mt * 16, Ws + O_WHH, LDM, n0, lane);
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int node = t0 + mt * 16 + g + 8 * hrow;
                if (node < p.N) {
                    const size_t o = eoff + (size_t)node * D + u0 + 2 * t;
                    *reinterpret_cast<float2*>(p.out_agg + o) = make_float2(ca[0][0][2 * hrow], ca[0][0][2 * hrow + 1]);
                    *reinterpret_cast<float2*>(p.out_part + o) =
                        make_float2(cp[0][0][2 * hrow] + dh_direct[2 * hrow], cp[0][0][2 * hrow + 1] + dh_direct[2 * hrow + 1]);
                }
            }
        }
        // ---- weight gradients (tile buffers are read-only here): d Wcx += d gi^T [agg || x], d Whh += d gh^T h
        mgv_warp_gemm<6, 1, TM / 8, true, false>(acc_cx, DGI, LDG, wm0, As, LDC, wn0, lane);
        mgv_warp_gemm<6, 1, TM / 8, true, false>(acc_hh, DGH, LDG, wm0, Hs, LDM, wn0, lane);
        if (warp < 12) mgv_warp_gemm<1, 1, TM / 8, true, false>(acc_f, DGI, LDG, warp * 16, As, LDC, fn0, lane);
        if (tid < G3) {
            for (int row = 0; row < TM; ++row) {
                const float dgi = DGI[row * LDG + tid];
                s_bih += dgi;
                s_bc = fmaf(dgi, Dg[row], s_bc);
                s_bhh += DGH[row * LDG + tid];
            }
        }
        __syncthreads();
    }
    // ---- flush this CTA's accumulators into its private partial block (summed over the 8 launches)
    float* part = p.partial + (((size_t)enc * gridDim.x + blockIdx.x) * 2 + p.dir) * SGRAD;
#pragma unroll
    for (int m = 0; m < 6; ++m) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int o = wm0 + 16 * m + g + ((e & 2) ? 8 : 0);
            const int c = wn0[0] + 2 * t + (e & 1);
            part[O_WCX + o * LDC + c] += acc_cx[m][0][e];
            part[O_WHH + o * LDM + c] += acc_hh[m][0][e];
        }
    }
    if (warp < 12) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int o = warp * 16 + g + ((e & 2) ? 8 : 0);
            part[O_WCX + o * LDC + D + 2 * t + (e & 1)] += acc_f[0][0][e];
        }
    }
    if (tid < G3) {
        part[O_BC + tid] += s_bc;
        part[O_BIH + tid] += s_bih;
        part[O_BHH + tid] += s_bhh;
    }
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
    p.next = nullptr;
    p.enc_stride = enc_stride;
}

}  // namespace

extern "C" int mgv_struct_bwd_grid(void) {
    int gx = 0;
    if (persistent_gx(1, 1 << 30, &gx) != MGV_OK) return -1;
    return gx;
}

extern "C" size_t mgv_struct_bwd_workspace_bytes(int64_t N, int32_t num_enc) {
    int gx = mgv_struct_bwd_grid();
    if (gx < 1) gx = 256;
    size_t b = 0;
    b += 4 * mgv_align_up((size_t)num_enc * N * D * 4 + 256, 256);              // part/agg ping-pong
    b += mgv_align_up((size_t)gx * 2 * SGRAD * 4 + 256, 256);                    // partial: num_enc * (sms / num_enc) <= sms CTAs
    return b + 1024;
}

extern "C" int mgv_struct_encoder_bwd(const mgv_schedule* sch, int32_t num_enc, int32_t rounds, int32_t layernorm,
                                      int32_t feat, const float* x, const float* weights, const float* states,
                                      const float* gout, float* grads, void* ws, size_t ws_bytes,
                                      mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_args(sch, num_enc, rounds, feat);
    if (rc != MGV_OK) return rc;
    const int N = sch->N;
    MGV_CUDA(cudaMemsetAsync(grads, 0, (size_t)num_enc * 2 * SGRAD * sizeof(float), st));
    if (N == 0) return MGV_OK;
    if (ws_bytes < mgv_struct_bwd_workspace_bytes(N, num_enc)) {
        mgv_set_error("mgv_struct_encoder_bwd: workspace %zu < %zu bytes", ws_bytes,
                      mgv_struct_bwd_workspace_bytes(N, num_enc));
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
    const size_t smem = (size_t)B_SMEM_FLOATS * sizeof(float);
    MGV_CUDA(cudaFuncSetAttribute((const void*)struct_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int k = steps; k >= 1; --k) {
        StepDev p{};
        fill_step(p, sch, k, steps, layernorm, feat, x, weights, states, slot, enc_stride);
        p.gout = gout;
        p.in_part = part[k & 1]; p.in_agg = agg[k & 1];
        p.out_part = part[(k - 1) & 1]; p.out_agg = agg[(k - 1) & 1];
        p.partial = partial;
        struct_bwd_kernel<<<dim3(gx, num_enc), THREADS, smem, st>>>(p);
        mgv_count_launches(1);
    }
    const size_t total = (size_t)num_enc * 2 * SGRAD;
    struct_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, gx, grads, num_enc);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_struct_encoder_bwd");
}
