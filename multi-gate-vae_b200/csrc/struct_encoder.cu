// Struct encoder: MultiGCNEncoder.forward (digae_layer.py:257-277) with AggConv
// (arch/gcn_conv.py:30-42), one fused kernel per half-round step:
//   gather-sum of neighbour states -> msg = W agg + deg b -> GRU_{70->64}([msg || x], state) -> LayerNorm
// Step k = 1..2R uses in-neighbours when k is odd (aggr/update) and out-neighbours when k is
// even (aggr_r/update_r); LayerNorm parameters are shared by both directions (digae_layer.py:270,275).
// source_conv and target_conv (digae_layer.py:294-297) run batched: blockIdx.y = encoder.
// Backward recomputes each step from the saved post-LN states and keeps the weight-gradient
// accumulators of a CTA in shared memory for the whole launch.
#include "mgv_common.cuh"

namespace {

constexpr int D = MGV_D;              // 64
constexpr int G3 = 3 * D;             // 192
constexpr int KX = D + MGV_MAX_FEAT;  // 72: [msg || x || pad]
constexpr int SPACK = MGV_STRUCT_PACK_FLOATS;
constexpr int SGRAD = MGV_STRUCT_GRAD_FLOATS;
constexpr int O_WT = 0, O_B = 4096, O_WIHT = 4160, O_WHHT = 17984, O_BIH = 30272, O_BHH = 30464, O_LNW = 30656, O_LNB = 30720;
constexpr int O_W = 30912, O_WIH = 35008, O_WHH = 48832;
constexpr int G_W = 0, G_B = 4096, G_WIH = 4160, G_WHH = 17984, G_BIH = 30272, G_BHH = 30464, G_LNW = 30656, G_LNB = 30720;
constexpr int NODE_MASK = (1 << MGV_CODE_SHIFT) - 1;

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int LDM = D + 4;            // 68
constexpr int LDK = KX + 4;           // 76
constexpr int LDG = G3 + 4;           // 196
constexpr float LN_EPS = 1e-5f;

struct StepDev {
    int N, feat, layernorm, first, last;
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
    int dir;
};

// Sum of neighbour rows (and optionally of a second array over the same neighbours); the two
// half-warps take alternate neighbours, 16 lanes x float4 cover a 64-wide row.
template <bool TWO>
__device__ __forceinline__ void gather_sum(const StepDev& p, const float* a, const float* b, int node, int lane,
                                           float4& sa, float4& sb, int& deg) {
    const int beg = p.ptr[node], end = p.ptr[node + 1];
    deg = end - beg;
    const int half = lane >> 4, l16 = lane & 15;
    sa = make_float4(0.f, 0.f, 0.f, 0.f);
    sb = sa;
    for (int q0 = beg; q0 < end; q0 += 4) {
        float4 va[2], vb[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int q = q0 + 2 * i + half;
            va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            vb[i] = va[i];
            if (q < end) {
                const int j = p.idx[q] & NODE_MASK;
                va[i] = mgv_ld4(a + (size_t)j * D + 4 * l16);
                if (TWO) vb[i] = mgv_ld4(b + (size_t)j * D + 4 * l16);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            sa.x += va[i].x; sa.y += va[i].y; sa.z += va[i].z; sa.w += va[i].w;
            if (TWO) { sb.x += vb[i].x; sb.y += vb[i].y; sb.z += vb[i].z; sb.w += vb[i].w; }
        }
    }
    sa.x += __shfl_xor_sync(0xffffffffu, sa.x, 16); sa.y += __shfl_xor_sync(0xffffffffu, sa.y, 16);
    sa.z += __shfl_xor_sync(0xffffffffu, sa.z, 16); sa.w += __shfl_xor_sync(0xffffffffu, sa.w, 16);
    if (TWO) {
        sb.x += __shfl_xor_sync(0xffffffffu, sb.x, 16); sb.y += __shfl_xor_sync(0xffffffffu, sb.y, 16);
        sb.z += __shfl_xor_sync(0xffffffffu, sb.z, 16); sb.w += __shfl_xor_sync(0xffffffffu, sb.w, 16);
    }
}

// ======================================================================================= forward step
constexpr int FTM = 32;
constexpr int F_SMEM_FLOATS = 3 * FTM * LDM + FTM * LDK + FTM;

__global__ void __launch_bounds__(THREADS, 2) struct_fwd_kernel(const StepDev p) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                    // [32][68] neighbour sum
    float* Hs = As + FTM * LDM;          // [32][68] own state
    float* Os = Hs + FTM * LDM;          // [32][68] GRU output (pre-LN)
    float* Ms = Os + FTM * LDM;          // [32][76] [msg || x || 0]
    float* Dg = Ms + FTM * LDK;          // [32] degree
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int enc = blockIdx.y;
    const float* W = p.weights + (size_t)enc * 2 * SPACK;
    const float* prev = p.prev + (size_t)enc * p.enc_stride;
    float* next = p.next + (size_t)enc * p.enc_stride;
    const int t0 = blockIdx.x * FTM;
    const int col = tid & 63, rg = tid >> 6;

    for (int row = warp; row < FTM; row += WARPS) {
        const int node = t0 + row;
        float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa, h4 = sa;
        int deg = 0;
        float xf = 0.f;
        if (node < p.N) {
            gather_sum<false>(p, prev, nullptr, node, lane, sa, sb, deg);
            if (lane >= 16) h4 = mgv_ld4(prev + (size_t)node * D + 4 * (lane - 16));
            if (lane < p.feat) xf = p.x[(size_t)node * p.feat + lane];
        }
        if (lane < 16) mgv_st4(As + row * LDM + 4 * lane, sa);
        else mgv_st4(Hs + row * LDM + 4 * (lane - 16), h4);
        if (lane < MGV_MAX_FEAT) Ms[row * LDK + D + lane] = xf;
        if (lane == 0) Dg[row] = (float)deg;
    }
    __syncthreads();
    {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        mgv_gemm_col<8, D>(As + rg * 8 * LDM, LDM, W + O_WT, D, col, acc);
        const float b = __ldg(W + O_B + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) Ms[(rg * 8 + i) * LDK + col] = fmaf(b, Dg[rg * 8 + i], acc[i]);
    }
    __syncthreads();
    {
        float ar[8], az[8], an[8], hr[8], hz[8], hn[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { ar[i] = az[i] = an[i] = 0.f; hr[i] = hz[i] = hn[i] = 0.f; }
        mgv_gemm_col3<8, KX>(Ms + rg * 8 * LDK, LDK, W + O_WIHT, col, ar, az, an);
        mgv_gemm_col3<8, D>(Hs + rg * 8 * LDM, LDM, W + O_WHHT, col, hr, hz, hn);
        const float bir = __ldg(W + O_BIH + col), biz = __ldg(W + O_BIH + D + col), bin = __ldg(W + O_BIH + 2 * D + col);
        const float bhr = __ldg(W + O_BHH + col), bhz = __ldg(W + O_BHH + D + col), bhn = __ldg(W + O_BHH + 2 * D + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = rg * 8 + i;
            const float rr = mgv_sigmoid(ar[i] + bir + hr[i] + bhr);
            const float zz = mgv_sigmoid(az[i] + biz + hz[i] + bhz);
            const float nn = tanhf(an[i] + bin + rr * (hn[i] + bhn));
            const float o = (1.0f - zz) * nn + zz * Hs[row * LDM + col];
            if (p.layernorm) Os[row * LDM + col] = o;
            else if (t0 + row < p.N) next[(size_t)(t0 + row) * D + col] = o;
        }
    }
    if (!p.layernorm) return;
    __syncthreads();
    const float g0 = __ldg(W + O_LNW + lane), g1 = __ldg(W + O_LNW + 32 + lane);
    const float b0 = __ldg(W + O_LNB + lane), b1 = __ldg(W + O_LNB + 32 + lane);
    for (int row = warp; row < FTM; row += WARPS) {
        const int node = t0 + row;
        if (node >= p.N) continue;
        const float v0 = Os[row * LDM + lane], v1 = Os[row * LDM + 32 + lane];
        const float mean = mgv_warp_sum(v0 + v1) * (1.0f / D);
        const float d0 = v0 - mean, d1 = v1 - mean;
        const float var = mgv_warp_sum(d0 * d0 + d1 * d1) * (1.0f / D);
        const float rstd = 1.0f / sqrtf(var + LN_EPS);
        next[(size_t)node * D + lane] = d0 * rstd * g0 + b0;
        next[(size_t)node * D + 32 + lane] = d1 * rstd * g1 + b1;
    }
}

__global__ void fill_ones_kernel(float* p, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 1.0f;
}

// ======================================================================================= backward step
constexpr int BTM = 16;
constexpr int B_SMEM_FLOATS = SGRAD + 5 * BTM * LDM + BTM * LDK + 2 * BTM * LDG + 2 * BTM;

__global__ void __launch_bounds__(THREADS, 1) struct_bwd_kernel(const StepDev p) {
    extern __shared__ __align__(16) float smem[];
    float* ACC = smem;                   // [SGRAD]
    float* As = ACC + SGRAD;             // [16][68] neighbour sum of state_{k-1}
    float* Hs = As + BTM * LDM;          // [16][68] state_{k-1} of the node
    float* Gs = Hs + BTM * LDM;          // [16][68] d state_k, then d (pre-LN GRU output)
    float* Xh = Gs + BTM * LDM;          // [16][68] pre-LN output, then xhat
    float* DMs = Xh + BTM * LDM;         // [16][68] d msg
    float* Ms = DMs + BTM * LDM;         // [16][76] [msg || x || 0]
    float* DGI = Ms + BTM * LDK;         // [16][196]
    float* DGH = DGI + BTM * LDG;        // [16][196]
    float* Dg = DGH + BTM * LDG;         // [16]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int enc = blockIdx.y;
    const float* W = p.weights + (size_t)enc * 2 * SPACK;
    const float* prev = p.prev + (size_t)enc * p.enc_stride;
    const size_t eoff = (size_t)enc * p.N * D;
    const int col = tid & 63, rg = tid >> 6;
    const int ntiles = (p.N + BTM - 1) / BTM;

    for (int i = tid; i < SGRAD; i += THREADS) ACC[i] = 0.f;
    __syncthreads();

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t0 = tile * BTM;
        // ---- phase A: gathers
        for (int row = warp; row < BTM; row += WARPS) {
            const int node = t0 + row;
            float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa, h4 = sa, g4 = sa;
            int deg = 0;
            float xf = 0.f;
            if (node < p.N) {
                if (p.last) {
                    gather_sum<false>(p, prev, nullptr, node, lane, sa, sb, deg);
                    if (lane < 16) g4 = mgv_ld4(p.gout + eoff + (size_t)node * D + 4 * lane);
                } else {
                    gather_sum<true>(p, prev, p.in_agg + eoff, node, lane, sa, sb, deg);
                    if (lane < 16) {
                        const float4 pt = mgv_ld4(p.in_part + eoff + (size_t)node * D + 4 * lane);
                        g4 = make_float4(pt.x + sb.x, pt.y + sb.y, pt.z + sb.z, pt.w + sb.w);
                    }
                }
                if (lane >= 16) h4 = mgv_ld4(prev + (size_t)node * D + 4 * (lane - 16));
                if (lane < p.feat) xf = p.x[(size_t)node * p.feat + lane];
            }
            if (lane < 16) { mgv_st4(As + row * LDM + 4 * lane, sa); mgv_st4(Gs + row * LDM + 4 * lane, g4); }
            else mgv_st4(Hs + row * LDM + 4 * (lane - 16), h4);
            if (lane < MGV_MAX_FEAT) Ms[row * LDK + D + lane] = xf;
            if (lane == 0) Dg[row] = (float)deg;
        }
        __syncthreads();
        // ---- phase B: msg
        {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            mgv_gemm_col<4, D>(As + rg * 4 * LDM, LDM, W + O_WT, D, col, acc);
            const float b = __ldg(W + O_B + col);
#pragma unroll
            for (int i = 0; i < 4; ++i) Ms[(rg * 4 + i) * LDK + col] = fmaf(b, Dg[rg * 4 + i], acc[i]);
        }
        __syncthreads();
        // ---- phase C: GRU recompute (registers), LayerNorm backward, GRU backward
        float rr[4], zz[4], nn[4], hnb[4];
        {
            float ar[4], az[4], an[4], hr[4], hz[4], hn[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { ar[i] = az[i] = an[i] = 0.f; hr[i] = hz[i] = hn[i] = 0.f; }
            mgv_gemm_col3<4, KX>(Ms + rg * 4 * LDK, LDK, W + O_WIHT, col, ar, az, an);
            mgv_gemm_col3<4, D>(Hs + rg * 4 * LDM, LDM, W + O_WHHT, col, hr, hz, hn);
            const float bir = __ldg(W + O_BIH + col), biz = __ldg(W + O_BIH + D + col), bin = __ldg(W + O_BIH + 2 * D + col);
            const float bhr = __ldg(W + O_BHH + col), bhz = __ldg(W + O_BHH + D + col), bhn = __ldg(W + O_BHH + 2 * D + col);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = rg * 4 + i;
                rr[i] = mgv_sigmoid(ar[i] + bir + hr[i] + bhr);
                zz[i] = mgv_sigmoid(az[i] + biz + hz[i] + bhz);
                hnb[i] = hn[i] + bhn;
                nn[i] = tanhf(an[i] + bin + rr[i] * hnb[i]);
                if (p.layernorm) Xh[row * LDM + col] = (1.0f - zz[i]) * nn[i] + zz[i] * Hs[row * LDM + col];
            }
        }
        if (p.layernorm) {
            __syncthreads();
            const float g0 = __ldg(W + O_LNW + lane), g1 = __ldg(W + O_LNW + 32 + lane);
            for (int row = warp; row < BTM; row += WARPS) {
                const float v0 = Xh[row * LDM + lane], v1 = Xh[row * LDM + 32 + lane];
                const float mean = mgv_warp_sum(v0 + v1) * (1.0f / D);
                const float d0 = v0 - mean, d1 = v1 - mean;
                const float var = mgv_warp_sum(d0 * d0 + d1 * d1) * (1.0f / D);
                const float rstd = 1.0f / sqrtf(var + LN_EPS);
                const float x0 = d0 * rstd, x1 = d1 * rstd;
                const float gy0 = Gs[row * LDM + lane], gy1 = Gs[row * LDM + 32 + lane];
                atomicAdd(ACC + G_LNW + lane, gy0 * x0);
                atomicAdd(ACC + G_LNW + 32 + lane, gy1 * x1);
                atomicAdd(ACC + G_LNB + lane, gy0);
                atomicAdd(ACC + G_LNB + 32 + lane, gy1);
                const float dx0 = gy0 * g0, dx1 = gy1 * g1;
                const float c1 = mgv_warp_sum(dx0 + dx1) * (1.0f / D);
                const float c2 = mgv_warp_sum(dx0 * x0 + dx1 * x1) * (1.0f / D);
                Gs[row * LDM + lane] = rstd * (dx0 - c1 - x0 * c2);
                Gs[row * LDM + 32 + lane] = rstd * (dx1 - c1 - x1 * c2);
            }
            __syncthreads();
        }
        float dh_direct[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = rg * 4 + i;
            const float g = Gs[row * LDM + col];
            const float hp = Hs[row * LDM + col];
            const float dn = g * (1.0f - zz[i]);
            const float dz = g * (hp - nn[i]);
            const float dnpre = dn * (1.0f - nn[i] * nn[i]);
            const float drpre = dnpre * hnb[i] * rr[i] * (1.0f - rr[i]);
            const float dzpre = dz * zz[i] * (1.0f - zz[i]);
            dh_direct[i] = g * zz[i];
            DGI[row * LDG + col] = drpre;
            DGI[row * LDG + D + col] = dzpre;
            DGI[row * LDG + 2 * D + col] = dnpre;
            DGH[row * LDG + col] = drpre;
            DGH[row * LDG + D + col] = dzpre;
            DGH[row * LDG + 2 * D + col] = dnpre * rr[i];
        }
        __syncthreads();
        // ---- phase D: d msg = d gi . Wih[:, :64] ; d part = g z + d gh . Whh
        {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            mgv_gemm_col<4, G3>(DGI + rg * 4 * LDG, LDG, W + O_WIH, KX, col, acc);
#pragma unroll
            for (int i = 0; i < 4; ++i) DMs[(rg * 4 + i) * LDM + col] = acc[i];
            if (!p.first) {
                float acch[4] = {0.f, 0.f, 0.f, 0.f};
                mgv_gemm_col<4, G3>(DGH + rg * 4 * LDG, LDG, W + O_WHH, D, col, acch);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int node = t0 + rg * 4 + i;
                    if (node < p.N) p.out_part[eoff + (size_t)node * D + col] = dh_direct[i] + acch[i];
                }
            }
        }
        __syncthreads();
        // ---- phase E: d agg = d msg . W
        if (!p.first) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            mgv_gemm_col<4, D>(DMs + rg * 4 * LDM, LDM, W + O_W, D, col, acc);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int node = t0 + rg * 4 + i;
                if (node < p.N) p.out_agg[eoff + (size_t)node * D + col] = acc[i];
            }
        }
        // ---- phase G: weight gradients (tile buffers are read-only here)
        {
            const int og = rg * 48;
            float acc[48];
#pragma unroll
            for (int i = 0; i < 48; ++i) acc[i] = 0.f;
            for (int row = 0; row < BTM; ++row) {
                const float mv = Ms[row * LDK + col];
#pragma unroll
                for (int i = 0; i < 48; i += 4) {
                    const float4 d4 = mgv_ld4(DGI + row * LDG + og + i);
                    acc[i] = fmaf(d4.x, mv, acc[i]); acc[i + 1] = fmaf(d4.y, mv, acc[i + 1]);
                    acc[i + 2] = fmaf(d4.z, mv, acc[i + 2]); acc[i + 3] = fmaf(d4.w, mv, acc[i + 3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 48; ++i) ACC[G_WIH + (og + i) * KX + col] += acc[i];
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
        {
            // dW[c][j] += sum_row dmsg[row][c] agg[row][j]   (c in [16 rg, 16 rg + 16), j = col)
            const int cg = rg * 16;
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
            for (int row = 0; row < BTM; ++row) {
                const float av = As[row * LDM + col];
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 d4 = mgv_ld4(DMs + row * LDM + cg + i);
                    acc[i] = fmaf(d4.x, av, acc[i]); acc[i + 1] = fmaf(d4.y, av, acc[i + 1]);
                    acc[i + 2] = fmaf(d4.z, av, acc[i + 2]); acc[i + 3] = fmaf(d4.w, av, acc[i + 3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) ACC[G_W + (cg + i) * D + col] += acc[i];
        }
        if (tid < G3) {
            float sgi = 0.f, sgh = 0.f;
            float fx[MGV_MAX_FEAT];
#pragma unroll
            for (int f = 0; f < MGV_MAX_FEAT; ++f) fx[f] = 0.f;
            for (int row = 0; row < BTM; ++row) {
                const float dgi = DGI[row * LDG + tid];
                sgi += dgi;
                sgh += DGH[row * LDG + tid];
#pragma unroll
                for (int f = 0; f < MGV_MAX_FEAT; ++f) fx[f] = fmaf(dgi, Ms[row * LDK + D + f], fx[f]);
            }
            ACC[G_BIH + tid] += sgi;
            ACC[G_BHH + tid] += sgh;
#pragma unroll
            for (int f = 0; f < MGV_MAX_FEAT; ++f) ACC[G_WIH + tid * KX + D + f] += fx[f];
        } else {
            const int c = tid - G3;
            float s = 0.f;
            for (int row = 0; row < BTM; ++row) s = fmaf(DMs[row * LDM + c], Dg[row], s);
            ACC[G_B + c] += s;
        }
        __syncthreads();
    }
    float* part = p.partial + (((size_t)enc * gridDim.x + blockIdx.x) * 2 + p.dir) * SGRAD;
    for (int i = tid; i < SGRAD; i += THREADS) part[i] += ACC[i];
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

int struct_bwd_gx(int num_enc, int* gx_out) {
    int dev = 0, sms = 0;
    MGV_CUDA(cudaGetDevice(&dev));
    MGV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int gx = sms / (num_enc > 0 ? num_enc : 1);
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

}  // namespace

extern "C" int mgv_struct_encoder_fwd(const mgv_schedule* sch, int32_t num_enc, int32_t rounds, int32_t layernorm,
                                      int32_t feat, const float* x, const float* weights, float* states,
                                      mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_args(sch, num_enc, rounds, feat);
    if (rc != MGV_OK) return rc;
    const int N = sch->N;
    if (N == 0) return MGV_OK;
    const int steps = 2 * rounds;
    const size_t slot = (size_t)N * D;
    const size_t enc_stride = (size_t)(steps + 1) * slot;
    const size_t smem = (size_t)F_SMEM_FLOATS * sizeof(float);
    MGV_CUDA(cudaFuncSetAttribute((const void*)struct_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int e = 0; e < num_enc; ++e)
        fill_ones_kernel<<<(unsigned)((slot + 255) / 256), 256, 0, st>>>(states + e * enc_stride, slot);
        mgv_count_launches(1);
    for (int k = 1; k <= steps; ++k) {
        StepDev p{};
        const int dir = (k & 1) ? 0 : 1;
        p.N = N; p.feat = feat; p.layernorm = layernorm; p.first = (k == 1); p.last = (k == steps); p.dir = dir;
        p.ptr = dir == 0 ? sch->in_ptr : sch->out_ptr;
        p.idx = dir == 0 ? sch->in_src : sch->out_pack;
        p.x = x;
        p.weights = weights + (size_t)dir * SPACK;
        p.prev = states + (size_t)(k - 1) * slot;
        p.next = states + (size_t)k * slot;
        p.enc_stride = enc_stride;
        dim3 grid((N + FTM - 1) / FTM, num_enc);
        struct_fwd_kernel<<<grid, THREADS, smem, st>>>(p);
        mgv_count_launches(1);
    }
    return mgv_check_cuda(cudaGetLastError(), "mgv_struct_encoder_fwd");
}

extern "C" int mgv_struct_bwd_grid(void) {
    int gx = 0;
    if (struct_bwd_gx(1, &gx) != MGV_OK) return -1;
    return gx;
}

extern "C" size_t mgv_struct_bwd_workspace_bytes(int64_t N, int32_t num_enc) {
    int gx = 0;
    if (struct_bwd_gx(1, &gx) != MGV_OK) gx = 256;
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
    rc = struct_bwd_gx(num_enc, &gx);
    if (rc != MGV_OK) return rc;
    const int ntiles = (N + BTM - 1) / BTM;
    if (gx > ntiles) gx = ntiles;
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
        const int dir = (k & 1) ? 0 : 1;
        p.N = N; p.feat = feat; p.layernorm = layernorm; p.first = (k == 1); p.last = (k == steps); p.dir = dir;
        p.ptr = dir == 0 ? sch->in_ptr : sch->out_ptr;
        p.idx = dir == 0 ? sch->in_src : sch->out_pack;
        p.x = x;
        p.weights = weights + (size_t)dir * SPACK;
        p.prev = states + (size_t)(k - 1) * slot;
        p.next = nullptr;
        p.enc_stride = enc_stride;
        p.gout = gout;
        p.in_part = part[k & 1]; p.in_agg = agg[k & 1];
        p.out_part = part[(k - 1) & 1]; p.out_agg = agg[(k - 1) & 1];
        p.partial = partial;
        dim3 grid(gx, num_enc);
        struct_bwd_kernel<<<grid, THREADS, smem, st>>>(p);
        mgv_count_launches(1);
    }
    const size_t total = (size_t)num_enc * 2 * SGRAD;
    struct_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, gx, grads, num_enc);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_struct_encoder_bwd");
}
