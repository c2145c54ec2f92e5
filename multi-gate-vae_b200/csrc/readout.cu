// Fused readout head: the probability MLP of Model.pred_prob (dg_ae_model_mig.py:44,150-152: MLP(64, 32, 1, num_layer=3,
// p_drop=0.2, batchnorm, relu), arch/mlp.py:14-56), the clamp to [0, 1] and the L1 loss of trainer.py:154-156 -- forward in ONE
// launch, backward in ONE launch (the reference runs 9 modules + clamp + loss, ~45 kernels per step with their backward).
//
//   y1 = W1 x + b1 -> BatchNorm(32) -> ReLU -> Dropout -> y2 = W2 h1 + b2 -> BatchNorm(32) -> ReLU -> Dropout -> y3 = w3 . h2 + b3
//   pred = clamp(y3, 0, 1);  loss = mean |pred - target|
//
// Layout: a warp walks nodes (groups of NB consecutive nodes, the same groups in every pass: a pass only reads scratch rows its
// own warp wrote), lane = hidden channel (32 channels = 32 lanes).  Batch statistics of a training-mode BatchNorm
// are per-lane running sums (double), combined over the grid between the three passes by a grid barrier (cooperative
// launch); the pre-normalisation activations y1 | y2 are kept in a [N][64] scratch so a pass never redoes a matrix product.
// Dropout masks come from a counter-based hash of (seed, node, layer, channel): the backward regenerates them.
#include "mgv_common.cuh"

namespace {

constexpr int DI = MGV_D, DH = 32;
constexpr int RD_THREADS = 256, RD_WARPS = RD_THREADS / 32;
constexpr int NB = 4;                      // nodes a warp stages at a time

struct RdParams {
    const float *W1, *b1, *g1, *be1, *W2, *b2, *g2, *be2, *W3, *b3;
    float *rm1, *rv1, *rm2, *rv2;          // running statistics (in: evaluation mode; in/out: training mode)
};
struct RdFwd {
    RdParams P;
    const float* x; long long N;
    int training; float p_drop; unsigned long long seed; float momentum, eps;
    const float* target;                   // [N] or null
    float* pred; float* loss;              // [N]; [1] (mean |pred - target|, when target is given)
    float* saved;                          // [N][64]: y1 | y2
    float* stats;                          // [4][32]: mean1, invstd1, mean2, invstd2 (what the backward needs)
    unsigned* mask;                        // optional [N][2] keep-bit masks of the two dropout layers (tests)
    double* acc;                           // [4][32] sums, [128] loss sum  (zeroed by the host)
    unsigned* bar;
};
struct RdBwd {
    RdParams P;
    const float* x; long long N;
    int training; float p_drop; unsigned long long seed;
    const float* target; const float* saved; const float* stats;
    const float* g_pred;                   // [N] or null: d L / d pred
    const float* g_loss;                   // device scalar or null: d L / d loss
    float* gx;                             // [N][64] out
    float* grads;                          // flat out (zeroed by the host): see mgv_b200.h
    float* dz1;                            // [N][32] scratch
    float* partial;                        // [blocks][3297] per-block parameter-gradient partial sums
    double* acc;                           // [4][32]: S2a, S2b, S1a, S1b
    unsigned* bar;
};
constexpr int G_W1 = 0, G_B1 = 2048, G_G1 = 2080, G_BE1 = 2112, G_W2 = 2144, G_B2 = 3168, G_G2 = 3200, G_BE2 = 3232, G_W3 = 3264, G_B3 = 3296;

__device__ __forceinline__ float keep_scale(unsigned long long seed, long long n, int layer, int c, float p, float sc) {
    if (p <= 0.f) return 1.0f;
    // counter-based: murmur3's 32-bit finaliser over the element index, keyed by both halves of the seed
    uint32_t h = (uint32_t)(n * 64 + layer * 32 + c) ^ (uint32_t)seed;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    h += (uint32_t)(seed >> 32) + (uint32_t)(n >> 26);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    const float u = (float)(h >> 8) * (1.0f / 16777216.0f);
    return u >= p ? sc : 0.f;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// y[c] = b + sum_k xs[k] w[k], xs in shared memory (broadcast reads), w = this lane's weight row in registers
__device__ __forceinline__ float dot64(const float* xs, const float (&w)[DI], float b) {
    float a0 = b, a1 = 0.f;
#pragma unroll
    for (int k = 0; k < DI; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(xs + k);
        a0 = fmaf(v.x, w[k], a0); a1 = fmaf(v.y, w[k + 1], a1);
        a0 = fmaf(v.z, w[k + 2], a0); a1 = fmaf(v.w, w[k + 3], a1);
    }
    return a0 + a1;
}
// y[c] = b + sum_j h[j] w[j], h[j] held by lane j
__device__ __forceinline__ float dot32(float h, const float (&w)[DH], float b) {
    float a0 = b, a1 = 0.f;
#pragma unroll
    for (int j = 0; j < DH; j += 2) {
        a0 = fmaf(__shfl_sync(0xffffffffu, h, j), w[j], a0);
        a1 = fmaf(__shfl_sync(0xffffffffu, h, j + 1), w[j + 1], a1);
    }
    return a0 + a1;
}

__global__ void __launch_bounds__(RD_THREADS, 2) readout_fwd_kernel(const RdFwd p) {
    __shared__ __align__(16) float xs[RD_WARPS][NB][DI];
    __shared__ double red[RD_WARPS][2][DH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * RD_WARPS + warp, nw = (long long)gridDim.x * RD_WARPS;
    const RdParams& P = p.P;
    const float sc = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
    const bool train = p.training != 0;
    float w[DI];
#pragma unroll
    for (int k = 0; k < DI; ++k) w[k] = __ldg(P.W1 + lane * DI + k);
    const float b1 = __ldg(P.b1 + lane);
    // ------------------------------------------------------------------ pass 1: y1 (+ batch statistics)
    double s = 0.0, q = 0.0;
    for (long long n0 = gw * NB; n0 < p.N; n0 += nw * NB) {
#pragma unroll
        for (int i = 0; i < NB; ++i)
            if (n0 + i < p.N) *reinterpret_cast<float2*>(&xs[warp][i][2 * lane]) = __ldg(reinterpret_cast<const float2*>(p.x + (n0 + i) * DI) + lane);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if (n0 + i < p.N) {
                const float y = dot64(xs[warp][i], w, b1);
                p.saved[(n0 + i) * 64 + lane] = y;
                s += (double)y; q += (double)y * (double)y;
            }
        }
        __syncwarp();
    }
    float mean1, is1;
    if (train) {
        red[warp][0][lane] = s; red[warp][1][lane] = q;
        __syncthreads();
        if (warp == 0) {
            double a = 0.0, b = 0.0;
            for (int v = 0; v < RD_WARPS; ++v) { a += red[v][0][lane]; b += red[v][1][lane]; }
            atomicAdd(p.acc + lane, a); atomicAdd(p.acc + DH + lane, b);
        }
        mgv_grid_sync(p.bar, gridDim.x);
        const double m = __ldcg(p.acc + lane) / (double)p.N;
        const double var = fmax(__ldcg(p.acc + DH + lane) / (double)p.N - m * m, 0.0);
        mean1 = (float)m; is1 = (float)(1.0 / sqrt(var + (double)p.eps));
        if (blockIdx.x == 0 && warp == 0) {
            p.stats[lane] = mean1; p.stats[DH + lane] = is1;
            const double unb = p.N > 1 ? var * (double)p.N / (double)(p.N - 1) : var;
            P.rm1[lane] = (1.0f - p.momentum) * P.rm1[lane] + p.momentum * mean1;
            P.rv1[lane] = (1.0f - p.momentum) * P.rv1[lane] + p.momentum * (float)unb;
        }
    } else {
        mean1 = __ldg(P.rm1 + lane); is1 = rsqrtf(__ldg(P.rv1 + lane) + p.eps);
        if (blockIdx.x == 0 && warp == 0) { p.stats[lane] = mean1; p.stats[DH + lane] = is1; }
    }
    // ------------------------------------------------------------------ pass 2: h1 -> y2 (+ batch statistics)
    float w2[DH];
#pragma unroll
    for (int j = 0; j < DH; ++j) w2[j] = __ldg(P.W2 + lane * DH + j);
    const float g1 = __ldg(P.g1 + lane) * is1, o1 = __ldg(P.be1 + lane) - mean1 * g1, b2 = __ldg(P.b2 + lane);
    s = 0.0; q = 0.0;
    for (long long n0 = gw * NB; n0 < p.N; n0 += nw * NB) {        // four nodes per trip: their loads and hashes overlap
        float y1[NB], k1[NB], y2[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) y1[i] = (n0 + i < p.N) ? p.saved[(n0 + i) * 64 + lane] : 0.f;
#pragma unroll
        for (int i = 0; i < NB; ++i) k1[i] = train ? keep_scale(p.seed, n0 + i, 0, lane, p.p_drop, sc) : 1.0f;
#pragma unroll
        for (int i = 0; i < NB; ++i) y2[i] = dot32(fmaxf(fmaf(y1[i], g1, o1), 0.f) * k1[i], w2, b2);
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if (n0 + i < p.N) {
                p.saved[(n0 + i) * 64 + DH + lane] = y2[i];
                s += (double)y2[i]; q += (double)y2[i] * (double)y2[i];
            }
            if (p.mask) {
                const unsigned bits = __ballot_sync(0xffffffffu, k1[i] != 0.f);
                if (lane == 0 && n0 + i < p.N) p.mask[(n0 + i) * 2] = bits;
            }
        }
    }
    float mean2, is2;
    if (train) {
        __syncthreads();
        red[warp][0][lane] = s; red[warp][1][lane] = q;
        __syncthreads();
        if (warp == 0) {
            double a = 0.0, b = 0.0;
            for (int v = 0; v < RD_WARPS; ++v) { a += red[v][0][lane]; b += red[v][1][lane]; }
            atomicAdd(p.acc + 2 * DH + lane, a); atomicAdd(p.acc + 3 * DH + lane, b);
        }
        mgv_grid_sync(p.bar, gridDim.x);
        const double m = __ldcg(p.acc + 2 * DH + lane) / (double)p.N;
        const double var = fmax(__ldcg(p.acc + 3 * DH + lane) / (double)p.N - m * m, 0.0);
        mean2 = (float)m; is2 = (float)(1.0 / sqrt(var + (double)p.eps));
        if (blockIdx.x == 0 && warp == 0) {
            p.stats[2 * DH + lane] = mean2; p.stats[3 * DH + lane] = is2;
            const double unb = p.N > 1 ? var * (double)p.N / (double)(p.N - 1) : var;
            P.rm2[lane] = (1.0f - p.momentum) * P.rm2[lane] + p.momentum * mean2;
            P.rv2[lane] = (1.0f - p.momentum) * P.rv2[lane] + p.momentum * (float)unb;
        }
    } else {
        mean2 = __ldg(P.rm2 + lane); is2 = rsqrtf(__ldg(P.rv2 + lane) + p.eps);
        if (blockIdx.x == 0 && warp == 0) { p.stats[2 * DH + lane] = mean2; p.stats[3 * DH + lane] = is2; }
    }
    // ------------------------------------------------------------------ pass 3: h2 -> y3 -> clamp -> |pred - target|
    const float g2 = __ldg(P.g2 + lane) * is2, o2 = __ldg(P.be2 + lane) - mean2 * g2, w3 = __ldg(P.W3 + lane), b3 = __ldg(P.b3);
    double lsum = 0.0;
    for (long long n0 = gw * NB; n0 < p.N; n0 += nw * NB) {
        float y2[NB], k2[NB], y3[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) y2[i] = (n0 + i < p.N) ? p.saved[(n0 + i) * 64 + DH + lane] : 0.f;
#pragma unroll
        for (int i = 0; i < NB; ++i) k2[i] = train ? keep_scale(p.seed, n0 + i, 1, lane, p.p_drop, sc) : 1.0f;
#pragma unroll
        for (int i = 0; i < NB; ++i) y3[i] = fmaxf(fmaf(y2[i], g2, o2), 0.f) * k2[i] * w3;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < NB; ++i) y3[i] += __shfl_xor_sync(0xffffffffu, y3[i], o);
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const long long n = n0 + i;
            const float pr = fminf(fmaxf(y3[i] + b3, 0.f), 1.f);
            if (p.mask) {
                const unsigned bits = __ballot_sync(0xffffffffu, k2[i] != 0.f);
                if (lane == 0 && n < p.N) p.mask[n * 2 + 1] = bits;
            }
            if (lane == 0 && n < p.N) {
                p.pred[n] = pr;
                if (p.target) lsum += (double)fabsf(pr - __ldg(p.target + n));
            }
        }
    }
    if (p.target) {
        __syncthreads();
        if (lane == 0) red[warp][0][0] = lsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0;
            for (int v = 0; v < RD_WARPS; ++v) a += red[v][0][0];
            atomicAdd(p.acc + 4 * DH, a);
        }
        mgv_grid_sync(p.bar, gridDim.x);
        if (blockIdx.x == 0 && threadIdx.x == 0) p.loss[0] = (float)(__ldcg(p.acc + 4 * DH) / (double)(p.N > 0 ? p.N : 1));
    }
}

// Sum of every warp's per-lane accumulators v[NV] over the RD_WARPS warps of the block, as a tree through `scratch`
// ([RD_WARPS / 2][NV][32] floats): three rounds of (upper half writes, lower half adds) instead of RD_WARPS serial turns or
// shared-memory float atomics (CAS loops).  Afterwards warp 0 holds the block sums.  Deterministic.
template <int NV>
__device__ __forceinline__ void block_tree_add(float (&v)[NV], float* scratch, int warp, int lane) {
#pragma unroll 1
    for (int half = RD_WARPS / 2; half >= 1; half >>= 1) {
        if (warp >= half && warp < 2 * half) {
#pragma unroll
            for (int i = 0; i < NV; ++i) scratch[((warp - half) * NV + i) * 32 + lane] = v[i];
        }
        __syncthreads();
        if (warp < half) {
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] += scratch[(warp * NV + i) * 32 + lane];
        }
        __syncthreads();
    }
}
constexpr int RD_SCRATCH_BYTES = (RD_WARPS / 2) * 65 * 32 * 4;          // pass C: d W1 row (64) + d b1 per lane

__global__ void __launch_bounds__(RD_THREADS, 2) readout_bwd_kernel(const RdBwd p) {
    __shared__ __align__(16) float xs[RD_WARPS][NB][DI];
    __shared__ __align__(16) float W1s[DH][DI];
    __shared__ double red[RD_WARPS][2][DH];
    extern __shared__ __align__(16) float scratch[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * RD_WARPS + warp, nw = (long long)gridDim.x * RD_WARPS;
    const RdParams& P = p.P;
    const bool train = p.training != 0;
    const float sc = (train && p.p_drop > 0.f) ? 1.0f / (1.0f - p.p_drop) : 1.0f;
    const float pd = train ? p.p_drop : 0.f;
    float* part = p.partial + (size_t)blockIdx.x * 3297;        // this block's parameter-gradient partial sums
    for (int i = threadIdx.x; i < DH * DI; i += RD_THREADS) W1s[i / DI][i % DI] = __ldg(P.W1 + i);
    __syncthreads();
    const float mean1 = p.stats[lane], is1 = p.stats[DH + lane], mean2 = p.stats[2 * DH + lane], is2 = p.stats[3 * DH + lane];
    const float g1 = __ldg(P.g1 + lane), be1 = __ldg(P.be1 + lane), g2 = __ldg(P.g2 + lane), be2 = __ldg(P.be2 + lane);
    const float w3 = __ldg(P.W3 + lane), b3 = __ldg(P.b3);
    const float gl = p.g_loss ? __ldg(p.g_loss) / (float)(p.N > 0 ? p.N : 1) : 0.f;
    const double invN = 1.0 / (double)(p.N > 0 ? p.N : 1);

    // d z2 of the NB nodes n0 .. (the gradient at the second BatchNorm's output, after ReLU / dropout), with xhat2, h2 and d pred;
    // the nodes' loads, hashes and reductions are interleaved (independent dependency chains)
    auto dz2_group = [&](long long n0, float (&xh2)[NB], float (&h2)[NB], float (&dp)[NB], float (&dz2)[NB]) {
        float y2[NB], k2[NB], d[NB], tg[NB], a2[NB], y3[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const long long n = n0 + i;
            const bool ok = n < p.N;
            y2[i] = ok ? p.saved[n * 64 + DH + lane] : 0.f;
            d[i] = (ok && p.g_pred) ? __ldg(p.g_pred + n) : 0.f;
            tg[i] = (ok && p.target) ? __ldg(p.target + n) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) k2[i] = keep_scale(p.seed, n0 + i, 1, lane, pd, sc);
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            xh2[i] = (y2[i] - mean2) * is2;
            a2[i] = fmaf(xh2[i], g2, be2);
            h2[i] = fmaxf(a2[i], 0.f) * k2[i];
            y3[i] = h2[i] * w3;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < NB; ++i) y3[i] += __shfl_xor_sync(0xffffffffu, y3[i], o);
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const float y = y3[i] + b3;
            const float pr = fminf(fmaxf(y, 0.f), 1.f);
            float dd = d[i];
            if (p.target) {
                const float diff = pr - tg[i];
                dd += gl * (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f));
            }
            dp[i] = (n0 + i < p.N && y >= 0.f && y <= 1.f) ? dd : 0.f;
            dz2[i] = a2[i] > 0.f ? dp[i] * w3 * k2[i] : 0.f;
        }
    };
    // ------------------------------------------------------------------ pass A: d W3, d b3, sums of the second BatchNorm
    {
        double sa = 0.0, sb = 0.0;
        float dw3 = 0.f, db3 = 0.f;
        for (long long n0 = gw * NB; n0 < p.N; n0 += nw * NB) {
            float xh2[NB], h2[NB], dp[NB], dz2[NB];
            dz2_group(n0, xh2, h2, dp, dz2);
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                dw3 = fmaf(dp[i], h2[i], dw3); db3 += dp[i];
                sa += (double)dz2[i]; sb += (double)dz2[i] * (double)xh2[i];
            }
        }
        {
            float v[4] = {dw3, db3, (float)sa, (float)sb};
            block_tree_add<4>(v, scratch, warp, lane);
            if (warp == 0) {
                part[G_W3 + lane] = v[0];
                if (lane == 0) part[G_B3] = v[1];              // d pred is warp-uniform: every lane holds the same sum
                part[G_BE2 + lane] = v[2]; part[G_G2 + lane] = v[3];
            }
        }
        red[warp][0][lane] = sa; red[warp][1][lane] = sb;
        __syncthreads();
        if (warp == 0) {
            double a = 0.0, b = 0.0;
            for (int v = 0; v < RD_WARPS; ++v) { a += red[v][0][lane]; b += red[v][1][lane]; }
            atomicAdd(p.acc + lane, a); atomicAdd(p.acc + DH + lane, b);
        }
        mgv_grid_sync(p.bar, gridDim.x);
    }
    // ------------------------------------------------------------------ pass B: through the second Linear, sums of the first BatchNorm
    {
        const float s2a = train ? (float)(__ldcg(p.acc + lane) * invN) : 0.f, s2b = train ? (float)(__ldcg(p.acc + DH + lane) * invN) : 0.f;
        float w2t[DH], dw2[DH];
#pragma unroll
        for (int c = 0; c < DH; ++c) { w2t[c] = __ldg(P.W2 + c * DH + lane); dw2[c] = 0.f; }
        float db2 = 0.f;
        double sa = 0.0, sb = 0.0;
        for (long long n0 = gw * NB; n0 < p.N; n0 += nw * NB) {
            float xh2[NB], h2[NB], dp[NB], dz2[NB];
            float y1[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) y1[i] = (n0 + i < p.N) ? p.saved[(n0 + i) * 64 + lane] : 0.f;
            dz2_group(n0, xh2, h2, dp, dz2);
            float dy2[NB], xh1[NB], a1[NB], k1[NB], h1[NB], dh1[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                dy2[i] = (n0 + i < p.N) ? g2 * is2 * (dz2[i] - s2a - xh2[i] * s2b) : 0.f;
                db2 += dy2[i];
                xh1[i] = (y1[i] - mean1) * is1;
                a1[i] = fmaf(xh1[i], g1, be1);
                k1[i] = keep_scale(p.seed, n0 + i, 0, lane, pd, sc);
                h1[i] = fmaxf(a1[i], 0.f) * k1[i];
                dh1[i] = 0.f;
            }
#pragma unroll
            for (int c = 0; c < DH; ++c) {
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const float dyc = __shfl_sync(0xffffffffu, dy2[i], c);
                    dh1[i] = fmaf(dyc, w2t[c], dh1[i]);            // d h1[lane] = sum_c d y2[c] W2[c][lane]
                    dw2[c] = fmaf(dyc, h1[i], dw2[c]);             // d W2[c][lane] += d y2[c] h1[lane]
                }
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                if (n0 + i < p.N) {
                    const float dz1 = a1[i] > 0.f ? dh1[i] * k1[i] : 0.f;
                    p.dz1[(n0 + i) * DH + lane] = dz1;
                    sa += (double)dz1; sb += (double)dz1 * (double)xh1[i];
                }
            }
        }
        {
            float v[DH + 3];
#pragma unroll
            for (int c = 0; c < DH; ++c) v[c] = dw2[c];
            v[DH] = db2; v[DH + 1] = (float)sa; v[DH + 2] = (float)sb;
            block_tree_add<DH + 3>(v, scratch, warp, lane);
            if (warp == 0) {
#pragma unroll
                for (int c = 0; c < DH; ++c) part[G_W2 + c * DH + lane] = v[c];
                part[G_B2 + lane] = v[DH]; part[G_BE1 + lane] = v[DH + 1]; part[G_G1 + lane] = v[DH + 2];
            }
        }
        red[warp][0][lane] = sa; red[warp][1][lane] = sb;
        __syncthreads();
        if (warp == 0) {
            double a = 0.0, b = 0.0;
            for (int v = 0; v < RD_WARPS; ++v) { a += red[v][0][lane]; b += red[v][1][lane]; }
            atomicAdd(p.acc + 2 * DH + lane, a); atomicAdd(p.acc + 3 * DH + lane, b);
        }
        mgv_grid_sync(p.bar, gridDim.x);
    }
    // ------------------------------------------------------------------ pass C: through the first Linear -> d x, d W1, d b1
    {
        const float s1a = train ? (float)(__ldcg(p.acc + 2 * DH + lane) * invN) : 0.f, s1b = train ? (float)(__ldcg(p.acc + 3 * DH + lane) * invN) : 0.f;
        float dw1[DI];
#pragma unroll
        for (int k = 0; k < DI; ++k) dw1[k] = 0.f;
        float db1 = 0.f;
        for (long long n0 = gw * NB; n0 < p.N; n0 += nw * NB) {
#pragma unroll
            for (int i = 0; i < NB; ++i)
                if (n0 + i < p.N) *reinterpret_cast<float2*>(&xs[warp][i][2 * lane]) = __ldg(reinterpret_cast<const float2*>(p.x + (n0 + i) * DI) + lane);
            __syncwarp();
#pragma unroll 1
            for (int i = 0; i < NB; ++i) {
                const long long n = n0 + i;
                if (n >= p.N) break;
                const float xh1 = (p.saved[n * 64 + lane] - mean1) * is1;
                const float dy1 = g1 * is1 * (p.dz1[n * DH + lane] - s1a - xh1 * s1b);
                db1 += dy1;
#pragma unroll
                for (int k = 0; k < DI; k += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(&xs[warp][i][k]);
                    dw1[k] = fmaf(dy1, v.x, dw1[k]); dw1[k + 1] = fmaf(dy1, v.y, dw1[k + 1]);
                    dw1[k + 2] = fmaf(dy1, v.z, dw1[k + 2]); dw1[k + 3] = fmaf(dy1, v.w, dw1[k + 3]);
                }
                float2 dx = make_float2(0.f, 0.f);             // d x[2 lane], d x[2 lane + 1] = sum_c d y1[c] W1[c][.]
#pragma unroll
                for (int c = 0; c < DH; ++c) {
                    const float dyc = __shfl_sync(0xffffffffu, dy1, c);
                    const float2 wv = *reinterpret_cast<const float2*>(&W1s[c][2 * lane]);
                    dx.x = fmaf(dyc, wv.x, dx.x); dx.y = fmaf(dyc, wv.y, dx.y);
                }
                *reinterpret_cast<float2*>(p.gx + n * DI + 2 * lane) = dx;
            }
            __syncwarp();
        }
        {
            float v[DI + 1];
#pragma unroll
            for (int k = 0; k < DI; ++k) v[k] = dw1[k];
            v[DI] = db1;
            block_tree_add<DI + 1>(v, scratch, warp, lane);
            if (warp == 0) {
#pragma unroll
                for (int k = 0; k < DI; ++k) part[G_W1 + lane * DI + k] = v[k];
                part[G_B1 + lane] = v[DI];
            }
        }
    }
    // per-block partial sums are in global memory (no atomics): every gradient element is summed over the blocks
    mgv_grid_sync(p.bar, gridDim.x);
    for (int i = blockIdx.x * RD_THREADS + threadIdx.x; i < 3297; i += gridDim.x * RD_THREADS) {
        float t = 0.f;
        for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(p.partial + (size_t)b * 3297 + i);
        p.grads[i] = t;
    }
}

int fill_params(RdParams& P, const void* const* params) {
    MGV_REQUIRE(params != nullptr, "readout: null parameter table");
    for (int i = 0; i < 14; ++i) MGV_REQUIRE(params[i] != nullptr, "readout: parameter %d is null", i);
    P.W1 = (const float*)params[0]; P.b1 = (const float*)params[1]; P.g1 = (const float*)params[2]; P.be1 = (const float*)params[3];
    P.rm1 = (float*)params[4]; P.rv1 = (float*)params[5];
    P.W2 = (const float*)params[6]; P.b2 = (const float*)params[7]; P.g2 = (const float*)params[8]; P.be2 = (const float*)params[9];
    P.rm2 = (float*)params[10]; P.rv2 = (float*)params[11];
    P.W3 = (const float*)params[12]; P.b3 = (const float*)params[13];
    return MGV_OK;
}

int coop_blocks(const void* kern, long long N, int* out, size_t dyn_smem = 0) {
    int dev = 0, sms = 0, occ = 0;
    MGV_CUDA(cudaGetDevice(&dev));
    MGV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (dyn_smem) MGV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    MGV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RD_THREADS, dyn_smem));
    MGV_REQUIRE(occ >= 1, "readout: kernel does not fit on an SM");
    long long want = (N + RD_WARPS * NB - 1) / (RD_WARPS * NB);
    const long long cap = (long long)sms * (occ > 2 ? 2 : occ);
    *out = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    return MGV_OK;
}

}  // namespace

constexpr int RD_MAX_BLOCKS = 1024;        // upper bound of the cooperative grid (2 blocks per SM)
extern "C" size_t mgv_readout_workspace_bytes(int64_t N) {
    return mgv_align_up((size_t)(N > 0 ? N : 1) * DH * 4 + 256, 256) + mgv_align_up(192 * sizeof(double), 256) +
           mgv_align_up((size_t)RD_MAX_BLOCKS * 3297 * 4 + 256, 256) + 1024;
}

extern "C" int mgv_readout_fwd(const float* x, int64_t N, const void* const* params, int32_t training, float p_drop, uint64_t seed,
                               float momentum, float eps, const float* target, float* pred, float* loss, float* saved, float* stats,
                               uint32_t* mask, void* ws, size_t ws_bytes, int32_t* sync, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MGV_REQUIRE(N >= 0 && pred && saved && stats && sync && (N == 0 || x), "mgv_readout_fwd: bad argument");
    MGV_REQUIRE(!target || loss, "mgv_readout_fwd: a target needs the loss output");
    MGV_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "mgv_readout_fwd: p_drop must be in [0, 1)");
    if (ws_bytes < mgv_readout_workspace_bytes(N)) { mgv_set_error("mgv_readout_fwd: workspace too small"); return MGV_ERR_WORKSPACE; }
    RdFwd p{};
    int rc = fill_params(p.P, params);
    if (rc != MGV_OK) return rc;
    MgvArena a(ws, ws_bytes);
    (void)a.take<float>((size_t)(N > 0 ? N : 1) * DH);
    p.acc = a.take<double>(192);
    p.x = x; p.N = N; p.training = training; p.p_drop = training ? p_drop : 0.f; p.seed = seed; p.momentum = momentum; p.eps = eps;
    p.target = target; p.pred = pred; p.loss = loss; p.saved = saved; p.stats = stats; p.mask = mask;
    p.bar = reinterpret_cast<unsigned*>(sync);
    MGV_CUDA(cudaMemsetAsync(p.acc, 0, 192 * sizeof(double), st));
    MGV_CUDA(cudaMemsetAsync(sync, 0, sizeof(int32_t), st));
    if (N == 0) {
        if (loss) MGV_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
        return MGV_OK;
    }
    int blocks = 0;
    rc = coop_blocks((const void*)readout_fwd_kernel, N, &blocks);
    if (rc != MGV_OK) return rc;
    void* args[] = {&p};
    MGV_CUDA(cudaLaunchCooperativeKernel((const void*)readout_fwd_kernel, dim3(blocks), dim3(RD_THREADS), args, 0, st));
    mgv_count_launches(1);
    return MGV_OK;
}

extern "C" int mgv_readout_bwd(const float* x, int64_t N, const void* const* params, int32_t training, float p_drop, uint64_t seed,
                               const float* target, const float* saved, const float* stats, const float* g_pred, const float* g_loss,
                               float* gx, float* grads, void* ws, size_t ws_bytes, int32_t* sync, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MGV_REQUIRE(N >= 0 && saved && stats && gx && grads && sync && (N == 0 || x), "mgv_readout_bwd: bad argument");
    MGV_REQUIRE(!g_loss || target, "mgv_readout_bwd: a loss gradient needs the target");
    if (ws_bytes < mgv_readout_workspace_bytes(N)) { mgv_set_error("mgv_readout_bwd: workspace too small"); return MGV_ERR_WORKSPACE; }
    RdBwd p{};
    int rc = fill_params(p.P, params);
    if (rc != MGV_OK) return rc;
    MgvArena a(ws, ws_bytes);
    p.dz1 = a.take<float>((size_t)(N > 0 ? N : 1) * DH);
    p.acc = a.take<double>(192);
    p.partial = a.take<float>((size_t)RD_MAX_BLOCKS * 3297);
    p.x = x; p.N = N; p.training = training; p.p_drop = p_drop; p.seed = seed;
    p.target = g_loss ? target : nullptr; p.saved = saved; p.stats = stats; p.g_pred = g_pred; p.g_loss = g_loss;
    p.gx = gx; p.grads = grads; p.bar = reinterpret_cast<unsigned*>(sync);
    MGV_CUDA(cudaMemsetAsync(p.acc, 0, 192 * sizeof(double), st));
    MGV_CUDA(cudaMemsetAsync(sync, 0, sizeof(int32_t), st));
    if (N == 0) {
        MGV_CUDA(cudaMemsetAsync(grads, 0, 3297 * sizeof(float), st));
        return MGV_OK;
    }
    int blocks = 0;
    rc = coop_blocks((const void*)readout_bwd_kernel, N, &blocks, (size_t)RD_SCRATCH_BYTES);
    if (rc != MGV_OK) return rc;
    MGV_REQUIRE(blocks <= RD_MAX_BLOCKS, "mgv_readout_bwd: grid of %d blocks exceeds the partial buffer", blocks);
    void* args[] = {&p};
    MGV_CUDA(cudaLaunchCooperativeKernel((const void*)readout_bwd_kernel, dim3(blocks), dim3(RD_THREADS), args, (size_t)RD_SCRATCH_BYTES, st));
    mgv_count_launches(1);
    return MGV_OK;
}
