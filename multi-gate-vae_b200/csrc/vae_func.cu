// Fused reparameterisation + KL + truth-table-similarity ("func") loss, one launch each way.
//   reparam : z = mu + exp(logstd) * eps                                   digvae_model.py:138-141
//   KL      : sum_{x in {s,t}} -0.5/N * mean_i sum_d (1 + 2 ls - mu^2 - exp(ls)^2)      trainer.py:145-148
//   func    : dis_p = 1 - cos(hf[a_p], hf[b_p]) (each norm clamped to 1e-8, torch >= 2 semantics);
//             z(x) = (x - mean) / std_unbiased ; loss = mean_p | z(dis)_p - z(tt_sim)_p |
//                                                           trainer.py:157-163, utils/utils.py:32-36
// Blocks [0, nb_vae) do the elementwise VAE part, blocks [nb_vae, grid) one pair per warp; the
// last block to finish (ticket counter) normalises and reduces the func loss.
#include "mgv_common.cuh"

namespace {
constexpr int D = MGV_D;
constexpr int THREADS = 256;
constexpr float COS_EPS = 1e-8f;

struct VfWs {               // layout of the caller-provided, zero-initialised workspace
    double sums[8];         // 0 kl_s, 1 kl_t, 2 sum dis, 3 sum dis^2, 4 sum tt, 5 sum tt^2
    unsigned ticket;
    unsigned pad[15];
    // followed by float dis[P], float na[P], float nb[P], float dot[P]
};

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < THREADS / 32) t = sh[threadIdx.x];
    if (w == 0) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;       // valid in thread 0
}

__global__ void __launch_bounds__(THREADS) vae_func_fwd_kernel(
    const float* __restrict__ mu, const float* __restrict__ ls, const float* __restrict__ eps, float* __restrict__ z,
    long long N, const float* __restrict__ hf, const long long* __restrict__ pair, const float* __restrict__ tt,
    long long P, float* __restrict__ out, VfWs* ws, int nb_vae) {
    __shared__ double sh[THREADS / 32];
    __shared__ bool is_last;
    float* dis = reinterpret_cast<float*>(ws + 1);
    float* na = dis + P;
    float* nb = na + P;
    float* dt = nb + P;
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < nb_vae) {
        // ---- reparam + KL partial sums: element e of [2][N][64]
        const long long per = N * D;
        double acc_s = 0.0, acc_t = 0.0;
        for (long long e = ((long long)blockIdx.x * THREADS + tid) * 4; e < 2 * per; e += (long long)nb_vae * THREADS * 4) {
            const float4 m4 = mgv_ldg4(mu + e), l4 = mgv_ldg4(ls + e), e4 = mgv_ldg4(eps + e);
            const float s0 = expf(l4.x), s1 = expf(l4.y), s2 = expf(l4.z), s3 = expf(l4.w);
            mgv_st4(z + e, make_float4(fmaf(s0, e4.x, m4.x), fmaf(s1, e4.y, m4.y), fmaf(s2, e4.z, m4.z), fmaf(s3, e4.w, m4.w)));
            const float k = (1.f + 2.f * l4.x - m4.x * m4.x - s0 * s0) + (1.f + 2.f * l4.y - m4.y * m4.y - s1 * s1) +
                            (1.f + 2.f * l4.z - m4.z * m4.z - s2 * s2) + (1.f + 2.f * l4.w - m4.w * m4.w - s3 * s3);
            if (e < per) acc_s += (double)k; else acc_t += (double)k;
        }
        const double bs = block_sum_d(acc_s, sh);
        const double bt = block_sum_d(acc_t, sh);
        if (tid == 0) { atomicAdd(&ws->sums[0], bs); atomicAdd(&ws->sums[1], bt); }
    } else {
        // ---- cosine distance, one pair per warp
        const int lane = tid & 31, warp = tid >> 5;
        const long long nwarps = (long long)(gridDim.x - nb_vae) * (THREADS / 32);
        double sd = 0.0, sd2 = 0.0, st = 0.0, st2 = 0.0;
        for (long long q = (long long)(blockIdx.x - nb_vae) * (THREADS / 32) + warp; q < P; q += nwarps) {
            const long long a = pair[q], b = pair[P + q];
            const float a0 = hf[a * D + lane], a1 = hf[a * D + 32 + lane];
            const float b0 = hf[b * D + lane], b1 = hf[b * D + 32 + lane];
            const float dot = mgv_warp_sum(a0 * b0 + a1 * b1);
            const float n_a = sqrtf(mgv_warp_sum(a0 * a0 + a1 * a1));
            const float n_b = sqrtf(mgv_warp_sum(b0 * b0 + b1 * b1));
            const float d = 1.0f - dot / (fmaxf(n_a, COS_EPS) * fmaxf(n_b, COS_EPS));
            if (lane == 0) {
                dis[q] = d; na[q] = n_a; nb[q] = n_b; dt[q] = dot;
                const float t = tt[q];
                sd += (double)d; sd2 += (double)d * (double)d;
                st += (double)t; st2 += (double)t * (double)t;
            }
        }
        const double r0 = block_sum_d(sd, sh), r1 = block_sum_d(sd2, sh), r2 = block_sum_d(st, sh), r3 = block_sum_d(st2, sh);
        if (tid == 0) {
            atomicAdd(&ws->sums[2], r0); atomicAdd(&ws->sums[3], r1);
            atomicAdd(&ws->sums[4], r2); atomicAdd(&ws->sums[5], r3);
        }
    }
    // ---- last block finalises
    __threadfence();
    if (tid == 0) is_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    volatile double* sums = ws->sums;
    float kl = 0.f;
    if (N > 0) {
        const double nn = (double)N;
        kl = (float)(-0.5 / nn * (sums[0] / nn) - 0.5 / nn * (sums[1] / nn));
    }
    double mean_d = 0, std_d = 1, mean_t = 0, std_t = 1;
    if (P > 0) {
        const double pp = (double)P;
        mean_d = sums[2] / pp; mean_t = sums[4] / pp;
        std_d = sqrt(fmax(sums[3] - sums[2] * sums[2] / pp, 0.0) / (pp - 1.0));
        std_t = sqrt(fmax(sums[5] - sums[4] * sums[4] / pp, 0.0) / (pp - 1.0));
    }
    // loss = mean |zd - zt| ; also the two reductions the backward needs: mean g, sum g*zd with g = sign/P
    double l1 = 0.0, sg = 0.0, sgz = 0.0;
    for (long long q = tid; q < P; q += THREADS) {
        const double zd = ((double)dis[q] - mean_d) / std_d;
        const double zt = ((double)tt[q] - mean_t) / std_t;
        const double df = zd - zt;
        l1 += fabs(df);
        const double g = (df > 0.0) ? 1.0 : ((df < 0.0) ? -1.0 : 0.0);
        sg += g; sgz += g * zd;
    }
    const double t0 = block_sum_d(l1, sh), t1 = block_sum_d(sg, sh), t2 = block_sum_d(sgz, sh);
    if (tid == 0) {
        const double pp = P > 0 ? (double)P : 1.0;
        out[0] = kl;
        out[1] = P > 0 ? (float)(t0 / pp) : 0.f;
        out[2] = (float)mean_d; out[3] = (float)std_d; out[4] = (float)mean_t; out[5] = (float)std_t;
        out[6] = (float)(t2 / pp);      // sum_p g_p zd_p with g_p = sign_p / P
        out[7] = (float)(t1 / pp / pp); // mean_p g_p
    }
}

__global__ void __launch_bounds__(THREADS) vae_func_bwd_kernel(
    const float* __restrict__ g_out, const float* __restrict__ gz, const float* __restrict__ mu,
    const float* __restrict__ ls, const float* __restrict__ eps, float* __restrict__ gmu, float* __restrict__ gls,
    long long N, const float* __restrict__ hf, const long long* __restrict__ pair, const float* __restrict__ tt,
    long long P, const float* __restrict__ out, const VfWs* ws, float* ghf, int nb_vae) {
    const float* dis = reinterpret_cast<const float*>(ws + 1);
    const float* na = dis + P;
    const float* nb = na + P;
    const float* dt = nb + P;
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < nb_vae) {
        const float gkl = g_out[0];
        const long long per = N * D;
        const float c = gkl / ((float)N * (float)N);           // d kl / d mu = mu / N^2 ; d kl / d ls = -(1 - exp(2 ls)) / N^2
        for (long long e = ((long long)blockIdx.x * THREADS + tid) * 4; e < 2 * per; e += (long long)nb_vae * THREADS * 4) {
            const float4 m4 = mgv_ldg4(mu + e), l4 = mgv_ldg4(ls + e), e4 = mgv_ldg4(eps + e), g4 = mgv_ldg4(gz + e);
            const float s0 = expf(l4.x), s1 = expf(l4.y), s2 = expf(l4.z), s3 = expf(l4.w);
            mgv_st4(gmu + e, make_float4(g4.x + c * m4.x, g4.y + c * m4.y, g4.z + c * m4.z, g4.w + c * m4.w));
            mgv_st4(gls + e, make_float4(g4.x * s0 * e4.x - c * (1.f - s0 * s0), g4.y * s1 * e4.y - c * (1.f - s1 * s1),
                                         g4.z * s2 * e4.z - c * (1.f - s2 * s2), g4.w * s3 * e4.w - c * (1.f - s3 * s3)));
        }
    } else {
        const int lane = tid & 31, warp = tid >> 5;
        const long long nwarps = (long long)(gridDim.x - nb_vae) * (THREADS / 32);
        const float gl = g_out[1];
        const float mean_d = out[2], std_d = out[3], mean_t = out[4], std_t = out[5], sgz = out[6], mg = out[7];
        const float invP = 1.0f / (float)P;
        for (long long q = (long long)(blockIdx.x - nb_vae) * (THREADS / 32) + warp; q < P; q += nwarps) {
            const float zd = (dis[q] - mean_d) / std_d;
            const float zt = (tt[q] - mean_t) / std_t;
            const float df = zd - zt;
            const float g = (df > 0.f ? invP : (df < 0.f ? -invP : 0.f));
            // d loss / d dis_q  (z-normalisation backward, unbiased std)
            const float ddis = gl * (g - mg - zd * sgz / ((float)P - 1.0f)) / std_d;
            const float dcos = -ddis;
            const long long a = pair[q], b = pair[P + q];
            const float n_a = na[q], n_b = nb[q], dot = dt[q];
            const float ca = fmaxf(n_a, COS_EPS), cb = fmaxf(n_b, COS_EPS);
            const float inv = 1.0f / (ca * cb);
            // d cos / d a = b/(ca cb) - [n_a > eps] dot a / (n_a^2 ca cb) ; same for b
            const float ka = (n_a > COS_EPS) ? dot * inv / (n_a * ca) : 0.f;
            const float kb = (n_b > COS_EPS) ? dot * inv / (n_b * cb) : 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = lane + 32 * h;
                const float av = hf[a * D + c], bv = hf[b * D + c];
                atomicAdd(ghf + a * D + c, dcos * (bv * inv - ka * av));
                atomicAdd(ghf + b * D + c, dcos * (av * inv - kb * bv));
            }
        }
    }
}

void pick_grid(long long N, long long P, int* nb_vae, int* grid) {
    long long v = N > 0 ? (2 * N * D / 4 + THREADS - 1) / THREADS : 0;
    if (v > 592) v = 592;
    long long f = P > 0 ? (P + THREADS / 32 - 1) / (THREADS / 32) : 0;
    if (f > 592) f = 592;
    if (v + f == 0) f = 1;
    *nb_vae = (int)v;
    *grid = (int)(v + f);
}
}  // namespace

extern "C" size_t mgv_vae_func_workspace_bytes(int64_t P) { return sizeof(VfWs) + (size_t)(P > 0 ? P : 0) * 4 * sizeof(float) + 256; }

extern "C" int mgv_vae_func_loss_fwd(const float* mu, const float* logstd, const float* eps, float* z, int64_t N,
                                     const float* hf, const int64_t* pair, const float* tt_sim, int64_t P,
                                     float* out, void* ws, size_t ws_bytes, mgv_stream_t stream) {
    MGV_REQUIRE(N >= 0 && P >= 0 && out && ws, "mgv_vae_func_loss_fwd: bad argument");
    MGV_REQUIRE(P != 1, "mgv_vae_func_loss_fwd: the unbiased std of one pair is undefined (reference yields NaN)");
    if (ws_bytes < mgv_vae_func_workspace_bytes(P)) {
        mgv_set_error("mgv_vae_func_loss_fwd: workspace too small");
        return MGV_ERR_WORKSPACE;
    }
    int nb_vae, grid;
    pick_grid(N, P, &nb_vae, &grid);
    vae_func_fwd_kernel<<<grid, THREADS, 0, (cudaStream_t)stream>>>(mu, logstd, eps, z, (long long)N, hf,
                                                                   (const long long*)pair, tt_sim, (long long)P, out,
                                                                   (VfWs*)ws, nb_vae);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_vae_func_loss_fwd");
}

extern "C" int mgv_vae_func_loss_bwd(const float* g_out, const float* gz, const float* mu, const float* logstd,
                                     const float* eps, float* gmu, float* glogstd, int64_t N,
                                     const float* hf, const int64_t* pair, const float* tt_sim, int64_t P,
                                     const float* out, const void* ws, float* ghf, mgv_stream_t stream) {
    MGV_REQUIRE(N >= 0 && P >= 0 && g_out && out && ws, "mgv_vae_func_loss_bwd: bad argument");
    int nb_vae, grid;
    pick_grid(N, P, &nb_vae, &grid);
    vae_func_bwd_kernel<<<grid, THREADS, 0, (cudaStream_t)stream>>>(g_out, gz, mu, logstd, eps, gmu, glogstd,
                                                                   (long long)N, hf, (const long long*)pair, tt_sim,
                                                                   (long long)P, out, (const VfWs*)ws, ghf, nb_vae);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_vae_func_loss_bwd");
}
