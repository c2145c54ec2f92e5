// Reconstruction loss of Model.recon_loss (dg_ae_model_mig.py:169-191) with the directed inner-product decoder
// (digae_layer.py:26-33), fused, plus a device-side negative sampler standing in for
// torch_geometric.utils.negative_sampling at dg_ae_model_mig.py:177-180:
//   st = hs_decompose(hs) = [s | t]  ([N][128]);  value(u -> v) = sigmoid(s_u . t_v)
//   loss = -mean_pos log(value + EPS) - mean_neg log(1 - value + EPS),   EPS = 1e-15 (dg_ae_model_mig.py:18)
// One half-warp per edge: two coalesced 256-byte row reads, shuffle reduction of the dot product.
#include "mgv_common.cuh"

namespace {

constexpr int D = MGV_D;
constexpr int D2 = 2 * D;
constexpr float EPS = 1e-15f;
constexpr int NODE_MASK = (1 << MGV_CODE_SHIFT) - 1;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {            // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// neg[0][e], neg[1][e]: a uniformly drawn ordered pair (u, v), u != v, that is not an edge u -> v (rejection, <= 16 tries).
__global__ void neg_sample_kernel(const int* __restrict__ out_ptr, const int* __restrict__ out_pack, int n, int64_t count,
                                  uint64_t seed, int64_t* __restrict__ neg) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    int u = 0, v = 0;
    for (int attempt = 0; attempt < 16; ++attempt) {
        const uint64_t h = mix64(seed ^ mix64((uint64_t)e * 16 + attempt));
        u = (int)((h & 0xffffffffull) % (uint64_t)n);
        v = (int)((h >> 32) % (uint64_t)n);
        if (u == v) continue;
        bool hit = false;
        for (int q = out_ptr[u]; q < out_ptr[u + 1]; ++q) hit |= ((out_pack[q] & NODE_MASK) == v);
        if (!hit) break;
    }
    neg[e] = u;
    neg[count + e] = v;
}

// Keyed pseudo-random bijection of [0, 2^(2 hb)): four Feistel rounds with the splitmix finaliser as round function.
__device__ __forceinline__ uint64_t feistel4(uint64_t x, int hb, uint64_t seed) {
    const uint64_t mask = (1ull << hb) - 1ull;
    uint64_t l = x >> hb, r = x & mask;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
        const uint64_t t = l ^ (mix64(r ^ seed ^ ((uint64_t)round << 56)) & mask);
        l = r;
        r = t;
    }
    return (l << hb) | r;
}

// out[:, i] = ei[:, pi(i)], pi = the Feistel bijection of the smallest even-width power-of-two domain >= E, cycle-walked back
// into [0, E) (a walk that starts inside [0, E) returns to it: pi restricted this way is a bijection of [0, E); the domain is
// < 4 E, so a walk takes < 4 steps on average).
__global__ void permute_edges_kernel(const int64_t* __restrict__ ei, int64_t E, int hb, uint64_t seed, int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    uint64_t x = (uint64_t)i;
    do x = feistel4(x, hb, seed); while (x >= (uint64_t)E);
    out[i] = ei[x];
    out[E + i] = ei[E + x];
}

struct ReconDev {
    const float* st;
    const int64_t* pos; int64_t ep;
    const int64_t* neg; int64_t en;
    int n;
    float* sig; int32_t* pred; double* sums;      // forward outputs
    const float* gl; float* gst;                  // backward
    int* err;
};

__device__ __forceinline__ bool edge_of(const ReconDev& p, int64_t e, int& u, int& v, bool& is_pos) {
    is_pos = e < p.ep;
    const int64_t* ei = is_pos ? p.pos : p.neg;
    const int64_t cnt = is_pos ? p.ep : p.en, k = is_pos ? e : e - p.ep;
    const int64_t uu = ei[k], vv = ei[cnt + k];
    u = (int)uu; v = (int)vv;
    return uu >= 0 && uu < p.n && vv >= 0 && vv < p.n;
}

__global__ void __launch_bounds__(256) recon_fwd_kernel(const ReconDev p) {
    __shared__ double s_pos[8], s_neg[8];
    const int lane = threadIdx.x & 31, l16 = lane & 15, warp = threadIdx.x >> 5;
    const int64_t total = p.ep + p.en;
    const int64_t hw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4, nhw = ((int64_t)gridDim.x * blockDim.x) >> 4;
    double acc_pos = 0.0, acc_neg = 0.0;
    const unsigned hmask = (lane & 16) ? 0xffff0000u : 0x0000ffffu;       // the two half-warps leave the loop independently
    for (int64_t e = hw; e < total; e += nhw) {
        int u, v; bool is_pos;
        float d = 0.f;
        const bool ok = edge_of(p, e, u, v, is_pos);
        if (ok) d = mgv_dot4(mgv_ld4(p.st + (size_t)u * D2 + 4 * l16), mgv_ld4(p.st + (size_t)v * D2 + D + 4 * l16));
        else if (l16 == 0 && p.err) atomicOr(p.err, 4);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(hmask, d, o);
        if (l16 == 0) {
            const float s = 1.0f / (1.0f + expf(-d));
            p.sig[e] = s;
            p.pred[e] = s > 0.5f ? 1 : 0;
            if (is_pos) acc_pos += (double)(-logf(s + EPS));
            else acc_neg += (double)(-logf(1.0f - s + EPS));
        }
    }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_pos += __shfl_xor_sync(0xffffffffu, acc_pos, o);
        acc_neg += __shfl_xor_sync(0xffffffffu, acc_neg, o);
    }
    if (lane == 0) { s_pos[warp] = acc_pos; s_neg[warp] = acc_neg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += s_pos[w]; b += s_neg[w]; }
        atomicAdd(p.sums, a);
        atomicAdd(p.sums + 1, b);
    }
}

__global__ void recon_finish_kernel(const double* sums, int64_t ep, int64_t en, float* out) {
    const double a = ep > 0 ? sums[0] / (double)ep : 0.0, b = en > 0 ? sums[1] / (double)en : 0.0;
    out[0] = (float)(a + b);
    out[1] = (float)a;
    out[2] = (float)b;
}

__device__ __forceinline__ void red_add4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(256) recon_bwd_kernel(const ReconDev p) {
    const int l16 = threadIdx.x & 15;
    const int64_t total = p.ep + p.en;
    const int64_t hw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4, nhw = ((int64_t)gridDim.x * blockDim.x) >> 4;
    const float gl = *p.gl;
    for (int64_t e = hw; e < total; e += nhw) {
        int u, v; bool is_pos;
        if (!edge_of(p, e, u, v, is_pos)) continue;
        const float s = p.sig[e];
        const float ds = s * (1.0f - s);
        const float coef = is_pos ? -gl * ds / (s + EPS) / (float)p.ep : gl * ds / (1.0f - s + EPS) / (float)p.en;
        const float4 s4 = mgv_ld4(p.st + (size_t)u * D2 + 4 * l16), t4 = mgv_ld4(p.st + (size_t)v * D2 + D + 4 * l16);
        red_add4(p.gst + (size_t)u * D2 + 4 * l16, make_float4(coef * t4.x, coef * t4.y, coef * t4.z, coef * t4.w));
        red_add4(p.gst + (size_t)v * D2 + D + 4 * l16, make_float4(coef * s4.x, coef * s4.y, coef * s4.z, coef * s4.w));
    }
}

int grid_for(int64_t half_warps) {
    int64_t blocks = (half_warps * 16 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    return blocks < 1 ? 1 : (int)blocks;
}

}  // namespace

extern "C" int mgv_negative_sample(const int32_t* out_ptr, const int32_t* out_pack, int32_t N, int64_t count, uint64_t seed,
                                   int64_t* neg, mgv_stream_t stream) {
    MGV_REQUIRE(out_ptr && out_pack && neg && N >= 2 && count >= 0, "mgv_negative_sample: bad argument (needs N >= 2)");
    if (count == 0) return MGV_OK;
    neg_sample_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out_ptr, out_pack, N, count, seed, neg);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_negative_sample");
}

extern "C" int mgv_permute_edges(const int64_t* edge_index, int64_t E, uint64_t seed, int64_t* out, mgv_stream_t stream) {
    MGV_REQUIRE(E >= 0 && (E == 0 || (edge_index && out && edge_index != out)), "mgv_permute_edges: bad argument (in place is not supported)");
    if (E == 0) return MGV_OK;
    int hb = 1;
    while ((1ull << (2 * hb)) < (uint64_t)E) ++hb;
    permute_edges_kernel<<<(unsigned)((E + 255) / 256), 256, 0, (cudaStream_t)stream>>>(edge_index, E, hb, seed, out);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_permute_edges");
}

extern "C" int mgv_recon_loss_fwd(const float* st, int32_t N, const int64_t* pos, int64_t Ep, const int64_t* neg, int64_t En,
                                  float* out, float* sig, int32_t* pred, void* ws, size_t ws_bytes, int32_t* err_flag,
                                  mgv_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    MGV_REQUIRE(st && out && sig && pred && ws && ws_bytes >= 16 && Ep >= 0 && En >= 0, "mgv_recon_loss_fwd: bad argument");
    ReconDev p{};
    p.st = st; p.pos = pos; p.ep = Ep; p.neg = neg; p.en = En; p.n = N; p.sig = sig; p.pred = pred;
    p.sums = reinterpret_cast<double*>(ws); p.err = err_flag;
    MGV_CUDA(cudaMemsetAsync(ws, 0, 16, s));
    if (Ep + En > 0) {
        recon_fwd_kernel<<<grid_for(Ep + En), 256, 0, s>>>(p);
        mgv_count_launches(1);
    }
    recon_finish_kernel<<<1, 1, 0, s>>>(p.sums, Ep, En, out);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_recon_loss_fwd");
}

extern "C" int mgv_recon_loss_bwd(const float* st, int32_t N, const int64_t* pos, int64_t Ep, const int64_t* neg, int64_t En,
                                  const float* sig, const float* g_loss, float* gst, mgv_stream_t stream) {
    MGV_REQUIRE(st && sig && g_loss && gst, "mgv_recon_loss_bwd: bad argument");
    if (Ep + En == 0) return MGV_OK;
    ReconDev p{};
    p.st = st; p.pos = pos; p.ep = Ep; p.neg = neg; p.en = En; p.n = N; p.sig = const_cast<float*>(sig); p.gl = g_loss; p.gst = gst;
    recon_bwd_kernel<<<grid_for(Ep + En), 256, 0, (cudaStream_t)stream>>>(p);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_recon_loss_bwd");
}
