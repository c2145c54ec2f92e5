// Parameter packing / gradient unpacking between the reference's nn.Module parameters (raw device pointers, handed over
// in small pointer tables) and the kernels' weight / gradient blocks (include/mgv_b200.h).  One launch each instead of
// dozens of small framework ops per training step:
//   struct encoder  Wc = W_ih[:, :64] W_msg,  bc = W_ih[:, :64] b_msg   (AggConv composed into the GRU input weights,
//                   digae_layer.py:266-268 + arch/gcn_conv.py:36-42) and the chain rule back to W_msg, b_msg, W_ih;
//   level sweep     u = msg_k.weight^T attn_lin.weight[0, 64:]  (arch/tfmlp.py:39-42) and back to attn_lin / msg_k.
#include "sweep_layout.cuh"

namespace {

constexpr int D = MGV_D, G3 = 3 * MGV_D, D2 = 2 * MGV_D;
constexpr int SPACK = MGV_STRUCT_PACK_FLOATS;
constexpr int O_WCX = 0, O_WHH = 14592, O_BC = 27648, O_BIH = 27840, O_BHH = 28032, O_LNW = 28224, O_LNB = 28288;
constexpr int NLDC = 76, NLDM = 68;
constexpr int SW_PACK = MGV_SWEEP_PACK_FLOATS, SW_GRAD = MGV_SWEEP_GRAD_FLOATS;
constexpr int SO_U = 0, SO_BV = 8320, SO_BIH = 32960, SO_BHH = 33152, SO_WV = 33344, SO_WIH = 41536, SO_WHH = 53824;
constexpr int SG_U = 0, SG_WV = 128, SG_BV = 8320, SG_WIH = 8384, SG_WHH = 20672, SG_BIH = 32960, SG_BHH = 33152;

struct StructParams {                 // per (encoder, direction): msg.weight, msg.bias, weight_ih, weight_hh, bias_ih, bias_hh
    const float* p[2][2][6];
    const float* ln[2][2];            // per encoder: ln.weight, ln.bias (null without layernorm)
    int num_enc, feat, layernorm;
};

__global__ void struct_pack_kernel(const StructParams sp, float* __restrict__ pack) {
    const int blk = blockIdx.y, enc = blk >> 1, dir = blk & 1;
    const float* w = sp.p[enc][dir][0]; const float* b = sp.p[enc][dir][1]; const float* wih = sp.p[enc][dir][2];
    const float* whh = sp.p[enc][dir][3]; const float* bih = sp.p[enc][dir][4]; const float* bhh = sp.p[enc][dir][5];
    float* P = pack + (size_t)blk * SPACK;
    const int ldw = D + sp.feat;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < SPACK; i += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (i < O_WHH) {
            const int o = i / NLDC, c = i % NLDC;
            if (c < D) {                                              // Wc[o][c] = sum_k wih[o][k] w[k][c]
                float acc = 0.f;
#pragma unroll 32
                for (int k = 0; k < D; ++k) acc = fmaf(__ldg(wih + o * ldw + k), __ldg(w + k * D + c), acc);
                v = acc;
            } else if (c < D + sp.feat) v = __ldg(wih + o * ldw + c);
        } else if (i < O_BC) {
            const int j = i - O_WHH, o = j / NLDM, c = j % NLDM;
            if (c < D) v = __ldg(whh + o * D + c);
        } else if (i < O_BIH) {
            const int o = i - O_BC;
            float acc = 0.f;
#pragma unroll 32
            for (int k = 0; k < D; ++k) acc = fmaf(__ldg(wih + o * ldw + k), __ldg(b + k), acc);
            v = acc;
        } else if (i < O_BHH) v = __ldg(bih + i - O_BIH);
        else if (i < O_LNW) v = __ldg(bhh + i - O_BHH);
        else if (i < O_LNB) v = sp.layernorm ? __ldg(sp.ln[enc][0] + i - O_LNW) : 1.0f;
        else if (i < O_LNB + D) v = sp.layernorm ? __ldg(sp.ln[enc][1] + i - O_LNB) : 0.0f;
        P[i] = v;
    }
}

// Output (per encoder, floats): for dir 0 then dir 1: d msg.weight [64][64], d msg.bias [64], d weight_ih [192][64+feat],
// d weight_hh [192][64], d bias_ih [192], d bias_hh [192]; then (layernorm) d ln.weight [64], d ln.bias [64].
__global__ void struct_unpack_kernel(const StructParams sp, const float* __restrict__ grads, float* __restrict__ out, int per_enc) {
    const int blk = blockIdx.y, enc = blk >> 1, dir = blk & 1;
    const float* w = sp.p[enc][dir][0]; const float* b = sp.p[enc][dir][1]; const float* wih = sp.p[enc][dir][2];
    const float* G = grads + (size_t)blk * SPACK;
    const int ldw = D + sp.feat;
    const int per_dir = D * D + D + G3 * ldw + G3 * D + 2 * G3;
    float* O = out + (size_t)enc * per_enc + (size_t)dir * per_dir;
    // One output per thread (the host sizes the grid: 256-thread blocks over per_dir).  The parameters are cold in L2 by the end
    // of the backward and the dot products walk them at a row stride, so everything a block reads of them is staged in shared
    // memory by ONE wave of independent loads first (the version that read weight_ih inside the 192-step loops was a chain of
    // DRAM misses: 62 us per call).  Blocks 0 .. 15 own d msg.weight rows k0 .. k0 + 3 and, from the same staged columns, d msg.bias.
    __shared__ float wT[D][D + 1];          // msg.weight transposed: the d weight_ih products run over c with lanes over k
    __shared__ float s_a[G3][4];            // weight_ih[:, k0 .. k0 + 3]
    __shared__ float s_b[D];
    const int i0 = blockIdx.x * blockDim.x, i = i0 + threadIdx.x;
    const bool dw_block = i0 < D * D;                                          // block-uniform (D * D is a multiple of 256)
    const bool dih_block = i0 + (int)blockDim.x > D * D + D && i0 < D * D + D + G3 * ldw;
    const int k0 = i0 / D;
    if (dw_block)
        for (int t = threadIdx.x; t < G3 * 4; t += blockDim.x) s_a[t >> 2][t & 3] = __ldg(wih + (t >> 2) * ldw + k0 + (t & 3));
    if (dih_block) {
        for (int t = threadIdx.x; t < D * D; t += blockDim.x) wT[t % D][t / D] = __ldg(w + t);
        if (threadIdx.x < D) s_b[threadIdx.x] = __ldg(b + threadIdx.x);
    }
    __syncthreads();
    if (dw_block && threadIdx.x < 128) {                                       // d b[k0 + q] = sum_o wih[o][k0 + q] gbc[o], warp q
        const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < G3 / 32; ++t) acc = fmaf(s_a[lane + 32 * t][q], __ldg(G + O_BC + lane + 32 * t), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) O[D * D + k0 + q] = acc;
    }
    if (i < per_dir) {
        float v;
        int j = i;
        bool store = true;
        if (j < D * D) {                                             // d w[k][c] = sum_o wih[o][k] gWc[o][c]
            const int k = j / D, c = j % D;
            float acc = 0.f;
#pragma unroll 16
            for (int o = 0; o < G3; ++o) acc = fmaf(s_a[o][k - k0], __ldg(G + O_WCX + o * NLDC + c), acc);
            v = acc;
        } else if ((j -= D * D) < D) {                               // d b: written by the blocks above
            v = 0.f;
            store = false;
        } else if ((j -= D) < G3 * ldw) {                            // d wih[o][k<64] = sum_c gWc[o][c] w[k][c] + gbc[o] b[k]
            const int o = j / ldw, k = j % ldw;
            if (k < D) {
                float acc = __ldg(G + O_BC + o) * s_b[k];
#pragma unroll 16
                for (int c = 0; c < D; ++c) acc = fmaf(__ldg(G + O_WCX + o * NLDC + c), wT[c][k], acc);
                v = acc;
            } else v = __ldg(G + O_WCX + o * NLDC + k);
        } else if ((j -= G3 * ldw) < G3 * D) v = __ldg(G + O_WHH + (j / D) * NLDM + j % D);
        else if ((j -= G3 * D) < G3) v = __ldg(G + O_BIH + j);
        else v = __ldg(G + O_BHH + j - G3);
        if (store) O[i] = v;
    }
    if (sp.layernorm && dir == 0 && blockIdx.x == 0 && threadIdx.x < 2 * D) {
        const float* G1 = G + SPACK;
        out[(size_t)enc * per_enc + 2 * (size_t)per_dir + threadIdx.x] = G[O_LNW + threadIdx.x] + G1[O_LNW + threadIdx.x];
    }
}

struct SweepParams {                  // per listed code: attn_lin.weight [1][128], msg_k.weight [64][128], msg_v.weight [64][128],
    const float* p[MGV_NCODE][8];     // msg_v.bias, weight_ih [192][64], weight_hh [192][64], bias_ih, bias_hh
    int code[MGV_NCODE];
    int n;
};

// Blocks [0, SWEEP_NAT_BLOCKS) write the natural block; the blocks behind them the tensor-core weight image of the code
// (sweep_layout.cuh): Wc = W_ih W_v as fp16 hi/lo planes in the UMMA layout + the fp32 tail (u, composed biases).
constexpr int SWEEP_NAT_BLOCKS = 32;
constexpr int SWEEP_IMG_ITEMS = G3 * 16 + 384;
template <bool LOWP>
__global__ void sweep_pack_kernel(const SweepParams sp, float* __restrict__ pack) {
    const int q = blockIdx.y;
    if (q >= sp.n) return;
    const float* aw = sp.p[q][0]; const float* kw = sp.p[q][1]; const float* vw = sp.p[q][2]; const float* vb = sp.p[q][3];
    const float* wih = sp.p[q][4]; const float* whh = sp.p[q][5]; const float* bih = sp.p[q][6]; const float* bhh = sp.p[q][7];
    float* P = pack + (size_t)sp.code[q] * SW_PACK;
    if (blockIdx.x >= SWEEP_NAT_BLOCKS) {
        using namespace sweep_layout;
        uint8_t* img = reinterpret_cast<uint8_t*>(pack) + IMG_OFFSET + (size_t)sp.code[q] * IMG_PAD;
        const int i = (blockIdx.x - SWEEP_NAT_BLOCKS) * blockDim.x + threadIdx.x;
        if (i < G3 * 16) {                                   // Wc row g, 8 consecutive input columns: sum_k wih[g][k] vw[k][.]
            const int g = i >> 4, c = i & 15;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.f;
#pragma unroll 8
            for (int k = 0; k < D; ++k) {
                const float w = __ldg(wih + g * D + k);
                const float4 a = mgv_ldg4(vw + k * D2 + c * 8), b = mgv_ldg4(vw + k * D2 + c * 8 + 4);
                v[0] = fmaf(w, a.x, v[0]); v[1] = fmaf(w, a.y, v[1]); v[2] = fmaf(w, a.z, v[2]); v[3] = fmaf(w, a.w, v[3]);
                v[4] = fmaf(w, b.x, v[4]); v[5] = fmaf(w, b.y, v[5]); v[6] = fmaf(w, b.z, v[6]); v[7] = fmaf(w, b.w, v[7]);
            }
            uint4 hi, lo;
            tc::split8p<LOWP>(v, hi, lo);
            const uint32_t off = (uint32_t)(c >> 3) * KB_W + tc::sw128_off(g, c & 7);
            *reinterpret_cast<uint4*>(img + I_WC_HI + off) = hi;
            *reinterpret_cast<uint4*>(img + I_WC_LO + off) = lo;
        } else if (i < SWEEP_IMG_ITEMS) {
            const int j = i - G3 * 16;
            float v = 0.f;
            if (j < 128) {                                   // u[k] = sum_o aw[64 + o] kw[o][k]
#pragma unroll 32
                for (int o = 0; o < D; ++o) v = fmaf(__ldg(aw + D + o), __ldg(kw + o * D2 + j), v);
            } else {
                const int t = (j - 128) >> 6, u = (j - 128) & 63;          // 0 b_r, 1 b_z, 2 b_in, 3 b_hn
                if (t == 3) v = __ldg(bhh + 2 * D + u);
                else {
                    const int g = t * D + u;
                    v = __ldg(bih + g) + (t < 2 ? __ldg(bhh + g) : 0.f);
#pragma unroll 32
                    for (int k = 0; k < D; ++k) v = fmaf(__ldg(wih + g * D + k), __ldg(vb + k), v);
                }
            }
            reinterpret_cast<float*>(img + I_F32)[j] = v;
        }
        return;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < SW_PACK; i += SWEEP_NAT_BLOCKS * blockDim.x) {
        float v = 0.f;
        if (i < 128) {                                               // u[k] = sum_o aw[64 + o] kw[o][k]
            float acc = 0.f;
            for (int o = 0; o < D; ++o) acc = fmaf(__ldg(aw + D + o), __ldg(kw + o * D2 + i), acc);
            v = acc;
        } else if (i >= SO_BV && i < SO_BV + D) v = __ldg(vb + i - SO_BV);
        else if (i >= SO_BIH && i < SO_BIH + G3) v = __ldg(bih + i - SO_BIH);
        else if (i >= SO_BHH && i < SO_BHH + G3) v = __ldg(bhh + i - SO_BHH);
        else if (i >= SO_WV && i < SO_WIH) v = __ldg(vw + i - SO_WV);
        else if (i >= SO_WIH && i < SO_WHH) v = __ldg(wih + i - SO_WIH);
        else if (i >= SO_WHH && i < SO_WHH + G3 * D) v = __ldg(whh + i - SO_WHH);
        P[i] = v;                                                    // (the transposed copies of the first generation stay zero)
    }
}

// Output per listed code (floats): d attn_lin.weight [128] (first 64 zero), d msg_k.weight [64][128], then the block's
// natural gradients are used in place (d msg_v.weight, d msg_v.bias, d weight_ih, d weight_hh, d bias_ih, d bias_hh).
__global__ void sweep_unpack_kernel(const SweepParams sp, const float* __restrict__ grads, float* __restrict__ out) {
    const int q = blockIdx.y;
    if (q >= sp.n) return;
    const float* aw = sp.p[q][0]; const float* kw = sp.p[q][1];
    const float* du = grads + (size_t)sp.code[q] * SW_GRAD + SG_U;
    float* O = out + (size_t)q * (128 + D * D2);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 128 + D * D2; i += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (i < 128) {
            if (i >= D) {                                            // d aw[64 + o] = sum_k kw[o][k] du[k]
                const int o = i - D;
                for (int k = 0; k < D2; ++k) v = fmaf(__ldg(kw + o * D2 + k), __ldg(du + k), v);
            }
        } else {
            const int j = i - 128, o = j / D2, k = j % D2;           // d kw[o][k] = aw[64 + o] du[k]
            v = __ldg(aw + D + o) * __ldg(du + k);
        }
        O[i] = v;
    }
}

}  // namespace

extern "C" int mgv_struct_pack(const void* const* params, int32_t num_enc, int32_t layernorm, int32_t feat, float* pack,
                               mgv_stream_t stream) {
    MGV_REQUIRE(params && pack && num_enc >= 1 && num_enc <= 2 && feat >= 0 && feat <= MGV_MAX_FEAT, "mgv_struct_pack: bad argument");
    StructParams sp{};
    sp.num_enc = num_enc; sp.feat = feat; sp.layernorm = layernorm;
    const int per = 12 + (layernorm ? 2 : 0);
    for (int e = 0; e < num_enc; ++e) {
        for (int d = 0; d < 2; ++d)
            for (int k = 0; k < 6; ++k) sp.p[e][d][k] = (const float*)params[e * per + d * 6 + k];
        if (layernorm) { sp.ln[e][0] = (const float*)params[e * per + 12]; sp.ln[e][1] = (const float*)params[e * per + 13]; }
    }
    struct_pack_kernel<<<dim3(28, num_enc * 2), 256, 0, (cudaStream_t)stream>>>(sp, pack);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_struct_pack");
}

extern "C" int mgv_struct_unpack_grads(const void* const* params, int32_t num_enc, int32_t layernorm, int32_t feat,
                                       const float* grads, float* out, mgv_stream_t stream) {
    MGV_REQUIRE(params && grads && out && num_enc >= 1 && num_enc <= 2, "mgv_struct_unpack_grads: bad argument");
    StructParams sp{};
    sp.num_enc = num_enc; sp.feat = feat; sp.layernorm = layernorm;
    const int per = 12 + (layernorm ? 2 : 0);
    for (int e = 0; e < num_enc; ++e)
        for (int d = 0; d < 2; ++d)
            for (int k = 0; k < 6; ++k) sp.p[e][d][k] = (const float*)params[e * per + d * 6 + k];
    const int ldw = D + feat;
    const int per_dir = D * D + D + G3 * ldw + G3 * D + 2 * G3;
    const int per_enc = 2 * per_dir + (layernorm ? 2 * D : 0);
    struct_unpack_kernel<<<dim3((per_dir + 255) / 256, num_enc * 2), 256, 0, (cudaStream_t)stream>>>(sp, grads, out, per_enc);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_struct_unpack_grads");
}

extern "C" size_t mgv_sweep_pack_bytes(void) { return sweep_layout::PACK_TOTAL_BYTES; }

extern "C" int mgv_sweep_pack(const void* const* params, const int32_t* codes, int32_t n, float* pack, int32_t precision,
                              mgv_stream_t stream) {
    MGV_REQUIRE(params && codes && pack && n >= 0 && n <= MGV_NCODE, "mgv_sweep_pack: bad argument");
    MGV_REQUIRE(precision == 0 || precision == 1, "mgv_sweep_pack: precision must be 0 (fp32-accurate) or 1 (bf16)");
    if (n == 0) return MGV_OK;
    SweepParams sp{};
    sp.n = n;
    for (int q = 0; q < n; ++q) {
        MGV_REQUIRE(codes[q] >= 1 && codes[q] <= 6, "mgv_sweep_pack: gate code %d outside 1..6", codes[q]);
        sp.code[q] = codes[q];
        for (int k = 0; k < 8; ++k) sp.p[q][k] = (const float*)params[q * 8 + k];
    }
    const dim3 grid(SWEEP_NAT_BLOCKS + (SWEEP_IMG_ITEMS + 255) / 256, n);
    if (precision == 1) sweep_pack_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(sp, pack);
    else sweep_pack_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(sp, pack);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_sweep_pack");
}

extern "C" int mgv_sweep_unpack_grads(const void* const* params, const int32_t* codes, int32_t n, const float* grads, float* out,
                                      mgv_stream_t stream) {
    MGV_REQUIRE(params && codes && grads && out && n >= 0 && n <= MGV_NCODE, "mgv_sweep_unpack_grads: bad argument");
    if (n == 0) return MGV_OK;
    SweepParams sp{};
    sp.n = n;
    for (int q = 0; q < n; ++q) {
        sp.code[q] = codes[q];
        for (int k = 0; k < 8; ++k) sp.p[q][k] = (const float*)params[q * 8 + k];
    }
    sweep_unpack_kernel<<<dim3(8, n), 256, 0, (cudaStream_t)stream>>>(sp, grads, out);
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_sweep_unpack_grads");
}
