// Weight / bias gradient of the small nn.Linear layers around the path (hs_linear 128->64, hs_decompose 64->128
// (dg_ae_model_mig.py:46-47) and the readout MLP 64->32->32->1 (arch/mlp.py:14-56)):
//     dW[O][I] = sum_n gy[n][O] x[n][I],   db[O] = sum_n gy[n][O],   N = nodes of the batch (10^4 .. 10^7), O, I <= 128.
// The output is at most 128 x 128, so a library GEMM runs it as ONE tile on ONE SM with K = N (measured: 133 us per
// layer at N = 65 818, cuBLAS SIMT sgemm); here the node dimension is split over a persistent grid, each CTA
// accumulates its rows in registers (fp32 FFMA, deterministic order) and a second pass sums the per-CTA partials.
#include "mgv_common.cuh"

namespace {

constexpr int RG = 8;              // partial groups summed in parallel by the reduce kernel

// Thread (ty, tx) owns the contiguous TO x TI output block (TO ty .., TI tx ..): 128-bit shared-memory reads,
// TO + TI loaded floats per TO * TI FMAs.
template <int TO, int TI, int LR>
__global__ void __launch_bounds__(256) linear_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, long long N,
                                                           int I, int O, int IP, int OP, float* __restrict__ partial) {
    extern __shared__ __align__(16) float sm[];
    const int stage_floats = LR * (IP + OP);           // two stages: [LR][IP] rows of x, then [LR][OP] rows of gy
    const int tid = threadIdx.x, nt = blockDim.x;
    const int ntx = IP / TI, tx = tid % ntx, ty = tid / ntx;
    const bool active = ty < OP / TO;
    float acc[TO][TI];
    float bsum[TO];
#pragma unroll
    for (int a = 0; a < TO; ++a) {
        bsum[a] = 0.f;
#pragma unroll
        for (int b = 0; b < TI; ++b) acc[a][b] = 0.f;
    }
    const long long trips = (N + LR - 1) / LR;
    const long long t0 = trips * blockIdx.x / gridDim.x, t1 = trips * (blockIdx.x + 1) / gridDim.x;
    const bool vec = (I == IP) && (O == OP) && (I % 4 == 0) && (O % 4 == 0);
    // stage the rows of trip t (asynchronous 16-byte copies when the rows are dense and aligned; rows past N read 0 bytes)
    auto stage = [&](long long t, int buf) {
        float* Xs = sm + buf * stage_floats;
        float* Gs = Xs + LR * IP;
        const long long r0 = t * LR;
        const int rows = (int)((N - r0 < LR) ? (N - r0) : LR);
        if (vec) {
            const float* xs = x + r0 * I;
            const float* gs = gy + r0 * O;
            for (int q = tid; q < LR * I / 4; q += nt) {
                const unsigned src = (q < rows * I / 4) ? 16u : 0u;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(Xs + 4 * q)),
                             "l"(src ? xs + 4 * q : x), "r"(src) : "memory");
            }
            for (int q = tid; q < LR * O / 4; q += nt) {
                const unsigned src = (q < rows * O / 4) ? 16u : 0u;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(Gs + 4 * q)),
                             "l"(src ? gs + 4 * q : gy), "r"(src) : "memory");
            }
        } else {
            for (int q = tid; q < LR * IP; q += nt) {
                const int r = q / IP, c = q % IP;
                Xs[q] = (r < rows && c < I) ? x[(r0 + r) * I + c] : 0.f;
            }
            for (int q = tid; q < LR * OP; q += nt) {
                const int r = q / OP, c = q % OP;
                Gs[q] = (r < rows && c < O) ? gy[(r0 + r) * O + c] : 0.f;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (t0 < t1) stage(t0, 0);
    for (long long t = t0; t < t1; ++t) {
        const int buf = (int)((t - t0) & 1);
        if (t + 1 < t1) {
            stage(t + 1, buf ^ 1);                     // the other stage was released by the barrier at the end of the last trip
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const float* Xs = sm + buf * stage_floats;
        const float* Gs = Xs + LR * IP;
        if (active) {
#pragma unroll 4
            for (int r = 0; r < LR; ++r) {
                float g[TO], xv[TI];
#pragma unroll
                for (int a = 0; a < TO; a += 4) *reinterpret_cast<float4*>(&g[a]) = *reinterpret_cast<const float4*>(Gs + r * OP + TO * ty + a);
                // x columns of this thread: 4 tx .. 4 tx + 3 and (8-wide tiles) IP / 2 + 4 tx ..: a 16-byte lane stride
#pragma unroll
                for (int b = 0; b < TI; b += 4) *reinterpret_cast<float4*>(&xv[b]) = *reinterpret_cast<const float4*>(Xs + r * IP + (b / 4) * (IP / 2) + 4 * tx);
#pragma unroll
                for (int a = 0; a < TO; ++a) {
                    bsum[a] += g[a];
#pragma unroll
                    for (int b = 0; b < TI; ++b) acc[a][b] = fmaf(g[a], xv[b], acc[a][b]);
                }
            }
        }
        __syncthreads();
    }
    if (!active) return;
    float* P = partial + (size_t)blockIdx.x * ((size_t)O * I + O);
#pragma unroll
    for (int a = 0; a < TO; ++a) {
        const int o = TO * ty + a;
        if (o < O) {
#pragma unroll
            for (int b = 0; b < TI; ++b) {
                const int i = (b / 4) * (IP / 2) + 4 * tx + (b & 3);
                if (i < I) P[(size_t)o * I + i] = acc[a][b];
            }
            if (tx == 0) P[(size_t)O * I + o] = bsum[a];
        }
    }
}

// 256 threads = RG groups x 32 consecutive outputs; group g sums partial blocks g, g + RG, ...
__global__ void __launch_bounds__(256) linear_wgrad_reduce_kernel(const float* __restrict__ partial, int nblk, int I, int O,
                                                                  float* __restrict__ dW, float* __restrict__ db) {
    __shared__ float red[RG][32];
    const int total = O * I + O;
    const int j = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + j;
    float s = 0.f;
    if (i < total) {
#pragma unroll 4
        for (int b = g; b < nblk; b += RG) s += partial[(size_t)b * total + i];
    }
    red[g][j] = s;
    __syncthreads();
    if (g == 0 && i < total) {
#pragma unroll
        for (int k = 1; k < RG; ++k) s += red[k][j];
        if (i < O * I) dW[i] = s;
        else if (db) db[i - O * I] = s;
    }
}

int grid_blocks(long long N) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long trips = (N + 127) / 128;
    long long g = 2LL * sms;
    if (g > trips) g = trips;
    return g < 1 ? 1 : (int)g;
}

template <int TO, int TI, int LR>
int launch(int nblk, cudaStream_t st, const float* x, const float* gy, long long N, int I, int O, float* partial) {
    const int IP = (I + TI - 1) / TI * TI, OP = (O + TO - 1) / TO * TO;
    int nt = (IP / TI) * (OP / TO);
    nt = nt < 256 ? 256 : (nt + 31) / 32 * 32;          // idle compute threads still help stage the rows
    const size_t smem = 2 * (size_t)LR * (IP + OP) * sizeof(float);
    MGV_CUDA(cudaFuncSetAttribute((const void*)linear_wgrad_kernel<TO, TI, LR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    linear_wgrad_kernel<TO, TI, LR><<<nblk, nt, smem, st>>>(x, gy, N, I, O, IP, OP, partial);
    return MGV_OK;
}

}  // namespace

extern "C" size_t mgv_linear_wgrad_workspace_bytes(int64_t N, int32_t I, int32_t O) {
    return (size_t)grid_blocks(N) * ((size_t)O * I + O) * sizeof(float) + 256;
}

extern "C" int mgv_linear_wgrad(const float* x, const float* gy, int64_t N, int32_t I, int32_t O, float* dW, float* db,
                                void* ws, size_t ws_bytes, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MGV_REQUIRE(I >= 1 && I <= 128 && O >= 1 && O <= 128, "mgv_linear_wgrad: in/out features must be in 1..128 (got %d, %d)", I, O);
    MGV_REQUIRE(N >= 0 && dW != nullptr, "mgv_linear_wgrad: bad argument");
    if (N == 0) {
        MGV_CUDA(cudaMemsetAsync(dW, 0, (size_t)O * I * sizeof(float), st));
        if (db) MGV_CUDA(cudaMemsetAsync(db, 0, (size_t)O * sizeof(float), st));
        return MGV_OK;
    }
    if (ws_bytes < mgv_linear_wgrad_workspace_bytes(N, I, O)) {
        mgv_set_error("mgv_linear_wgrad: workspace %zu < %zu bytes", ws_bytes, mgv_linear_wgrad_workspace_bytes(N, I, O));
        return MGV_ERR_WORKSPACE;
    }
    const int nblk = grid_blocks(N);
    float* partial = (float*)ws;
    // thread tile: 8 wide along a dimension of >= 64 features, else 4 (<= 16 x 16 threads)
    const bool o8 = O > 32, i8 = I > 32;
    int rc;
    if (o8 && i8) rc = launch<8, 8, 64>(nblk, st, x, gy, N, I, O, partial);
    else if (o8) rc = launch<8, 4, 128>(nblk, st, x, gy, N, I, O, partial);
    else if (i8) rc = launch<4, 8, 128>(nblk, st, x, gy, N, I, O, partial);
    else rc = launch<4, 4, 128>(nblk, st, x, gy, N, I, O, partial);
    if (rc != MGV_OK) return rc;
    const int total = O * I + O;
    linear_wgrad_reduce_kernel<<<(total + 31) / 32, 256, 0, st>>>(partial, nblk, I, O, dW, db);
    mgv_count_launches(2);
    return mgv_check_cuda(cudaGetLastError(), "mgv_linear_wgrad");
}
