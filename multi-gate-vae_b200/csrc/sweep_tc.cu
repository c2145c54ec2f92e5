// Level sweep on the 5th-generation tensor cores: the level-synchronous TFMlpAggr + GRU propagation of Model.forward
// (dg_ae_model_mig.py:84-129; aig :70-97, xmg :95-147, xag :88-121; arch/tfmlp.py:31-46) and its backward, each ONE
// persistent cooperative kernel over all levels, for the single-round sweep (num_rounds = 1, the reference default:
// h = hf[node] = 0 when a node is updated).  Multi-round sweeps stay on the mma.sync kernels of sweep.cu.
//
// Math per node i of gate code T at level >= 1:
//   x_j = [hs_j || hf_j] over the predecessors j (ascending edge id);  score_j = u_T . x_j   (u = msg_k.weight^T attn_lin.weight[64:];
//   the query part, msg_k.bias and attn_lin.bias are constant inside a softmax group and cancel)
//   alpha = softmax(score) (PyG: exp(a - max) / (sum + 1e-16));  xbar = sum_j alpha_j x_j
//   GRU(x = W_v xbar + b_v, h = 0):  [r z n]_pre = Wc xbar + bias,  Wc = W_ih W_v (192 x 128, composed once per call),
//   bias = W_ih b_v + b_ih (+ b_hh for r, z);  r = sig, z = sig, n = tanh(n_pre + r b_hn);  hf_i = (1 - z) n
//
// A level moves ~2 KB per gate and has only hundreds to thousands of gates, so a level is LATENCY bound: barrier -> index
// chain -> row gather -> dense products -> store -> barrier.  The design removes what can be removed from that chain:
//   * every level is spread over ALL CTAs of its gate code (tile = ceil(segment / CTAs) rows, at most 64), so the gather of
//     a level runs on every SM at once;
//   * the products are issued TRANSPOSED -- D^T[gate unit][node] = Wc[gate unit][k] xbar^T[k][node] -- so the tensor-core
//     time of a tile is proportional to its NODE count (UMMA N = rows rounded to 16) instead of a fixed 128-row M tile;
//     the accumulator lives in tensor memory with TMEM lane = gate unit, which makes the biases per-thread constants;
//   * the static part of the index chain (order -> in_ptr -> in_src) is prefetched BEFORE the level barrier is awaited;
//   * the grid barrier is polled by one thread that then releases the workers through a shared-memory mbarrier.
// Operands are fp16 hi/lo planes with three products per K step (mgv_tc.cuh): fp32-accurate.
#include <stdlib.h>
#include <string.h>
#include "sweep_layout.cuh"

namespace sweep_tc {
using namespace sweep_layout;

constexpr int D = MGV_D, D2 = 2 * MGV_D, G3 = 3 * MGV_D;
constexpr int PACK = MGV_SWEEP_PACK_FLOATS, GRAD = MGV_SWEEP_GRAD_FLOATS;
// natural weight block (include/mgv_b200.h)
constexpr int O_U = 0, O_BV = 8320, O_BIH = 32960, O_BHH = 33152, O_WV = 33344, O_WIH = 41536;
constexpr int G_U = 0, G_WV = 128, G_BV = 8320, G_WIH = 8384, G_WHH = 20672, G_BIH = 32960, G_BHH = 33152;
constexpr int NODE_MASK = (1 << MGV_CODE_SHIFT) - 1;

constexpr int WORKERS = 16;                     // worker warps; warp WORKERS issues the MMAs and polls the grid barrier
constexpr int NWT = WORKERS * 32;
constexpr int NTHREADS = NWT + 32;

struct SweepTC {
    int N, L;
    unsigned handled;
    const int* order; const int* seg_ptr; const int* in_ptr; const int* in_src;
    const int* out_ptr; const int* out_pack; const int* out_slot;
    const uint8_t* image;      // [MGV_NCODE][IMG_PAD] (behind the natural blocks in the pack buffer)
    const float* weights;      // natural blocks (u of every code, for the pulls)
    const float* hs;
    float* hf;                 // [N][64]
    int cta_start[MGV_NCODE + 1];
    unsigned* bar;             // [0] grid barrier counter
    // backward
    float* ghs; float* ghf; float* dxb; float* alpha; float* dscore; float* raw;
};

// ------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ float4 ldcg4(const float* p) {          // L2-coherent load: rows written by other CTAs during the kernel
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldcg1(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t a) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void red_add1(float* dst, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst), "f"(v) : "memory");
}
template <bool LOWP>
__device__ __forceinline__ void split2p(float a, float b, uint32_t& hi, uint32_t& lo) {
    if (LOWP) { hi = tc::pack_bf16x2(a, b); lo = 0u; }
    else tc::split2(a, b, hi, lo);
}
__device__ __forceinline__ float sigmoid_fast(float x) {
    return __fdividef(1.0f, 1.0f + __expf(-fminf(fmaxf(x, -28.f), 28.f)));
}
__device__ __forceinline__ float tanh_fast(float x) {
    const float y = fminf(fmaxf(x, -14.f), 14.f);
    return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * y));
}

__device__ __forceinline__ void find_role(const SweepTC& p, int& code, int& rank, int& nct) {
    code = -1; rank = 0; nct = 1;
    const int b = blockIdx.x;
#pragma unroll
    for (int c = 0; c < MGV_NCODE; ++c) {
        if (b >= p.cta_start[c] && b < p.cta_start[c + 1]) {
            code = c; rank = b - p.cta_start[c]; nct = p.cta_start[c + 1] - p.cta_start[c];
        }
    }
}

// Tiles of one (level, code) segment for CTA `rank` of `nct`: the segment is cut into ceil(n / nct) rows per tile (at most
// NT), tile k belongs to CTA k % nct -- every CTA of the code gets a share of every level.
struct TileIter {
    int sbeg, n, per, ntiles, k;
    __device__ __forceinline__ void start(const int* __restrict__ seg_ptr, int lvl, int code, int rank, int nct) {
        sbeg = __ldg(seg_ptr + lvl * MGV_NCODE + code);
        n = __ldg(seg_ptr + lvl * MGV_NCODE + code + 1) - sbeg;
        per = n > 0 ? min(NT, (n + nct - 1) / nct) : 1;
        ntiles = (n + per - 1) / per;
        k = rank;
    }
    __device__ __forceinline__ bool next(int nct, int& t0, int& rows) {
        if (k >= ntiles) return false;
        t0 = sbeg + k * per;
        rows = min(per, n - k * per);
        k += nct;
        return true;
    }
};

// Lane-distributed descriptors of the (at most 4) rows warp `warp` owns in a tile: row i = warp + 16 k is held by the 8
// lanes of slot k = lane / 8; lane q = lane % 8 of the slot additionally holds the q-th predecessor id.  All of it is static
// schedule data, so it is loaded before the level barrier is awaited.
struct RowRegs { int node, beg, cnt, src; };
__device__ __forceinline__ RowRegs prefetch_rows(const SweepTC& p, int t0, int rows, int warp, int lane) {
    RowRegs rr;
    rr.node = -1; rr.beg = 0; rr.cnt = 0; rr.src = 0;
    const int i = warp + WORKERS * (lane >> 3), q = lane & 7;
    if (i < rows) {
        rr.node = __ldg(p.order + t0 + i);
        rr.beg = __ldg(p.in_ptr + rr.node);
        rr.cnt = __ldg(p.in_ptr + rr.node + 1) - rr.beg;
        if (q < rr.cnt) rr.src = __ldg(p.in_src + rr.beg + q);
    }
    return rr;
}

// Gather + additive attention of ONE node by one warp.  Lane l owns columns 4 (l % 16) .. + 3 of the hs part (l < 16) or the
// hf part (l >= 16) of the 128-wide row.  `al` receives the first four attention weights (all lanes), used again by the
// backward's attention phase.  STORE_ALPHA: the weights also go to p.alpha (indexed by in-CSR slot) for the pulls.
template <bool STORE_ALPHA, bool HF_CG>
__device__ __forceinline__ void attend_row(const SweepTC& p, const float* __restrict__ hf, const float4& u4, int slot, const RowRegs& rr,
                                           int lane, float4& xbar, float (&al)[4]) {
    const unsigned full = 0xffffffffu;
    const int cnt = __shfl_sync(full, rr.cnt, slot * 8), beg = __shfl_sync(full, rr.beg, slot * 8);
    const int off = 4 * (lane & 15);
    const float* base = (lane < 16) ? p.hs : hf;
    float4 x[4];
    float sc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = __shfl_sync(full, rr.src, slot * 8 + q);
        x[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < cnt) {
            const float* row = base + (size_t)j * D + off;
            x[q] = (HF_CG && lane >= 16) ? ldcg4(row) : mgv_ld4(row);
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) sc[q] = mgv_warp_sum(mgv_dot4(x[q], u4));
    float mx = -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (q < cnt) mx = fmaxf(mx, sc[q]);
    float sum = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        al[q] = 0.f;
        if (q < cnt) {
            al[q] = expf(sc[q] - mx);
            sum += al[q];
            mgv_fma4(acc, al[q], x[q]);
        }
    }
    if (cnt > 4) {
        // rare: more than four predecessors -- online softmax over the rest, one at a time (scores parked in p.alpha)
        if (STORE_ALPHA && lane < 4) p.alpha[beg + lane] = sc[lane];
        for (int q = 4; q < cnt; ++q) {
            const int j = (q < 8) ? __shfl_sync(full, rr.src, slot * 8 + q) : __ldg(p.in_src + beg + q);
            const float* row = base + (size_t)j * D + off;
            const float4 xq = (HF_CG && lane >= 16) ? ldcg4(row) : mgv_ld4(row);
            const float s = mgv_warp_sum(mgv_dot4(xq, u4));
            if (STORE_ALPHA && lane == 0) p.alpha[beg + q] = s;
            const float nmx = fmaxf(mx, s);
            const float f = expf(mx - nmx), e = expf(s - nmx);
            sum = sum * f + e;
            acc.x = acc.x * f + e * xq.x; acc.y = acc.y * f + e * xq.y; acc.z = acc.z * f + e * xq.z; acc.w = acc.w * f + e * xq.w;
#pragma unroll
            for (int t = 0; t < 4; ++t) al[t] *= f;
            mx = nmx;
        }
    }
    const float inv = 1.0f / (sum + 1e-16f);
    xbar = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
#pragma unroll
    for (int q = 0; q < 4; ++q) al[q] *= inv;
    if (STORE_ALPHA) {
        if (cnt <= 4) {
            const float mine = lane == 0 ? al[0] : (lane == 1 ? al[1] : (lane == 2 ? al[2] : al[3]));
            if (lane < cnt) p.alpha[beg + lane] = mine;
        } else {
            __syncwarp();
            for (int q = lane; q < cnt; q += 32) p.alpha[beg + q] = expf(p.alpha[beg + q] - mx) * inv;
            __syncwarp();
        }
    }
}

// The lane's 4 values of row `row` -> fp16 hi/lo planes of a node tile (SW128, K-major; K block = hs / hf part).
template <bool LOWP>
__device__ __forceinline__ void store_xbar(uint32_t xb_hi, uint32_t xb_lo, int row, int lane, const float4& v) {
    uint32_t h0, l0, h1, l1;
    split2p<LOWP>(v.x, v.y, h0, l0);
    split2p<LOWP>(v.z, v.w, h1, l1);
    const uint32_t off = (uint32_t)(lane >> 4) * KB_X + tc::sw128_off(row, (lane & 15) >> 1) + (uint32_t)(lane & 1) * 8u;
    st_shared_v2(xb_hi + off, h0, h1);
    if (!LOWP) st_shared_v2(xb_lo + off, l0, l1);
}

// ======================================================================================= forward
constexpr uint32_t F_XB_HI = IMG_PAD, F_XB_LO = F_XB_HI + 2 * KB_X;
constexpr uint32_t F_ZX = F_XB_LO + 2 * KB_X;                    // z pre-activations [NT][64] fp32 (lane exchange)
constexpr uint32_t F_IDS = F_ZX + NT * D * 4;
constexpr uint32_t F_BAR = F_IDS + NT * 4;                       // mbarriers: w, x_full, acc_full, level
constexpr uint32_t F_TMEM = F_BAR + 64;
constexpr uint32_t F_SMEM = F_TMEM + 64 + 1024;                  // + alignment slack
constexpr uint32_t TF_ACC1 = 0, TF_ACC2 = NT, TF_COLS = 128;     // tensor memory: [r | z] lanes, [n | -] lanes, NT node columns each

// The recompute / forward products of one tile: acc1 = Wc[0:128] xbar^T, acc2 = Wc[128:256] xbar^T (rows 192.. are whatever
// follows the image: those accumulator lanes are never read).
template <bool LOWP>
__device__ __forceinline__ void issue_gate_mmas(uint32_t sbase, uint32_t xb_hi, uint32_t xb_lo, uint32_t acc1, uint32_t acc2, int npad) {
    const uint32_t idesc = tc::make_idesc(128, npad, false, false, LOWP);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t wo = (uint32_t)kb * KB_W + 32u * j, xo = (uint32_t)kb * KB_X + 32u * j;
            const uint64_t b_hi = tc::desc_k_sw128(xb_hi + xo), b_lo = tc::desc_k_sw128(xb_lo + xo);
            const uint32_t accf = (kb | j) ? 1u : 0u;
            tc::mma3p<LOWP>(acc1, tc::desc_k_sw128(sbase + I_WC_HI + wo), tc::desc_k_sw128(sbase + I_WC_LO + wo), b_hi, b_lo, idesc, accf);
            tc::mma3p<LOWP>(acc2, tc::desc_k_sw128(sbase + I_WC_HI + wo + 16384u), tc::desc_k_sw128(sbase + I_WC_LO + wo + 16384u), b_hi, b_lo,
                            idesc, accf);
        }
}

// Grid barrier, split: the workers' thread 0 arrives (release), the MMA warp's thread polls (acquire) and releases the
// workers through a shared-memory mbarrier.
__device__ __forceinline__ void grid_arrive(unsigned* counter) {
    __threadfence();
    atomicAdd(counter, 1u);
}
__device__ __forceinline__ void grid_poll(const unsigned* counter, unsigned target) {
    while (mgv_ld_acquire(counter) < target) __nanosleep(20);
    __threadfence();
}

template <bool LOWP>
__global__ void __launch_bounds__(NTHREADS, 1) sweep_fwd_tc_kernel(const SweepTC p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - tc::smem_u32(smem_raw));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_w = sbase + F_BAR, bar_x_full = bar_w + 8, bar_acc_full = bar_w + 16, bar_level = bar_w + 24;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + F_TMEM);
    float* ZX = reinterpret_cast<float*>(sgen + F_ZX);
    int* IDS = reinterpret_cast<int*>(sgen + F_IDS);
    int code, rank, nct;
    find_role(p, code, rank, nct);
    if (tid == 0) {
        tc::mbar_init(bar_w, 1);
        tc::mbar_init(bar_x_full, NWT);
        tc::mbar_init(bar_acc_full, 1);
        tc::mbar_init(bar_level, 1);
        tc::fence_barrier_init();
        if (code >= 0) {
            const uint8_t* img = p.image + (size_t)code * IMG_PAD;
            tc::mbar_expect_tx(bar_w, IMG_BYTES);
#pragma unroll 1
            for (uint32_t o = 0; o < I_F32; o += 16384u) tc::bulk_g2s(sbase + o, img + o, 16384u, bar_w);
            tc::bulk_g2s(sbase + I_F32, img + I_F32, IMG_BYTES - I_F32, bar_w);
        }
    }
    if (warp == WORKERS) tc::tmem_alloc(tmem_slot, TF_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int nsteps = p.L - 1;

    if (warp < WORKERS) {
        // ===================================================================== workers: gather -> [MMA] -> epilogue
        float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const int qd = warp & 3, cg = warp >> 2;
        const int eu = (qd & 1) * 32 + lane;                   // gate unit of this thread's tensor-memory lane
        float b_r = 0.f, b_z = 0.f, b_in = 0.f, b_hn = 0.f;
        if (code >= 0) {
            tc::mbar_wait_warp(bar_w, 0u, lane);
            const float* F = reinterpret_cast<const float*>(sgen + I_F32);
            u4 = *reinterpret_cast<const float4*>(F + 4 * lane);
            b_r = F[128 + eu]; b_z = F[192 + eu]; b_in = F[256 + eu]; b_hn = F[320 + eu];
        }
        (void)b_z;
        const uint32_t tl = tmem + ((uint32_t)(qd * 32) << 16);
        uint32_t it = 0;
        for (int step = 0; step < nsteps; ++step) {
            const int lvl = step + 1;
            TileIter ti;
            int t0 = 0, rows = 0;
            bool have = false;
            RowRegs rr;
            if (code >= 0) {
                ti.start(p.seg_ptr, lvl, code, rank, nct);
                have = ti.next(nct, t0, rows);
                if (have) rr = prefetch_rows(p, t0, rows, warp, lane);
            }
            if (step > 0) tc::mbar_wait_warp(bar_level, (uint32_t)((step - 1) & 1), lane, 32);     // level step - 1 is complete everywhere
            while (have) {
                const int npad = (rows + 15) & ~15;
                // ---- gather + attention: rows warp, warp + 16, ..
#pragma unroll
                for (int k = 0; k < NT / WORKERS; ++k) {
                    const int i = warp + WORKERS * k;
                    if (i < rows) {
                        float4 xbar;
                        float al[4];
                        attend_row<false, true>(p, p.hf, u4, k, rr, lane, xbar, al);
                        store_xbar<LOWP>(sbase + F_XB_HI, sbase + F_XB_LO, i, lane, xbar);
                    }
                }
                // (lane 0 is q = 0 of slot 0 only: fetch each slot's node id from its first lane)
#pragma unroll
                for (int k = 0; k < NT / WORKERS; ++k) {
                    const int nd = __shfl_sync(0xffffffffu, rr.node, k * 8);
                    if (lane == 0 && warp + WORKERS * k < rows) IDS[warp + WORKERS * k] = nd;
                }
                tc::fence_async_smem();
                tc::mbar_arrive(bar_x_full);
                tc::mbar_wait_warp(bar_acc_full, it & 1u, lane, 32);
                tc::fence_after_sync();
                // ---- epilogue: TMEM lane = gate unit.  Lanes 64..127 hold z: hand it to the r / n lanes through shared memory
                const int c0 = cg * 16;
                if (c0 < npad && qd >= 2) {
                    float z[16];
                    tc::tmem_ld16(tl + TF_ACC1 + c0, z);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 16; ++c) ZX[(c0 + c) * D + eu] = z[c];
                }
                tc::named_bar_sync(2, NWT);
                if (c0 < npad && qd < 2) {
                    float rp[16], np[16];
                    tc::tmem_ld16(tl + TF_ACC1 + c0, rp);
                    tc::tmem_ld16(tl + TF_ACC2 + c0, np);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        if (c0 + c < rows) {
                            const float r = sigmoid_fast(rp[c] + b_r);
                            const float z = sigmoid_fast(ZX[(c0 + c) * D + eu] + b_z);
                            const float n = tanh_fast(np[c] + b_in + r * b_hn);
                            p.hf[(size_t)IDS[c0 + c] * D + eu] = n - z * n;               // (1 - z) n + z h,  h = 0
                        }
                    }
                }
                tc::fence_before_sync();
                tc::named_bar_sync(1, NWT);                    // all stores of the tile issued; ZX / IDS / the node tile are free
                ++it;
                have = ti.next(nct, t0, rows);
                if (have) rr = prefetch_rows(p, t0, rows, warp, lane);
            }
            if (tid == 0 && step + 1 < nsteps) grid_arrive(p.bar);
        }
    } else if (lane == 0) {
        // ===================================================================== MMA issue + grid barrier polling (one thread)
        if (code >= 0) tc::mbar_wait_sleep(bar_w, 0u);
        uint32_t it = 0;
        for (int step = 0; step < nsteps; ++step) {
            if (code >= 0) {
                TileIter ti;
                ti.start(p.seg_ptr, step + 1, code, rank, nct);
                int t0, rows;
                while (ti.next(nct, t0, rows)) {
                    tc::mbar_wait_sleep(bar_x_full, it & 1u, 32);
                    tc::fence_after_sync();
                    issue_gate_mmas<LOWP>(sbase, sbase + F_XB_HI, sbase + F_XB_LO, tmem + TF_ACC1, tmem + TF_ACC2, (rows + 15) & ~15);
                    tc::mma_commit(bar_acc_full);
                    ++it;
                }
            }
            if (step + 1 < nsteps) {
                grid_poll(p.bar, (unsigned)(step + 1) * gridDim.x);
                tc::mbar_arrive(bar_level);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == WORKERS) tc::tmem_dealloc(tmem, TF_COLS);
}


// ======================================================================================= backward
// Reverse sweep, same tiles and the same transposed products.  Per tile (level l, rows of one gate code):
//   R  recompute: gather + attention (alpha -> HBM for the pulls) -> xbar planes -> [r | z], [n] pre-activations (tensor memory).
//      Depends on forward values only, so it runs BEFORE the level barrier is awaited (it hides behind the wait).
//   P  pull (needs the barrier): d(hs, hf)_i += sum over out-edges (i -> k) of alpha_e dxbar_k + dscore_e u_code(k); the hs half is
//      accumulated into ghs, the hf half + the incoming d hf is the GRU's output gradient g.
//   W  pointwise GRU backward (thread = gate unit, TMEM lane): d r, d z, d n -> power-of-two scaled fp16 hi/lo planes DG [node][gate]
//   MMA  dxbar^T[f][node] = Wc^T[f][g] DG^T[g][node]   (A = the forward's weight image read MN-major, B = DG K-major)
//        dWc^T[f][g]    += xbar^T[f][node] DG[node][g] (both MN-major from the node tiles; accumulator persistent in tensor memory)
//   X  dxbar^T -> shared memory [node][f] (transpose)   A  attention backward per node: d alpha_j = dxbar . x_j,
//      d score_j = alpha_j (d alpha_j - sum alpha d alpha) -> HBM with dxbar for the pulls of the predecessors; d u += sum d score_j x_j
// Nodes that are only pulled (level 0, codes without an aggregator) are handled after the last level by all CTAs.
// Weight gradients leave the kernel as raw blocks [d Wc | d u | d b_r d b_z d b_in d b_hn] (vector reductions from all CTAs of a
// code); sweep_chain_kernel maps them back to the reference's parameters (W_v, b_v, W_ih, b_ih, b_hh).
constexpr int RAWF = G3 * D2 + D2 + 4 * D;                       // 24960 floats per code
constexpr int R_WC = 0, R_U = G3 * D2, R_B = G3 * D2 + D2;
constexpr uint32_t B_XB_HI = IMG_PAD, B_XB_LO = B_XB_HI + 2 * KB_X;
constexpr uint32_t B_DG_HI = B_XB_LO + 2 * KB_X, B_DG_LO = B_DG_HI + 3 * KB_X;       // d gates [3 K blocks r z n][NT][64]
constexpr uint32_t B_ST = B_DG_LO + 3 * KB_X;                    // staging: GS [NT][64] + ZX [NT][64], later DXS [NT][128] (fp32)
constexpr uint32_t B_IDS = B_ST + NT * D2 * 4;
constexpr uint32_t B_MISC = B_IDS + NT * 4;                      // amax[2], rescale flag
constexpr uint32_t B_BAR = B_MISC + 64;                          // mbarriers
constexpr uint32_t B_TMEM = B_BAR + 128;
constexpr uint32_t B_SMEM = B_TMEM + 64 + 1024;
static_assert(B_SMEM <= 227 * 1024, "sweep backward: shared memory");
constexpr uint32_t TB_ACC1 = 0, TB_ACC2 = NT, TB_DX = 2 * NT, TB_DW = 3 * NT, TB_COLS = 512;     // DW: 192 columns

__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0u;
}
// Power of two s with amax * s in [2^8, 2^9) (amax > 0), else 1; kept while amax * cur stays inside [2^2, 2^13].
__device__ __forceinline__ float pow2_scale_keep(float amax, float cur) {
    const float v = amax * cur;
    if (v >= 4.0f && v <= 8192.0f) return cur;
    if (!(amax > 0.f) || !isfinite(amax)) return cur;
    const int e = (int)((__float_as_uint(amax) >> 23) & 0xff) - 127;
    int k = 8 - e;
    k = k < -100 ? -100 : (k > 100 ? 100 : k);
    return __uint_as_float((uint32_t)(k + 127) << 23);
}

// sum over out-edges e = (v -> k) of  alpha_e dxbar_k + dscore_e u_code(k)   (128-wide, lane chunk of 4).  Fan-out is heavy
// tailed and the per-edge chain out_pack -> out_slot -> alpha / dscore -> dxbar row is three dependent loads, so edges are taken a
// warp at a time: every lane fetches the metadata of one edge, rows are gathered 8 at a time with the ids broadcast by
// shuffles, and sum_e dscore_e u_code(e) is accumulated per code and applied once at the end.
__device__ __forceinline__ float4 pull_out_edges(const SweepTC& p, int v, int lane) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int beg = __ldg(p.out_ptr + v), end = __ldg(p.out_ptr + v + 1);
    float sds[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    unsigned present = 0u;
    for (int q0 = beg; q0 < end; q0 += 32) {
        const int q = q0 + lane;
        int kk = -1, c = 0;
        float a = 0.f, ds = 0.f;
        if (q < end) {
            const int pk = __ldg(p.out_pack + q);
            c = (pk >> MGV_CODE_SHIFT) & 7;
            if ((p.handled >> c) & 1u) {
                kk = pk & NODE_MASK;
                const int slot = __ldg(p.out_slot + q);
                a = ldcg1(p.alpha + slot);
                ds = ldcg1(p.dscore + slot);
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
                if (i + j < ne && kj >= 0) dx[j] = ldcg4(p.dxb + (size_t)kj * D2 + 4 * lane);
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

__device__ __forceinline__ void pull_into_ghs(const SweepTC& p, int node, int lane) {
    const float4 pl = pull_out_edges(p, node, lane);
    if (lane < 16) {
        float* gp = p.ghs + (size_t)node * D + 4 * lane;
        float4 cur = ldcg4(gp);
        cur.x += pl.x; cur.y += pl.y; cur.z += pl.z; cur.w += pl.w;
        mgv_st4(gp, cur);
    }
}

template <bool LOWP>
__global__ void __launch_bounds__(NTHREADS, 1) sweep_bwd_tc_kernel(const SweepTC p, const unsigned pull_only_codes) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - tc::smem_u32(smem_raw));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_w = sbase + B_BAR, bar_x_full = bar_w + 8, bar_acc_full = bar_w + 16, bar_level = bar_w + 24;
    const uint32_t bar_dg_full = bar_w + 32, bar_dx_full = bar_w + 40, bar_wg_done = bar_w + 48, bar_rescaled = bar_w + 56;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + B_TMEM);
    float* GS = reinterpret_cast<float*>(sgen + B_ST);
    float* ZX = GS + NT * D;
    float* DXS = GS;
    int* IDS = reinterpret_cast<int*>(sgen + B_IDS);
    unsigned* s_misc = reinterpret_cast<unsigned*>(sgen + B_MISC);      // [0], [1] tile amax (float bits), [2] rescale flag
    int code, rank, nct;
    find_role(p, code, rank, nct);
    if (tid == 0) {
        tc::mbar_init(bar_w, 1);
        tc::mbar_init(bar_x_full, NWT);
        tc::mbar_init(bar_acc_full, 1);
        tc::mbar_init(bar_level, 1);
        tc::mbar_init(bar_dg_full, NWT);
        tc::mbar_init(bar_dx_full, 1);
        tc::mbar_init(bar_wg_done, 1);
        tc::mbar_init(bar_rescaled, NWT);
        tc::fence_barrier_init();
        if (code >= 0) {
            const uint8_t* img = p.image + (size_t)code * IMG_PAD;
            tc::mbar_expect_tx(bar_w, IMG_BYTES);
#pragma unroll 1
            for (uint32_t o = 0; o < I_F32; o += 16384u) tc::bulk_g2s(sbase + o, img + o, 16384u, bar_w);
            tc::bulk_g2s(sbase + I_F32, img + I_F32, IMG_BYTES - I_F32, bar_w);
        }
        s_misc[0] = 0u; s_misc[1] = 0u; s_misc[2] = 0u;
    }
    if (warp == WORKERS) tc::tmem_alloc(tmem_slot, TB_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int nsteps = p.L - 1;

    if (warp < WORKERS) {
        // ===================================================================== workers
        float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const int qd = warp & 3, cg = warp >> 2;
        const int eu = (qd & 1) * 32 + lane;                   // gate unit of this thread's tensor-memory lane (phase W)
        const int ef = qd * 32 + lane;                         // input feature of this thread's lane (phase X, flush)
        float b_r = 0.f, b_z = 0.f, b_in = 0.f, b_hn = 0.f;
        if (code >= 0) {
            tc::mbar_wait_warp(bar_w, 0u, lane);
            const float* F = reinterpret_cast<const float*>(sgen + I_F32);
            u4 = *reinterpret_cast<const float4*>(F + 4 * lane);
            b_r = F[128 + eu]; b_z = F[192 + eu]; b_in = F[256 + eu]; b_hn = F[320 + eu];
        }
        const uint32_t tl = tmem + ((uint32_t)(qd * 32) << 16);
        float4 du4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float sb_r = 0.f, sb_z = 0.f, sb_n = 0.f, sb_hn = 0.f;   // bias-gradient partial sums of unit eu (threads with qd < 2)
        float acc_scale = 1.0f;
        uint32_t it = 0;
        for (int step = 0; step < nsteps; ++step) {
            const int lvl = p.L - 1 - step;
            TileIter ti;
            int t0 = 0, rows = 0;
            bool have = false, waited = (step == 0);
            if (code >= 0) {
                ti.start(p.seg_ptr, lvl, code, rank, nct);
                have = ti.next(nct, t0, rows);
            }
            while (have) {
                const int npad = (rows + 15) & ~15;
                const RowRegs rr = prefetch_rows(p, t0, rows, warp, lane);
                // the previous tile's weight-gradient MMAs have read its node tile and d-gate planes
                if (it > 0) tc::mbar_wait_warp(bar_wg_done, (it - 1) & 1u, lane, 32);
                // ---- R: recompute gather + attention (alphas kept in registers for phase A, and stored for the pulls)
                float al[NT / WORKERS][4];
#pragma unroll
                for (int k = 0; k < NT / WORKERS; ++k) {
                    const int i = warp + WORKERS * k;
#pragma unroll
                    for (int q = 0; q < 4; ++q) al[k][q] = 0.f;
                    if (i < rows) {
                        float4 xbar;
                        attend_row<true, false>(p, p.hf, u4, k, rr, lane, xbar, al[k]);
                        store_xbar<LOWP>(sbase + B_XB_HI, sbase + B_XB_LO, i, lane, xbar);
                    }
                }
#pragma unroll
                for (int k = 0; k < NT / WORKERS; ++k) {
                    const int nd = __shfl_sync(0xffffffffu, rr.node, k * 8);
                    if (lane == 0 && warp + WORKERS * k < rows) IDS[warp + WORKERS * k] = nd;
                }
                // rows rows .. npad - 1 are the K padding of the weight-gradient product: zeros
                for (int idx = tid; idx < (npad - rows) * 32; idx += NWT) {
                    const int row = rows + (idx >> 5), j = idx & 31;
                    const uint32_t off = (uint32_t)((j >> 3) & 1) * KB_X + tc::sw128_off(row, j & 7);
                    tc::st_shared_v4(sbase + ((j >> 4) ? B_XB_LO : B_XB_HI) + off, make_uint4(0u, 0u, 0u, 0u));
                }
                tc::fence_async_smem();
                tc::mbar_arrive(bar_x_full);
                // ---- the successors' levels are complete everywhere
                if (!waited) { tc::mbar_wait_warp(bar_level, (uint32_t)((step - 1) & 1), lane, 32); waited = true; }
                // ---- P: pull
                float gmax = 0.f;
#pragma unroll 1
                for (int k = 0; k < NT / WORKERS; ++k) {
                    const int i = warp + WORKERS * k;
                    if (i < rows) {
                        const int node = __shfl_sync(0xffffffffu, rr.node, k * 8);
                        const float4 pl = pull_out_edges(p, node, lane);
                        if (lane < 16) {
                            float* gp = p.ghs + (size_t)node * D + 4 * lane;
                            float4 cur = ldcg4(gp);
                            cur.x += pl.x; cur.y += pl.y; cur.z += pl.z; cur.w += pl.w;
                            mgv_st4(gp, cur);
                        } else {
                            const float4 gin = mgv_ld4(p.ghf + (size_t)node * D + 4 * (lane - 16));
                            const float4 g4 = make_float4(gin.x + pl.x, gin.y + pl.y, gin.z + pl.z, gin.w + pl.w);
                            mgv_st4(GS + i * D + 4 * (lane - 16), g4);
                            gmax = fmaxf(gmax, fmaxf(fmaxf(fabsf(g4.x), fabsf(g4.y)), fmaxf(fabsf(g4.z), fabsf(g4.w))));
                        }
                    }
                }
                // tile-wide bound of the gate gradients (|d n|, |d z|, |d r| <= max |g| for |b_hn| <= 4): fixes the power-of-two scale
                // of the d-gate planes BEFORE the pointwise pass, which then writes them straight away
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
                if (lane == 0 && gmax > 0.f) atomicMax(s_misc + (it & 1u), __float_as_uint(gmax));
                tc::mbar_wait_warp(bar_acc_full, it & 1u, lane, 32);
                tc::fence_after_sync();
                // ---- W: pointwise GRU backward.  z lives on TMEM lanes 64..127: through shared memory to the r / n lanes
                const int c0 = cg * 16;
                if (c0 < npad && qd >= 2) {
                    float z[16];
                    tc::tmem_ld16(tl + TB_ACC1 + c0, z);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 16; ++c) ZX[(c0 + c) * D + eu] = z[c];
                }
                tc::named_bar_sync(2, NWT);                    // GS (pull), ZX and the tile's gradient bound complete
                const float scale = pow2_scale_keep(__uint_as_float(s_misc[it & 1u]), acc_scale);
                const bool rescale = it > 0 && scale != acc_scale;
                if (tid == 0) { s_misc[(it + 1) & 1u] = 0u; s_misc[2] = rescale ? 1u : 0u; }
                if (c0 < npad && qd < 2) {
                    float rp[16], np[16];
                    tc::tmem_ld16(tl + TB_ACC1 + c0, rp);
                    tc::tmem_ld16(tl + TB_ACC2 + c0, np);
                    tc::tmem_ld_wait();
                    // unit pairs (eu even, eu + 1): the even lane writes the hi plane word, the odd lane the lo plane word
                    const int ue = eu & ~1;
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        float dg[3] = {0.f, 0.f, 0.f};                                  // d r, d z, d n (pre-activations)
                        if (c0 + c < rows) {
                            const float r = sigmoid_fast(rp[c] + b_r);
                            const float z = sigmoid_fast(ZX[(c0 + c) * D + eu] + b_z);
                            const float n = tanh_fast(np[c] + b_in + r * b_hn);
                            const float g = GS[(c0 + c) * D + eu];
                            const float dni = g * (1.0f - z) * (1.0f - n * n);
                            dg[2] = dni;
                            dg[0] = dni * b_hn * r * (1.0f - r);                        // n_pre = gi_n + r gh_n, gh_n = b_hn (h = 0)
                            dg[1] = -g * n * z * (1.0f - z);                            // h' = (1 - z) n + z h, h = 0
                            sb_r += dg[0]; sb_z += dg[1]; sb_n += dni; sb_hn += dni * r;
                        }
#pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            const float mine = dg[t] * scale;
                            const float other = __shfl_xor_sync(0xffffffffu, mine, 1);
                            uint32_t hi, lo;
                            split2p<LOWP>((lane & 1) ? other : mine, (lane & 1) ? mine : other, hi, lo);
                            const uint32_t off = (uint32_t)t * KB_X + tc::sw128_off(c0 + c, ue >> 3) + (uint32_t)(ue & 7) * 2u;
                            if (!(lane & 1)) st_shared_b32(sbase + B_DG_HI + off, hi);
                            else if (!LOWP) st_shared_b32(sbase + B_DG_LO + off, lo);
                        }
                    }
                }
                tc::fence_async_smem();
                tc::mbar_arrive(bar_dg_full);
                // ---- the persistent weight-gradient accumulator changes scale (rare): exact power-of-two rescale in place
                if (rescale) {
                    const float f = scale / acc_scale;
#pragma unroll 1
                    for (int cc = 0; cc < 3; ++cc) {
                        float v[16];
                        tc::tmem_ld16(tl + TB_DW + cg * 48 + 16 * cc, v);
                        tc::tmem_ld_wait();
                        uint32_t w[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) w[e] = __float_as_uint(v[e] * f);
                        tc::tmem_st16(tl + TB_DW + cg * 48 + 16 * cc, w);
                    }
                    tc::tmem_st_wait();
                    tc::fence_before_sync();
                    tc::mbar_arrive(bar_rescaled);
                }
                acc_scale = scale;
                // ---- X: dxbar^T (TMEM lane = input feature) -> [node][feature] in shared memory
                tc::mbar_wait_warp(bar_dx_full, it & 1u, lane, 32);
                tc::fence_after_sync();
                if (c0 < npad) {
                    float v[16];
                    tc::tmem_ld16(tl + TB_DX + c0, v);
                    tc::tmem_ld_wait();
                    const float inv = 1.0f / scale;
#pragma unroll
                    for (int c = 0; c < 16; ++c) DXS[(c0 + c) * D2 + ef] = v[c] * inv;
                }
                tc::fence_before_sync();
                tc::named_bar_sync(2, NWT);
                // ---- A: attention backward, warp per node
#pragma unroll
                for (int k = 0; k < NT / WORKERS; ++k) {
                    const int i = warp + WORKERS * k;
                    if (i < rows) {
                        const unsigned full = 0xffffffffu;
                        const int node = __shfl_sync(full, rr.node, k * 8), cnt = __shfl_sync(full, rr.cnt, k * 8), beg = __shfl_sync(full, rr.beg, k * 8);
                        const float4 dxb4 = *reinterpret_cast<const float4*>(DXS + i * D2 + 4 * lane);
                        mgv_st4(p.dxb + (size_t)node * D2 + 4 * lane, dxb4);
                        const int off = 4 * (lane & 15);
                        const float* base = (lane < 16) ? p.hs : p.hf;
                        if (cnt <= 4) {
                            float4 x[4];
                            float dal[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int j = __shfl_sync(full, rr.src, k * 8 + q);
                                x[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (q < cnt) x[q] = mgv_ld4(base + (size_t)j * D + off);
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) dal[q] = mgv_warp_sum(mgv_dot4(dxb4, x[q]));
                            float A = 0.f;
#pragma unroll
                            for (int q = 0; q < 4; ++q) A = fmaf(al[k][q], dal[q], A);
                            float mine = 0.f;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float ds = al[k][q] * (dal[q] - A);
                                mgv_fma4(du4, ds, x[q]);
                                if (lane == q) mine = ds;
                            }
                            if (lane < cnt) p.dscore[beg + lane] = mine;
                        } else {
                            // rare: more than four predecessors -- two passes over the rows (d alpha parked in p.dscore)
                            float A = 0.f;
                            for (int q = 0; q < cnt; ++q) {
                                const int j = __ldg(p.in_src + beg + q);
                                const float4 xq = mgv_ld4(base + (size_t)j * D + off);
                                const float dalq = mgv_warp_sum(mgv_dot4(dxb4, xq));
                                A = fmaf(p.alpha[beg + q], dalq, A);
                                if (lane == 0) p.dscore[beg + q] = dalq;
                            }
                            __syncwarp();
                            for (int q = 0; q < cnt; ++q) {
                                const int j = __ldg(p.in_src + beg + q);
                                const float4 xq = mgv_ld4(base + (size_t)j * D + off);
                                const float ds = p.alpha[beg + q] * (p.dscore[beg + q] - A);
                                mgv_fma4(du4, ds, xq);
                                __syncwarp();
                                if (lane == 0) p.dscore[beg + q] = ds;
                            }
                        }
                    }
                }
                tc::named_bar_sync(1, NWT);                    // all stores of the tile issued; the staging buffers are free
                ++it;
                have = ti.next(nct, t0, rows);
            }
            if (!waited) tc::mbar_wait_warp(bar_level, (uint32_t)((step - 1) & 1), lane, 32);    // stay in phase with the barrier
            if (tid == 0) grid_arrive(p.bar);
        }
        // ---- every level is done: nodes that are only pulled (level 0; codes without an aggregator at any level)
        if (nsteps > 0) tc::mbar_wait_warp(bar_level, (uint32_t)((nsteps - 1) & 1), lane, 32);
        {
            const int gw = blockIdx.x * WORKERS + warp, nw = gridDim.x * WORKERS;
            const int end0 = __ldg(p.seg_ptr + MGV_NCODE);
            for (int tt = gw; tt < end0; tt += nw) pull_into_ghs(p, __ldg(p.order + tt), lane);
            for (int c = 0; c < MGV_NCODE; ++c) {
                if (!((pull_only_codes >> c) & 1u)) continue;
                for (int lvl = 1; lvl < p.L; ++lvl) {
                    const int sbeg = __ldg(p.seg_ptr + lvl * MGV_NCODE + c), send = __ldg(p.seg_ptr + lvl * MGV_NCODE + c + 1);
                    for (int tt = sbeg + gw; tt < send; tt += nw) pull_into_ghs(p, __ldg(p.order + tt), lane);
                }
            }
        }
        // ---- flush: weight gradients (tensor memory), bias and attention-vector gradients (registers) -> raw block of the code
        if (code >= 0 && it > 0) {
            float* raw = p.raw + (size_t)code * RAWF;
            tc::mbar_wait_warp(bar_wg_done, (it - 1) & 1u, lane, 32);
            tc::fence_after_sync();
            const float un = 1.0f / acc_scale;
#pragma unroll 1
            for (int cc = 0; cc < 3; ++cc) {
                float v[16];
                tc::tmem_ld16(tl + TB_DW + cg * 48 + 16 * cc, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) red_add1(raw + R_WC + (size_t)(cg * 48 + 16 * cc + e) * D2 + ef, v[e] * un);
            }
            if (qd < 2) {
                red_add1(raw + R_B + eu, sb_r); red_add1(raw + R_B + D + eu, sb_z);
                red_add1(raw + R_B + 2 * D + eu, sb_n); red_add1(raw + R_B + 3 * D + eu, sb_hn);
            }
            red_add1(raw + R_U + 4 * lane, du4.x); red_add1(raw + R_U + 4 * lane + 1, du4.y);
            red_add1(raw + R_U + 4 * lane + 2, du4.z); red_add1(raw + R_U + 4 * lane + 3, du4.w);
        }
    } else if (lane == 0) {
        // ===================================================================== MMA issue + grid barrier polling (one thread)
        if (code >= 0) tc::mbar_wait_sleep(bar_w, 0u);
        uint32_t it = 0, rs_it = 0;
        bool acc_has = false;
        for (int step = 0; step < nsteps; ++step) {
            bool lvl_done = (step == 0);
            const unsigned target = (unsigned)step * gridDim.x;
            TileIter ti;
            int t0 = 0, rows = 0;
            bool have = false;
            if (code >= 0) {
                ti.start(p.seg_ptr, p.L - 1 - step, code, rank, nct);
                have = ti.next(nct, t0, rows);
            }
            while (have) {
                const int npad = (rows + 15) & ~15;
                // recompute products as soon as the node tile is full -- before or after the level barrier completes
                bool rc_done = false;
                while (!rc_done || !lvl_done) {
                    if (!rc_done && mbar_try(bar_x_full, it & 1u)) {
                        tc::fence_after_sync();
                        issue_gate_mmas<LOWP>(sbase, sbase + B_XB_HI, sbase + B_XB_LO, tmem + TB_ACC1, tmem + TB_ACC2, npad);
                        tc::mma_commit(bar_acc_full);
                        rc_done = true;
                    }
                    if (!lvl_done && mgv_ld_acquire(p.bar) >= target) {
                        __threadfence();
                        tc::mbar_arrive(bar_level);
                        lvl_done = true;
                    }
                    if (!rc_done || !lvl_done) __nanosleep(20);
                }
                tc::mbar_wait_sleep(bar_dg_full, it & 1u, 32);
                tc::fence_after_sync();
                {   // dxbar^T = Wc^T DG^T : A = weight image MN-major (M = 128 input features = the two K blocks), K = 192 gates
                    const uint32_t idesc = tc::make_idesc(128, npad, true, false, LOWP);
#pragma unroll
                    for (int s = 0; s < 12; ++s) {
                        const uint32_t go = (uint32_t)(s >> 2) * KB_X + 32u * (s & 3);
                        tc::mma3p<LOWP>(tmem + TB_DX, tc::desc_mn_sw128(sbase + I_WC_HI + 2048u * s, KB_W), tc::desc_mn_sw128(sbase + I_WC_LO + 2048u * s, KB_W),
                                        tc::desc_k_sw128(sbase + B_DG_HI + go), tc::desc_k_sw128(sbase + B_DG_LO + go), idesc, s ? 1u : 0u);
                    }
                    tc::mma_commit(bar_dx_full);
                }
                if (*reinterpret_cast<volatile unsigned*>(s_misc + 2)) {
                    tc::mbar_wait_sleep(bar_rescaled, rs_it & 1u, 32);
                    tc::fence_after_sync();
                    ++rs_it;
                }
                {   // dWc^T += xbar^T DG : both MN-major, K = nodes
                    const uint32_t idesc = tc::make_idesc(128, G3, true, true, LOWP);
                    for (int s = 0; s < npad / 16; ++s)
                        tc::mma3p<LOWP>(tmem + TB_DW, tc::desc_mn_sw128(sbase + B_XB_HI + 2048u * s, KB_X), tc::desc_mn_sw128(sbase + B_XB_LO + 2048u * s, KB_X),
                                        tc::desc_mn_sw128(sbase + B_DG_HI + 2048u * s, KB_X), tc::desc_mn_sw128(sbase + B_DG_LO + 2048u * s, KB_X), idesc,
                                        (acc_has || s) ? 1u : 0u);
                    tc::mma_commit(bar_wg_done);
                    acc_has = true;
                }
                ++it;
                have = ti.next(nct, t0, rows);
            }
            if (!lvl_done) {
                grid_poll(p.bar, target);
                tc::mbar_arrive(bar_level);
            }
        }
        if (nsteps > 0) {
            grid_poll(p.bar, (unsigned)nsteps * gridDim.x);
            tc::mbar_arrive(bar_level);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == WORKERS) tc::tmem_dealloc(tmem, TB_COLS);
}

// raw blocks -> gradient blocks in the layout of include/mgv_b200.h (chain rule of Wc = W_ih W_v, b = W_ih b_v + b_ih (+ b_hh)):
//   d W_v = W_ih^T d Wc,  d b_v = W_ih^T d c,  d W_ih = d Wc W_v^T + d c b_v^T,  d b_ih = d c,  d b_hh = [d c_r, d c_z, d b_hn],  d W_hh = 0
__global__ void sweep_chain_kernel(const float* __restrict__ pack, const float* __restrict__ raw, float* __restrict__ grads, unsigned handled) {
    const int code = blockIdx.y;
    float* G = grads + (size_t)code * GRAD;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= GRAD) return;
    if (!((handled >> code) & 1u)) { G[i] = 0.f; return; }
    const float* W = pack + (size_t)code * PACK;
    const float* R = raw + (size_t)code * RAWF;
    float v = 0.f;
    if (i < G_WV) v = R[R_U + i];
    else if (i < G_BV) {                                   // d W_v[k][f] = sum_g W_ih[g][k] d Wc[g][f]
        const int k = (i - G_WV) / D2, f = (i - G_WV) % D2;
        for (int g = 0; g < G3; ++g) v = fmaf(__ldg(W + O_WIH + g * D + k), __ldg(R + R_WC + g * D2 + f), v);
    } else if (i < G_WIH) {                                // d b_v[k] = sum_g W_ih[g][k] d c[g]
        const int k = i - G_BV;
        for (int g = 0; g < G3; ++g) v = fmaf(__ldg(W + O_WIH + g * D + k), __ldg(R + R_B + g), v);
    } else if (i < G_WHH) {                                // d W_ih[g][k] = sum_f d Wc[g][f] W_v[k][f] + d c[g] b_v[k]
        const int g = (i - G_WIH) / D, k = (i - G_WIH) % D;
        v = __ldg(R + R_B + g) * __ldg(W + O_BV + k);
        for (int f = 0; f < D2; ++f) v = fmaf(__ldg(R + R_WC + g * D2 + f), __ldg(W + O_WV + k * D2 + f), v);
    } else if (i < G_BIH) v = 0.f;                         // d W_hh: h = 0 in a single-round sweep
    else if (i < G_BHH) v = R[R_B + (i - G_BIH)];
    else {
        const int g = i - G_BHH;
        v = g < 2 * D ? R[R_B + g] : R[R_B + 3 * D + (g - 2 * D)];
    }
    G[i] = v;
}

}  // namespace sweep_tc

// ======================================================================================= host side (called from sweep.cu)
using namespace sweep_tc;

int mgv_sweep_tc_fwd(const mgv_schedule* sch, unsigned handled, const int* cta_start, int grid, const float* weights,
                     const float* hs, float* hf, int32_t* sync, int precision, cudaStream_t st) {
    SweepTC d{};
    d.N = sch->N; d.L = sch->L; d.handled = handled;
    d.order = sch->order; d.seg_ptr = sch->seg_ptr; d.in_ptr = sch->in_ptr; d.in_src = sch->in_src;
    d.out_ptr = sch->out_ptr; d.out_pack = sch->out_pack; d.out_slot = sch->out_slot;
    d.image = reinterpret_cast<const uint8_t*>(weights) + IMG_OFFSET; d.weights = weights; d.hs = hs; d.hf = hf;
    for (int c = 0; c <= MGV_NCODE; ++c) d.cta_start[c] = cta_start[c];
    d.bar = reinterpret_cast<unsigned*>(sync);
    const void* kern = precision == 1 ? (const void*)sweep_fwd_tc_kernel<true> : (const void*)sweep_fwd_tc_kernel<false>;
    MGV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM));
    void* args[] = {&d};
    MGV_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(NTHREADS), args, (size_t)F_SMEM, st));
    mgv_count_launches(1);
    return MGV_OK;
}

int mgv_sweep_tc_grid(int* grid_out) {
    int dev = 0, sms = 0, occ = 0;
    MGV_CUDA(cudaGetDevice(&dev));
    MGV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const void* kern = (const void*)sweep_fwd_tc_kernel<false>;
    MGV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM));
    MGV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHREADS, (size_t)F_SMEM));
    MGV_REQUIRE(occ >= 1, "level sweep: the tensor-core forward kernel does not fit on an SM");
    *grid_out = sms;                     // one CTA per SM (tensor memory and the weight image are per CTA)
    return MGV_OK;
}


bool mgv_sweep_tc_bwd_available() { return true; }

size_t mgv_sweep_tc_bwd_workspace_bytes(int64_t N, int64_t E) {
    size_t b = 0;
    b += mgv_align_up((size_t)N * D2 * 4 + 256, 256);          // dxb
    b += 2 * mgv_align_up((size_t)E * 4 + 256, 256);           // alpha, dscore
    b += mgv_align_up((size_t)MGV_NCODE * RAWF * 4 + 256, 256);
    return b + 1024;
}

int mgv_sweep_tc_bwd(const mgv_schedule* sch, unsigned handled, const int* cta_start, int grid, const float* weights,
                     const float* hs, const float* hf, float* ghs, float* ghf, float* grads, void* ws, size_t ws_bytes,
                     int32_t* sync, int precision, cudaStream_t st) {
    SweepTC d{};
    d.N = sch->N; d.L = sch->L; d.handled = handled;
    d.order = sch->order; d.seg_ptr = sch->seg_ptr; d.in_ptr = sch->in_ptr; d.in_src = sch->in_src;
    d.out_ptr = sch->out_ptr; d.out_pack = sch->out_pack; d.out_slot = sch->out_slot;
    d.image = reinterpret_cast<const uint8_t*>(weights) + IMG_OFFSET; d.weights = weights; d.hs = hs; d.hf = const_cast<float*>(hf);
    for (int c = 0; c <= MGV_NCODE; ++c) d.cta_start[c] = cta_start[c];
    d.bar = reinterpret_cast<unsigned*>(sync);
    MgvArena a(ws, ws_bytes);
    d.dxb = a.take<float>((size_t)sch->N * D2);
    d.alpha = a.take<float>((size_t)sch->E + 1);
    d.dscore = a.take<float>((size_t)sch->E + 1);
    d.raw = a.take<float>((size_t)MGV_NCODE * RAWF);
    MGV_REQUIRE(a.ok(), "mgv_sweep_tc_bwd: workspace too small");
    d.ghs = ghs; d.ghf = ghf;
    unsigned pull_only = 0u;                                  // codes without an aggregator that do occur at a level >= 1
    for (int c = 0; c < MGV_NCODE; ++c)
        if (!((handled >> c) & 1u) && sch->code_count[c] > 0) pull_only |= 1u << c;
    MGV_CUDA(cudaMemsetAsync(d.raw, 0, (size_t)MGV_NCODE * RAWF * sizeof(float), st));
    const void* kern = precision == 1 ? (const void*)sweep_bwd_tc_kernel<true> : (const void*)sweep_bwd_tc_kernel<false>;
    MGV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B_SMEM));
    void* args[] = {&d, &pull_only};
    MGV_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(NTHREADS), args, (size_t)B_SMEM, st));
    sweep_chain_kernel<<<dim3((GRAD + 255) / 256, MGV_NCODE), 256, 0, st>>>(weights, d.raw, grads, handled);
    mgv_count_launches(2);
    return mgv_check_cuda(cudaGetLastError(), "mgv_sweep_tc_bwd");
}
