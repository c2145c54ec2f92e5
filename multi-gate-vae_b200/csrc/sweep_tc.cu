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
// A level moves ~2 KB per gate and has only hundreds to thousands of gates, so ONE level is a latency chain: barrier ->
// row gather -> dense products -> pointwise -> store -> barrier.  The design (a) shortens the chain and (b) runs several
// chains at once:
//   * STREAMS: the circuits of a batch are independent, so the schedule cuts the batch into S circuit sets ("streams",
//     mgv_schedule.streams) with their own (level, code) node lists and their own grid barrier.  A CTA runs S worker groups
//     (one per stream: 16 / S warps + one tensor-core / barrier warp each) that share the weight image; while one stream
//     waits on its barrier, its products or a gather, the warp scheduler runs the other -- the level chains overlap;
//   * every level of a stream is spread over ALL CTAs of its gate code in equal tiles (at most 64 / S rows);
//   * the products are issued TRANSPOSED -- D^T[gate unit][node] = Wc[gate unit][k] xbar^T[k][node] -- so the tensor-core
//     time of a tile is proportional to its NODE count (UMMA N = rows rounded to 16) instead of a fixed 128-row M tile;
//     the accumulator lives in tensor memory with TMEM lane = gate unit, which makes the biases per-thread constants;
//   * gathers run a HALF WARP per node (lane = 4 hs + 4 hf columns), two nodes per warp instruction, with the rows of all
//     the warp's nodes requested before the first is reduced; the static index chains (order -> in_ptr -> in_src, out_ptr ->
//     out_pack / out_slot) of the NEXT tile are fetched while the tensor core works on the current one;
//   * the grid barrier of a stream is polled by its tensor-core warp, which releases the workers through an mbarrier.
// Operands are fp16 hi/lo planes with three products per K step (mgv_tc.cuh): fp32-accurate.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <type_traits>
#include "sweep_layout.cuh"

namespace sweep_tc {
using namespace sweep_layout;

constexpr int D = MGV_D, D2 = 2 * MGV_D, G3 = 3 * MGV_D;
constexpr int PACK = MGV_SWEEP_PACK_FLOATS, GRAD = MGV_SWEEP_GRAD_FLOATS;
// natural weight block (include/mgv_b200.h)
constexpr int O_U = 0, O_BV = 8320, O_BIH = 32960, O_BHH = 33152, O_WV = 33344, O_WIH = 41536;
constexpr int G_U = 0, G_WV = 128, G_BV = 8320, G_WIH = 8384, G_WHH = 20672, G_BIH = 32960, G_BHH = 33152;
constexpr int NODE_MASK = (1 << MGV_CODE_SHIFT) - 1;

constexpr int WORKERS = 16;                     // worker warps of a CTA, split evenly over the streams
constexpr int NWT = WORKERS * 32;
constexpr int MAX_STREAMS = 2;
constexpr int BAR_STRIDE = 32;                  // ints between the grid barrier counters of two streams (own 128-byte line)

// Geometry of a CTA that serves S streams.
template <int S>
struct Geo {
    static constexpr int WPS = WORKERS / S;                // worker warps per stream
    static constexpr int NSW = WPS * 32;                   // worker threads per stream
    static constexpr int NTS = 4 * WPS;                    // rows (nodes) of a tile, at most: four per worker warp = UMMA N
    static constexpr int NTHREADS = NWT + 32 * S;          // + one tensor-core / barrier warp per stream
    static constexpr uint32_t KBX = NTS * 128u;            // one 64-column K block of a node tile (SW128)
    // forward, per stream: xbar hi / lo planes (2 K blocks each), z pre-activations [NTS][64] fp32, node ids
    static constexpr uint32_t F_XB_HI = 0, F_XB_LO = 2 * KBX, F_ZX = 4 * KBX, F_IDS = F_ZX + NTS * D * 4;
    static constexpr uint32_t F_STRIDE = (F_IDS + NTS * 4 + 1023u) & ~1023u;
    static constexpr uint32_t F_BAR = IMG_PAD + S * F_STRIDE;                   // mbarriers: w | per stream x_full, acc_full, level
    static constexpr uint32_t F_TMEM = F_BAR + 128;
    static constexpr uint32_t F_SMEM = F_TMEM + 64 + 1024;                      // + alignment slack
    static constexpr uint32_t TF_STRIDE = 2 * NTS, TF_COLS = 128;               // tensor memory per stream: [r | z] lanes, [n | -] lanes
    // backward, per stream: xbar planes, d-gate planes [3 K blocks r z n][NTS][64], staging GS + ZX / DXS (fp32), ids, misc
    static constexpr uint32_t B_XB_HI = 0, B_XB_LO = 2 * KBX, B_DG_HI = 4 * KBX, B_DG_LO = 7 * KBX, B_ST = 10 * KBX;
    static constexpr uint32_t B_IDS = B_ST + NTS * D2 * 4, B_MISC = B_IDS + NTS * 4;
    static constexpr uint32_t B_STRIDE = (B_MISC + 64 + 1023u) & ~1023u;
    static constexpr uint32_t B_BAR = IMG_PAD + S * B_STRIDE;                   // w | per stream 7 mbarriers
    static constexpr uint32_t B_TMEM = B_BAR + 256;
    static constexpr uint32_t B_SMEM = B_TMEM + 64 + 1024;
    // tensor memory per stream: [r | z] (re-used for dxbar^T once the gates are consumed), [n | -], d Wc^T (192 columns)
    static constexpr uint32_t TB_ACC1 = 0, TB_ACC2 = NTS, TB_DX = 0, TB_DW = 2 * NTS, TB_STRIDE = 256, TB_COLS = 512;
    static_assert(2 * NTS + G3 <= 256 || S == 1, "sweep backward: tensor memory per stream");
    static_assert(B_SMEM <= 227 * 1024 && F_SMEM <= 227 * 1024, "sweep: shared memory");
};

struct SweepTC {
    int N, L, S;
    unsigned handled;
    const int* order; const int* seg_ptr; const int* in_ptr; const int* in_src;
    const int* out_ptr; const int* out_pack; const int* out_slot;
    const int* desc;           // [N][8] per row of `order`: node, in_beg, in_cnt, out_beg, out_cnt, src0, src1, src2 (mgv_build_sweep_desc)
    const uint8_t* image;      // [MGV_NCODE][IMG_PAD] (behind the natural blocks in the pack buffer)
    const float* weights;      // natural blocks (u of every code, for the pulls)
    const float* hs;
    float* hf;                 // [N][64]
    int cta_start[MGV_NCODE + 1];
    int cl_size, cl_code[MGV_NCODE];   // cluster mode: CTAs per cluster (= gate codes with nodes) and the code of each cluster rank
    unsigned* bar;             // [s * BAR_STRIDE] grid barrier counter of stream s
    // backward
    float* ghs; float* ghf; float* dxb; float* alpha; float* anode; float* sds; float* raw;   // anode [N]: A = dxbar . xbar; sds [N][8]
    long long* trace;          // MGV_SWEEP_TRACE builds: [CTA][16] accumulated clock64 cycles per phase of worker thread 0
};
#ifdef MGV_SWEEP_TRACE
#define SWT_DECL(n) long long tr_acc[n] = {}, tr_last = clock64()
#define SWT(slot) do { if (p.trace && tid == 0) { const long long now_ = clock64(); tr_acc[slot] += now_ - tr_last; tr_last = now_; } } while (0)
#define SWT_FLUSH(n, tiles) do { if (p.trace && tid == 0) { for (int i_ = 0; i_ < (n); ++i_) p.trace[(size_t)blockIdx.x * 32 + i_] = tr_acc[i_]; \
                                 p.trace[(size_t)blockIdx.x * 32 + (n)] = (tiles); p.trace[(size_t)blockIdx.x * 32 + 15] = code; } } while (0)
#define SWT_SUB(slot) do { if (p.trace && threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd((unsigned long long*)p.trace + (size_t)blockIdx.x * 32 + (slot), (unsigned long long)(now_ - sub_last)); sub_last = now_; } } while (0)
#define SWT_SUB_BEGIN long long sub_last = clock64()
#else
#define SWT_DECL(n) do { } while (0)
#define SWT(slot) do { } while (0)
#define SWT_FLUSH(n, tiles) do { } while (0)
#define SWT_SUB(slot) do { } while (0)
#define SWT_SUB_BEGIN do { } while (0)
#endif

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
template <int CW>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float (&v)[CW]) {
    if constexpr (CW == 16) tc::tmem_ld16(taddr, v);
    else tc::tmem_ld8(taddr, v);
}
// Columns of a tile each column group of a stream takes in the pointwise phases: 16, or 8 when a 16-column tile is shared by
// two column groups (the phases are instruction-latency bound per thread: half the columns, half the time).
template <int NCG>
__device__ __forceinline__ int cols_per_group(int npad) { return (NCG == 2 && npad == 16) ? 8 : 16; }
__device__ __forceinline__ float half_sum(float v) {              // sum over the 16 lanes of a half warp
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float amax4(const float4& a) { return fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))); }

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

// Tiles of one (stream, level, code) segment for CTA `rank` of `nct`: the segment is cut into m * nct EQUAL tiles, m = the
// smallest count that keeps a tile within `nts` rows -- every CTA of the code gets the same share of every level.  The first
// tile moves with the level and the stream so that short segments (fewer tiles than CTAs) do not always land on the same CTAs.
struct TileIter {
    int sbeg, n, per, ntiles, k, left;
    __device__ __forceinline__ void start(const SweepTC& p, int s, int lvl, int code, int rank, int nct, int nts) {
        const int idx = (s * p.L + lvl) * MGV_NCODE + code;
        sbeg = __ldg(p.seg_ptr + idx);
        n = __ldg(p.seg_ptr + idx + 1) - sbeg;
        ntiles = 0; per = 1;
        if (n > 0) {
            const int m = (n + nct * nts - 1) / (nct * nts);
            per = (n + m * nct - 1) / (m * nct);
            ntiles = (n + per - 1) / per;
        }
        const int rot = (lvl * 37 + s * (nct >> 1)) % nct;
        k = rank - rot; if (k < 0) k += nct;
        left = 0;
    }
    __device__ __forceinline__ bool next(int nct, int& t0, int& rows) {
        if (k >= ntiles) return false;
        t0 = sbeg + k * per;
        rows = min(per, n - k * per);
        k += nct;
        return true;
    }
};

// Lane-distributed static data of the four rows worker warp `ws` owns in a tile: row i = ws + WPS * slot is held by the 8 lanes
// of slot = lane / 8, lane w = lane % 8 holding word w of the row's schedule descriptor (ONE coalesced load, no index chain):
//   0 node   1 first in-CSR position   2 fan-in   3 first out-CSR position   4 fan-out   5..7 the first three predecessors
// Out-edges (backward): lane q holds the q-th successor (out_pack = node | code << 28) and that edge's in-CSR position.
constexpr int W_NODE = 0, W_BEG = 1, W_CNT = 2, W_OBEG = 3, W_OCNT = 4, W_SRC = 5;
constexpr int FAST_FANIN = 3, FAST_FANOUT = 8;
struct RowRegs { int w; };
struct OutRegs { int cnt, pk, slot; };
template <int WPS>
__device__ __forceinline__ RowRegs prefetch_rows(const SweepTC& p, int t0, int rows, int ws, int lane) {
    RowRegs rr;
    const int i = ws + WPS * (lane >> 3), w = lane & 7;
    rr.w = w == W_NODE ? -1 : 0;
    if (i < rows) rr.w = __ldg(p.desc + (size_t)(t0 + i) * 8 + w);
    return rr;
}
__device__ __forceinline__ int row_word(const RowRegs& rr, int slot, int w) { return __shfl_sync(0xffffffffu, rr.w, slot * 8 + w); }
__device__ __forceinline__ OutRegs prefetch_out(const SweepTC& p, const RowRegs& rr, int lane) {
    OutRegs o;
    const int slot = lane >> 3, q = lane & 7;
    const int beg = row_word(rr, slot, W_OBEG);
    o.cnt = row_word(rr, slot, W_OCNT);
    o.pk = -1; o.slot = 0;
    if (q < o.cnt) { o.pk = __ldg(p.out_pack + beg + q); o.slot = __ldg(p.out_slot + beg + q); }
    return o;
}

// ---------------------------------------------------------------------------- gather + additive attention, HALF warp per node
// Lane lp = lane % 16 of a half owns columns 4 lp .. + 3 of the hs part AND of the hf part of the 128-wide row.  Both halves
// of a warp work on different nodes (slots 2 it + half) in the same instruction.  Rows with more than three predecessors
// take the whole-warp path below (attend_row_wide); the caller decides per warp.
struct HalfRow { float4 xs[FAST_FANIN], xf[FAST_FANIN]; int cnt, beg; };
template <bool HF_CG>
__device__ __forceinline__ void load_half_row(const SweepTC& p, const float* __restrict__ hf, int slot, const RowRegs& rr, int lp, HalfRow& r) {
    r.cnt = row_word(rr, slot, W_CNT);
    r.beg = row_word(rr, slot, W_BEG);
#pragma unroll
    for (int q = 0; q < FAST_FANIN; ++q) {
        const int j = row_word(rr, slot, W_SRC + q);
        r.xs[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        r.xf[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < r.cnt) {
            r.xs[q] = mgv_ld4(p.hs + (size_t)j * D + 4 * lp);
            r.xf[q] = HF_CG ? ldcg4(hf + (size_t)j * D + 4 * lp) : mgv_ld4(hf + (size_t)j * D + 4 * lp);
        }
    }
}
// softmax over the (at most three) loaded rows; xbar halves out, the attention weights in al[].
template <bool STORE_ALPHA>
__device__ __forceinline__ void attend_half(const SweepTC& p, const HalfRow& r, const float4& us, const float4& uf, int lp,
                                            float4& xbs, float4& xbf, float (&al)[FAST_FANIN]) {
    float sc[FAST_FANIN];
#pragma unroll
    for (int q = 0; q < FAST_FANIN; ++q) sc[q] = half_sum(mgv_dot4(r.xs[q], us) + mgv_dot4(r.xf[q], uf));
    float mx = -INFINITY;
#pragma unroll
    for (int q = 0; q < FAST_FANIN; ++q)
        if (q < r.cnt) mx = fmaxf(mx, sc[q]);
    float sum = 0.f;
    xbs = make_float4(0.f, 0.f, 0.f, 0.f);
    xbf = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < FAST_FANIN; ++q) {
        al[q] = 0.f;
        if (q < r.cnt) {
            al[q] = expf(sc[q] - mx);
            sum += al[q];
            mgv_fma4(xbs, al[q], r.xs[q]);
            mgv_fma4(xbf, al[q], r.xf[q]);
        }
    }
    const float inv = 1.0f / (sum + 1e-16f);
    xbs = make_float4(xbs.x * inv, xbs.y * inv, xbs.z * inv, xbs.w * inv);
    xbf = make_float4(xbf.x * inv, xbf.y * inv, xbf.z * inv, xbf.w * inv);
#pragma unroll
    for (int q = 0; q < FAST_FANIN; ++q) al[q] *= inv;
    if (STORE_ALPHA) {
        const float mine = lp == 0 ? al[0] : (lp == 1 ? al[1] : al[2]);
        if (lp < r.cnt) p.alpha[r.beg + lp] = mine;
    }
}
// The half warp's 4 + 4 values of row `row` -> fp16 hi/lo planes of a node tile (SW128, K-major; K block 0 = hs part, 1 = hf part).
template <bool LOWP>
__device__ __forceinline__ void store_xbar_half(uint32_t xb_hi, uint32_t xb_lo, uint32_t kbx, int row, int lp, const float4& vs, const float4& vf) {
    const uint32_t off = tc::sw128_off(row, lp >> 1) + (uint32_t)(lp & 1) * 8u;
    uint32_t h0, l0, h1, l1;
    split2p<LOWP>(vs.x, vs.y, h0, l0);
    split2p<LOWP>(vs.z, vs.w, h1, l1);
    st_shared_v2(xb_hi + off, h0, h1);
    if (!LOWP) st_shared_v2(xb_lo + off, l0, l1);
    split2p<LOWP>(vf.x, vf.y, h0, l0);
    split2p<LOWP>(vf.z, vf.w, h1, l1);
    st_shared_v2(xb_hi + kbx + off, h0, h1);
    if (!LOWP) st_shared_v2(xb_lo + kbx + off, l0, l1);
}

// Whole-warp variant for any fan-in (online softmax past four predecessors; predecessor ids from the in-CSR).  Lane l owns columns 4 (l % 16) .. + 3 of the hs part
// (l < 16) or the hf part (l >= 16).  `al` receives the first four attention weights.
template <bool STORE_ALPHA, bool HF_CG>
__device__ __forceinline__ void attend_row_wide(const SweepTC& p, const float* __restrict__ hf, const float4& u4, int slot, const RowRegs& rr,
                                                int lane, float4& xbar, float (&al)[4]) {
    const int cnt = row_word(rr, slot, W_CNT), beg = row_word(rr, slot, W_BEG);
    const int off = 4 * (lane & 15);
    const float* base = (lane < 16) ? p.hs : hf;
    float4 x[4];
    float sc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        x[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < cnt) {
            const int j = __ldg(p.in_src + beg + q);
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
        // more than four predecessors -- online softmax over the rest, one at a time (scores parked in p.alpha)
        if (STORE_ALPHA && lane < 4) p.alpha[beg + lane] = sc[lane];
        for (int q = 4; q < cnt; ++q) {
            const int j = __ldg(p.in_src + beg + q);
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
template <bool LOWP>
__device__ __forceinline__ void store_xbar_wide(uint32_t xb_hi, uint32_t xb_lo, uint32_t kbx, int row, int lane, const float4& v) {
    uint32_t h0, l0, h1, l1;
    split2p<LOWP>(v.x, v.y, h0, l0);
    split2p<LOWP>(v.z, v.w, h1, l1);
    const uint32_t off = (uint32_t)(lane >> 4) * kbx + tc::sw128_off(row, (lane & 15) >> 1) + (uint32_t)(lane & 1) * 8u;
    st_shared_v2(xb_hi + off, h0, h1);
    if (!LOWP) st_shared_v2(xb_lo + off, l0, l1);
}

// Gather + attention of a whole tile by one worker warp (its four rows) -> xbar planes and node ids.  Rows npad > i >= rows are
// written as exact zeros (the K padding of the backward's weight-gradient product).  al2[it][q]: attention weights of the row
// of slot 2 it + half (fast path only; the wide path of the backward re-reads p.alpha).  Returns true when the wide path ran
// (some row of the warp has more than FAST_FANIN predecessors).
template <bool LOWP, bool STORE_ALPHA, bool HF_CG, int WPS>
__device__ __forceinline__ bool gather_tile(const SweepTC& p, const float* __restrict__ hf, const float* F32,
                                            const RowRegs& rr, int rows, int npad, int ws, int lane, uint32_t xb_hi, uint32_t xb_lo,
                                            uint32_t kbx, int* IDS, float (&al2)[2][FAST_FANIN]) {
    const unsigned full = 0xffffffffu;
    const bool wide = __any_sync(full, (lane & 7) == W_CNT && rr.w > FAST_FANIN);
    if (!wide) {
        const int half = lane >> 4, lp = lane & 15;
        const bool second = npad > 2 * WPS;                // rows 2 WPS .. exist (slots 2, 3): warp-uniform
        HalfRow r0, r1;
        load_half_row<HF_CG>(p, hf, half, rr, lp, r0);
        if (second) load_half_row<HF_CG>(p, hf, 2 + half, rr, lp, r1);
        // the attention vector comes from the weight image in shared memory each time (registers are the scarce resource here)
        const float4 us = *reinterpret_cast<const float4*>(F32 + 4 * lp), uf = *reinterpret_cast<const float4*>(F32 + D + 4 * lp);
        float4 xbs, xbf;
        attend_half<STORE_ALPHA>(p, r0, us, uf, lp, xbs, xbf, al2[0]);
        int i = ws + WPS * half;
        if (i < npad) store_xbar_half<LOWP>(xb_hi, xb_lo, kbx, i, lp, xbs, xbf);
        if (second) {
            attend_half<STORE_ALPHA>(p, r1, us, uf, lp, xbs, xbf, al2[1]);
            i = ws + WPS * (2 + half);
            if (i < npad) store_xbar_half<LOWP>(xb_hi, xb_lo, kbx, i, lp, xbs, xbf);
        }
    } else {
        const float4 u4 = *reinterpret_cast<const float4*>(F32 + 4 * lane);      // lane l of the wide layout owns columns 4 l .. + 3 of the 128
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const int i = ws + WPS * k;
            if (i < npad) {
                float4 xbar;
                float al[4];
                attend_row_wide<STORE_ALPHA, HF_CG>(p, hf, u4, k, rr, lane, xbar, al);
                store_xbar_wide<LOWP>(xb_hi, xb_lo, kbx, i, lane, xbar);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int nd = row_word(rr, k, W_NODE);
        if (lane == 0 && ws + WPS * k < rows) IDS[ws + WPS * k] = nd;
    }
    return wide;
}

// The recompute / forward products of one tile: acc1 = Wc[0:128] xbar^T, acc2 = Wc[128:256] xbar^T (rows 192.. are whatever
// follows the image: those accumulator lanes are never read).
template <bool LOWP>
__device__ __forceinline__ void issue_gate_mmas(uint32_t sbase, uint32_t xb_hi, uint32_t xb_lo, uint32_t kbx, uint32_t acc1, uint32_t acc2, int npad) {
    const uint32_t idesc = tc::make_idesc(128, npad, false, false, LOWP);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t wo = (uint32_t)kb * KB_W + 32u * j, xo = (uint32_t)kb * kbx + 32u * j;
            const uint64_t b_hi = tc::desc_k_sw128(xb_hi + xo), b_lo = tc::desc_k_sw128(xb_lo + xo);
            const uint32_t accf = (kb | j) ? 1u : 0u;
            tc::mma3p<LOWP>(acc1, tc::desc_k_sw128(sbase + I_WC_HI + wo), tc::desc_k_sw128(sbase + I_WC_LO + wo), b_hi, b_lo, idesc, accf);
            tc::mma3p<LOWP>(acc2, tc::desc_k_sw128(sbase + I_WC_HI + wo + 16384u), tc::desc_k_sw128(sbase + I_WC_LO + wo + 16384u), b_hi, b_lo,
                            idesc, accf);
        }
}

// Grid barrier of a stream, split: one worker thread arrives (release), the stream's tensor-core warp polls (acquire) and
// releases the workers through a shared-memory mbarrier.
__device__ __forceinline__ void grid_arrive(unsigned* counter) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0u;
}
// Whole warp: lane 0 polls the counter and, once the target is reached, releases the workers; the result is warp-uniform.
__device__ __forceinline__ bool grid_try(const unsigned* counter, unsigned target, uint32_t bar_level, int lane) {
    unsigned ok = 0u;
    if (lane == 0) {
        ok = mgv_ld_acquire(counter) >= target ? 1u : 0u;
        if (ok) { __threadfence(); tc::mbar_arrive(bar_level); }
    }
    return __shfl_sync(0xffffffffu, ok, 0) != 0u;
}

// Cluster barrier, split (all threads of all CTAs of the cluster execute both): release / acquire at cluster scope cover the
// global-memory rows the CTAs of a cluster hand to each other between levels.
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// ======================================================================================= forward
// CL = false: grid mode -- S (1 or 2) circuit-set streams, every level of a stream spread over all CTAs of its gate code, one grid
//             barrier per stream and level (cooperative launch).  For batches of few large circuits.
// CL = true : cluster mode -- the batch is cut into many circuit sets (mgv_schedule.streams > 2); a thread-block CLUSTER of one
//             CTA per gate code walks whole streams on its own, levels separated by the hardware cluster barrier: no grid-wide
//             synchronisation at all, clusters never wait for each other.  For batches of many small circuits (the reference's
//             training batches), where a level is a fraction of a tile per CTA and the grid barrier is 40 % of a level.
template <bool LOWP, int S, bool CL>
__global__ void __launch_bounds__(Geo<S>::NTHREADS, 1) sweep_fwd_tc_kernel(const SweepTC p) {
    static_assert(!CL || S == 1, "cluster mode runs one stream at a time per CTA");
    using G = Geo<S>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - tc::smem_u32(smem_raw));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_w = sbase + G::F_BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + G::F_TMEM);
    int code, rank, nct;
    int s_first = 0, s_step = 1, s_end = 1;                 // schedule streams this CTA walks (grid mode: its warps' own one)
    if constexpr (CL) {
        code = p.cl_code[blockIdx.x % p.cl_size]; rank = 0; nct = 1;
        s_first = blockIdx.x / p.cl_size; s_step = gridDim.x / p.cl_size; s_end = p.S;
    } else {
        find_role(p, code, rank, nct);
    }
    if (tid == 0) {
        tc::mbar_init(bar_w, 1);
        for (int s = 0; s < S; ++s) {
            tc::mbar_init(bar_w + 8 + 24 * s, G::NSW);     // x_full
            tc::mbar_init(bar_w + 16 + 24 * s, 1);         // acc_full
            tc::mbar_init(bar_w + 24 + 24 * s, 1);         // level
        }
        tc::fence_barrier_init();
        if (code >= 0) {
            const uint8_t* img = p.image + (size_t)code * IMG_PAD;
            tc::mbar_expect_tx(bar_w, IMG_BYTES);
#pragma unroll 1
            for (uint32_t o = 0; o < I_F32; o += 16384u) tc::bulk_g2s(sbase + o, img + o, 16384u, bar_w);
            tc::bulk_g2s(sbase + I_F32, img + I_F32, IMG_BYTES - I_F32, bar_w);
        }
    }
    if (warp == WORKERS) tc::tmem_alloc(tmem_slot, G::TF_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int nsteps = p.L - 1;
    const int s = CL ? 0 : (warp < WORKERS ? warp / G::WPS : warp - WORKERS);    // stream slot (resources) of this warp
    if constexpr (!CL) { s_first = s; s_end = s + 1; }
    const uint32_t bar_x_full = bar_w + 8 + 24 * s, bar_acc_full = bar_x_full + 8, bar_level = bar_x_full + 16;
    const uint32_t sstream = sbase + IMG_PAD + (uint32_t)s * G::F_STRIDE;
    const uint32_t xb_hi = sstream + G::F_XB_HI, xb_lo = sstream + G::F_XB_LO;
    const uint32_t acc1 = tmem + (uint32_t)s * G::TF_STRIDE, acc2 = acc1 + G::NTS;
    unsigned* gbar = p.bar + s * BAR_STRIDE;

    if (warp < WORKERS) {
        // ===================================================================== workers of stream slot s: gather -> [MMA] -> epilogue
        const int ws = warp % G::WPS;
        float* ZX = reinterpret_cast<float*>(sgen + IMG_PAD + s * G::F_STRIDE + G::F_ZX);
        int* IDS = reinterpret_cast<int*>(sgen + IMG_PAD + s * G::F_STRIDE + G::F_IDS);
        const int qd = warp & 3, cg = ws >> 2;
        const int eu = (qd & 1) * 32 + lane;                   // gate unit of this thread's tensor-memory lane
        const float* F = reinterpret_cast<const float*>(sgen + I_F32);        // fp32 tail of the weight image: u | b_r | b_z | b_in | b_hn
        float b_r = 0.f, b_z = 0.f, b_in = 0.f, b_hn = 0.f;
        if (code >= 0) {
            tc::mbar_wait_warp(bar_w, 0u, lane);
            b_r = F[128 + eu]; b_z = F[192 + eu]; b_in = F[256 + eu]; b_hn = F[320 + eu];
        }
        const uint32_t tl1 = acc1 + ((uint32_t)(qd * 32) << 16), tl2 = acc2 + ((uint32_t)(qd * 32) << 16);
        uint32_t it = 0;
        SWT_DECL(8);
        for (int sx = s_first; sx < s_end; sx += s_step) {     // schedule stream (grid mode: one; cluster mode: every s_step-th)
        TileIter ti;
        int t0 = 0, rows = 0;
        bool have = false;
        RowRegs rr;
        rr.w = (lane & 7) == W_NODE ? -1 : 0;
        if (code >= 0 && nsteps > 0) {
            ti.start(p, sx, 1, code, rank, nct, G::NTS);
            have = ti.next(nct, t0, rows);
            if (have) rr = prefetch_rows<G::WPS>(p, t0, rows, ws, lane);
        }
        for (int step = 0; step < nsteps; ++step) {
            SWT(0);
            if (step > 0) {                                    // level `step` of this stream is complete everywhere
                if constexpr (CL) cluster_wait();
                else tc::mbar_wait_warp(bar_level, (uint32_t)((step - 1) & 1), lane, 0);
            }
            SWT(1);
            // the next tile of this CTA (same level, else the first of the next level): its static data is fetched while the
            // tensor core works on the current tile
            int t0n = 0, rowsn = 0;
            bool haven = false, crossed = false;
            RowRegs rrn = rr;
            while (have) {
                const int npad = (rows + 15) & ~15;
                float al2[2][FAST_FANIN];
                gather_tile<LOWP, false, true, G::WPS>(p, p.hf, F, rr, rows, npad, ws, lane, xb_hi, xb_lo, G::KBX, IDS, al2);
                tc::fence_async_smem();
                tc::mbar_arrive(bar_x_full);
                SWT(2);
                haven = ti.next(nct, t0n, rowsn);
                if (!haven && !crossed && step + 1 < nsteps) {
                    ti.start(p, sx, step + 2, code, rank, nct, G::NTS);
                    crossed = true;
                    haven = ti.next(nct, t0n, rowsn);
                    if (haven) rrn = prefetch_rows<G::WPS>(p, t0n, rowsn, ws, lane);
                    have = false;                              // leave the level after this tile
                } else if (haven) {
                    rrn = prefetch_rows<G::WPS>(p, t0n, rowsn, ws, lane);
                } else {
                    have = false;
                }
                tc::mbar_wait_warp(bar_acc_full, it & 1u, lane, 0);
                tc::fence_after_sync();
                SWT(3);
                // ---- epilogue: TMEM lane = gate unit.  Lanes 64..127 hold z: those warps hand 1 - z to the r / n lanes through shared
                // memory.  All columns are computed branch-free first (independent dependency chains per thread), then stored.
                auto epilogue = [&](auto cw_tag) {
                    constexpr int CW = decltype(cw_tag)::value;
                    const int c0 = cg * CW;
                    if (c0 < npad && qd >= 2) {
                        float z[CW];
                        tmem_ld_cols<CW>(tl1 + c0, z);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < CW; ++c) z[c] = 1.0f - sigmoid_fast(z[c] + b_z);
#pragma unroll
                        for (int c = 0; c < CW; ++c) ZX[(c0 + c) * D + eu] = z[c];
                    }
                    tc::named_bar_sync(2 + 2 * s, G::NSW);
                    if (c0 < npad && qd < 2) {
                        float rp[CW], np[CW];
                        tmem_ld_cols<CW>(tl1 + c0, rp);
                        tmem_ld_cols<CW>(tl2 + c0, np);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < CW; ++c) {
                            const float r = sigmoid_fast(rp[c] + b_r);
                            const float n = tanh_fast(np[c] + b_in + r * b_hn);
                            rp[c] = n * ZX[(c0 + c) * D + eu];                            // (1 - z) n + z h,  h = 0
                        }
#pragma unroll
                        for (int c = 0; c < CW; ++c)
                            if (c0 + c < rows) p.hf[(size_t)IDS[c0 + c] * D + eu] = rp[c];
                    }
                };
                if (cols_per_group<G::WPS / 4>(npad) == 8) epilogue(std::integral_constant<int, 8>{});
                else epilogue(std::integral_constant<int, 16>{});
                tc::fence_before_sync();
                SWT(4);
                tc::named_bar_sync(1 + 2 * s, G::NSW);         // all stores of the tile issued; ZX / IDS / the node tile are free
                SWT(5);
                ++it;
                if (have) { t0 = t0n; rows = rowsn; rr = rrn; }
            }
            if (!crossed && step + 1 < nsteps && code >= 0) {  // no tile of this CTA at this level
                ti.start(p, sx, step + 2, code, rank, nct, G::NTS);
                haven = ti.next(nct, t0n, rowsn);
                if (haven) rrn = prefetch_rows<G::WPS>(p, t0n, rowsn, ws, lane);
            }
            if (step + 1 < nsteps) {
                if constexpr (CL) { __syncwarp(); cluster_arrive(); }
                else if (ws == 0 && lane == 0) grid_arrive(gbar);
            }
            have = haven && step + 1 < nsteps;
            t0 = t0n; rows = rowsn; rr = rrn;
            SWT(6);
        }
        }
        SWT_FLUSH(8, it);
    } else {
        // ===================================================================== tensor-core issue + grid barrier polling of stream s:
        // the whole warp runs the control flow (descriptors stay in uniform registers), one elected lane issues
        if (code >= 0) tc::mbar_wait_warp(bar_w, 0u, lane);
        uint32_t it = 0;
        for (int sx = s_first; sx < s_end; sx += s_step)
        for (int step = 0; step < nsteps; ++step) {
            if (code >= 0) {
                TileIter ti;
                ti.start(p, sx, step + 1, code, rank, nct, G::NTS);
                int t0, rows;
                while (ti.next(nct, t0, rows)) {
                    tc::mbar_wait_warp(bar_x_full, it & 1u, lane, 0);
                    tc::fence_after_sync();
                    if (tc::elect_one()) {
                        issue_gate_mmas<LOWP>(sbase, xb_hi, xb_lo, G::KBX, acc1, acc2, (rows + 15) & ~15);
                        tc::mma_commit(bar_acc_full);
                    }
                    __syncwarp();
                    ++it;
                }
            }
            if (step + 1 < nsteps) {
                if constexpr (CL) {
                    __syncwarp();
                    cluster_arrive();
                    cluster_wait();
                } else {
                    const unsigned target = (unsigned)(step + 1) * gridDim.x;
                    tc::WaitGuard wg;
                    while (!grid_try(gbar, target, bar_level, lane)) { __nanosleep(20); wg.tick(); }
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == WORKERS) tc::tmem_dealloc(tmem, G::TF_COLS);
}

#include "sweep_tc_bwd.inc"

// d u_code = sum over ALL nodes v of sds[v][code - 1] x_v, x_v = [hs_v || hf_v]  (sds: the per-code sums of dscore over v's out-edges,
// written by the backward's pulls) -> added to the raw blocks.  One streaming pass over the embeddings, warp per node, lane = 4 columns.
__global__ void __launch_bounds__(256) sweep_du_kernel(const float* __restrict__ sds, const float* __restrict__ hs, const float* __restrict__ hf,
                                                       int n, float* __restrict__ raw, unsigned handled) {
    __shared__ float red[8][6][D2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 acc[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* base = lane < 16 ? hs : hf;
    const int off = 4 * (lane & 15);
    for (int v = blockIdx.x * 8 + warp; v < n; v += gridDim.x * 8) {
        const float4 s0 = mgv_ldg4(sds + (size_t)v * 8), s1 = mgv_ldg4(sds + (size_t)v * 8 + 4);
        const float sv[6] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y};
        bool any = false;
#pragma unroll
        for (int c = 0; c < 6; ++c) any |= sv[c] != 0.f;
        if (!any) continue;
        const float4 x = mgv_ldg4(base + (size_t)v * D + off);
#pragma unroll
        for (int c = 0; c < 6; ++c) mgv_fma4(acc[c], sv[c], x);
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) *reinterpret_cast<float4*>(&red[warp][c][4 * lane]) = acc[c];
    __syncthreads();
    for (int i = threadIdx.x; i < 6 * D2; i += 256) {
        const int c = i / D2;
        if (!((handled >> (c + 1)) & 1u)) continue;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][c][i % D2];
        if (t != 0.f) red_add1(raw + (size_t)(c + 1) * RAWF + R_U + (i % D2), t);
    }
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
#pragma unroll 32
        for (int g = 0; g < G3; ++g) v = fmaf(__ldg(W + O_WIH + g * D + k), __ldg(R + R_WC + g * D2 + f), v);
    } else if (i < G_WIH) {                                // d b_v[k] = sum_g W_ih[g][k] d c[g]
        const int k = i - G_BV;
#pragma unroll 32
        for (int g = 0; g < G3; ++g) v = fmaf(__ldg(W + O_WIH + g * D + k), __ldg(R + R_B + g), v);     // 64 threads: keep the loads in flight
    } else if (i < G_WHH) {                                // d W_ih[g][k] = sum_f d Wc[g][f] W_v[k][f] + d c[g] b_v[k]
        const int g = (i - G_WIH) / D, k = (i - G_WIH) % D;
        v = __ldg(R + R_B + g) * __ldg(W + O_BV + k);
#pragma unroll 32
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
long long* mgv_debug_trace();

namespace {
int fill(SweepTC& d, const mgv_schedule* sch, unsigned handled, const int* cta_start, const float* weights, const float* hs, float* hf,
         int32_t* sync) {
    d.N = sch->N; d.L = sch->L; d.S = sch->streams > 1 ? sch->streams : 1; d.handled = handled;
    d.order = sch->order; d.seg_ptr = sch->seg_ptr; d.in_ptr = sch->in_ptr; d.in_src = sch->in_src;
    d.out_ptr = sch->out_ptr; d.out_pack = sch->out_pack; d.out_slot = sch->out_slot;
    MGV_REQUIRE(sch->sweep_desc != nullptr, "level sweep: the schedule has no row descriptors (mgv_build_sweep_desc)");
    d.desc = sch->sweep_desc;
    d.image = reinterpret_cast<const uint8_t*>(weights) + IMG_OFFSET; d.weights = weights; d.hs = hs; d.hf = hf;
    for (int c = 0; c <= MGV_NCODE; ++c) d.cta_start[c] = cta_start[c];
    d.bar = reinterpret_cast<unsigned*>(sync);
    d.trace = nullptr;
    return MGV_OK;
}
template <int S, bool CL>
const void* fwd_kernel(int precision) {
    return precision == 1 ? (const void*)sweep_fwd_tc_kernel<true, S, CL> : (const void*)sweep_fwd_tc_kernel<false, S, CL>;
}
// Cluster mode: one CTA per gate code that has nodes; as many clusters as fit (at most one per stream).
int cluster_setup(SweepTC& d, const mgv_schedule* sch, unsigned handled) {
    d.cl_size = 0;
    for (int c = 0; c < MGV_NCODE; ++c)
        if (((handled >> c) & 1u) && sch->code_count[c] > 0) d.cl_code[d.cl_size++] = c;
    return d.cl_size;
}
int cluster_launch(const void* kern, int threads, size_t smem, int cl_size, int streams, void** args, cudaStream_t st) {
    MGV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cl_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(cl_size);
    int fit = 0;
    MGV_CUDA(cudaOccupancyMaxActiveClusters(&fit, kern, &cfg));
    MGV_REQUIRE(fit >= 1, "level sweep: a cluster of %d CTAs does not fit on the device", cl_size);
    const int clusters = streams < fit ? streams : fit;
    cfg.gridDim = dim3(clusters * cl_size);
    MGV_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
    return MGV_OK;
}
template <int S, bool CL>
const void* bwd_kernel(int precision) {
    return precision == 1 ? (const void*)sweep_bwd_tc_kernel<true, S, CL> : (const void*)sweep_bwd_tc_kernel<false, S, CL>;
}
}  // namespace

int mgv_sweep_tc_fwd(const mgv_schedule* sch, unsigned handled, const int* cta_start, int grid, const float* weights,
                     const float* hs, float* hf, int32_t* sync, int precision, cudaStream_t st) {
    SweepTC d{};
    int rc = fill(d, sch, handled, cta_start, weights, hs, hf, sync);
    if (rc != MGV_OK) return rc;
#ifdef MGV_SWEEP_TRACE
    d.trace = getenv("MGV_TRACE_SWEEP_FWD") ? mgv_debug_trace() : nullptr;
#endif
    void* args[] = {&d};
    if (d.S > MAX_STREAMS) {                              // many circuit sets: clusters walk them independently
        if (cluster_setup(d, sch, handled) == 0) return MGV_OK;
        rc = cluster_launch(fwd_kernel<1, true>(precision), Geo<1>::NTHREADS, (size_t)Geo<1>::F_SMEM, d.cl_size, d.S, args, st);
        if (rc != MGV_OK) return rc;
        mgv_count_launches(1);
        return MGV_OK;
    }
    const void* kern = d.S == 2 ? fwd_kernel<2, false>(precision) : fwd_kernel<1, false>(precision);
    const size_t smem = d.S == 2 ? (size_t)Geo<2>::F_SMEM : (size_t)Geo<1>::F_SMEM;
    const int threads = d.S == 2 ? Geo<2>::NTHREADS : Geo<1>::NTHREADS;
    MGV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MGV_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(threads), args, smem, st));
    mgv_count_launches(1);
    return MGV_OK;
}

int mgv_sweep_tc_grid(int* grid_out) {
    int dev = 0, sms = 0, occ = 0;
    MGV_CUDA(cudaGetDevice(&dev));
    MGV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const void* kern = bwd_kernel<2, false>(0);
    MGV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<2>::B_SMEM));
    MGV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Geo<2>::NTHREADS, (size_t)Geo<2>::B_SMEM));
    MGV_REQUIRE(occ >= 1, "level sweep: the tensor-core kernels do not fit on an SM");
    *grid_out = sms;                     // one CTA per SM (tensor memory and the weight image are per CTA)
    return MGV_OK;
}

bool mgv_sweep_tc_bwd_available() { return true; }

size_t mgv_sweep_tc_bwd_workspace_bytes(int64_t N, int64_t E) {
    size_t b = 0;
    b += mgv_align_up((size_t)N * D2 * 4 + 256, 256);          // dxb
    b += mgv_align_up((size_t)E * 4 + 256, 256);               // alpha
    b += mgv_align_up((size_t)N * 4 + 256, 256);               // A per node
    b += mgv_align_up((size_t)N * 8 * 4 + 256, 256);           // sds per node and code
    b += mgv_align_up((size_t)MGV_NCODE * RAWF * 4 + 256, 256);
    return b + 1024;
}

int mgv_sweep_tc_bwd(const mgv_schedule* sch, unsigned handled, const int* cta_start, int grid, const float* weights,
                     const float* hs, const float* hf, float* ghs, float* ghf, float* grads, void* ws, size_t ws_bytes,
                     int32_t* sync, int precision, cudaStream_t st) {
    SweepTC d{};
    int rc = fill(d, sch, handled, cta_start, weights, hs, const_cast<float*>(hf), sync);
    if (rc != MGV_OK) return rc;
    MgvArena a(ws, ws_bytes);
    d.dxb = a.take<float>((size_t)sch->N * D2);
    d.alpha = a.take<float>((size_t)sch->E + 1);
    d.anode = a.take<float>((size_t)sch->N + 1);
    d.sds = a.take<float>((size_t)sch->N * 8 + 8);
    d.raw = a.take<float>((size_t)MGV_NCODE * RAWF);
    MGV_REQUIRE(a.ok(), "mgv_sweep_tc_bwd: workspace too small");
    d.ghs = ghs; d.ghf = ghf;
#ifdef MGV_SWEEP_TRACE
    d.trace = getenv("MGV_TRACE_SWEEP_FWD") ? nullptr : mgv_debug_trace();
#endif
    unsigned pull_only = 0u;                                  // codes without an aggregator that do occur at a level >= 1
    for (int c = 0; c < MGV_NCODE; ++c)
        if (!((handled >> c) & 1u) && sch->code_count[c] > 0) pull_only |= 1u << c;
    MGV_CUDA(cudaMemsetAsync(d.raw, 0, (size_t)MGV_NCODE * RAWF * sizeof(float), st));
    MGV_CUDA(cudaMemsetAsync(d.sds, 0, (size_t)sch->N * 8 * sizeof(float), st));      // every node is pulled exactly once and writes its row; belt and braces
    void* args[] = {&d, &pull_only};
    if (d.S > MAX_STREAMS) {                              // many circuit sets: clusters walk them independently
        MGV_REQUIRE(cluster_setup(d, sch, handled) > 0, "level sweep backward: no gate code to propagate");
        rc = cluster_launch(bwd_kernel<1, true>(precision), Geo<1>::NTHREADS, (size_t)Geo<1>::B_SMEM, d.cl_size, d.S, args, st);
        if (rc != MGV_OK) return rc;
    } else {
        const void* kern = d.S == 2 ? bwd_kernel<2, false>(precision) : bwd_kernel<1, false>(precision);
        const size_t smem = d.S == 2 ? (size_t)Geo<2>::B_SMEM : (size_t)Geo<1>::B_SMEM;
        const int threads = d.S == 2 ? Geo<2>::NTHREADS : Geo<1>::NTHREADS;
        MGV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MGV_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(threads), args, smem, st));
    }
    {
        int sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int blocks = (int)std::min<int64_t>((int64_t)sms * 4, ((int64_t)sch->N + 7) / 8);
        if (blocks > 0) sweep_du_kernel<<<blocks, 256, 0, st>>>(d.sds, hs, hf, sch->N, d.raw, handled);
    }
    sweep_chain_kernel<<<dim3((GRAD + 255) / 256, MGV_NCODE), 256, 0, st>>>(weights, d.raw, grads, handled);
    mgv_count_launches(3);
    return mgv_check_cuda(cudaGetLastError(), "mgv_sweep_tc_bwd");
}
