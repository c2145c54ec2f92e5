// tcgen05 (5th-generation tensor core) building blocks for the mgv_b200 kernels, sm_100a only.
//
// Numerics: every dense contraction of the hot path runs as  fp16 x fp16 -> fp32 (TMEM accumulator)
// with BOTH operands split into two fp16 planes, x = hi + lo (hi = rn_f16(x), lo = rn_f16(x - hi)),
// and three MMAs per K step: hi.hi + lo.hi + hi.lo.  The dropped lo.lo term is 2^-22 relative, so the
// result is fp32-accurate (measured against the fp64 reference: same error as plain fp32).
// fp16 instead of TF32 because (a) the planes cost 4 bytes per element in shared memory, like ONE fp32
// copy -- weights + operand tiles of a 128-node tile fit in 227 KB, TF32 hi/lo planes (8 B) do not --
// and (b) kind::f16 issues at twice the kind::tf32 rate.  Gradient operands are scaled by a power of
// two into fp16's range before the split (see mgv_grad_scale) and un-scaled in the epilogue.
//
// Shared-memory operand tiles are written by ordinary threads (the gather phases) directly in the
// canonical UMMA layouts, so the same tile can be consumed K-major (rows = M/N index) or MN-major
// (rows = K index) by a different descriptor -- forward and transposed (weight-gradient) products
// share one copy:
//   SW128 tile  : rows of 128 bytes (64 fp16), 8-row groups of 1024 bytes, 16-byte chunk index XOR (row & 7)
//   plain tile  : rows of 32 bytes (16 fp16) as 8x16-byte core matrices (no swizzle), used for the small
//                 [x | deg | 1] block
// Descriptor bit layouts follow the PTX ISA "matrix descriptor" / "instruction descriptor" tables
// (restated in CUTLASS cute/arch/mma_sm100_desc.hpp).
#pragma once
#include <cuda_fp16.h>
#include "mgv_common.cuh"

namespace tc {

// ------------------------------------------------------------------------------------------ addresses
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Byte offset of 16-byte chunk `c16` (0..7) of row `r` inside a SW128 tile (any number of rows).
__device__ __forceinline__ uint32_t sw128_off(int r, int c16) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}
// Byte offset of chunk `c16` (0..1) of row `r` inside a plain (no-swizzle) 16-column tile.
__device__ __forceinline__ uint32_t plain16_off(int r, int c16) {
    return (uint32_t)((r >> 3) * 256 + c16 * 128 + (r & 7) * 16);
}

// ------------------------------------------------------------------------------------------ fp16 split
// Two floats -> packed hi pair and lo pair (element 0 in the low half-word).  Saturating converts:
// |x| > 65504 clamps (hi) and the remainder clamps again (lo) instead of producing inf.
__device__ __forceinline__ uint32_t pack_f16x2_sat(float e0, float e1) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
    return r;
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_f16x2_sat(a, b);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    lo = pack_f16x2_sat(a - hf.x, b - hf.y);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float e0, float e1) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
    return r;
}
// bf16 mode: one plane of bf16 operands (the lo plane is not written / not multiplied).
template <bool LOWP>
__device__ __forceinline__ void split8p(const float (&v)[8], uint4& hi, uint4& lo) {
    if (LOWP) {
        hi.x = pack_bf16x2(v[0], v[1]); hi.y = pack_bf16x2(v[2], v[3]); hi.z = pack_bf16x2(v[4], v[5]); hi.w = pack_bf16x2(v[6], v[7]);
        lo = make_uint4(0u, 0u, 0u, 0u);
    } else {
        split2(v[0], v[1], hi.x, lo.x); split2(v[2], v[3], hi.y, lo.y); split2(v[4], v[5], hi.z, lo.z); split2(v[6], v[7], hi.w, lo.w);
    }
}
// 8 consecutive K elements -> one 16-byte chunk in each plane.
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
    split2(v[0], v[1], hi.x, lo.x);
    split2(v[2], v[3], hi.y, lo.y);
    split2(v[4], v[5], hi.z, lo.z);
    split2(v[6], v[7], hi.w, lo.w);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
// Bounded waits: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.  The bound is wall-clock time
// (%globaltimer, checked every 1024 polls), not a poll count, so kernels slowed down 10-100x by a profiler or
// compute-sanitizer still run to completion.
constexpr unsigned long long WAIT_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct WaitGuard {
    unsigned long long t0 = 0ull;
    uint32_t polls = 0u;
    __device__ __forceinline__ void tick() {
        if ((++polls & 1023u) == 0u) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0ull) t0 = now;
            else if (now - t0 > WAIT_TIMEOUT_NS) __trap();
        }
    }
};
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0u;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    WaitGuard g;
#pragma unroll 1
    while (!mbar_try_wait(bar, parity)) g.tick();
}

// Whole-warp wait that keeps the issue slots of the SM sub-partition free for the warps doing work: lane 0 polls
// with a sleep between attempts, the other lanes park at the warp barrier.  (A 32-lane try_wait spin loop in the
// waiting roles starves the epilogue warp sharing their scheduler: measured ~8 cycles per instruction.)
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int lane, unsigned sleep_ns = 64) {
    if (lane == 0) {
        WaitGuard g;
#pragma unroll 1
        while (!mbar_try_wait(bar, parity)) {
            if (sleep_ns) __nanosleep(sleep_ns);           // 0: try_wait itself suspends the thread for a hardware-defined time
            g.tick();
        }
    }
    __syncwarp();
}
// Single-thread variant (the MMA-issue thread).
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, unsigned sleep_ns = 64) {
    WaitGuard g;
#pragma unroll 1
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(sleep_ns);
        g.tick();
    }
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
// Bulk asynchronous copy global -> shared (TMA engine, no tensor map); completes `bytes` on `bar`.
// 16-byte aligned addresses, size a multiple of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Bulk asynchronous copy shared -> global (TMA engine), tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
// L2 eviction-priority hints for data that is written once and read once by the NEXT kernel (the struct-backward hand-off
// buffers): evict-first keeps the 126 MB L2 for the rows the gathers re-read.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_s2g_hint(void* dst, uint32_t src, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_v4_hint(void* dst, const uint4& v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed groups of this thread have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... have completed (global writes performed)
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}

// ------------------------------------------------------------------------------------------ fences
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ TMEM
// One full warp allocates `ncols` (power of two >= 32) columns; the base address lands in *slot (shared).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
// The calling warp's 32 TMEM lanes (lane field of taddr = 32 * (warp % 4)), 32 / 16 / 8 consecutive columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Registers -> the calling warp's 32 TMEM lanes, 16 / 8 / 4 consecutive columns (asynchronous: tmem_st_wait before
// the data is read back or handed to another thread).
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st8f(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ descriptors
constexpr uint32_t LAYOUT_NONE = 0, LAYOUT_SW128 = 2;
// Shared-memory matrix descriptor: start address, leading / stride byte offsets (all >> 4), version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
// K-major SW128 tile (rows = M or N index): 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr) { return make_desc(saddr, 16, 1024, LAYOUT_SW128); }
// The same kind of tile read MN-major (rows = K index): `mn_block_bytes` between 64-element blocks along M/N.
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t mn_block_bytes) {
    return make_desc(saddr, mn_block_bytes, 1024, LAYOUT_SW128);
}
// Plain 16-column tile, K-major: core matrices 128 bytes apart along K, 8-row groups 256 bytes apart.
__device__ __forceinline__ uint64_t desc_k_plain16(uint32_t saddr) { return make_desc(saddr, 128, 256, LAYOUT_NONE); }
// ... read MN-major (rows = K index): 8-element M/N chunks 128 bytes apart (SBO), 8-row K groups 256 bytes apart (LBO).
__device__ __forceinline__ uint64_t desc_mn_plain16(uint32_t saddr) { return make_desc(saddr, 256, 128, LAYOUT_NONE); }

// Instruction descriptor, kind::f16, fp16 (or bf16) operands, fp32 accumulate.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn_major, bool b_mn_major, bool bf16 = false) {
    return (1u << 4)                                  // D format f32
           | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10)      // A, B format: 0 = f16, 1 = bf16
           | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16)
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows = lanes, 16 fp16 of K = 8 packed 32-bit columns per
// K step) lives in tensor memory, written there by tcgen05.st.
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
template <bool LOWP>
__device__ __forceinline__ void mma3p_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                         uint32_t idesc, uint32_t accumulate) {
    mma_ts(d_tmem, a_hi, b_hi, idesc, accumulate);
    if (!LOWP) {
        mma_ts(d_tmem, a_lo, b_hi, idesc, 1u);
        mma_ts(d_tmem, a_hi, b_lo, idesc, 1u);
    }
}
// hi.hi + lo.hi + hi.lo for one K step; `step` = descriptor start-address increment (bytes >> 4 applied here).
__device__ __forceinline__ void mma3(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                     uint32_t idesc, uint32_t accumulate) {
    mma(d_tmem, a_hi, b_hi, idesc, accumulate);
    mma(d_tmem, a_lo, b_hi, idesc, 1u);
    mma(d_tmem, a_hi, b_lo, idesc, 1u);
}
template <bool LOWP>
__device__ __forceinline__ void mma3p(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                      uint32_t idesc, uint32_t accumulate) {
    if (LOWP) mma(d_tmem, a_hi, b_hi, idesc, accumulate);
    else mma3(d_tmem, a_hi, a_lo, b_hi, b_lo, idesc, accumulate);
}
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// One lane of a fully converged warp (the elect.sync leader): tcgen05.mma / commit are issued under this predicate while the
// whole warp runs the surrounding control flow, so descriptors stay in uniform registers (a lane == 0 branch makes every
// operand a per-thread value: ptxas then wraps each UTCHMMA in an R2UR + ELECT loop, ~60 cycles per instruction).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0u;
}

// All previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// First index t in [0, n] with a[t] - adj * t >= target, the adjusted sequence non-decreasing with its last element
// >= target.  Whole warp, ~log32(n) dependent loads instead of log2(n).  `adj` lowers the per-tile fixed cost of the
// tile cost model for kernels whose per-neighbour work is larger (the backward gathers two rows per neighbour).
__device__ __forceinline__ int warp_lower_bound(const unsigned* __restrict__ a, int n, unsigned target, int lane, unsigned adj = 0u) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int span = hi - lo;
        if (span <= 32) {
            const int m = lo + lane;
            const bool ge = (m < hi) ? (a[m] - adj * (unsigned)m >= target) : true;
            const unsigned bal = __ballot_sync(0xffffffffu, ge);
            // span == 32 with no entry >= target: every lane probed a real element, the ballot is empty -> hi
            return bal == 0u ? hi : lo + __ffs(bal) - 1;
        }
        const int m = lo + (int)(((long long)span * (lane + 1)) / 33);
        const bool ge = a[m] - adj * (unsigned)m >= target;
        const unsigned bal = __ballot_sync(0xffffffffu, ge);
        if (bal == 0u) {
            lo = lo + (int)(((long long)span * 32) / 33) + 1;
        } else {
            const int f = __ffs(bal) - 1;
            const int nlo = f ? lo + (int)(((long long)span * f) / 33) + 1 : lo;
            hi = lo + (int)(((long long)span * (f + 1)) / 33);
            lo = nlo;
        }
    }
    return lo;
}

}  // namespace tc
