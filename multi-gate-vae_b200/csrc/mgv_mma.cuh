// Warp-level tensor-core tiles with fp32 accuracy: mma.sync m16n8k8 TF32 with the 3xTF32 split done
// in registers (a = a_hi + a_lo, b = b_hi + b_lo; d += a_lo b_hi + a_hi b_lo + a_hi b_hi).
// Operands are ordinary fp32 tiles in SHARED memory (ldmatrix for the row-major cases), read straight into
// MMA fragments -- either orientation of a matrix can be consumed from ONE copy, which is what lets
// a gate code's / encoder's weights stay resident in shared memory once (fp32, 108-138 KB) next to
// the activation tiles.  (A tcgen05 path would need hi and lo copies of every weight matrix in the
// canonical smem layout -- 2x the bytes -- and does not fit next to the tiles; it is the natural
// path for the bf16 configuration, see DESIGN.md.)
//
// Fragment ownership (PTX ISA, mma.m16n8k8 .tf32): g = lane / 4, t = lane % 4
//   A (16x8, row): a0 (g, t)   a1 (g+8, t)   a2 (g, t+4)   a3 (g+8, t+4)
//   B (8x8,  col): b0 (k = t, n = g)   b1 (k = t+4, n = g)
//   C (16x8)     : c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1)
#pragma once
#include "mgv_common.cuh"

__device__ __forceinline__ uint32_t mgv_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// hi = x with the 13 low mantissa bits cleared (exactly what the tensor core reads from a 32-bit
// operand), lo = x - hi (exact in fp32; the hardware truncates it to tf32 again: residual 2^-21).
__device__ __forceinline__ void mgv_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mgv_ldmatrix_x4(const float* p, uint32_t (&r)[4]) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mgv_ldmatrix_x2(const float* p, uint32_t (&r)[2]) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void mgv_mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// A(m, k): row-major A[m * lda + k], or (TRANS) A[k * lda + m]
template <bool TRANS>
__device__ __forceinline__ void mgv_load_a(const float* A, int lda, int m0, int k0, int g, int t,
                                           uint32_t (&hi)[4], uint32_t (&lo)[4]) {
    float v0, v1, v2, v3;
    if (!TRANS) {
        // one ldmatrix: an 8x8 b16 matrix row is 16 bytes = 4 tf32; lanes 8i..8i+7 address matrix i
        // (0: rows 0-7 / k 0-3, 1: rows 8-15 / k 0-3, 2: rows 0-7 / k 4-7, 3: rows 8-15 / k 4-7)
        const int lane = g * 4 + t, mi = lane >> 3, r = lane & 7;
        uint32_t v[4];
        mgv_ldmatrix_x4(A + (m0 + (mi & 1) * 8 + r) * lda + k0 + (mi >> 1) * 4, v);
        v0 = __uint_as_float(v[0]); v1 = __uint_as_float(v[1]); v2 = __uint_as_float(v[2]); v3 = __uint_as_float(v[3]);
    } else {
        const float* p = A + (k0 + t) * lda + m0 + g;
        v0 = p[0]; v1 = p[8]; v2 = p[4 * lda]; v3 = p[4 * lda + 8];
    }
    mgv_split(v0, hi[0], lo[0]); mgv_split(v1, hi[1], lo[1]);
    mgv_split(v2, hi[2], lo[2]); mgv_split(v3, hi[3], lo[3]);
}
// B(k, n): (NK) B[n * ldb + k]  -- e.g. an nn.Linear weight [out][in] used as y = x W^T --
//          or (!NK) B[k * ldb + n]
template <bool NK>
__device__ __forceinline__ void mgv_load_b(const float* B, int ldb, int k0, int n0, int g, int t,
                                           uint32_t (&hi)[2], uint32_t (&lo)[2]) {
    float v0, v1;
    if (NK) {
        const int lane = g * 4 + t, mi = (lane >> 3) & 1, r = lane & 7;   // lanes 16..31 repeat valid addresses
        uint32_t v[2];
        mgv_ldmatrix_x2(B + (n0 + r) * ldb + k0 + mi * 4, v);
        v0 = __uint_as_float(v[0]); v1 = __uint_as_float(v[1]);
    } else {
        const float* p = B + (k0 + t) * ldb + n0 + g;
        v0 = p[0]; v1 = p[4 * ldb];
    }
    mgv_split(v0, hi[0], lo[0]); mgv_split(v1, hi[1], lo[1]);
}

// c[mt][nt] += A[m0 + 16 mt .. , 0..8 KSTEPS) * B[0..8 KSTEPS, n0[nt] ..)   for one warp.
template <int MT, int NT, int KSTEPS, bool A_TRANS, bool B_NK>
__device__ __forceinline__ void mgv_warp_gemm(float (&c)[MT][NT][4], const float* A, int lda, int m0,
                                              const float* B, int ldb, const int (&n0)[NT], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (int ks = 0; ks < KSTEPS; ++ks) {
        const int k0 = ks * 8;
        uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) mgv_load_a<A_TRANS>(A, lda, m0 + 16 * mt, k0, g, t, ahi[mt], alo[mt]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            uint32_t bhi[2], blo[2];
            mgv_load_b<B_NK>(B, ldb, k0, n0[nt], g, t, bhi, blo);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                mgv_mma_tf32(c[mt][nt], alo[mt], bhi);
                mgv_mma_tf32(c[mt][nt], ahi[mt], blo);
                mgv_mma_tf32(c[mt][nt], ahi[mt], bhi);
            }
        }
    }
}

template <int MT, int NT>
__device__ __forceinline__ void mgv_zero_frag(float (&c)[MT][NT][4]) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) { c[i][j][0] = 0.f; c[i][j][1] = 0.f; c[i][j][2] = 0.f; c[i][j][3] = 0.f; }
}
