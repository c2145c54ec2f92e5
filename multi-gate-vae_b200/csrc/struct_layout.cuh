// Shared-memory / weight-image layout and small device helpers shared by the struct-encoder tensor-core kernels
// (struct_tc.cu forward, struct_bwd_tc.cu backward).
#pragma once
#include "mgv_tc.cuh"

namespace struct_layout {

constexpr int D = MGV_D;              // 64
constexpr int G3 = 3 * D;             // 192
constexpr int TM = 128;               // nodes per tile
constexpr int SPACK = MGV_STRUCT_PACK_FLOATS;
// natural fp32 weight block (include/mgv_b200.h)
constexpr int O_WCX = 0, O_WHH = 14592, O_BC = 27648, O_BIH = 27840, O_BHH = 28032, O_LNW = 28224, O_LNB = 28288;
constexpr int LDC = 76, LDM = 68;
constexpr int NODE_MASK = (1 << MGV_CODE_SHIFT) - 1;
constexpr float LN_EPS = 1e-5f;

// ---- shared-memory / weight-image layout (bytes).  The first IMG_W bytes are the per-(encoder, direction)
// weight image prepared by struct_image_kernel; it is copied verbatim.
constexpr uint32_t WC_HI = 0, WC_LO = 24576, WHH_HI = 49152, WHH_LO = 73728, WX_HI = 98304, WX_LO = 106496;
constexpr uint32_t IMG_W = 114688;                  // weight planes
constexpr uint32_t IMG_BYTES = IMG_W + 512;         // + ln_w[64], ln_b[64] fp32
constexpr uint32_t A_AGG_HI = IMG_W, A_AGG_LO = A_AGG_HI + 16384, A_H_HI = A_AGG_LO + 16384, A_H_LO = A_H_HI + 16384;
constexpr uint32_t A_X_HI = A_H_LO + 16384, A_X_LO = A_X_HI + 4096;
constexpr uint32_t A_TILE_BYTES = A_X_LO + 4096 - A_AGG_HI;      // 73728: [agg hi | agg lo | h hi | h lo | x hi | x lo]

__device__ __forceinline__ void ldg8(const float* p, float (&v)[8]) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void stg8(float* p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
// GRU gates with 5 MUFU ops per unit (3 ex2 + 2 rcp): r and z share one reciprocal.  Pre-activations are
// clamped to +-28 / +-14 (sigmoid / tanh saturate to 1 - 1e-12 there) so the shared product cannot overflow.
__device__ __forceinline__ void gru_gates(float gr, float gz, float gi, float gh, float& r, float& z, float& n, float& hnb) {
    const float a = __expf(-fminf(fmaxf(gr, -28.f), 28.f));
    const float b = __expf(-fminf(fmaxf(gz, -28.f), 28.f));
    const float inv = __fdividef(1.0f, (1.0f + a) * (1.0f + b));
    r = (1.0f + b) * inv;
    z = (1.0f + a) * inv;
    hnb = gh;
    const float y = fminf(fmaxf(fmaf(r, gh, gi), -14.f), 14.f);
    n = 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * y));
}


}  // namespace struct_layout
