// Weight-image layout of the tensor-core level sweep (sweep_tc.cu), shared with the packing kernel (pack.cu).
#pragma once
#include "mgv_tc.cuh"

namespace sweep_layout {
// weight image of one gate code (bytes): Wc = W_ih W_v as two 64-column K blocks of 192 rows (SW128, K-major), hi and lo
// planes, then fp32 u[128] | b_r[64] | b_z[64] | b_in[64] | b_hn[64]  (b = W_ih b_v + b_ih (+ b_hh for r, z); b_hn = b_hh[n])
constexpr uint32_t KB_W = 3 * MGV_D * 128;      // 24576
constexpr uint32_t I_WC_HI = 0, I_WC_LO = 2 * KB_W, I_F32 = 4 * KB_W;
constexpr uint32_t IMG_BYTES = I_F32 + 384 * 4; // 99840
constexpr uint32_t IMG_PAD = 100352;            // next multiple of 1024
// the buffer mgv_sweep_pack fills: natural blocks [MGV_NCODE][MGV_SWEEP_PACK_FLOATS] fp32, then the images [MGV_NCODE][IMG_PAD]
constexpr size_t IMG_OFFSET = (size_t)MGV_NCODE * MGV_SWEEP_PACK_FLOATS * sizeof(float);
constexpr size_t PACK_TOTAL_BYTES = IMG_OFFSET + (size_t)MGV_NCODE * IMG_PAD;
}  // namespace sweep_layout
