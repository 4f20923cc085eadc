// The 16-bit storage type of activations and weights is a build-time choice: the same sources build
// libspegnet_b200_fp16.so (-DSPG_FP16, IEEE half: 10-bit mantissa) and libspegnet_b200_bf16.so (bfloat16:
// 7-bit mantissa).  Both run kind::f16 tcgen05 MMAs at the same rate with fp32 accumulation; only the
// operand format bits, the pack / unpack conversions and the mma.sync type differ.
#pragma once
#include <cstdint>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace spg {

#ifdef SPG_FP16
using h16 = __half;
using h162 = __half2;
constexpr int kHalfIsFp16 = 1;
#define SPG_MMA_TYPE "f16"
#else
using h16 = __nv_bfloat16;
using h162 = __nv_bfloat162;
constexpr int kHalfIsFp16 = 0;
#define SPG_MMA_TYPE "bf16"
#endif

// Every 16-bit store saturates to the largest finite value (F2FP.SATFINITE, same single instruction): an fp16
// activation beyond 65504 clamps instead of becoming inf and then NaN in the next LayerNorm / softmax
// (tests/test_gpu_ops.py::test_fp16_stores_saturate, tests/test_gpu_model.py::test_large_activations_stay_finite).
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    uint32_t r;
#ifdef SPG_FP16
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
#else
    asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
#endif
    return r;
}
__device__ __forceinline__ float h_lo(uint32_t u) {
#ifdef SPG_FP16
    return __half2float(__ushort_as_half(static_cast<unsigned short>(u & 0xFFFFu)));
#else
    return __uint_as_float(u << 16);
#endif
}
__device__ __forceinline__ float h_hi(uint32_t u) {
#ifdef SPG_FP16
    return __half2float(__ushort_as_half(static_cast<unsigned short>(u >> 16)));
#else
    return __uint_as_float(u & 0xFFFF0000u);
#endif
}
__device__ __forceinline__ float h_to_float(uint16_t u) { return h_lo(u); }
__device__ __forceinline__ float h_to_float(h16 v) {
#ifdef SPG_FP16
    return __half2float(v);
#else
    return __bfloat162float(v);
#endif
}
__device__ __forceinline__ h16 float_to_h(float f) {
    unsigned short r;
#ifdef SPG_FP16
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(f));
    return __ushort_as_half(r);
#else
    asm("cvt.rn.satfinite.bf16.f32 %0, %1;" : "=h"(r) : "f"(f));
    return __ushort_as_bfloat16(r);
#endif
}
__device__ __forceinline__ uint32_t max_h2(uint32_t a, uint32_t b) {
    h162 x = *reinterpret_cast<h162*>(&a), y = *reinterpret_cast<h162*>(&b);
    h162 m = __hmax2(x, y);
    return *reinterpret_cast<uint32_t*>(&m);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = h_lo(u.x); f[1] = h_hi(u.x); f[2] = h_lo(u.y); f[3] = h_hi(u.y);
    f[4] = h_lo(u.z); f[5] = h_hi(u.z); f[6] = h_lo(u.w); f[7] = h_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}

// ---- packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE operations per issue slot, same roundings as the
// scalar forms).  A pair lives in a 64-bit register.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_make(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_split(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// one 16-byte vector of 8 halves <-> four fp32 pairs
__device__ __forceinline__ void unpack8_f2(const uint4& u, f32x2 (&f)[4]) {
    f[0] = f2_make(h_lo(u.x), h_hi(u.x));
    f[1] = f2_make(h_lo(u.y), h_hi(u.y));
    f[2] = f2_make(h_lo(u.z), h_hi(u.z));
    f[3] = f2_make(h_lo(u.w), h_hi(u.w));
}
__device__ __forceinline__ uint32_t pack_f2(f32x2 v) {
    float lo, hi;
    f2_split(v, lo, hi);
    return pack2(lo, hi);
}
__device__ __forceinline__ uint4 pack8_f2(const f32x2 (&f)[4]) { return make_uint4(pack_f2(f[0]), pack_f2(f[1]), pack_f2(f[2]), pack_f2(f[3])); }

}  // namespace spg
