// Row statistics of the LayerNorm that a residual GEMM applies to its own output (gemm_tcgen05.cu, kLn == 3) -- and of
// the stand-alone kernel that reproduces them BIT FOR BIT (pointwise.cu, layernorm_sliced_kernel), so that the host may
// choose per launch which of the two runs (fused for large batches, separate in the latency regime) without the result
// depending on that choice.  Both kernels therefore share these functions and the slice table below.
//
// A row of N channels is cut into slices (one per (n-tile, epilogue column split) of the GEMM tiling for that N, which
// is fixed per N and never a function of M).  Per slice, sequentially in column order with shift = the slice's first
// value:  s1 = sum (v - shift),  s2 = sum (v - shift)^2  ->  {mean_i, M2_i};  the slices are merged in column order with
// Chan's formula.
#pragma once
#include <cuda_runtime.h>

namespace spg {

constexpr int kLnMaxSlices = 6;

struct LnSlices {
    int count;
    int bound[kLnMaxSlices + 1];  // slice q = columns [bound[q], bound[q+1])
};

// Slice table of an N-channel row (host): the GEMM tiling (block_n = the largest multiple of 16 <= 192 that divides N
// into <= 3 tiles) and the split of a tile's 16-column chunks between the two warps of a TMEM lane quarter.
// Returns count == 0 when N has no such tiling.
inline LnSlices ln_slices_for(int N) {
    LnSlices s{};
    int block_n = 0;
    for (int bn = 192; bn >= 16; bn -= 16)
        if (N % bn == 0 && N / bn <= kLnMaxSlices / 2) {
            block_n = bn;
            break;
        }
    if (block_n == 0) return s;
    const int chunks = block_n >> 4;
    const int first = (chunks + 1) / 2;  // chunks of the first column split (group = 1, two splits)
    int q = 0;
    for (int n0 = 0; n0 < N; n0 += block_n) {
        s.bound[q++] = n0;
        s.bound[q++] = n0 + first * 16;
    }
    s.bound[q] = N;
    s.count = q;
    return s;
}

#ifdef __CUDACC__
// accumulate one value into a slice's shifted sums
__device__ __forceinline__ void ln_accumulate(float v, float shift, float& s1, float& s2) {
    const float d = v - shift;
    s1 += d;
    s2 = fmaf(d, d, s2);
}
// {mean, M2} of a slice of n values from its shifted sums
__device__ __forceinline__ float2 ln_slice_stats(float shift, float s1, float s2, float n) {
    return make_float2(shift + s1 / n, fmaxf(s2 - s1 * s1 / n, 0.f));
}
// merge `count` slices (stats(q) -> {mean, M2}, cols(q) -> number of values) into {rstd, -mean * rstd}
template <typename StatsFn, typename ColsFn>
__device__ __forceinline__ float2 ln_merge(int count, StatsFn stats, ColsFn cols, float inv_cols, float eps) {
    float mean = 0.f;
    for (int q = 0; q < count; ++q) mean = fmaf(cols(q), stats(q).x, mean);
    mean *= inv_cols;
    float m2 = 0.f;
    for (int q = 0; q < count; ++q) {
        const float2 st = stats(q);
        const float d = st.x - mean;
        m2 += fmaf(cols(q) * d, d, st.y);
    }
    const float rstd = rsqrtf(fmaf(m2, inv_cols, eps));
    return make_float2(rstd, -mean * rstd);
}
// y = (v - mean) * rstd * gamma + beta with rm = {rstd, -mean * rstd}
__device__ __forceinline__ float ln_normalise(float v, float2 rm, float gamma, float beta) {
    return fmaf(fmaf(v, rm.x, rm.y), gamma, beta);
}
#endif

}  // namespace spg
