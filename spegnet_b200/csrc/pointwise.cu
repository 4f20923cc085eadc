// HBM-bound kernels of the SPEGNet forward: everything that is not a GEMM / conv / attention.
// All are vectorised (8 or 16 bytes per access), coalesced over the channel (innermost NHWC) axis and
// keep their arithmetic in fp32.  Entry points are declared in include/spegnet_b200.h.
#include <atomic>

#include "common.h"
#include "half16.cuh"
#include "ln_stats.cuh"

namespace spg {
extern std::atomic<long long> g_launches;

namespace {

#define SPG_LAUNCHED()                                        \
    do {                                                      \
        g_launches.fetch_add(1, std::memory_order_relaxed);   \
        SPG_CHECK_LAUNCH();                                   \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the channel axis: fp32 residual stream in, bf16 GEMM operand out. One warp per token,
// the row lives in registers (two-pass variance), C <= 1152.
// ------------------------------------------------------------------------------------------------
constexpr int kLnMaxC = 1152;

// LPR = lanes cooperating on one row (32, or 16 for C <= 256 so that a 144-channel row does not idle half a warp);
// each warp walks over `rows_per_warp` consecutive row groups so that a block moves >= 64 KB.
template <int LPR, int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, uint16_t* __restrict__ y,
                                                        int M, int C, float eps, int rows_per_warp, int reverse) {
    pdl_prologue();
    constexpr int RPW = 32 / LPR;  // rows processed concurrently by one warp
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane % LPR, grp = lane / LPR;
    const int nvec = C >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    const unsigned bid = reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;  // see common.h "Traversal direction"
    const long long first = (static_cast<long long>(bid) * 8 + warp) * rows_per_warp * RPW;
    for (int it = 0; it < rows_per_warp; ++it) {
        const long long row = first + it * RPW + grp;
        const bool ok = row < M;
        const float4* xr = reinterpret_cast<const float4*>(x + row * C);
        float4 v[MAXV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int j = sub + LPR * i;
            if (ok && j < nvec) {
                v[i] = __ldcs(xr + j);  // streaming: this is the last read of the fp32 row before the block's GEMMs; keep L2 for y
                s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int j = sub + LPR * i;
            if (ok && j < nvec) {
                const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
                q += (a * a + b * b) + (c * c + d * d);
            }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q / C + eps);
        uint2* yr = reinterpret_cast<uint2*>(y + row * C);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int j = sub + LPR * i;
            if (ok && j < nvec) {
                const float4 g = __ldg(g4 + j), b = __ldg(b4 + j);
                yr[j] = make_uint2(pack2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y),
                                   pack2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Patchify: NCHW fp32 image -> im2col rows for the 7x7 / stride 4 / pad 3 patch embedding (bf16).
// Column order k = (ky*3 + c)*8 + kx (kx = 0..6, slot 7 is zero): 168 columns.  The 7 taps of one (ky, c) sit at
// ix = 4*ox - 3 .. 4*ox + 3, i.e. elements 1..7 of the 16-byte aligned pair x[4*ox - 4 .. 4*ox + 3]: one thread =
// two float4 loads + one 16-byte store, with no per-element index arithmetic (550 -> 220 us at batch 64).  The weight
// matrix is packed in the same column order on the host (model.py).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ x, uint4* __restrict__ cols, int B, int S) {
    pdl_prologue();
    const int G = S >> 2;
    const long long total = static_cast<long long>(B) * G * G * 21;
    for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += gridDim.x * 256ll) {
        const int kc = idx % 21;  // ky*3 + c
        const long long row = idx / 21;
        const int ox = row % G, oy = (row / G) % G, b = row / (static_cast<long long>(G) * G);
        const int ky = kc / 3, c = kc - ky * 3;
        const int iy = oy * 4 + ky - 3;
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (iy >= 0 && iy < S) {
            const float4* line = reinterpret_cast<const float4*>(x + ((static_cast<size_t>(b) * 3 + c) * S + iy) * S);
            const float4 hi = __ldg(line + ox);  // ix = 4*ox .. 4*ox + 3
            f[3] = hi.x; f[4] = hi.y; f[5] = hi.z; f[6] = hi.w;
            if (ox > 0) {
                const float4 lo = __ldg(line + ox - 1);  // ix = 4*ox - 4 .. 4*ox - 1 (the first is not a tap)
                f[0] = lo.y; f[1] = lo.z; f[2] = lo.w;
            }
        }
        cols[idx] = pack8(f);
    }
}

// ------------------------------------------------------------------------------------------------
// 2x2 / stride 2 max pool on an fp32 NHWC map (the pooled shortcut of blocks 2 / 8 / 44).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_kernel(const float4* __restrict__ x, float4* __restrict__ y, int B, int H,
                                                      int W, int C4) {
    pdl_prologue();
    const int Ho = H >> 1, Wo = W >> 1;
    const long long total = static_cast<long long>(B) * Ho * Wo * C4;
    for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += gridDim.x * 256ll) {
        const int c = idx % C4;
        const long long p = idx / C4;
        const int ox = p % Wo, oy = (p / Wo) % Ho, b = p / (static_cast<long long>(Wo) * Ho);
        const size_t base = ((static_cast<size_t>(b) * H + 2 * oy) * W + 2 * ox) * C4 + c;
        const float4 a = x[base], bb = x[base + C4], cc = x[base + static_cast<size_t>(W) * C4],
                     d = x[base + static_cast<size_t>(W) * C4 + C4];
        y[idx] = make_float4(fmaxf(fmaxf(a.x, bb.x), fmaxf(cc.x, d.x)), fmaxf(fmaxf(a.y, bb.y), fmaxf(cc.y, d.y)),
                             fmaxf(fmaxf(a.z, bb.z), fmaxf(cc.z, d.z)), fmaxf(fmaxf(a.w, bb.w), fmaxf(cc.w, d.w)));
    }
}

// fp32 -> bf16 cast, 8 elements per thread.
__global__ void __launch_bounds__(256) cast_kernel(const float4* __restrict__ x, uint4* __restrict__ y, long long n8) {
    pdl_prologue();
    for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < n8; idx += gridDim.x * 256ll) {
        const float4 a = x[2 * idx], b = x[2 * idx + 1];
        y[idx] = make_uint4(pack2(a.x, a.y), pack2(a.z, a.w), pack2(b.x, b.y), pack2(b.z, b.w));
    }
}

// ------------------------------------------------------------------------------------------------
// Bilinear (align_corners=False) source coordinates exactly as ATen's upsample_bilinear2d:
// src = max(0, (dst + 0.5) * in/out - 0.5), i1 = min(i0 + 1, in - 1).
// ------------------------------------------------------------------------------------------------
struct Lerp {
    int i0, i1;
    float w0, w1;
};
__device__ __forceinline__ Lerp lerp_coord(int dst, int in, int out) {
    const float scale = static_cast<float>(in) / static_cast<float>(out);
    float src = (dst + 0.5f) * scale - 0.5f;
    src = src < 0.f ? 0.f : src;
    Lerp l;
    l.i0 = static_cast<int>(src);
    l.i1 = l.i0 + (l.i0 < in - 1 ? 1 : 0);
    l.w1 = src - l.i0;
    l.w0 = 1.f - l.w1;
    return l;
}

// out[b,y,x,:] = concat(bilinear(src0)[C0], bilinear(src1)[C1]) in bf16 NHWC; src1 may be absent (C1 = 0).
// One block per PAIR of output rows; a thread produces a 2x2 block of output pixels for one 8-channel vector from a
// 3x3 patch of its source (for an integer scale >= 2 the four outputs read at most 3 source rows x 3 source columns):
// 2.25 16-byte loads per 16-byte store instead of 4, and one set of interpolation coordinates per 4 outputs
// (measured: 2.1 -> 2.9 TB/s of output).  Arithmetic order per output is ATen's: wy0 (wx0 a + wx1 b) + wy1 (wx0 c + wx1 d).
__global__ void __launch_bounds__(256)
upcat_kernel(const uint4* __restrict__ s0, int h0, int w0, int c0v, const uint4* __restrict__ s1, int h1, int w1,
             int c1v, uint4* __restrict__ out, int Ho, int Wo) {
    pdl_prologue();
    const int cv = c0v + c1v;
    const int by = blockIdx.x, b = blockIdx.y;
    const int half_w = Wo >> 1;
    const int total = half_w * cv;
    uint4* orow0 = out + (static_cast<size_t>(b) * Ho + 2 * by) * Wo * cv;
    uint4* orow1 = orow0 + static_cast<size_t>(Wo) * cv;
    for (int t = threadIdx.x; t < total; t += 256) {
        const int bx = t / cv;
        const int c = t - bx * cv;
        const uint4* src;
        int h, w, ncv, cc;
        if (c < c0v) {
            src = s0; h = h0; w = w0; ncv = c0v; cc = c;
        } else {
            src = s1; h = h1; w = w1; ncv = c1v; cc = c - c0v;
        }
        const Lerp ya = lerp_coord(2 * by, h, Ho), yb = lerp_coord(2 * by + 1, h, Ho);
        const Lerp xa = lerp_coord(2 * bx, w, Wo), xb = lerp_coord(2 * bx + 1, w, Wo);
        // patch rows {ya.i0, ya.i1, yb.i1} and columns {xa.i0, xa.i1, xb.i1}; yb.i0 / xb.i0 coincide with one of the first two
        const int rows[3] = {ya.i0, ya.i1, yb.i1};
        const int cols[3] = {xa.i0, xa.i1, xb.i1};
        const size_t img = static_cast<size_t>(b) * h * w;
        float p[3][3][8];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q)
                unpack8(__ldg(src + (img + static_cast<size_t>(rows[r]) * w + cols[q]) * ncv + cc), p[r][q]);
        const bool yb0_first = yb.i0 == ya.i0, xb0_first = xb.i0 == xa.i0;
        float o[2][2][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            // horizontal interpolation of the three patch rows at the two output columns
            float hx[3][2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                hx[r][0] = xa.w0 * p[r][0][i] + xa.w1 * p[r][1][i];
                const float left = xb0_first ? p[r][0][i] : p[r][1][i];
                hx[r][1] = xb.w0 * left + xb.w1 * p[r][2][i];
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                o[0][q][i] = ya.w0 * hx[0][q] + ya.w1 * hx[1][q];
                const float top = yb0_first ? hx[0][q] : hx[1][q];
                o[1][q][i] = yb.w0 * top + yb.w1 * hx[2][q];
            }
        }
        const size_t off = static_cast<size_t>(2 * bx) * cv + c;
        orow0[off] = pack8(o[0][0]);
        orow0[off + cv] = pack8(o[0][1]);
        orow1[off] = pack8(o[1][0]);
        orow1[off + cv] = pack8(o[1][1]);
    }
}

// The decoder's own geometry, specialised: src0 is upsampled x2, src1 x2 (kS1 == 2) or x4 (kS1 == 4).  For these ratios the
// bilinear taps are compile-time constants -- x2: output 2j reads sources (j-1, j) with weights (0.25, 0.75), output
// 2j+1 reads (j, j+1) with (0.75, 0.25); x4: outputs 4k .. 4k+3 read (k-1, k) with lambda 0.625 / 0.875 and (k, k+1)
// with 0.125 / 0.375 -- and clamping the source index at the border reproduces ATen's clamped coordinate exactly
// (w0 v + w1 v == v for these weights).  No coordinate arithmetic, no selects, and the interpolation runs on packed
// fp32 pairs (FMUL2 / FFMA2): a third of the generic kernel's instructions, which was bound by instruction issue.
// A thread still produces a 2x2 block of output pixels of one 8-channel vector (x2: 3x3 source patch; x4: 2x2).
template <int kS>
__device__ __forceinline__ void upcat_block(const uint4* __restrict__ src, int b, int h, int w, int ncv, int cc, int by, int bx,
                                            f32x2 (&o)[2][2][4]) {
    const size_t img = static_cast<size_t>(b) * h * w;
    if (kS == 2) {
        const int rows[3] = {by > 0 ? by - 1 : 0, by, by + 1 < h ? by + 1 : h - 1};
        const int cols[3] = {bx > 0 ? bx - 1 : 0, bx, bx + 1 < w ? bx + 1 : w - 1};
        f32x2 p[3][3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q)
                unpack8_f2(__ldg(src + (img + static_cast<size_t>(rows[r]) * w + cols[q]) * ncv + cc), p[r][q]);
        const f32x2 k25 = f2_make(0.25f, 0.25f), k75 = f2_make(0.75f, 0.75f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f32x2 hx[3][2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const f32x2 mid = f2_mul(p[r][1][i], k75);
                hx[r][0] = f2_fma(p[r][0][i], k25, mid);
                hx[r][1] = f2_fma(p[r][2][i], k25, mid);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const f32x2 mid = f2_mul(hx[1][q], k75);
                o[0][q][i] = f2_fma(hx[0][q], k25, mid);
                o[1][q][i] = f2_fma(hx[2][q], k25, mid);
            }
        }
    } else {  // x4: both outputs of a pair share their two sources
        const int r0 = (by >> 1) - 1 + (by & 1), c0 = (bx >> 1) - 1 + (bx & 1);
        const int rows[2] = {r0 > 0 ? r0 : 0, r0 + 1 < h ? r0 + 1 : h - 1};
        const int cols[2] = {c0 > 0 ? c0 : 0, c0 + 1 < w ? c0 + 1 : w - 1};
        f32x2 p[2][2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 2; ++q)
                unpack8_f2(__ldg(src + (img + static_cast<size_t>(rows[r]) * w + cols[q]) * ncv + cc), p[r][q]);
        const float lya = (by & 1) ? 0.125f : 0.625f, lxa = (bx & 1) ? 0.125f : 0.625f;  // lambda of the first output; second = + 0.25
        const f32x2 wx1[2] = {f2_make(lxa, lxa), f2_make(lxa + 0.25f, lxa + 0.25f)};
        const f32x2 wx0[2] = {f2_make(1.f - lxa, 1.f - lxa), f2_make(0.75f - lxa, 0.75f - lxa)};
        const f32x2 wy1[2] = {f2_make(lya, lya), f2_make(lya + 0.25f, lya + 0.25f)};
        const f32x2 wy0[2] = {f2_make(1.f - lya, 1.f - lya), f2_make(0.75f - lya, 0.75f - lya)};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f32x2 hx[2][2];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int q = 0; q < 2; ++q) hx[r][q] = f2_fma(p[r][0][i], wx0[q], f2_mul(p[r][1][i], wx1[q]));
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int q = 0; q < 2; ++q) o[a][q][i] = f2_fma(hx[0][q], wy0[a], f2_mul(hx[1][q], wy1[a]));
        }
    }
}

// The two sources are walked in separate loops: with one loop over the concatenated channel vectors a warp straddles both
// sources and executes both interpolation paths (measured: 3.9 TB/s of output for 256 + 64 channels against 5.8 TB/s for
// a single 128-channel source).
template <int kS, int kOther>
__device__ __forceinline__ void upcat_source(const uint4* __restrict__ src, int h, int w, int ncv, int c_off, int cv, int b, int by,
                                             uint4* __restrict__ orow0, uint4* __restrict__ orow1, int Wo) {
    const int total = (Wo >> 1) * ncv;
    const int shift = (ncv & (ncv - 1)) == 0 ? 31 - __clz(ncv) : -1;  // channel-vector counts are powers of two in SPEGNet
    for (int t = threadIdx.x; t < total; t += 256) {
        const int bx = shift >= 0 ? t >> shift : t / ncv;
        const int c = t - bx * ncv;
        f32x2 o[2][2][4];
        upcat_block<kS>(src, b, h, w, ncv, c, by, bx, o);
        const size_t off = static_cast<size_t>(2 * bx) * cv + c_off + c;
        orow0[off] = pack8_f2(o[0][0]);
        orow0[off + cv] = pack8_f2(o[0][1]);
        orow1[off] = pack8_f2(o[1][0]);
        orow1[off + cv] = pack8_f2(o[1][1]);
    }
}

template <int kS1>
__global__ void __launch_bounds__(256)
upcat_fixed_kernel(const uint4* __restrict__ s0, int h0, int w0, int c0v, const uint4* __restrict__ s1, int h1, int w1,
                   int c1v, uint4* __restrict__ out, int Ho, int Wo) {
    pdl_prologue();
    const int cv = c0v + c1v;
    const int by = blockIdx.x, b = blockIdx.y;
    uint4* orow0 = out + (static_cast<size_t>(b) * Ho + 2 * by) * Wo * cv;
    uint4* orow1 = orow0 + static_cast<size_t>(Wo) * cv;
    upcat_source<2, 0>(s0, h0, w0, c0v, 0, cv, b, by, orow0, orow1, Wo);
    if (c1v > 0) upcat_source<kS1, 1>(s1, h1, w1, c1v, c0v, cv, b, by, orow0, orow1, Wo);
}

// ------------------------------------------------------------------------------------------------
// CFI fusion tail.  The 1x1 conv over concat(f2, up2(f3), up4(f4)) is linear, so it is evaluated at
// each source's native resolution (g2, g3, g4 = per-scale GEMM outputs, fp32, BN scale folded) and
// combined here: fused = relu(g2 + up2(g3) + up4(g4) + bias).  One CTA per (image, output row);
// thread = 4 channels, looping over the row's pixels; also emits per-row channel sums for the SE squeeze.
// ------------------------------------------------------------------------------------------------
// The x2 / x4 taps are constants (see upcat_block); the row interpolation of a source column is done once and shared by
// the outputs that read the column (a sliding window of 4 g3 columns and 3 g4 columns along x), on packed fp32 pairs:
// 10 16-byte loads per 4 outputs instead of 36 and no coordinate arithmetic.
struct F4 {
    f32x2 lo, hi;
};
__device__ __forceinline__ F4 f4_load(const float4* p) {
    const float4 v = __ldg(p);
    return F4{f2_make(v.x, v.y), f2_make(v.z, v.w)};
}
__device__ __forceinline__ F4 f4_lerp(const F4& u, float wu, const F4& v, float wv) {  // wu * u + wv * v
    const f32x2 a = f2_make(wu, wu), b = f2_make(wv, wv);
    return F4{f2_fma(u.lo, a, f2_mul(v.lo, b)), f2_fma(u.hi, a, f2_mul(v.hi, b))};
}
// A CTA is one output row: 128 channel quads x `nchunk` chunks of the row (512 threads for rows of >= 16 source columns);
// the chunks' partial row sums are added in chunk order through shared memory, so the SE squeeze input is the same
// for every batch size.  (One chunk per row left a 16-step dependent loop per thread: 41 us at batch 1.)
__global__ void __launch_bounds__(512)
fusion_combine_kernel(const float4* __restrict__ g2, const float4* __restrict__ g3, const float4* __restrict__ g4,
                      const float4* __restrict__ bias, uint2* __restrict__ fused, float4* __restrict__ partial, int Hs,
                      int C4, int nchunk) {
    pdl_prologue();
    __shared__ float4 chunk_sum[4][128];
    const int b = blockIdx.y, y = blockIdx.x;
    const int H3 = Hs >> 1, H4 = Hs >> 2;
    const int lane_c = threadIdx.x & 127, chunk = threadIdx.x >> 7;
    const int k_per = H4 / nchunk, k_begin = chunk * k_per, k_end = k_begin + k_per;
    // source rows and their weights (clamped index == ATen's clamped coordinate for these taps)
    const int j3 = y >> 1, k4 = y >> 2;
    const int r30 = (y & 1) ? j3 : (j3 > 0 ? j3 - 1 : 0), r31 = (y & 1) ? (j3 + 1 < H3 ? j3 + 1 : H3 - 1) : j3;
    const float wy31 = (y & 1) ? 0.25f : 0.75f, wy30 = 1.f - wy31;
    const int q4 = y & 3;
    const int r40 = q4 < 2 ? (k4 > 0 ? k4 - 1 : 0) : k4, r41 = q4 < 2 ? k4 : (k4 + 1 < H4 ? k4 + 1 : H4 - 1);
    const float wy41 = q4 == 0 ? 0.625f : (q4 == 1 ? 0.875f : (q4 == 2 ? 0.125f : 0.375f)), wy40 = 1.f - wy41;
    const float4* g3a = g3 + (static_cast<size_t>(b) * H3 + r30) * H3 * C4;
    const float4* g3b = g3 + (static_cast<size_t>(b) * H3 + r31) * H3 * C4;
    const float4* g4a = g4 + (static_cast<size_t>(b) * H4 + r40) * H4 * C4;
    const float4* g4b = g4 + (static_cast<size_t>(b) * H4 + r41) * H4 * C4;
    const float4* g2r = g2 + (static_cast<size_t>(b) * Hs + y) * Hs * C4;
    uint2* out = fused + (static_cast<size_t>(b) * Hs + y) * Hs * C4;
    for (int cbase = 0; cbase < C4; cbase += 128) {
        const int c = cbase + lane_c;
        f32x2 acc_lo = f2_make(0.f, 0.f), acc_hi = acc_lo;
        if (c < C4) {
            auto col3 = [&](int j) {  // row-interpolated g3 column j (clamped)
                j = j < 0 ? 0 : (j < H3 ? j : H3 - 1);
                return f4_lerp(f4_load(g3a + static_cast<size_t>(j) * C4 + c), wy30, f4_load(g3b + static_cast<size_t>(j) * C4 + c), wy31);
            };
            auto col4 = [&](int k) {
                k = k < 0 ? 0 : (k < H4 ? k : H4 - 1);
                return f4_lerp(f4_load(g4a + static_cast<size_t>(k) * C4 + c), wy40, f4_load(g4b + static_cast<size_t>(k) * C4 + c), wy41);
            };
            const F4 bi = f4_load(bias + c);
            F4 a3 = col3(2 * k_begin - 1), b3 = col3(2 * k_begin), a4 = col4(k_begin - 1), b4 = col4(k_begin);
            for (int k = k_begin; k < k_end; ++k) {
                const F4 c3 = col3(2 * k + 1), d3 = col3(2 * k + 2), c4 = col4(k + 1);
                F4 x2[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) x2[i] = f4_load(g2r + static_cast<size_t>(4 * k + i) * C4 + c);
                const F4 t3[4] = {f4_lerp(a3, 0.25f, b3, 0.75f), f4_lerp(b3, 0.75f, c3, 0.25f), f4_lerp(b3, 0.25f, c3, 0.75f),
                                  f4_lerp(c3, 0.75f, d3, 0.25f)};
                const F4 t4[4] = {f4_lerp(a4, 0.375f, b4, 0.625f), f4_lerp(a4, 0.125f, b4, 0.875f), f4_lerp(b4, 0.875f, c4, 0.125f),
                                  f4_lerp(b4, 0.625f, c4, 0.375f)};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const f32x2 lo = f2_add(f2_add(f2_add(x2[i].lo, t3[i].lo), t4[i].lo), bi.lo);
                    const f32x2 hi = f2_add(f2_add(f2_add(x2[i].hi, t3[i].hi), t4[i].hi), bi.hi);
                    float v0, v1, v2, v3;
                    f2_split(lo, v0, v1);
                    f2_split(hi, v2, v3);
                    v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
                    acc_lo = f2_add(acc_lo, f2_make(v0, v1));
                    acc_hi = f2_add(acc_hi, f2_make(v2, v3));
                    out[static_cast<size_t>(4 * k + i) * C4 + c] = make_uint2(pack2(v0, v1), pack2(v2, v3));
                }
                a3 = c3; b3 = d3; a4 = b4; b4 = c4;
            }
        }
        float s0, s1, s2, s3;
        f2_split(acc_lo, s0, s1);
        f2_split(acc_hi, s2, s3);
        chunk_sum[chunk][lane_c] = make_float4(s0, s1, s2, s3);
        __syncthreads();
        if (chunk == 0 && c < C4) {
            float4 t = chunk_sum[0][lane_c];
            for (int q = 1; q < nchunk; ++q) {
                const float4 u = chunk_sum[q][lane_c];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            partial[(static_cast<size_t>(b) * Hs + y) * C4 + c] = t;
        }
        __syncthreads();
    }
}

// Per-row channel sums of a bf16 NHWC map (global-average-pool partials): one CTA per (image, row).
__global__ void __launch_bounds__(128)
row_sums_kernel(const uint2* __restrict__ x, float4* __restrict__ partial, int H, int W, int C4) {
    pdl_prologue();
    const int b = blockIdx.y, y = blockIdx.x;
    for (int c = threadIdx.x; c < C4; c += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int xx = 0; xx < W; ++xx) {
            const uint2 u = x[((static_cast<size_t>(b) * H + y) * W + xx) * C4 + c];
            acc.x += h_lo(u.x); acc.y += h_hi(u.x); acc.z += h_lo(u.y); acc.w += h_hi(u.y);
        }
        partial[(static_cast<size_t>(b) * H + y) * C4 + c] = acc;
    }
}

// Squeeze-excite gate (or, with `sigmoid_out`=0 and bias, the e-ASPP global branch): one CTA per image.
//   mean[c] = sum_rows partial[b,r,c] / count
//   se:      gate = sigmoid(W2 @ relu(W1 @ mean))           W1 [R,C], W2 [C,R]
//   global:  out  = relu(W1 @ mean + bias)                  W1 [C,C] (BN folded), W2 == nullptr
__global__ void __launch_bounds__(512)
pooled_mlp_kernel(const float* __restrict__ partial, int rows, float inv_count, const float* __restrict__ W1,
                  const float* __restrict__ b1, int R, const float* __restrict__ W2, float* __restrict__ out, int C) {
    pdl_prologue();
    extern __shared__ float sm[];
    float* mean = sm;          // [C]
    float* hidden = sm + C;    // [R]
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += partial[(static_cast<size_t>(b) * rows + r) * C + c];
        mean[c] = s * inv_count;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < R; j += nwarps) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += W1[static_cast<size_t>(j) * C + c] * mean[c];
        s = warp_sum(s);
        if (lane == 0) {
            if (b1 != nullptr) s += b1[j];
            s = fmaxf(s, 0.f);
            if (W2 == nullptr) out[static_cast<size_t>(b) * R + j] = s;
            hidden[j] = s;
        }
    }
    if (W2 == nullptr) return;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int j = 0; j < R; ++j) s += W2[static_cast<size_t>(c) * R + j] * hidden[j];
        out[static_cast<size_t>(b) * C + c] = 1.f / (1.f + __expf(-s));
    }
}

// x[b, p, c] *= gate[b, c]  (bf16 NHWC in place, 8 channels per thread)
__global__ void __launch_bounds__(256)
scale_channels_kernel(uint4* __restrict__ x, const float* __restrict__ gate, long long total, int HW, int C8) {
    pdl_prologue();
    const long long idx = blockIdx.x * 256ll + threadIdx.x;
    if (idx >= total) return;
    const int c = idx % C8;
    const long long b = idx / (static_cast<long long>(HW) * C8);
    float f[8];
    unpack8(x[idx], f);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + b * C8 * 8) + 2 * c);
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + b * C8 * 8) + 2 * c + 1);
    f[0] *= g0.x; f[1] *= g0.y; f[2] *= g0.z; f[3] *= g0.w;
    f[4] *= g1.x; f[5] *= g1.y; f[6] *= g1.z; f[7] *= g1.w;
    x[idx] = pack8(f);
}

// ------------------------------------------------------------------------------------------------
// e-ASPP core: four depth-wise dilated 3x3 branches (+BN+ReLU), the broadcast global branch, the
// 640-channel concat and the grouped 1x1 "fusion" conv (+BN+ReLU) in ONE pass, nothing materialised.
// Concat channel c = branch*128 + ch feeds output channel c/5 with weight wf[c/5][c%5]
// (models/feature_integration.py:349-357,411-412).  One thread owns 40 consecutive concat channels
// (= five 8-channel vectors, each inside one branch) of one pixel and produces 8 output channels.
// ------------------------------------------------------------------------------------------------
struct AsppParams {
    const uint4* x;        // [B,H,W,128] bf16 (reduce output)
    const float* dw;       // [4][9][128] depth-wise weights, BN scale folded
    const float* dw_bias;  // [4][128]
    const float* gvec;     // [B][128] global branch output (already BN+ReLU)
    const float* wf;       // [128][5] grouped-conv weights, BN scale folded
    const float* wf_bias;  // [128]
    uint4* y;              // [B,H,W,128] bf16
    int B, H, W;
    int dil[4];
};

// One CTA = (band of rows, slice t, image): 2048 pixels, 4 per thread.  For each of the slice's five 8-channel vectors
// the rows the dilated taps can reach are staged in shared memory ONCE ([rows][W] 16-byte vectors of that channel
// group) and all nine taps of every pixel are served from there; the first version fetched every tap from L2
// (36 x 16 B per thread for 16 B of output, 2.4 GB of L2 traffic per launch at batch 64: 740 us, L2-bound).
// kAsppPix pixels per thread: 4 (band of 2048 pixels) when the batch fills the machine, 1 (512-pixel bands, 4x the CTAs,
// more halo rows staged per output) for small batches, where the kernel is a chain of staging latencies on 32 CTAs
// (50 us at batch 1).  Every output pixel is computed by the same instruction sequence either way.
// Tried and dropped (late round 2): staging only the three row groups y - d, y, y + d of a 4-row band for all five
// vectors at once (thread = pixel, one staging phase): correct, 663 us -- its halo redundancy is a constant 3x where a
// 32-row band averages 1.6x, and the staging traffic is what the kernel pays for.
constexpr int kAsppThreads = 512;

template <int kAsppPix>
__global__ void __launch_bounds__(kAsppThreads, 2) aspp_kernel(const AsppParams p, int band_rows) {
    pdl_prologue();
    extern __shared__ uint4 slab[];  // [rows][W]
    const int t = blockIdx.y;        // which 40-channel slice of the 640-channel concat -> output channels 8t .. 8t+7
    const int b = blockIdx.z;
    const int y_band = blockIdx.x * band_rows;
    const int W = p.W, H = p.H;
    const size_t img = static_cast<size_t>(b) * H * W;
    float out_acc[kAsppPix][8];
#pragma unroll
    for (int q = 0; q < kAsppPix; ++q)
#pragma unroll
        for (int g = 0; g < 8; ++g) out_acc[q][g] = 0.f;

#pragma unroll
    for (int v = 0; v < 5; ++v) {
        const int c0 = 40 * t + 8 * v;
        const int br = c0 >> 7, ch0 = c0 & 127;
        float wfv[8];  // grouped-conv weight of concat channel c0 + i: wf[(c0 + i) / 5][(c0 + i) % 5] = wf[c0 + i] (row-major [128][5])
#pragma unroll
        for (int i = 0; i < 8; ++i) wfv[i] = __ldg(p.wf + c0 + i);
        if (br == 4) {  // broadcast global branch (already BN + ReLU)
            float gv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) gv[i] = __ldg(p.gvec + b * 128 + ch0 + i);
#pragma unroll
            for (int q = 0; q < kAsppPix; ++q)
#pragma unroll
                for (int i = 0; i < 8; ++i) out_acc[q][(8 * v + i) / 5] = fmaf(wfv[i], gv[i], out_acc[q][(8 * v + i) / 5]);
            continue;
        }
        const int d = p.dil[br];
        const int y_lo = max(0, y_band - d), y_hi = min(H, y_band + band_rows + d);
        __syncthreads();  // the previous vector's slab is no longer read
        for (int i = threadIdx.x; i < (y_hi - y_lo) * W; i += kAsppThreads)
            slab[i] = __ldg(p.x + (img + static_cast<size_t>(y_lo) * W + i) * 16 + (ch0 >> 3));
        __syncthreads();
        // the nine tap weights of this channel group are the same for every thread: they stay in global memory / L1
        // (uniform float4 loads per tap) instead of 72 registers per thread, which held the kernel at one CTA per SM
        const float4* wt4 = reinterpret_cast<const float4*>(p.dw + br * 9 * 128 + ch0);
        float bias[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) bias[i] = __ldg(p.dw_bias + br * 128 + ch0 + i);
#pragma unroll
        for (int q = 0; q < kAsppPix; ++q) {
            const int pix = threadIdx.x + q * kAsppThreads;  // pixel inside the band, x fastest
            const int y0 = y_band + pix / W, x0 = pix % W;
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            if (y0 < H && pix < band_rows * W) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int yy = y0 + (tap / 3 - 1) * d, xx = x0 + (tap % 3 - 1) * d;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    float f[8];
                    unpack8(slab[(yy - y_lo) * W + xx], f);
                    const float4 w0 = __ldg(wt4 + tap * 32), w1 = __ldg(wt4 + tap * 32 + 1);
                    acc[0] = fmaf(f[0], w0.x, acc[0]); acc[1] = fmaf(f[1], w0.y, acc[1]);
                    acc[2] = fmaf(f[2], w0.z, acc[2]); acc[3] = fmaf(f[3], w0.w, acc[3]);
                    acc[4] = fmaf(f[4], w1.x, acc[4]); acc[5] = fmaf(f[5], w1.y, acc[5]);
                    acc[6] = fmaf(f[6], w1.z, acc[6]); acc[7] = fmaf(f[7], w1.w, acc[7]);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float cat = fmaxf(acc[i] + bias[i], 0.f);
                out_acc[q][(8 * v + i) / 5] = fmaf(wfv[i], cat, out_acc[q][(8 * v + i) / 5]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < kAsppPix; ++q) {
        const int pix = threadIdx.x + q * kAsppThreads;
        const int y0 = y_band + pix / W, x0 = pix % W;
        if (y0 >= H || pix >= band_rows * W) continue;  // (the band holds band_rows * W <= 2048 pixels)
        float o[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) o[g] = fmaxf(out_acc[q][g] + __ldg(p.wf_bias + 8 * t + g), 0.f);
        p.y[(img + static_cast<size_t>(y0) * W + x0) * 16 + t] = pack8(o);
    }
}

// ------------------------------------------------------------------------------------------------
// Operand of the border-column correction GEMM of the fused "bilinear x2 -> 3x3 conv" (spg_conv3x3_up2_h16):
// for side s (0: column 0, 1: column W-1), image b, row y:   A[s][b*H + y][(cls*3 + dy)*C + c] = x[b, y+dy-1, col, c]
// inside the block cls = row class of y (0 top, 1 interior, 2 bottom) and zero in the other two class blocks and
// outside the image, so that ONE weight matrix [N, 9C] carries the three row-class variants.  8 channels / thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
up2_border_gather_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int B, int H, int W, int C8) {
    pdl_prologue();
    const long long total = 2ll * B * H * 9 * C8;
    const long long idx = blockIdx.x * 256ll + threadIdx.x;
    if (idx >= total) return;
    const int c = idx % C8;
    const int blk = (idx / C8) % 9;       // cls * 3 + dy
    const long long r = idx / (9ll * C8);  // side * B*H + b*H + y
    const int y = r % H;
    const int b = (r / H) % B;
    const int side = r / (static_cast<long long>(B) * H);
    const int cls = y == 0 ? 0 : (y == H - 1 ? 2 : 1);
    const int yy = y + blk % 3 - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (blk / 3 == cls && yy >= 0 && yy < H)
        v = x[((static_cast<size_t>(b) * H + yy) * W + (side ? W - 1 : 0)) * C8 + c];
    out[idx] = v;
}

// bf16 NHWC -> fp32 NCHW (for the lazily materialised `features` entries of the output dict).
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const uint16_t* __restrict__ x, float* __restrict__ y, int HW, int C, long long total) {
    pdl_prologue();
    const long long idx = blockIdx.x * 256ll + threadIdx.x;
    if (idx >= total) return;
    const int p = idx % HW;
    const int c = (idx / HW) % C;
    const long long b = idx / (static_cast<long long>(HW) * C);
    y[idx] = h_to_float(x[(b * HW + p) * C + c]);
}

// ------------------------------------------------------------------------------------------------
// Mask quantisation exactly as the reference's metric wrapper (utils/metrics.py:205-210): q = uint8(sigmoid(x)
// * 255) with truncation (optionally sigmoid twice: the evaluator path, engine/evaluator.py:544), plus the
// per-image integer statistics from which MAE after min-max normalisation follows exactly:
//   stats[b] = { 255 - min q, max q, #gt_fg, sum q over gt background, sum q over gt foreground }
// Integer atomics only, so the result is independent of scheduling.  4 pixels per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mask_stats_kernel(const float4* __restrict__ logits, const uchar4* __restrict__ gt, uchar4* __restrict__ mask,
                  unsigned* __restrict__ stats, int HW4, int double_sigmoid) {
    pdl_prologue();
    const int b = blockIdx.y;
    unsigned inv_min = 0, mx = 0, nfg = 0, sbg = 0, sfg = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < HW4; i += gridDim.x * 256) {
        const float4 x = logits[static_cast<size_t>(b) * HW4 + i];
        const uchar4 g = gt[static_cast<size_t>(b) * HW4 + i];
        const float xs[4] = {x.x, x.y, x.z, x.w};
        const unsigned char gs[4] = {g.x, g.y, g.z, g.w};
        unsigned char qs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float p = 1.f / (1.f + expf(-xs[k]));
            if (double_sigmoid) p = 1.f / (1.f + expf(-p));
            const unsigned q = static_cast<unsigned>(p * 255.f);
            qs[k] = static_cast<unsigned char>(q);
            inv_min = max(inv_min, 255u - q);
            mx = max(mx, q);
            if (gs[k] > 128) { ++nfg; sfg += q; } else { sbg += q; }
        }
        mask[static_cast<size_t>(b) * HW4 + i] = make_uchar4(qs[0], qs[1], qs[2], qs[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        inv_min = max(inv_min, __shfl_xor_sync(0xffffffffu, inv_min, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        nfg += __shfl_xor_sync(0xffffffffu, nfg, o);
        sbg += __shfl_xor_sync(0xffffffffu, sbg, o);
        sfg += __shfl_xor_sync(0xffffffffu, sfg, o);
    }
    if ((threadIdx.x & 31) == 0) {
        unsigned* st = stats + b * 8;
        atomicMax(st + 0, inv_min);
        atomicMax(st + 1, mx);
        atomicAdd(st + 2, nfg);
        atomicAdd(st + 3, sbg);
        atomicAdd(st + 4, sfg);
    }
}

// Same, one pixel per thread iteration: image sizes that are not a multiple of 4 (original-resolution ground truth).
__global__ void __launch_bounds__(256)
mask_stats_scalar_kernel(const float* __restrict__ logits, const unsigned char* __restrict__ gt,
                         unsigned char* __restrict__ mask, unsigned* __restrict__ stats, int HW, int double_sigmoid) {
    pdl_prologue();
    const int b = blockIdx.y;
    unsigned inv_min = 0, mx = 0, nfg = 0, sbg = 0, sfg = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) {
        float p = 1.f / (1.f + expf(-logits[static_cast<size_t>(b) * HW + i]));
        if (double_sigmoid) p = 1.f / (1.f + expf(-p));
        const unsigned q = static_cast<unsigned>(p * 255.f);
        mask[static_cast<size_t>(b) * HW + i] = static_cast<unsigned char>(q);
        inv_min = max(inv_min, 255u - q);
        mx = max(mx, q);
        if (gt[static_cast<size_t>(b) * HW + i] > 128) { ++nfg; sfg += q; } else { sbg += q; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        inv_min = max(inv_min, __shfl_xor_sync(0xffffffffu, inv_min, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        nfg += __shfl_xor_sync(0xffffffffu, nfg, o);
        sbg += __shfl_xor_sync(0xffffffffu, sbg, o);
        sfg += __shfl_xor_sync(0xffffffffu, sfg, o);
    }
    if ((threadIdx.x & 31) == 0) {
        unsigned* st = stats + b * 8;
        atomicMax(st + 0, inv_min);
        atomicMax(st + 1, mx);
        atomicAdd(st + 2, nfg);
        atomicAdd(st + 3, sbg);
        atomicAdd(st + 4, sfg);
    }
}

// LayerNorm whose result is BIT-IDENTICAL to what a residual GEMM with spg_epilogue_t.ln_apply_* stores (same slice
// table, same sequential per-slice sums, same merge and normalisation: ln_stats.cuh).  One warp per row: lane q < count
// walks slice q of the row in column order straight from global memory (16-byte loads; no shared memory: six lanes on a
// stride of 96 floats would all hit one bank), the partials are merged by every lane, the normalisation re-reads the
// row (L1 / L2 hit) coalesced.  Used in the latency regime, where the fused form does not pay: 4.4 us for 1024 rows of
// 576 channels against 2.7 us for the free-order kernel above (the 96-long dependent chains are the price of matching
// the GEMM epilogue's summation order; packing several rows into a warp measured slower, 6.4 us: divergent row reads).
__global__ void __launch_bounds__(128) layernorm_sliced_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, uint16_t* __restrict__ y, int M,
                                                               int C, float eps, LnSlices sl) {
    pdl_prologue();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float inv_cols = 1.0f / static_cast<float>(C);
    int lo = 0, hi = 0;
#pragma unroll
    for (int q = 0; q < kLnMaxSlices; ++q)  // static indexing: the table stays in the constant bank
        if (q == lane && q < sl.count) {
            lo = sl.bound[q];
            hi = sl.bound[q + 1];
        }
    for (int r = blockIdx.x * 4 + warp; r < M; r += gridDim.x * 4) {
        const float* xr = x + static_cast<size_t>(r) * C;
        float2 mine = make_float2(0.f, 0.f);
        if (hi > lo) {
            const float4* src = reinterpret_cast<const float4*>(xr + lo);
            const float shift = __ldg(xr + lo);
            float s1 = 0.f, s2 = 0.f;
            for (int j = 0; j < (hi - lo) >> 2; ++j) {
                const float4 v = __ldg(src + j);
                ln_accumulate(v.x, shift, s1, s2);
                ln_accumulate(v.y, shift, s1, s2);
                ln_accumulate(v.z, shift, s1, s2);
                ln_accumulate(v.w, shift, s1, s2);
            }
            mine = ln_slice_stats(shift, s1, s2, static_cast<float>(hi - lo));
        }
        const float2 rm = ln_merge(
            sl.count,
            [&](int q) { return make_float2(__shfl_sync(0xffffffffu, mine.x, q), __shfl_sync(0xffffffffu, mine.y, q)); },
            [&](int q) { return static_cast<float>(__shfl_sync(0xffffffffu, hi - lo, q)); }, inv_cols, eps);
        uint16_t* yr = y + static_cast<size_t>(r) * C;
        for (int i = lane * 8; i < C; i += 256) {
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(xr + i)), a1 = __ldg(reinterpret_cast<const float4*>(xr + i) + 1);
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + i)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + i) + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + i)), b1 = __ldg(reinterpret_cast<const float4*>(beta + i) + 1);
            float v[8];
            v[0] = ln_normalise(a0.x, rm, g0.x, b0.x); v[1] = ln_normalise(a0.y, rm, g0.y, b0.y);
            v[2] = ln_normalise(a0.z, rm, g0.z, b0.z); v[3] = ln_normalise(a0.w, rm, g0.w, b0.w);
            v[4] = ln_normalise(a1.x, rm, g1.x, b1.x); v[5] = ln_normalise(a1.y, rm, g1.y, b1.y);
            v[6] = ln_normalise(a1.z, rm, g1.z, b1.z); v[7] = ln_normalise(a1.w, rm, g1.w, b1.w);
            *reinterpret_cast<uint4*>(yr + i) = pack8(v);
        }
    }
}

// dst[b, y, x, :] = src[b, y, x, :] for y < H, x < W between two NHWC layouts [B, Hs, Ws, C] -> [B, Hd, Wd, C] (16-byte
// vectors): the zero-padded token grid that window attention needs when the grid does not tile into windows
// (HF:modeling_sam2.py:395-399 pads after norm1), and the crop after it (:435-437).  The rest of dst is not touched.
__global__ void __launch_bounds__(256) copy_grid_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int H, int W,
                                                        int Hs, int Ws, int Hd, int Wd, int cv, long long total) {
    pdl_prologue();
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
        const int v = static_cast<int>(i % cv);
        long long t = i / cv;
        const int x = static_cast<int>(t % W);
        t /= W;
        const int y = static_cast<int>(t % H);
        const long long b = t / H;
        dst[((b * Hd + y) * Wd + x) * cv + v] = __ldg(src + ((b * Hs + y) * Ws + x) * cv + v);
    }
}

inline unsigned blocks_for(long long total, int per_block = 256) {
    return static_cast<unsigned>((total + per_block - 1) / per_block);
}
// grid-stride kernels: enough blocks to fill the machine several times over, far fewer than one per 256 elements
inline unsigned capped_blocks(long long total) {
    const long long want = (total + 255) / 256;
    const long long cap = static_cast<long long>(sm_count()) * 16;
    return static_cast<unsigned>(want < cap ? want : cap);
}

}  // namespace
}  // namespace spg

using namespace spg;

extern "C" int spg_layernorm_f32_h16(const float* x, const float* gamma, const float* beta, void* y, int M, int C,
                                      float eps, const spg_launch_t* launch) {
    SPG_CHECK_ARG(x && gamma && beta && y, "null pointer");
    SPG_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0 && C <= kLnMaxC, "LayerNorm needs C %% 4 == 0 and C <= 1152 (C=%d)", C);
    const LaunchCtx st(launch);
    uint16_t* yo = static_cast<uint16_t*>(y);
    // rows per warp so that one block (8 warps) streams roughly 64 KB+ and the grid stays far below the block-launch rate
    // small M (batch 1): fewer rows per warp so that the grid still covers the machine
    auto fit = [&](int rpw, int rows_per_iter) {
        while (rpw > 1 && (M + 8 * rpw * rows_per_iter - 1) / (8 * rpw * rows_per_iter) < 2 * sm_count()) rpw >>= 1;
        return rpw;
    };
    if (C <= 256) {
        const int rpw = fit(8, 2);  // 8 * 2 rows per warp
        const int rows_per_block = 8 * rpw * 2;
        SPG_CHECK_CUDA((launch_pdl(layernorm_kernel<16, 4>, (M + rows_per_block - 1) / rows_per_block, 256, 0, st, x, gamma, beta, yo, M, C, eps, rpw, st.reverse ? 1 : 0)));
    } else {
        // register footprint follows the row length (float4 per lane): 3 for C <= 384, 5 for C <= 640, else 9
        const int rpw = fit(4, 1);
        const int rows_per_block = 8 * rpw;
        const unsigned grid = (M + rows_per_block - 1) / rows_per_block;
        if (C <= 384)
            SPG_CHECK_CUDA((launch_pdl(layernorm_kernel<32, 3>, grid, 256, 0, st, x, gamma, beta, yo, M, C, eps, rpw, st.reverse ? 1 : 0)));
        else if (C <= 640)
            SPG_CHECK_CUDA((launch_pdl(layernorm_kernel<32, 5>, grid, 256, 0, st, x, gamma, beta, yo, M, C, eps, rpw, st.reverse ? 1 : 0)));
        else
            SPG_CHECK_CUDA((launch_pdl(layernorm_kernel<32, 9>, grid, 256, 0, st, x, gamma, beta, yo, M, C, eps, rpw, st.reverse ? 1 : 0)));
    }
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_copy_grid_h16(const void* src, int Hs, int Ws, void* dst, int Hd, int Wd, int B, int H, int W, int C,
                                 const spg_launch_t* launch) {
    SPG_CHECK_ARG(src && dst, "null pointer");
    SPG_CHECK_ARG(B > 0 && H > 0 && W > 0 && H <= Hs && W <= Ws && H <= Hd && W <= Wd, "bad grid copy %dx%d from %dx%d to %dx%d", H, W, Hs, Ws, Hd, Wd);
    SPG_CHECK_ARG(C % 8 == 0, "C=%d must be a multiple of 8 (16-byte vectors)", C);
    const long long total = static_cast<long long>(B) * H * W * (C / 8);
    SPG_CHECK_CUDA((launch_pdl(copy_grid_kernel, capped_blocks(total), 256, 0, LaunchCtx(launch), static_cast<const uint4*>(src),
                               static_cast<uint4*>(dst), H, W, Hs, Ws, Hd, Wd, C / 8, total)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_layernorm_matched_f32_h16(const float* x, const float* gamma, const float* beta, void* y, int M, int C,
                                             float eps, const spg_launch_t* launch) {
    SPG_CHECK_ARG(x && gamma && beta && y, "null pointer");
    SPG_CHECK_ARG(M > 0 && C > 0 && C % 8 == 0, "bad LayerNorm shape M=%d C=%d", M, C);
    const LnSlices sl = ln_slices_for(C);
    SPG_CHECK_ARG(sl.count > 0, "C=%d has no producer-side LayerNorm tiling (supported: 144, 288, 576)", C);
    const LaunchCtx st(launch);
    const unsigned want = static_cast<unsigned>((M + 3) / 4), cap = static_cast<unsigned>(sm_count()) * 16u;
    SPG_CHECK_CUDA((launch_pdl(layernorm_sliced_kernel, want < cap ? want : cap, 128, 0, st, x, gamma, beta,
                               static_cast<uint16_t*>(y), M, C, eps, sl)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_patchify_7x7s4(const float* x, void* cols, int B, int S, const spg_launch_t* launch) {
    SPG_CHECK_ARG(x && cols, "null pointer");
    SPG_CHECK_ARG(B > 0 && S > 0 && S % 4 == 0, "bad image size S=%d", S);
    SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
    const long long total = static_cast<long long>(B) * (S / 4) * (S / 4) * 21;
    SPG_CHECK_CUDA((launch_pdl(patchify_kernel, capped_blocks(total), 256, 0, LaunchCtx(launch), x, static_cast<uint4*>(cols), B, S)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_maxpool2x2_f32(const float* x, float* y, int B, int H, int W, int C, const spg_launch_t* launch) {
    SPG_CHECK_ARG(x && y, "null pointer");
    SPG_CHECK_ARG(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "maxpool needs even H, W and C %% 4 == 0");
    const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * (C / 4);
    SPG_CHECK_CUDA((launch_pdl(maxpool_kernel, capped_blocks(total), 256, 0, LaunchCtx(launch), reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), B, H, W, C / 4)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_cast_f32_h16(const float* x, void* y, long long n, const spg_launch_t* launch) {
    SPG_CHECK_ARG(x && y, "null pointer");
    SPG_CHECK_ARG(n > 0 && n % 8 == 0, "cast needs n %% 8 == 0");
    SPG_CHECK_CUDA((launch_pdl(cast_kernel, capped_blocks(n / 8), 256, 0, LaunchCtx(launch), reinterpret_cast<const float4*>(x), static_cast<uint4*>(y), n / 8)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_upsample_concat_h16(const void* src0, int h0, int w0, int c0, const void* src1, int h1, int w1,
                                        int c1, void* out, int B, int Ho, int Wo, const spg_launch_t* launch) {
    SPG_CHECK_ARG(src0 && out, "null pointer");
    SPG_CHECK_ARG(c0 % 8 == 0 && c1 % 8 == 0 && c0 > 0 && c1 >= 0, "channel counts must be multiples of 8");
    SPG_CHECK_ARG(c1 == 0 || src1 != nullptr, "src1 is NULL but c1 > 0");
    SPG_CHECK_ARG(Ho % 2 == 0 && Wo % 2 == 0, "output size must be even");
    SPG_CHECK_ARG(Ho % h0 == 0 && Ho / h0 >= 2 && Wo % w0 == 0 && Wo / w0 >= 2, "src0 must be upsampled by an integer factor >= 2");
    SPG_CHECK_ARG(c1 == 0 || (Ho % h1 == 0 && Ho / h1 >= 2 && Wo % w1 == 0 && Wo / w1 >= 2),
                  "src1 must be upsampled by an integer factor >= 2");
    // the decoder's ratios (src0 x2; src1 absent, x2 or x4) run the constant-tap kernel, anything else the generic one
    const bool x2_0 = Ho == 2 * h0 && Wo == 2 * w0;
    const int s1 = c1 == 0 ? 2 : ((Ho == 2 * h1 && Wo == 2 * w1) ? 2 : ((Ho == 4 * h1 && Wo == 4 * w1) ? 4 : 0));
    static const int fixed_env = [] { const char* e = getenv("SPG_UPCAT_FIXED"); return e ? atoi(e) : 1; }();
    if (fixed_env && x2_0 && s1 == 2)
        SPG_CHECK_CUDA((launch_pdl(upcat_fixed_kernel<2>, dim3(Ho / 2, B), 256, 0, LaunchCtx(launch), static_cast<const uint4*>(src0), h0, w0, c0 / 8,
                                   static_cast<const uint4*>(src1), h1, w1, c1 / 8, static_cast<uint4*>(out), Ho, Wo)));
    else if (fixed_env && x2_0 && s1 == 4)
        SPG_CHECK_CUDA((launch_pdl(upcat_fixed_kernel<4>, dim3(Ho / 2, B), 256, 0, LaunchCtx(launch), static_cast<const uint4*>(src0), h0, w0, c0 / 8,
                                   static_cast<const uint4*>(src1), h1, w1, c1 / 8, static_cast<uint4*>(out), Ho, Wo)));
    else
        SPG_CHECK_CUDA((launch_pdl(upcat_kernel, dim3(Ho / 2, B), 256, 0, LaunchCtx(launch), static_cast<const uint4*>(src0), h0, w0, c0 / 8, static_cast<const uint4*>(src1), h1, w1, c1 / 8,
            static_cast<uint4*>(out), Ho, Wo)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_fusion_combine(const float* g2, const float* g3, const float* g4, const float* bias, void* fused,
                                  float* row_sums, int B, int Hs, int C, const spg_launch_t* launch) {
    SPG_CHECK_ARG(g2 && g3 && g4 && bias && fused && row_sums, "null pointer");
    SPG_CHECK_ARG(Hs % 4 == 0 && C % 4 == 0, "fusion_combine needs Hs %% 4 == 0 and C %% 4 == 0");
    // chunks of a row per CTA: a function of the geometry only (never of B), so the row sums are batch-invariant
    const int H4 = Hs / 4;
    const int nchunk = H4 % 4 == 0 ? 4 : (H4 % 2 == 0 ? 2 : 1);
    SPG_CHECK_CUDA((launch_pdl(fusion_combine_kernel, dim3(Hs, B), 128 * nchunk, 0, LaunchCtx(launch), reinterpret_cast<const float4*>(g2), reinterpret_cast<const float4*>(g3), reinterpret_cast<const float4*>(g4),
        reinterpret_cast<const float4*>(bias), static_cast<uint2*>(fused), reinterpret_cast<float4*>(row_sums), Hs, C / 4, nchunk)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_row_sums_h16(const void* x, float* row_sums, int B, int H, int W, int C, const spg_launch_t* launch) {
    SPG_CHECK_ARG(x && row_sums, "null pointer");
    SPG_CHECK_ARG(C % 4 == 0, "row_sums needs C %% 4 == 0");
    SPG_CHECK_CUDA((launch_pdl(row_sums_kernel, dim3(H, B), 128, 0, LaunchCtx(launch), static_cast<const uint2*>(x), reinterpret_cast<float4*>(row_sums), H, W, C / 4)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_pooled_mlp(const float* row_sums, int rows, int count, const float* W1, const float* b1, int R,
                              const float* W2, float* out, int B, int C, const spg_launch_t* launch) {
    SPG_CHECK_ARG(row_sums && W1 && out, "null pointer");
    SPG_CHECK_ARG(rows > 0 && count > 0 && R > 0 && C > 0, "bad pooled_mlp shape");
    SPG_CHECK_CUDA((launch_pdl(pooled_mlp_kernel, B, 512, (C + R) * sizeof(float), LaunchCtx(launch), row_sums, rows, 1.0f / count, W1, b1, R, W2, out, C)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_scale_channels_h16(void* x, const float* gate, int B, int HW, int C, const spg_launch_t* launch) {
    SPG_CHECK_ARG(x && gate, "null pointer");
    SPG_CHECK_ARG(C % 8 == 0, "scale_channels needs C %% 8 == 0");
    const long long total = static_cast<long long>(B) * HW * (C / 8);
    SPG_CHECK_CUDA((launch_pdl(scale_channels_kernel, blocks_for(total), 256, 0, LaunchCtx(launch), static_cast<uint4*>(x), gate, total, HW, C / 8)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_easpp_branches(const void* x, const float* dw, const float* dw_bias, const float* gvec,
                                  const float* wf, const float* wf_bias, void* y, int B, int H, int W,
                                  const int* dilations, const spg_launch_t* launch) {
    SPG_CHECK_ARG(x && dw && dw_bias && gvec && wf && wf_bias && y && dilations, "null pointer");
    AsppParams p{static_cast<const uint4*>(x), dw, dw_bias, gvec, wf, wf_bias, static_cast<uint4*>(y), B, H, W,
                 {dilations[0], dilations[1], dilations[2], dilations[3]}};
    // band = 2048 pixels (4 per thread), or 512 (1 per thread) when 2048-pixel bands would leave most SMs without a CTA;
    // shared memory holds the band plus the largest dilation above and below
    SPG_CHECK_ARG(W <= 2048, "e-ASPP needs W <= 2048 (W=%d)", W);
    const long long ctas4 = static_cast<long long>((H * W + 2047) / 2048) * 16 * B;
    const int pix = (ctas4 < 2 * sm_count() && kAsppThreads / W >= 1) ? 1 : 4;
    const int band_rows = kAsppThreads * pix / W;
    int dmax = 0;
    for (int i = 0; i < 4; ++i) dmax = dilations[i] > dmax ? dilations[i] : dmax;
    const int rows = (band_rows + 2 * dmax) < H ? (band_rows + 2 * dmax) : H;
    const size_t smem = static_cast<size_t>(rows) * W * sizeof(uint4);
    SPG_CHECK_ARG(smem <= 200 * 1024, "e-ASPP band does not fit shared memory (W=%d, dilation %d)", W, dmax);
    static PerDeviceOnce attr_set;
    if (attr_set.needed()) {
        SPG_CHECK_CUDA(cudaFuncSetAttribute(aspp_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SPG_CHECK_CUDA(cudaFuncSetAttribute(aspp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set.done();
    }
    const dim3 grid((H + band_rows - 1) / band_rows, 16, B);
    if (pix == 1) SPG_CHECK_CUDA((launch_pdl(aspp_kernel<1>, grid, kAsppThreads, smem, LaunchCtx(launch), p, band_rows)));
    else SPG_CHECK_CUDA((launch_pdl(aspp_kernel<4>, grid, kAsppThreads, smem, LaunchCtx(launch), p, band_rows)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_mask_stats_u8(const float* logits, const unsigned char* gt, unsigned char* mask, unsigned* stats,
                                 int B, int HW, int double_sigmoid, const spg_launch_t* launch) {
    SPG_CHECK_ARG(logits && gt && mask && stats, "null pointer");
    SPG_CHECK_ARG(B > 0 && HW > 0, "bad shape B=%d HW=%d", B, HW);
    const LaunchCtx st(launch);
    SPG_CHECK_CUDA(cudaMemsetAsync(stats, 0, static_cast<size_t>(B) * 8 * sizeof(unsigned), st));
    if (HW % 4 != 0 || (reinterpret_cast<uintptr_t>(logits) & 15) || (reinterpret_cast<uintptr_t>(gt) & 3) ||
        (reinterpret_cast<uintptr_t>(mask) & 3)) {
        SPG_CHECK_CUDA((launch_pdl(mask_stats_scalar_kernel, dim3(min(64, (HW + 255) / 256), B), 256, 0, st, logits, gt, mask, stats, HW, double_sigmoid)));
        SPG_LAUNCHED();
        return SPG_OK;
    }
    const int per_img = min(64, (HW / 4 + 255) / 256);
    SPG_CHECK_CUDA((launch_pdl(mask_stats_kernel, dim3(per_img, B), 256, 0, st, reinterpret_cast<const float4*>(logits),
                                                        reinterpret_cast<const uchar4*>(gt),
                                                        reinterpret_cast<uchar4*>(mask), stats, HW / 4, double_sigmoid)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_up2_border_gather_h16(const void* x, void* out, int B, int H, int W, int C, const spg_launch_t* launch) {
    SPG_CHECK_ARG(x && out, "null pointer");
    SPG_CHECK_ARG(B > 0 && H >= 2 && W >= 2 && C % 8 == 0, "bad shape B=%d H=%d W=%d C=%d", B, H, W, C);
    const long long total = 2ll * B * H * 9 * (C / 8);
    SPG_CHECK_CUDA((launch_pdl(up2_border_gather_kernel, blocks_for(total), 256, 0, LaunchCtx(launch), static_cast<const uint4*>(x), static_cast<uint4*>(out), B, H, W, C / 8)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_nhwc_h16_to_nchw_f32(const void* x, float* y, int B, int HW, int C, const spg_launch_t* launch) {
    SPG_CHECK_ARG(x && y, "null pointer");
    const long long total = static_cast<long long>(B) * HW * C;
    SPG_CHECK_CUDA((launch_pdl(nhwc_to_nchw_kernel, blocks_for(total), 256, 0, LaunchCtx(launch), static_cast<const uint16_t*>(x), y, HW, C, total)));
    SPG_LAUNCHED();
    return SPG_OK;
}
