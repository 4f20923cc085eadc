// Camouflaged-object scores on the GPU: S-alpha, weighted F-beta, MAE, adaptive E-phi and the mean of the F-beta
// curve for a batch of (uint8 prediction, uint8 ground truth) pairs, in fp64, without the per-image device->host
// copy + process pool of the reference (utils/metrics.py:142-167,199-231; algorithms: py_sod_metrics, restated on
// the CPU in oracle/sod_metrics.py).
//
// Structure (all HBM-bound integer / fp64 work, 1 byte per pixel per operand):
//   spg_sod_gt_prepare_u8   GT only (cacheable per dataset): exact Euclidean feature transform = index of the
//                           nearest foreground pixel for every pixel, with the SAME tie-breaking as
//                           scipy.ndimage.distance_transform_edt (a column scan, then scipy's Voronoi row scan in
//                           integer arithmetic), plus {#fg, sum y, sum x} for the S-measure centroid.
//   spg_sod_scores_u8       pass 1: joint histogram  hist[quadrant(4)][fg(2)][grey level(256)]  per image.  After the
//                           min-max normalisation a prediction takes <= 256 distinct values, so MAE, S-alpha (object
//                           + 4-quadrant SSIM), adaptive E-phi and the 256-threshold F-beta curve are exact functions
//                           of this histogram; pass 2: the weighted F-beta pixel pass (error transfer through the
//                           feature transform, 7x7 Gaussian, distance weights) with per-tile fp64 partials summed in
//                           a fixed order; pass 3: one CTA per image turns histogram + partials into the five scores.
#include <atomic>
#include <cmath>

#include "common.h"

namespace spg {
extern std::atomic<long long> g_launches;

namespace {

#define SPG_LAUNCHED()                                        \
    do {                                                      \
        g_launches.fetch_add(1, std::memory_order_relaxed);   \
        SPG_CHECK_LAUNCH();                                   \
    } while (0)

constexpr double kEps = 2.220446049250313e-16;  // np.spacing(1)
constexpr int kTile = 32;                       // weighted-F tile edge
constexpr int kHalo = 3;                        // 7x7 Gaussian
constexpr int kTileH = kTile + 2 * kHalo;

__constant__ double c_gauss7[49];

// ---------------------------------------------------------------------------------------------------------
// GT statistics: {#fg, sum of fg rows, sum of fg columns}
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gt_stats_kernel(const uint8_t* __restrict__ gt, int H, int W,
                                                       unsigned long long* __restrict__ stats) {
    pdl_prologue();
    const int b = blockIdx.y;
    const int HW = H * W;
    unsigned long long n = 0, sy = 0, sx = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) {
        if (gt[static_cast<size_t>(b) * HW + i] > 128) {
            ++n;
            sy += i / W;
            sx += i % W;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
    }
    if ((threadIdx.x & 31) == 0 && n) {
        atomicAdd(stats + b * 4 + 0, n);
        atomicAdd(stats + b * 4 + 1, sy);
        atomicAdd(stats + b * 4 + 2, sx);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Feature transform, pass 1: for every pixel the row of the nearest foreground pixel IN ITS COLUMN (-1: none);
// equidistant candidates resolve to the smaller row, as scipy's 1-D scan does.  Thread = column (coalesced rows).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ft_columns_kernel(const uint8_t* __restrict__ gt, int H, int W,
                                                         short* __restrict__ colfeat) {
    pdl_prologue();
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int b = blockIdx.y;
    if (x >= W) return;
    const uint8_t* g = gt + static_cast<size_t>(b) * H * W + x;
    short* f = colfeat + static_cast<size_t>(b) * H * W + x;
    int last = -1;
    for (int y = 0; y < H; ++y) {  // nearest foreground row at or above
        if (g[static_cast<size_t>(y) * W] > 128) last = y;
        f[static_cast<size_t>(y) * W] = static_cast<short>(last);
    }
    int next = -1;
    for (int y = H - 1; y >= 0; --y) {  // combine with the nearest at or below; ties keep the one above
        if (g[static_cast<size_t>(y) * W] > 128) next = y;
        const int up = f[static_cast<size_t>(y) * W];
        int best = up;
        if (next >= 0 && (up < 0 || next - y < y - up)) best = next;
        f[static_cast<size_t>(y) * W] = static_cast<short>(best);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Feature transform, pass 2: scipy's `_VoronoiFT` along a row, restated in integer arithmetic.  Sites are the
// columns ii whose column feature exists, at vertical offset fy[ii] - y.  A site stays on the stack unless it is
// strictly dominated (scipy's `c*vR - b*uR - a*wR - a*b*c <= 0 -> keep`); the query scan advances only to a
// strictly closer site.  Both rules decide which of several equidistant foreground pixels is reported, and the
// weighted F-measure reads the prediction error AT that pixel.  Thread = row; the stack lives in `stack`.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) ft_rows_kernel(const short* __restrict__ colfeat, int H, int W,
                                                     short* __restrict__ stack, int* __restrict__ nearest) {
    pdl_prologue();
    const int y = blockIdx.x * 64 + threadIdx.x;
    const int b = blockIdx.y;
    if (y >= H) return;
    const size_t row = (static_cast<size_t>(b) * H + y) * W;
    const short* fy = colfeat + row;
    short* g = stack + row;
    int* out = nearest + row;
    int l = -1;
    for (int ii = 0; ii < W; ++ii) {
        const int f = fy[ii];
        if (f < 0) continue;
        const long long wR = static_cast<long long>(f - y) * (f - y);
        while (l >= 1) {
            const int i1 = g[l], i2 = g[l - 1];
            const long long a = i1 - i2, bb = ii - i1, c = a + bb;
            const long long d2 = fy[i2] - y, d1 = fy[i1] - y;
            const long long uR = d2 * d2, vR = d1 * d1;
            if (c * vR - bb * uR - a * wR - a * bb * c <= 0) break;
            --l;
        }
        g[++l] = static_cast<short>(ii);
    }
    const int maxl = l;
    if (maxl < 0) {
        for (int ii = 0; ii < W; ++ii) out[ii] = -1;
        return;
    }
    l = 0;
    for (int ii = 0; ii < W; ++ii) {
        int gx = g[l];
        long long dy = fy[gx] - y, dx = gx - ii;
        long long delta1 = dy * dy + dx * dx;
        while (l < maxl) {
            const int nx = g[l + 1];
            dy = fy[nx] - y;
            dx = nx - ii;
            const long long delta2 = dy * dy + dx * dx;
            if (delta1 <= delta2) break;
            delta1 = delta2;
            ++l;
            gx = nx;
        }
        out[ii] = static_cast<int>(fy[gx]) * W + gx;
    }
}

// centroid of the ground truth as py_sod_metrics' Smeasure computes it (round half to even, then + 1)
__device__ __forceinline__ void centroid(const unsigned long long* st, int H, int W, int& cx, int& cy) {
    const unsigned long long n = st[0];
    if (n == 0) {
        cx = static_cast<int>(rint(W / 2.0)) + 1;
        cy = static_cast<int>(rint(H / 2.0)) + 1;
    } else {
        cy = static_cast<int>(rint(static_cast<double>(st[1]) / static_cast<double>(n))) + 1;
        cx = static_cast<int>(rint(static_cast<double>(st[2]) / static_cast<double>(n))) + 1;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Pass 1: joint histogram hist[b][quadrant][fg][grey level] (uint32).  Shared-memory privatised, 16 pixels / thread
// per iteration (one 16-byte load of each operand).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sod_hist_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                                                       const unsigned long long* __restrict__ gt_stats, int H, int W,
                                                       unsigned* __restrict__ hist) {
    pdl_prologue();
    __shared__ unsigned sh[8 * 256];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) sh[i] = 0;
    int cx, cy;
    centroid(gt_stats + b * 4, H, W, cx, cy);
    __syncthreads();
    const int HW = H * W;
    const uint8_t* p = pred + static_cast<size_t>(b) * HW;
    const uint8_t* g = gt + static_cast<size_t>(b) * HW;
    if ((W & 15) == 0) {
        const int n16 = HW >> 4;
        for (int v = blockIdx.x * 256 + threadIdx.x; v < n16; v += gridDim.x * 256) {
            const uint4 pv = reinterpret_cast<const uint4*>(p)[v];
            const uint4 gv = reinterpret_cast<const uint4*>(g)[v];
            const int i0 = v << 4;
            const int r = i0 / W, c0 = i0 - r * W;
            const int qrow = (r >= cy) ? 2 : 0;
            const unsigned pw[4] = {pv.x, pv.y, pv.z, pv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const unsigned q = (pw[k >> 2] >> (8 * (k & 3))) & 255u;
                const unsigned gg = (gw[k >> 2] >> (8 * (k & 3))) & 255u;
                const int quad = qrow + ((c0 + k >= cx) ? 1 : 0);
                atomicAdd(&sh[((quad << 1) + (gg > 128 ? 1 : 0)) * 256 + q], 1u);
            }
        }
    } else {
        for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) {
            const int r = i / W, c = i - r * W;
            const int quad = ((r >= cy) ? 2 : 0) + ((c >= cx) ? 1 : 0);
            atomicAdd(&sh[((quad << 1) + (g[i] > 128 ? 1 : 0)) * 256 + p[i]], 1u);
        }
    }
    __syncthreads();
    unsigned* out = hist + static_cast<size_t>(b) * 8 * 256;
    for (int i = threadIdx.x; i < 8 * 256; i += 256)
        if (sh[i]) atomicAdd(out + i, sh[i]);
}

// grey level -> normalised prediction value, exactly the sequence of IEEE operations of `prepare` in the oracle
// (pred / 255, then (pred - min) / (max - min) unless constant)
__device__ __forceinline__ double level_value(int q, int lo, int hi) {
    const double v = q / 255.0;
    if (hi == lo) return v;
    const double l = lo / 255.0, h = hi / 255.0;
    return (v - l) / (h - l);
}

// min / max occupied grey level of image b from its histogram (block-wide, 256 threads)
__device__ __forceinline__ void level_range(const unsigned* hist_b, int& lo, int& hi, int* s_lo, int* s_hi) {
    const int t = threadIdx.x;
    if (t == 0) {
        *s_lo = 256;
        *s_hi = -1;
    }
    __syncthreads();
    if (t < 256) {
        unsigned any = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) any |= hist_b[k * 256 + t];
        if (any) {
            atomicMin(s_lo, t);
            atomicMax(s_hi, t);
        }
    }
    __syncthreads();
    lo = *s_lo;
    hi = *s_hi;
}

// ---------------------------------------------------------------------------------------------------------
// Pass 2: weighted F-measure pixel pass (Margolin et al.; oracle/sod_metrics.py::weighted_f).  One CTA per 32x32
// tile: Et (the error, copied from the nearest foreground pixel for background pixels) for the tile + 3-pixel halo
// goes to shared memory, every thread filters 4 pixels with the 7x7 Gaussian, applies min(E, EA) on the foreground
// and the distance weight 2 - exp(ln(.5)/5 * dist) on the background, and the tile's two sums are written as fp64
// partials (no floating-point atomics: the final order of summation is fixed).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sod_wfm_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                                                      const int* __restrict__ nearest, const unsigned* __restrict__ hist,
                                                      const unsigned long long* __restrict__ gt_stats, int H, int W,
                                                      double* __restrict__ partials) {
    pdl_prologue();
    __shared__ double s_val[256];
    __shared__ double s_et[kTileH][kTileH + 1];
    __shared__ double s_red[2][8];
    __shared__ int s_lo, s_hi;
    const int b = blockIdx.z;
    if (gt_stats[b * 4] == 0) return;  // empty ground truth: weighted F is 0 by definition, `nearest` holds -1
    const int HW = H * W;
    const uint8_t* p = pred + static_cast<size_t>(b) * HW;
    const uint8_t* g = gt + static_cast<size_t>(b) * HW;
    const int* nn = nearest + static_cast<size_t>(b) * HW;
    int lo, hi;
    level_range(hist + static_cast<size_t>(b) * 8 * 256, lo, hi, &s_lo, &s_hi);
    s_val[threadIdx.x] = level_value(threadIdx.x, lo, hi);
    __syncthreads();
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    for (int i = threadIdx.x; i < kTileH * kTileH; i += 256) {
        const int ty = i / kTileH, tx = i - ty * kTileH;
        const int y = y0 + ty - kHalo, x = x0 + tx - kHalo;
        double et = 0.0;  // zero padding of scipy.ndimage.convolve(mode="constant")
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const int idx = y * W + x;
            const int src = g[idx] > 128 ? idx : nn[idx];
            et = 1.0 - s_val[p[src]];  // |pred - 1| at a foreground pixel
        }
        s_et[ty][tx] = et;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, tyb = threadIdx.x >> 5;
    double sum_fg = 0.0, sum_bg = 0.0;
    const double kdecay = log(0.5) / 5.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int ty = tyb + 8 * r;
        const int y = y0 + ty, x = x0 + tx;
        if (y < H && x < W) {
            double ea = 0.0;
#pragma unroll
            for (int i = 0; i < 7; ++i)
#pragma unroll
                for (int j = 0; j < 7; ++j) ea += c_gauss7[i * 7 + j] * s_et[ty + i][tx + j];
            const int idx = y * W + x;
            const double pv = s_val[p[idx]];
            if (g[idx] > 128) {
                const double e = 1.0 - pv;
                sum_fg += ea < e ? ea : e;
            } else {
                const int src = nn[idx];
                const int fy = src / W, fx = src - fy * W;
                const double d2 = static_cast<double>((fy - y) * (fy - y) + (fx - x) * (fx - x));
                sum_bg += pv * (2.0 - exp(kdecay * sqrt(d2)));
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum_fg += __shfl_xor_sync(0xffffffffu, sum_fg, o);
        sum_bg += __shfl_xor_sync(0xffffffffu, sum_bg, o);
    }
    if (tx == 0) {
        s_red[0][tyb] = sum_fg;
        s_red[1][tyb] = sum_bg;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int k = 0; k < 8; ++k) {
            a += s_red[0][k];
            c += s_red[1][k];
        }
        const size_t tile = (static_cast<size_t>(b) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        partials[2 * tile] = a;
        partials[2 * tile + 1] = c;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Pass 3: histogram + partials -> {S-alpha, weighted F, MAE, adaptive E, mean of the F curve}.  One CTA per image,
// thread = grey level; the handful of 256-term fp64 sums are block reductions in a fixed tree order.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* s_buf) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s_buf[k];
    return t;
}

// area-weighted SSIM of one quadrant (py_sod_metrics Smeasure.ssim) from its two 256-bin histograms
__device__ double quadrant_ssim(const unsigned* hq /* [2][256] */, double val, double* s_buf) {
    const int t = threadIdx.x;
    const double cb = hq[t], cf = hq[256 + t];
    const double n = block_sum(cb + cf, s_buf);
    if (n < 2.0) return 0.0;  // empty or single-pixel quadrant: skipped by the oracle
    const double nf = block_sum(cf, s_buf);
    const double x = block_sum((cb + cf) * val, s_buf) / n;
    const double y = nf / n;
    const double dx = val - x;
    const double sx = block_sum((cb + cf) * dx * dx, s_buf) / (n - 1.0);
    const double sy = ((n - nf) * y * y + nf * (1.0 - y) * (1.0 - y)) / (n - 1.0);
    const double sxy = block_sum(dx * (cf * (1.0 - y) - cb * y), s_buf) / (n - 1.0);
    const double alpha = 4.0 * x * y * sxy;
    const double beta = (x * x + y * y) * (sx + sy);
    if (alpha != 0.0) return alpha / (beta + kEps);
    return beta == 0.0 ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(256) sod_finalize_kernel(const unsigned* __restrict__ hist,
                                                           const unsigned long long* __restrict__ gt_stats,
                                                           const double* __restrict__ partials, int tiles, int H, int W,
                                                           double* __restrict__ scores) {
    pdl_prologue();
    __shared__ double s_buf[8];
    __shared__ double s_fgh[256], s_bgh[256];
    __shared__ double s_f[256];
    __shared__ int s_lo, s_hi;
    const int b = blockIdx.x, t = threadIdx.x;
    const unsigned* hb = hist + static_cast<size_t>(b) * 8 * 256;
    int lo, hi;
    level_range(hb, lo, hi, &s_lo, &s_hi);
    const double val = level_value(t, lo, hi);
    double cb = 0.0, cf = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        cb += hb[(2 * q) * 256 + t];
        cf += hb[(2 * q + 1) * 256 + t];
    }
    const double N = static_cast<double>(H) * W;
    const double nf = block_sum(cf, s_buf);
    const double nb = N - nf;
    const double sum_p = block_sum((cb + cf) * val, s_buf);
    const double mean_p = sum_p / N;

    // ---- MAE
    const double mae = block_sum(cb * val + cf * (1.0 - val), s_buf) / N;

    // ---- S-measure (alpha = 0.5)
    double sm;
    if (nf == 0.0) {
        sm = 1.0 - mean_p;
    } else if (nf == N) {
        sm = mean_p;
    } else {
        const double mu_f = block_sum(cf * val, s_buf) / nf;
        const double sig_f = sqrt(block_sum(cf * (val - mu_f) * (val - mu_f), s_buf) / (nf - 1.0));
        const double o_f = 2.0 * mu_f / (mu_f * mu_f + 1.0 + sig_f + kEps);
        const double mu_b = block_sum(cb * (1.0 - val), s_buf) / nb;
        const double db = (1.0 - val) - mu_b;
        const double sig_b = sqrt(block_sum(cb * db * db, s_buf) / (nb - 1.0));
        const double o_b = 2.0 * mu_b / (mu_b * mu_b + 1.0 + sig_b + kEps);
        const double u = nf / N;
        const double s_obj = u * o_f + (1.0 - u) * o_b;
        int cx, cy;
        centroid(gt_stats + b * 4, H, W, cx, cy);
        const double area = N;
        const double w1 = static_cast<double>(cx) * cy / area;
        const double w2 = static_cast<double>(cy) * (W - cx) / area;
        const double w3 = static_cast<double>(H - cy) * cx / area;
        const double w4 = 1.0 - w1 - w2 - w3;
        double s_reg = 0.0;
        s_reg += w1 * quadrant_ssim(hb + 0 * 512, val, s_buf);
        s_reg += w2 * quadrant_ssim(hb + 1 * 512, val, s_buf);
        s_reg += w3 * quadrant_ssim(hb + 2 * 512, val, s_buf);
        s_reg += w4 * quadrant_ssim(hb + 3 * 512, val, s_buf);
        sm = 0.5 * s_obj + 0.5 * s_reg;
        sm = sm > 0.0 ? sm : 0.0;
    }

    // ---- adaptive E-measure
    const double thr = fmin(2.0 * mean_p, 1.0);
    const bool on = val >= thr;
    const double fg_fg = block_sum(on ? cf : 0.0, s_buf);
    const double fg_bg = block_sum(on ? cb : 0.0, s_buf);
    double em;
    {
        const double n_pred_fg = fg_fg + fg_bg, n_pred_bg = N - n_pred_fg;
        double total;
        if (nf == 0.0) {
            total = n_pred_bg;
        } else if (nf == N) {
            total = n_pred_fg;
        } else {
            const double bg_fg = nf - fg_fg, bg_bg = n_pred_bg - bg_fg;
            const double mp = n_pred_fg / N, mg = nf / N;
            const double a4[4] = {1.0 - mp, 1.0 - mp, -mp, -mp};
            const double b4[4] = {1.0 - mg, -mg, 1.0 - mg, -mg};
            const double c4[4] = {fg_fg, fg_bg, bg_fg, bg_bg};
            total = 0.0;
            for (int k = 0; k < 4; ++k) {
                const double align = 2.0 * a4[k] * b4[k] / (a4[k] * a4[k] + b4[k] * b4[k] + kEps);
                total += (align + 1.0) * (align + 1.0) / 4.0 * c4[k];
            }
        }
        em = total / (N - 1.0 + kEps);
    }

    // ---- F-measure curve (beta^2 = 0.3): 256-bin histograms of uint8(pred * 255) over fg / bg
    s_fgh[t] = 0.0;
    s_bgh[t] = 0.0;
    __syncthreads();
    {
        const int bin = static_cast<int>(val * 255.0) & 255;  // astype(np.uint8): truncation
        if (cf != 0.0) atomicAdd(&s_fgh[bin], cf);             // integer-valued doubles: order-independent
        if (cb != 0.0) atomicAdd(&s_bgh[bin], cb);
    }
    __syncthreads();
    {
        double tp = 0.0, ps = 0.0;  // reversed cumulative sums up to bin 255 - t
        for (int k = 255; k >= 255 - t; --k) {
            tp += s_fgh[k];
            ps += s_bgh[k];
        }
        double pos = tp + ps;
        if (pos == 0.0) pos = 1.0;
        const double tot = nf > 1.0 ? nf : 1.0;
        const double precision = tp / pos, recall = tp / tot;
        const double num = (1.0 + 0.3) * precision * recall;
        const double den = num == 0.0 ? 1.0 : 0.3 * precision + recall;
        s_f[t] = num / den;
    }
    __syncthreads();

    if (t == 0) {
        double fm = 0.0;
        for (int k = 0; k < 256; ++k) fm += s_f[k];
        fm /= 256.0;
        // ---- weighted F (beta^2 = 1)
        double wfm = 0.0;
        if (nf > 0.0) {
            double ew_fg = 0.0, ew_bg = 0.0;
            const double* pp = partials + static_cast<size_t>(b) * tiles * 2;
            for (int k = 0; k < tiles; ++k) {
                ew_fg += pp[2 * k];
                ew_bg += pp[2 * k + 1];
            }
            const double tpw = nf - ew_fg, fpw = ew_bg;
            const double recall = 1.0 - ew_fg / nf;
            const double precision = tpw / (tpw + fpw + kEps);
            wfm = 2.0 * recall * precision / (recall + precision + kEps);
        }
        double* o = scores + static_cast<size_t>(b) * 5;
        o[0] = sm;
        o[1] = wfm;
        o[2] = mae;
        o[3] = em;
        o[4] = fm;
    }
}

// matlab-style fspecial('gaussian', 7, 5), as oracle/sod_metrics.py::_gauss7
int upload_gauss7(cudaStream_t st) {
    static PerDeviceOnce once;  // c_gauss7 lives in each device's context
    if (!once.needed()) return SPG_OK;
    double k[49], mx = 0.0, sum = 0.0;
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) {
            const double y = i - 3, x = j - 3;
            k[i * 7 + j] = std::exp(-(x * x + y * y) / (2.0 * 5.0 * 5.0));
            mx = k[i * 7 + j] > mx ? k[i * 7 + j] : mx;
        }
    for (int i = 0; i < 49; ++i) {
        if (k[i] < kEps * mx) k[i] = 0.0;
        sum += k[i];
    }
    if (sum != 0.0)
        for (int i = 0; i < 49; ++i) k[i] /= sum;
    SPG_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_gauss7, k, sizeof(k), 0, cudaMemcpyHostToDevice, st));
    once.done();
    return SPG_OK;
}

}  // namespace
}  // namespace spg

using namespace spg;

extern "C" size_t spg_sod_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    const size_t hw = static_cast<size_t>(B) * H * W;
    const size_t tiles = static_cast<size_t>(B) * ((H + kTile - 1) / kTile) * ((W + kTile - 1) / kTile);
    // gt_prepare: column features + row stacks (2 x int16 per pixel); scores: histogram + tile partials
    const size_t a = hw * 2 * sizeof(short);
    const size_t b = static_cast<size_t>(B) * 8 * 256 * sizeof(unsigned) + tiles * 2 * sizeof(double);
    return (a > b ? a : b) + 256;
}

extern "C" int spg_sod_gt_prepare_u8(const unsigned char* gt, int B, int H, int W, int* nearest,
                                     unsigned long long* gt_stats, void* workspace, size_t ws_bytes,
                                     const spg_launch_t* launch) {
    SPG_CHECK_ARG(gt && nearest && gt_stats && workspace, "null pointer");
    SPG_CHECK_ARG(B > 0 && H > 0 && W > 0 && H <= 32767 && W <= 32767, "bad ground-truth shape B=%d H=%d W=%d", B, H, W);
    SPG_CHECK_ARG(static_cast<long long>(H) * W < (1ll << 31), "image too large");
    SPG_CHECK_ARG(ws_bytes >= spg_sod_workspace_bytes(B, H, W), "workspace too small (%zu bytes)", ws_bytes);
    const LaunchCtx st(launch);
    const size_t hw = static_cast<size_t>(B) * H * W;
    short* colfeat = static_cast<short*>(workspace);
    short* stack = colfeat + hw;
    SPG_CHECK_CUDA(cudaMemsetAsync(gt_stats, 0, static_cast<size_t>(B) * 4 * sizeof(unsigned long long), st));
    const int per_img = min(32, (H * W + 256 * 16 - 1) / (256 * 16));
    SPG_CHECK_CUDA((launch_pdl(gt_stats_kernel, dim3(per_img, B), 256, 0, st, gt, H, W, gt_stats)));
    SPG_LAUNCHED();
    SPG_CHECK_CUDA((launch_pdl(ft_columns_kernel, dim3((W + 127) / 128, B), 128, 0, st, gt, H, W, colfeat)));
    SPG_LAUNCHED();
    SPG_CHECK_CUDA((launch_pdl(ft_rows_kernel, dim3((H + 63) / 64, B), 64, 0, st, colfeat, H, W, stack, nearest)));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_sod_scores_u8(const unsigned char* pred, const unsigned char* gt, const int* nearest,
                                 const unsigned long long* gt_stats, int B, int H, int W, double* scores,
                                 void* workspace, size_t ws_bytes, const spg_launch_t* launch) {
    SPG_CHECK_ARG(pred && gt && nearest && gt_stats && scores && workspace, "null pointer");
    SPG_CHECK_ARG(B > 0 && H > 0 && W > 0, "bad shape B=%d H=%d W=%d", B, H, W);
    SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(pred) & 15) == 0 && (reinterpret_cast<uintptr_t>(gt) & 15) == 0,
                  "pred / gt must be 16-byte aligned");
    SPG_CHECK_ARG(ws_bytes >= spg_sod_workspace_bytes(B, H, W), "workspace too small (%zu bytes)", ws_bytes);
    const LaunchCtx st(launch);
    if (int rc = upload_gauss7(st)) return rc;
    unsigned* hist = static_cast<unsigned*>(workspace);
    const size_t hist_bytes = static_cast<size_t>(B) * 8 * 256 * sizeof(unsigned);
    double* partials = reinterpret_cast<double*>(static_cast<char*>(workspace) + ((hist_bytes + 255) & ~size_t(255)));
    const int tx = (W + kTile - 1) / kTile, ty = (H + kTile - 1) / kTile;
    SPG_CHECK_CUDA(cudaMemsetAsync(hist, 0, hist_bytes, st));
    const int per_img = min(32, (H * W + 256 * 16 - 1) / (256 * 16));
    SPG_CHECK_CUDA((launch_pdl(sod_hist_kernel, dim3(per_img, B), 256, 0, st, pred, gt, gt_stats, H, W, hist)));
    SPG_LAUNCHED();
    SPG_CHECK_CUDA((launch_pdl(sod_wfm_kernel, dim3(tx, ty, B), 256, 0, st, pred, gt, nearest, hist, gt_stats, H, W, partials)));
    SPG_LAUNCHED();
    SPG_CHECK_CUDA((launch_pdl(sod_finalize_kernel, B, 256, 0, st, hist, gt_stats, partials, tx * ty, H, W, scores)));
    SPG_LAUNCHED();
    return SPG_OK;
}
