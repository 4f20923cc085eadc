// Input / output side of the inference path: the reference's image preprocessing and prediction post-processing on
// the GPU (HBM-bound byte / fp32 work, one pass per stage).
//
//   spg_preprocess_rgb_u8     CODImageProcessor.process_image (utils/image_processor.py:114-134): uint8 HWC RGB ->
//                             /255 -> antialiased bilinear resize to S x S (F.interpolate(..., antialias=True)) ->
//                             (x - mean) / std, fp32 CHW.  Restates ATen's separable anti-aliasing resampler
//                             (aten/src/ATen/native/cpu/UpSampleKernel.cpp, _compute_indices_min_size_weights_aa:
//                             triangle filter of support max(in/out, 1), width first, then height), including the
//                             float / double promotion of its index arithmetic, so that window bounds and weights
//                             are the ones ATen computes.
//   spg_resize_bilinear_f32   F.interpolate(logits, size=(h, w), mode='bilinear', align_corners=False) [+ sigmoid]:
//                             the per-image resize of the finest prediction to the original / ground-truth size
//                             (engine/predictor.py:350-365, engine/evaluator.py:539-544).
#include <atomic>

#include "common.h"

namespace spg {
extern std::atomic<long long> g_launches;

namespace {

#define SPG_LAUNCHED()                                        \
    do {                                                      \
        g_launches.fetch_add(1, std::memory_order_relaxed);   \
        SPG_CHECK_LAUNCH();                                   \
    } while (0)

// Window [xmin, xmin + xsize) and filter parameters of output index i, exactly as ATen computes them for
// scalar_t = float: the "+ 0.5" literals are doubles, so those expressions are evaluated in double and rounded.
struct AaWindow {
    int xmin, xsize;
    float center, invscale;
};
__device__ __forceinline__ AaWindow aa_window(int i, int in_size, float scale) {
    AaWindow w;
    const float support = scale >= 1.0f ? static_cast<float>(1.0 * static_cast<double>(scale)) : 1.0f;
    w.center = static_cast<float>(static_cast<double>(scale) * (static_cast<double>(i) + 0.5));
    w.invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
    const long long lo = static_cast<long long>(static_cast<double>(w.center) - static_cast<double>(support) + 0.5);
    const long long hi = static_cast<long long>(static_cast<double>(w.center) + static_cast<double>(support) + 0.5);
    w.xmin = static_cast<int>(lo > 0 ? lo : 0);
    w.xsize = static_cast<int>((hi < in_size ? hi : in_size) - w.xmin);
    return w;
}
__device__ __forceinline__ float aa_weight(const AaWindow& w, int j) {
    const double arg = (static_cast<double>(static_cast<float>(static_cast<long long>(j) + w.xmin) - w.center) + 0.5) *
                       static_cast<double>(w.invscale);
    const float x = fabsf(static_cast<float>(arg));
    return x < 1.0f ? 1.0f - x : 0.0f;
}

// Horizontal pass: uint8 HWC -> fp32 [3][H][So] (values / 255 resampled along x).  Thread = (y, xo), 3 channels.
__global__ void __launch_bounds__(256) aa_rows_kernel(const uint8_t* __restrict__ img, int H, int W, float* __restrict__ tmp,
                                                      int So, float scale) {
    pdl_prologue();
    const long long idx = blockIdx.x * 256ll + threadIdx.x;
    if (idx >= static_cast<long long>(H) * So) return;
    const int xo = idx % So, y = idx / So;
    const uint8_t* row = img + static_cast<size_t>(y) * W * 3;
    float r = 0.f, g = 0.f, b = 0.f;
    if (W == So) {  // ATen skips a dimension whose size does not change
        r = row[3 * xo] / 255.0f;
        g = row[3 * xo + 1] / 255.0f;
        b = row[3 * xo + 2] / 255.0f;
    } else {
        const AaWindow w = aa_window(xo, W, scale);
        float total = 0.f;
        for (int j = 0; j < w.xsize; ++j) total += aa_weight(w, j);
        for (int j = 0; j < w.xsize; ++j) {
            float wt = aa_weight(w, j);
            if (total != 0.f) wt /= total;
            const uint8_t* px = row + 3 * (w.xmin + j);
            r += (px[0] / 255.0f) * wt;
            g += (px[1] / 255.0f) * wt;
            b += (px[2] / 255.0f) * wt;
        }
    }
    const size_t plane = static_cast<size_t>(H) * So;
    tmp[idx] = r;
    tmp[plane + idx] = g;
    tmp[2 * plane + idx] = b;
}

// Vertical pass + normalisation: fp32 [3][H][So] -> fp32 [3][So][So].  Thread = (c, yo, xo), coalesced along x.
__global__ void __launch_bounds__(256) aa_cols_kernel(const float* __restrict__ tmp, int H, int So, float* __restrict__ out,
                                                      float scale, float m0, float m1, float m2, float s0, float s1,
                                                      float s2) {
    pdl_prologue();
    const long long idx = blockIdx.x * 256ll + threadIdx.x;
    if (idx >= 3ll * So * So) return;
    const int xo = idx % So;
    const int yo = (idx / So) % So;
    const int c = idx / (static_cast<long long>(So) * So);
    const float* plane = tmp + static_cast<size_t>(c) * H * So;
    float v = 0.f;
    if (H == So) {
        v = plane[static_cast<size_t>(yo) * So + xo];
    } else {
        const AaWindow w = aa_window(yo, H, scale);
        float total = 0.f;
        for (int j = 0; j < w.xsize; ++j) total += aa_weight(w, j);
        for (int j = 0; j < w.xsize; ++j) {
            float wt = aa_weight(w, j);
            if (total != 0.f) wt /= total;
            v += plane[static_cast<size_t>(w.xmin + j) * So + xo] * wt;
        }
    }
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
    const float sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    out[idx] = (v - mean) / sd;
}

// Bilinear resize of fp32 maps, align_corners=False, ATen's coordinates: src = max(0, (dst + 0.5) * in/out - 0.5).
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float* __restrict__ src, int hi, int wi,
                                                              float* __restrict__ dst, int ho, int wo, int apply_sigmoid) {
    pdl_prologue();
    const int b = blockIdx.y;
    const long long idx = blockIdx.x * 256ll + threadIdx.x;
    if (idx >= static_cast<long long>(ho) * wo) return;
    const int x = idx % wo, y = idx / wo;
    const float sy = static_cast<float>(hi) / static_cast<float>(ho), sx = static_cast<float>(wi) / static_cast<float>(wo);
    float fy = (y + 0.5f) * sy - 0.5f, fx = (x + 0.5f) * sx - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const int y1 = y0 + (y0 < hi - 1 ? 1 : 0), x1 = x0 + (x0 < wi - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0;
    const float* s = src + static_cast<size_t>(b) * hi * wi;
    float v = (1.f - ly) * ((1.f - lx) * s[static_cast<size_t>(y0) * wi + x0] + lx * s[static_cast<size_t>(y0) * wi + x1]) +
              ly * ((1.f - lx) * s[static_cast<size_t>(y1) * wi + x0] + lx * s[static_cast<size_t>(y1) * wi + x1]);
    if (apply_sigmoid) v = 1.f / (1.f + expf(-v));
    dst[static_cast<size_t>(b) * ho * wo + idx] = v;
}

}  // namespace
}  // namespace spg

using namespace spg;

extern "C" size_t spg_preprocess_workspace_bytes(int H, int W, int S) {
    if (H <= 0 || W <= 0 || S <= 0) return 0;
    return 3ull * H * S * sizeof(float);
}

extern "C" int spg_preprocess_rgb_u8(const unsigned char* img, int H, int W, float* out, int S, const float* mean3,
                                     const float* std3, void* workspace, size_t ws_bytes, const spg_launch_t* launch) {
    SPG_CHECK_ARG(img && out && mean3 && std3 && workspace, "null pointer");
    SPG_CHECK_ARG(H > 0 && W > 0 && S > 0, "bad image shape H=%d W=%d S=%d", H, W, S);
    SPG_CHECK_ARG(ws_bytes >= spg_preprocess_workspace_bytes(H, W, S), "workspace too small (%zu bytes)", ws_bytes);
    const LaunchCtx st(launch);
    float* tmp = static_cast<float*>(workspace);
    const float sx = static_cast<float>(W) / static_cast<float>(S), sy = static_cast<float>(H) / static_cast<float>(S);
    const long long n1 = static_cast<long long>(H) * S;
    SPG_CHECK_CUDA((launch_pdl(aa_rows_kernel, static_cast<unsigned>((n1 + 255) / 256), 256, 0, st, img, H, W, tmp, S, sx)));
    SPG_LAUNCHED();
    const long long n2 = 3ll * S * S;
    SPG_CHECK_CUDA((launch_pdl(aa_cols_kernel, static_cast<unsigned>((n2 + 255) / 256), 256, 0, st, tmp, H, S, out, sy, mean3[0], mean3[1], mean3[2],
                                                                           std3[0], std3[1], std3[2])));
    SPG_LAUNCHED();
    return SPG_OK;
}

extern "C" int spg_resize_bilinear_f32(const float* src, int B, int Hi, int Wi, float* dst, int Ho, int Wo,
                                       int apply_sigmoid, const spg_launch_t* launch) {
    SPG_CHECK_ARG(src && dst, "null pointer");
    SPG_CHECK_ARG(B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "bad shape");
    const long long n = static_cast<long long>(Ho) * Wo;
    SPG_CHECK_CUDA((launch_pdl(resize_bilinear_kernel, dim3(static_cast<unsigned>((n + 255) / 256), B), 256, 0, LaunchCtx(launch), src, Hi, Wi, dst, Ho, Wo, apply_sigmoid)));
    SPG_LAUNCHED();
    return SPG_OK;
}
