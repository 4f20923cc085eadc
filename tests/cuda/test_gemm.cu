// Standalone GPU check of the tcgen05 GEMM / conv engine against naive CUDA-core reference kernels
// (no torch: starts in milliseconds on a fresh box).  Prints one line per case with max error and
// the measured throughput; exit code = number of failed cases.
//   build: make -C spegnet_b200/csrc test_gemm       run: spegnet_b200/csrc/build/test_gemm_{fp16,bf16} [--perf]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/spegnet_b200.h"
#ifdef SPG_TRACE
extern "C" int spg_debug_trace_dump(long long*, int);
#endif

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(99);                                                                  \
        }                                                                              \
    } while (0)

// 16-bit storage type follows the library variant this binary is linked against (see half16.cuh)
#ifdef SPG_FP16
typedef __half h16;
static inline h16 to_h16(float f) { return __float2half(f); }
static inline float from_h16(h16 v) { return __half2float(v); }
__device__ inline float dev_from_h16(h16 v) { return __half2float(v); }
#else
typedef __nv_bfloat16 h16;
static inline h16 to_h16(float f) { return __float2bfloat16(f); }
static inline float from_h16(h16 v) { return __bfloat162float(v); }
__device__ inline float dev_from_h16(h16 v) { return __bfloat162float(v); }
#endif

static uint32_t g_seed = 12345u;
static float frand() {
    g_seed = g_seed * 1664525u + 1013904223u;
    return ((g_seed >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}

static h16* dev_bf16(size_t n, float scale) {
    std::vector<h16> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = to_h16(frand() * scale);
    h16* d;
    CK(cudaMalloc(&d, n * sizeof(h16)));
    CK(cudaMemcpy(d, h.data(), n * sizeof(h16), cudaMemcpyHostToDevice));
    return d;
}
static float* dev_f32(size_t n, float scale) {
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = frand() * scale;
    float* d;
    CK(cudaMalloc(&d, n * sizeof(float)));
    CK(cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    return d;
}

__device__ float ref_act(float v, int act) {
    if (act == SPG_ACT_RELU) return fmaxf(v, 0.f);
    if (act == SPG_ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    return v;
}

__global__ void ref_gemm(const h16* A, const h16* W, int M, int N, int K, const float* bias,
                         int act, const float* res, int res_rows, float* out) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)M * N) return;
    const int m = idx / N, n = idx % N;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += dev_from_h16(A[(size_t)m * K + k]) * dev_from_h16(W[(size_t)n * K + k]);
    if (bias) acc += bias[n];
    acc = ref_act(acc, act);
    if (res) acc += res[(size_t)(res_rows > 0 ? m % res_rows : m) * N + n];
    out[idx] = acc;
}

__global__ void ref_conv(const h16* X, const h16* Wt, int B, int H, int Wd, int Cin, int Cout,
                         const float* bias, int act, float* out) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)B * H * Wd * Cout) return;
    const int co = idx % Cout;
    const long long pix = idx / Cout;
    const int x = pix % Wd, y = (pix / Wd) % H, b = pix / ((long long)Wd * H);
    float acc = 0.f;
    for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
            const int yy = y + ky - 1, xx = x + kx - 1;
            if (yy < 0 || yy >= H || xx < 0 || xx >= Wd) continue;
            const h16* xp = X + (((size_t)b * H + yy) * Wd + xx) * Cin;
            const h16* wp = Wt + (size_t)co * 9 * Cin + (ky * 3 + kx) * Cin;
            for (int c = 0; c < Cin; ++c) acc += dev_from_h16(xp[c]) * dev_from_h16(wp[c]);
        }
    if (bias) acc += bias[co];
    out[idx] = ref_act(acc, act);
}

__global__ void ref_head(const float* full, int M, int N, const float* hw, float hb, float* out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float a = hb;
    for (int n = 0; n < N; ++n) a += full[(size_t)m * N + n] * hw[n];
    out[m] = a;
}

struct Stats {
    double max_abs = 0, max_ref = 0;
    long long bad = 0;
};

static Stats compare(const std::vector<float>& got, const std::vector<float>& ref, double atol, double rtol) {
    Stats s;
    for (size_t i = 0; i < got.size(); ++i) {
        const double d = fabs((double)got[i] - ref[i]);
        if (!(d <= atol + rtol * fabs(ref[i]))) {
            if (s.bad < 5) printf("    mismatch @%zu got %.6f ref %.6f\n", i, got[i], ref[i]);
            ++s.bad;
        }
        if (d > s.max_abs || d != d) s.max_abs = d;
        if (fabs(ref[i]) > s.max_ref) s.max_ref = fabs(ref[i]);
    }
    return s;
}

static std::vector<float> fetch_f32(const float* d, size_t n) {
    std::vector<float> h(n);
    CK(cudaMemcpy(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost));
    return h;
}
static std::vector<float> fetch_bf16(const h16* d, size_t n) {
    std::vector<h16> h(n);
    CK(cudaMemcpy(h.data(), d, n * sizeof(h16), cudaMemcpyDeviceToHost));
    std::vector<float> f(n);
    for (size_t i = 0; i < n; ++i) f[i] = from_h16(h[i]);
    return f;
}

static int g_fail = 0;
static bool g_perf = false;
static bool g_check = true;  // --perf-only skips the naive reference (for ncu / timing runs)

template <class F>
static float bench(F&& f, int iters) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) f();
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}

static void case_gemm(int M, int N, int K, int act, bool with_bias, bool with_res, int res_rows, bool out_f32,
                      bool with_head) {
    h16* A = dev_bf16((size_t)M * K, 1.0f);
    h16* W = dev_bf16((size_t)N * K, 1.0f / sqrtf((float)K));
    float* bias = with_bias ? dev_f32(N, 0.5f) : nullptr;
    float* res = with_res ? dev_f32((size_t)(res_rows > 0 ? res_rows : M) * N, 1.0f) : nullptr;
    float* hw = with_head ? dev_f32(N, 0.2f) : nullptr;
    float *ref, *ref_h = nullptr, *head_out = nullptr;
    CK(cudaMalloc(&ref, (size_t)M * N * 4));
    void* out;
    CK(cudaMalloc(&out, (size_t)M * N * (out_f32 ? 4 : 2)));
    CK(cudaMemset(out, 0xFF, (size_t)M * N * (out_f32 ? 4 : 2)));
    if (with_head) {
        CK(cudaMalloc(&ref_h, M * 4));
        CK(cudaMalloc(&head_out, M * 4));
    }
    const long long total = (long long)M * N;
    if (g_check) ref_gemm<<<(unsigned)((total + 255) / 256), 256>>>(A, W, M, N, K, bias, act, res, res_rows, ref);
    if (g_check && with_head) ref_head<<<(M + 127) / 128, 128>>>(ref, M, N, hw, 0.25f, ref_h);
    CK(cudaGetLastError());
    spg_epilogue_t ep{};
    ep.bias = bias;
    ep.act = act;
    ep.residual = res;
    ep.res_rows = res_rows;
    ep.out = out;
    ep.out_dtype = out_f32 ? SPG_F32 : SPG_H16;
    ep.head_w = hw;
    ep.head_b = 0.25f;
    ep.head_out = head_out;
    int rc = spg_linear_h16(A, W, M, N, K, &ep, nullptr);
    if (rc != SPG_OK) {
        printf("FAIL gemm M=%d N=%d K=%d: rc=%d %s\n", M, N, K, rc, spg_last_error());
        ++g_fail;
        return;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("FAIL gemm M=%d N=%d K=%d: kernel error %s\n", M, N, K, cudaGetErrorString(e));
        exit(100 + g_fail);
    }
    Stats s, sh;
    if (g_check) {
        auto r = fetch_f32(ref, (size_t)M * N);
        auto g = out_f32 ? fetch_f32((float*)out, (size_t)M * N) : fetch_bf16((h16*)out, (size_t)M * N);
        s = compare(g, r, out_f32 ? 2e-3 : 2e-2, out_f32 ? 1e-3 : 1e-2);
        if (with_head) sh = compare(fetch_f32(head_out, M), fetch_f32(ref_h, M), 5e-3, 2e-3);
    }
    float ms = 0;
    if (g_perf) ms = bench([&] { spg_linear_h16(A, W, M, N, K, &ep, nullptr); }, 10);
    const bool ok = s.bad == 0 && sh.bad == 0;
    printf("%s gemm M=%-6d N=%-5d K=%-5d act=%d bias=%d res=%d/%d f32=%d head=%d  max_abs=%.3e (ref max %.2f) head_err=%.3e",
           ok ? "PASS" : "FAIL", M, N, K, act, with_bias, with_res, res_rows, out_f32, with_head, s.max_abs, s.max_ref,
           sh.max_abs);
    if (g_perf) printf("  %.3f ms  %.1f TFLOP/s", ms, 2.0 * M * N * K / ms * 1e-9);
    printf("\n");
    if (!ok) ++g_fail;
    cudaFree(A); cudaFree(W); cudaFree(bias); cudaFree(res); cudaFree(hw); cudaFree(ref); cudaFree(out);
    cudaFree(ref_h); cudaFree(head_out);
}

static void case_conv(int B, int H, int Wd, int Cin, int Cout, int act, bool with_head, bool store_out) {
    const size_t M = (size_t)B * H * Wd;
    h16* X = dev_bf16(M * Cin, 1.0f);
    h16* Wt = dev_bf16((size_t)Cout * 9 * Cin, 1.0f / sqrtf(9.0f * Cin));
    float* bias = dev_f32(Cout, 0.5f);
    float* hw = with_head ? dev_f32(Cout, 0.2f) : nullptr;
    float *ref, *ref_h = nullptr, *head_out = nullptr;
    CK(cudaMalloc(&ref, M * Cout * 4));
    h16* out = nullptr;
    if (store_out) {
        CK(cudaMalloc(&out, M * Cout * 2));
        CK(cudaMemset(out, 0xFF, M * Cout * 2));
    }
    if (with_head) {
        CK(cudaMalloc(&ref_h, M * 4));
        CK(cudaMalloc(&head_out, M * 4));
    }
    const long long total = (long long)M * Cout;
    if (g_check) ref_conv<<<(unsigned)((total + 255) / 256), 256>>>(X, Wt, B, H, Wd, Cin, Cout, bias, act, ref);
    if (g_check && with_head) ref_head<<<(unsigned)((M + 127) / 128), 128>>>(ref, (int)M, Cout, hw, -0.1f, ref_h);
    CK(cudaGetLastError());
    spg_epilogue_t ep{};
    ep.bias = bias;
    ep.act = act;
    ep.out = out;
    ep.out_dtype = SPG_H16;
    ep.head_w = hw;
    ep.head_b = -0.1f;
    ep.head_out = head_out;
    int rc = spg_conv3x3_h16(X, Wt, B, H, Wd, Cin, Cout, &ep, nullptr);
    if (rc != SPG_OK) {
        printf("FAIL conv B=%d H=%d W=%d Cin=%d Cout=%d: rc=%d %s\n", B, H, Wd, Cin, Cout, rc, spg_last_error());
        ++g_fail;
        return;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("FAIL conv: kernel error %s\n", cudaGetErrorString(e));
        exit(100 + g_fail);
    }
    Stats s, sh;
    if (g_check && store_out) s = compare(fetch_bf16(out, M * Cout), fetch_f32(ref, M * Cout), 2e-2, 1e-2);
    if (g_check && with_head) sh = compare(fetch_f32(head_out, M), fetch_f32(ref_h, M), 5e-3, 2e-3);
    float ms = 0;
    if (g_perf) ms = bench([&] { spg_conv3x3_h16(X, Wt, B, H, Wd, Cin, Cout, &ep, nullptr); }, 10);
    const bool ok = s.bad == 0 && sh.bad == 0;
    printf("%s conv B=%d H=%-3d W=%-3d Cin=%-3d Cout=%-3d act=%d head=%d store=%d  max_abs=%.3e head_err=%.3e",
           ok ? "PASS" : "FAIL", B, H, Wd, Cin, Cout, act, with_head, store_out, s.max_abs, sh.max_abs);
    if (g_perf) printf("  %.3f ms  %.1f TFLOP/s", ms, 2.0 * M * Cout * 9.0 * Cin / ms * 1e-9);
    printf("\n");
    if (!ok) ++g_fail;
    cudaFree(X); cudaFree(Wt); cudaFree(bias); cudaFree(hw); cudaFree(ref); cudaFree(out); cudaFree(ref_h);
    cudaFree(head_out);
}

int main(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--perf")) g_perf = true;
        if (!strcmp(argv[i], "--perf-only")) { g_perf = true; g_check = false; }
    }
    if (spg_device_check() != SPG_OK) {
        printf("device check failed: %s\n", spg_last_error());
        return 98;
    }
    printf("libspegnet_b200 version %d\n", spg_version());
    // ad-hoc timing of one GEMM shape (ncu / tuning):  --shape M N K act(0|1|2) res(0|1) f32(0|1)
    for (int i = 1; i + 6 < argc; ++i) {
        if (!strcmp(argv[i], "--shape")) {
            g_perf = true;
            g_check = false;
            case_gemm(atoi(argv[i + 1]), atoi(argv[i + 2]), atoi(argv[i + 3]), atoi(argv[i + 4]), true, atoi(argv[i + 5]) != 0, 0,
                      atoi(argv[i + 6]) != 0, false);
#ifdef SPG_TRACE
            {   // SM-clock timeline of CTA 0 (last launch): producer / MMA issuer / epilogue warp 2 per tile
                static long long tr[64 * 12];
                spg_debug_trace_dump(tr, 64 * 12);
                const long long t0 = tr[0];
                printf("tile | prod start, prod issued | mma start, acc free, first chunk landed, all issued | epi start, bias staged, acc full, acc read | issuer waited   (SM cycles since CTA start)\n");
                for (int t = 0; t < 24; ++t) {
                    printf("%4d |", t);
                    for (int k = 0; k < 10; ++k) printf(" %8lld%s", tr[t * 12 + k] - t0, (k == 1 || k == 5) ? " |" : "");
                    printf(" | %lld\n", tr[t * 12 + 10]);
                }
            }
#endif
            return g_fail;
        }
    }
    if (g_check) {
    // smallest possible: one tile, one k-chunk
    case_gemm(128, 64, 64, SPG_ACT_NONE, false, false, 0, true, false);
    case_gemm(128, 16, 64, SPG_ACT_NONE, false, false, 0, true, false);
    case_gemm(256, 64, 256, SPG_ACT_NONE, true, false, 0, false, false);
    // K tails (144 = 2*64+16, 288 = 4*64+32), N = 144 tiles
    case_gemm(384, 144, 144, SPG_ACT_NONE, true, false, 0, false, false);
    case_gemm(512, 432, 144, SPG_ACT_NONE, true, false, 0, false, false);
    case_gemm(512, 864, 288, SPG_ACT_NONE, true, false, 0, false, false);
    // M tail
    case_gemm(300, 192, 128, SPG_ACT_RELU, true, false, 0, false, false);
    // stage-3 shapes: qkv, proj(+residual fp32), fc1(+GELU), fc2(+residual)
    case_gemm(2048, 1728, 576, SPG_ACT_NONE, true, false, 0, false, false);
    case_gemm(2048, 576, 576, SPG_ACT_NONE, true, true, 0, true, false);
    case_gemm(2048, 2304, 576, SPG_ACT_GELU, true, false, 0, false, false);
    case_gemm(2048, 576, 2304, SPG_ACT_NONE, true, true, 0, true, false);
    // patch-embed shape with broadcast residual (pos-embed): K=160, rows modulo 1024
    case_gemm(4096, 144, 160, SPG_ACT_NONE, true, true, 1024, true, false);
    // fusion-style K=2016 (tail 32) and fused head
    case_gemm(1024, 512, 2016, SPG_ACT_RELU, true, false, 0, false, false);
    case_gemm(1024, 256, 128, SPG_ACT_RELU, true, false, 0, false, true);
    // many tiles per CTA (persistence + both accumulator stages + phase wrap)
    case_gemm(148 * 128 * 3 + 128, 64, 64, SPG_ACT_NONE, false, false, 0, false, false);

    case_conv(2, 16, 16, 64, 64, SPG_ACT_RELU, false, true);
    case_conv(1, 64, 64, 256, 64, SPG_ACT_RELU, true, true);
    case_conv(2, 128, 128, 320, 256, SPG_ACT_RELU, true, true);
    case_conv(1, 256, 256, 128, 128, SPG_ACT_RELU, false, true);
    case_conv(1, 512, 512, 64, 64, SPG_ACT_RELU, true, false);
    }

    if (g_perf) {
        // headline shapes at batch 64
        g_perf = true;
        case_gemm(65536, 1728, 576, SPG_ACT_NONE, true, false, 0, false, false);
        case_gemm(65536, 2304, 576, SPG_ACT_GELU, true, false, 0, false, false);
        case_gemm(65536, 576, 2304, SPG_ACT_NONE, true, true, 0, true, false);
        case_conv(8, 256, 256, 320, 128, SPG_ACT_RELU, false, true);
        case_conv(4, 512, 512, 128, 64, SPG_ACT_RELU, false, true);
        case_gemm(65536, 576, 576, SPG_ACT_NONE, true, true, 0, true, false);
        case_gemm(262144, 432, 144, SPG_ACT_NONE, true, false, 0, false, false);
        case_gemm(262144, 576, 144, SPG_ACT_GELU, true, false, 0, false, false);
        case_conv(16, 128, 128, 320, 256, SPG_ACT_RELU, true, true);
        case_conv(8, 128, 128, 256, 256, SPG_ACT_RELU, true, true);
        case_conv(4, 512, 512, 64, 64, SPG_ACT_RELU, true, false);
        // tile-shape sweep (long K, light epilogue): block_n 256 / 192 / 144 / 128 / 64
        case_gemm(65536, 2304, 2304, SPG_ACT_NONE, false, false, 0, false, false);
        case_gemm(65536, 1728, 2304, SPG_ACT_NONE, false, false, 0, false, false);
        case_gemm(65536, 1152, 2304, SPG_ACT_NONE, false, false, 0, false, false);
        case_gemm(65536, 144 * 8, 2304, SPG_ACT_NONE, false, false, 0, false, false);
        case_gemm(65536, 128, 2304, SPG_ACT_NONE, false, false, 0, false, false);
        case_gemm(65536, 64, 2304, SPG_ACT_NONE, false, false, 0, false, false);
    }
    printf("%d case(s) failed\n", g_fail);
    return g_fail;
}
