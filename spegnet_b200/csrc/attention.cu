// Windowed / global multi-head attention of the Hiera trunk (head_dim = 72), flash-style:
//   out = softmax(q k^T / sqrt(72)) v   per (window, head), optional 2x2 max-pooled queries.
//
// Window partition / un-partition (HF:modeling_sam2.py:378-438) and the q-pool (:317-321) are pure
// addressing here: tokens are gathered straight out of the fused qkv GEMM output [tokens, 3*D] and
// the context is scattered to its final token position, nothing is permuted or copied in HBM.
//
// Round-1 implementation: warp-level mma.sync (m16n8k16, fp32 accumulate) with the window's K / V tiles
// brought into shared memory by TMA -- a window is one box {72 ch, ws, ws, 1} of the qkv tensor viewed as
// [B, H, W, 3D], so the "window partition" costs one descriptor-driven copy and no address arithmetic --
// and the online softmax in registers.  Attention is 4.7 % of the model's FLOPs;
// moving the two GEMMs of it to tcgen05 is listed as next work in DESIGN.md.
#include <atomic>
#include <cstdlib>

#include "common.h"
#include "half16.cuh"
#include "ptx.cuh"

namespace spg {
extern std::atomic<long long> g_launches;
namespace {

constexpr int kHd = 72;          // head dim
constexpr int kRowsSmem = 256;   // keys staged per CTA pass
constexpr int kThreads = 128;    // 4 warps x 16 query rows

struct AttnParams {
    const h16* qkv;  // [B*H*W, 3*D]
    h16* out;        // [B*Ho*Wo, D]
    int B, H, W, D, heads;
    int ws;      // window edge (tokens); global attention = ws == H == W
    int qpool;   // 1: queries are 2x2 max-pooled inside the window
    int nwx, nwy;
    int Nk, Nq;  // keys / queries per window
    int wpc;     // windows per CTA (Nq < 64) else 1
    int qtiles;  // 64-row query tiles per window (Nq >= 64) else 1
    int box_h;   // window rows per K/V TMA box (box = 72 ch x ws x box_h tokens, <= 256 tokens)
    // floor(2^32 / d) + 1 for d = ws and d = the query-grid edge: n / d == __umulhi(n, magic) for n, d < 2^16.  The kernel
    // is short and integer-bound, and a runtime division costs ~25 instructions (there were 21 of them)
    unsigned ws_magic, wq_magic;
    int reverse;    // walk the window groups in descending order (common.h "Traversal direction")
    int rows_smem;  // key rows staged per pass = min(256, keys this CTA sees): sizes the dynamic shared memory
    float scale_log2e;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x1(uint32_t addr, uint32_t& r0) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];" : "=r"(r0) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32." SPG_MMA_TYPE "." SPG_MMA_TYPE ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 2^x on the MUFU alone (exp2f adds a range-scaling sequence around the same instruction; the arguments here are <= 0 and
// results below 2^-126 may flush to zero)
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// n / d through the host-computed reciprocal (magic == 0 encodes d == 1)
__device__ __forceinline__ int fast_div(int n, unsigned magic) {
    return magic ? static_cast<int>(__umulhi(static_cast<unsigned>(n), magic)) : n;
}
// token index (row of the qkv matrix) of local position `li` inside window `win` of image `b`
__device__ __forceinline__ long long window_token(const AttnParams& p, int b, int wy, int wx, int li) {
    const int iy = fast_div(li, p.ws_magic), ix = li - iy * p.ws;
    return (static_cast<long long>(b) * p.H + wy * p.ws + iy) * p.W + wx * p.ws + ix;
}

template <int KB>  // keys consumed per softmax step (16 or 64)
__global__ void __launch_bounds__(kThreads)
window_attention_kernel(const __grid_constant__ CUtensorMap tmap_kv, const AttnParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    // K/V arrive in up to four independent units (64 keys, or one small window) with one mbarrier each, so
    // the first QK^T tiles start while the rest of the window is still in flight
    __shared__ __align__(8) uint64_t kv_bar[4];
    const uint32_t Ks_u = (smem_u32(smem) + 127u) & ~127u;  // TMA destinations: 128-byte aligned
    const uint32_t Vs_u = Ks_u + p.rows_smem * kHd * 2;
    const uint32_t bar_u = smem_u32(&kv_bar[0]);
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_kv);
        for (int u = 0; u < 4; ++u) mbar_init(bar_u + 8u * u, 1);
        fence_barrier_init();
    }
    pdl_launch_dependents();
    __syncthreads();
    pdl_wait();  // everything above touched only shared memory / the tensor map; qkv is read below

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    // heads vary FASTEST over the grid: the CTAs of the heads of one window group run at the same time, so the rows of
    // qkv they share come from HBM once.  (With heads as the slow grid dimension every head pass re-streamed the tensor:
    // a head's 144-byte slices of q, k and v touch nearly every DRAM burst of the 864 ... 3456-byte token row -- ncu:
    // 1.74 GB read for 0.91 GB of qkv at d = 144.)
    const int head = static_cast<int>(blockIdx.x) % p.heads;
    const int wins_per_img = p.nwx * p.nwy;

    // ---- which windows / query rows this CTA owns
    const int nblk = static_cast<int>(gridDim.x) / p.heads, bidx = static_cast<int>(blockIdx.x) / p.heads;
    const int blk = p.reverse ? nblk - 1 - bidx : bidx;
    int win0, qt;
    if (p.wpc > 1) {
        win0 = blk * p.wpc;
        qt = 0;
    } else {
        win0 = blk / p.qtiles;
        qt = blk - win0 * p.qtiles;
    }
    const int my_win = win0 + (p.wpc > 1 ? warp : 0);
    const int b = my_win / wins_per_img;
    const int wrem = my_win - b * wins_per_img;
    const int wy = wrem / p.nwx, wx = wrem - wy * p.nwx;
    const int q0 = (p.wpc > 1 ? 0 : qt * 64 + warp * 16);  // first query row (window-local) of this warp
    const bool warp_active = q0 < p.Nq;

    // ---- Q fragments (registers), with the optional 2x2 max pool over the window's token grid
    uint32_t qa[5][4];
    const size_t ld = static_cast<size_t>(3) * p.D;
    {
        const int wq = p.qpool ? p.ws / 2 : p.ws;  // query grid edge
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int qi = q0 + g + 8 * half;
            const bool ok = warp_active && qi < p.Nq;
            const int qy = fast_div(qi, p.wq_magic), qx = qi - qy * wq;
            const h16* src[4];
            int nsrc = 1;
            if (p.qpool) {
                nsrc = 4;
#pragma unroll
                for (int s = 0; s < 4; ++s)
                    src[s] = p.qkv + window_token(p, b, wy, wx, (2 * qy + (s >> 1)) * p.ws + 2 * qx + (s & 1)) * ld + head * kHd;
            } else {
                src[0] = p.qkv + window_token(p, b, wy, wx, qi) * ld + head * kHd;
            }
#pragma unroll
            for (int kk = 0; kk < 5; ++kk) {
#pragma unroll
                for (int hi = 0; hi < 2; ++hi) {
                    const int d = kk * 16 + hi * 8 + 2 * t;
                    uint32_t v = 0u;
                    if (ok && d < kHd) {
                        v = *reinterpret_cast<const uint32_t*>(src[0] + d);
                        for (int s = 1; s < nsrc; ++s) v = max_h2(v, *reinterpret_cast<const uint32_t*>(src[s] + d));
                    }
                    qa[kk][half + 2 * hi] = v;
                }
            }
        }
    }

    float o[9][4];
#pragma unroll
    for (int i = 0; i < 9; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    const int rows_total = p.wpc > 1 ? p.wpc * p.Nk : p.Nk;  // key rows this CTA must see in total
    for (int kbase = 0; kbase < rows_total; kbase += kRowsSmem) {
        const int rows = min(kRowsSmem, rows_total - kbase);
        if (kbase > 0) __syncthreads();  // everyone is done with the previous K/V pass
        // ---- stage K and V rows with TMA: unit u = one box per K and per V, signalling barrier u
        const uint32_t pass_parity = static_cast<uint32_t>(kbase / kRowsSmem) & 1u;
        if (threadIdx.x == 0) {
            const int unit_rows = p.ws * p.box_h;  // 64 keys, or a whole small window
            const int nunits = rows / unit_rows;
            for (int u = 0; u < nunits; ++u) {
                // quad mode: unit = window win0+u; otherwise unit = box_h consecutive rows of this window
                const int kw = p.wpc > 1 ? win0 + u : win0;
                const int kb = kw / wins_per_img;
                const int kwr = kw - kb * wins_per_img;
                const int kwy = kwr / p.nwx, kwx = kwr - kwy * p.nwx;
                const int y = kwy * p.ws + (p.wpc > 1 ? 0 : (kbase / p.ws + u * p.box_h));
                const uint32_t dst_off = static_cast<uint32_t>(u) * unit_rows * kHd * 2u;
                const uint32_t bar = bar_u + 8u * u;
                mbar_arrive_expect_tx(bar, static_cast<uint32_t>(unit_rows) * kHd * 2u * 2u);
                tma_load_4d(Ks_u + dst_off, &tmap_kv, bar, p.D + head * kHd, kwx * p.ws, y, kb);
                tma_load_4d(Vs_u + dst_off, &tmap_kv, bar, 2 * p.D + head * kHd, kwx * p.ws, y, kb);
            }
        }

        // ---- this warp's slice of the staged rows
        int kbeg = 0, kend = rows;
        if (p.wpc > 1) {
            kbeg = warp * p.Nk;
            kend = kbeg + p.Nk;
        }
        if (warp_active) {
            for (int k0 = kbeg; k0 < kend; k0 += KB) {
                constexpr int NT = KB / 8;
                mbar_wait(bar_u + 8u * (p.wpc > 1 ? warp : (k0 >> 6)), pass_parity);  // this unit has landed
                float s[NT][4];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
                    const int key = k0 + nt * 8 + (lane & 7);
                    const uint32_t row_addr = Ks_u + key * (kHd * 2);
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4(row_addr + (lane >> 3) * 16, b0, b1, b2, b3);  // dims 0..31
                    mma_bf16(s[nt], qa[0][0], qa[0][1], qa[0][2], qa[0][3], b0, b1);
                    mma_bf16(s[nt], qa[1][0], qa[1][1], qa[1][2], qa[1][3], b2, b3);
                    ldsm_x4(row_addr + 64 + (lane >> 3) * 16, b0, b1, b2, b3);  // dims 32..63
                    mma_bf16(s[nt], qa[2][0], qa[2][1], qa[2][2], qa[2][3], b0, b1);
                    mma_bf16(s[nt], qa[3][0], qa[3][1], qa[3][2], qa[3][3], b2, b3);
                    ldsm_x1(row_addr + 128, b0);  // dims 64..71; dims 72..79 are zero padding
                    mma_bf16(s[nt], qa[4][0], qa[4][1], qa[4][2], qa[4][3], b0, 0u);
                }
                // ---- online softmax (rows g and g+8)
                float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
                    mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
                }
                float corr[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
                    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
                    const float m_new = fmaxf(m_run[h], mx[h] * p.scale_log2e);
                    corr[h] = fast_exp2(m_run[h] - m_new);  // first step: 2^(-inf) = 0, and o / l are still zero
                    m_run[h] = m_new;
                }
                float rs[2] = {0.f, 0.f};
                uint32_t pa[NT][2];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float p0 = fast_exp2(fmaf(s[nt][0], p.scale_log2e, -m_run[0]));
                    const float p1 = fast_exp2(fmaf(s[nt][1], p.scale_log2e, -m_run[0]));
                    const float p2 = fast_exp2(fmaf(s[nt][2], p.scale_log2e, -m_run[1]));
                    const float p3 = fast_exp2(fmaf(s[nt][3], p.scale_log2e, -m_run[1]));
                    rs[0] += p0 + p1;
                    rs[1] += p2 + p3;
                    pa[nt][0] = pack2(p0, p1);
                    pa[nt][1] = pack2(p2, p3);
                }
                if (k0 == kbeg && kbase == 0) {  // nothing accumulated yet (the only step of a <= 64-key window)
                    l_run[0] = rs[0];
                    l_run[1] = rs[1];
                } else {
                    l_run[0] = l_run[0] * corr[0] + rs[0];
                    l_run[1] = l_run[1] * corr[1] + rs[1];
#pragma unroll
                    for (int i = 0; i < 9; ++i) {
                        o[i][0] *= corr[0]; o[i][1] *= corr[0];
                        o[i][2] *= corr[1]; o[i][3] *= corr[1];
                    }
                }
                // ---- O += P V   (k = keys in steps of 16, n = 72 dims in 9 tiles of 8)
#pragma unroll
                for (int j = 0; j < KB / 16; ++j) {
                    const uint32_t a0 = pa[2 * j][0], a1 = pa[2 * j][1], a2 = pa[2 * j + 1][0], a3 = pa[2 * j + 1][1];
                    const int key = k0 + 16 * j + ((lane >> 3) & 1) * 8 + (lane & 7);
                    const uint32_t row_addr = Vs_u + key * (kHd * 2) + (lane >> 4) * 16;
#pragma unroll
                    for (int dt = 0; dt < 4; ++dt) {
                        uint32_t b0, b1, b2, b3;
                        ldsm_x4_t(row_addr + dt * 32, b0, b1, b2, b3);
                        mma_bf16(o[2 * dt], a0, a1, a2, a3, b0, b1);
                        mma_bf16(o[2 * dt + 1], a0, a1, a2, a3, b2, b3);
                    }
                    uint32_t b0, b1;
                    ldsm_x2_t(Vs_u + key * (kHd * 2) + 128, b0, b1);
                    mma_bf16(o[8], a0, a1, a2, a3, b0, b1);
                }
            }
        }
    }

    if (!warp_active) return;
    // ---- normalise and scatter to the token's final position
    const int wq = p.qpool ? p.ws / 2 : p.ws;
    const int Wo = p.qpool ? p.W / 2 : p.W, Ho = p.qpool ? p.H / 2 : p.H;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float l = l_run[half];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        const int qi = q0 + g + 8 * half;
        if (qi >= p.Nq) continue;
        const float inv = 1.f / l;
        const int qy = fast_div(qi, p.wq_magic), qx = qi - qy * wq;
        const long long tok = (static_cast<long long>(b) * Ho + wy * wq + qy) * Wo + wx * wq + qx;
        h16* dst = p.out + tok * p.D + head * kHd + 2 * t;
#pragma unroll
        for (int i = 0; i < 9; ++i)
            *reinterpret_cast<uint32_t*>(dst + 8 * i) = pack2(o[i][2 * half] * inv, o[i][2 * half + 1] * inv);
    }
}

// Tiny-window fallback on CUDA cores (block 8 of Hiera-L: 4 pooled queries x 16 keys per window):
// one warp per (window, head, query), lanes over keys for the scores and over dims for the output.
__global__ void __launch_bounds__(128) tiny_attention_kernel(const AttnParams p) {
    pdl_prologue();
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int total = p.B * p.nwx * p.nwy * p.heads * p.Nq;
    if (warp_global >= total) return;
    const int qi = warp_global % p.Nq;
    const int head = (warp_global / p.Nq) % p.heads;
    const int win = warp_global / (p.Nq * p.heads);
    const int wins_per_img = p.nwx * p.nwy;
    const int b = win / wins_per_img, wrem = win % wins_per_img;
    const int wy = wrem / p.nwx, wx = wrem % p.nwx;
    const size_t ld = static_cast<size_t>(3) * p.D;
    const int wq = p.qpool ? p.ws / 2 : p.ws;
    const int qy = qi / wq, qx = qi % wq;
    // q (pooled) : lanes hold dims lane, lane+32, lane+64
    float q[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int d = lane + 32 * i;
        float v = 0.f;
        if (d < kHd) {
            if (p.qpool) {
                v = -INFINITY;
                for (int s = 0; s < 4; ++s)
                    v = fmaxf(v, h_to_float(p.qkv[window_token(p, b, wy, wx, (2 * qy + (s >> 1)) * p.ws + 2 * qx + (s & 1)) * ld + head * kHd + d]));
            } else {
                v = h_to_float(p.qkv[window_token(p, b, wy, wx, qi) * ld + head * kHd + d]);
            }
        }
        q[i] = v;
    }
    float m = -INFINITY, l = 0.f, acc[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < p.Nk; ++k) {
        const h16* kr = p.qkv + window_token(p, b, wy, wx, k) * ld + p.D + head * kHd;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int d = lane + 32 * i;
            if (d < kHd) s += q[i] * h_to_float(kr[d]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s *= p.scale_log2e;
        const float m_new = fmaxf(m, s);
        const float corr = exp2f(m - m_new), pr = exp2f(s - m_new);
        // probabilities are rounded to bf16 before P.V, like the tensor-core path
        const float prb = h_to_float(float_to_h(pr));
        l = l * corr + pr;
        const h16* vr = kr + p.D;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int d = lane + 32 * i;
            acc[i] = acc[i] * corr + (d < kHd ? prb * h_to_float(vr[d]) : 0.f);
        }
        m = m_new;
    }
    const int Wo = p.qpool ? p.W / 2 : p.W, Ho = p.qpool ? p.H / 2 : p.H;
    const long long tok = (static_cast<long long>(b) * Ho + wy * wq + qy) * Wo + wx * wq + qx;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int d = lane + 32 * i;
        if (d < kHd) p.out[tok * p.D + head * kHd + d] = float_to_h(acc[i] / l);
    }
}

}  // namespace
}  // namespace spg

extern "C" int spg_window_attention_tc_h16(const void* qkv, void* out, int B, int H, int W, int D, int heads,
                                           int window, int q_pool, const spg_launch_t* launch);

extern "C" int spg_window_attention_h16(const void* qkv, void* out, int B, int H, int W, int D, int heads,
                                         int window, int q_pool, const spg_launch_t* launch) {
    using namespace spg;
    SPG_CHECK_ARG(qkv && out, "null pointer");
    {
        // 16x16 windows without query pooling (32 of Hiera-L's 48 blocks) run on tcgen05 / TMEM (attention_tc.cu)
        static const int tc_env = [] { const char* e = getenv("SPG_ATTN_TC"); return e ? atoi(e) : 1; }();
        if (tc_env && window == 16 && !q_pool && H % 16 == 0 && W % 16 == 0 && D == heads * 72)
            return spg_window_attention_tc_h16(qkv, out, B, H, W, D, heads, window, q_pool, launch);
        // global blocks (window == 0) on a 32 / 64 / 128-wide token grid: two-pass tcgen05 kernel (SPG_ATTN_TC=2: windows only)
        // (at every batch size: an image's result must not depend on the batch it is in, bit for bit)
        if (tc_env == 1 && window == 0 && !q_pool && H == W && (W == 32 || W == 64 || W == 128) && D == heads * 72)
            return spg_window_attention_tc_h16(qkv, out, B, H, W, D, heads, window, q_pool, launch);
    }
    SPG_CHECK_ARG(heads > 0 && D == heads * kHd, "attention is specialised for head_dim 72 (D=%d heads=%d)", D, heads);
    int ws = window;
    if (ws == 0) {
        SPG_CHECK_ARG(H == W, "global attention needs a square token grid");
        ws = H;
    }
    SPG_CHECK_ARG(H % ws == 0 && W % ws == 0, "token grid %dx%d does not tile into %dx%d windows", H, W, ws, ws);
    SPG_CHECK_ARG(!q_pool || ws % 2 == 0, "q_pool needs an even window");
    AttnParams p{};
    p.qkv = static_cast<const h16*>(qkv);
    p.out = static_cast<h16*>(out);
    p.B = B; p.H = H; p.W = W; p.D = D; p.heads = heads; p.ws = ws; p.qpool = q_pool ? 1 : 0;
    SPG_CHECK_ARG(ws >= 1 && ws < 65536, "window edge out of range");
    p.ws_magic = ws == 1 ? 0u : static_cast<unsigned>((1ull << 32) / static_cast<unsigned>(ws)) + 1u;
    {
        const unsigned wq = static_cast<unsigned>(q_pool ? ws / 2 : ws);
        p.wq_magic = wq <= 1 ? 0u : static_cast<unsigned>((1ull << 32) / wq) + 1u;
    }
    p.nwx = W / ws; p.nwy = H / ws;
    p.Nk = ws * ws;
    p.Nq = q_pool ? p.Nk / 4 : p.Nk;
    p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(kHd));
    const LaunchCtx st(launch);
    p.reverse = st.reverse ? 1 : 0;
    const int nwin = B * p.nwx * p.nwy;
    // tensor-core path: 64-row query tiles over <=256-key passes, or four 16-query windows per CTA
    const bool big = p.Nq % 64 == 0 && p.Nk % 64 == 0 && (p.Nk <= kRowsSmem || p.Nk % kRowsSmem == 0);
    const bool quad = (p.Nq == 16 || p.Nq == 4) && (p.Nk == 16 || p.Nk == 64) && nwin % 4 == 0;
    const bool mma_ok = big || quad;
    if (!mma_ok) {
        const long long warps = static_cast<long long>(nwin) * heads * p.Nq;
        SPG_CHECK_CUDA((launch_pdl(tiny_attention_kernel, static_cast<unsigned>((warps * 32 + 127) / 128), 128, 0, st, p)));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        SPG_CHECK_LAUNCH();
        return SPG_OK;
    }
    p.wpc = p.Nq < 64 ? 4 : 1;  // one window per warp; rows >= Nq of a warp's 16-row tile are masked off
    p.qtiles = p.Nq >= 64 ? p.Nq / 64 : 1;
    p.box_h = p.Nk <= 64 ? ws : 64 / ws;  // a unit is 64 keys (or the whole window when it is smaller)
    SPG_CHECK_ARG(ws <= 64 && p.box_h >= 1, "window %d too wide for the K/V staging", ws);
    CUtensorMap tmap;
    if (int rc = make_tmap_qkv_window(&tmap, qkv, B, H, W, 3 * D, kHd, ws, p.box_h)) return rc;
    // stage only what a CTA consumes per pass (64 keys = 18 KB of K + V for the small windows).  Occupancy is then set
    // by registers (135 -> 3 CTAs / SM); capping them at 128 for a 4th CTA spills and measured slower (global blocks
    // 600 -> 710 us), so the kernel keeps its natural register count
    p.rows_smem = (p.wpc > 1 ? p.wpc * p.Nk : p.Nk) < kRowsSmem ? (p.wpc > 1 ? p.wpc * p.Nk : p.Nk) : kRowsSmem;
    const int smem = 2 * p.rows_smem * kHd * 2 + 128;
    static PerDeviceOnce attr_set;
    if (attr_set.needed()) {
        const int smem_max = 2 * kRowsSmem * kHd * 2 + 128;
        SPG_CHECK_CUDA(cudaFuncSetAttribute(window_attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        SPG_CHECK_CUDA(cudaFuncSetAttribute(window_attention_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        attr_set.done();
    }
    const long long nblocks = static_cast<long long>(p.wpc > 1 ? nwin / p.wpc : nwin * p.qtiles) * heads;
    SPG_CHECK_ARG(nblocks < (1ll << 31), "attention grid too large");
    dim3 grid(static_cast<unsigned>(nblocks));  // block = (window group, head), head fastest
    if (p.Nk == 16)
        SPG_CHECK_CUDA((launch_pdl(window_attention_kernel<16>, grid, kThreads, smem, st, tmap, p)));
    else
        SPG_CHECK_CUDA((launch_pdl(window_attention_kernel<64>, grid, kThreads, smem, st, tmap, p)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    SPG_CHECK_LAUNCH();
    return SPG_OK;
}
