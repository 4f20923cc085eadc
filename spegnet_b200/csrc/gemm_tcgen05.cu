// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] * W[N,K]^T )        bf16 operands, fp32 accumulation in TMEM
//
// One CTA per SM loops over 128 x block_n output tiles (n fastest, so concurrently running CTAs
// share A tiles through L2).  Roles (320 threads):
//   warp 0      TMA producer   - fills a ring of {A 128x64, W block_n x 64} bf16 stages (128B swizzle)
//   warp 1      MMA issuer     - one thread issues tcgen05.mma (M=128, N=block_n, K=16) x4 per stage,
//                                tcgen05.commit releases the stage / publishes the accumulator
//   warps 2..9  epilogue       - tcgen05.ld the accumulator (thread == output row, two warps per TMEM
//                                lane quarter split the columns), bias / GELU / ReLU / fp32 residual /
//                                fused N->1 head, vectorised stores
// The accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps
// the MMAs of tile i+1.
//
// Convolution mode: the A tile of k-chunk (tap, channel-chunk) is a 4-D TMA box
// {64 ch, tile_w, tile_h, 1 image} of the NHWC input at (x0+dx-1, y0+dy-1); the zero halo comes
// from TMA out-of-bounds fill, so the same MMA / epilogue pipeline serves both modes.
#include <atomic>

#include "common.h"
#include "half16.cuh"
#include "ptx.cuh"

namespace spg {

extern std::atomic<long long> g_launches;

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kAStageBytes = kBlockM * kBlockK * 2;
constexpr int kThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kTmemCols = 512;
constexpr int kAccStageCols = 256;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kEpiScratch = (2 * 256 + 256 + 128) * 4;  // bias[2][256], head weights[256], head partials[128]

struct GemmArgs {
    int M, N;
    int block_n;
    int num_m_tiles, num_n_tiles, num_k_chunks;
    int stages;
    // conv mode
    int conv;
    int H, W, cin_chunks, tile_w;
    // epilogue
    const float* bias;
    int act;
    const float* residual;
    int res_rows;
    void* out;
    int out_f32;
    const float* head_w;
    float head_b;
    float* head_out;
};

__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) { return pack2(lo, hi); }

__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b, const GemmArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
    const uint32_t b_stage_bytes = static_cast<uint32_t>(p.block_n) * 128u;
    const uint32_t stage_bytes = kAStageBytes + b_stage_bytes;
    const uint32_t bar_addr = tiles_addr + static_cast<uint32_t>(p.stages) * stage_bytes;
    auto full_bar = [&](int s) { return bar_addr + 8u * s; };
    auto empty_bar = [&](int s) { return bar_addr + 8u * (p.stages + s); };
    auto tmem_full_bar = [&](int a) { return bar_addr + 8u * (2 * p.stages + a); };
    auto tmem_empty_bar = [&](int a) { return bar_addr + 8u * (2 * p.stages + 2 + a); };
    const uint32_t tmem_slot = bar_addr + 8u * (2 * p.stages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.num_m_tiles * p.num_n_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), 8);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const int hw = p.H * p.W;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m_blk = tile / p.num_n_tiles;
                const int n_blk = tile - m_blk * p.num_n_tiles;
                int img = 0, y0 = 0, x0 = 0;
                if (p.conv) {
                    const int m0 = m_blk * kBlockM;
                    img = m0 / hw;
                    const int rem = m0 - img * hw;
                    y0 = rem / p.W;
                    x0 = rem - y0 * p.W;
                }
                for (int kc = 0; kc < p.num_k_chunks; ++kc) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t a_dst = tiles_addr + stage * stage_bytes;
                    const uint32_t b_dst = a_dst + kAStageBytes;
                    mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
                    if (p.conv) {
                        const int tap = kc / p.cin_chunks;
                        const int cc = kc - tap * p.cin_chunks;
                        const int dy = tap / 3 - 1;
                        const int dx = tap - (tap / 3) * 3 - 1;
                        tma_load_4d(a_dst, &tmap_a, full_bar(stage), cc * kBlockK, x0 + dx, y0 + dy, img);
                    } else {
                        tma_load_2d(a_dst, &tmap_a, full_bar(stage), kc * kBlockK, m_blk * kBlockM);
                    }
                    tma_load_2d(b_dst, &tmap_b, full_bar(stage), kc * kBlockK, n_blk * p.block_n);
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16_f32(kBlockM, p.block_n);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kAccStageCols;
                for (int kc = 0; kc < p.num_k_chunks; ++kc) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_addr = tiles_addr + stage * stage_bytes;
                    const uint64_t a_desc = make_sw128_kmajor_desc(a_addr);
                    const uint64_t b_desc = make_sw128_kmajor_desc(a_addr + kAStageBytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        // advance 32 B (= 16 bf16) along K inside the 128 B swizzle atom
                        const uint64_t koff = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
                        umma_bf16_ss(d_tmem, a_desc + koff, b_desc + koff, idesc, (kc | k) != 0);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(tmem_full_bar(acc));
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        // Two warps per TMEM lane quarter: both own the same 32 output rows, each takes half of the
        // tile's 16-column chunks.  Per tile: bias (and head weights) are staged in smem BEFORE the
        // accumulator is awaited, TMEM loads are double-buffered against the arithmetic, and the fp32
        // residual of the next chunk is prefetched, so no global / TMEM latency sits on the critical path.
        const int ewarp = warp - 2;                // 0..7
        const int quarter = warp & 3;              // TMEM lane quarter this warp may access
        const int half = ewarp >> 2;               // which half of the chunks
        const int etid = threadIdx.x - 64;         // 0..255
        const int row_in_tile = quarter * 32 + lane;
        const int chunks = p.block_n >> 4;
        const int c_begin = half == 0 ? 0 : (chunks + 1) >> 1;
        const int c_end = half == 0 ? (chunks + 1) >> 1 : chunks;
        float* bias_s = reinterpret_cast<float*>(smem_raw + (tiles_addr - raw_addr) + p.stages * stage_bytes + 256);
        float* headw_s = bias_s + 2 * 256;         // [2][256] bias, [256] head weights, [128] head partials
        float* headp_s = headw_s + 256;
        if (p.head_w != nullptr && etid < p.block_n) headw_s[etid] = __ldg(p.head_w + etid);  // num_n_tiles == 1
        int acc = 0;
        uint32_t acc_phase = 0;
        int buf = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m_blk = tile / p.num_n_tiles;
            const int n_blk = tile - m_blk * p.num_n_tiles;
            const int n0 = n_blk * p.block_n;
            float* bs = bias_s + buf * 256;
            if (etid < p.block_n) bs[etid] = p.bias != nullptr ? __ldg(p.bias + n0 + etid) : 0.f;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int row = m_blk * kBlockM + row_in_tile;
            const bool row_ok = row < p.M;
            const size_t out_off = static_cast<size_t>(row) * p.N + n0;
            const float* res_row = nullptr;
            if (p.residual != nullptr && row_ok) {
                const int rr = p.res_rows > 0 ? (row % p.res_rows) : row;
                res_row = p.residual + static_cast<size_t>(rr) * p.N + n0;
            }
            float4 rnext[4];
            if (res_row != nullptr) {
                const float4* r4 = reinterpret_cast<const float4*>(res_row + c_begin * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) rnext[i] = r4[i];
            }
            mbar_wait(tmem_full_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kAccStageCols;
            float head_acc = 0.f;

            auto process = [&](const uint32_t (&raw)[16], int c) {
                const int col = c * 16;  // column inside the tile
                float v[16];
                const float4* b4 = reinterpret_cast<const float4*>(bs + col);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b = b4[i];
                    v[4 * i + 0] = __uint_as_float(raw[4 * i + 0]) + b.x;
                    v[4 * i + 1] = __uint_as_float(raw[4 * i + 1]) + b.y;
                    v[4 * i + 2] = __uint_as_float(raw[4 * i + 2]) + b.z;
                    v[4 * i + 3] = __uint_as_float(raw[4 * i + 3]) + b.w;
                }
                if (p.act == SPG_ACT_RELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                } else if (p.act == SPG_ACT_GELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
                }
                if (res_row != nullptr) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[4 * i + 0] += rnext[i].x;
                        v[4 * i + 1] += rnext[i].y;
                        v[4 * i + 2] += rnext[i].z;
                        v[4 * i + 3] += rnext[i].w;
                    }
                    if (c + 1 < c_end) {  // prefetch the next chunk's residual
                        const float4* r4 = reinterpret_cast<const float4*>(res_row + col + 16);
#pragma unroll
                        for (int i = 0; i < 4; ++i) rnext[i] = r4[i];
                    }
                }
                if (p.head_w != nullptr) {
                    const float4* w4 = reinterpret_cast<const float4*>(headw_s + col);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 w = w4[i];
                        head_acc = fmaf(v[4 * i + 0], w.x, head_acc);
                        head_acc = fmaf(v[4 * i + 1], w.y, head_acc);
                        head_acc = fmaf(v[4 * i + 2], w.z, head_acc);
                        head_acc = fmaf(v[4 * i + 3], w.w, head_acc);
                    }
                }
                if (p.out != nullptr && row_ok) {
                    if (p.out_f32) {
                        float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(p.out) + out_off + col);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    } else {
                        uint4* o4 = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.out) + out_off + col);
                        o4[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                           pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                        o4[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]),
                                           pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
                    }
                }
            };

            uint32_t ra[16], rb[16];
            if (c_begin < c_end) tmem_ld16(taddr + c_begin * 16, ra);
            for (int c = c_begin; c < c_end; c += 2) {
                tmem_ld_wait();
                if (c + 1 < c_end) tmem_ld16(taddr + (c + 1) * 16, rb);
                process(ra, c);
                if (c + 1 < c_end) {
                    tmem_ld_wait();
                    if (c + 2 < c_end) tmem_ld16(taddr + (c + 2) * 16, ra);
                    process(rb, c + 1);
                }
            }
            // accumulator fully read: hand the TMEM stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
            if (p.head_w != nullptr) {
                // combine the two column halves of the fused N->1 head through smem
                if (half == 1) headp_s[row_in_tile] = head_acc;
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (half == 0 && row_ok) p.head_out[row] = head_acc + headp_s[row_in_tile] + p.head_b;
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
            buf ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

int pick_block_n(int N) {
    for (int bn = 256; bn >= 16; bn -= 16)
        if (N % bn == 0) return bn;
    return 0;
}

int launch(const CUtensorMap& ta, const CUtensorMap& tb, GemmArgs& a, cudaStream_t stream) {
    const int stage_bytes = kAStageBytes + a.block_n * 128;
    int stages = (kSmemBudget - 1024 - 256 - kEpiScratch) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages > a.num_k_chunks + 1) stages = a.num_k_chunks + 1;  // no point in a deeper ring
    if (stages < 2) stages = 2;
    a.stages = stages;
    const int smem = stages * stage_bytes + 1024 + 256 + kEpiScratch;
    static bool attr_set = false;
    if (!attr_set) {
        SPG_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
        attr_set = true;
    }
    const int total = a.num_m_tiles * a.num_n_tiles;
    const int sms = sm_count();
    const int grid = total < sms ? total : sms;
    gemm_tcgen05_kernel<<<grid, kThreads, smem, stream>>>(ta, tb, a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    SPG_CHECK_LAUNCH();
    return SPG_OK;
}

int fill_epilogue(GemmArgs& a, const spg_epilogue_t* ep, int N) {
    SPG_CHECK_ARG(ep != nullptr, "epilogue descriptor is NULL");
    SPG_CHECK_ARG(ep->out != nullptr || ep->head_out != nullptr, "epilogue has neither out nor head_out");
    SPG_CHECK_ARG(ep->act >= SPG_ACT_NONE && ep->act <= SPG_ACT_GELU, "unknown activation %d", ep->act);
    SPG_CHECK_ARG(ep->out_dtype == SPG_H16 || ep->out_dtype == SPG_F32, "unknown out_dtype %d", ep->out_dtype);
    SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(ep->out) & 15) == 0, "out must be 16-byte aligned");
    SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(ep->bias) & 15) == 0, "bias must be 16-byte aligned");
    SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(ep->residual) & 15) == 0, "residual must be 16-byte aligned");
    SPG_CHECK_ARG(ep->res_rows >= 0, "res_rows must be >= 0");
    if (ep->head_w != nullptr) {
        SPG_CHECK_ARG(ep->head_out != nullptr, "head_w given without head_out");
        SPG_CHECK_ARG(a.block_n == N, "fused head needs the whole row in one tile (N=%d <= 256, N %% 16 == 0)", N);
        SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(ep->head_w) & 15) == 0, "head_w must be 16-byte aligned");
    }
    a.bias = ep->bias;
    a.act = ep->act;
    a.residual = ep->residual;
    a.res_rows = ep->res_rows;
    a.out = ep->out;
    a.out_f32 = ep->out_dtype == SPG_F32;
    a.head_w = ep->head_w;
    a.head_b = ep->head_b;
    a.head_out = ep->head_w != nullptr ? ep->head_out : nullptr;
    return SPG_OK;
}

}  // namespace
}  // namespace spg

extern "C" int spg_linear_h16(const void* A, const void* W, int M, int N, int K,
                               const spg_epilogue_t* ep, spg_stream_t stream) {
    using namespace spg;
    SPG_CHECK_ARG(A != nullptr && W != nullptr, "A / W is NULL");
    SPG_CHECK_ARG(M > 0 && N > 0 && K > 0, "bad GEMM shape M=%d N=%d K=%d", M, N, K);
    SPG_CHECK_ARG(K % 8 == 0, "K=%d must be a multiple of 8 (16-byte TMA row pitch)", K);
    SPG_CHECK_ARG(N % 16 == 0, "N=%d must be a multiple of 16 (UMMA N granularity at M=128)", N);
    GemmArgs a{};
    a.M = M;
    a.N = N;
    a.block_n = pick_block_n(N);
    a.num_m_tiles = (M + kBlockM - 1) / kBlockM;
    a.num_n_tiles = N / a.block_n;
    a.num_k_chunks = (K + kBlockK - 1) / kBlockK;
    a.conv = 0;
    a.H = a.W = 1;
    a.cin_chunks = 1;
    a.tile_w = 1;
    if (int rc = fill_epilogue(a, ep, N)) return rc;
    CUtensorMap ta, tb;
    if (int rc = make_tmap_2d(&ta, A, M, K, static_cast<uint64_t>(K) * 2, kBlockM)) return rc;
    if (int rc = make_tmap_2d(&tb, W, N, K, static_cast<uint64_t>(K) * 2, a.block_n)) return rc;
    return launch(ta, tb, a, static_cast<cudaStream_t>(stream));
}

extern "C" int spg_conv3x3_h16(const void* x, const void* w, int B, int H, int W, int Cin, int Cout,
                                const spg_epilogue_t* ep, spg_stream_t stream) {
    using namespace spg;
    SPG_CHECK_ARG(x != nullptr && w != nullptr, "x / w is NULL");
    SPG_CHECK_ARG(B > 0 && H > 0 && W > 0, "bad conv shape B=%d H=%d W=%d", B, H, W);
    SPG_CHECK_ARG(Cin % kBlockK == 0, "Cin=%d must be a multiple of 64", Cin);
    SPG_CHECK_ARG(Cout % 16 == 0, "Cout=%d must be a multiple of 16", Cout);
    const int tile_w = W < kBlockM ? W : kBlockM;
    SPG_CHECK_ARG(kBlockM % tile_w == 0 && W % tile_w == 0, "W=%d must divide or be a multiple of 128", W);
    const int tile_h = kBlockM / tile_w;
    SPG_CHECK_ARG(H % tile_h == 0, "H=%d must be a multiple of the tile height %d", H, tile_h);
    GemmArgs a{};
    a.M = B * H * W;
    a.N = Cout;
    a.block_n = pick_block_n(Cout);
    a.num_m_tiles = a.M / kBlockM;
    a.num_n_tiles = Cout / a.block_n;
    a.cin_chunks = Cin / kBlockK;
    a.num_k_chunks = 9 * a.cin_chunks;
    a.conv = 1;
    a.H = H;
    a.W = W;
    a.tile_w = tile_w;
    if (int rc = fill_epilogue(a, ep, Cout)) return rc;
    CUtensorMap ta, tb;
    if (int rc = make_tmap_nhwc(&ta, x, B, H, W, Cin, tile_h, tile_w)) return rc;
    if (int rc = make_tmap_2d(&tb, w, Cout, 9ull * Cin, 18ull * Cin, a.block_n)) return rc;
    return launch(ta, tb, a, static_cast<cudaStream_t>(stream));
}
