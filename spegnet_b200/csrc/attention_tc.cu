// tcgen05 / TMEM window attention for the stage-3 windowed blocks of Hiera-L (16x16 windows = 256 keys,
// 256 queries, head_dim 72):  out = softmax(q k^T / sqrt(72)) v  per (image, window, head).
//
// One persistent CTA per SM loops over (window, head) work items; a work item is two 128-query tiles.
//   warp 0      TMA producer: 5-D boxes {d, 1 head, 1 of q|k|v, 16 x, 8 y} of the qkv tensor viewed as
//               [72, heads, 3, W, B*H].  The 72-wide head dim is split into a 64-element SWIZZLE_128B box and a
//               16-element SWIZZLE_32B tail box whose elements 72..79 are out of bounds in dim 0 and therefore
//               ZERO-FILLED by the TMA unit (free K padding to 80 = 5 UMMA k-steps).  Q/K are double-buffered
//               across work items, V single-buffered (it is consumed last).
//   warp 1      MMA issuer: S = Q K^T as 4 + 1 SS-MMAs (M=128, N=256) into TMEM; then O = P V as TS-MMAs with the
//               fp16 probabilities read from TMEM as the A operand and V as an MN-major B operand (N = 64 dims main
//               + N = 16 tail), 16 k-steps of 16 keys.
//   warps 2..9  two softmax / epilogue warpgroups, one per query tile, thread == query row: two passes over the S row
//               in TMEM (max, then exp2 / sum) with double-buffered tcgen05.ld, P written back to TMEM in place as
//               packed 16-bit pairs (tcgen05.st), O read back, scaled by 1/sum and stored.  No shuffles: every
//               reduction is along a thread's own row.
// TMEM: one 256-column region per query tile.  S occupies [0,256); P overwrites the consumed S columns [0,128);
// O accumulates in [128,208).  The MMA warp issues S0, S1, PV0, PV1 per work item, so the tensor pipe works on one
// tile while the other tile's warpgroup is in its softmax / epilogue.
#include <atomic>

#include "common.h"
#include "half16.cuh"
#include "ptx.cuh"

namespace spg {
extern std::atomic<long long> g_launches;
namespace {

constexpr int kHd = 72;
constexpr int kWs = 16;                 // window edge
constexpr int kKeys = kWs * kWs;        // 256
constexpr int kThreadsTc = 320;  // global kernel: warp 0 TMA, warp 1 MMA, warps 2..5 / 6..9 softmax of query tile 0 / 1
constexpr int kThreadsWin = 352; // windowed kernel: + warp 10 = the MMA issuer of query tile 1 (warp 1 serves tile 0)
constexpr uint32_t kQMain = 128 * 128;  // one 128-query tile, dims 0..63   (SWIZZLE_128B rows of 128 B)
constexpr uint32_t kQTail = 128 * 32;   // dims 64..79                      (SWIZZLE_32B rows of 32 B)
constexpr uint32_t kKMain = kKeys * 128, kKTail = kKeys * 32;
// shared memory map (bytes, all 1 KB aligned)
constexpr uint32_t kOffK = 0;                         // per stage: K main | K tail | Q main x2 | Q tail x2
constexpr uint32_t kOffKt = kOffK + kKMain;
constexpr uint32_t kOffQ = kOffKt + kKTail;
constexpr uint32_t kOffQt = kOffQ + 2 * kQMain;
constexpr uint32_t kStageQK = kOffQt + 2 * kQTail;    // 81920
constexpr uint32_t kOffV = 2 * kStageQK;              // V main | V tail (single buffer)
constexpr uint32_t kOffVt = kOffV + kKMain;
constexpr uint32_t kOffBar = kOffVt + kKTail;         // 204800
constexpr uint32_t kSmemTc = kOffBar + 256 + 1024;    // + barriers + alignment slack

// Softmax arithmetic on Blackwell's packed / 3-input fp32 forms (same roundings as the scalar code they replace, so the
// results are bit-identical): FMNMX3 halves the instructions and the dependent chain of the row maximum, FFMA2 / FADD2
// process two probabilities per issue slot -- the softmax warps are the busy resource of these kernels.
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// 32 raw fp32 scores -> running maxima (two independent chains)
__device__ __forceinline__ void row_max32(const uint32_t (&raw)[32], float& m0, float& m1) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        m0 = max3(m0, __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
        m1 = max3(m1, __uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3]));
    }
}
// {p0, p1} = 2^({s0, s1} * scale - m); sum2 += {p0, p1} (packed); returns the 16-bit pair
__device__ __forceinline__ uint32_t exp2_pair(uint32_t s0, uint32_t s1, unsigned long long scale2, unsigned long long negm2,
                                              unsigned long long& sum2) {
    unsigned long long x, t;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(s0), "r"(s1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(x), "l"(scale2), "l"(negm2));
    float t0, t1, p0, p1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(t1));
    unsigned long long pp;
    asm("mov.b64 %0, {%1, %2};" : "=l"(pp) : "f"(p0), "f"(p1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(sum2) : "l"(sum2), "l"(pp));
    return pack2(p0, p1);
}
__device__ __forceinline__ unsigned long long splat2(float v) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float hsum2(unsigned long long v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a + b;
}

struct AttnTcParams {
    h16* out;  // [B*H*W, D]
    int B, H, W, D, heads;
    int nwx, nwy;
    int items;  // B * nwy * nwx * heads
    int reverse;  // walk the items in descending order (common.h "Traversal direction")
    float scale_log2e;
};

__global__ void __launch_bounds__(kThreadsWin, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_main, const __grid_constant__ CUtensorMap tmap_tail,
                    const AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + kOffBar;
    // barriers: qk_full[2], qk_empty[2], v_full, v_empty, and per TMEM region r: s_full[r], p_full[r], o_full[r],
    // region_free[r]
    const uint32_t qk_full = bar, qk_empty = bar + 16, v_full = bar + 32, v_empty = bar + 40, s_full = bar + 48,
                   p_full = bar + 64, o_full = bar + 80, region_free = bar + 96, tmem_slot = bar + 112;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_main);
        tma_prefetch_desc(&tmap_tail);
        for (int s = 0; s < 2; ++s) {
            mbar_init(qk_full + 8 * s, 1);
            mbar_init(qk_empty + 8 * s, 2);  // one commit per MMA issuer (query tile)
        }
        mbar_init(v_full, 1);
        mbar_init(v_empty, 2);
        for (int r = 0; r < 2; ++r) {
            mbar_init(s_full + 8 * r, 1);
            mbar_init(p_full + 8 * r, 4);       // one arrive per softmax warp of the region's warpgroup
            mbar_init(o_full + 8 * r, 1);
            mbar_init(region_free + 8 * r, 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();  // barriers / TMEM are set up; qkv (the previous kernel's output) is first read below
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int wins_per_img = p.nwx * p.nwy;
    auto decode = [&](int item, int& head, int& y0, int& x0) {
        if (p.reverse) item = p.items - 1 - item;
        head = item % p.heads;
        const int win = item / p.heads;
        const int b = win / wins_per_img;
        const int wr = win - b * wins_per_img;
        const int wy = wr / p.nwx, wx = wr - wy * p.nwx;
        y0 = b * p.H + wy * kWs;  // row of the [B*H, W] token grid
        x0 = wx * kWs;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int n = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
                int head, y0, x0;
                decode(item, head, y0, x0);
                const int s = n & 1;
                const uint32_t st = base + s * kStageQK;
                const uint32_t fb = qk_full + 8 * s;
                mbar_wait(qk_empty + 8 * s, ((n >> 1) & 1) ^ 1u);
                mbar_arrive_expect_tx(fb, kStageQK);
                for (int half = 0; half < 2; ++half) {  // 8 window rows = 128 tokens per box
                    tma_load_5d(st + kOffK + half * (kKMain / 2), &tmap_main, fb, 0, head, 1, x0, y0 + 8 * half);
                    tma_load_5d(st + kOffKt + half * (kKTail / 2), &tmap_tail, fb, 64, head, 1, x0, y0 + 8 * half);
                    tma_load_5d(st + kOffQ + half * kQMain, &tmap_main, fb, 0, head, 0, x0, y0 + 8 * half);
                    tma_load_5d(st + kOffQt + half * kQTail, &tmap_tail, fb, 64, head, 0, x0, y0 + 8 * half);
                }
                mbar_wait(v_empty, (n & 1) ^ 1u);
                mbar_arrive_expect_tx(v_full, kKMain + kKTail);
                for (int half = 0; half < 2; ++half) {
                    tma_load_5d(base + kOffV + half * (kKMain / 2), &tmap_main, v_full, 0, head, 2, x0, y0 + 8 * half);
                    tma_load_5d(base + kOffVt + half * (kKTail / 2), &tmap_tail, v_full, 64, head, 2, x0, y0 + 8 * half);
                }
            }
        }
    } else if (warp == 1 || warp == 10) {
        // ===================== MMA issuers: warp 1 serves query tile / TMEM region 0, warp 10 region 1 =====================
        // Each region is its own chain S -> softmax -> PV -> epilogue -> next S, and the chain's latency is the kernel's
        // time (ncu: the softmax warps wait on s_full / o_full for half of their cycles).  With ONE issuing thread a
        // region's ready step queued behind the other region's 32-MMA PV issue; one issuer per region answers at once.
        // tcgen05.commit tracks the issuing thread's own MMAs, so the shared Q/K and V buffers are released by one
        // commit from each issuer (barrier count 2).
        if (lane == 0) {
            const int r = warp == 1 ? 0 : 1;
#ifdef SPG_FP16
            constexpr uint32_t kFmt = 0u;
#else
            constexpr uint32_t kFmt = 1u;
#endif
            constexpr uint32_t kIdescBase = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((128u >> 4) << 24);
            constexpr uint32_t idesc_s = kIdescBase | ((256u >> 3) << 17);                      // K-major A and B
            constexpr uint32_t idesc_pv64 = kIdescBase | (1u << 16) | ((64u >> 3) << 17);       // B MN-major
            constexpr uint32_t idesc_pv16 = kIdescBase | (1u << 16) | ((16u >> 3) << 17);
            const uint32_t d = tmem_base + 256u * r;
            int n = 0;
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
                const uint32_t st = base + (n & 1) * kStageQK;
                const uint32_t qm = st + kOffQ + r * kQMain, qt = st + kOffQt + r * kQTail;
                mbar_wait(qk_full + 8 * (n & 1), (n >> 1) & 1);
                mbar_wait(region_free + 8 * r, (n & 1) ^ 1u);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(d, make_smem_desc(qm + 32u * k, 2, 1024, 16),
                                 make_smem_desc(st + kOffK + 32u * k, 2, 1024, 16), idesc_s, k != 0);
                // dims 64..79 from the 32B-swizzled tails (72..79 are TMA zero fill)
                umma_bf16_ss(d, make_smem_desc(qt, 6, 256, 16), make_smem_desc(st + kOffKt, 6, 256, 16), idesc_s, 1);
                umma_commit(s_full + 8 * r);
                umma_commit(qk_empty + 8 * (n & 1));  // this tile's share of "Q / K of item n consumed"
                mbar_wait(v_full, n & 1);
                mbar_wait(p_full + 8 * r, n & 1);
                tc_fence_after();
#pragma unroll 4
                for (int k = 0; k < kKeys / 16; ++k) {
                    const uint32_t a_tmem = d + 8u * k;  // 16 keys = 8 packed fp16-pair columns of P
                    umma_ts(d + 128, a_tmem, make_smem_desc(base + kOffV + 2048u * k, 2, 1024, 16), idesc_pv64, k != 0);
                    umma_ts(d + 192, a_tmem, make_smem_desc(base + kOffVt + 512u * k, 6, 256, 16), idesc_pv16, k != 0);
                }
                umma_commit(o_full + 8 * r);
                umma_commit(v_empty);  // this tile's share of "V of item n consumed"
            }
        }
    } else {
        // ============ softmax + epilogue: warps 2..5 own query tile 0, warps 6..9 query tile 1; thread == query row
        const int r = (warp - 2) >> 2;        // query tile / TMEM region of this warpgroup
        const int quarter = warp & 3;         // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;  // query row inside the 128-row tile
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + 256u * r;
        int n = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
            int head, y0, x0;
            decode(item, head, y0, x0);
            const uint32_t par = n & 1;
            mbar_wait(s_full + 8 * r, par);
            tc_fence_after();
            uint32_t ra[32], rb[32];
            // ---- pass 1: row maximum (32-column TMEM loads, double-buffered against the reductions)
            float mx0 = -INFINITY, mx1 = -INFINITY;
            tmem_ld32(lane_addr, ra);
#pragma unroll 1
            for (int c = 0; c < kKeys / 32; c += 2) {
                tmem_ld_wait();
                tmem_ld32(lane_addr + 32 * (c + 1), rb);
                row_max32(ra, mx0, mx1);
                tmem_ld_wait();
                if (c + 2 < kKeys / 32) tmem_ld32(lane_addr + 32 * (c + 2), ra);
                row_max32(rb, mx0, mx1);
            }
            const float m = fmaxf(mx0, mx1) * p.scale_log2e;
            // ---- pass 2: p = 2^(s * scale - m), row sum, pack to 16 bit, write back over the consumed columns
            unsigned long long sum2 = 0ull;  // {sum of even keys, sum of odd keys}
            const unsigned long long scale2 = splat2(p.scale_log2e), negm2 = splat2(-m);
            auto exp_pack_store = [&](const uint32_t (&raw)[32], int c) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = exp2_pair(raw[2 * i], raw[2 * i + 1], scale2, negm2, sum2);
                tmem_st16(lane_addr + 16 * c, pk);
            };
            tmem_ld32(lane_addr, ra);
#pragma unroll 1
            for (int c = 0; c < kKeys / 32; c += 2) {
                tmem_ld_wait();
                tmem_ld32(lane_addr + 32 * (c + 1), rb);
                exp_pack_store(ra, c);
                tmem_ld_wait();
                if (c + 2 < kKeys / 32) tmem_ld32(lane_addr + 32 * (c + 2), ra);
                exp_pack_store(rb, c + 1);
            }
            const float sum = hsum2(sum2);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full + 8 * r);
            // ---- epilogue: O / sum -> out[token, head*72 .. +72)
            const int q = r * 128 + row;
            const long long tok = (static_cast<long long>(y0) + (q >> 4)) * p.W + x0 + (q & 15);
            uint4* dst = reinterpret_cast<uint4*>(p.out + tok * p.D + head * kHd);
            mbar_wait(o_full + 8 * r, par);
            tc_fence_after();
            const float inv = 1.f / sum;
            // the whole O row moves to registers first and the region is handed back at once: the next item's S MMAs
            // run while this thread scales, packs and stores
            uint32_t o[5][16];
#pragma unroll
            for (int c = 0; c < 5; ++c) tmem_ld16(lane_addr + 128 + 16 * c, o[c]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(region_free + 8 * r);
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; i += 2)  // packed multiply (FMUL2), same roundings
                    f2_split(f2_mul(f2_make(__uint_as_float(o[c][i]), __uint_as_float(o[c][i + 1])), f2_make(inv, inv)), v[i], v[i + 1]);
                dst[2 * c] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                if (c < 4)  // dims 72..79 of the last chunk are padding
                    dst[2 * c + 1] = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]),
                                                pack2(v[14], v[15]));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}


// ------------------------------------------------------------------------------------------------------------------
// Global attention (window == 0: the three global blocks of stage 3, 1024 keys at 512^2 and 4096 at 1024^2) on
// tcgen05 / TMEM.  A work item is (image, head, 256 consecutive queries) = two 128-query tiles, one TMEM region each;
// the keys are streamed in blocks of 128 through two 3-stage TMA rings (K and V).  The softmax is exact and two-pass:
//   pass A  for every PAIR of key blocks: S = Q [K_j | K_j+1]^T (SS-MMAs, 256 columns = the whole region, O is not live
//           yet) -> the row's warp group folds it into the running maximum: one handshake per 256 keys (a handshake
//           MMA -> barrier -> TMEM load -> barrier costs ~1.5 us whatever the tile size, so fewer, larger units win:
//           64-key units with double-buffered S measured 803 us against 486 us for 128-key units);
//   pass B  for every key block: S again -> p = 2^(s * scale - m) written back to TMEM as packed 16-bit pairs over the
//           consumed S columns -> O += P V_j (TS-MMA, V MN-major) accumulating in TMEM across ALL key blocks.
// Recomputing S costs 1.5x the MMA work of an online softmax but needs no rescaling of O in TMEM, and the tensor pipe
// is the idle resource of this kernel (the exp2 on the MUFU and the TMEM round trips are the busy ones).  tcgen05.mma
// instructions issued by one thread execute in order, so S_{j+1} may be issued right behind PV_j although it overwrites
// the P_j columns.  Region layout (256 columns each): S / P in [0,128), O dims 0..63 in [128,192), dims 64..79 in
// [192,208).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGKeys = 128;                    // keys per block
constexpr int kGStages = 3;                    // K / V ring depth
constexpr uint32_t kGTileMain = 128 * 128;     // 128 tokens x dims 0..63 (SWIZZLE_128B)
constexpr uint32_t kGTileTail = 128 * 32;      // dims 64..79 (SWIZZLE_32B, 72..79 zero-filled)
constexpr uint32_t kGTile = kGTileMain + kGTileTail;
constexpr uint32_t kGOffQ = 0;                               // 2 query tiles: mains, then tails
constexpr uint32_t kGOffK = 2 * kGTile;                      // kGStages x (main | tail)
constexpr uint32_t kGOffV = kGOffK + kGStages * kGTile;
constexpr uint32_t kGOffBar = kGOffV + kGStages * kGTile;    // 163840
constexpr uint32_t kSmemG = kGOffBar + 512 + 1024;

struct AttnGlobalParams {
    h16* out;
    int B, H, W, D, heads;
    int rows_per_tile;  // grid rows per 128-token tile (128 / W)
    int tiles_per_img;  // H*W / 128
    int items;          // B * heads * tiles_per_img / 2
    int nkb;            // key blocks per image (= tiles_per_img)
    int reverse;
    float scale_log2e;
};

__global__ void __launch_bounds__(kThreadsTc, 1)
attention_tc_global_kernel(const __grid_constant__ CUtensorMap tmap_main, const __grid_constant__ CUtensorMap tmap_tail,
                           const AttnGlobalParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + kGOffBar;
    const uint32_t q_full = bar, q_empty = bar + 8, k_full = bar + 16, k_empty = bar + 40, v_full = bar + 64,
                   v_empty = bar + 88, s_full = bar + 112, sa_free = bar + 128, p_full = bar + 144, pv_done = bar + 160,
                   o_read = bar + 176, tmem_slot = bar + 192;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_main);
        tma_prefetch_desc(&tmap_tail);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < kGStages; ++s) {
            mbar_init(k_full + 8 * s, 1);
            mbar_init(k_empty + 8 * s, 1);
            mbar_init(v_full + 8 * s, 1);
            mbar_init(v_empty + 8 * s, 1);
        }
        for (int r = 0; r < 2; ++r) {
            mbar_init(s_full + 8 * r, 1);
            mbar_init(sa_free + 8 * r, 4);  // one arrive per warp of the region's warp group
            mbar_init(p_full + 8 * r, 4);
            mbar_init(pv_done + 8 * r, 1);
            mbar_init(o_read + 8 * r, 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int pairs_per_img = p.tiles_per_img / 2;
    // item -> (head, first grid row of the image, first query tile of the pair)
    auto decode = [&](int item, int& head, int& img_row0, int& qtile0) {
        if (p.reverse) item = p.items - 1 - item;
        head = item % p.heads;
        const int rest = item / p.heads;
        const int b = rest / pairs_per_img;
        qtile0 = 2 * (rest - b * pairs_per_img);
        img_row0 = b * p.H;
    };
    const int nkb = p.nkb;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int n = 0;
            uint32_t kload = 0, vload = 0;
            auto load_tile = [&](uint32_t dst, uint32_t fb, int head, int which, int row) {
                tma_load_5d(dst, &tmap_main, fb, 0, head, which, 0, row);
                tma_load_5d(dst + kGTileMain, &tmap_tail, fb, 64, head, which, 0, row);
            };
            auto load_k = [&](int head, int img_row0, int j) {
                const uint32_t s = kload % kGStages;
                mbar_wait(k_empty + 8 * s, ((kload / kGStages) & 1u) ^ 1u);
                mbar_arrive_expect_tx(k_full + 8 * s, kGTile);
                load_tile(base + kGOffK + s * kGTile, k_full + 8 * s, head, 1, img_row0 + j * p.rows_per_tile);
                ++kload;
            };
            for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
                int head, img_row0, qtile0;
                decode(item, head, img_row0, qtile0);
                mbar_wait(q_empty, (n & 1u) ^ 1u);
                mbar_arrive_expect_tx(q_full, 2 * kGTile);
                for (int r = 0; r < 2; ++r) {
                    tma_load_5d(base + kGOffQ + r * kGTileMain, &tmap_main, q_full, 0, head, 0, 0,
                                img_row0 + (qtile0 + r) * p.rows_per_tile);
                    tma_load_5d(base + kGOffQ + 2 * kGTileMain + r * kGTileTail, &tmap_tail, q_full, 64, head, 0, 0,
                                img_row0 + (qtile0 + r) * p.rows_per_tile);
                }
                for (int j = 0; j < nkb; ++j) load_k(head, img_row0, j);  // pass A
                for (int j = 0; j < nkb; ++j) {                            // pass B: K_j then V_j, in consumption order
                    load_k(head, img_row0, j);
                    const uint32_t s = vload % kGStages;
                    mbar_wait(v_empty + 8 * s, ((vload / kGStages) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(v_full + 8 * s, kGTile);
                    load_tile(base + kGOffV + s * kGTile, v_full + 8 * s, head, 2, img_row0 + j * p.rows_per_tile);
                    ++vload;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (event driven, one state machine per region) =====================
        if (lane == 0) {
#ifdef SPG_FP16
            constexpr uint32_t kFmt = 0u;
#else
            constexpr uint32_t kFmt = 1u;
#endif
            constexpr uint32_t kIdescBase = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((128u >> 4) << 24);
            constexpr uint32_t idesc_s = kIdescBase | ((128u >> 3) << 17);                      // N = 128 keys
            constexpr uint32_t idesc_pv64 = kIdescBase | (1u << 16) | ((64u >> 3) << 17);       // B MN-major
            constexpr uint32_t idesc_pv16 = kIdescBase | (1u << 16) | ((16u >> 3) << 17);
            const int my_items = (p.items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
            // per region: item, pass (0 = A, 1 = B), key block, whether the next step is the PV of that block
            int it_[2] = {0, 0}, ph_[2] = {0, 0}, j_[2] = {0, 0}, need_pv[2] = {0, 0};
            uint32_t kc[2] = {0, 0}, vc[2] = {0, 0};        // K / V blocks consumed so far (ring position)
            uint32_t sa_cnt[2] = {0, 0}, pf_cnt[2] = {0, 0}, or_cnt[2] = {0, 0};  // barrier completions consumed
            int k_uses[kGStages] = {0, 0, 0}, v_uses[kGStages] = {0, 0, 0};
            int q_uses = 0;
            uint32_t spins = 0;
            while (it_[0] < my_items || it_[1] < my_items) {
                bool progressed = false;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (it_[r] >= my_items) continue;
                    const uint32_t d = tmem_base + 256u * r;
                    if (!need_pv[r] && ph_[r] == 0) {
                        // ---- pass A: S = Q_r [K_j | K_j+1]^T, 256 keys per handshake, over the WHOLE region (O is not
                        // live in this pass; the previous item's epilogue must have read it: o_read)
                        const uint32_t ks0 = kc[r] % kGStages, ks1 = (kc[r] + 1) % kGStages;
                        if (!mbar_test(k_full + 8 * ks0, (kc[r] / kGStages) & 1u)) continue;
                        if (!mbar_test(k_full + 8 * ks1, ((kc[r] + 1) / kGStages) & 1u)) continue;
                        if (j_[r] == 0 && !mbar_test(q_full, it_[r] & 1u)) continue;
                        if (j_[r] == 0 && it_[r] > 0 && !mbar_test(o_read + 8 * r, or_cnt[r] & 1u)) continue;
                        if (j_[r] > 0 && !mbar_test(sa_free + 8 * r, sa_cnt[r] & 1u)) continue;
                        if (j_[r] == 0 && it_[r] > 0) ++or_cnt[r];
                        if (j_[r] > 0) ++sa_cnt[r];
                        tc_fence_after();
                        const uint32_t qm = base + kGOffQ + r * kGTileMain, qt = base + kGOffQ + 2 * kGTileMain + r * kGTileTail;
#pragma unroll
                        for (int hb = 0; hb < 2; ++hb) {
                            const uint32_t km = base + kGOffK + (hb ? ks1 : ks0) * kGTile;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_ss(d + 128u * hb, make_smem_desc(qm + 32u * k, 2, 1024, 16),
                                             make_smem_desc(km + 32u * k, 2, 1024, 16), idesc_s, k != 0);
                            umma_bf16_ss(d + 128u * hb, make_smem_desc(qt, 6, 256, 16), make_smem_desc(km + kGTileMain, 6, 256, 16),
                                         idesc_s, 1);
                        }
                        umma_commit(s_full + 8 * r);
                        if (++k_uses[ks0] == 2) {
                            k_uses[ks0] = 0;
                            umma_commit(k_empty + 8 * ks0);
                        }
                        if (++k_uses[ks1] == 2) {
                            k_uses[ks1] = 0;
                            umma_commit(k_empty + 8 * ks1);
                        }
                        kc[r] += 2;
                        j_[r] += 2;
                        if (j_[r] == nkb) {
                            ph_[r] = 1;
                            j_[r] = 0;
                        }
                        progressed = true;
                    } else if (!need_pv[r]) {
                        // ---- pass B: S = Q_r K_j^T
                        const uint32_t ks = kc[r] % kGStages;
                        if (!mbar_test(k_full + 8 * ks, (kc[r] / kGStages) & 1u)) continue;
                        // the region's S columns: pass A block j > 0 and the first block of pass B wait for the warp
                        // group to have read the previous pass-A scores; everything else is ordered by the MMA pipe
                        const bool after_a = j_[r] == 0;  // first block of pass B: the last pass-A scores have been read
                        if (after_a && !mbar_test(sa_free + 8 * r, (sa_cnt[r] & 1u))) continue;
                        if (after_a) ++sa_cnt[r];
                        tc_fence_after();
                        const uint32_t qm = base + kGOffQ + r * kGTileMain, qt = base + kGOffQ + 2 * kGTileMain + r * kGTileTail;
                        const uint32_t km = base + kGOffK + ks * kGTile;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_ss(d, make_smem_desc(qm + 32u * k, 2, 1024, 16), make_smem_desc(km + 32u * k, 2, 1024, 16),
                                         idesc_s, k != 0);
                        umma_bf16_ss(d, make_smem_desc(qt, 6, 256, 16), make_smem_desc(km + kGTileMain, 6, 256, 16), idesc_s, 1);
                        umma_commit(s_full + 8 * r);
                        if (++k_uses[ks] == 2) {
                            k_uses[ks] = 0;
                            umma_commit(k_empty + 8 * ks);  // both query tiles have used this K block
                        }
                        ++kc[r];
                        need_pv[r] = 1;
                        if (j_[r] == nkb - 1 && ++q_uses == 2) {
                            q_uses = 0;
                            umma_commit(q_empty);  // last S of the item issued for both tiles
                        }
                        progressed = true;
                    } else {
                        // ---- O_r += P_j V_j
                        const uint32_t vs = vc[r] % kGStages;
                        if (!mbar_test(v_full + 8 * vs, (vc[r] / kGStages) & 1u)) continue;
                        if (!mbar_test(p_full + 8 * r, pf_cnt[r] & 1u)) continue;
                        ++pf_cnt[r];  // (O was released by the previous item's epilogue before pass A started)
                        tc_fence_after();
                        const uint32_t vm = base + kGOffV + vs * kGTile;
#pragma unroll 4
                        for (int k = 0; k < kGKeys / 16; ++k) {
                            const uint32_t acc = (j_[r] | k) != 0;
                            umma_ts(d + 128, d + 8u * k, make_smem_desc(vm + 2048u * k, 2, 1024, 16), idesc_pv64, acc);
                            umma_ts(d + 192, d + 8u * k, make_smem_desc(vm + kGTileMain + 512u * k, 6, 256, 16), idesc_pv16, acc);
                        }
                        if (j_[r] == nkb - 1) umma_commit(pv_done + 8 * r);
                        if (++v_uses[vs] == 2) {
                            v_uses[vs] = 0;
                            umma_commit(v_empty + 8 * vs);
                        }
                        ++vc[r];
                        need_pv[r] = 0;
                        if (++j_[r] == nkb) {
                            j_[r] = 0;
                            ph_[r] = 0;
                            ++it_[r];
                        }
                        progressed = true;
                    }
                }
                if (progressed) {
                    spins = 0;
                } else if (++spins > (1u << 27)) {
                    printf("spg: attention_tc_global MMA issuer stuck block=%d items=(%d,%d) pass=(%d,%d) j=(%d,%d)\n",
                           (int)blockIdx.x, it_[0], it_[1], ph_[0], ph_[1], j_[0], j_[1]);
                    __trap();
                }
            }
        }
    } else {
        // ============ softmax + epilogue: warps 2..5 own query tile 0, warps 6..9 query tile 1; thread == query row
        const int r = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + 256u * r;
        uint32_t s_cnt = 0;  // s_full completions consumed
        int n = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
            int head, img_row0, qtile0;
            decode(item, head, img_row0, qtile0);
            uint32_t ra[32], rb[32];
            // ---- pass A: running maximum of the raw scores over all key blocks
            float mx0 = -INFINITY, mx1 = -INFINITY;
            for (int j = 0; j < nkb; j += 2) {  // 256 keys per handshake
                mbar_wait(s_full + 8 * r, s_cnt & 1u);
                ++s_cnt;
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 256; c += 64) {
                    tmem_ld32(lane_addr + c, ra);
                    tmem_ld32(lane_addr + c + 32, rb);
                    tmem_ld_wait();
                    row_max32(ra, mx0, mx1);
                    row_max32(rb, mx0, mx1);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(sa_free + 8 * r);
            }
            const float m = fmaxf(mx0, mx1) * p.scale_log2e;
            // ---- pass B: probabilities of every key block, packed in place; O accumulates in TMEM
            unsigned long long sum2 = 0ull;  // {sum of even keys, sum of odd keys}
            const unsigned long long scale2 = splat2(p.scale_log2e), negm2 = splat2(-m);
            auto exp_pack_store = [&](const uint32_t (&raw)[32], int c) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = exp2_pair(raw[2 * i], raw[2 * i + 1], scale2, negm2, sum2);
                tmem_st16(lane_addr + 16 * c, pk);
            };
            for (int j = 0; j < nkb; ++j) {
                mbar_wait(s_full + 8 * r, s_cnt & 1u);
                ++s_cnt;
                tc_fence_after();
                tmem_ld32(lane_addr, ra);
                tmem_ld_wait();
                tmem_ld32(lane_addr + 32, rb);
                exp_pack_store(ra, 0);
                tmem_ld_wait();
                tmem_ld32(lane_addr + 64, ra);
                exp_pack_store(rb, 1);
                tmem_ld_wait();
                tmem_ld32(lane_addr + 96, rb);
                exp_pack_store(ra, 2);
                tmem_ld_wait();
                exp_pack_store(rb, 3);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full + 8 * r);
            }
            // ---- epilogue: O / sum -> out[token, head*72 .. +72)
            const long long tok = (static_cast<long long>(img_row0) * p.W) + static_cast<long long>(qtile0 + r) * 128 + row;
            uint4* dst = reinterpret_cast<uint4*>(p.out + tok * p.D + head * kHd);
            mbar_wait(pv_done + 8 * r, n & 1u);
            tc_fence_after();
            const float inv = 1.f / (hsum2(sum2));
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                uint32_t raw[16];
                tmem_ld16(lane_addr + 128 + 16 * c, raw);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; i += 2)
                    f2_split(f2_mul(f2_make(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1])), f2_make(inv, inv)), v[i], v[i + 1]);
                dst[2 * c] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                if (c < 4)
                    dst[2 * c + 1] = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]),
                                                pack2(v[14], v[15]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_read + 8 * r);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace spg

// tcgen05 path of spg_window_attention_h16 for window == 16 without query pooling (the 32 windowed stage-3 blocks).
// Returns SPG_ERR_UNSUPPORTED for any other geometry; the caller then uses the generic kernel.
extern "C" int spg_window_attention_tc_h16(const void* qkv, void* out, int B, int H, int W, int D, int heads,
                                           int window, int q_pool, const spg_launch_t* launch) {
    using namespace spg;
    SPG_CHECK_ARG(qkv && out, "null pointer");
    if (window == 0 && !q_pool && D == heads * kHd && H == W && (W == 32 || W == 64 || W == 128) && (H * W) % 256 == 0) {
        // global attention: two-pass tcgen05 kernel over 128-key blocks
        AttnGlobalParams g{};
        g.out = static_cast<h16*>(out);
        g.B = B; g.H = H; g.W = W; g.D = D; g.heads = heads;
        g.rows_per_tile = 128 / W;
        g.tiles_per_img = H * W / 128;
        g.nkb = g.tiles_per_img;
        g.items = B * heads * (g.tiles_per_img / 2);
        g.reverse = LaunchCtx(launch).reverse ? 1 : 0;
        g.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(kHd));
        CUtensorMap gm, gt;
        if (int rc = make_tmap_qkv_5d(&gm, qkv, static_cast<uint64_t>(B) * H, W, heads, 64, W, g.rows_per_tile, 128)) return rc;
        if (int rc = make_tmap_qkv_5d(&gt, qkv, static_cast<uint64_t>(B) * H, W, heads, 16, W, g.rows_per_tile, 32)) return rc;
        static PerDeviceOnce g_attr_set;
        if (g_attr_set.needed()) {
            SPG_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_global_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemG));
            g_attr_set.done();
        }
        const int ggrid = g.items < sm_count() ? g.items : sm_count();
        SPG_CHECK_CUDA((launch_pdl(attention_tc_global_kernel, ggrid, kThreadsTc, kSmemG, LaunchCtx(launch), gm, gt, g)));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        SPG_CHECK_LAUNCH();
        return SPG_OK;
    }
    if (window != kWs || q_pool || H % kWs || W % kWs || D != heads * kHd)
        return fail(SPG_ERR_UNSUPPORTED, "tcgen05 attention covers 16x16 windows and global attention, without query pooling");
    AttnTcParams p{};
    p.out = static_cast<h16*>(out);
    p.B = B; p.H = H; p.W = W; p.D = D; p.heads = heads;
    p.nwx = W / kWs; p.nwy = H / kWs;
    p.items = B * p.nwx * p.nwy * heads;
    p.reverse = LaunchCtx(launch).reverse ? 1 : 0;
    p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(kHd));
    CUtensorMap tmain, ttail;
    if (int rc = make_tmap_qkv_5d(&tmain, qkv, static_cast<uint64_t>(B) * H, W, heads, 64, kWs, 8, 128)) return rc;
    if (int rc = make_tmap_qkv_5d(&ttail, qkv, static_cast<uint64_t>(B) * H, W, heads, 16, kWs, 8, 32)) return rc;
    static PerDeviceOnce attr_set;
    if (attr_set.needed()) {
        SPG_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTc));
        attr_set.done();
    }
    const int grid = p.items < sm_count() ? p.items : sm_count();
    SPG_CHECK_CUDA((launch_pdl(attention_tc_kernel, grid, kThreadsWin, kSmemTc, LaunchCtx(launch), tmain, ttail, p)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    SPG_CHECK_LAUNCH();
    return SPG_OK;
}
