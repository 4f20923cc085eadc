// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] * W[N,K]^T )        16-bit operands (fp16 / bf16), fp32 accumulation in TMEM
//
// One CTA per SM loops over 128 x block_n output tiles (n fastest, so concurrently running CTAs
// share A tiles through L2).  Roles (320 threads):
//   warp 0      TMA producer   - fills a ring of {A 128x64, W block_n x 64} stages (128B swizzle)
//   warp 1      MMA issuer     - one thread issues tcgen05.mma (M=128, N=block_n, K=16) x4 per stage,
//                                tcgen05.commit releases the stage / publishes the accumulator
//   warps 2..9  epilogue       - two warps per TMEM lane quarter split the tile's 16-column chunks;
//                                thread == output row.  Per chunk: tcgen05.ld (double-buffered) -> bias
//                                (staged in smem per tile) -> GELU / ReLU -> fp32 residual -> fused N->1
//                                head -> swizzled smem staging -> TMA store.  The fp32 residual chunk is
//                                TMA-loaded into the same staging buffer two chunks ahead, so neither
//                                the residual read nor the output write touches the LSU with a
//                                row-per-thread (32 cache lines per instruction) access pattern.
// The accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps
// the MMAs of tile i+1.
//
// Convolution mode: the A tile of k-chunk (tap, channel-chunk) is a 4-D TMA box
// {64 ch, tile_w, tile_h, 1 image} of the NHWC input at (x0+dx-1, y0+dy-1); the zero halo comes
// from TMA out-of-bounds fill, so the same MMA / epilogue pipeline serves both modes.
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "half16.cuh"
#include "ln_stats.cuh"
#include "ptx.cuh"

namespace spg {

extern std::atomic<long long> g_launches;

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kAStageBytes = kBlockM * kBlockK * 2;
constexpr int kHaloABytes = 17 * 1024;  // 130 halo pixels x 128 B = 16640 B, padded to a 1 KB multiple
// warp 0 TMA, warp 1 MMA, warps 2.. epilogue.  The kernel is written for 4 * k epilogue warps (k warps per TMEM lane
// quarter splitting the tile's columns); every instance runs 8.  16 warps were measured on B200 and gain nothing
// (fc1+GELU 963 vs 971 TFLOP/s, stage-1 fc1 470 vs 468 us): short-K shapes are bound by the TMA round trip of a
// smem-limited ring, not by epilogue latency hiding.
constexpr int kEpiWarpsDefault = 8;
constexpr int kTmemCols = 512;
constexpr int kAccStageCols = 256;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kBarBytes = 1024;                                // pipeline barriers + residual barriers
// Epilogue scratch, sized per launch in whole KB (every shared-memory piece is a 1 KB multiple, so nothing is lost to
// alignment and the operand ring gets what is left): bias[2][256] | LayerNorm modes: column sums or gamma [2][256],
// beta [2][256] | fused head: weights[256], partials[3][128]
constexpr int kScratchBias = 2 * 256 * 4;
constexpr int kScratchLn = 4 * 256 * 4;
constexpr int kScratchHead = 3 * 1024;
inline int epi_scratch_bytes(int ln_mode, bool head) { return kScratchBias + (ln_mode ? kScratchLn : 0) + (head ? kScratchHead : 0); }
constexpr int kLnRec = 32;                                      // floats per LayerNorm row record {c, P, (s1, s2) x P}
constexpr int kLnBufBytes = 32 * 32;                            // staging buffer of the 16-bit centred copy: 32 rows x 16 columns
// LayerNorm applied by the producer (kLn == 3): staging ring of the normalised 16-bit rows per epilogue warp, and the
// exchange area of the per-(n-tile, column slice) row statistics {mean, M2} that the CTAs of a cluster write into each
// other's shared memory (double-buffered across tiles)
constexpr int kLnSlots = 2;
constexpr int kLnMaxParts = kLnMaxSlices;
constexpr int kLnXBytes = 2 * kLnMaxParts * kBlockM * 8;
#ifndef SPG_RES_SLOTS
#define SPG_RES_SLOTS 3
#endif
// staging ring depth; with a residual, kResSlots - 1 residual groups are in flight per epilogue warp
constexpr int kResSlots = SPG_RES_SLOTS;
static_assert(kResSlots >= 3 && kResSlots <= 8, "staging ring depth");

struct GemmArgs {
    int M, N;
    int block_n;
    int num_m_tiles, num_n_tiles, num_k_chunks;
    int last_chunk_ksteps;  // 16-deep MMA steps that hold real data in the LAST 64-wide k-chunk (K tail; 4 when K % 64 == 0)
    int stages;
    int reverse;     // walk the tiles in descending order (see common.h "Traversal direction")
    int l2_prefetch; // linear mode: prefetch the next tile's A rows into L2 one tile ahead
    int pair;        // launch as CTA pairs (cta_group::2): decided on the host, selects the kPair kernel instance
    int halo;        // conv only: one stage = a 130-pixel halo row of A + the 3 dx-tap weight tiles (A reuse x3)
    // linear, short K (<= 3 k-chunks): the CTA's weight tile stays RESIDENT in shared memory (the grid is a multiple of
    // the n-tiles, so a CTA keeps one n-block for all its tiles) and the ring holds A chunks only: 2-3 tiles in flight
    // per TMA round trip instead of 1.3 -- these GEMMs are bound by that round trip, not by the tensor pipe
    int b_resident;
    int b_res_tiles;  // weight tiles held resident: k-chunks (linear) or 9 taps x channel chunks (row-halo conv)
    // ln_mode 3 (the CTAs / pairs of a cluster work on the n-tiles of ONE row block at the same time): every A chunk is
    // fetched from L2 by one CTA of the cluster and TMA-multicast to the CTAs that need the same rows (k-chunk kc by
    // cluster position kc mod #n-tiles): the L2 -> SM traffic of A drops by #n-tiles.  A stage is then released by the
    // MMA issuers of ALL n-tiles (commit multicast to every CTA of the cluster).
    int a_mcast;
    int scratch_bytes;  // epilogue scratch of this launch (epi_scratch_bytes)
    // conv mode
    int conv;
    int H, W, cin_chunks, tile_w;
    // conv modes: an m-tile is tile_h x tile_w pixels (tile_h * tile_w = 128) of one image; tiles_x tiles cover an image
    // row band (tiles_x * tile_w >= W), tiles_img = tiles_x * H / tile_h tiles an image.  `ragged` (W % tile_w != 0): the
    // pixels with x >= W of the last tile column do not exist -- their A rows are TMA zero fill, their outputs are
    // clipped by a 4-D store map / skipped by the fused head.
    int tiles_x, tiles_img, ragged, tile_h;
    // fused bilinear x2 upsample (kUp2 instances): the conv runs on the LOW-resolution grid with 4 output phases stacked
    // along N; w holds one [N, K] weight set per row class (top / interior / bottom image row), corr the pre-activation
    // corrections of the first / last image column, the store is a pixel shuffle (see spg_conv3x3_up2_h16)
    int up2;
    int cout;            // channels per phase (N = 4 * cout)
    int bh;              // B * H
    const float* corr;   // [2][B*H][N] fp32
    // epilogue
    const float* bias;
    int act;
    int has_res;
    int res_rows;
    int has_out;
    int out_f32;
    int epi_warps;    // 8 or 16 epilogue warps (selects the kernel instance)
    int group;        // 16-column chunks per staging buffer / TMA op (1 or 2)
    int row_bytes;    // bytes per staging row = group * 16 * sizeof(out)   (32 / 64 / 128)
    int buf_bytes;    // 32 * row_bytes
    int piece_shift, piece_mask;  // swizzle of the 16-byte pieces of a staging row
    const float* head_w;
    float head_b;
    float* head_out;
    // LayerNorm folding (see spg_epilogue_t): consumer = fold_rec / fold_cw, producer = emit_rec / prev_rec (+ tmap_ln)
    const float* ln_fold_rec;
    const float* ln_fold_cw;
    float* ln_emit_rec;
    const float* ln_prev_rec;
    float ln_inv_cols, ln_eps;
    int ln_mode;      // 0 none, 1 consumer (fold), 2 producer (emit), 3 producer that applies the LayerNorm itself
    // ln_mode 3: out = fp32 residual stream as usual, plus y = LayerNorm(out) * gamma + beta stored as 16 bit through
    // tmap_ln.  The CTAs that hold the n-tiles of one 128-row block form a cluster and exchange row statistics.
    const float* ln_gamma;
    const float* ln_beta;
};

#ifdef SPG_TRACE
// Timeline instrumentation (variant builds only: tools/build_variant.sh trace -DSPG_TRACE): SM-clock stamps of CTA 0's
// producer / MMA issuer / epilogue warp 2 for its first 64 tiles + the cycles the issuer spent waiting for operands,
// dumped by spg_debug_trace_dump (tests/cuda/test_gemm.cu --shape prints them).  profiles/r01_gemm_timeline.md.
__device__ long long g_trace[64 * 12];
#define SPG_STAMP(tile_i, slot)                                                           \
    do {                                                                                  \
        if (blockIdx.x == 0 && (tile_i) < 64) g_trace[(tile_i) * 12 + (slot)] = clock64(); \
    } while (0)
#else
#define SPG_STAMP(tile_i, slot) do {} while (0)
#endif

// Exact-erf GELU without libdevice erff (which costs ~4x the issue slots because both of its branches are
// predicated).  With Phi the normal CDF, gelu(x) = x Phi(x) = max(x, 0) - |x| (1 - Phi(|x|)) and
//   1 - Phi(t) = erfc(t / sqrt 2) / 2 = 2^P(t),   P = degree-4 fit of log2(erfc(t / sqrt 2)) - 1 on [0, 5.65] that
// minimises the error of the PRODUCT t 2^P(t) (weighted minimax), t = min(|x|, 5.65) (beyond it the term is
// < 1e-8 |x|).  5 FMA + 2 FMNMX + one MUFU.EX2 per element; max |error| 6.2e-6 over all x when evaluated in fp32,
// 20x below the rounding of the fp16 result at |gelu| ~ 0.25 (the degree-6 form it replaces reached 3.1e-7 for two more
// FMAs per element in an epilogue that is bound by instruction issue; fit + check: DESIGN.md "Numerics").
__device__ __forceinline__ float gelu_erf(float x) {
    const float ax = fabsf(x);
    const float t = fminf(ax, 5.65f);
    float q = 0.0038662622682750225f;
    q = fmaf(q, t, -0.044079262763261795f);
    q = fmaf(q, t, -0.46801483631134033f);
    q = fmaf(q, t, -1.1473722457885742f);
    q = fmaf(q, t, -1.0004795789718628f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
    return fmaf(-ax, e, fmaxf(x, 0.f));
}

// The same function on two elements at once: the polynomial and the final multiply-add on packed fp32 pairs (FFMA2 =
// two IEEE FMAs per issue slot, so the results are bit-identical to gelu_erf); |x|, min, max and the MUFU stay scalar.
// 13 issue slots per pair instead of 18: the GELU is the largest part of an epilogue that costs 11 % of the fc1 launch
// under the power cap (profiles/r02_sustained_gemm_probe.md).
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
    const float a0 = fabsf(x0), a1 = fabsf(x1);
    const f32x2 t = f2_make(fminf(a0, 5.65f), fminf(a1, 5.65f));
    f32x2 q = f2_make(0.0038662622682750225f, 0.0038662622682750225f);
    q = f2_fma(q, t, f2_make(-0.044079262763261795f, -0.044079262763261795f));
    q = f2_fma(q, t, f2_make(-0.46801483631134033f, -0.46801483631134033f));
    q = f2_fma(q, t, f2_make(-1.1473722457885742f, -1.1473722457885742f));
    q = f2_fma(q, t, f2_make(-1.0004795789718628f, -1.0004795789718628f));
    float q0, q1, e0, e1;
    f2_split(q, q0, q1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
    const f32x2 r = f2_fma(f2_make(-a0, -a1), f2_make(e0, e1), f2_make(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
    f2_split(r, x0, x1);
}

// Epilogue traits are compile-time constants for the combinations the model launches (the hot loop then
// carries no flag tests); -1 selects the run-time value from GemmArgs (generic fallback instance).
// kPair = 1: the kernel runs as clusters of two CTAs (tcgen05 cta_group::2).  A pair owns a 256 x block_n
// tile: each CTA TMA-loads its own 128 A rows and HALF of the weight tile, the leader (cluster rank 0) issues
// M=256 MMAs that read both CTAs' smem and fill both CTAs' TMEM, every CTA runs the epilogue of its 128 rows.
// Staging half of W per CTA shrinks the stage (28 KB instead of 40 KB at N=192): more stages in flight per
// TMA round trip and 30 % less L2->smem traffic per FLOP.
template <int kAct, int kOutF32, int kHasRes, int kHasHead, int kHasOut, int kPair, int kEpiWarps, int kUp2 = 0, int kLn = 0>
__global__ void __launch_bounds__(64 + 32 * kEpiWarps, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                    const __grid_constant__ CUtensorMap tmap_ln, const GemmArgs p) {
    const int act = kAct < 0 ? p.act : kAct;
    const bool out_f32 = kOutF32 < 0 ? p.out_f32 != 0 : kOutF32 != 0;
    const bool has_res = kHasRes < 0 ? p.has_res != 0 : kHasRes != 0;
    const bool has_head = kHasHead < 0 ? p.head_w != nullptr : kHasHead != 0;
    const bool has_out = kHasOut < 0 ? p.has_out != 0 : kHasOut != 0;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
    // cluster rank; a cluster is one CTA pair, or (kLn == 3) every CTA / pair holding an n-tile of the same row block
    const uint32_t crank = (kPair || kLn == 3) ? cluster_ctarank() : 0u;
    const uint32_t rank = kPair ? (crank & 1u) : 0u;  // position inside the CTA pair
    const bool leader = rank == 0;
    // bytes of ONE weight tile held by this CTA (pair mode: half of the block_n rows)
    const uint32_t b_stage_bytes = static_cast<uint32_t>(kPair ? p.block_n / 2 : p.block_n) * 128u;
    // halo mode: A stage = 130 pixels x 128 B (padded to 17 KB so stages stay 1 KB aligned) + three weight tiles
    const uint32_t a_stage_bytes = p.halo ? kHaloABytes : kAStageBytes;
    const uint32_t stage_bytes = a_stage_bytes + (p.b_resident ? 0u : (p.halo ? 3u : 1u) * b_stage_bytes);
    const uint32_t b_res_addr = tiles_addr + static_cast<uint32_t>(p.stages) * stage_bytes;  // resident weight tile (b_resident)
    const uint32_t bar_addr = b_res_addr + (p.b_resident ? static_cast<uint32_t>(p.b_res_tiles) * b_stage_bytes : 0u);
    auto full_bar = [&](int s) { return bar_addr + 8u * s; };
    auto empty_bar = [&](int s) { return bar_addr + 8u * (p.stages + s); };
    auto tmem_full_bar = [&](int a) { return bar_addr + 8u * (2 * p.stages + a); };
    auto tmem_empty_bar = [&](int a) { return bar_addr + 8u * (2 * p.stages + 2 + a); };
    const uint32_t tmem_slot = bar_addr + 8u * (2 * p.stages + 4);
    const uint32_t b_res_bar = bar_addr + 8u * (2 * p.stages + 5);  // the resident weight tile has landed
    auto res_bar = [&](int ew, int slot) { return bar_addr + 256u + 8u * (ew * kResSlots + slot); };
    auto ln_bar = [&](int b) { return bar_addr + 768u + 8u * b; };  // kLn == 3: statistics of tile parity b have arrived
    const uint32_t scratch_addr = bar_addr + kBarBytes;
    const uint32_t staging_addr = (scratch_addr + static_cast<uint32_t>(p.scratch_bytes) + 1023u) & ~1023u;  // 128B-swizzle atoms: 1 KB aligned

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // work distribution: a "unit" is a CTA (1-CTA mode) or a CTA pair; its m-block is 128 or 256 rows
    const int total_tiles = (kPair ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles) * p.num_n_tiles;
    const int unit = kPair ? blockIdx.x >> 1 : blockIdx.x;
    const int nunits = kPair ? gridDim.x >> 1 : gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if (has_out) tma_prefetch_desc(&tmap_out);
        if (has_res) tma_prefetch_desc(&tmap_res);
        if (kLn == 2) tma_prefetch_desc(&tmap_ln);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), (kLn == 3 && p.a_mcast) ? static_cast<uint32_t>(p.num_n_tiles) : 1u);
        }
        mbar_init(b_res_bar, 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), kPair ? 2 * kEpiWarps : kEpiWarps);  // one arrive per epilogue warp (both CTAs)
        }
        for (int ew = 0; ew < kEpiWarps; ++ew)
            for (int s = 0; s < kResSlots; ++s) mbar_init(res_bar(ew, s), 1);
        if (kLn == 3)  // every epilogue WARP of every CTA holding an n-tile of the row block arrives once per tile
            for (int b = 0; b < 2; ++b) mbar_init(ln_bar(b), static_cast<uint32_t>(p.num_n_tiles) * kEpiWarps);
        fence_barrier_init();
    }
    if (warp == 1) {
        if (kPair) {
            tmem_alloc_pair(tmem_slot, kTmemCols);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_slot, kTmemCols);
            tmem_relinquish();
        }
    }
    pdl_launch_dependents();
    tc_fence_before();
    if (kPair || kLn == 3) cluster_sync_all();  // the peers' barriers / TMEM must exist before any remote arrive or 2-CTA MMA
    else __syncthreads();
    tc_fence_after();
    pdl_wait();  // prologue done (barriers, TMEM, descriptor prefetch); operands of the previous kernel are read below
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            // pair mode: both CTAs load; all transaction bytes are signalled on the leader's full barrier
            const uint32_t n_half = kPair ? rank * (p.block_n / 2) : 0u;  // this CTA's slice of the weight tile
            // a_mcast: position of this CTA (pair) among the n-tiles of the cluster, and the CTAs that share its A rows
            const bool mcast = kLn == 3 && p.a_mcast != 0;
            const int cpos = static_cast<int>(kPair ? crank >> 1 : crank);
            uint16_t a_mask = 0;
            for (int q = 0; q < p.num_n_tiles; ++q) a_mask |= static_cast<uint16_t>(1u << (kPair ? 2 * q + static_cast<int>(rank) : q));
            int a_turn = 0;  // k-chunk counter mod #n-tiles (runs across tiles: whole chunks, any K)
            [[maybe_unused]] int trace_i = 0;
            if (p.b_resident && unit < total_tiles) {
                // this CTA's n-block never changes (grid % n-tiles == 0): load its weight tile once
                const int t0 = p.reverse ? total_tiles - 1 - unit : unit;
                const int n_fixed = t0 % p.num_n_tiles;
                // (linear: tile kc = k-chunk kc; row-halo conv: tile (tap, cc) = columns ((tap * cin_chunks + cc) * 64 ..) of w)
                mbar_arrive_expect_tx(b_res_bar, static_cast<uint32_t>(p.b_res_tiles) * b_stage_bytes);
                for (int kc = 0; kc < p.b_res_tiles; ++kc)
                    tma_load_2d(b_res_addr + kc * b_stage_bytes, &tmap_b, b_res_bar, kc * kBlockK, n_fixed * p.block_n);
            }
            for (int tile = unit; tile < total_tiles; tile += nunits, ++trace_i) {
                SPG_STAMP(trace_i, 0);
                const int t_idx = p.reverse ? total_tiles - 1 - tile : tile;
                const int m_unit = t_idx / p.num_n_tiles;
                const int n_blk = t_idx - m_unit * p.num_n_tiles;
                const int m_blk = kPair ? 2 * m_unit + static_cast<int>(rank) : m_unit;  // this CTA's 128-row block
                int img = 0, y0 = 0, x0 = 0;
                if (p.conv) {
                    img = m_blk / p.tiles_img;
                    const int rem = m_blk - img * p.tiles_img;
                    const int ty = rem / p.tiles_x;
                    y0 = ty * p.tile_h;
                    x0 = (rem - ty * p.tiles_x) * p.tile_w;
                }
                // The first touch of an m-block's A rows comes from HBM (the weights and every later n-tile hit L2), and the
                // ring only looks `stages` k-chunks (~1.5 us) ahead: each tile would start with an HBM-latency bubble
                // (measured as a fixed ~1.6 us per tile, which is what separates K = 576 from long-K efficiency).  So the
                // A rows of this CTA's NEXT tile are prefetched into L2 now, one whole tile ahead; the CTAs working on the
                // same m-block share the job (chunk kc goes to the CTA whose n-block is kc mod #n-tiles).
                if (p.l2_prefetch) {
                    const int nt_raw = tile + nunits;
                    if (nt_raw < total_tiles) {
                        const int nt = p.reverse ? total_tiles - 1 - nt_raw : nt_raw;
                        const int nm_unit = nt / p.num_n_tiles;
                        const int nn_blk = nt - nm_unit * p.num_n_tiles;
                        const int nm_blk = kPair ? 2 * nm_unit + static_cast<int>(rank) : nm_unit;
                        for (int kc = nn_blk; kc < p.num_k_chunks; kc += p.num_n_tiles)
                            tma_prefetch_l2_2d(&tmap_a, kc * kBlockK, nm_blk * kBlockM);
                    }
                }
                // kUp2: the tile is part of one image row; rows 0 / H-1 use their own weight sets (stacked along N)
                const int w_row0 = kUp2 ? (y0 == 0 ? 0 : (y0 == p.H - 1 ? 2 * p.N : p.N)) : 0;
                for (int kc = 0; kc < p.num_k_chunks; ++kc) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t a_dst = tiles_addr + stage * stage_bytes;
                    const uint32_t b_dst = a_dst + a_stage_bytes;
                    if (p.halo) {
                        // k-chunk = (dy, 64-channel chunk): one 130-pixel halo row serves the taps dx = -1, 0, +1
                        const int dyi = kc / p.cin_chunks;
                        const int cc = kc - dyi * p.cin_chunks;
                        const uint32_t bytes = 130u * 128u + (p.b_resident ? 0u : 3u * b_stage_bytes);
                        if (leader) mbar_arrive_expect_tx(full_bar(stage), kPair ? 2u * bytes : bytes);
                        if (kPair) tma_load_4d_pair(a_dst, &tmap_a, full_bar(stage), cc * kBlockK, x0 - 1, y0 + dyi - 1, img);
                        else tma_load_4d(a_dst, &tmap_a, full_bar(stage), cc * kBlockK, x0 - 1, y0 + dyi - 1, img);
#pragma unroll
                        for (int dxi = 0; dxi < 3; ++dxi) {
                            if (p.b_resident) break;  // the nine tap tiles are resident
                            const int kcol = ((dyi * 3 + dxi) * p.cin_chunks + cc) * kBlockK;
                            const int nrow = w_row0 + n_blk * p.block_n + n_half;
                            if (kPair) tma_load_2d_pair(b_dst + dxi * b_stage_bytes, &tmap_b, full_bar(stage), kcol, nrow);
                            else tma_load_2d(b_dst + dxi * b_stage_bytes, &tmap_b, full_bar(stage), kcol, nrow);
                        }
                        if (++stage == p.stages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                        continue;
                    }
                    if (p.b_resident) {  // A chunk only; the weights are resident
                        mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
                        tma_load_2d(a_dst, &tmap_a, full_bar(stage), kc * kBlockK, m_blk * kBlockM);
                        if (++stage == p.stages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                        continue;
                    }
                    if (leader) mbar_arrive_expect_tx(full_bar(stage), kPair ? 2u * stage_bytes : stage_bytes);
                    if (p.conv) {
                        const int tap = kc / p.cin_chunks;
                        const int cc = kc - tap * p.cin_chunks;
                        const int dy = tap / 3 - 1;
                        const int dx = tap - (tap / 3) * 3 - 1;
                        if (kPair) tma_load_4d_pair(a_dst, &tmap_a, full_bar(stage), cc * kBlockK, x0 + dx, y0 + dy, img);
                        else tma_load_4d(a_dst, &tmap_a, full_bar(stage), cc * kBlockK, x0 + dx, y0 + dy, img);
                    } else if (mcast) {
                        if (a_turn == cpos) {
                            if (kPair) tma_load_2d_pair_mcast(a_dst, &tmap_a, full_bar(stage), kc * kBlockK, m_blk * kBlockM, a_mask);
                            else tma_load_2d_mcast(a_dst, &tmap_a, full_bar(stage), kc * kBlockK, m_blk * kBlockM, a_mask);
                        }
                        if (++a_turn == p.num_n_tiles) a_turn = 0;
                    } else {
                        if (kPair) tma_load_2d_pair(a_dst, &tmap_a, full_bar(stage), kc * kBlockK, m_blk * kBlockM);
                        else tma_load_2d(a_dst, &tmap_a, full_bar(stage), kc * kBlockK, m_blk * kBlockM);
                    }
                    if (kPair) tma_load_2d_pair(b_dst, &tmap_b, full_bar(stage), kc * kBlockK, w_row0 + n_blk * p.block_n + n_half);
                    else tma_load_2d(b_dst, &tmap_b, full_bar(stage), kc * kBlockK, w_row0 + n_blk * p.block_n);
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                SPG_STAMP(trace_i, 1);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0 && leader) {
            const uint32_t idesc = make_idesc_bf16_f32(kPair ? 2 * kBlockM : kBlockM, p.block_n);
            auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t accum) {
                if (kPair) umma_bf16_ss_pair(d, a, b, idesc, accum);
                else umma_bf16_ss(d, a, b, idesc, accum);
            };
            auto commit = [&](uint32_t bar) {
                if (kPair) umma_commit_pair(bar, static_cast<uint16_t>(3u << (crank & ~1u)));
                else umma_commit(bar);
            };
            // a_mcast: the stage is shared by the whole cluster -- release it on every CTA's empty barrier
            const bool mcast = kLn == 3 && p.a_mcast != 0;
            const uint16_t all_mask = static_cast<uint16_t>((1u << ((kPair ? 2 : 1) * p.num_n_tiles)) - 1u);
            auto release = [&](uint32_t bar) {
                if (!mcast) commit(bar);
                else if (kPair) umma_commit_pair(bar, all_mask);
                else umma_commit_mcast(bar, all_mask);
            };
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            [[maybe_unused]] int trace_i = 0;
            if (p.b_resident && unit < total_tiles) {
                mbar_wait(b_res_bar, 0);
                tc_fence_after();
            }
            for (int tile = unit; tile < total_tiles; tile += nunits, ++trace_i) {
                SPG_STAMP(trace_i, 2);
                mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                SPG_STAMP(trace_i, 3);
                const uint32_t d_tmem = tmem_base + acc * kAccStageCols;
#ifdef SPG_TRACE
                long long starved = 0;
#endif
                for (int kc = 0; kc < p.num_k_chunks; ++kc) {
#ifdef SPG_TRACE
                    const long long w0 = clock64();
#endif
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
#ifdef SPG_TRACE
                    if (kc > 0) starved += clock64() - w0;  // cycles the issuer waited for operands (first chunk excluded)
                    if (kc == 0) SPG_STAMP(trace_i, 4);
#endif
                    const uint32_t a_addr = tiles_addr + stage * stage_bytes;
                    if (p.halo) {
                        // tap dx reads the same halo rows shifted by dx pixels: start address + dx * 128 B.  The 128 B
                        // swizzle is a function of absolute smem address bits, so a start that is not 1024 B aligned
                        // needs no base offset in the descriptor (measured: base_offset 0 is the convention that
                        // reproduces the reference, tests/cuda/test_gemm.cu)
#pragma unroll
                        for (int dxi = 0; dxi < 3; ++dxi) {
                            const uint64_t a_desc = make_sw128_kmajor_desc(a_addr + dxi * 128u);
                            // k-chunk kc = (dy, channel chunk): tap (dy, dx) of chunk cc is resident tile (dy*3 + dx) * cin_chunks + cc
                            const int dyi = kc / p.cin_chunks;
                            const uint64_t b_desc = make_sw128_kmajor_desc(
                                p.b_resident ? b_res_addr + static_cast<uint32_t>((dyi * 3 + dxi) * p.cin_chunks + (kc - dyi * p.cin_chunks)) * b_stage_bytes
                                             : a_addr + a_stage_bytes + dxi * b_stage_bytes);
#pragma unroll
                            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                                const uint64_t koff = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
                                mma(d_tmem, a_desc + koff, b_desc + koff, (kc | dxi | k) != 0);
                            }
                        }
                    } else {
                        const uint64_t a_desc = make_sw128_kmajor_desc(a_addr);
                        const uint64_t b_desc = make_sw128_kmajor_desc(p.b_resident ? b_res_addr + kc * b_stage_bytes : a_addr + kAStageBytes);
                        // K tail (K = 144 / 288 / 168 in stages 1-2 and the patch embedding): the zero-filled k-steps of
                        // the last chunk are not issued at all (25 % / 10 % of those GEMMs' tensor work)
                        const int ksteps = kc == p.num_k_chunks - 1 ? p.last_chunk_ksteps : kBlockK / kUmmaK;
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            if (k >= ksteps) break;
                            // advance 32 B (= 16 elements) along K inside the 128 B swizzle atom
                            const uint64_t koff = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
                            mma(d_tmem, a_desc + koff, b_desc + koff, (kc | k) != 0);
                        }
                    }
                    release(empty_bar(stage));
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                commit(tmem_full_bar(acc));
                SPG_STAMP(trace_i, 5);
#ifdef SPG_TRACE
                if (blockIdx.x == 0 && trace_i < 64) g_trace[trace_i * 12 + 10] = starved;
#endif
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        constexpr int kParts = kEpiWarps / 4;  // warps sharing one TMEM lane quarter split the tile's columns
        constexpr int kEpiThreads = 32 * kEpiWarps;
        const int ewarp = warp - 2;      // 0..kEpiWarps-1
        const int quarter = warp & 3;    // TMEM lane quarter this warp may access
        const int part = ewarp >> 2;     // which slice of the chunks
        const int etid = threadIdx.x - 64;
        const int row_in_tile = quarter * 32 + lane;
        float* bias_s = reinterpret_cast<float*>(smem_raw + (scratch_addr - raw_addr));
        float* cw_s = bias_s + 2 * 256;   // [2][256] column sums of W' (kLn == 1) / LayerNorm gamma (kLn == 3)
        float* lnb_s = cw_s + 2 * 256;    // [2][256] LayerNorm beta (kLn == 3)
        float* headw_s = bias_s + 2 * 256 + (p.ln_mode ? 4 * 256 : 0);  // fused head (see epi_scratch_bytes)
        float* headp_s = headw_s + 256;
        const uint32_t my_staging = staging_addr + ewarp * kResSlots * p.buf_bytes;
        // kLn == 2: ring of 1 KB buffers for the 16-bit centred copy, slot-locked to the main ring
        const uint32_t my_ln_staging = staging_addr + kEpiWarps * kResSlots * p.buf_bytes + ewarp * kResSlots * kLnBufBytes;
        uint8_t* my_ln_staging_ptr = smem_raw + (my_ln_staging - raw_addr);
        // kLn == 3: own ring of kLnSlots 1 KB buffers for the normalised rows, then the statistics exchange area
        const uint32_t ln3_base = staging_addr + kEpiWarps * kResSlots * p.buf_bytes;
        const uint32_t my_ln3_staging = ln3_base + ewarp * kLnSlots * kLnBufBytes;
        uint8_t* my_ln3_staging_ptr = smem_raw + (my_ln3_staging - raw_addr);
        const uint32_t ln_x_addr = ln3_base + kEpiWarps * kLnSlots * kLnBufBytes;  // float2 [2][kLnMaxParts][128]
        const float2* ln_x_ptr = reinterpret_cast<const float2*>(smem_raw + (ln_x_addr - raw_addr));
        int ln_buf = 0, ln_slot = 0;
        uint32_t ln_phase = 0;  // bit b = parity of ln_bar(b)
        uint8_t* my_staging_ptr = smem_raw + (my_staging - raw_addr);
        // Staging buffer = 32 rows x (group x 16) columns of the output type, rows of 32 / 64 / 128 bytes in
        // the matching TMA swizzle (32B / 64B / 128B): 16-byte piece j of row r sits at j ^ ((r >> shift) & mask),
        // which makes the row-per-thread accesses below bank-conflict free.
        const uint32_t row_off = lane * p.row_bytes;
        const uint32_t row_xor = (lane >> p.piece_shift) & p.piece_mask;
        const uint32_t pieces_per_chunk = out_f32 ? 4u : 2u;
        const bool pair = p.group == 2;  // two 16-column chunks share one staging buffer / TMA op
        if (has_head && etid < p.block_n) headw_s[etid] = __ldg(p.head_w + etid);  // num_n_tiles == 1

        int acc = 0;
        uint32_t acc_phase = 0;
        int buf = 0;
        // staging ring state: slot of the group being filled, per-slot residual-barrier parity (bit i = slot i),
        // and the slot / parity bookkeeping of the residual loads running two groups ahead
        int slot = 0;
        uint32_t res_parity = 0;
        // the accumulator stage is handed back on the LEADER's tmem_empty barrier (the leader issues the MMAs)
        const uint32_t tmem_empty_remote0 = kPair ? mapa_shared(tmem_empty_bar(0), crank & ~1u) : 0u;
        [[maybe_unused]] int trace_i = 0;
        for (int tile = unit; tile < total_tiles; tile += nunits, ++trace_i) {
            if (ewarp == 0 && lane == 0) SPG_STAMP(trace_i, 6);
            const int t_idx = p.reverse ? total_tiles - 1 - tile : tile;
            const int m_unit = t_idx / p.num_n_tiles;
            const int n_blk = t_idx - m_unit * p.num_n_tiles;
            const int m_blk = kPair ? 2 * m_unit + static_cast<int>(rank) : m_unit;
            const int n0 = n_blk * p.block_n;
            // ragged last n-tile (N not a multiple of block_n): only the valid 16-column chunks are processed; the
            // MMA still runs block_n wide on TMA zero-filled weight rows
            const int chunks = min(p.block_n, p.N - n0) >> 4;
            // balanced split in units of one staging group (a group is never shared between warps); the first part
            // takes the remainder of an odd chunk count
            const int units = (chunks + p.group - 1) / p.group;
            const int c_begin = part == 0 ? 0 : min(chunks, p.group * ((units * part + kParts - 1) / kParts));
            const int c_end = part == kParts - 1 ? chunks : min(chunks, p.group * ((units * (part + 1) + kParts - 1) / kParts));
            const int n_my = c_end - c_begin;
            const int row0 = m_blk * kBlockM + quarter * 32;  // first output row of this warp
            const int res_row0 = p.res_rows > 0 ? row0 % p.res_rows : row0;
            // kUp2: position of this warp's 32 low-resolution pixels, and the correction row of a border-column pixel
            int up_img_row = 0, up_x = 0;
            const float* corr_row = nullptr;
            // conv modes: image / row / column of this warp's 32 pixels (they lie in one image row: tile_w >= 32)
            int cv_img = 0, cv_y = 0, cv_x = 0;
            // (exact tilings address their outputs by the flat pixel index `row`; ragged convolutions run the generic
            // instance, so that the specialised ones carry none of this)
            constexpr bool kGenericInst = kAct < 0;
            if (kUp2 || (kGenericInst && p.ragged)) {
                cv_img = m_blk / p.tiles_img;
                const int rem = m_blk - cv_img * p.tiles_img;
                const int ty = rem / p.tiles_x;
                const int pix0 = quarter * 32;
                cv_y = ty * p.tile_h + pix0 / p.tile_w;
                cv_x = (rem - ty * p.tiles_x) * p.tile_w + pix0 % p.tile_w;
            }
            if (kUp2) {
                up_x = cv_x;
                up_img_row = cv_img * p.H + cv_y;
                const int x = up_x + lane;
                if (x == 0) corr_row = p.corr + static_cast<size_t>(up_img_row) * p.N + n0;
                else if (x == p.W - 1) corr_row = p.corr + (static_cast<size_t>(p.bh) + up_img_row) * p.N + n0;
            }
            float* bs = bias_s + buf * 256;
            if (etid < p.block_n) bs[etid] = (p.bias != nullptr && n0 + etid < p.N) ? __ldg(p.bias + n0 + etid) : 0.f;
            float* cws = cw_s + buf * 256;
            if (kLn == 1 && etid < p.block_n) cws[etid] = n0 + etid < p.N ? __ldg(p.ln_fold_cw + n0 + etid) : 0.f;
            float* lnbs = lnb_s + buf * 256;
            if (kLn == 3 && etid < p.block_n) {
                cws[etid] = __ldg(p.ln_gamma + n0 + etid);  // N % block_n == 0 in this mode
                lnbs[etid] = __ldg(p.ln_beta + n0 + etid);
            }
            float ln3_shift = 0.f;
            // LayerNorm folding: this thread's row statistics (consumer) or centre (producer), from the row records
            const int ln_row = row0 + lane;
            float ln_rs = 1.f, ln_rm = 0.f, ln_c = 0.f, ln_s1 = 0.f, ln_s2 = 0.f;
            if (kLn == 1 && ln_row < p.M) {
                const float* rec = p.ln_fold_rec + static_cast<size_t>(ln_row) * kLnRec;
                const int parts = static_cast<int>(__ldg(rec + 1));
                float s1 = 0.f, s2 = 0.f;
                for (int i = 0; i < parts; ++i) {
                    const float2 pr = __ldg(reinterpret_cast<const float2*>(rec + 2) + i);
                    s1 += pr.x;
                    s2 += pr.y;
                }
                const float m = s1 * p.ln_inv_cols;
                const float var = fmaxf(s2 * p.ln_inv_cols - m * m, 0.f);
                ln_rs = rsqrtf(var + p.ln_eps);
                ln_rm = ln_rs * m;
            }
            if (kLn == 2 && p.ln_prev_rec != nullptr && ln_row < p.M) {
                const float* rec = p.ln_prev_rec + static_cast<size_t>(ln_row % (p.res_rows > 0 ? p.res_rows : p.M)) * kLnRec;
                const int parts = static_cast<int>(__ldg(rec + 1));
                float s1 = 0.f;
                for (int i = 0; i < parts; ++i) s1 += __ldg(rec + 2 + 2 * i);
                ln_c = __ldg(rec) + s1 * p.ln_inv_cols;  // mean of the residual input row
            }

            auto issue_res_load = [&](int sl, int c) {  // lane 0 only; c = first chunk of the group
                const uint32_t bar = res_bar(ewarp, sl);
                mbar_arrive_expect_tx(bar, p.buf_bytes);
                tma_load_2d(my_staging + sl * p.buf_bytes, &tmap_res, bar, n0 + c * 16, res_row0);
            };
            if (has_res && lane == 0) {
                // the kResSlots - 1 slots after the last published one: every store but the most recent has read them
                tma_store_wait_read<1>();
                int sl = slot;
#pragma unroll
                for (int i = 0; i < kResSlots - 1; ++i) {
                    if (i * p.group < n_my) issue_res_load(sl, c_begin + i * p.group);
                    sl = sl == kResSlots - 1 ? 0 : sl + 1;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // bias staged (double-buffered across tiles)

            if (ewarp == 0 && lane == 0) SPG_STAMP(trace_i, 7);
            mbar_wait(tmem_full_bar(acc), acc_phase);
            tc_fence_after();
            if (ewarp == 0 && lane == 0) SPG_STAMP(trace_i, 8);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kAccStageCols;
            const int row = row0 + lane;
            const bool row_ok = row < p.M;
            float head_acc = 0.f;

            // first / last: position of this chunk inside its staging group
            auto process = [&](const uint32_t (&raw)[16], int c, bool first, bool last) {
                const int col = c * 16;  // column inside the tile
                uint8_t* stg = my_staging_ptr + slot * p.buf_bytes + row_off;
                const uint32_t piece0 = first ? 0u : pieces_per_chunk;
                if (first) {
                    if (has_res) {
                        mbar_wait(res_bar(ewarp, slot), (res_parity >> slot) & 1u);
                        res_parity ^= 1u << slot;
                    } else if (has_out) {
                        // the slot was last read by the store issued kResSlots groups ago
                        if (lane == 0) tma_store_wait_read<kResSlots - 1>();
                        __syncwarp();
                    }
                }
                float v[16];
                const float4* b4 = reinterpret_cast<const float4*>(bs + col);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b = b4[i];
                    if (kLn == 1) {  // rstd * (acc - m * colsum(W')) + (b + W beta)
                        const float4 cw = reinterpret_cast<const float4*>(cws + col)[i];
                        v[4 * i + 0] = fmaf(__uint_as_float(raw[4 * i + 0]), ln_rs, fmaf(-ln_rm, cw.x, b.x));
                        v[4 * i + 1] = fmaf(__uint_as_float(raw[4 * i + 1]), ln_rs, fmaf(-ln_rm, cw.y, b.y));
                        v[4 * i + 2] = fmaf(__uint_as_float(raw[4 * i + 2]), ln_rs, fmaf(-ln_rm, cw.z, b.z));
                        v[4 * i + 3] = fmaf(__uint_as_float(raw[4 * i + 3]), ln_rs, fmaf(-ln_rm, cw.w, b.w));
                        continue;
                    }
                    // packed adds (FADD2: two IEEE adds per issue slot, same results)
                    f2_split(f2_add(f2_make(__uint_as_float(raw[4 * i + 0]), __uint_as_float(raw[4 * i + 1])), f2_make(b.x, b.y)),
                             v[4 * i + 0], v[4 * i + 1]);
                    f2_split(f2_add(f2_make(__uint_as_float(raw[4 * i + 2]), __uint_as_float(raw[4 * i + 3])), f2_make(b.z, b.w)),
                             v[4 * i + 2], v[4 * i + 3]);
                }
                if (kUp2 && corr_row != nullptr) {  // first / last image column: bilinear clamp + conv zero padding
                    const float4* c4 = reinterpret_cast<const float4*>(corr_row + col);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 cc = __ldg(c4 + i);
                        v[4 * i + 0] += cc.x;
                        v[4 * i + 1] += cc.y;
                        v[4 * i + 2] += cc.z;
                        v[4 * i + 3] += cc.w;
                    }
                }
                if (act == SPG_ACT_RELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                } else if (act == SPG_ACT_GELU) {
#pragma unroll
                    for (int i = 0; i < 16; i += 2) gelu_erf2(v[i], v[i + 1]);
                }
                if (has_res) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 r = *reinterpret_cast<const float4*>(stg + (((piece0 + i) ^ row_xor) << 4));
                        f2_split(f2_add(f2_make(v[4 * i + 0], v[4 * i + 1]), f2_make(r.x, r.y)), v[4 * i + 0], v[4 * i + 1]);
                        f2_split(f2_add(f2_make(v[4 * i + 2], v[4 * i + 3]), f2_make(r.z, r.w)), v[4 * i + 2], v[4 * i + 3]);
                    }
                }
                if (kLn == 3) {  // v = the row's new residual-stream values: statistics now, normalisation in pass 2
                    if (c == c_begin) ln3_shift = v[0];
                    uint32_t wb[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        ln_accumulate(v[i], ln3_shift, ln_s1, ln_s2);  // shifted sums: no cancellation in M2
                        wb[i] = __float_as_uint(v[i]);
                    }
                    tmem_st16(taddr + c * 16, wb);  // park v over the consumed accumulator columns
                }
                if (kLn == 2) {  // centred 16-bit copy for the consumer GEMMs + this thread's partial row statistics
                    float vc[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        vc[i] = v[i] - ln_c;
                        ln_s1 += vc[i];
                        ln_s2 = fmaf(vc[i], vc[i], ln_s2);
                    }
                    uint8_t* lst = my_ln_staging_ptr + slot * kLnBufBytes + lane * 32;
                    const uint32_t lx = (lane >> 2) & 1u;  // 32-byte swizzle: the two 16-byte pieces swap on rows 4..7 of 8
                    *reinterpret_cast<uint4*>(lst + ((0u ^ lx) << 4)) =
                        make_uint4(pack2(vc[0], vc[1]), pack2(vc[2], vc[3]), pack2(vc[4], vc[5]), pack2(vc[6], vc[7]));
                    *reinterpret_cast<uint4*>(lst + ((1u ^ lx) << 4)) =
                        make_uint4(pack2(vc[8], vc[9]), pack2(vc[10], vc[11]), pack2(vc[12], vc[13]), pack2(vc[14], vc[15]));
                }
                if (has_head) {
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
                if (has_out) {
                    if (out_f32) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            *reinterpret_cast<float4*>(stg + (((piece0 + i) ^ row_xor) << 4)) =
                                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    } else {
                        *reinterpret_cast<uint4*>(stg + (((piece0 + 0u) ^ row_xor) << 4)) =
                            make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                        *reinterpret_cast<uint4*>(stg + (((piece0 + 1u) ^ row_xor) << 4)) =
                            make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
                    }
                }
                if (last) {  // group complete: publish it
                    const int c_first = first ? c : c - 1;
                    if (has_out) {
                        fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
                        __syncwarp();
                        if (lane == 0) {
                            if (kUp2) {
                                const int n = n0 + c_first * 16;
                                const int phase = n / p.cout;  // (row phase, column phase) = (phase >> 1, phase & 1)
                                tma_store_5d(&tmap_out, my_staging + slot * p.buf_bytes, n - phase * p.cout, phase & 1, up_x,
                                             phase >> 1, up_img_row);
                            } else if (kAct < 0 && p.ragged) {  // [C, W, H, B] map: columns x >= W of the last tile column are clipped
                                tma_store_4d(&tmap_out, my_staging + slot * p.buf_bytes, n0 + c_first * 16, cv_x, cv_y, cv_img);
                            } else {
                                tma_store_2d(&tmap_out, my_staging + slot * p.buf_bytes, n0 + c_first * 16, row0);
                            }
                            if (kLn == 2) tma_store_2d(&tmap_ln, my_ln_staging + slot * kLnBufBytes, n0 + c_first * 16, row0);
                            tma_store_commit();
                        }
                    }
                    if (has_res && lane == 0 && c + 1 + (kResSlots - 2) * p.group < c_end) {
                        // the group kResSlots - 1 ahead reuses the slot of the previous group: its store must be done reading
                        tma_store_wait_read<1>();
                        issue_res_load(slot == 0 ? kResSlots - 1 : slot - 1, c + 1 + (kResSlots - 2) * p.group);
                    }
                    slot = slot == kResSlots - 1 ? 0 : slot + 1;
                }
            };

            uint32_t ra[16], rb[16];
            if (n_my > 0) tmem_ld16(taddr + c_begin * 16, ra);
            for (int c = c_begin; c < c_end; c += 2) {
                tmem_ld_wait();
                if (c + 1 < c_end) tmem_ld16(taddr + (c + 1) * 16, rb);
                process(ra, c, true, !pair);
                if (c + 1 < c_end) {
                    tmem_ld_wait();
                    if (c + 2 < c_end) tmem_ld16(taddr + (c + 2) * 16, ra);
                    process(rb, c + 1, !pair, true);
                }
            }
            if (kLn == 3) {
                // ---- exchange {mean, M2} of this thread's column slice with every CTA that holds an n-tile of the row
                // block (cluster peers; in pair mode the CTAs with the same pair rank), combine in fixed (n-tile, slice)
                // order -> statistics that do not depend on M, on the grid or on which CTA computed what
                const int chunks_all = p.block_n >> 4;
                const int units_all = (chunks_all + p.group - 1) / p.group;
                auto slice_cols = [&](int pt) {
                    const int b = pt == 0 ? 0 : min(chunks_all, p.group * ((units_all * pt + kParts - 1) / kParts));
                    const int e = pt == kParts - 1 ? chunks_all : min(chunks_all, p.group * ((units_all * (pt + 1) + kParts - 1) / kParts));
                    return static_cast<float>((e - b) * 16);
                };
                const float2 mine = n_my > 0 ? ln_slice_stats(ln3_shift, ln_s1, ln_s2, static_cast<float>(n_my * 16)) : make_float2(0.f, 0.f);
                const float mean_i = mine.x, m2_i = mine.y;
                const uint32_t my_x = ln_x_addr + static_cast<uint32_t>(((ln_buf * kLnMaxParts + n_blk * kParts + part) * kBlockM + row_in_tile) * 8);
                for (int j = 0; j < p.num_n_tiles; ++j) {
                    const uint32_t peer = kPair ? static_cast<uint32_t>(2 * j) + rank : static_cast<uint32_t>(j);
                    st_cluster_f32x2(mapa_shared(my_x, peer), mean_i, m2_i);
                }
                // one release-arrive per warp and peer (lane 0, after the warp barrier has ordered the other lanes' remote
                // stores before it): a release at cluster scope per THREAD showed up as 2.8 membar stalls per issue in ncu
                __syncwarp();
                if (lane == 0)
                    for (int j = 0; j < p.num_n_tiles; ++j) {
                        const uint32_t peer = kPair ? static_cast<uint32_t>(2 * j) + rank : static_cast<uint32_t>(j);
                        mbar_arrive_cluster(mapa_shared(ln_bar(ln_buf), peer));
                    }
                tmem_st_wait();
                mbar_wait_cluster(ln_bar(ln_buf), (ln_phase >> ln_buf) & 1u);
                ln_phase ^= 1u << ln_buf;
                const float2* px = ln_x_ptr + (ln_buf * kLnMaxParts) * kBlockM + row_in_tile;
                const float2 rm = ln_merge(p.num_n_tiles * kParts, [&](int q) { return px[q * kBlockM]; },
                                           [&](int q) { return slice_cols(q % kParts); }, p.ln_inv_cols, p.ln_eps);
                ln_buf ^= 1;
                // ---- pass 2: y = (v - mean) * rstd * gamma + beta, 16 bit, staged per 16-column chunk and TMA-stored
                uint32_t ya[16], yb[16];
                auto emit = [&](const uint32_t (&raw)[16], int c) {
                    const int col = c * 16;
                    if (lane == 0) tma_store_wait_read<kLnSlots - 1>();  // the store that last read this slot is done
                    __syncwarp();
                    float y[16];
                    // ln_normalise on packed pairs (two FFMA2 per two elements; the same two roundings per element)
                    const f32x2 rstd2 = f2_make(rm.x, rm.x), shift2 = f2_make(rm.y, rm.y);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 g = reinterpret_cast<const float4*>(cws + col)[i];
                        const float4 b = reinterpret_cast<const float4*>(lnbs + col)[i];
                        const f32x2 v01 = f2_make(__uint_as_float(raw[4 * i + 0]), __uint_as_float(raw[4 * i + 1]));
                        const f32x2 v23 = f2_make(__uint_as_float(raw[4 * i + 2]), __uint_as_float(raw[4 * i + 3]));
                        f2_split(f2_fma(f2_fma(v01, rstd2, shift2), f2_make(g.x, g.y), f2_make(b.x, b.y)), y[4 * i + 0], y[4 * i + 1]);
                        f2_split(f2_fma(f2_fma(v23, rstd2, shift2), f2_make(g.z, g.w), f2_make(b.z, b.w)), y[4 * i + 2], y[4 * i + 3]);
                    }
                    uint8_t* lst = my_ln3_staging_ptr + ln_slot * kLnBufBytes + lane * 32;
                    const uint32_t lx = (lane >> 2) & 1u;  // 32-byte swizzle: the two 16-byte pieces swap on rows 4..7 of 8
                    *reinterpret_cast<uint4*>(lst + ((0u ^ lx) << 4)) =
                        make_uint4(pack2(y[0], y[1]), pack2(y[2], y[3]), pack2(y[4], y[5]), pack2(y[6], y[7]));
                    *reinterpret_cast<uint4*>(lst + ((1u ^ lx) << 4)) =
                        make_uint4(pack2(y[8], y[9]), pack2(y[10], y[11]), pack2(y[12], y[13]), pack2(y[14], y[15]));
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmap_ln, my_ln3_staging + ln_slot * kLnBufBytes, n0 + col, row0);
                        tma_store_commit();
                    }
                    ln_slot = ln_slot == kLnSlots - 1 ? 0 : ln_slot + 1;
                };
                if (n_my > 0) tmem_ld16(taddr + c_begin * 16, ya);
                for (int c = c_begin; c < c_end; c += 2) {
                    tmem_ld_wait();
                    if (c + 1 < c_end) tmem_ld16(taddr + (c + 1) * 16, yb);
                    emit(ya, c);
                    if (c + 1 < c_end) {
                        tmem_ld_wait();
                        if (c + 2 < c_end) tmem_ld16(taddr + (c + 2) * 16, ya);
                        emit(yb, c + 1);
                    }
                }
            }
            if (kLn == 2 && ln_row < p.M) {  // row record: fixed slot per (n-tile, column slice) -> deterministic statistics
                float* rec = p.ln_emit_rec + static_cast<size_t>(ln_row) * kLnRec;
                reinterpret_cast<float2*>(rec + 2)[n_blk * kParts + part] = make_float2(ln_s1, ln_s2);
                if (n_blk == 0 && part == 0) *reinterpret_cast<float2*>(rec) = make_float2(ln_c, static_cast<float>(p.num_n_tiles * kParts));
            }
            // accumulator fully read: hand the TMEM stage back to the MMA warp
            if (ewarp == 0 && lane == 0) SPG_STAMP(trace_i, 9);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (kPair) mbar_arrive_cluster(tmem_empty_remote0 + 8u * acc);
                else mbar_arrive(tmem_empty_bar(acc));
            }
            if (has_head) {
                // combine the column slices of the fused N->1 head through smem
                if (part > 0) headp_s[(part - 1) * 128 + row_in_tile] = head_acc;
                asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads) : "memory");
                if (part == 0 && row_ok) {
#pragma unroll
                    for (int k = 1; k < kParts; ++k) head_acc += headp_s[(k - 1) * 128 + row_in_tile];
                    if (!(kAct < 0 && p.ragged)) p.head_out[row] = head_acc + p.head_b;
                    else if (cv_x + lane < p.W)  // head_out is the compact [B, H, W] map
                        p.head_out[(static_cast<size_t>(cv_img) * p.H + cv_y) * p.W + cv_x + lane] = head_acc + p.head_b;
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
            buf ^= 1;
        }
        if (lane == 0) tma_store_wait<0>();  // staging smem must outlive the bulk stores
    }

    tc_fence_before();
    if (kPair) {
        cluster_sync_all();  // the peer may still be reading this CTA's smem / signalling its barriers
        if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
    } else {
        if (kLn == 3) cluster_sync_all();  // peers may still be writing statistics into this CTA's shared memory
        else __syncthreads();
        if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
    }
}

// Tile width.  256-wide tiles run the tensor pipe ~15 % faster than 192-wide ones (measured 1347 vs 1150 TFLOP/s at
// long K), so when N is not a multiple of 256 but the padding of a ragged last tile costs <= 12 % (N = 1728, 3456,
// 1152: 3.7 / 3.7 / 11 %) the launch uses 256-wide tiles with a ragged tail; otherwise the largest divisor of N.
// max_n_tiles bounds the number of n-tiles (fused head: 1; LayerNorm producer: 7 = the partial slots of a row record)
int pick_block_n(int N, int num_m_tiles = 1 << 20, int max_n_tiles = 1 << 20) {
    static const int ragged_env = [] { const char* e = getenv("SPG_GEMM_RAGGED"); return e ? atoi(e) : 1; }();
    static const int small_env0 = [] { const char* e = getenv("SPG_GEMM_SMALLM"); return e ? atoi(e) : 1; }();
    const bool small = small_env0 && num_m_tiles * ((N + 255) / 256) * 2 <= sm_count();
    if (ragged_env && !small && N > 256 && N % 256 != 0 && N % 64 == 0) {
        const int padded = (N + 255) / 256 * 256;
        if ((padded - N) * 100 <= 12 * N) return 256;
    }
    static const int max_bn_env = [] { const char* e = getenv("SPG_GEMM_MAX_BN"); return e ? atoi(e) : 256; }();  // tuning
    int best = 0;
    for (int bn = max_bn_env; bn >= 16; bn -= 16)
        if (N % bn == 0) {
            best = bn;
            break;
        }
    // Small problems (batch 1: M = 1024 / 256 rows in stages 3 / 4) fill less than half of the SMs with the widest
    // tile, and a tile's time is then set by the TMA round trips of its k-loop rather than by its width: take the
    // narrowest tile (>= 64 columns) that still fits one wave, i.e. the most CTAs working in parallel.
    static const int small_env = [] { const char* e = getenv("SPG_GEMM_SMALLM"); return e ? atoi(e) : 1; }();
    if (small_env && best > 64 && num_m_tiles * ((N + best - 1) / best) * 2 <= sm_count()) {
        for (int bn = 64; bn < best; bn += 16)
            if (N % bn == 0 && num_m_tiles * (N / bn) <= sm_count() && N / bn <= max_n_tiles) return bn;
    }
    return best;
}

// CTA pairs (cta_group::2).  Must be decided before the weight tensor map is encoded: in pair mode each CTA's
// TMA box is half of the block_n weight rows.  Measured on B200 (tests/cuda/test_gemm.cu --perf, B=64 shapes):
// pairs gain 3-9 % on wide tiles with a long reduction (N=2304 K=2304: 1236 -> 1347 TFLOP/s, conv N=128: 1147 ->
// 1237) and lose 3-18 % on short reductions and on N=64 tiles, where the per-tile handshake across the two CTAs is
// not amortised; so they are used for block_n >= 128 with >= 16 k-steps and enough tiles to fill every SM pair.
void decide_pair(GemmArgs& a) {
    static const int pair_env = [] { const char* e = getenv("SPG_GEMM_PAIR"); return e ? atoi(e) : 1; }();
    const int ksteps = a.num_k_chunks * (a.halo ? 3 : 1);
    const bool fits = a.block_n % 32 == 0 && a.num_m_tiles * a.num_n_tiles >= 2 * sm_count();
    // 256-wide tiles without a GELU epilogue also gain at K = 576 (QKV: 999 -> 1079 TFLOP/s); with GELU the
    // epilogue is the longer phase and the cross-CTA handshake only adds to it (935 -> 926)
    const bool wide_short = a.block_n == 256 && ksteps >= 8 && a.act != SPG_ACT_GELU && !a.has_res;
    a.pair = (pair_env == 2 && fits) || (pair_env == 1 && fits && ((a.block_n >= 128 && ksteps >= 16) || wide_short)) ? 1 : 0;
}

struct EpiMaps {
    CUtensorMap out, res, ln;
};

int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const EpiMaps& em, GemmArgs& a, const LaunchCtx& ctx) {
    const int total_1cta = a.num_m_tiles * a.num_n_tiles;
    const bool pair = a.pair != 0;
    a.reverse = ctx.reverse ? 1 : 0;
    const int bn_cta = pair ? a.block_n / 2 : a.block_n;  // weight rows staged per CTA
    const int staging_all = (a.has_out ? a.epi_warps * kResSlots * a.buf_bytes : 0) +
                            (a.ln_mode == 2 ? a.epi_warps * kResSlots * kLnBufBytes : 0) +
                            (a.ln_mode == 3 ? a.epi_warps * kLnSlots * kLnBufBytes + kLnXBytes : 0);
    a.scratch_bytes = epi_scratch_bytes(a.ln_mode, a.head_w != nullptr);
    // 1024: the dynamic shared memory base is only 16 B aligned by contract; everything after it is a 1 KB multiple
    const int fixed_base = 1024 + kBarBytes + a.scratch_bytes + ((bn_cta * 128) % 1024 ? 1024 : 0);
    if (a.b_resident && !a.halo &&
        kSmemBudget - (fixed_base + staging_all) - a.num_k_chunks * bn_cta * 128 < 4 * kAStageBytes)
        a.b_resident = 0;  // the resident weight tile would leave fewer than four A stages
    a.b_res_tiles = a.b_resident ? (a.halo ? 9 * a.cin_chunks : a.num_k_chunks) : 0;
    // off by default: built and correct (tests/test_gpu_modes.py runs the LayerNorm-producer tests with it on), but the
    // step is power-capped, not L2-feed bound -- same-box A/B 1006.4 (on) vs 1006.5 img/s (off), DESIGN.md 4.1b2
    static const int amcast_env = [] { const char* e = getenv("SPG_GEMM_AMCAST"); return e ? atoi(e) : 0; }();
    a.a_mcast = (amcast_env && a.ln_mode == 3 && a.num_n_tiles > 1 && !a.b_resident && !a.conv) ? 1 : 0;
    const int b_res_bytes = a.b_res_tiles * bn_cta * 128;
    const int stage_bytes = a.b_resident ? (a.halo ? kHaloABytes : kAStageBytes)
                                         : (a.halo ? kHaloABytes + 3 * bn_cta * 128 : kAStageBytes + bn_cta * 128);
    const int staging = (a.has_out ? a.epi_warps * kResSlots * a.buf_bytes : 0) +
                        (a.ln_mode == 2 ? a.epi_warps * kResSlots * kLnBufBytes : 0) +
                        (a.ln_mode == 3 ? a.epi_warps * kLnSlots * kLnBufBytes + kLnXBytes : 0);
    const int fixed = fixed_base + staging + b_res_bytes;
    int stages = (kSmemBudget - fixed) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (!a.b_resident && stages > a.num_k_chunks + 1) stages = a.num_k_chunks + 1;  // no point in a deeper ring
    if (stages < 2) stages = 2;
    a.stages = stages;
    const int smem = stages * stage_bytes + fixed;
    const int sms = sm_count();
    // ln_mode 3: the CTAs / pairs holding the n-tiles of one row block form the cluster
    const int cluster = (pair ? 2 : 1) * (a.ln_mode == 3 ? a.num_n_tiles : 1);
    int grid;
    if (pair) {
        const int units = ((a.num_m_tiles + 1) / 2) * a.num_n_tiles;
        grid = 2 * (units < sms / 2 ? units : sms / 2);
    } else {
        grid = total_1cta < sms ? total_1cta : sms;
    }
    grid -= grid % cluster;  // whole clusters only (total tiles are a multiple of the n-tiles per row block)
    if (a.b_resident) {
        grid -= grid % a.num_n_tiles;  // a CTA keeps one n-block for all its tiles
        if (grid == 0) return fail(SPG_ERR_INVALID, "resident-weight mode needs at least one CTA per n-tile");
    }
    const bool head = a.head_w != nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(64 + 32 * a.epi_warps);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx.stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see common.h "Programmatic dependent launch"
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx.pdl ? 2 : 1;
    // the combinations SPEGNet launches get a compile-time specialised epilogue; anything else runs the generic one
#define SPG_LAUNCH_ONE(ACT, F32, RES, HEAD, OUT, PAIR, EW)                                                         \
    do {                                                                                                           \
        auto kern = gemm_tcgen05_kernel<ACT, F32, RES, HEAD, OUT, PAIR, EW>;                                       \
        static PerDeviceOnce attr_set;                                                                              \
        if (attr_set.needed()) {                                                                                           \
            SPG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));  \
            attr_set.done();                                                                                       \
        }                                                                                                          \
        SPG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, em.out, em.res, em.ln, a));                                 \
    } while (0)
#define SPG_LAUNCH_ONE_UP2(PAIR)                                                                                   \
    do {                                                                                                           \
        auto kern = gemm_tcgen05_kernel<SPG_ACT_RELU, 0, 0, 0, 1, PAIR, kEpiWarpsDefault, 1>;                       \
        static PerDeviceOnce attr_set;                                                                              \
        if (attr_set.needed()) {                                                                                           \
            SPG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));  \
            attr_set.done();                                                                                       \
        }                                                                                                          \
        SPG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, em.out, em.res, em.ln, a));                                 \
    } while (0)
#define SPG_LAUNCH_ONE_LN(ACT, F32, RES, PAIR, LN)                                                                 \
    do {                                                                                                           \
        auto kern = gemm_tcgen05_kernel<ACT, F32, RES, 0, 1, PAIR, kEpiWarpsDefault, 0, LN>;                        \
        static PerDeviceOnce attr_set;                                                                              \
        if (attr_set.needed()) {                                                                                           \
            SPG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));  \
            attr_set.done();                                                                                       \
        }                                                                                                          \
        SPG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, em.out, em.res, em.ln, a));                          \
    } while (0)
#define SPG_LAUNCH_LN(ACT, F32, RES, LN)                       \
    do {                                                       \
        if (pair) SPG_LAUNCH_ONE_LN(ACT, F32, RES, 1, LN);     \
        else SPG_LAUNCH_ONE_LN(ACT, F32, RES, 0, LN);          \
    } while (0)
#define SPG_LAUNCH(ACT, F32, RES, HEAD, OUT, EW)                     \
    do {                                                             \
        if (pair) SPG_LAUNCH_ONE(ACT, F32, RES, HEAD, OUT, 1, EW);   \
        else SPG_LAUNCH_ONE(ACT, F32, RES, HEAD, OUT, 0, EW);        \
    } while (0)
#define SPG_LAUNCH_EW(ACT, F32, RES, HEAD, OUT) SPG_LAUNCH(ACT, F32, RES, HEAD, OUT, kEpiWarpsDefault)
    if (a.ln_mode == 1) {  // LayerNorm consumer: qkv (no act), fc1 (GELU), dim-change proj (fp32 out)
        if (a.act == SPG_ACT_GELU) SPG_LAUNCH_LN(SPG_ACT_GELU, 0, 0, 1);
        else if (a.out_f32) SPG_LAUNCH_LN(SPG_ACT_NONE, 1, 0, 1);
        else SPG_LAUNCH_LN(SPG_ACT_NONE, 0, 0, 1);
    } else if (a.ln_mode == 2) {  // LayerNorm producer: fp32 residual GEMM that also emits the centred copy + records
        SPG_LAUNCH_LN(SPG_ACT_NONE, 1, 1, 2);
    } else if (a.ln_mode == 3) {  // LayerNorm producer that normalises itself (clusters over the n-tiles of a row block)
        if (cluster > 2) {
            // a persistent grid must be co-resident: cap it at what the GPCs can hold of this cluster shape
            static int max_clusters[2][9][128] = {};
            int dev = 0;
            cudaGetDevice(&dev);
            int& cached = max_clusters[pair ? 1 : 0][cluster][dev & 127];
            if (cached == 0) {
                cfg.gridDim = dim3(static_cast<unsigned>(sms - sms % cluster));
                int n = 0;
                auto k1 = gemm_tcgen05_kernel<SPG_ACT_NONE, 1, 1, 0, 1, 1, kEpiWarpsDefault, 0, 3>;
                auto k0 = gemm_tcgen05_kernel<SPG_ACT_NONE, 1, 1, 0, 1, 0, kEpiWarpsDefault, 0, 3>;
                SPG_CHECK_CUDA(cudaFuncSetAttribute(pair ? k1 : k0, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
                SPG_CHECK_CUDA(cudaFuncSetAttribute(pair ? k1 : k0, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                SPG_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, pair ? k1 : k0, &cfg));
                cached = n > 0 ? n : 1;
            }
            if (grid > cached * cluster) grid = cached * cluster;
            cfg.gridDim = dim3(grid);
        }
        SPG_LAUNCH_LN(SPG_ACT_NONE, 1, 1, 3);
    } else if (a.up2) {
        if (pair) SPG_LAUNCH_ONE_UP2(1);
        else SPG_LAUNCH_ONE_UP2(0);
    } else if (a.ragged) {  // ragged conv tile columns: only the generic instance carries the clipping code
        SPG_LAUNCH(-1, -1, -1, -1, -1, kEpiWarpsDefault);
    } else if (a.act == SPG_ACT_NONE && !a.out_f32 && !a.has_res && !head && a.has_out) SPG_LAUNCH_EW(SPG_ACT_NONE, 0, 0, 0, 1);
    else if (a.act == SPG_ACT_NONE && a.out_f32 && a.has_res && !head && a.has_out) SPG_LAUNCH(SPG_ACT_NONE, 1, 1, 0, 1, kEpiWarpsDefault);
    else if (a.act == SPG_ACT_NONE && a.out_f32 && !a.has_res && !head && a.has_out) SPG_LAUNCH(SPG_ACT_NONE, 1, 0, 0, 1, kEpiWarpsDefault);
    else if (a.act == SPG_ACT_GELU && !a.out_f32 && !a.has_res && !head && a.has_out) SPG_LAUNCH_EW(SPG_ACT_GELU, 0, 0, 0, 1);
    else if (a.act == SPG_ACT_RELU && !a.out_f32 && !a.has_res && !head && a.has_out) SPG_LAUNCH_EW(SPG_ACT_RELU, 0, 0, 0, 1);
    else if (a.act == SPG_ACT_RELU && !a.out_f32 && !a.has_res && head && a.has_out) SPG_LAUNCH_EW(SPG_ACT_RELU, 0, 0, 1, 1);
    else if (a.act == SPG_ACT_RELU && !a.has_res && head && !a.has_out) SPG_LAUNCH_EW(SPG_ACT_RELU, 0, 0, 1, 0);
    else SPG_LAUNCH(-1, -1, -1, -1, -1, kEpiWarpsDefault);
#undef SPG_LAUNCH_EW
#undef SPG_LAUNCH_LN
#undef SPG_LAUNCH_ONE_LN
#undef SPG_LAUNCH_ONE_UP2
#undef SPG_LAUNCH
#undef SPG_LAUNCH_ONE
    g_launches.fetch_add(1, std::memory_order_relaxed);
    SPG_CHECK_LAUNCH();
    return SPG_OK;
}

int fill_epilogue(GemmArgs& a, EpiMaps& em, const spg_epilogue_t* ep, int M, int N) {
    SPG_CHECK_ARG(ep != nullptr, "epilogue descriptor is NULL");
    SPG_CHECK_ARG(ep->out != nullptr || ep->head_out != nullptr, "epilogue has neither out nor head_out");
    SPG_CHECK_ARG(ep->act >= SPG_ACT_NONE && ep->act <= SPG_ACT_GELU, "unknown activation %d", ep->act);
    SPG_CHECK_ARG(ep->out_dtype == SPG_H16 || ep->out_dtype == SPG_F32, "unknown out_dtype %d", ep->out_dtype);
    SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(ep->out) & 15) == 0, "out must be 16-byte aligned");
    SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(ep->bias) & 15) == 0, "bias must be 16-byte aligned");
    SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(ep->residual) & 15) == 0, "residual must be 16-byte aligned");
    SPG_CHECK_ARG(ep->res_rows >= 0 && ep->res_rows % 32 == 0, "res_rows must be a non-negative multiple of 32");
    if (ep->residual != nullptr)
        SPG_CHECK_ARG(ep->out != nullptr && ep->out_dtype == SPG_F32, "a residual needs an fp32 output (the residual stream is fp32)");
    if (ep->head_w != nullptr) {
        SPG_CHECK_ARG(ep->head_out != nullptr, "head_w given without head_out");
        SPG_CHECK_ARG(a.block_n == N, "fused head needs the whole row in one tile (N=%d <= 256, N %% 16 == 0)", N);
        SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(ep->head_w) & 15) == 0, "head_w must be 16-byte aligned");
    }
    a.bias = ep->bias;
    a.act = ep->act;
    a.has_res = ep->residual != nullptr;
    a.res_rows = ep->res_rows;
    a.has_out = ep->out != nullptr;
    a.out_f32 = ep->out_dtype == SPG_F32;
    a.epi_warps = kEpiWarpsDefault;
    // two chunks per staging buffer / TMA op when every warp's slice of the tile is a whole number of pairs
    // (with a residual: 2 KB buffers -> one more mainloop stage)
    a.group = (a.block_n % 64 == 0 && !a.has_res) ? 2 : 1;

    a.row_bytes = a.group * 16 * (a.out_f32 ? 4 : 2);
    a.buf_bytes = 32 * a.row_bytes;
    a.piece_shift = a.row_bytes == 128 ? 0 : (a.row_bytes == 64 ? 1 : 2);
    a.piece_mask = a.row_bytes / 16 - 1;
    a.head_w = ep->head_w;
    a.head_b = ep->head_b;
    a.head_out = ep->head_w != nullptr ? ep->head_out : nullptr;
    // LayerNorm folding
    a.ln_mode = 0;
    if (ep->ln_fold_rec != nullptr) {
        SPG_CHECK_ARG(ep->ln_fold_cw != nullptr && ep->ln_cols > 0, "ln_fold_rec needs ln_fold_cw and ln_cols");
        SPG_CHECK_ARG(ep->head_w == nullptr && ep->residual == nullptr && ep->out != nullptr && ep->act != SPG_ACT_RELU &&
                      !(ep->act == SPG_ACT_GELU && ep->out_dtype == SPG_F32),
                      "LayerNorm folding is built for: no act / GELU with 16-bit out, no act with fp32 out; no residual, no head");
        SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(ep->ln_fold_rec) & 15) == 0 && (reinterpret_cast<uintptr_t>(ep->ln_fold_cw) & 15) == 0,
                      "ln_fold_rec / ln_fold_cw must be 16-byte aligned");
        a.ln_mode = 1;
        a.ln_fold_rec = ep->ln_fold_rec;
        a.ln_fold_cw = ep->ln_fold_cw;
    }
    if (ep->ln_emit_out != nullptr) {
        SPG_CHECK_ARG(a.ln_mode == 0, "a GEMM cannot both consume and produce a folded LayerNorm");
        SPG_CHECK_ARG(ep->ln_emit_rec != nullptr && ep->ln_cols == N, "ln_emit_out needs ln_emit_rec and ln_cols == N");
        SPG_CHECK_ARG(ep->residual != nullptr && ep->out_dtype == SPG_F32 && ep->act == SPG_ACT_NONE && ep->head_w == nullptr,
                      "the LayerNorm producer is the fp32 residual GEMM (no activation, no head)");
        SPG_CHECK_ARG(a.num_n_tiles * 2 <= (kLnRec - 2) / 2, "too many n-tiles (%d) for a LayerNorm row record", a.num_n_tiles);
        SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(ep->ln_emit_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(ep->ln_emit_rec) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(ep->ln_prev_rec) & 15) == 0, "LayerNorm buffers must be 16-byte aligned");
        a.ln_mode = 2;
        a.ln_emit_rec = ep->ln_emit_rec;
        a.ln_prev_rec = ep->ln_prev_rec;
    }
    if (ep->ln_apply_out != nullptr) {
        SPG_CHECK_ARG(a.ln_mode == 0, "ln_apply_out cannot be combined with ln_fold_rec / ln_emit_out");
        SPG_CHECK_ARG(ep->ln_apply_gamma != nullptr && ep->ln_apply_beta != nullptr, "ln_apply_out needs gamma and beta");
        SPG_CHECK_ARG(ep->residual != nullptr && ep->out_dtype == SPG_F32 && ep->act == SPG_ACT_NONE && ep->head_w == nullptr,
                      "the LayerNorm producer is the fp32 residual GEMM (no activation, no head)");
        SPG_CHECK_ARG(N % a.block_n == 0 && a.num_n_tiles * (kEpiWarpsDefault / 4) <= kLnMaxParts,
                      "ln_apply_out: N=%d does not split into <= %d equal n-tiles", N, kLnMaxParts / (kEpiWarpsDefault / 4));
        SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(ep->ln_apply_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(ep->ln_apply_gamma) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(ep->ln_apply_beta) & 15) == 0, "LayerNorm buffers must be 16-byte aligned");
        a.ln_mode = 3;
        a.ln_gamma = ep->ln_apply_gamma;
        a.ln_beta = ep->ln_apply_beta;
    }
    a.ln_inv_cols = ep->ln_cols > 0 ? 1.0f / static_cast<float>(ep->ln_cols) : 0.f;
    if (a.ln_mode == 3) a.ln_inv_cols = 1.0f / static_cast<float>(N);
    a.ln_eps = ep->ln_eps;
    memset(&em, 0, sizeof(em));
    if (a.has_out)
        if (int rc = make_tmap_epilogue(&em.out, ep->out, M, N, a.out_f32, a.group * 16)) return rc;
    if (a.has_res)
        if (int rc = make_tmap_epilogue(&em.res, ep->residual, ep->res_rows > 0 ? ep->res_rows : M, N, 1, a.group * 16)) return rc;
    if (a.ln_mode == 2)
        if (int rc = make_tmap_epilogue(&em.ln, ep->ln_emit_out, M, N, 0, 16)) return rc;
    if (a.ln_mode == 3)
        if (int rc = make_tmap_epilogue(&em.ln, ep->ln_apply_out, M, N, 0, 16)) return rc;
    return SPG_OK;
}

}  // namespace
}  // namespace spg

extern "C" int spg_linear_h16(const void* A, const void* W, int M, int N, int K,
                              const spg_epilogue_t* ep, const spg_launch_t* launch) {
    using namespace spg;
    SPG_CHECK_ARG(A != nullptr && W != nullptr, "A / W is NULL");
    SPG_CHECK_ARG(M > 0 && N > 0 && K > 0, "bad GEMM shape M=%d N=%d K=%d", M, N, K);
    SPG_CHECK_ARG(K % 8 == 0, "K=%d must be a multiple of 8 (16-byte TMA row pitch)", K);
    SPG_CHECK_ARG(N % 16 == 0, "N=%d must be a multiple of 16 (UMMA N granularity at M=128)", N);
    GemmArgs a{};
    EpiMaps em;
    a.M = M;
    a.N = N;
    a.num_m_tiles = (M + kBlockM - 1) / kBlockM;
    a.block_n = pick_block_n(N, a.num_m_tiles, ep != nullptr && ep->head_w != nullptr ? 1 : (ep != nullptr && ep->ln_emit_out != nullptr ? 7 : 1 << 20));
    if (ep != nullptr && ep->ln_apply_out != nullptr) {
        // LayerNorm producer: a FIXED tiling per N (never a function of M: the row statistics are combined per
        // (n-tile, column slice), and results must not depend on the batch size) with at most 3 equal n-tiles
        const LnSlices sl = ln_slices_for(N);
        a.block_n = sl.count > 0 ? 2 * N / sl.count : 0;
        SPG_CHECK_ARG(a.block_n != 0, "ln_apply_out: N=%d has no tiling into <= 3 equal n-tiles of <= 192 columns", N);
    }
    a.num_n_tiles = (N + a.block_n - 1) / a.block_n;
    a.num_k_chunks = (K + kBlockK - 1) / kBlockK;
    a.last_chunk_ksteps = (K - (a.num_k_chunks - 1) * kBlockK + kUmmaK - 1) / kUmmaK;
    a.conv = 0;
    a.H = a.W = 1;
    a.cin_chunks = 1;
    a.tile_w = 1;
    if (int rc = fill_epilogue(a, em, ep, M, N)) return rc;
    decide_pair(a);
    // LayerNorm clusters: 3 n-tiles x CTA pairs = 6-CTA clusters, of which a B200 holds 22 (132 of 148 SMs); without
    // pairs the clusters are 3 CTAs and 46 fit (138 SMs).  SPG_LN_PAIR=0 selects that (experiment knob)
    static const int ln_pair_env = [] { const char* e = getenv("SPG_LN_PAIR"); return e ? atoi(e) : 1; }();
    if (a.ln_mode == 3 && !ln_pair_env) a.pair = 0;
    // resident weights: short K, plain tiling (N a multiple of block_n), enough tiles that the ring depth matters
    static const int bres_env = [] { const char* e = getenv("SPG_GEMM_BRES"); return e ? atoi(e) : 1; }();
    // (K <= 192: 2-3 tiles in flight instead of 1.3; K <= 320 with a 144-wide tile: same ring depth in tiles, but no
    // weight refill traffic on kernels whose MMAs are paced by shared-memory bandwidth -- launch_gemm drops the mode again
    // if fewer than four A stages would fit next to the resident tile)
    a.b_resident = (bres_env && !a.pair && a.num_k_chunks <= (bres_env >= 2 ? 3 : 5) && N % a.block_n == 0 && a.num_n_tiles <= 16 &&
                    a.num_m_tiles * a.num_n_tiles >= 4 * sm_count() && a.ln_mode != 1 && a.ln_mode != 2) ? 1 : 0;
    static const int l2pf_env = [] { const char* e = getenv("SPG_GEMM_L2PF"); return e ? atoi(e) : 1; }();
    // pays on short reductions without a residual stream (fc1 +3.5 %, QKV +2 %); with K = 2304 and an fp32 residual
    // in flight the extra requests cost 9 %
    a.l2_prefetch = l2pf_env && !a.has_res && a.num_k_chunks <= 18 && a.num_m_tiles * a.num_n_tiles > 2 * sm_count() ? 1 : 0;
    CUtensorMap ta, tb;
    if (int rc = make_tmap_2d(&ta, A, M, K, static_cast<uint64_t>(K) * 2, kBlockM)) return rc;
    if (int rc = make_tmap_2d(&tb, W, N, K, static_cast<uint64_t>(K) * 2, a.pair ? a.block_n / 2 : a.block_n)) return rc;
    return launch_gemm(ta, tb, em, a, LaunchCtx(launch));
}

extern "C" int spg_conv3x3_h16(const void* x, const void* w, int B, int H, int W, int Cin, int Cout,
                               const spg_epilogue_t* ep, const spg_launch_t* launch) {
    using namespace spg;
    SPG_CHECK_ARG(x != nullptr && w != nullptr, "x / w is NULL");
    SPG_CHECK_ARG(B > 0 && H > 0 && W > 0, "bad conv shape B=%d H=%d W=%d", B, H, W);
    SPG_CHECK_ARG(Cin % kBlockK == 0, "Cin=%d must be a multiple of 64", Cin);
    SPG_CHECK_ARG(Cout % 16 == 0, "Cout=%d must be a multiple of 16", Cout);
    // tile = tile_h x tile_w pixels of one image.  Widths that divide 128 or are multiples of 128 tile exactly; any other
    // width (e.g. 44 / 88 / 176 / 352 at a 352 x 352 input) takes 64- or 128-pixel tile columns whose last one is ragged
    int tile_w = W < kBlockM ? W : kBlockM;
    if (kBlockM % tile_w != 0 || W % tile_w != 0) tile_w = W <= 64 ? 64 : kBlockM;
    const int tile_h = kBlockM / tile_w;
    SPG_CHECK_ARG(H % tile_h == 0, "H=%d must be a multiple of the tile height %d", H, tile_h);
    GemmArgs a{};
    EpiMaps em;
    a.tiles_x = (W + tile_w - 1) / tile_w;
    a.tile_h = tile_h;
    a.tiles_img = a.tiles_x * (H / tile_h);
    a.ragged = W % tile_w != 0 ? 1 : 0;
    SPG_CHECK_ARG(!a.ragged || tile_w >= 32, "ragged conv tiles need >= 32 pixels per row");
    a.num_m_tiles = B * a.tiles_img;
    a.M = a.num_m_tiles * kBlockM;  // tile pixels (== B*H*W unless ragged)
    a.N = Cout;
    a.block_n = pick_block_n(Cout);
    a.num_n_tiles = Cout / a.block_n;
    a.cin_chunks = Cin / kBlockK;
    a.num_k_chunks = 9 * a.cin_chunks;
    a.last_chunk_ksteps = kBlockK / kUmmaK;
    a.conv = 1;
    a.H = H;
    a.W = W;
    a.tile_w = tile_w;
    // Row-halo mode (tile = 128 pixels of one image row, narrow N): the ring is bound by bytes in flight over a
    // ~2-3 us TMA round trip, so A is loaded once per (dy, channel chunk) as a 130-pixel halo row and reused by
    // the three dx taps through shifted descriptors: 1.5-1.9x fewer bytes per FLOP than one A tile per tap.
    a.halo = (tile_h == 1 && a.block_n <= 128) ? 1 : 0;
    if (a.halo) a.num_k_chunks = 3 * a.cin_chunks;
    if (int rc = fill_epilogue(a, em, ep, B * H * W, Cout)) return rc;
    if (a.ragged && a.has_out) {  // pixels of the last tile column beyond W must be clipped: address the output as NHWC
        SPG_CHECK_ARG(!a.out_f32, "ragged conv widths support 16-bit outputs only");
        if (int rc = make_tmap_epilogue_nhwc(&em.out, ep->out, B, H, W, Cout, a.group * 16)) return rc;
    }
    decide_pair(a);
    // narrow row-halo convs whose nine tap tiles fit next to a deep A ring keep the weights resident (Cout = 64,
    // Cin = 64: 72 KB): no weight refill traffic on a kernel that is bound by shared-memory bandwidth
    static const int conv_bres_env = [] { const char* e = getenv("SPG_CONV_BRES"); return e ? atoi(e) : 1; }();
    a.b_resident = (conv_bres_env && a.halo && !a.pair && 9 * a.cin_chunks * a.block_n * 128 <= 80 * 1024 &&
                    a.num_m_tiles * a.num_n_tiles >= 4 * sm_count()) ? 1 : 0;
    CUtensorMap ta, tb;
    if (int rc = make_tmap_nhwc(&ta, x, B, H, W, Cin, tile_h, a.halo ? 130 : tile_w)) return rc;
    if (int rc = make_tmap_2d(&tb, w, Cout, 9ull * Cin, 18ull * Cin, a.pair ? a.block_n / 2 : a.block_n)) return rc;
    return launch_gemm(ta, tb, em, a, LaunchCtx(launch));
}

extern "C" int spg_conv3x3_up2_h16(const void* x, const void* w_phase, const float* corr, int B, int H, int W, int Cin,
                                   int Cout, const float* bias4, void* out, const spg_launch_t* launch) {
    using namespace spg;
    SPG_CHECK_ARG(x != nullptr && w_phase != nullptr && corr != nullptr && out != nullptr, "x / w_phase / corr / out is NULL");
    SPG_CHECK_ARG(B > 0 && H >= 2 && W > 0, "bad shape B=%d H=%d W=%d", B, H, W);
    SPG_CHECK_ARG(Cin % kBlockK == 0, "Cin=%d must be a multiple of 64", Cin);
    SPG_CHECK_ARG(Cout % 32 == 0 && 4 * Cout <= 256, "Cout=%d must be a multiple of 32 and <= 64 (4 phases in one 256-wide tile)", Cout);
    SPG_CHECK_ARG(W >= 32, "W=%d: a tile is 128 pixels of one low-resolution row (the last tile of a row may be ragged)", W);
    SPG_CHECK_ARG((reinterpret_cast<uintptr_t>(corr) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias4) & 15) == 0,
                  "corr / bias4 must be 16-byte aligned");
    GemmArgs a{};
    EpiMaps em;
    a.tiles_x = (W + kBlockM - 1) / kBlockM;
    a.tile_h = 1;
    a.tiles_img = a.tiles_x * H;
    a.ragged = W % kBlockM != 0 ? 1 : 0;  // the pixel-shuffle store map clips x >= W by itself
    a.num_m_tiles = B * a.tiles_img;
    a.M = a.num_m_tiles * kBlockM;
    a.N = 4 * Cout;
    a.block_n = a.N;
    a.num_n_tiles = 1;
    a.cin_chunks = Cin / kBlockK;
    a.num_k_chunks = 9 * a.cin_chunks;
    a.last_chunk_ksteps = kBlockK / kUmmaK;
    a.conv = 1;
    a.H = H;
    a.W = W;
    a.tile_w = kBlockM;
    a.halo = 0;
    a.up2 = 1;
    a.cout = Cout;
    a.bh = B * H;
    a.corr = corr;
    a.bias = bias4;
    a.act = SPG_ACT_RELU;
    a.has_out = 1;
    a.epi_warps = kEpiWarpsDefault;
    a.group = 2;  // 32 columns: never straddles a phase (Cout % 32 == 0)
    a.row_bytes = 64;
    a.buf_bytes = 32 * 64;
    a.piece_shift = 1;
    a.piece_mask = 3;
    memset(&em, 0, sizeof(em));
    if (int rc = make_tmap_up2_out(&em.out, out, static_cast<uint64_t>(B) * H, W, Cout, 32)) return rc;
    decide_pair(a);
    if (a.tiles_x % 2 != 0) a.pair = 0;  // both CTAs of a pair must sit in the same image row (same weight set)
    CUtensorMap ta, tb;
    if (int rc = make_tmap_nhwc(&ta, x, B, H, W, Cin, 1, kBlockM)) return rc;
    if (int rc = make_tmap_2d(&tb, w_phase, 3ull * a.N, 9ull * Cin, 18ull * Cin, a.pair ? a.block_n / 2 : a.block_n)) return rc;
    return launch_gemm(ta, tb, em, a, LaunchCtx(launch));
}

#ifdef SPG_TRACE
extern "C" int spg_debug_trace_dump(long long* host, int n) {
    using namespace spg;
    SPG_CHECK_CUDA(cudaDeviceSynchronize());
    SPG_CHECK_CUDA(cudaMemcpyFromSymbol(host, g_trace, sizeof(long long) * (n < 64 * 12 ? n : 64 * 12)));
    return SPG_OK;
}
#endif
