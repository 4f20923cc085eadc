/*
 * spegnet_b200 -- C-ABI of the B200-native SPEGNet inference forward pass.
 *
 * The reference (Baber-Jan/SPEGNet) has no FFI of its own: its hot path is the Python
 * `SPEGNet.forward` (models/spegnet.py:137-206), which dispatches torch.nn modules to ATen /
 * cuDNN / cuBLAS.  Each entry point below replaces one group of those library dispatches with a
 * hand-written sm_100a kernel; the comment on every function cites the reference lines it stands
 * in for.  The Python drop-in (spegnet_b200/model.py) binds these with ctypes.
 *
 * Conventions (all functions):
 *   - plain device pointers + sizes, caller-owned memory, NO allocation, NO synchronisation and NO
 *     global state inside (the only process-wide datum is the diagnostic launch counter); work is enqueued on
 *     `launch->stream` with the per-call flags of `launch` (spg_launch_t below).
 *   - return SPG_OK (0) or a negative SPG_ERR_* code; spg_last_error() gives the text (thread-local).
 *   - activations are NHWC ("tokens x channels") in the library's 16-bit storage type "h16" unless
 *     stated: IEEE half in libspegnet_b200_fp16.so, bfloat16 in libspegnet_b200_bf16.so (same sources,
 *     same entry points, spg_half_is_fp16() tells which; both use kind::f16 tcgen05 MMAs at the same
 *     rate).  Comments below say "bf16" for brevity.  The residual stream and all logits are fp32;
 *     accumulation is fp32 (TMEM).
 *   - pointers handed to TMA-fed kernels (GEMM / conv operands) must be 16-byte aligned with a
 *     row pitch that is a multiple of 16 bytes.
 */
#ifndef SPEGNET_B200_H
#define SPEGNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPG_OK 0
#define SPG_ERR_INVALID (-1)     /* bad argument (shape, alignment, null pointer) */
#define SPG_ERR_CUDA (-2)        /* CUDA runtime / driver error while enqueueing */
#define SPG_ERR_UNSUPPORTED (-3) /* device is not sm_100 or shape outside the supported set */

#define SPG_ACT_NONE 0
#define SPG_ACT_RELU 1
#define SPG_ACT_GELU 2 /* exact erf GELU, as torch.nn.GELU() in the sam2 trunk MLP */

#define SPG_H16 0 /* the library's 16-bit storage type (fp16 or bf16, see above) */
#define SPG_F32 1

/*
 * Per-call launch descriptor, the last argument of every entry point that enqueues work.  NULL = the default stream
 * with no flags.  Nothing about a launch is process-global: two host threads / streams / models cannot disturb each
 * other, and a CUDA graph captures exactly the flags its calls were made with.
 *   SPG_LAUNCH_PDL      the kernels carry cudaLaunchAttributeProgrammaticStreamSerialization and overlap their
 *                       global-data-free prologue with the tail of their predecessor in the stream
 *                       (griddepcontrol.launch_dependents / wait); pays for small batches (launch-latency regime).
 *   SPG_LAUNCH_REVERSE  walk rows / tiles / work items in descending order.  Consecutive kernels of the forward are
 *                       producer -> consumer pairs over tensors larger than the 126 MB L2; a host that alternates this
 *                       flag launch by launch lets each consumer start on the rows its producer wrote last.  Honoured
 *                       by the GEMM / conv engine, LayerNorm and the attention kernels; results do not depend on it.
 */
#define SPG_LAUNCH_PDL 1u
#define SPG_LAUNCH_REVERSE 2u
typedef struct spg_launch {
    void* stream;   /* cudaStream_t */
    unsigned flags; /* SPG_LAUNCH_* */
} spg_launch_t;

/* Library version (major*10000 + minor*100 + patch). */
int spg_version(void);
/* Text of the last error raised on the calling thread ("" if none). */
const char* spg_last_error(void);
/* SPG_OK iff the current CUDA device is compute capability 10.x (B200). */
int spg_device_check(void);
/* 1 if this build stores activations / weights as IEEE fp16, 0 if bfloat16. */
int spg_half_is_fp16(void);
/* Number of kernels this library has launched since load / since the last reset (all threads). */
long long spg_launch_count(void);
void spg_launch_count_reset(void);

/*
 * Epilogue applied to a GEMM / convolution accumulator tile while it is read out of TMEM:
 *   v = acc + bias[n];  v = act(v);  v += residual[row % res_rows or row][n];  out[row][n] = v
 * and, optionally, a fused 1x1 "N -> 1" projection  head_out[row] = sum_n v[n]*head_w[n] + head_b
 * (replaces the separate nn.Conv2d(C,1,1) prediction heads, models/object_detection.py:126-130,306,339).
 */
typedef struct spg_epilogue {
    const float* bias;     /* [N] fp32 or NULL */
    int act;               /* SPG_ACT_* */
    const float* residual; /* fp32 [M, N] (or [res_rows, N]) or NULL; may alias `out` when out is fp32 */
    int res_rows;          /* 0: one residual row per output row; >0: row index modulo res_rows */
    void* out;             /* [M, N] row-major, bf16 or fp32; NULL = do not store (head only) */
    int out_dtype;         /* SPG_H16 / SPG_F32 */
    const float* head_w;   /* [N] fp32 or NULL; needs N <= 256 and N % 16 == 0 */
    float head_b;
    float* head_out;       /* [M] fp32 */
    /*
     * LayerNorm folded into the GEMMs around it (blocks.{i}.norm1 / norm2, HF:modeling_sam2.py:495,528).  Every
     * LayerNorm input of the trunk is the fp32 output of a residual GEMM and is read only by GEMMs, so:
     *  - the PRODUCER (ln_emit_out != NULL, needs residual + fp32 out) also stores xc = out - c as a 16-bit copy, c =
     *    the row's previous mean (from ln_prev_rec, 0 if NULL), and writes the row record ln_emit_rec[row][32] =
     *    {c, P, (sum xc, sum xc^2) x P partials}: one partial per (n-tile, epilogue half), plain stores in fixed slots,
     *    so the statistics are deterministic;
     *  - the CONSUMER (ln_fold_rec != NULL) runs on A = xc with W' = W diag(gamma) and applies, per output row,
     *    out = rstd * (acc - m * ln_fold_cw[n]) + bias[n]   (m, rstd from the record, over ln_cols channels; bias must
     *    already contain W beta; ln_fold_cw[n] = sum_k W'[n,k]), then the activation.
     */
    const float* ln_fold_rec; /* consumer: [M][32] fp32 row records of A */
    const float* ln_fold_cw;  /* consumer: [N] fp32 */
    int ln_cols;              /* channels of the normalised row (consumer: K, producer: N) */
    float ln_eps;
    float* ln_emit_rec;       /* producer: [M][32] fp32 row records of out */
    const float* ln_prev_rec; /* producer: records of the residual input rows, or NULL */
    void* ln_emit_out;        /* producer: [M, N] 16-bit centred copy of out */
    /*
     * LayerNorm APPLIED by the producer (the default trunk path): besides out (the fp32 residual stream, needs
     * residual != NULL) the GEMM stores ln_apply_out[M, N] = LayerNorm(out) * ln_apply_gamma + ln_apply_beta as 16 bit
     * (eps = ln_eps) -- the operand of the next block's qkv / MLP GEMMs, so no separate LayerNorm pass over HBM exists.
     * The CTAs (or CTA pairs) that hold the n-tiles of one 128-row block are launched as one thread-block cluster: each
     * epilogue thread parks its final fp32 values in the consumed TMEM accumulator columns, publishes {mean, M2} of its
     * column slice to its peers' shared memory (st.shared::cluster + mbarrier), combines the slices in fixed column
     * order (Chan's formula; independent of M and of the grid) and normalises in a second pass over TMEM.
     * N must be 144, 288 or 576 (one, two or three n-tiles); no activation, no head.
     */
    const float* ln_apply_gamma; /* [N] fp32 or NULL */
    const float* ln_apply_beta;  /* [N] fp32 */
    void* ln_apply_out;          /* [M, N] 16-bit */
} spg_epilogue_t;

/*
 * out[M,N] = epilogue(A[M,K] @ W[N,K]^T), bf16 operands, tcgen05 / TMEM accumulate, TMA-fed.
 * Replaces every nn.Linear of the Hiera trunk (qkv / proj / mlp.layers.{0,1} / blocks.{2,8,44}.proj;
 * HF:modeling_sam2.py:307-345,348-375,499-500) and every 1x1 nn.Conv2d (+ folded BatchNorm + ReLU)
 * of the head (models/feature_integration.py:198,239-241,310-314,363-367).
 * K and N may be any multiple of 8 / 16; K tails are zero-filled by TMA.
 */
int spg_linear_h16(const void* A, const void* W, int M, int N, int K, const spg_epilogue_t* ep,
                    const spg_launch_t* launch);

/*
 * 3x3, stride 1, zero-pad 1 convolution as an implicit GEMM: x is NHWC bf16 [B,H,W,Cin]
 * (Cin % 64 == 0), w is [Cout, 9*Cin] bf16 with k = (ky*3+kx)*Cin + ci; the halo is produced by
 * TMA out-of-bounds zero fill, nothing is materialised.  out is [B*H*W, Cout].  A tile is 128 pixels (tile_h x tile_w) of
 * one image: W may divide 128, be a multiple of 128, or be anything else >= 32 (e.g. 44 / 88 / 176 at a 352 x 352 input),
 * in which case 64- / 128-pixel tile columns are used and the last one is ragged (its missing pixels are zero fill on the
 * way in and clipped by an NHWC store map on the way out; 16-bit outputs only); H must be a multiple of tile_h.
 * Replaces nn.Conv2d(k=3,p=1) + BatchNorm2d(eval) + ReLU in EdgeDetectionModule and DecoderBlock
 * (models/object_detection.py:115-123,150-152,193-198,230-236).
 */
int spg_conv3x3_h16(const void* x, const void* w, int B, int H, int W, int Cin, int Cout,
                     const spg_epilogue_t* ep, const spg_launch_t* launch);

/*
 * Bilinear x2 upsample (align_corners=False) FUSED into the 3x3 convolution that consumes it:
 *   out[B, 2H, 2W, Cout] = relu(conv3x3_pad1(up2(x))[..., co] + bias)      x: NHWC bf16 [B,H,W,Cin]
 * Replaces F.interpolate(x, size=2x) followed by conv1 + bn1 + relu of the last DecoderBlock (no edge branch:
 * models/object_detection.py:219,230-232 with models/spegnet.py:187-191), without materialising the upsampled map.
 * Both operators are linear, so each of the 4 output phases (row phase a, column phase b of the x2 grid) is a 3x3
 * convolution on the LOW-resolution grid with its own folded weights: an implicit GEMM with N = 4*Cout (one 256-wide
 * tcgen05 tile for Cout = 64 instead of four times as many 64-wide ones) and a pixel-shuffle TMA store.
 *   w_phase [3][4*Cout][9*Cin] bf16: one weight set per ROW CLASS of the low-resolution pixel (0: first image row,
 *       1: interior, 2: last row), n = (a*2 + b)*Cout + co, k = (dy*3 + dx)*Cin + ci.  The first / last row differ
 *       because the bilinear clamp and the conv's zero padding meet there.
 *   corr [2][B*H][4*Cout] fp32: pre-activation corrections of the first (side 0) / last (side 1) image COLUMN, the
 *       same border effect along x: corr = spg_up2_border_gather_h16(x) @ delta_w^T via spg_linear_h16.
 *   bias4 [4*Cout] fp32 (the folded BatchNorm shift repeated per phase).
 * W >= 32 (a tile is 128 pixels of one low-resolution row; the last tile of a row may be ragged), Cin % 64 == 0,
 * Cout % 32 == 0, Cout <= 64.  The host-side weight folding is spegnet_b200/model.py
 * (`up2_phase_weights`), pinned against F.interpolate + F.conv2d in tests/test_host.py and tests/test_gpu_ops.py.
 */
int spg_conv3x3_up2_h16(const void* x, const void* w_phase, const float* corr, int B, int H, int W, int Cin, int Cout,
                        const float* bias4, void* out, const spg_launch_t* launch);

/* Left operand of the border-column correction GEMM above: out [2][B*H][9*C] bf16,
 * out[s][b*H+y][(cls*3+dy)*C + c] = x[b, y+dy-1, s ? W-1 : 0, c] in the block of y's row class, zero elsewhere. */
int spg_up2_border_gather_h16(const void* x, void* out, int B, int H, int W, int C, const spg_launch_t* launch);

/*
 * y[M,C] (bf16) = LayerNorm(x[M,C] (fp32 residual stream)) * gamma + beta, eps as given (1e-6 in Hiera).
 * Replaces blocks.{i}.norm1 / norm2 (HF:modeling_sam2.py:495,528).  C % 4 == 0, C <= 1152.
 */
int spg_layernorm_f32_h16(const float* x, const float* gamma, const float* beta, void* y, int M, int C,
                           float eps, const spg_launch_t* launch);

/*
 * dst[b, y, x, :] = src[b, y, x, :] for y < H, x < W between two NHWC token grids [B, Hs, Ws, C] -> [B, Hd, Wd, C]
 * (h16, C % 8 == 0); the rest of dst is left untouched.  Window attention on a grid that does not tile into windows
 * (input sizes that are not multiples of 256): the norm1 output is copied into a zero-initialised padded grid before
 * the qkv projection -- the padded tokens are zeros there and take part in the softmax as keys, exactly as
 * window_partition does (HF:modeling_sam2.py:395-399) -- and the attention output is cropped back (:435-437).
 */
int spg_copy_grid_h16(const void* src, int Hs, int Ws, void* dst, int Hd, int Wd, int B, int H, int W, int C,
                      const spg_launch_t* launch);

/*
 * The same LayerNorm, BIT-IDENTICAL to what a residual GEMM with spg_epilogue_t.ln_apply_* stores for the same rows
 * (same per-slice sequential statistics, same merge, same normalisation): the host may pick the fused form for large
 * batches and this kernel in the latency regime without the results depending on the choice.  C in {144, 288, 576}.
 */
int spg_layernorm_matched_f32_h16(const float* x, const float* gamma, const float* beta, void* y, int M, int C,
                                  float eps, const spg_launch_t* launch);

/*
 * im2col for the 7x7 / stride 4 / pad 3 patch embedding: x fp32 NCHW [B,3,S,S] (16-byte aligned) -> cols bf16
 * [B*(S/4)^2, 168], column k = (ky*3 + c)*8 + kx for kx < 7, zero at kx = 7 (each (ky, c) group is one aligned
 * 16-byte gather).  The projection itself is spg_linear_h16 over a weight matrix packed in the same column order, with
 * the positional embedding as a broadcast residual (res_rows = (S/4)^2).
 * Replaces PatchEmbed's Conv2d(3,144,7,4,3) + permute (HF:modeling_sam2.py:138-148).
 */
int spg_patchify_7x7s4(const float* x, void* cols, int B, int S, const spg_launch_t* launch);

/* 2x2 / stride-2 max pool of an fp32 NHWC map: the pooled shortcut of blocks 2 / 8 / 44
 * (do_pool(self.proj(x)), HF:modeling_sam2.py:271-279,499-500). */
int spg_maxpool2x2_f32(const float* x, float* y, int B, int H, int W, int C, const spg_launch_t* launch);

/* fp32 -> bf16 copy (stage outputs of the residual stream become GEMM operands of the head). n % 8 == 0. */
int spg_cast_f32_h16(const float* x, void* y, long long n, const spg_launch_t* launch);

/*
 * Multi-head attention over non-overlapping window x window token tiles of the NHWC grid [B,H,W]
 * (window = 0: global), head_dim 72, softmax scale 1/sqrt(72).  qkv is the fused projection output
 * bf16 [B*H*W, 3*D] with columns ordered [q|k|v][head][72]; with q_pool the queries are 2x2 max-pooled
 * inside each window and `out` is the [B, H/2, W/2, D] grid.  Window partition, q-pool and
 * un-partition are addressing only.  Replaces window_partition / MultiScaleAttention / window_unpartition
 * (HF:modeling_sam2.py:307-345,378-438,503-525).
 */
int spg_window_attention_h16(const void* qkv, void* out, int B, int H, int W, int D, int heads, int window,
                              int q_pool, const spg_launch_t* launch);

/*
 * The tcgen05 / TMEM implementation of the same operation for window == 16 without query pooling (the 32 windowed
 * stage-3 blocks of Hiera-L): S = Q K^T and O = P V as tcgen05.mma tiles with the softmax probabilities kept in
 * TMEM.  spg_window_attention_h16 dispatches to it automatically; other geometries return SPG_ERR_UNSUPPORTED.
 */
int spg_window_attention_tc_h16(const void* qkv, void* out, int B, int H, int W, int D, int heads, int window,
                                int q_pool, const spg_launch_t* launch);

/*
 * out[b,y,x,:] = concat(bilinear(src0 [B,h0,w0,c0]), bilinear(src1 [B,h1,w1,c1])) resized to Ho x Wo,
 * align_corners=False, bf16 NHWC; src1 may be NULL with c1 = 0.  Replaces F.interpolate + torch.cat in
 * DecoderBlock.forward (models/object_detection.py:219-227).  Channel counts % 8 == 0.
 */
int spg_upsample_concat_h16(const void* src0, int h0, int w0, int c0, const void* src1, int h1, int w1, int c1,
                             void* out, int B, int Ho, int Wo, const spg_launch_t* launch);

/*
 * CFI fusion tail.  Conv1x1(concat(f2, up2(f3), up4(f4))) is linear, so the three per-scale products
 * g2 [B,Hs,Hs,C], g3 [B,Hs/2,Hs/2,C], g4 [B,Hs/4,Hs/4,C] (fp32, BatchNorm scale folded into the weights)
 * are computed at native resolution by spg_linear_h16 and combined here:
 *   fused = relu(g2 + up2(g3) + up4(g4) + bias)   (bf16 NHWC),  row_sums[b,y,:] = sum_x fused[b,y,x,:]
 * Replaces F.interpolate x2 + torch.cat + conv1x1 + bn + relu (models/feature_integration.py:229-241)
 * with 4x fewer MACs and no 2016-channel concat; row_sums feeds the SE squeeze (:147).
 */
int spg_fusion_combine(const float* g2, const float* g3, const float* g4, const float* bias, void* fused,
                       float* row_sums, int B, int Hs, int C, const spg_launch_t* launch);

/* row_sums[b,y,:] = sum_x x[b,y,x,:] for a bf16 NHWC map (AdaptiveAvgPool2d(1) partials,
 * models/feature_integration.py:336,401). */
int spg_row_sums_h16(const void* x, float* row_sums, int B, int H, int W, int C, const spg_launch_t* launch);

/*
 * Pooled-vector MLP, one CTA per image.  mean = sum_rows(row_sums) / count, then
 *   W2 != NULL: out[b, C] = sigmoid(W2[C,R] @ relu(W1[R,C] @ mean (+ b1)))      (SE gate, :121-126,149)
 *   W2 == NULL: out[b, R] = relu(W1[R,C] @ mean + b1)          (e-ASPP global branch, :335-345,401)
 */
int spg_pooled_mlp(const float* row_sums, int rows, int count, const float* W1, const float* b1, int R,
                   const float* W2, float* out, int B, int C, const spg_launch_t* launch);

/* x[b,p,c] *= gate[b,c] in place, bf16 NHWC (SE rescale, models/feature_integration.py:151). */
int spg_scale_channels_h16(void* x, const float* gate, int B, int HW, int C, const spg_launch_t* launch);

/*
 * e-ASPP core in one pass over x bf16 [B,H,W,128]: four depth-wise dilated 3x3 branches (+BN+ReLU),
 * the broadcast global branch gvec [B,128], the 640-channel concat and the grouped 1x1 fusion conv
 * (+BN+ReLU): y[b,y,x,g] = relu(sum_j wf[g][j] * cat[5g+j] + wf_bias[g]), cat[c] = branch[c/128][c%128].
 * dw is [4][9][128] (tap-major, BN scale folded), dw_bias [4][128], wf [128][5], dilations int[4].
 * Replaces models/feature_integration.py:397-412.
 */
int spg_easpp_branches(const void* x, const float* dw, const float* dw_bias, const float* gvec, const float* wf,
                       const float* wf_bias, void* y, int B, int H, int W, const int* dilations,
                       const spg_launch_t* launch);

/*
 * Mask quantisation of the reference's metric wrapper on the GPU (utils/metrics.py:205-210): mask = uint8(
 * sigmoid(logit) * 255) with truncation (any HW; double_sigmoid = 1: sigmoid applied twice, the evaluator path,
 * engine/evaluator.py:544 + utils/metrics.py:209), gt is uint8 with foreground > 128.  stats[b][8] (uint32) =
 * {255 - min q, max q, #gt foreground, sum q over gt background, sum q over gt foreground, 0, 0, 0}: integer
 * partials from which MAE after the min-max normalisation of py_sod_metrics follows exactly; they are what the
 * ranks all_gather in the sharded evaluation (8 x 4 bytes per image instead of the mask).
 */
int spg_mask_stats_u8(const float* logits, const unsigned char* gt, unsigned char* mask, unsigned* stats, int B,
                      int HW, int double_sigmoid, const spg_launch_t* launch);

/*
 * The five camouflaged-object scores of the reference's metric wrapper, per image, on the GPU in fp64:
 * replaces MetricsProcessor._process_single_sample (utils/metrics.py:142-167: py_sod_metrics Smeasure,
 * WeightedFmeasure, MAE, Emeasure["adp"], Fmeasure["curve"].mean()) and the device->host copy + process pool
 * around it (utils/metrics.py:224-231).  Inputs are the uint8 pairs that wrapper builds (:209-220): pred = uint8
 * mask (e.g. from spg_mask_stats_u8), gt = uint8 with foreground > 128, both [B,H,W] contiguous.
 *
 * spg_sod_gt_prepare_u8: ground truth only (cache it per dataset): nearest[b,y,x] = y'*W + x' of the nearest
 *   foreground pixel (-1 when the image has none), bit-identical to scipy.ndimage.distance_transform_edt(
 *   return_indices=True) including its choice among equidistant pixels, and gt_stats[b][4] = {#fg, sum of fg
 *   rows, sum of fg columns, 0} (the S-measure centroid).
 * spg_sod_scores_u8: scores[b][5] = {S-alpha, weighted F-beta, MAE, adaptive E-phi, mean F-beta} as fp64.
 * Both need a workspace of spg_sod_workspace_bytes(B,H,W) bytes (contents are scratch).
 */
size_t spg_sod_workspace_bytes(int B, int H, int W);
int spg_sod_gt_prepare_u8(const unsigned char* gt, int B, int H, int W, int* nearest, unsigned long long* gt_stats,
                          void* workspace, size_t ws_bytes, const spg_launch_t* launch);
int spg_sod_scores_u8(const unsigned char* pred, const unsigned char* gt, const int* nearest,
                      const unsigned long long* gt_stats, int B, int H, int W, double* scores, void* workspace,
                      size_t ws_bytes, const spg_launch_t* launch);

/*
 * Image preprocessing of the reference on the GPU: CODImageProcessor.process_image (utils/image_processor.py:114-134):
 * img uint8 HWC RGB [H,W,3] (device) -> /255 -> F.interpolate(size=(S,S), mode='bilinear', align_corners=False,
 * antialias=True) -> (x - mean) / std -> out fp32 CHW [3,S,S] (the model's input layout).  The resampler restates ATen's
 * separable anti-aliasing kernel (triangle filter of support max(in/out, 1); width, then height) with its index
 * arithmetic, so windows and weights are ATen's.  mean3 / std3 are HOST arrays of 3 floats (ImageNet statistics in the
 * reference); workspace = spg_preprocess_workspace_bytes(H, W, S) bytes of device scratch.
 */
size_t spg_preprocess_workspace_bytes(int H, int W, int S);
int spg_preprocess_rgb_u8(const unsigned char* img, int H, int W, float* out, int S, const float* mean3,
                          const float* std3, void* workspace, size_t ws_bytes, const spg_launch_t* launch);

/*
 * dst[b] = F.interpolate(src[b], size=(Ho,Wo), mode='bilinear', align_corners=False) for fp32 maps [B,Hi,Wi], followed
 * by sigmoid when apply_sigmoid != 0: the per-image resize of the finest prediction / edge map to the original or
 * ground-truth size (engine/predictor.py:350-365, engine/evaluator.py:539-554).
 */
int spg_resize_bilinear_f32(const float* src, int B, int Hi, int Wi, float* dst, int Ho, int Wo, int apply_sigmoid,
                            const spg_launch_t* launch);

/* bf16 NHWC [B,HW,C] -> fp32 NCHW [B,C,HW] (materialises `features` entries of the output dict on demand). */
int spg_nhwc_h16_to_nchw_f32(const void* x, float* y, int B, int HW, int C, const spg_launch_t* launch);

#ifdef __cplusplus
}
#endif
#endif /* SPEGNET_B200_H */
