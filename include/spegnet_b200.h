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
 *     global state inside; work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - return SPG_OK (0) or a negative SPG_ERR_* code; spg_last_error() gives the text (thread-local).
 *   - activations are NHWC ("tokens x channels") bf16 unless stated; the residual stream and all
 *     logits are fp32; accumulation is fp32 (TMEM).
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

#define SPG_BF16 0
#define SPG_F32 1

typedef void* spg_stream_t; /* cudaStream_t */

/* Library version (major*10000 + minor*100 + patch). */
int spg_version(void);
/* Text of the last error raised on the calling thread ("" if none). */
const char* spg_last_error(void);
/* SPG_OK iff the current CUDA device is compute capability 10.x (B200). */
int spg_device_check(void);
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
    int out_dtype;         /* SPG_BF16 / SPG_F32 */
    const float* head_w;   /* [N] fp32 or NULL; needs N <= 256 and N % 16 == 0 */
    float head_b;
    float* head_out;       /* [M] fp32 */
} spg_epilogue_t;

/*
 * out[M,N] = epilogue(A[M,K] @ W[N,K]^T), bf16 operands, tcgen05 / TMEM accumulate, TMA-fed.
 * Replaces every nn.Linear of the Hiera trunk (qkv / proj / mlp.layers.{0,1} / blocks.{2,8,44}.proj;
 * HF:modeling_sam2.py:307-345,348-375,499-500) and every 1x1 nn.Conv2d (+ folded BatchNorm + ReLU)
 * of the head (models/feature_integration.py:198,239-241,310-314,363-367).
 * K and N may be any multiple of 8 / 16; K tails are zero-filled by TMA.
 */
int spg_linear_bf16(const void* A, const void* W, int M, int N, int K, const spg_epilogue_t* ep,
                    spg_stream_t stream);

/*
 * 3x3, stride 1, zero-pad 1 convolution as an implicit GEMM: x is NHWC bf16 [B,H,W,Cin]
 * (Cin % 64 == 0), w is [Cout, 9*Cin] bf16 with k = (ky*3+kx)*Cin + ci; the halo is produced by
 * TMA out-of-bounds zero fill, nothing is materialised.  out is [B*H*W, Cout].
 * Replaces nn.Conv2d(k=3,p=1) + BatchNorm2d(eval) + ReLU in EdgeDetectionModule and DecoderBlock
 * (models/object_detection.py:115-123,150-152,193-198,230-236).
 */
int spg_conv3x3_bf16(const void* x, const void* w, int B, int H, int W, int Cin, int Cout,
                     const spg_epilogue_t* ep, spg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPEGNET_B200_H */
