// Host-side helpers shared by every translation unit of libspegnet_b200.so:
// error reporting behind the C-ABI (include/spegnet_b200.h) and TMA tensor-map encoding.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/spegnet_b200.h"

namespace spg {

// Thread-local last-error text returned by spg_last_error().
char* error_buffer();
int fail(int code, const char* fmt, ...);

#define SPG_CHECK_ARG(cond, ...)                                  \
    do {                                                          \
        if (!(cond)) return ::spg::fail(SPG_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define SPG_CHECK_CUDA(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return ::spg::fail(SPG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                   \
    } while (0)

// Launch-error check that never synchronises (the C-ABI contract: no sync inside).
#define SPG_CHECK_LAUNCH() SPG_CHECK_CUDA(cudaPeekAtLastError())

// bf16 K-major 2-D operand [rows, cols] (cols contiguous), box = [box_rows, 64 cols], 128B swizzle.
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                 uint64_t row_pitch_bytes, uint32_t box_rows);

// bf16 NHWC activation [B, H, W, C] for implicit-GEMM convolution: box = [1, box_h, box_w, 64 ch];
// out-of-bounds coordinates (the conv halo) are zero-filled by the TMA unit.
int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C,
                   uint32_t box_h, uint32_t box_w);

// [rows, cols] row-major epilogue tile map (output store / fp32 residual load): box = 32 rows x box_cols
// columns (16 or 32); the swizzle mode equals the box row size (32 / 64 / 128 bytes), which makes the
// row-per-thread staging accesses conflict free.
int make_tmap_epilogue(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, int elem_is_f32,
                       uint32_t box_cols);

// The same 32-pixel x box_cols epilogue tile, but addressed as pixels of an NHWC map [B, H, W, C] (16-bit): box =
// {box_cols, 32 pixels of one image row, 1, 1}; pixels with x >= W are clipped by the TMA unit.  Used by convolutions
// whose width is not a multiple of their tile width.
int make_tmap_epilogue_nhwc(CUtensorMap* out, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C,
                            uint32_t box_cols);

// Pixel-shuffle store map of the fused "bilinear x2 -> 3x3 conv" kernel: out [B, 2H, 2W, C] viewed as the 5-D tensor
// [C, 2 (column phase), W, 2 (row phase), B*H]; a box {box_cols, 1, 32, 1, 1} is 32 low-resolution pixels of one
// output phase -- in shared memory the same 32 x box_cols tile the 2-D epilogue map stores.
int make_tmap_up2_out(CUtensorMap* out, const void* base, uint64_t BH, uint64_t W, uint64_t C, uint32_t box_cols);

// qkv [B*H*W, ld] viewed as [B, H, W, ld]: box = {box_c channels, box_w, box_h, 1 image}, no swizzle
// (dense rows of box_c elements in shared memory).  One box = the K or V tile of one attention window / head.
int make_tmap_qkv_window(CUtensorMap* out, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t ld,
                         uint32_t box_c, uint32_t box_w, uint32_t box_h);

// qkv [B*H*W, 3*heads*72] viewed as the 5-D tensor [72 (head dim), heads, 3 (q|k|v), W, B*H]: a box
// {box_d, 1, 1, box_w, box_h} is one head's slice of a window's tokens.  Because dim 0 has extent 72, a box that
// starts at d = 64 with box_d = 16 gets dims 72..79 ZERO-FILLED by the TMA unit: the zero padding of the head
// dimension to a multiple of the UMMA K step costs nothing.  swizzle_bytes = 128 (box_d = 64) or 32 (box_d = 16).
int make_tmap_qkv_5d(CUtensorMap* out, const void* base, uint64_t BH, uint64_t W, uint64_t heads, uint32_t box_d,
                     uint32_t box_w, uint32_t box_h, int swizzle_bytes);

int sm_count();  // of the CURRENT device (cached per device)

// One-time-per-DEVICE guard for state that lives in a CUDA context (cudaFuncSetAttribute, __constant__ uploads): a
// process that drives several GPUs must repeat it on each of them.  `if (once.needed()) { ...; once.done(); }`
struct PerDeviceOnce {
    unsigned long long mask[2] = {0, 0};  // up to 128 devices; benign race (the guarded work is idempotent)
    static int device() {
        int d = 0;
        cudaGetDevice(&d);
        return d & 127;
    }
    bool needed() const {
        const int d = device();
        return ((mask[d >> 6] >> (d & 63)) & 1ull) == 0;
    }
    void done() {
        const int d = device();
        mask[d >> 6] |= 1ull << (d & 63);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and begins with
//     griddepcontrol.launch_dependents   -- the NEXT kernel in the stream may be scheduled as soon as every CTA of this
//                                           one has started (its CTAs fill SMs as ours drain)
//     griddepcontrol.wait                -- block until the PREVIOUS kernel has completed and its writes are visible
// placed after the part of the prologue that touches no global data (barrier init, TMEM allocation, tensor-map
// prefetch).  All global reads AND writes of a kernel come after its wait, so the stream's dependency semantics are
// unchanged (no RAW / WAR hazard); what is hidden is the launch latency, the prologue and the tail of the previous
// kernel -- 371 serialised launches per forward.  Requested per call (SPG_LAUNCH_PDL); without the attribute the
// instructions are no-ops.
// ---------------------------------------------------------------------------------------------------------------
// SPG_PDL=0|1 in the environment pins the switch for every launch (debugging); -1 = not pinned.
int pdl_pinned();

// Per-call launch state decoded from the C-ABI descriptor (spg_launch_t): stream, programmatic dependent launch,
// traversal direction.  Nothing here is process-global.
//
// Traversal direction: consecutive kernels of the forward are producer -> consumer pairs over tensors larger than the
// 126 MB L2, so a consumer that walks its rows in the OPPOSITE order of its producer starts on the lines that are
// still resident.  The GEMM / conv engine, LayerNorm and the attention kernels honour the flag (tile / block / item
// index i -> n-1-i); the host alternates it launch by launch.
struct LaunchCtx {
    cudaStream_t stream;
    bool pdl;
    bool reverse;
    explicit LaunchCtx(const spg_launch_t* l)
        : stream(l != nullptr ? static_cast<cudaStream_t>(l->stream) : nullptr),
          pdl(pdl_pinned() >= 0 ? pdl_pinned() != 0 : (l != nullptr && (l->flags & SPG_LAUNCH_PDL) != 0)),
          reverse(l != nullptr && (l->flags & SPG_LAUNCH_REVERSE) != 0) {}
    operator cudaStream_t() const { return stream; }
};

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
    pdl_launch_dependents();
    pdl_wait();
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, const LaunchCtx& st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = st.pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace spg
