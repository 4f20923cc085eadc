// Error plumbing, device queries and TMA tensor-map encoding for libspegnet_b200.so.
#include "common.h"
#include "half16.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace spg {

std::atomic<long long> g_launches{0};

char* error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int sm_count() {
    static int cached[128] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int& c = cached[dev & 127];
    if (c == 0) {
        int n = 0;
        c = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return c;
}

int pdl_pinned() {
    static const int pinned = [] { const char* e = getenv("SPG_PDL"); return e == nullptr ? -1 : (atoi(e) != 0 ? 1 : 0); }();
    return pinned;
}

namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

// libcuda is not linked (the .so must load on a GPU-less build box); the encoder is looked up lazily.
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// Memo of encoded tensor maps (thread-local, direct-mapped): a forward re-encodes the same ~1000 (pointer, shape, box)
// combinations every call -- workspaces and weights do not move -- and cuTensorMapEncodeTiled costs 1-2 us each.  Pure
// memoisation of a pure function of its arguments: no effect on results, nothing shared between threads.
struct TmapKey {
    const void* base;
    cuuint64_t dims[5];
    cuuint64_t strides[4];
    cuuint32_t box[5];
    uint32_t rank;
    int dtype;
    int swz;
    bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapSlot {
    TmapKey key;
    CUtensorMap map;
    bool valid;
};
constexpr int kTmapCacheSlots = 4096;

int encode(CUtensorMap* out, const void* base, uint32_t rank, const cuuint64_t* dims,
           const cuuint64_t* strides, const cuuint32_t* box,
           int dtype = -1 /* -1: the library's 16-bit type, 1: fp32 */,
           CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    static thread_local TmapSlot* cache = nullptr;
    if (cache == nullptr) cache = static_cast<TmapSlot*>(calloc(kTmapCacheSlots, sizeof(TmapSlot)));
    TmapKey key;
    memset(&key, 0, sizeof(key));
    key.base = base;
    key.rank = rank;
    key.dtype = dtype;
    key.swz = static_cast<int>(swz);
    for (uint32_t i = 0; i < rank; ++i) {
        key.dims[i] = dims[i];
        key.box[i] = box[i];
        if (i + 1 < rank) key.strides[i] = strides[i];
    }
    uint64_t h = 1469598103934665603ull;
    const unsigned char* kb = reinterpret_cast<const unsigned char*>(&key);
    for (size_t i = 0; i < sizeof(key); i += 4) h = (h ^ *reinterpret_cast<const uint32_t*>(kb + i)) * 1099511628211ull;
    TmapSlot* slot = cache != nullptr ? &cache[(h >> 20) & (kTmapCacheSlots - 1)] : nullptr;
    if (slot != nullptr && slot->valid && slot->key == key) {
        *out = slot->map;
        return SPG_OK;
    }
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return fail(SPG_ERR_CUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(SPG_ERR_INVALID, "TMA operand must be 16-byte aligned");
    cuuint32_t elem_strides[5] = {1, 1, 1, 1, 1};
    const CUtensorMapDataType dt = dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                   : (kHalfIsFp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
    CUresult r = fn(out, dt, rank, const_cast<void*>(base), dims, strides, box,
                    elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SPG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    if (slot != nullptr) {
        slot->key = key;
        slot->map = *out;
        slot->valid = true;
    }
    return SPG_OK;
}

}  // namespace

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes,
                 uint32_t box_rows) {
    if (row_pitch_bytes % 16 != 0) return fail(SPG_ERR_INVALID, "row pitch %llu B is not a multiple of 16", (unsigned long long)row_pitch_bytes);
    if (box_rows == 0 || box_rows > 256) return fail(SPG_ERR_INVALID, "TMA box rows %u out of range", box_rows);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_pitch_bytes};
    cuuint32_t box[2] = {64, box_rows};
    return encode(out, base, 2, dims, strides, box);
}

int make_tmap_epilogue(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, int elem_is_f32,
                       uint32_t box_cols) {
    const uint64_t esize = elem_is_f32 ? 4 : 2;
    const uint64_t row_bytes = box_cols * esize;
    if (row_bytes != 32 && row_bytes != 64 && row_bytes != 128) return fail(SPG_ERR_INVALID, "bad epilogue box");
    if ((cols * esize) % 16 != 0) return fail(SPG_ERR_INVALID, "epilogue row pitch must be a multiple of 16 bytes");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * esize};
    cuuint32_t box[2] = {box_cols, 32};
    return encode(out, base, 2, dims, strides, box, elem_is_f32 ? 1 : -1,
                  row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B));
}

int make_tmap_epilogue_nhwc(CUtensorMap* out, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C,
                            uint32_t box_cols) {
    const uint64_t row_bytes = box_cols * 2;
    if (row_bytes != 32 && row_bytes != 64 && row_bytes != 128) return fail(SPG_ERR_INVALID, "bad epilogue box");
    if ((C * 2) % 16 != 0) return fail(SPG_ERR_INVALID, "channel pitch must be a multiple of 16 bytes");
    cuuint64_t dims[4] = {C, W, H, B};
    cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    cuuint32_t box[4] = {box_cols, 32, 1, 1};
    return encode(out, base, 4, dims, strides, box, -1,
                  row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B));
}

int make_tmap_up2_out(CUtensorMap* out, const void* base, uint64_t BH, uint64_t W, uint64_t C, uint32_t box_cols) {
    const uint64_t row_bytes = box_cols * 2;
    if (row_bytes != 32 && row_bytes != 64 && row_bytes != 128) return fail(SPG_ERR_INVALID, "bad epilogue box");
    if ((C * 2) % 16 != 0 || C % box_cols != 0) return fail(SPG_ERR_INVALID, "up2 store: C must be a multiple of the box width");
    cuuint64_t dims[5] = {C, 2, W, 2, BH};
    cuuint64_t strides[4] = {C * 2, 2 * C * 2, 2 * W * C * 2, 2 * 2 * W * C * 2};
    cuuint32_t box[5] = {box_cols, 1, 32, 1, 1};
    return encode(out, base, 5, dims, strides, box, -1,
                  row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B));
}

int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C,
                   uint32_t box_h, uint32_t box_w) {
    if ((C * 2) % 16 != 0) return fail(SPG_ERR_INVALID, "channel pitch must be a multiple of 16 bytes");
    cuuint64_t dims[4] = {C, W, H, B};
    cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    cuuint32_t box[4] = {64, box_w, box_h, 1};
    return encode(out, base, 4, dims, strides, box);
}

int make_tmap_qkv_window(CUtensorMap* out, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t ld,
                         uint32_t box_c, uint32_t box_w, uint32_t box_h) {
    if ((ld * 2) % 16 != 0 || (box_c * 2) % 16 != 0) return fail(SPG_ERR_INVALID, "qkv row pitch / box must be multiples of 16 bytes");
    if (box_w > 256 || box_h > 256 || box_c > 256) return fail(SPG_ERR_INVALID, "qkv window box too large");
    cuuint64_t dims[4] = {ld, W, H, B};
    cuuint64_t strides[3] = {ld * 2, W * ld * 2, H * W * ld * 2};
    cuuint32_t box[4] = {box_c, box_w, box_h, 1};
    return encode(out, base, 4, dims, strides, box, -1, CU_TENSOR_MAP_SWIZZLE_NONE);
}

int make_tmap_qkv_5d(CUtensorMap* out, const void* base, uint64_t BH, uint64_t W, uint64_t heads, uint32_t box_d,
                     uint32_t box_w, uint32_t box_h, int swizzle_bytes) {
    const uint64_t D = heads * 72;
    cuuint64_t dims[5] = {72, heads, 3, W, BH};
    cuuint64_t strides[4] = {72 * 2, D * 2, 3 * D * 2, W * 3 * D * 2};
    cuuint32_t box[5] = {box_d, 1, 1, box_w, box_h};
    return encode(out, base, 5, dims, strides, box, -1,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B);
}

}  // namespace spg

extern "C" int spg_version(void) { return 100; /* 0.1.0 */ }

extern "C" int spg_half_is_fp16(void) { return spg::kHalfIsFp16; }

extern "C" const char* spg_last_error(void) { return spg::error_buffer(); }

extern "C" int spg_device_check(void) {
    int dev = 0;
    SPG_CHECK_CUDA(cudaGetDevice(&dev));
    int major = 0, minor = 0;
    SPG_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    SPG_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10) return spg::fail(SPG_ERR_UNSUPPORTED, "device is sm_%d%d; this library is sm_100a only", major, minor);
    return SPG_OK;
}

extern "C" long long spg_launch_count(void) { return spg::g_launches.load(std::memory_order_relaxed); }
extern "C" void spg_launch_count_reset(void) { spg::g_launches.store(0, std::memory_order_relaxed); }
