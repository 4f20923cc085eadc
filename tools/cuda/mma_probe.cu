// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS operands in shared memory) as a function of N, of
// the number of independent TMEM accumulators, and of HOW the single issuing thread is selected:
//   mode 0: `if (threadIdx.x == 0)` -- divergent branch; ptxas wraps every UTCHMMA in an ELECT / BRA.U.ANY loop
//   mode 1: the whole warp runs the loop, the MMA sits under an `elect.sync` predicate (the CUTLASS idiom)
// One CTA, no TMA, operands resident: isolates the issue path + tensor pipe from the memory system.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I spegnet_b200/csrc -o tools/cuda/mma_probe tools/cuda/mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define SPG_FP16
#include "half16.cuh"
#include "ptx.cuh"
using namespace spg;

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

template <int kMode>
__global__ void __launch_bounds__(128, 1) probe(int n, int accs, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t a_addr = base, b_addr = base + 16384;  // A 128 x 64 (16 KB), B 256 x 64 (32 KB), 128B-swizzled K-major
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + (base - smem_u32(smem)))[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) {
        tmem_alloc(smem_u32(&tmem_slot), 512);
        tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t idesc = make_idesc_bf16_f32(128, n);
    const uint64_t a_desc = make_sw128_kmajor_desc(a_addr), b_desc = make_sw128_kmajor_desc(b_addr);
    const uint32_t stride = accs > 1 ? 256u : 0u;  // accs == 2: alternate between the two 256-column halves
    if (kMode == 0) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            for (int i = 0; i < reps; i += 4) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem_base + ((i >> 2) & 1) * stride, a_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
            }
            const long long t1 = clock64();
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            out[0] = t1 - t0;
            out[1] = clock64() - t0;
        }
    } else {
        if (threadIdx.x < 32) {
            const long long t0 = clock64();
            for (int i = 0; i < reps; i += 4) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_base + ((i >> 2) & 1) * stride, a_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
                }
                __syncwarp();
            }
            const long long t1 = clock64();
            if (elect_one()) umma_commit(smem_u32(&bar));
            __syncwarp();
            mbar_wait(smem_u32(&bar), 0);
            if (threadIdx.x == 0) {
                out[0] = t1 - t0;
                out[1] = clock64() - t0;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int reps = 4096;
    printf("mode N accumulators | cycles per MMA (issue loop) | cycles per MMA (until complete) | ideal N/256*128\n");
    const int ns[] = {64, 128, 144, 192, 256};
    for (int mode = 0; mode < 2; ++mode)
        for (int n : ns)
            for (int accs : {1, 2}) {
                if (mode == 0) probe<0><<<1, 128, 64 * 1024>>>(n, accs, reps, d);
                else probe<1><<<1, 128, 64 * 1024>>>(n, accs, reps, d);
                long long h[2];
                if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                printf("%d  %3d  %d | %7.1f | %7.1f | %5.1f\n", mode, n, accs, (double)h[0] / reps, (double)h[1] / reps, n / 256.0 * 128);
            }
    return 0;
}
