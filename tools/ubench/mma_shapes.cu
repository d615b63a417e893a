// tcgen05 MMA issue patterns under the power cap: what would B-concatenation buy?  
//
// One persistent CTA per SM, operands resident in shared memory in the K-major SWIZZLE_NONE layout of conv_tc.cu (no
// TMA traffic: this measures the MMA side alone), accumulators in TMEM.  Per "k-step" of an EXACT (fp16 hi/lo)
// convolution the patterns issue:
//   0  3 x (M128 N128 K16)                a_hi*w_hi, a_hi*w_lo, a_lo*w_hi into one accumulator   (conv_tc.cu today)
//   1  1 x (M128 N256 K16) + 1 x N128     a_hi*[w_hi | w_lo] into (main | cross), a_lo*w_hi into cross (B-concatenation)
//   2  1 x (M128 N128 K16)                FAST mode
//   3  3 x (M128 N32 K16)                 context model today
//   4  1 x (M128 N64 K16) + 1 x N32       context model with B-concatenation
//   5  1 x (M128 N128 K16) + 1 x N64      context model, two output depth slices per input slice + B-concatenation
// and print k-steps per microsecond per SM and the equivalent algorithmic TFLOP/s of the whole GPU.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../imgcomp_cvpr_b200/csrc -I ../../include \
//        -o mma_shapes mma_shapes.cu && ./mma_shapes [iterations per launch, default 20000]
// Run each pattern long enough (>= 50 ms) for the power cap to settle; read nvidia-smi clocks beside it.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "tc_ptx.cuh"

using namespace ic::tc;

constexpr int kTile = 128 * 16 * 2;          // one M128 x K16 fp16 operand tile = 4 KB (also N128 x K16)
constexpr int kBufs = 8;                     // operand tiles cycled through (mimics the stage ring)

__device__ __forceinline__ uint32_t idesc(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) mma_pattern_kernel(int pattern, int iters, unsigned long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // [A hi x kBufs][A lo x kBufs][B (hi | lo interleaved per 16-byte row group: N up to 256) x kBufs]
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + kBufs * kTile;
    uint8_t* b = a_lo + kBufs * kTile;                      // 8 KB per buffer: rows 0..127 = w_hi, 128..255 = w_lo
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (2 * kBufs * kTile + kBufs * 2 * kTile) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u ^ ((i * 2654435761u) & 0x03ff03ffu);     // fp16 values near 1
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        // K-major no-swizzle: 8 rows x 16 B core matrices; LBO = stride between the two 16-byte K chunks, SBO = 128 B
        const uint32_t lbo_a = 128 * 16, lbo_b = 256 * 16;
        const unsigned long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int s = it % kBufs;
            const uint64_t ah = make_desc(smem_u32(a_hi + s * kTile), lbo_a, 128);
            const uint64_t al = make_desc(smem_u32(a_lo + s * kTile), lbo_a, 128);
            const uint64_t bh = make_desc(smem_u32(b + s * 2 * kTile), lbo_b, 128);                 // rows 0.. (hi first)
            const uint64_t bl = make_desc(smem_u32(b + s * 2 * kTile) + 128 * 16, lbo_b, 128);      // rows 128.. (lo)
            const uint32_t acc = it ? 1u : 0u;
            switch (pattern) {
                case 0:
                    umma_f16(tmem, ah, bh, idesc(128), acc);
                    umma_f16(tmem, ah, bl, idesc(128), 1u);
                    umma_f16(tmem, al, bh, idesc(128), 1u);
                    break;
                case 1:
                    umma_f16(tmem, ah, bh, idesc(256), acc);
                    umma_f16(tmem + 128, al, bh, idesc(128), 1u);
                    break;
                case 2:
                    umma_f16(tmem, ah, bh, idesc(128), acc);
                    break;
                case 3:
                    umma_f16(tmem, ah, bh, idesc(32), acc);
                    umma_f16(tmem, ah, bl, idesc(32), 1u);
                    umma_f16(tmem, al, bh, idesc(32), 1u);
                    break;
                case 4:
                    umma_f16(tmem, ah, bh, idesc(64), acc);
                    umma_f16(tmem + 32, al, bh, idesc(32), 1u);
                    break;
                default:
                    umma_f16(tmem, ah, bh, idesc(128), acc);
                    umma_f16(tmem + 64, al, bh, idesc(64), 1u);
                    break;
            }
            if ((it & 63) == 63) {                       // bound the number of MMAs in flight like a stage hand-over does
                umma_commit(smem_u32(&bar));
                mbar_wait(smem_u32(&bar), (it >> 6) & 1);
            }
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), (iters >> 6) & 1);          // iters / 64 commits so far: this is the next phase
        cycles[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = 2 * kBufs * kTile + kBufs * 2 * kTile + 1024;
    cudaFuncSetAttribute(mma_pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned long long* d_cycles;
    cudaMalloc(&d_cycles, sizeof(unsigned long long) * sms);
    // algorithmic MACs per k-step: one (pixels x cout x 16) product; patterns 3-5: cout = 24 of 32 (x2 slices for 5)
    const double macs[6] = {128.0 * 128 * 16, 128.0 * 128 * 16, 128.0 * 128 * 16, 128.0 * 24 * 16, 128.0 * 24 * 16, 2 * 128.0 * 24 * 16};
    const char* names[6] = {"exact 3 x N128 (today)", "exact N256 + N128 (B-concat)", "fast 1 x N128", "pc 3 x N32 (today)",
                            "pc N64 + N32 (B-concat)", "pc N128 + N64 (2 slices + B-concat)"};
    for (int p = 0; p < 6; ++p) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        mma_pattern_kernel<<<sms, 128, smem>>>(p, iters / 10, d_cycles);      // warm-up
        cudaEventRecord(e0);
        for (int r = 0; r < 5; ++r) mma_pattern_kernel<<<sms, 128, smem>>>(p, iters, d_cycles);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) {
            printf("pattern %d: %s\n", p, cudaGetErrorString(err));
            return 1;
        }
        const double ksteps = 5.0 * iters;
        const double us = ms * 1e3;
        printf("%-38s %8.3f k-steps/us/SM   %7.1f algorithmic TFLOP/s (%d SMs)   %.2f ms\n", names[p], ksteps / us,
               2.0 * macs[p] * ksteps * sms / (us * 1e-6) / 1e12, sms, ms);
    }
    cudaFree(d_cycles);
    return 0;
}
