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

template <int P>
__device__ __forceinline__ void issue_pattern(uint32_t tmem, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, uint32_t acc) {
    if (P == 0) {
        umma_f16(tmem, ah, bh, idesc(128), acc);
        umma_f16(tmem, ah, bl, idesc(128), 1u);
        umma_f16(tmem, al, bh, idesc(128), 1u);
    } else if (P == 1) {
        umma_f16(tmem, ah, bh, idesc(256), acc);
        umma_f16(tmem + 128, al, bh, idesc(128), 1u);
    } else if (P == 2) {
        umma_f16(tmem, ah, bh, idesc(128), acc);
    } else if (P == 3) {
        umma_f16(tmem, ah, bh, idesc(32), acc);
        umma_f16(tmem, ah, bl, idesc(32), 1u);
        umma_f16(tmem, al, bh, idesc(32), 1u);
    } else if (P == 4) {
        umma_f16(tmem, ah, bh, idesc(64), acc);
        umma_f16(tmem + 32, al, bh, idesc(32), 1u);
    } else {
        umma_f16(tmem, ah, bh, idesc(128), acc);
        umma_f16(tmem + 64, al, bh, idesc(64), 1u);
    }
}

// OCC: CTAs per SM (2: half the operand buffers and 256 TMEM columns each).  chains: accumulator regions the k-steps rotate
// through (1: every MMA accumulates onto the columns the previous one wrote; 2: two independent chains in one CTA).
// amode: layout of the A operand.  0: dense tile (core matrices 128-byte aligned, SBO = 128).  1 / 2: as conv_tc.cu reads it out
// of a halo tile -- 8-pixel core matrices at the pitch of a 10- / 18-pixel halo row (SBO = 160 / 288 bytes) and a start address
// that moves by the tap shift (dx 16 bytes + dy one halo row), i.e. core matrices that straddle 128-byte lines.  3: dense
// pitch, start shifted by dx 16 bytes only.
template <int P, int OCC = 1>
__global__ void __launch_bounds__(128, OCC) mma_pattern_kernel(int iters, unsigned long long* cycles, int amode, int chains) {
    constexpr int kBufs = OCC == 2 ? 4 : ::kBufs;
    constexpr int kCols = OCC == 2 ? 256 : 512;
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
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        // K-major no-swizzle: 8 rows x 16 B core matrices; LBO = stride between the two 16-byte K chunks, SBO = 128 B.
        // Descriptors of buffer s = base + s * constant: the issue loop is integer adds + MMAs, like conv_tc.cu's.
        const uint32_t sbo_a = amode == 1 ? 160 : (amode == 2 ? 288 : 128);
        const uint32_t lbo_a = amode == 1 ? 18 * 160 : (amode == 2 ? 18 * 288 : 128 * 16), lbo_b = 256 * 16;
        const uint64_t ah0 = make_desc(smem_u32(a_hi), lbo_a, sbo_a), al0 = make_desc(smem_u32(a_lo), lbo_a, sbo_a);
        const uint64_t bh0 = make_desc(smem_u32(b), lbo_b, 128), bl0 = make_desc(smem_u32(b) + 128 * 16, lbo_b, 128);
        const unsigned long long t0 = clock64();
        uint32_t ph = 0;
        for (int it = 0; it < iters; it += 32) {
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                const uint64_t sb = (uint64_t)((u % kBufs) * (2 * kTile >> 4));
                const uint64_t sa = amode == 0 ? (uint64_t)((u % kBufs) * (kTile >> 4))
                                               : (uint64_t)((u % kBufs) * (1024 >> 4) + (u % 3) + (amode == 3 ? 0 : ((u / 3) % 3) * (sbo_a >> 4)));
                issue_pattern<P>(tmem + (uint32_t)(u & (chains - 1)) * (kCols / 2), ah0 + sa, al0 + sa, bh0 + sb, bl0 + sb, (it | (u >> (chains - 1))) ? 1u : 0u);
            }
            umma_commit(smem_u32(&bar));          // bound the number of MMAs in flight like a stage hand-over does
            mbar_wait(smem_u32(&bar), ph);
            ph ^= 1;
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kCols));
}

// ---- the same question for 2-CTA pairs (cta_group::2: M = 256 over both CTAs, each CTA holds half of the B rows)
//   6  3 x (M256 N128 K16)                today's IC_CONV_PAIR kernel
//   7  1 x (M256 N256 K16) + 1 x N128     pair + B-concatenation (each CTA reads A 4 KB + B 4 KB, then A 4 KB + B 2 KB)
//   8  1 x (M256 N128 K16)                pair, FAST mode
__device__ __forceinline__ uint32_t idesc2(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((256u >> 4) << 24);
}

__device__ __forceinline__ void commit_2sm_leader(uint32_t bar_local) {      // completion of the pair's MMAs -> the leader's barrier only
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_local),
                 "h"((uint16_t)1)
                 : "memory");
}

template <int P>
__device__ __forceinline__ void issue_pair(uint32_t tmem, uint64_t ah, uint64_t al, uint64_t b128, uint64_t b128l, uint64_t b256, uint32_t acc) {
    if (P == 6) {
        umma_f16_2sm(tmem, ah, b128, idesc2(128), acc);
        umma_f16_2sm(tmem, ah, b128l, idesc2(128), 1u);
        umma_f16_2sm(tmem, al, b128, idesc2(128), 1u);
    } else if (P == 7) {
        umma_f16_2sm(tmem, ah, b256, idesc2(256), acc);
        umma_f16_2sm(tmem + 128, al, b128l, idesc2(128), 1u);
    } else {
        umma_f16_2sm(tmem, ah, b128, idesc2(128), acc);
    }
}

template <int P>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
mma_pattern_pair_kernel(int iters, unsigned long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + kBufs * kTile;
    uint8_t* b = a_lo + kBufs * kTile;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < (2 * kBufs * kTile + kBufs * 2 * kTile) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u ^ ((i * 2654435761u) & 0x03ff03ffu);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0 && rank == 0) {
        // per CTA: A = its own 128 rows; B = half of the N rows.  N128: 64 rows per CTA (LBO 64*16); N256: 128 rows
        const uint64_t ah0 = make_desc(smem_u32(a_hi), 128 * 16, 128), al0 = make_desc(smem_u32(a_lo), 128 * 16, 128);
        const uint64_t b128_0 = make_desc(smem_u32(b), 64 * 16, 128), b128l_0 = make_desc(smem_u32(b) + 2 * 64 * 16, 64 * 16, 128);
        const uint64_t b256_0 = make_desc(smem_u32(b), 128 * 16, 128);
        const unsigned long long t0 = clock64();
        uint32_t ph = 0;
        for (int it = 0; it < iters; it += 32) {
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                const uint64_t sa = (uint64_t)((u % kBufs) * (kTile >> 4)), sb = (uint64_t)((u % kBufs) * (2 * kTile >> 4));
                issue_pair<P>(tmem, ah0 + sa, al0 + sa, b128_0 + sb, b128l_0 + sb, b256_0 + sb, (it | u) ? 1u : 0u);
            }
            commit_2sm_leader(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), ph);
            ph ^= 1;
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int P, int OCC = 1>
void launch_pattern(int grid, size_t smem, int iters, unsigned long long* d_cycles, int amode = 0, int chains = 1) {
    if (P >= 6) {
        cudaFuncSetAttribute(mma_pattern_pair_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mma_pattern_pair_kernel<P><<<grid, 128, smem>>>(iters, d_cycles);
    } else {
        const size_t sm = OCC == 2 ? (2 * 4 * kTile + 4 * 2 * kTile + 1024) : smem;
        cudaFuncSetAttribute(mma_pattern_kernel<P, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        mma_pattern_kernel<P, OCC><<<grid * OCC, 128, sm>>>(iters, d_cycles, amode, chains);
    }
}

void launch_any(int p, int grid, size_t smem, int iters, unsigned long long* d) {
    switch (p) {
        case 0: launch_pattern<0>(grid, smem, iters, d); break;
        case 1: launch_pattern<1>(grid, smem, iters, d); break;
        case 2: launch_pattern<2>(grid, smem, iters, d); break;
        case 3: launch_pattern<3>(grid, smem, iters, d); break;
        case 4: launch_pattern<4>(grid, smem, iters, d); break;
        case 5: launch_pattern<5>(grid, smem, iters, d); break;
        case 6: launch_pattern<6>(grid, smem, iters, d); break;
        case 7: launch_pattern<7>(grid, smem, iters, d); break;
        case 8: launch_pattern<8>(grid, smem, iters, d); break;
        case 9: launch_pattern<4>(grid, smem, iters, d, 1); break;
        case 10: launch_pattern<5>(grid, smem, iters, d, 1); break;
        case 11: launch_pattern<0>(grid, smem, iters, d, 2); break;
        case 12: launch_pattern<4>(grid, smem, iters, d, 3); break;
        case 13: launch_pattern<2>(grid, smem, iters, d, 2); break;
        case 14: launch_pattern<5, 2>(grid, smem, iters, d, 1); break;
        case 15: launch_pattern<5>(grid, smem, iters, d, 1, 2); break;
        case 16: launch_pattern<4, 2>(grid, smem, iters, d, 1); break;
        case 17: launch_pattern<4>(grid, smem, iters, d, 1, 2); break;
        case 18: launch_pattern<0>(grid, smem, iters, d, 2, 2); break;
        default: launch_pattern<5, 2>(grid, smem, iters, d, 1, 2); break;
    }
}

int main(int argc, char** argv) {
    const int iters = (argc > 1 ? atoi(argv[1]) : 20000) / 32 * 32;
    const int only = argc > 2 ? atoi(argv[2]) : -1;
    const int reps = argc > 3 ? atoi(argv[3]) : 20;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = 2 * kBufs * kTile + kBufs * 2 * kTile + 1024;
    unsigned long long* d_cycles;
    cudaMalloc(&d_cycles, sizeof(unsigned long long) * sms * 2);
    // algorithmic MACs per k-step: one (pixels x cout x 16) product; patterns 3-5: cout = 24 of 32 (x2 slices for 5)
    const int occ[20] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 2, 1, 1, 2};
    const double macs[20] = {128.0 * 128 * 16, 128.0 * 128 * 16, 128.0 * 128 * 16, 128.0 * 24 * 16, 128.0 * 24 * 16, 2 * 128.0 * 24 * 16,
                             128.0 * 128 * 16, 128.0 * 128 * 16, 128.0 * 128 * 16, 128.0 * 24 * 16, 2 * 128.0 * 24 * 16, 128.0 * 128 * 16,
                             128.0 * 24 * 16, 128.0 * 128 * 16, 2 * 128.0 * 24 * 16, 2 * 128.0 * 24 * 16, 128.0 * 24 * 16, 128.0 * 24 * 16,
                             128.0 * 128 * 16, 2 * 128.0 * 24 * 16};
    const char* names[20] = {"exact 3 x N128 (today)", "exact N256 + N128 (B-concat)", "fast 1 x N128", "pc 3 x N32 (today)",
                             "pc N64 + N32 (B-concat)", "pc N128 + N64 (2 slices + B-concat)",
                             "pair exact 3 x M256 N128", "pair exact M256 N256 + N128 (B-concat)", "pair fast 1 x M256 N128",
                             "pc N64 + N32, A from a 10-px halo", "pc N128 + N64, A from a 10-px halo", "exact 3 x N128, A from an 18-px halo",
                             "pc N64 + N32, dense A shifted by dx", "fast 1 x N128, A from an 18-px halo",
                             "pc N128 + N64 halo A, TWO CTAs per SM", "pc N128 + N64 halo A, two chains in one CTA",
                             "pc N64 + N32 halo A, TWO CTAs per SM", "pc N64 + N32 halo A, two chains in one CTA",
                             "exact 3 x N128 halo A, two chains in one CTA", "pc N128 + N64 halo A, two CTAs x two chains"};
    for (int p = 0; p < 20; ++p) {
        if (only >= 0 && p != only) continue;
        const bool pair = p >= 6;
        const int grid = pair ? (sms / 2) * 2 : sms;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int r = 0; r < reps / 2; ++r) launch_any(p, grid, smem, iters, d_cycles);      // warm-up until the power cap has settled
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) launch_any(p, grid, smem, iters, d_cycles);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) {
            printf("pattern %d: %s\n", p, cudaGetErrorString(err));
            return 1;
        }
        unsigned long long cyc = 0;
        cudaMemcpy(&cyc, d_cycles, sizeof(cyc), cudaMemcpyDeviceToHost);
        const double ksteps = (double)reps * iters * occ[p];      // per SM
        const double us = ms * 1e3;
        printf("%-46s %7.3f k-steps/us/SM  %6.1f CTA-cycles/k-step (last launch)  %7.1f algorithmic TFLOP/s (%d SMs)  %.0f ms\n", names[p],
               ksteps / us, (double)cyc / iters, 2.0 * macs[p] * ksteps * grid / (us * 1e-6) / 1e12, grid, ms);
        fflush(stdout);
    }
    cudaFree(d_cycles);
    return 0;
}
