// Training step, tensor-core leg: the 64 3x3 128->128 stride-1 SAME convolutions of the residual trunks
// (code/autoencoder.py:225-233,253-261,274-287), forward AND data gradient, on the tcgen05 kernel of conv_tc.cu in
// EXACT (fp16 hi/lo, three MMAs per product, fp32 TMEM accumulation) arithmetic.
//   forward        y  = conv3x3(x,  W)                       W  float32 HWIO [3][3][cin][cout]
//   data gradient  dx = conv3x3(dy, W')  with  W'[ky][kx][co][ci] = W[2-ky][2-kx][ci][co]
// (the gradient of a stride-1 SAME 3x3 convolution is the same convolution with flipped, transposed taps).
// Weights change every step, so the kernel's stage-ordered fp16 hi/lo weight image is re-packed ON THE DEVICE per
// call (148 K weights: two tiny kernels); activations go fp32 NHWC -> hi/lo planes -> conv -> planes -> fp32 NHWC
// because the batch-norm kernels around the conv work on float32 NHWC.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace ic {

namespace {

constexpr int kC = 128, kTaps = 9, kW = kTaps * kC * kC;       // 147 456 weights
constexpr int kMaxBlocks = 1024;
constexpr int kPlaneElems = 4 * kC * 8;                        // one plane of one stage: [4 chunks][128 rows][8 cin]
constexpr int kStages = (kC / 32) * kTaps;                     // 36

__global__ void __launch_bounds__(256) maxabs_partial_kernel(const float* __restrict__ w, int64_t n, float* __restrict__ partial) {
    __shared__ float red[8];
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
        partial[blockIdx.x] = m;
    }
}

// power of two that brings the largest magnitude into [2^(t-1), 2^t): hi AND lo parts of the fp16 split then stay normal
// fp16 numbers for everything within ~2^-10 of the maximum (t = 8 for weights: the rule of tc::pack_weights; t = 6 for
// activations / gradients, so that the conv OUTPUT, which stays in the activation's scale until it is merged back to
// float32, cannot overflow fp16); multiplying by a power of two is exact
__device__ __forceinline__ float pow2_scale(float mx, int t) {
    if (!(mx > 0.f) || !isfinite(mx)) return 1.f;
    int ex;
    frexpf(mx, &ex);
    return ldexpf(1.f, t - ex);
}

// params[k] = scale of tensor k (k < ntens; tensor 0 = the weights), params[4 + k] = its inverse;
// scale[c] = 1 / params[0] (epilogue of the conv: the output keeps the activation's scale), shift = 0
struct ScaleArgs {
    const float* partial[3];
    int count[3];
    int ntens;
};

__global__ void __launch_bounds__(256) finalize_scales_kernel(ScaleArgs a, float* __restrict__ params, float* __restrict__ scale,
                                                              float* __restrict__ shift, const float* __restrict__ kept = nullptr,
                                                              int both = 0) {
    __shared__ float red[256];
    __shared__ float sc[3];
    if (kept && threadIdx.x == 0) {        // forward-pass scales of the weights (tensor 0) and of x (tensor 2)
        sc[0] = kept[0];
        params[0] = kept[0];
        params[4] = kept[2];
        params[2] = kept[1];
        params[6] = kept[3];
    }
    for (int k = kept ? 1 : 0; k < a.ntens; ++k) {
        float m = 0.f;
        for (int i = threadIdx.x; i < a.count[k]; i += 256) m = fmaxf(m, a.partial[k][i]);
        red[threadIdx.x] = m;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            sc[k] = pow2_scale(red[0], k == 0 ? 8 : 6);
            params[k] = sc[k];
            params[4 + k] = 1.f / sc[k];
        }
        __syncthreads();
    }
    if (scale && threadIdx.x < kC) {
        // powers of two: exact.  both: the conv writes float32 itself, so it also undoes the input's pre-scale
        scale[threadIdx.x] = both ? (1.f / sc[0]) * (1.f / sc[1]) : 1.f / sc[0];
        shift[threadIdx.x] = 0.f;
    }
}

// Same stage order as tc::pack_weights (k = 3, stride 1): stage = (cin quarter q, tap), within a stage
// [plane hi|lo][4 chunks][128 cout rows][8 cin]; values pre-scaled by sc[0] (a power of two).
// pair = 1: the layout of tc::repack_pair, [stage][cout half][plane][4 chunks][64 rows][8 cin] (CTA-pair kernel, cta_group::2)
__device__ __forceinline__ void pack3x3_elem(const float* __restrict__ w, int data_grad, const float* __restrict__ sc,
                                             __half* __restrict__ packed, float* __restrict__ scale_out, float* __restrict__ shift_out,
                                             int pair, int i) {
    if (scale_out && i < kC) {             // epilogue of the conv: undo the weight pre-scale (a power of two: exact)
        scale_out[i] = 1.f / sc[0];
        shift_out[i] = 0.f;
    }
    if (i >= kStages * kPlaneElems) return;
    const int ei = i & 7, co = (i >> 3) & (kC - 1), ch = (i >> 10) & 3, s = i >> 12;
    const int q = s / kTaps, tap = s - q * kTaps;
    const int ci = q * 32 + ch * 8 + ei;                       // input channel of THIS convolution, co its output channel
    float v;
    if (!data_grad) {
        v = w[((size_t)tap * kC + ci) * kC + co];
    } else {
        v = w[((size_t)(kTaps - 1 - tap) * kC + co) * kC + ci];   // flipped tap, transposed channels
    }
    v *= sc[0];
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    if (pair) {
        const size_t base = (size_t)s * 2 * kPlaneElems + ((((size_t)(co >> 6) * 2 + 0) * 4 + ch) * 64 + (co & 63)) * 8 + ei;
        packed[base] = hi;
        packed[base + 4 * 64 * 8] = lo;       // next plane of the same half
        return;
    }
    const size_t base = (size_t)s * 2 * kPlaneElems + (size_t)(ch * kC + co) * 8 + ei;
    packed[base] = hi;
    packed[base + kPlaneElems] = lo;
}

__global__ void __launch_bounds__(256) pack3x3_kernel(const float* __restrict__ w, int data_grad, const float* __restrict__ sc,
                                                      __half* __restrict__ packed, float* __restrict__ scale_out = nullptr,
                                                      float* __restrict__ shift_out = nullptr, int pair = 0) {
    pack3x3_elem(w, data_grad, sc, packed, scale_out, shift_out, pair, blockIdx.x * blockDim.x + threadIdx.x);
}

// "prepared" weights of one trunk conv in one direction: the packed fp16 image followed by the epilogue's scale / shift vectors
constexpr size_t kPackedBytes = (size_t)kStages * 2 * kPlaneElems * sizeof(__half);
constexpr size_t kPreparedBytes = kPackedBytes + 2 * kC * sizeof(float);

// every trunk conv of the step, both directions, in ONE launch (the weights do not change between the forward and the
// backward pass of a step): blockIdx.y = layer * 2 + direction
__global__ void __launch_bounds__(256) pack3x3_all_kernel(const float* __restrict__ base, const int64_t* __restrict__ offsets,
                                                          const float* __restrict__ scales /* [n][4] */, uint8_t* __restrict__ prepared,
                                                          int pair) {
    const int layer = blockIdx.y >> 1, dir = blockIdx.y & 1;
    uint8_t* dst = prepared + (size_t)blockIdx.y * kPreparedBytes;
    float* sc = reinterpret_cast<float*>(dst + kPackedBytes);
    pack3x3_elem(base + offsets[layer], dir, scales + 4 * layer, reinterpret_cast<__half*>(dst), sc, sc + kC, pair,
                 blockIdx.x * blockDim.x + threadIdx.x);
}

// ------------------------------------------------------------------ filter gradient on tcgen05
// dW[ky][kx][ci][co] = sum over pixels p of x[p + (ky-1, kx-1)][ci] * dy[p][co]: a GEMM whose reduction runs over PIXELS.
// In the [chunk][row][pixel][8 channels] shared-memory tiles that TMA delivers from the NC/8HW8 planes, 8 channels are
// contiguous (16 B) and consecutive pixels are 16 B apart: that is the UMMA *MN-major* SWIZZLE_NONE canonical layout
// ((1,n),(8,k)):((X,SBO),(1,LBO)) in 16-byte units (cute/atom/mma_traits_sm100.hpp) with SBO = chunk pitch (M or N
// direction: channels) and LBO = 128 B (the next 8 pixels of the K direction), so both operands are used as they lie,
// and a filter tap is again a 16-byte-granular shift of A's start address.  One MMA (M = 128 cin, N = 128 cout, K = 16)
// consumes 16 consecutive pixels of one image row.
//   grid = 3 * S CTAs: CTA (ky, split) accumulates the three taps (ky, 0..2) -- 3 x 128 TMEM columns -- over the
//   4-row x 16-column pixel tiles split, split + S, ... and writes one float32 partial [3][128][128]; a fixed-order
//   reduction over the S splits follows (no atomics).  EXACT arithmetic: x_hi*dy_hi + x_hi*dy_lo + x_lo*dy_hi.
constexpr int WG_ROWS = 4, WG_COLS = 16, WG_STAGES = 3;
constexpr int WG_X_PLANE = 16 * WG_ROWS * (WG_COLS + 2) * 16;     // [16 chunks][4 rows][18 px][8] fp16 = 18 432 B
constexpr int WG_Y_PLANE = 16 * WG_ROWS * WG_COLS * 16;           // [16 chunks][4 rows][16 px][8] fp16 = 16 384 B
constexpr int WG_STAGE_BYTES = 2 * WG_X_PLANE + 2 * WG_Y_PLANE;   // hi + lo of both = 69 632 B
constexpr int WG_THREADS = 6 * 32;
constexpr uint32_t WG_IDESC = (1u << 4) /* D fp32 */ | (1u << 15) /* A MN-major */ | (1u << 16) /* B MN-major */ |
                              ((uint32_t)(kC >> 3) << 17) /* N */ | ((128u >> 4) << 24) /* M */;

struct __align__(8) WgBars {
    uint64_t full[WG_STAGES], empty[WG_STAGES], acc_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap dy_map, int N, int H, int W,
                int n_splits, float* __restrict__ partial) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    WgBars* bars = reinterpret_cast<WgBars*>(smem + WG_STAGES * WG_STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ky = blockIdx.x % 3, split = blockIdx.x / 3;
    const int tiles_x = (W + WG_COLS - 1) / WG_COLS, tiles_y = (H + WG_ROWS - 1) / WG_ROWS;
    const int n_tiles = N * tiles_y * tiles_x;
    const int my_tiles = split < n_tiles ? (n_tiles - split + n_splits - 1) / n_splits : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_STAGES; ++i) {
            mbar_init(smem_u32(&bars->full[i]), 1);
            mbar_init(smem_u32(&bars->empty[i]), 1);
        }
        mbar_init(smem_u32(&bars->acc_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < my_tiles; ++it) {
                const int tile = split + it * n_splits;
                const int n = tile / (tiles_y * tiles_x), r = tile - n * tiles_y * tiles_x;
                const int y0 = (r / tiles_x) * WG_ROWS, x0 = (r % tiles_x) * WG_COLS;
                const uint32_t slot = it % WG_STAGES, ph = (it / WG_STAGES) & 1;
                mbar_wait(smem_u32(&bars->empty[slot]), ph ^ 1);
                const uint32_t full = smem_u32(&bars->full[slot]);
                mbar_expect_tx(full, WG_STAGE_BYTES);
                const uint32_t xs = smem_u32(smem + slot * WG_STAGE_BYTES), ys = xs + 2 * WG_X_PLANE;
                for (int pl = 0; pl < 2; ++pl) {
                    // rows y0+ky-1 .. +3 and columns x0-1 .. x0+16 of the input: out-of-image elements read as zero (SAME)
                    tma_load_5d(xs + pl * WG_X_PLANE, &x_map, full, (x0 - 1) * 8, y0 + ky - 1, 0, n, pl);
                    tma_load_5d(ys + pl * WG_Y_PLANE, &dy_map, full, x0 * 8, y0, 0, n, pl);
                }
            }
        }
    } else if (warp == 1) {
        // the whole warp walks the loop, one elected lane issues: warp-uniform control flow keeps the descriptors in uniform
        // registers (see the MMA issuer of conv_tc.cu)
        {
            constexpr uint32_t kXSbo = WG_ROWS * (WG_COLS + 2) * 16, kYSbo = WG_ROWS * WG_COLS * 16, kLbo = 128;
            for (int it = 0; it < my_tiles; ++it) {
                const uint32_t slot = it % WG_STAGES, ph = (it / WG_STAGES) & 1;
                mbar_wait(smem_u32(&bars->full[slot]), ph);
                tc_fence_after();
                const uint32_t xs = smem_u32(smem + slot * WG_STAGE_BYTES), ys = xs + 2 * WG_X_PLANE;
                if (elect_one()) {
#pragma unroll
                for (int r = 0; r < WG_ROWS; ++r) {
                    const uint64_t b_hi = make_desc(ys + r * WG_COLS * 16, kLbo, kYSbo);
                    const uint64_t b_lo = make_desc(ys + WG_Y_PLANE + r * WG_COLS * 16, kLbo, kYSbo);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const uint32_t a_addr = xs + (r * (WG_COLS + 2) + kx) * 16;
                        const uint64_t a_hi = make_desc(a_addr, kLbo, kXSbo);
                        const uint64_t a_lo = make_desc(a_addr + WG_X_PLANE, kLbo, kXSbo);
                        const uint32_t d = tmem_base + kx * kC;
                        umma_f16(d, a_hi, b_hi, WG_IDESC, (it | r) == 0 ? 0u : 1u);
                        umma_f16(d, a_hi, b_lo, WG_IDESC, 1u);
                        umma_f16(d, a_lo, b_hi, WG_IDESC, 1u);
                    }
                }
                umma_commit(smem_u32(&bars->empty[slot]));
                if (it == my_tiles - 1) umma_commit(smem_u32(&bars->acc_full));
                }
                __syncwarp();
            }
        }
    } else {
        // epilogue: TMEM lane = cin row, columns = (kx, cout)
        const int q = warp & 3;
        const int ci = q * 32 + lane;
        if (my_tiles > 0) {
            mbar_wait(smem_u32(&bars->acc_full), 0);
            tc_fence_after();
        }
        for (int kx = 0; kx < 3; ++kx) {
            float* dst = partial + (((size_t)split * kTaps + ky * 3 + kx) * kC + ci) * kC;
            for (int cc = 0; cc < kC / 16; ++cc) {
                uint32_t rr[16];
                if (my_tiles > 0) {
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + kx * kC + cc * 16, rr);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) rr[e] = 0u;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    reinterpret_cast<float4*>(dst + cc * 16)[e] = make_float4(__uint_as_float(rr[4 * e]), __uint_as_float(rr[4 * e + 1]),
                                                                              __uint_as_float(rr[4 * e + 2]), __uint_as_float(rr[4 * e + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

// ---- the same GEMM over pixels for the other layers (strided convs through their space-to-depth form, context model):
// a CTA "kind" fixes the tap row ky, the 128-channel block of A (first chunk) and the image offset of A (context model:
// filter depth 1 reads the next depth slice); it accumulates the taps (ky, 0..2) in 3 x NB TMEM columns.  A has up to 128
// channels (missing chunks are TMA zero fill: M stays 128), B has NB = 32 or 128 columns.
constexpr int WG_MAX_KINDS = 8;
struct WgGenParams {
    int N, Ho, Wo;                 // B (output-gradient) grid: N images of Ho x Wo
    int ox, oy;                    // A tile origin relative to the B tile: -1 (SAME 3x3 on the s2d grid) or 0 (VALID)
    int nkinds, n_splits;
    int ky[WG_MAX_KINDS], chunk0[WG_MAX_KINDS], img_off[WG_MAX_KINDS];
};

template <int NB>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_gen_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap dy_map, const WgGenParams p,
                    float* __restrict__ partial) {
    using namespace tc;
    constexpr int Y_PLANE = (NB / 8) * WG_ROWS * WG_COLS * 16;
    constexpr int STAGE_BYTES = 2 * WG_X_PLANE + 2 * Y_PLANE;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t TCOLS = NB == 128 ? 512 : 128;
    extern __shared__ __align__(1024) uint8_t smem[];
    WgBars* bars = reinterpret_cast<WgBars*>(smem + WG_STAGES * STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kind = blockIdx.x % p.nkinds, split = blockIdx.x / p.nkinds;
    const int ky = p.ky[kind], chunk0 = p.chunk0[kind], img_off = p.img_off[kind];
    const int tiles_x = (p.Wo + WG_COLS - 1) / WG_COLS, tiles_y = (p.Ho + WG_ROWS - 1) / WG_ROWS;
    const int n_tiles = p.N * tiles_y * tiles_x;
    const int my_tiles = split < n_tiles ? (n_tiles - split + p.n_splits - 1) / p.n_splits : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_STAGES; ++i) {
            mbar_init(smem_u32(&bars->full[i]), 1);
            mbar_init(smem_u32(&bars->empty[i]), 1);
        }
        mbar_init(smem_u32(&bars->acc_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < my_tiles; ++it) {
                const int tile = split + it * p.n_splits;
                const int n = tile / (tiles_y * tiles_x), r = tile - n * tiles_y * tiles_x;
                const int y0 = (r / tiles_x) * WG_ROWS, x0 = (r % tiles_x) * WG_COLS;
                const uint32_t slot = it % WG_STAGES, ph = (it / WG_STAGES) & 1;
                mbar_wait(smem_u32(&bars->empty[slot]), ph ^ 1);
                const uint32_t full = smem_u32(&bars->full[slot]);
                mbar_expect_tx(full, STAGE_BYTES);
                const uint32_t xs = smem_u32(smem + slot * STAGE_BYTES), ys = xs + 2 * WG_X_PLANE;
                for (int pl = 0; pl < 2; ++pl) {
                    tma_load_5d(xs + pl * WG_X_PLANE, &x_map, full, (x0 + p.ox) * 8, y0 + ky + p.oy, chunk0, n + img_off, pl);
                    tma_load_5d(ys + pl * Y_PLANE, &dy_map, full, x0 * 8, y0, 0, n, pl);
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t kXSbo = WG_ROWS * (WG_COLS + 2) * 16, kYSbo = WG_ROWS * WG_COLS * 16, kLbo = 128;
        for (int it = 0; it < my_tiles; ++it) {
            const uint32_t slot = it % WG_STAGES, ph = (it / WG_STAGES) & 1;
            mbar_wait(smem_u32(&bars->full[slot]), ph);
            tc_fence_after();
            const uint32_t xs = smem_u32(smem + slot * STAGE_BYTES), ys = xs + 2 * WG_X_PLANE;
            if (elect_one()) {
#pragma unroll
                for (int r = 0; r < WG_ROWS; ++r) {
                    const uint64_t b_hi = make_desc(ys + r * WG_COLS * 16, kLbo, kYSbo);
                    const uint64_t b_lo = make_desc(ys + Y_PLANE + r * WG_COLS * 16, kLbo, kYSbo);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const uint32_t a_addr = xs + (r * (WG_COLS + 2) + kx) * 16;
                        const uint64_t a_hi = make_desc(a_addr, kLbo, kXSbo);
                        const uint64_t a_lo = make_desc(a_addr + WG_X_PLANE, kLbo, kXSbo);
                        const uint32_t d = tmem_base + kx * NB;
                        umma_f16(d, a_hi, b_hi, IDESC, (it | r) == 0 ? 0u : 1u);
                        umma_f16(d, a_hi, b_lo, IDESC, 1u);
                        umma_f16(d, a_lo, b_hi, IDESC, 1u);
                    }
                }
                umma_commit(smem_u32(&bars->empty[slot]));
                if (it == my_tiles - 1) umma_commit(smem_u32(&bars->acc_full));
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int ci = q * 32 + lane;
        if (my_tiles > 0) {
            mbar_wait(smem_u32(&bars->acc_full), 0);
            tc_fence_after();
        }
        for (int kx = 0; kx < 3; ++kx) {
            float* dst = partial + ((((size_t)split * p.nkinds + kind) * 3 + kx) * 128 + ci) * NB;
            for (int cc = 0; cc < NB / 16; ++cc) {
                uint32_t rr[16];
                if (my_tiles > 0) {
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + kx * NB + cc * 16, rr);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) rr[e] = 0u;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    reinterpret_cast<float4*>(dst + cc * 16)[e] = make_float4(__uint_as_float(rr[4 * e]), __uint_as_float(rr[4 * e + 1]),
                                                                              __uint_as_float(rr[4 * e + 2]), __uint_as_float(rr[4 * e + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS));
}

// dw[i] = (sum over splits, fixed order, of partial[split][map[i]]) / (scale_a * scale_b); map[i] < 0: no gradient (masked tap)
__global__ void __launch_bounds__(256) wgrad_gather_kernel(const float* __restrict__ partial, int n_splits, int64_t split_stride,
                                                           const int* __restrict__ map, int n, const float* __restrict__ params,
                                                           float* __restrict__ dw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int m = map[i];
    float s = 0.f;
    if (m >= 0)
        for (int k = 0; k < n_splits; ++k) s += partial[(size_t)k * split_stride + m];
    dw[i] = m >= 0 ? s / (params[0] * params[1]) : 0.f;
}

// dW = (sum over splits, fixed order) / (scale_x * scale_dy)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int n_splits, const float* __restrict__ params,
                                                           float* __restrict__ dw, const float* __restrict__ scale_b = nullptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kW) return;
    float s = 0.f;
    for (int k = 0; k < n_splits; ++k) s += partial[(size_t)k * kW + i];
    dw[i] = s / (params[0] * (scale_b ? scale_b[0] : params[1]));       // the two operand scales (adjacent, or from two places)
}

// ---- fused forward of a trunk layer (training): the conv reads the fp16 planes the previous batch-norm kernel wrote
// (no maximum search, no split pass), and the pass that merges its output planes to float32 also accumulates the batch-norm
// statistics of that output, in exactly the grouping of col_partial_kernel (train_ops.cu): 64 rows per block, a thread adds
// 8 rows (r0 + g, r0 + g + 8, ...) in float32 around the channel's row-0 value, the 8 row groups of a block are added in
// double -> the same mean / variance bits as the unfused path.
__global__ void __launch_bounds__(256) weight_scales_kernel(const float* __restrict__ base, const int64_t* __restrict__ offsets,
                                                            int64_t count, float* __restrict__ scales /* [n][4] */) {
    __shared__ float red[8];
    const float* w = base + offsets[blockIdx.x];
    float m = 0.f;
    for (int64_t i = threadIdx.x; i < count; i += 256) m = fmaxf(m, fabsf(w[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
        const float sc = pow2_scale(m, 8);
        float* o = scales + 4 * blockIdx.x;        // {scale_w, scale_x, 1/scale_w, 1/scale_x}: the layout conv3x3_tc_bwd_ex takes
        o[0] = sc;
        o[1] = 1.f;
        o[2] = 1.f / sc;
        o[3] = 1.f;
    }
}

// params / epilogue scale of a backward pass whose three scales are all known: kept = {s_w, s_x, 1/s_w, 1/s_x} (forward
// pass), dys = {s_dy, 1/s_dy} (ic_nn_bn_train_bwd_ex).  Same slots as finalize_scales_kernel: [0] w, [1] dy, [2] x, [4 + k] inverses.
__global__ void __launch_bounds__(128) set_params_kernel(const float* __restrict__ kept, const float* __restrict__ dys,
                                                         float* __restrict__ params, float* __restrict__ scale, float* __restrict__ shift) {
    if (threadIdx.x == 0) {
        params[0] = kept[0];
        params[4] = kept[2];
        params[2] = kept[1];
        params[6] = kept[3];
        params[1] = dys[0];
        params[5] = dys[1];
    }
    scale[threadIdx.x] = 1.f / kept[0];
    shift[threadIdx.x] = 0.f;
}

constexpr int MS_ROWS = 64;       // = CR_ROWS of train_ops.cu

__global__ void __launch_bounds__(128) merge_stats_kernel(const __half* __restrict__ in, int64_t hw, int64_t M, int64_t plane,
                                                          float* __restrict__ out, double* __restrict__ partial /* [blocks][128][2] */) {
    __shared__ float red[8][128][2];
    const int g = threadIdx.x >> 4, chunk = threadIdx.x & 15;       // row group (a warp of col_partial_kernel), 8-channel chunk
    const int64_t r0 = (int64_t)blockIdx.x * MS_ROWS;
    float piv[8];
    {
        const float4 h4 = *reinterpret_cast<const float4*>(in + (size_t)chunk * hw * 8);          // row 0 = image 0, pixel 0
        const float4 l4 = *reinterpret_cast<const float4*>(in + plane + (size_t)chunk * hw * 8);
        const __half2* h2 = reinterpret_cast<const __half2*>(&h4);
        const __half2* l2 = reinterpret_cast<const __half2*>(&l4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 a = __half22float2(h2[e]), b = __half22float2(l2[e]);
            piv[2 * e] = a.x + b.x;
            piv[2 * e + 1] = a.y + b.y;
        }
    }
    float4 hv[MS_ROWS / 8], lv[MS_ROWS / 8];
#pragma unroll
    for (int k = 0; k < MS_ROWS / 8; ++k) {
        const int64_t r = r0 + g + 8 * k;
        if (r < M) {
            const int64_t n = r / hw, rr = r - n * hw;
            const size_t off = (((size_t)n * 16 + chunk) * hw + rr) * 8;
            hv[k] = *reinterpret_cast<const float4*>(in + off);
            lv[k] = *reinterpret_cast<const float4*>(in + plane + off);
        }
    }
    float s1[8], s2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
#pragma unroll
    for (int k = 0; k < MS_ROWS / 8; ++k) {
        const int64_t r = r0 + g + 8 * k;
        if (r < M) {
            const __half2* h2 = reinterpret_cast<const __half2*>(&hv[k]);
            const __half2* l2 = reinterpret_cast<const __half2*>(&lv[k]);
            float v[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 a = __half22float2(h2[e]), b = __half22float2(l2[e]);
                v[2 * e] = a.x + b.x;
                v[2 * e + 1] = a.y + b.y;
            }
            float4* dst = reinterpret_cast<float4*>(out + ((size_t)r * 16 + chunk) * 8);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float q1 = v[e] - piv[e];
                const float q2 = q1 * q1;
                s1[e] += q1;
                s2[e] += q2;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        red[g][chunk * 8 + e][0] = s1[e];
        red[g][chunk * 8 + e][1] = s2[e];
    }
    __syncthreads();
    const int c = threadIdx.x;
    double t1 = 0, t2 = 0;
    for (int w = 0; w < 8; ++w) {
        t1 += (double)red[w][c][0];
        t2 += (double)red[w][c][1];
    }
    partial[((int64_t)blockIdx.x * 128 + c) * 2 + 0] = t1;
    partial[((int64_t)blockIdx.x * 128 + c) * 2 + 1] = t2;
}


// ---- tile-transposing forms of the planes -> float32 NHWC passes (C = 128; see bn_apply_planes_tt_kernel in train_ops.cu):
// phase 1 reads the planes with a warp per chunk (32 pixels = 512 contiguous bytes), phase 2 writes float32 with a warp per
// pixel (512 contiguous bytes).  A block moves 64 rows (two tiles of 32): warp g owns the rows r0 + g + 8 k, k = 0..7, in
// ascending order -- col_partial_kernel's grouping, so the batch-norm partial sums keep their bits.
constexpr int TT_PITCH = 264, TT_SMEM_FLOATS = 16 * TT_PITCH;

template <bool STATS>
__global__ void __launch_bounds__(256) merge_tt_kernel(const __half* __restrict__ in, int64_t hw, int64_t M, int64_t plane,
                                                       float* __restrict__ out, const float* __restrict__ mul,
                                                       const float* __restrict__ add, double* __restrict__ partial) {
    __shared__ __align__(16) float tile[2][TT_SMEM_FLOATS];
    __shared__ float red[STATS ? 8 : 1][128][2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * 64;
    const float g = mul ? mul[0] : 1.f;
    // phase 1: both tiles (64 rows x 16 chunks = 1024 (chunk, row) items, 4 per thread)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int item = threadIdx.x + 256 * j, tl = item >> 9, c = (item >> 5) & 15, p = item & 31;
        const int64_t r = r0 + tl * 32 + p;
        if (r < M) {
            const int64_t n = r / hw, rr = r - n * hw;
            const size_t off = (((size_t)n * 16 + c) * hw + rr) * 8;
            const float4 h4 = *reinterpret_cast<const float4*>(in + off);
            const float4 l4 = *reinterpret_cast<const float4*>(in + plane + off);
            const __half2* h2 = reinterpret_cast<const __half2*>(&h4);
            const __half2* l2 = reinterpret_cast<const __half2*>(&l4);
            float v[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 a = __half22float2(h2[e]), b = __half22float2(l2[e]);
                v[2 * e] = a.x + b.x;
                v[2 * e + 1] = a.y + b.y;
            }
            float* t = tile[tl] + c * TT_PITCH + p * 8;
            *reinterpret_cast<float4*>(t) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(t + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    __syncthreads();
    // phase 2: warp w writes rows r0 + w + 8 k (k ascending); lane = channels 4 lane .. 4 lane + 3
    float piv[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (STATS) {        // pivot = the tensor's row 0 (image 0, pixel 0), as col_partial_kernel's mode 0
        const size_t off = ((size_t)(lane >> 1) * hw) * 8 + (lane & 1) * 4;
        const uint2 hq = *reinterpret_cast<const uint2*>(in + off), lq = *reinterpret_cast<const uint2*>(in + plane + off);
        const __half2* h2 = reinterpret_cast<const __half2*>(&hq);
        const __half2* l2 = reinterpret_cast<const __half2*>(&lq);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float2 a = __half22float2(h2[e]), b = __half22float2(l2[e]);
            piv[2 * e] = a.x + b.x;
            piv[2 * e + 1] = a.y + b.y;
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int row = w + 8 * k, tl = row >> 5, p = row & 31;
        const int64_t r = r0 + row;
        if (r < M) {
            const float4 q = *reinterpret_cast<const float4*>(tile[tl] + (lane >> 1) * TT_PITCH + p * 8 + (lane & 1) * 4);
            float v[4] = {q.x, q.y, q.z, q.w};
            if (STATS) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float q1 = v[u] - piv[u];
                    const float q2 = q1 * q1;
                    s1[u] += q1;
                    s2[u] += q2;
                }
            }
            if (mul) {
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] *= g;
            }
            if (add) {
                const float4 a = reinterpret_cast<const float4*>(add)[r * 32 + lane];
                v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
            }
            reinterpret_cast<float4*>(out)[r * 32 + lane] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    if (STATS) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            red[w][4 * lane + u][0] = s1[u];
            red[w][4 * lane + u][1] = s2[u];
        }
        __syncthreads();
        if (threadIdx.x < 128) {
            const int c = threadIdx.x;
            double t1 = 0, t2 = 0;
            for (int ww = 0; ww < 8; ++ww) {
                t1 += (double)red[ww][c][0];
                t2 += (double)red[ww][c][1];
            }
            partial[((int64_t)blockIdx.x * 128 + c) * 2 + 0] = t1;
            partial[((int64_t)blockIdx.x * 128 + c) * 2 + 1] = t2;
        }
    }
}

bool use_tt() {
    static const bool tt = !(getenv("IC_TRAIN_TT") && atoi(getenv("IC_TRAIN_TT")) == 0);
    return tt;
}

tc::GroupTable g_gt;
bool g_gt_ready = false;

int ensure_group_table() {
    if (g_gt_ready) return IC_OK;
    std::vector<float> ones((size_t)kW, 1.f);
    std::vector<__half> packed;
    float inv;
    int rc = tc::pack_weights(ones.data(), 3, 1, kC, kC, kC, packed, g_gt, &inv);
    IC_REQUIRE(rc == IC_OK && g_gt.nstages == kStages, IC_ERR_STATE, "train_tc: group table");
    g_gt_ready = true;
    return IC_OK;
}

int maxabs(const float* p, int64_t n, float* partial, int* nb, cudaStream_t s) {
    *nb = (int)std::min<int64_t>(kMaxBlocks, (n + 2047) / 2048);
    maxabs_partial_kernel<<<*nb, 256, 0, s>>>(p, n, partial);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int wgrad_splits(int N, int H, int W) {
    const int n_tiles = N * ((H + WG_ROWS - 1) / WG_ROWS) * ((W + WG_COLS - 1) / WG_COLS);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return std::max(1, std::min(sms / 3, n_tiles));
}

// CTA pairs (cta_group::2) like the inference path (api.cu): IC_CONV_PAIR=0 selects the single-CTA kernel
bool use_pair(int W) {
    const char* e = getenv("IC_CONV_PAIR");
    return W > 8 && !(e && atoi(e) == 0);
}

// conv over planes `in` (already split and pre-scaled) with weights d_w scaled by params[0] -> float32 NHWC, multiplied
// by unscale[0] (the inverse of the input's pre-scale) when the planes are merged back
int conv_planes(const __half* in, const float* d_w, int data_grad, const float* params, const float* unscale, const float* scale,
                const float* shift, __half* wp, __half* bo, int N, int H, int W, float* d_y, cudaStream_t s,
                const float* d_add = nullptr, const void* prepared = nullptr) {
    const bool pair = use_pair(W);
    if (prepared) {          // packed once per step by ic_nn_pack3x3_all
        wp = reinterpret_cast<__half*>(const_cast<void*>(prepared));
        scale = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(prepared) + kPackedBytes);
        shift = scale + kC;
    } else {
        ProfScope ps(IC_PROF_ELEMENTWISE, s);
        pack3x3_kernel<<<cdiv(kStages * kPlaneElems, 256), 256, 0, s>>>(d_w, data_grad, params, wp, nullptr, nullptr, pair ? 1 : 0);
        IC_CHECK_LAUNCH();
    }
    tc::ConvTcArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in;
    a.Nimg = N;
    a.in_chunks = kC / 8;
    a.Hin = H;
    a.Win = W;
    a.weights = wp;
    a.weights_pair = pair ? wp : nullptr;
    a.groups = &g_gt;
    a.scale = scale;
    a.shift = shift;
    a.out = bo;
    a.N = N;
    a.H = H;
    a.W = W;
    a.relu = 0;
    a.cout = kC;
    a.nout = kC;
    a.halo0 = -1;
    a.img_mul = 1;
    a.head = -1;
    a.cpg = 4;
    a.exact = 1;
    a.prof_class = IC_PROF_CONV3X3;
    int rc = tc::launch_conv_tc(a, s);
    if (rc != IC_OK) return rc;
    if (use_tt()) {
        const int64_t M = (int64_t)N * H * W;
        ProfScope ps(IC_PROF_ELEMENTWISE, s);
        merge_tt_kernel<false><<<(unsigned)((M + 63) / 64), 256, 0, s>>>(bo, (int64_t)H * W, M, M * kC, d_y, unscale, d_add, nullptr);
        IC_CHECK_LAUNCH();
        return IC_OK;
    }
    return tc::launch_merge_to_nhwc(bo, N, H, W, kC, d_y, 1, s, unscale, d_add);
}


// ------------------------------------------------------------------ planned tensor-core convs (other layers of the step)
// The stride-2 convs / transposed convs around the residual trunks (h2, h12, h13) and the context-model layers run on the
// same tcgen05 kernels as at inference (conv_tc.cu: space-to-depth groups, depth-to-space columns, resident-weight
// B-concatenated context model), forward AND data gradient:
//     d/dx of conv2d(x, W, stride 2)            = conv2d_transpose(dy, W)        (the same filter array)
//     d/dx of conv2d_transpose(x, W, stride 2)  = conv2d(dy, W, stride 2)        (channel roles swapped)
//     d/dx of the VALID (2,3,3) conv3d          = the FULL correlation with flipped taps, filter depth 1 reading slice s-1
// Weights change every step, so what is fixed per layer is the MAP packed element -> element of the trainer's weight
// buffer.  It is derived from the host packers of conv_tc.cu themselves (no second copy of their layout rules): they
// are run on two synthetic filters whose values are the base-2047 digits (+1) of the source index -- integers <= 2048
// survive the power-of-two scaling and the fp16 hi/lo split exactly -- and the packed digits are read back.
struct TcPlan {
    int kind, data_grad;
    int cin_pad, cout_pad;        // channel strides of the float32 NHWC input / output of THIS conv
    int cout;                     // real output channels
    int nout;                     // kernel columns
    int w_elems;                  // elements of the trainer's weight buffer (maximum search)
    int half;                     // packed layout: element i is a lo part iff (i % (2 half)) >= half, its hi part is i - half
    size_t n_packed;
    tc::GroupTable gt;
    int* d_map;
};

enum { TCK_CONV5S2 = 0, TCK_TCONV5S2 = 1, TCK_TCONV5S2_IMG = 2, TCK_PC = 3, TCK_CONV5S2_F32 = 4 };
__host__ __device__ inline bool tck_strided(int k) { return k == TCK_CONV5S2 || k == TCK_CONV5S2_F32; }

__global__ void __launch_bounds__(256) pack_map_kernel(const float* __restrict__ w, const int* __restrict__ map, int total, int half,
                                                       const float* __restrict__ sc, __half* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const bool is_lo = (i % (2 * half)) >= half;
    const int src = map[is_lo ? i - half : i];
    const float v = src >= 0 ? w[src] * sc[0] : 0.f;
    const __half hi = __float2half_rn(v);
    packed[i] = is_lo ? __float2half_rn(v - __half2float(hi)) : hi;
}

template <class PackFn>
int build_map_by_digits(size_t n_expected, const std::vector<int>& src_of, PackFn pack, std::vector<int>& map, tc::GroupTable& gt) {
    std::vector<float> w(n_expected);
    std::vector<__half> p[2];
    float inv[2];
    for (int pass = 0; pass < 2; ++pass) {
        for (size_t e = 0; e < n_expected; ++e) w[e] = (float)((pass == 0 ? src_of[e] % 2047 : src_of[e] / 2047) + 1);
        int rc = pack(w.data(), p[pass], gt, &inv[pass]);
        if (rc != IC_OK) return rc;
    }
    IC_REQUIRE(p[0].size() == p[1].size(), IC_ERR_STATE, "tc plan: packer sizes differ");
    map.assign(p[0].size(), -1);
    for (size_t i = 0; i < p[0].size(); ++i) {
        const float a = __half2float(p[0][i]) * inv[0], b = __half2float(p[1][i]) * inv[1];
        if (a == 0.f && b == 0.f) continue;
        IC_REQUIRE(a >= 1.f && b >= 1.f && a == floorf(a) && b == floorf(b), IC_ERR_STATE, "tc plan: digit %g / %g at %zu", a, b, i);
        map[i] = ((int)b - 1) * 2047 + ((int)a - 1);
    }
    return IC_OK;
}

// context-model layer, trainer weights [2][3][3][ci_pad][co_pad] ("other" mask: taps (1, fy > 1) and (1, 1, fx > 1) unused).
// forward: input channels = ci, columns = co.  data gradient: input = dy (co channels), columns = ci, taps flipped.
int build_pc_map(int ci, int co, int ci_pad, int co_pad, int nout, int data_grad, std::vector<int>& map, tc::GroupTable& gt) {
    const int cin_e = data_grad ? co : ci, cout_e = data_grad ? ci : co;
    IC_REQUIRE(cin_e <= 32 && cout_e <= nout, IC_ERR_UNSUPPORTED, "tc plan: context-model layer %d -> %d", cin_e, cout_e);
    memset(&gt, 0, sizeof(gt));
    gt.ngroups = 2;
    const size_t plane_elems = (size_t)4 * nout * 8;
    map.clear();
    int nst = 0;
    for (int g = 0; g < 2; ++g) {
        const int fd = g;                           // group 0: filter depth 0 (9 taps), group 1: depth 1 (5 taps)
        gt.img_off[g] = (uint8_t)(data_grad ? 1 - fd : fd);     // data gradient: depth 0 reads slice s, depth 1 slice s - 1
        gt.chunk0[g] = 0;
        int nt = 0;
        for (int ty = 0; ty < 3; ++ty)
            for (int tx = 0; tx < 3; ++tx) {
                const int fy = data_grad ? 2 - ty : ty, fx = data_grad ? 2 - tx : tx;
                if (fd == 1 && (fy > 1 || (fy == 1 && fx > 1))) continue;
                gt.taps[g][nt++] = (uint8_t)(ty * 3 + tx);
                const size_t base = map.size();
                map.resize(base + 2 * plane_elems, -1);
                for (int c = 0; c < cin_e; ++c)
                    for (int o = 0; o < cout_e; ++o) {
                        const int wc = data_grad ? o : c, wo = data_grad ? c : o;      // [ci][co] position in the filter
                        const int src = ((((fd * 3 + fy) * 3 + fx) * ci_pad) + wc) * co_pad + wo;
                        map[base + ((size_t)(c / 8) * 2 * nout + o) * 8 + (c % 8)] = src;
                    }
                ++nst;
            }
        gt.ntaps[g] = (uint8_t)nt;
    }
    gt.nstages = nst;
    gt.eff_ksteps = nst * ((cin_e + 15) / 16);
    return IC_OK;
}

}  // namespace

}  // namespace ic

using namespace ic;

extern "C" {

size_t ic_nn_conv3x3_tc_workspace_bytes(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    const size_t planes = align_up((size_t)N * H * W * kC * 2 * sizeof(__half), 256);
    return 2 * planes + align_up((size_t)kStages * 2 * kPlaneElems * sizeof(__half), 256) + 3 * kMaxBlocks * sizeof(float) + 8192;
}

int ic_nn_conv3x3_tc(const float* d_x, const float* d_w, int N, int H, int W, int data_grad, float* d_y, void* d_workspace,
                     size_t workspace_bytes, void* stream) {
    return ic_nn_conv3x3_tc_ex(d_x, d_w, N, H, W, data_grad, d_y, nullptr, nullptr, d_workspace, workspace_bytes, stream);
}

/* d_x_planes_keep (optional, 2 * N*H*W*128 halves): the fp16 hi/lo planes of the pre-scaled input are written THERE
 * instead of into the workspace, and d_scales_keep (optional, 4 floats) receives {scale_w, scale_x, 1/scale_w, 1/scale_x}:
 * what ic_nn_conv3x3_tc_bwd_ex needs to skip re-deriving them (the weights and x do not change between the forward and
 * the backward pass of one training step). */
int ic_nn_conv3x3_tc_ex(const float* d_x, const float* d_w, int N, int H, int W, int data_grad, float* d_y, void* d_x_planes_keep,
                        float* d_scales_keep, void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(d_x && d_w && d_y && d_workspace, IC_ERR_INVALID, "ic_nn_conv3x3_tc: NULL argument");
    IC_REQUIRE(N > 0 && H > 0 && W > 0, IC_ERR_INVALID, "ic_nn_conv3x3_tc: bad shape");
    IC_REQUIRE(workspace_bytes >= ic_nn_conv3x3_tc_workspace_bytes(N, H, W), IC_ERR_WORKSPACE, "ic_nn_conv3x3_tc: workspace too small");
    int rc = ensure_group_table();
    if (rc != IC_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    Arena ar(d_workspace, workspace_bytes);
    const size_t elems = (size_t)N * H * W * kC * 2;
    __half* bi = ar.get<__half>(elems);
    __half* bo = ar.get<__half>(elems);
    __half* wp = ar.get<__half>((size_t)kStages * 2 * kPlaneElems);
    float* scale = ar.get<float>(kC);
    float* shift = ar.get<float>(kC);
    float* params = ar.get<float>(8);
    float* pw = ar.get<float>(kMaxBlocks);
    float* px = ar.get<float>(kMaxBlocks);
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "ic_nn_conv3x3_tc: workspace too small");
    if (d_x_planes_keep) bi = reinterpret_cast<__half*>(d_x_planes_keep);
    ScaleArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.ntens = 2;
    sa.partial[0] = pw;
    sa.partial[1] = px;
    {
        ProfScope ps(IC_PROF_ELEMENTWISE, s, 3);
        rc = maxabs(d_w, kW, pw, &sa.count[0], s);
        if (rc == IC_OK) rc = maxabs(d_x, (int64_t)N * H * W * kC, px, &sa.count[1], s);
        if (rc != IC_OK) return rc;
        finalize_scales_kernel<<<1, 256, 0, s>>>(sa, params, scale, shift);
        IC_CHECK_LAUNCH();
    }
    rc = tc::launch_split_from_nhwc(d_x, N, H, W, kC, 0, bi, 1, s, params + 1);
    if (rc != IC_OK) return rc;
    if (d_scales_keep) {
        IC_CHECK_CUDA(cudaMemcpyAsync(d_scales_keep, params, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
        IC_CHECK_CUDA(cudaMemcpyAsync(d_scales_keep + 2, params + 4, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    return conv_planes(bi, d_w, data_grad, params, params + 5, scale, shift, wp, bo, N, H, W, d_y, s);
}

size_t ic_nn_conv3x3_tc_bwd_workspace_bytes(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    const size_t planes = align_up((size_t)N * H * W * kC * 2 * sizeof(__half), 256);
    return 3 * planes + align_up((size_t)kStages * 2 * kPlaneElems * sizeof(__half), 256) +
           align_up((size_t)wgrad_splits(N, H, W) * kW * sizeof(float), 256) + 3 * kMaxBlocks * sizeof(float) + 8192;
}

/* backward of y = conv3x3(x, w): d_dx (optional) = data gradient, d_dw = filter gradient [3][3][128][128] */
int ic_nn_conv3x3_tc_bwd(const float* d_x, const float* d_dy, const float* d_w, int N, int H, int W, float* d_dx, float* d_dw,
                         void* d_workspace, size_t workspace_bytes, void* stream) {
    return ic_nn_conv3x3_tc_bwd_ex(d_x, d_dy, d_w, N, H, W, d_dx, d_dw, nullptr, nullptr, d_workspace, workspace_bytes, stream);
}

/* d_x_planes / d_scales (both or neither): what ic_nn_conv3x3_tc_ex kept of the forward pass; d_x may then be NULL */
int ic_nn_conv3x3_tc_bwd_ex(const float* d_x, const float* d_dy, const float* d_w, int N, int H, int W, float* d_dx, float* d_dw,
                            const void* d_x_planes, const float* d_scales, void* d_workspace, size_t workspace_bytes,
                            void* stream) {
    const bool cached = d_x_planes != nullptr && d_scales != nullptr;
    IC_REQUIRE((d_x || cached) && d_dy && d_w && d_dw && d_workspace, IC_ERR_INVALID, "ic_nn_conv3x3_tc_bwd: NULL argument");
    IC_REQUIRE(N > 0 && H > 0 && W > 0, IC_ERR_INVALID, "ic_nn_conv3x3_tc_bwd: bad shape");
    IC_REQUIRE(workspace_bytes >= ic_nn_conv3x3_tc_bwd_workspace_bytes(N, H, W), IC_ERR_WORKSPACE, "ic_nn_conv3x3_tc_bwd: workspace too small");
    int rc = ensure_group_table();
    if (rc != IC_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int S = wgrad_splits(N, H, W);
    Arena ar(d_workspace, workspace_bytes);
    const size_t elems = (size_t)N * H * W * kC * 2;
    __half* bx = ar.get<__half>(elems);
    __half* bdy = ar.get<__half>(elems);
    __half* bo = ar.get<__half>(elems);
    __half* wp = ar.get<__half>((size_t)kStages * 2 * kPlaneElems);
    float* partial = ar.get<float>((size_t)S * kW);
    float* scale = ar.get<float>(kC);
    float* shift = ar.get<float>(kC);
    float* params = ar.get<float>(8);        // [0] weights, [1] dy, [2] x
    float* pw = ar.get<float>(kMaxBlocks);
    float* pdy = ar.get<float>(kMaxBlocks);
    float* px = ar.get<float>(kMaxBlocks);
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "ic_nn_conv3x3_tc_bwd: workspace too small");
    const int64_t n_act = (int64_t)N * H * W * kC;
    ScaleArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.ntens = cached ? 2 : 3;
    sa.partial[0] = pw;
    sa.partial[1] = pdy;
    sa.partial[2] = px;
    {
        ProfScope ps(IC_PROF_ELEMENTWISE, s, cached ? 2 : 4);
        if (cached) {
            // scale_w from the forward pass: a one-element "partial maximum" whose pow2_scale is that scale again
            rc = maxabs(d_dy, n_act, pdy, &sa.count[1], s);
            sa.count[0] = 0;
        } else {
            rc = maxabs(d_w, kW, pw, &sa.count[0], s);
            if (rc == IC_OK) rc = maxabs(d_dy, n_act, pdy, &sa.count[1], s);
            if (rc == IC_OK) rc = maxabs(d_x, n_act, px, &sa.count[2], s);
        }
        if (rc != IC_OK) return rc;
        finalize_scales_kernel<<<1, 256, 0, s>>>(sa, params, scale, shift, cached ? d_scales : nullptr);
        IC_CHECK_LAUNCH();
    }
    rc = tc::launch_split_from_nhwc(d_dy, N, H, W, kC, 0, bdy, 1, s, params + 1);
    if (rc == IC_OK && !cached) rc = tc::launch_split_from_nhwc(d_x, N, H, W, kC, 0, bx, 1, s, params + 2);
    if (rc != IC_OK) return rc;
    if (cached) bx = reinterpret_cast<__half*>(const_cast<void*>(d_x_planes));
    // filter gradient
    CUtensorMap xmap, ymap;
    rc = tc::encode_planes_map(&xmap, bx, 2, N, kC / 8, H, W, WG_COLS + 2, WG_ROWS, kC / 8);
    if (rc == IC_OK) rc = tc::encode_planes_map(&ymap, bdy, 2, N, kC / 8, H, W, WG_COLS, WG_ROWS, kC / 8);
    if (rc != IC_OK) return rc;
    const size_t smem = (size_t)WG_STAGES * WG_STAGE_BYTES + sizeof(WgBars) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        IC_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    {
        ProfScope ps(IC_PROF_CONV3X3, s, 2);
        wgrad_tc_kernel<<<3 * S, WG_THREADS, smem, s>>>(xmap, ymap, N, H, W, S, partial);
        IC_CHECK_LAUNCH();
        wgrad_reduce_kernel<<<cdiv(kW, 256), 256, 0, s>>>(partial, S, params + 1, d_dw);     // / (s_dy * s_x)
        IC_CHECK_LAUNCH();
    }
    if (!d_dx) return IC_OK;
    return conv_planes(bdy, d_w, 1, params, params + 5, scale, shift, wp, bo, N, H, W, d_dx, s);
}


/* ---- planned tensor-core convs of the training step (see TcPlan above).  op_kind: the op of the forward graph
 * (0: conv2d 5x5 stride 2 SAME, 1: conv2d_transpose 5x5 stride 2 SAME, 3: masked (2,3,3) VALID conv3d of the context model on
 * a depth-major volume), data_grad: its data gradient instead.  Weights as the trainer stores them: [5][5][ceil4 Cin][ceil4
 * Cout] in the orientation of the op, [2][3][3][ceil4 Cin][ceil4 Cout] for the context model.  Returns IC_ERR_UNSUPPORTED for
 * shapes the tensor-core kernels do not cover (the caller keeps its FFMA path). */
struct ic_tc_plan {
    TcPlan p;
};

int ic_nn_tc_plan_create(int op_kind, int data_grad, int op_cin, int op_cout, ic_tc_plan_t** out) {
    IC_REQUIRE(out, IC_ERR_INVALID, "ic_nn_tc_plan_create: NULL argument");
    *out = nullptr;
    IC_REQUIRE(op_cin > 0 && op_cout > 0 && (op_kind == 0 || op_kind == 1 || op_kind == 3), IC_ERR_INVALID, "ic_nn_tc_plan_create: bad op");
    const int ci_pad = (int)align_up(op_cin, 4), co_pad = (int)align_up(op_cout, 4);
    TcPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.data_grad = data_grad ? 1 : 0;
    std::vector<int> map;
    int rc = IC_OK;
    if (op_kind == 3) {
        pl.kind = TCK_PC;
        const int cin_e = data_grad ? co_pad : ci_pad, cout_e = data_grad ? op_cin : op_cout;
        if (cin_e % 8 != 0 || cin_e > 32 || cout_e > 32) return IC_ERR_UNSUPPORTED;
        pl.nout = cout_e <= 16 ? 16 : 32;
        pl.cin_pad = cin_e;
        pl.cout_pad = data_grad ? ci_pad : co_pad;
        pl.cout = pl.cout_pad;            // padded channels are written too (zero weights -> zeros)
        pl.w_elems = 18 * ci_pad * co_pad;
        pl.half = pl.nout * 8;
        rc = build_pc_map(op_cin, op_cout, ci_pad, co_pad, pl.nout, pl.data_grad, map, pl.gt);
        if (rc != IC_OK) return rc;
    } else {
        // effective conv: a strided conv (forward of conv2d, data gradient of conv2d_transpose) or a transposed one
        const bool eff_strided = (op_kind == 0) != (data_grad != 0);
        const int cin_e = data_grad ? op_cout : op_cin, cout_e = data_grad ? op_cin : op_cout;
        pl.cin_pad = data_grad ? co_pad : ci_pad;
        pl.cout_pad = data_grad ? ci_pad : co_pad;
        pl.cout = cout_e;
        pl.w_elems = 25 * ci_pad * co_pad;
        if (cin_e % 32 != 0 || pl.cin_pad != cin_e) return IC_ERR_UNSUPPORTED;
        std::vector<int> src((size_t)25 * cin_e * cout_e);
        if (eff_strided) {
            if (cin_e % 64 != 0 || !(cout_e == 128 || cout_e <= 80)) return IC_ERR_UNSUPPORTED;
            // 128 columns: plane output + merge pass; narrower (to_bn: 33 -> 48 columns): the kernel writes float32 NHWC itself,
            // all cout_pad channels (the padding columns have zero weights)
            pl.kind = cout_e == 128 ? TCK_CONV5S2 : TCK_CONV5S2_F32;
            pl.nout = cout_e == 128 ? 128 : (cout_e <= 48 ? 48 : 80);
            if (pl.kind == TCK_CONV5S2_F32) pl.cout = pl.cout_pad;
            for (int t = 0; t < 25; ++t)
                for (int i = 0; i < cin_e; ++i)
                    for (int o = 0; o < cout_e; ++o)      // expected HWIO [t][i][o]
                        src[((size_t)t * cin_e + i) * cout_e + o] = data_grad ? (t * ci_pad + o) * co_pad + i : (t * ci_pad + i) * co_pad + o;
            rc = build_map_by_digits(src.size(), src, [&](const float* w, std::vector<__half>& packed, tc::GroupTable& gt, float* inv) {
                return tc::pack_weights(w, 5, 2, cin_e, cout_e, pl.nout, packed, gt, inv);
            }, map, pl.gt);
        } else {
            if (!(cout_e == 64 || cout_e == 3)) return IC_ERR_UNSUPPORTED;
            pl.kind = cout_e == 3 ? TCK_TCONV5S2_IMG : TCK_TCONV5S2;
            pl.nout = cout_e == 3 ? 16 : 256;
            for (int t = 0; t < 25; ++t)
                for (int o = 0; o < cout_e; ++o)
                    for (int i = 0; i < cin_e; ++i)       // expected [t][cout][cin] (TF conv2d_transpose filter)
                        src[((size_t)t * cout_e + o) * cin_e + i] = data_grad ? (t * ci_pad + o) * co_pad + i : (t * ci_pad + i) * co_pad + o;
            const int all[4] = {0, 1, 2, 3};
            rc = build_map_by_digits(src.size(), src, [&](const float* w, std::vector<__half>& packed, tc::GroupTable& gt, float* inv) {
                return tc::pack_weights_tconv(w, 5, cin_e, cout_e, all, 4, pl.nout, packed, gt, inv);
            }, map, pl.gt);
        }
        if (rc != IC_OK) return rc;
        pl.half = 4 * pl.nout * 8;
    }
    pl.n_packed = map.size();
    IC_CHECK_CUDA(cudaMalloc((void**)&pl.d_map, map.size() * sizeof(int)));
    IC_CHECK_CUDA(cudaMemcpy(pl.d_map, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice));
    ic_tc_plan* h = new ic_tc_plan;
    h->p = pl;
    *out = h;
    return IC_OK;
}

void ic_nn_tc_plan_destroy(ic_tc_plan_t* plan) {
    if (!plan) return;
    cudaFree(plan->p.d_map);
    delete plan;
}

/* host copy of the pack map (tests): n_packed entries, -1 = zero */
int64_t ic_nn_tc_plan_map(const ic_tc_plan_t* plan, int* h_map_out, int64_t capacity) {
    if (!plan) return -1;
    if (h_map_out && capacity >= (int64_t)plan->p.n_packed)
        if (cudaMemcpy(h_map_out, plan->p.d_map, plan->p.n_packed * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)plan->p.n_packed;
}

namespace {
struct PlanGeo {
    int64_t in_imgs, out_imgs;
    int Hi, Wi, Ho, Wo;
};
PlanGeo plan_geo(const TcPlan& p, int D, int N, int H, int W) {
    PlanGeo g;
    if (p.kind == TCK_PC) {
        g.in_imgs = (int64_t)D * N;
        g.out_imgs = (int64_t)(p.data_grad ? D + 1 : D - 1) * N;
        g.Hi = H; g.Wi = W;
        g.Ho = p.data_grad ? H + 2 : H - 2;
        g.Wo = p.data_grad ? W + 2 : W - 2;
    } else {
        g.in_imgs = g.out_imgs = N;
        g.Hi = H; g.Wi = W;
        g.Ho = tck_strided(p.kind) ? H / 2 : 2 * H;
        g.Wo = tck_strided(p.kind) ? W / 2 : 2 * W;
    }
    return g;
}
}  // namespace

size_t ic_nn_tc_plan_workspace_bytes(const ic_tc_plan_t* plan, int D, int N, int H, int W) {
    if (!plan || N <= 0 || H <= 0 || W <= 0) return 0;
    const TcPlan& p = plan->p;
    const PlanGeo g = plan_geo(p, D, N, H, W);
    if (g.out_imgs <= 0 || g.Ho <= 0 || g.Wo <= 0) return 0;
    size_t b = align_up((size_t)g.in_imgs * g.Hi * g.Wi * p.cin_pad * 2 * sizeof(__half), 256);
    if (p.kind == TCK_CONV5S2 || p.kind == TCK_TCONV5S2) b += align_up((size_t)g.out_imgs * g.Ho * g.Wo * p.cout * 2 * sizeof(__half), 256);
    if (p.kind == TCK_TCONV5S2_IMG) b += align_up((size_t)g.out_imgs * g.Ho * g.Wo * 3 * sizeof(float), 256);
    return b + align_up(p.n_packed * sizeof(__half), 256) + 3 * kMaxBlocks * sizeof(float) + 8192;
}

/* d_x: float32 NHWC input of THIS conv (the output gradient when the plan is a data gradient), dims (N, H, W, C) or, context
 * model, (D, N, H, W, C) depth-major; d_y: its float32 NHWC output (context model forward: (D-1, N, H-2, W-2, C'), data
 * gradient: (D+1, N, H+2, W+2, C')). */
int ic_nn_tc_plan_run(const ic_tc_plan_t* plan, const float* d_x, const float* d_w, int D, int N, int H, int W, float* d_y,
                      void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(plan && d_x && d_w && d_y && d_workspace, IC_ERR_INVALID, "ic_nn_tc_plan_run: NULL argument");
    const TcPlan& p = plan->p;
    IC_REQUIRE(N > 0 && H > 0 && W > 0 && (p.kind != TCK_PC || D > (p.data_grad ? 0 : 1)), IC_ERR_INVALID, "ic_nn_tc_plan_run: bad shape");
    IC_REQUIRE(!tck_strided(p.kind) || (H % 2 == 0 && W % 2 == 0), IC_ERR_INVALID, "ic_nn_tc_plan_run: odd size for a stride-2 conv");
    IC_REQUIRE(workspace_bytes >= ic_nn_tc_plan_workspace_bytes(plan, D, N, H, W), IC_ERR_WORKSPACE, "ic_nn_tc_plan_run: workspace too small");
    const PlanGeo g = plan_geo(p, D, N, H, W);
    IC_REQUIRE(g.Ho > 0 && g.Wo > 0 && g.out_imgs > 0, IC_ERR_INVALID, "ic_nn_tc_plan_run: empty output");
    cudaStream_t s = (cudaStream_t)stream;
    Arena ar(d_workspace, workspace_bytes);
    const size_t in_elems = (size_t)g.in_imgs * g.Hi * g.Wi * p.cin_pad;
    __half* bi = ar.get<__half>(2 * in_elems);
    __half* bo = nullptr;
    float* img = nullptr;
    if (p.kind == TCK_CONV5S2 || p.kind == TCK_TCONV5S2) bo = ar.get<__half>((size_t)g.out_imgs * g.Ho * g.Wo * p.cout * 2);
    if (p.kind == TCK_TCONV5S2_IMG) img = ar.get<float>((size_t)g.out_imgs * g.Ho * g.Wo * 3);
    __half* wp = ar.get<__half>(p.n_packed);
    float* scale = ar.get<float>(kC);
    float* shift = ar.get<float>(kC);
    float* params = ar.get<float>(8);
    float* pw = ar.get<float>(kMaxBlocks);
    float* px = ar.get<float>(kMaxBlocks);
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "ic_nn_tc_plan_run: workspace too small");
    const bool f32_out = p.kind == TCK_PC || p.kind == TCK_TCONV5S2_IMG || p.kind == TCK_CONV5S2_F32;
    ScaleArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.ntens = 2;
    sa.partial[0] = pw;
    sa.partial[1] = px;
    int rc;
    {
        ProfScope ps(IC_PROF_ELEMENTWISE, s, 4);
        rc = maxabs(d_w, p.w_elems, pw, &sa.count[0], s);
        if (rc == IC_OK) rc = maxabs(d_x, (int64_t)in_elems, px, &sa.count[1], s);
        if (rc != IC_OK) return rc;
        finalize_scales_kernel<<<1, 256, 0, s>>>(sa, params, scale, shift, nullptr, f32_out ? 1 : 0);
        IC_CHECK_LAUNCH();
        pack_map_kernel<<<cdiv((int64_t)p.n_packed, 256), 256, 0, s>>>(d_w, p.d_map, (int)p.n_packed, p.half, params, wp);
        IC_CHECK_LAUNCH();
    }
    // float32 NHWC (x sx) -> fp16 hi/lo planes; the strided conv reads its input in space-to-depth form
    rc = tc::launch_split_from_nhwc(d_x, (int)g.in_imgs, g.Hi, g.Wi, p.cin_pad, tck_strided(p.kind) ? 1 : 0, bi, 1, s, params + 1);
    if (rc != IC_OK) return rc;
    tc::ConvTcArgs a;
    memset(&a, 0, sizeof(a));
    a.in = bi;
    a.weights = wp;
    a.groups = &p.gt;
    a.scale = scale;
    a.shift = shift;
    a.N = (int)g.out_imgs;
    a.relu = 0;
    a.cout = p.cout;
    a.nout = p.nout;
    a.img_mul = 1;
    a.head = -1;
    a.cpg = 4;
    a.exact = 1;
    a.prof_class = p.kind == TCK_PC ? IC_PROF_PROBCLASS : IC_PROF_CONV_OTHER;
    if (tck_strided(p.kind)) {
        a.Nimg = N;
        a.in_chunks = 4 * p.cin_pad / 8;
        a.Hin = H / 2;
        a.Win = W / 2;
        a.H = g.Ho;
        a.W = g.Wo;
        a.halo0 = -1;
        if (p.kind == TCK_CONV5S2) a.out = bo;
        else a.out_f32 = d_y;
    } else if (p.kind == TCK_PC) {
        a.Nimg = (int)g.in_imgs;
        a.in_chunks = p.cin_pad / 8;
        a.Hin = H;
        a.Win = W;
        a.H = g.Ho;
        a.W = g.Wo;
        a.halo0 = p.data_grad ? -2 : 0;
        a.img_off_mul = N;
        a.img_base = p.data_grad ? -N : 0;
        a.pc_f32 = 1;
        a.pair_c2 = p.cin_pad <= 24 ? 1 : 0;      // chunks past the tensor's channel count are TMA zero fill
        a.out_f32 = d_y;
    } else {
        a.Nimg = N;
        a.in_chunks = p.cin_pad / 8;
        a.Hin = H;
        a.Win = W;
        a.H = H;                 // depth-to-space: the tile grid is the INPUT grid, columns = (output phase, channel)
        a.W = W;
        a.halo0 = -1;
        if (p.kind == TCK_TCONV5S2) {
            a.out = bo;
            a.d2s_cch = p.cout / 8;
            a.d2s_ph0 = 0;
        } else {
            a.out_f32 = img;     // NCHW (N, 3, 2H, 2W), no denormalisation
            a.denorm = 0;
        }
    }
    rc = tc::launch_conv_tc(a, s);
    if (rc != IC_OK) return rc;
    if (bo) return tc::launch_merge_to_nhwc(bo, (int)g.out_imgs, g.Ho, g.Wo, p.cout, d_y, 1, s, params + 5);
    if (img) return ic_nn_nchw_to_nhwc(img, N, 3, p.cout_pad, (int64_t)g.Ho * g.Wo, d_y, stream);
    return IC_OK;
}


/* ---- filter gradients of the same layers on tcgen05 (wgrad_tc_gen_kernel): op_kind as in ic_nn_tc_plan_create.
 * d_x: the op's float32 NHWC input, d_dy: the gradient w.r.t. its output, D/N/H/W: dimensions of d_x;
 * d_dw: the gradient in the layout of the op's weight array (masked taps of the context model: 0). */
struct ic_tc_wgrad_plan {
    int op_kind;
    int a_ch, b_ch;               // channel strides of the A source / B source tensors (float32 NHWC)
    int nb;                       // B columns (32 or 128)
    int a_s2d;                    // A source goes through space-to-depth (strided / transposed convs)
    int a_is_dy;                  // transposed conv: A = output gradient, B = input
    int n_dw;
    WgGenParams gp;               // kinds (N / Ho / Wo / n_splits / img_off scale filled per call)
    int* d_map;
};

int ic_nn_tc_wgrad_plan_create(int op_kind, int op_cin, int op_cout, ic_tc_wgrad_plan_t** out) {
    IC_REQUIRE(out, IC_ERR_INVALID, "ic_nn_tc_wgrad_plan_create: NULL argument");
    *out = nullptr;
    IC_REQUIRE(op_cin > 0 && op_cout > 0 && (op_kind == 0 || op_kind == 1 || op_kind == 3), IC_ERR_INVALID, "ic_nn_tc_wgrad_plan_create: bad op");
    const int ci_pad = (int)align_up(op_cin, 4), co_pad = (int)align_up(op_cout, 4);
    ic_tc_wgrad_plan pl;
    memset(&pl, 0, sizeof(pl));
    pl.op_kind = op_kind;
    std::vector<int> map;
    if (op_kind == 3) {
        if (ci_pad % 8 != 0 || ci_pad > 128 || co_pad % 8 != 0 || co_pad > 32) return IC_ERR_UNSUPPORTED;
        pl.a_ch = ci_pad;
        pl.b_ch = co_pad;
        pl.nb = 32;
        pl.n_dw = 18 * ci_pad * co_pad;
        pl.gp.nkinds = 5;
        for (int k = 0; k < 5; ++k) {
            pl.gp.ky[k] = k % 3;
            pl.gp.chunk0[k] = 0;
            pl.gp.img_off[k] = k / 3;          // x N per call: filter depth 1 reads the next depth slice
        }
        map.assign(pl.n_dw, -1);
        for (int fd = 0; fd < 2; ++fd)
            for (int fy = 0; fy < 3; ++fy)
                for (int fx = 0; fx < 3; ++fx) {
                    if (fd == 1 && (fy > 1 || (fy == 1 && fx > 1))) continue;          // "other" mask (code/probclass.py:164-176)
                    const int kind = fd * 3 + fy;
                    for (int c = 0; c < op_cin; ++c)
                        for (int o = 0; o < op_cout; ++o)
                            map[((((fd * 3 + fy) * 3 + fx) * ci_pad) + c) * co_pad + o] = ((kind * 3 + fx) * 128 + c) * 32 + o;
                }
    } else {
        // A = the 2x finer tensor in space-to-depth form (conv2d: the input; conv2d_transpose: the output gradient)
        const int fine = op_kind == 0 ? op_cin : op_cout, coarse = op_kind == 0 ? op_cout : op_cin;
        if (fine % 32 != 0 || 4 * fine > 256 || coarse != 128 || fine != (op_kind == 0 ? ci_pad : co_pad)) return IC_ERR_UNSUPPORTED;
        const int nhalves = 4 * fine / 128;
        pl.a_ch = fine;
        pl.b_ch = coarse;
        pl.nb = 128;
        pl.a_s2d = 1;
        pl.a_is_dy = op_kind == 1;
        pl.n_dw = 25 * ci_pad * co_pad;
        pl.gp.nkinds = 3 * nhalves;
        pl.gp.ox = pl.gp.oy = -1;
        for (int ty = 0; ty < 3; ++ty)
            for (int h = 0; h < nhalves; ++h) {
                pl.gp.ky[ty * nhalves + h] = ty;
                pl.gp.chunk0[ty * nhalves + h] = h * 16;
            }
        map.assign(pl.n_dw, -1);
        for (int ky = 0; ky < 5; ++ky)
            for (int kx = 0; kx < 5; ++kx) {
                const int ty = ((ky - 1) >> 1) + 1, py = (ky - 1) & 1, tx = ((kx - 1) >> 1) + 1, px = (kx - 1) & 1;
                for (int f = 0; f < fine; ++f)
                    for (int c = 0; c < coarse; ++c) {
                        const int sc = (py * 2 + px) * fine + f;                 // channel of the space-to-depth tensor
                        const int kind = ty * nhalves + sc / 128;
                        const int off = ((kind * 3 + tx) * 128 + sc % 128) * 128 + c;
                        // conv2d weights [t][cin = f][cout = c]; conv2d_transpose (op orientation) [t][cin = c][cout = f]
                        const int wi = op_kind == 0 ? ((ky * 5 + kx) * ci_pad + f) * co_pad + c : ((ky * 5 + kx) * ci_pad + c) * co_pad + f;
                        map[wi] = off;
                    }
            }
    }
    IC_CHECK_CUDA(cudaMalloc((void**)&pl.d_map, map.size() * sizeof(int)));
    IC_CHECK_CUDA(cudaMemcpy(pl.d_map, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice));
    *out = new ic_tc_wgrad_plan(pl);
    return IC_OK;
}

void ic_nn_tc_wgrad_plan_destroy(ic_tc_wgrad_plan_t* plan) {
    if (!plan) return;
    cudaFree(plan->d_map);
    delete plan;
}

namespace {
struct WgGeo {
    int64_t a_imgs, b_imgs;
    int Ha, Wa, Hb, Wb;           // A source / B source spatial sizes (float32 tensors)
    int S;
};
WgGeo wgrad_geo(const ic_tc_wgrad_plan& p, int D, int N, int H, int W) {
    WgGeo g;
    if (p.op_kind == 3) {
        g.a_imgs = (int64_t)D * N; g.b_imgs = (int64_t)(D - 1) * N;
        g.Ha = H; g.Wa = W; g.Hb = H - 2; g.Wb = W - 2;
    } else if (p.op_kind == 0) {
        g.a_imgs = g.b_imgs = N;
        g.Ha = H; g.Wa = W; g.Hb = H / 2; g.Wb = W / 2;
    } else {
        g.a_imgs = g.b_imgs = N;
        g.Ha = 2 * H; g.Wa = 2 * W; g.Hb = H; g.Wb = W;
    }
    const int64_t n_tiles = g.b_imgs * ((g.Hb + WG_ROWS - 1) / WG_ROWS) * ((g.Wb + WG_COLS - 1) / WG_COLS);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    g.S = (int)std::max<int64_t>(1, std::min<int64_t>(sms / p.gp.nkinds, n_tiles));
    return g;
}
}  // namespace

size_t ic_nn_tc_wgrad_plan_workspace_bytes(const ic_tc_wgrad_plan_t* plan, int D, int N, int H, int W) {
    if (!plan || N <= 0 || H <= 0 || W <= 0) return 0;
    const WgGeo g = wgrad_geo(*plan, D, N, H, W);
    if (g.b_imgs <= 0 || g.Hb <= 0 || g.Wb <= 0) return 0;
    return align_up((size_t)g.a_imgs * g.Ha * g.Wa * plan->a_ch * 2 * sizeof(__half), 256) +
           align_up((size_t)g.b_imgs * g.Hb * g.Wb * plan->b_ch * 2 * sizeof(__half), 256) +
           align_up((size_t)g.S * plan->gp.nkinds * 3 * 128 * plan->nb * sizeof(float), 256) + 3 * kMaxBlocks * sizeof(float) + 8192;
}

int ic_nn_tc_wgrad_plan_run(const ic_tc_wgrad_plan_t* plan, const float* d_x, const float* d_dy, int D, int N, int H, int W,
                            float* d_dw, void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(plan && d_x && d_dy && d_dw && d_workspace, IC_ERR_INVALID, "ic_nn_tc_wgrad_plan_run: NULL argument");
    const ic_tc_wgrad_plan& p = *plan;
    IC_REQUIRE(N > 0 && H > 0 && W > 0 && (p.op_kind != 3 || D > 1), IC_ERR_INVALID, "ic_nn_tc_wgrad_plan_run: bad shape");
    IC_REQUIRE(p.op_kind != 0 || (H % 2 == 0 && W % 2 == 0), IC_ERR_INVALID, "ic_nn_tc_wgrad_plan_run: odd size for a stride-2 conv");
    IC_REQUIRE(workspace_bytes >= ic_nn_tc_wgrad_plan_workspace_bytes(plan, D, N, H, W), IC_ERR_WORKSPACE, "ic_nn_tc_wgrad_plan_run: workspace too small");
    const WgGeo g = wgrad_geo(p, D, N, H, W);
    IC_REQUIRE(g.Hb > 0 && g.Wb > 0 && g.b_imgs > 0, IC_ERR_INVALID, "ic_nn_tc_wgrad_plan_run: empty output grid");
    cudaStream_t s = (cudaStream_t)stream;
    const float* a_src = p.a_is_dy ? d_dy : d_x;
    const float* b_src = p.a_is_dy ? d_x : d_dy;
    const size_t a_elems = (size_t)g.a_imgs * g.Ha * g.Wa * p.a_ch, b_elems = (size_t)g.b_imgs * g.Hb * g.Wb * p.b_ch;
    const size_t split_stride = (size_t)p.gp.nkinds * 3 * 128 * p.nb;
    Arena ar(d_workspace, workspace_bytes);
    __half* ba = ar.get<__half>(2 * a_elems);
    __half* bb = ar.get<__half>(2 * b_elems);
    float* partial = ar.get<float>((size_t)g.S * split_stride);
    float* params = ar.get<float>(8);
    float* pa = ar.get<float>(kMaxBlocks);
    float* pb = ar.get<float>(kMaxBlocks);
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "ic_nn_tc_wgrad_plan_run: workspace too small");
    ScaleArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.ntens = 2;
    sa.partial[0] = pa;
    sa.partial[1] = pb;
    int rc;
    {
        ProfScope ps(IC_PROF_ELEMENTWISE, s, 3);
        rc = maxabs(a_src, (int64_t)a_elems, pa, &sa.count[0], s);
        if (rc == IC_OK) rc = maxabs(b_src, (int64_t)b_elems, pb, &sa.count[1], s);
        if (rc != IC_OK) return rc;
        finalize_scales_kernel<<<1, 256, 0, s>>>(sa, params, nullptr, nullptr);
        IC_CHECK_LAUNCH();
    }
    rc = tc::launch_split_from_nhwc(a_src, (int)g.a_imgs, g.Ha, g.Wa, p.a_ch, p.a_s2d, ba, 1, s, params + 0);
    if (rc == IC_OK) rc = tc::launch_split_from_nhwc(b_src, (int)g.b_imgs, g.Hb, g.Wb, p.b_ch, 0, bb, 1, s, params + 1);
    if (rc != IC_OK) return rc;
    // A planes as the kernel sees them: space-to-depth -> [pl][N][4 a_ch / 8][Ha/2][Wa/2][8]
    const int a_chunks = (p.a_s2d ? 4 : 1) * p.a_ch / 8, Hs = p.a_s2d ? g.Ha / 2 : g.Ha, Ws = p.a_s2d ? g.Wa / 2 : g.Wa;
    CUtensorMap xmap, ymap;
    rc = tc::encode_planes_map(&xmap, ba, 2, (int)g.a_imgs, a_chunks, Hs, Ws, WG_COLS + 2, WG_ROWS, 16);
    if (rc == IC_OK) rc = tc::encode_planes_map(&ymap, bb, 2, (int)g.b_imgs, p.b_ch / 8, g.Hb, g.Wb, WG_COLS, WG_ROWS, p.nb / 8);
    if (rc != IC_OK) return rc;
    WgGenParams gp = p.gp;
    gp.N = (int)g.b_imgs;
    gp.Ho = g.Hb;
    gp.Wo = g.Wb;
    gp.n_splits = g.S;
    if (p.op_kind == 3)
        for (int k = 0; k < gp.nkinds; ++k) gp.img_off[k] *= N;
    {
        ProfScope ps(p.op_kind == 3 ? IC_PROF_PROBCLASS : IC_PROF_CONV_OTHER, s, 2);
        if (p.nb == 128) {
            const size_t smem = (size_t)WG_STAGES * (2 * WG_X_PLANE + 2 * 16 * WG_ROWS * WG_COLS * 16) + sizeof(WgBars) + 64;
            static bool attr_set = false;
            if (!attr_set) {
                IC_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_gen_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attr_set = true;
            }
            wgrad_tc_gen_kernel<128><<<gp.nkinds * g.S, WG_THREADS, smem, s>>>(xmap, ymap, gp, partial);
        } else {
            const size_t smem = (size_t)WG_STAGES * (2 * WG_X_PLANE + 2 * 4 * WG_ROWS * WG_COLS * 16) + sizeof(WgBars) + 64;
            static bool attr_set = false;
            if (!attr_set) {
                IC_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_gen_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attr_set = true;
            }
            wgrad_tc_gen_kernel<32><<<gp.nkinds * g.S, WG_THREADS, smem, s>>>(xmap, ymap, gp, partial);
        }
        IC_CHECK_LAUNCH();
        wgrad_gather_kernel<<<cdiv(p.n_dw, 256), 256, 0, s>>>(partial, g.S, (int64_t)split_stride, p.d_map, p.n_dw, params, d_dw);
        IC_CHECK_LAUNCH();
    }
    return IC_OK;
}


/* ---- fused forward of the trunk layers in the training step (see merge_stats_kernel) */
/* scales of n equally sized weight tensors at d_base + d_offsets[i] (count floats each): d_scales[4 i ..] = {2^e, 1, 2^-e, 1} with
 * 2^e the power of two that brings the largest magnitude into [2^7, 2^8) -- the rule of ic_nn_conv3x3_tc; one launch per step */
int ic_nn_weight_scales(const float* d_base, const int64_t* d_offsets, int n, int64_t count, float* d_scales, void* stream) {
    IC_REQUIRE(d_base && d_offsets && d_scales && n > 0 && count > 0, IC_ERR_INVALID, "ic_nn_weight_scales: bad argument");
    ProfScope ps(IC_PROF_ELEMENTWISE, (cudaStream_t)stream);
    weight_scales_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(d_base, d_offsets, count, d_scales);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

size_t ic_nn_conv3x3_tc_fused_workspace_bytes(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    return align_up((size_t)N * H * W * kC * 2 * sizeof(__half), 256) + align_up((size_t)kStages * 2 * kPlaneElems * sizeof(__half), 256) + 8192;
}

size_t ic_nn_bn_partial_bytes(int64_t M) { return M > 0 ? (size_t)((M + MS_ROWS - 1) / MS_ROWS) * kC * 2 * sizeof(double) : 0; }

/* every trunk conv's weights packed for the kernel, forward and data-gradient direction, in one launch per step:
 * d_prepared receives n x 2 entries of ic_nn_conv3x3_tc_prepared_bytes() (entry 2 i = forward of tensor i, 2 i + 1 = its data
 * gradient); W = the width of the images the convs will run on (selects the CTA-pair weight layout like the convs do) */
size_t ic_nn_conv3x3_tc_prepared_bytes(void) { return kPreparedBytes; }

int ic_nn_pack3x3_all(const float* d_base, const int64_t* d_offsets, const float* d_scales, int n, int W, void* d_prepared, void* stream) {
    IC_REQUIRE(d_base && d_offsets && d_scales && d_prepared && n > 0, IC_ERR_INVALID, "ic_nn_pack3x3_all: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    pack3x3_all_kernel<<<dim3(cdiv(kStages * kPlaneElems, 256), 2 * n), 256, 0, s>>>(d_base, d_offsets, d_scales, (uint8_t*)d_prepared,
                                                                                   use_pair(W) ? 1 : 0);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

/* y = conv3x3(x, w) for the 128 -> 128 trunk convs: d_x_planes = the UNSCALED fp16 hi/lo planes of x (what
 * ic_nn_bn_train_fwd_ex wrote), d_wscale4 = this layer's row of ic_nn_weight_scales, d_prepared (optional) = its forward
 * entry of ic_nn_pack3x3_all (else the weights are packed here).  d_bn_partial (ic_nn_bn_partial_bytes(N H W))
 * receives the batch-norm partial sums of y for ic_nn_bn_train_fwd_ex. */
int ic_nn_conv3x3_tc_fused(const void* d_x_planes, const float* d_w, const float* d_wscale4, const void* d_prepared, int N, int H,
                           int W, float* d_y, double* d_bn_partial, void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(d_x_planes && d_w && d_wscale4 && d_y && d_bn_partial && d_workspace, IC_ERR_INVALID, "ic_nn_conv3x3_tc_fused: NULL argument");
    IC_REQUIRE(N > 0 && H > 0 && W > 0, IC_ERR_INVALID, "ic_nn_conv3x3_tc_fused: bad shape");
    IC_REQUIRE(workspace_bytes >= ic_nn_conv3x3_tc_fused_workspace_bytes(N, H, W), IC_ERR_WORKSPACE, "ic_nn_conv3x3_tc_fused: workspace too small");
    int rc = ensure_group_table();
    if (rc != IC_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    Arena ar(d_workspace, workspace_bytes);
    const size_t elems = (size_t)N * H * W * kC;
    __half* bo = ar.get<__half>(2 * elems);
    __half* wp = ar.get<__half>((size_t)kStages * 2 * kPlaneElems);
    float* scale = ar.get<float>(kC);
    float* shift = ar.get<float>(kC);
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "ic_nn_conv3x3_tc_fused: workspace too small");
    const bool pair = use_pair(W);
    if (d_prepared) {
        wp = reinterpret_cast<__half*>(const_cast<void*>(d_prepared));
        scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(const_cast<void*>(d_prepared)) + kPackedBytes);
        shift = scale + kC;
    } else {
        ProfScope ps(IC_PROF_ELEMENTWISE, s);
        pack3x3_kernel<<<cdiv(kStages * kPlaneElems, 256), 256, 0, s>>>(d_w, 0, d_wscale4, wp, scale, shift, pair ? 1 : 0);
        IC_CHECK_LAUNCH();
    }
    tc::ConvTcArgs a;
    memset(&a, 0, sizeof(a));
    a.in = reinterpret_cast<const __half*>(d_x_planes);
    a.Nimg = N;
    a.in_chunks = kC / 8;
    a.Hin = H;
    a.Win = W;
    a.weights = wp;
    a.weights_pair = pair ? wp : nullptr;
    a.groups = &g_gt;
    a.scale = scale;
    a.shift = shift;
    a.out = bo;
    a.N = N;
    a.H = H;
    a.W = W;
    a.cout = kC;
    a.nout = kC;
    a.halo0 = -1;
    a.img_mul = 1;
    a.head = -1;
    a.cpg = 4;
    a.exact = 1;
    a.prof_class = IC_PROF_CONV3X3;
    rc = tc::launch_conv_tc(a, s);
    if (rc != IC_OK) return rc;
    const int64_t M = (int64_t)N * H * W;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    if (use_tt())
        merge_tt_kernel<true><<<(unsigned)((M + 63) / 64), 256, 0, s>>>(bo, (int64_t)H * W, M, (int64_t)elems, d_y, nullptr, nullptr, d_bn_partial);
    else
        merge_stats_kernel<<<(unsigned)((M + MS_ROWS - 1) / MS_ROWS), 128, 0, s>>>(bo, (int64_t)H * W, M, (int64_t)elems, d_y, d_bn_partial);
    IC_CHECK_LAUNCH();
    return IC_OK;
}


/* ic_nn_conv3x3_tc_bwd_ex for an output gradient that already exists as pre-scaled fp16 planes (ic_nn_bn_train_bwd_ex):
 * d_dy_planes with d_dy_scale = {s, 1/s}; d_x_planes / d_scales as in ic_nn_conv3x3_tc_bwd_ex (required).  No maximum
 * search, no split pass.  d_dx_add (optional, same shape as d_dx): d_dx = data gradient + d_dx_add, the gradient the input
 * already received through a residual connection (saves the separate accumulation pass). */
int ic_nn_conv3x3_tc_bwd_planes(const void* d_dy_planes, const float* d_dy_scale, const float* d_w, const void* d_prepared_dgrad,
                                int N, int H, int W, float* d_dx, const float* d_dx_add, float* d_dw, const void* d_x_planes,
                                const float* d_scales, void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(d_dy_planes && d_dy_scale && d_w && d_dw && d_x_planes && d_scales && d_workspace, IC_ERR_INVALID, "ic_nn_conv3x3_tc_bwd_planes: NULL argument");
    IC_REQUIRE(N > 0 && H > 0 && W > 0, IC_ERR_INVALID, "ic_nn_conv3x3_tc_bwd_planes: bad shape");
    IC_REQUIRE(workspace_bytes >= ic_nn_conv3x3_tc_bwd_workspace_bytes(N, H, W), IC_ERR_WORKSPACE, "ic_nn_conv3x3_tc_bwd_planes: workspace too small");
    int rc = ensure_group_table();
    if (rc != IC_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int S = wgrad_splits(N, H, W);
    Arena ar(d_workspace, workspace_bytes);
    const size_t elems = (size_t)N * H * W * kC * 2;
    __half* bo = ar.get<__half>(elems);
    __half* wp = ar.get<__half>((size_t)kStages * 2 * kPlaneElems);
    float* partial = ar.get<float>((size_t)S * kW);
    float* scale = ar.get<float>(kC);
    float* shift = ar.get<float>(kC);
    float* params = ar.get<float>(8);
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "ic_nn_conv3x3_tc_bwd_planes: workspace too small");
    if (!d_prepared_dgrad) {      // with prepared weights every scale is read where it lives: no parameter block to set up
        ProfScope ps(IC_PROF_ELEMENTWISE, s);
        set_params_kernel<<<1, 128, 0, s>>>(d_scales, d_dy_scale, params, scale, shift);
        IC_CHECK_LAUNCH();
    }
    const __half* bx = reinterpret_cast<const __half*>(d_x_planes);
    const __half* bdy = reinterpret_cast<const __half*>(d_dy_planes);
    CUtensorMap xmap, ymap;
    rc = tc::encode_planes_map(&xmap, bx, 2, N, kC / 8, H, W, WG_COLS + 2, WG_ROWS, kC / 8);
    if (rc == IC_OK) rc = tc::encode_planes_map(&ymap, bdy, 2, N, kC / 8, H, W, WG_COLS, WG_ROWS, kC / 8);
    if (rc != IC_OK) return rc;
    const size_t smem = (size_t)WG_STAGES * WG_STAGE_BYTES + sizeof(WgBars) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        IC_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    {
        ProfScope ps(IC_PROF_CONV3X3, s, 2);
        wgrad_tc_kernel<<<3 * S, WG_THREADS, smem, s>>>(xmap, ymap, N, H, W, S, partial);
        IC_CHECK_LAUNCH();
        if (d_prepared_dgrad) wgrad_reduce_kernel<<<cdiv(kW, 256), 256, 0, s>>>(partial, S, d_dy_scale, d_dw, d_scales + 1);
        else wgrad_reduce_kernel<<<cdiv(kW, 256), 256, 0, s>>>(partial, S, params + 1, d_dw);     // / (s_dy * s_x)
        IC_CHECK_LAUNCH();
    }
    if (!d_dx) return IC_OK;
    return conv_planes(bdy, d_w, 1, params, d_prepared_dgrad ? d_dy_scale + 1 : params + 5, scale, shift, wp, bo, N, H, W, d_dx, s, d_dx_add,
                       d_prepared_dgrad);
}

}  // extern "C"
