// h1 of _CVPR._encode (code/autoencoder.py:221-222): _normalize + slim.conv2d 5x5 stride 2, 3 -> 64, BN, ReLU,
// as ONE kernel that starts from the uint8 (or float32) NCHW image.
//
// Why a dedicated kernel.  On the generic grouped-tap kernel (conv_tc.cu) h1 ran as 9 taps x 32 padded channels
// (K = 288 for a real K of 75: 54 MMAs per 128 output pixels) behind a separate normalise / space-to-depth pass that
// wrote and re-read 32 B per input pixel.  Here the patch matrix is built on chip: K = 75 -> 80 (5 k-steps), 10 or 15
// MMAs per tile, HBM traffic = 3 B per input pixel in, 64 B per input pixel out (the 64-channel hi/lo planes): the
// layer is bound by that store stream (604 MB for 24 x 768 x 512).
//
// Per 16 x 8 output tile (M = 128 UMMA rows), warp-specialised and persistent (one CTA per SM):
//   warps 0-7  (thread = output pixel x plane: 128 build the hi operand, 128 the lo operand)
//              load the 35 x 19 x 3 input patch (TF SAME, stride 2, even size: pad_before 1, pad_after 2; zeros
//              outside the image apply to the NORMALISED input), normalise ((x - mean_c) / std_c with the same
//              intrinsics as prep_input_kernel), split into fp16 hi / lo, and gather the im2col tile
//              A[k/8][pixel][k%8], k = (ky*5 + kx)*3 + c, straight into the UMMA K-major no-swizzle layout
//              (core matrix = 8 pixels x 16 B); fence.proxy.async, then arrive on a_full
//   warp 8     tcgen05.mma, B-concatenated like conv_tc.cu's resident-weight kernels: a_hi * [w_hi | w_lo] (N = 128)
//              into TMEM columns [hh | x], a_lo * w_hi (N = 64) into x; weights (20 KB) stay resident in shared memory
//   warps 9-12 tcgen05.ld -> hh + x -> BN scale / shift -> ReLU -> hi / lo split -> 16-byte stores in the
//              space-to-depth order h2 reads ([plane][N][4 phases x 8 chunks][H/4][W/4][8])
// A tiles and accumulators are double buffered, so build / MMA / epilogue of consecutive tiles overlap.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace ic {
namespace tc {

namespace {

constexpr int H1_TH = 16, H1_TW = 8;                        // output tile: 16 rows x 8 pixels = M 128
constexpr int H1_PH = 2 * H1_TH + 3, H1_PW = 2 * H1_TW + 3;  // input patch 35 x 19
constexpr int H1_K = 75, H1_KCH = 10;                        // K = 5*5*3 -> 10 chunks of 8 (5 k-steps)
constexpr int H1_NOUT = 64;
constexpr int H1_A_PLANE = H1_KCH * 128 * 16;                // 20 KB: one plane of one A stage
constexpr int H1_W_BYTES = H1_KCH * 2 * H1_NOUT * 16;        // 20 KB: [10 chunks][hi 64 rows | lo 64 rows][8]
constexpr int H1_BUILD = 256;                               // builder threads: (pixel m, plane) pairs
constexpr int H1_THREADS = H1_BUILD + 5 * 32;
constexpr int H1_PATCH_ELEMS = 4000;                       // 2 planes x 3 x 35 x 19 = 3990 halves, padded to keep the barriers 8-byte aligned

__constant__ float c_h1_mean[3] = {121.853699f, 113.588608f, 100.637154f};
__constant__ float c_h1_std[3] = {68.8939514f, 66.7393417f, 69.3702698f};   // float32 sqrt(var + 1e-10)

struct __align__(8) H1Bars {
    uint64_t a_full[2], a_empty[2], acc_full[2], acc_empty[2], w_full;
    uint32_t tmem_base;
};

struct H1Params {
    const void* x;              // N,3,H,W uint8 or float32
    const uint8_t* weights;     // H1_W_BYTES, packed by pack_weights_h1_im2col
    const float* scale;         // [64] BN scale / weight pre-scale
    const float* shift;         // [64]
    __half* out;                // [2][N][32][H/4][W/4][8]
    int N, H, W;                // input size (even)
    int normalize, relu;
    float acc_gain;
};

template <typename TIn, bool EXACT>
__global__ void __launch_bounds__(H1_THREADS, 1) conv_h1_kernel(const H1Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_buf = smem;                                        // [2 stages][2 planes][H1_A_PLANE]
    uint8_t* w_buf = a_buf + 2 * 2 * H1_A_PLANE;                  // resident weights
    __half* patch = reinterpret_cast<__half*>(w_buf + H1_W_BYTES);   // [2 planes][3][35][19] normalised input, hi / lo
    float* s_scale = reinterpret_cast<float*>(patch + H1_PATCH_ELEMS);
    float* s_shift = s_scale + H1_NOUT;
    H1Bars* bars = reinterpret_cast<H1Bars*>(s_shift + H1_NOUT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Ho = p.H >> 1, Wo = p.W >> 1;
    const int tiles_x = (Wo + H1_TW - 1) / H1_TW, tiles_y = (Ho + H1_TH - 1) / H1_TH;
    const int n_tiles = p.N * tiles_y * tiles_x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars->a_full[i]), H1_BUILD);
            mbar_init(smem_u32(&bars->a_empty[i]), 1);
            mbar_init(smem_u32(&bars->acc_full[i]), 1);
            mbar_init(smem_u32(&bars->acc_empty[i]), 128);
        }
        mbar_init(smem_u32(&bars->w_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < H1_NOUT) {
        s_scale[threadIdx.x] = p.scale[threadIdx.x];
        s_shift[threadIdx.x] = p.shift[threadIdx.x];
    }
    if (warp == H1_BUILD / 32) {   // TMEM: 2 accumulator sets x [hh 64 | x 64]
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < H1_BUILD / 32) {
        // ===================== patch loader + im2col builder =====================
        const int m = threadIdx.x & 127;               // A row = output pixel of the tile
        const int my_pl = threadIdx.x >> 7;            // 0: builds the hi operand, 1: the lo operand
        const int r = m >> 3, q = m & 7;
        const int base = (2 * r) * H1_PW + 2 * q;      // top-left of this pixel's 5x5 window in the patch
        const TIn* x = reinterpret_cast<const TIn*>(p.x);
        const size_t hw = (size_t)p.H * p.W;
        constexpr int kPatch = 3 * H1_PH * H1_PW;                  // 1995 elements
        constexpr int kPerThread = (kPatch + H1_BUILD - 1) / H1_BUILD;   // 8
        // The patch of tile t+1 is fetched into registers (all loads issued back to back) while tile t is gathered:
        // one exposed global-memory latency per tile would cost more than the whole MMA time of the tile.
        TIn regs[kPerThread];
        auto fetch = [&](int tile) {
            const int n = tile / (tiles_y * tiles_x), rr = tile - n * tiles_y * tiles_x;
            const int iy0 = 2 * (rr / tiles_x) * H1_TH - 1, ix0 = 2 * (rr % tiles_x) * H1_TW - 1;
#pragma unroll
            for (int u = 0; u < kPerThread; ++u) {
                const int i = threadIdx.x + u * H1_BUILD;
                const int c = i / (H1_PH * H1_PW), rem = i - c * (H1_PH * H1_PW);
                const int py = rem / H1_PW, px = rem - py * H1_PW;
                const int iy = iy0 + py, ix = ix0 + px;
                const bool ok = i < kPatch && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                regs[u] = ok ? x[((size_t)n * 3 + c) * hw + (size_t)iy * p.W + ix] : (TIn)0;
            }
        };
        // zeros outside the image apply to the NORMALISED input: mark them by coordinates again when converting
        auto stash = [&](int tile) {
            const int rr = tile % (tiles_y * tiles_x);
            const int iy0 = 2 * (rr / tiles_x) * H1_TH - 1, ix0 = 2 * (rr % tiles_x) * H1_TW - 1;
#pragma unroll
            for (int u = 0; u < kPerThread; ++u) {
                const int i = threadIdx.x + u * H1_BUILD;
                if (i >= kPatch) break;
                const int c = i / (H1_PH * H1_PW), rem = i - c * (H1_PH * H1_PW);
                const int py = rem / H1_PW, px = rem - py * H1_PW;
                const int iy = iy0 + py, ix = ix0 + px;
                float v = 0.f;
                if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
                    const float f = (float)regs[u];
                    v = p.normalize ? __fdiv_rn(__fsub_rn(f, c_h1_mean[c]), c_h1_std[c]) : f;
                }
                const __half hi = __float2half_rn(v);
                patch[i] = hi;
                if (EXACT) patch[kPatch + i] = __float2half_rn(v - __half2float(hi));
            }
        };
        uint32_t it = 0;
        if ((int)blockIdx.x < n_tiles) fetch(blockIdx.x);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            stash(tile);                                            // [c][35][19] hi / lo halves in shared memory
            if (tile + (int)gridDim.x < n_tiles) fetch(tile + gridDim.x);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // -- im2col gather into stage `it & 1` once the MMAs of the tile before last have retired
            const uint32_t stage = it & 1;
            mbar_wait(smem_u32(&bars->a_empty[stage]), ((it >> 1) & 1) ^ 1);
            uint8_t* a_st = a_buf + stage * 2 * H1_A_PLANE;
            if (EXACT || my_pl == 0) {
                const int pl = my_pl;
                const __half* src = patch + pl * 3 * H1_PH * H1_PW + base;
#pragma unroll
                for (int j = 0; j < H1_KCH; ++j) {
                    __align__(16) __half v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int k = 8 * j + e;                 // compile-time: every offset below is a constant
                        if (k < H1_K) {
                            const int ky = k / 15, kx = (k % 15) / 3, c = k % 3;
                            v[e] = src[(c * H1_PH + ky) * H1_PW + kx];
                        } else {
                            v[e] = __float2half(0.f);
                        }
                    }
                    *reinterpret_cast<float4*>(a_st + pl * H1_A_PLANE + (j * 128 + m) * 16) = *reinterpret_cast<const float4*>(v);
                }
            }
            // generic-proxy stores -> visible to the tensor core's async-proxy reads
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(smem_u32(&bars->a_full[stage]));
            asm volatile("bar.sync 1, 256;" ::: "memory");       // everyone is done reading the patch
        }
    } else if (warp == H1_BUILD / 32) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            mbar_expect_tx(smem_u32(&bars->w_full), H1_W_BYTES);
            bulk_load(smem_u32(w_buf), p.weights, H1_W_BYTES, smem_u32(&bars->w_full));
        }
        mbar_wait(smem_u32(&bars->w_full), 0);
        tc_fence_after();
        constexpr uint32_t IDESC_N128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint32_t IDESC_N64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint64_t kAKs = (2 * 128 * 16) >> 4, kAPlane = H1_A_PLANE >> 4, kAStage = (2 * H1_A_PLANE) >> 4;
        constexpr uint64_t kWKs = (2 * 2 * H1_NOUT * 16) >> 4;
        const uint64_t a_desc0 = make_desc(smem_u32(a_buf), 128 * 16, 128);
        const uint64_t w_desc0 = make_desc(smem_u32(w_buf), 2 * H1_NOUT * 16, 128);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t stage = it & 1, ph = (it >> 1) & 1;
            mbar_wait(smem_u32(&bars->acc_empty[stage]), ph ^ 1);
            mbar_wait(smem_u32(&bars->a_full[stage]), ph);
            tc_fence_after();
            const uint64_t a_s = a_desc0 + (uint64_t)stage * kAStage;
            const uint32_t d = tmem_base + stage * 128;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < H1_KCH / 2; ++ks) {
                    const uint64_t a_hi = a_s + ks * kAKs, w = w_desc0 + ks * kWKs;
                    if (EXACT) {
                        umma_f16(d, a_hi, w, IDESC_N128, ks == 0 ? 0u : 1u);          // [hh | x]
                        umma_f16(d + H1_NOUT, a_hi + kAPlane, w, IDESC_N64, 1u);       // x += a_lo * w_hi
                    } else {
                        umma_f16(d, a_hi, w, IDESC_N64, ks == 0 ? 0u : 1u);
                    }
                }
                umma_commit(smem_u32(&bars->a_empty[stage]));
                umma_commit(smem_u32(&bars->acc_full[stage]));
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue (warps 9..12) =====================
        const int q4 = warp & 3;
        const int m = q4 * 32 + lane;
        const int ty = m >> 3, tx = m & 7;
        const int H4 = p.H >> 2, W4 = p.W >> 2;
        const size_t plane = (size_t)p.N * 32 * H4 * W4 * 8;
        const size_t chunk_stride = (size_t)H4 * W4 * 8;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t stage = it & 1, ph = (it >> 1) & 1;
            const int n = tile / (tiles_y * tiles_x), rr = tile - n * tiles_y * tiles_x;
            const int y = (rr / tiles_x) * H1_TH + ty, x = (rr % tiles_x) * H1_TW + tx;
            const bool inside = y < Ho && x < Wo;
            mbar_wait(smem_u32(&bars->acc_full[stage]), ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + stage * 128;
            const int phs = (y & 1) * 2 + (x & 1);
            const size_t pix_off = (((size_t)n * 32 + phs * 8) * H4 + (y >> 1)) * W4 * 8 + (size_t)(x >> 1) * 8;
#pragma unroll 1
            for (int cc = 0; cc < H1_NOUT / 16; ++cc) {
                uint32_t rh[16];
                tmem_ld16(taddr + cc * 16, rh);
                if (EXACT) {
                    uint32_t rx[16];
                    tmem_ld16(taddr + H1_NOUT + cc * 16, rx);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) rh[e] = __float_as_uint(fmaf(__uint_as_float(rh[e]), p.acc_gain, __uint_as_float(rx[e])));
                }
                tmem_ld_wait();
                if (inside) {
#pragma unroll
                    for (int hc = 0; hc < 2; ++hc) {
                        const int chunk = cc * 2 + hc;
                        __align__(16) __half2 hi[4];
                        __align__(16) __half2 lo[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float a = fmaf(__uint_as_float(rh[hc * 8 + 2 * e]), s_scale[chunk * 8 + 2 * e], s_shift[chunk * 8 + 2 * e]);
                            float b = fmaf(__uint_as_float(rh[hc * 8 + 2 * e + 1]), s_scale[chunk * 8 + 2 * e + 1], s_shift[chunk * 8 + 2 * e + 1]);
                            if (p.relu) {
                                a = fmaxf(a, 0.f);
                                b = fmaxf(b, 0.f);
                            }
                            hi[e] = __floats2half2_rn(a, b);
                            const float2 hf = __half22float2(hi[e]);
                            lo[e] = __floats2half2_rn(a - hf.x, b - hf.y);
                        }
                        const size_t off = pix_off + (size_t)chunk * chunk_stride;
                        *reinterpret_cast<float4*>(p.out + off) = *reinterpret_cast<const float4*>(hi);
                        if (EXACT) *reinterpret_cast<float4*>(p.out + plane + off) = *reinterpret_cast<const float4*>(lo);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->acc_empty[stage]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == H1_BUILD / 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
}

template <typename TIn, bool EXACT>
int launch_h1_t(const H1Params& p, cudaStream_t s) {
    const size_t smem = 2 * 2 * H1_A_PLANE + H1_W_BYTES + H1_PATCH_ELEMS * sizeof(__half) + 2 * H1_NOUT * sizeof(float) +
                        sizeof(H1Bars) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        IC_CHECK_CUDA(cudaFuncSetAttribute(conv_h1_kernel<TIn, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int Ho = p.H / 2, Wo = p.W / 2;
    const int n_tiles = p.N * ((Ho + H1_TH - 1) / H1_TH) * ((Wo + H1_TW - 1) / H1_TW);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    ProfScope ps(IC_PROF_CONV_OTHER, s);
    conv_h1_kernel<TIn, EXACT><<<n_tiles < sms ? n_tiles : sms, H1_THREADS, smem, s>>>(p);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

}  // namespace

// x: N,3,H,W uint8 / float32 (H, W multiples of 4) -> fp16 hi/lo planes [2][N][32][H/4][W/4][8] = h1's 64 channels at
// H/2 x W/2 in space-to-depth order (FAST mode writes the hi plane only)
int launch_conv_h1(const void* x, int x_is_u8, int N, int H, int W, int normalize, const __half* weights, const float* scale,
                   const float* shift, int relu, __half* out, int exact, cudaStream_t s) {
    IC_REQUIRE(x && weights && scale && shift && out, IC_ERR_INVALID, "conv_h1: NULL argument");
    IC_REQUIRE(N > 0 && H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0, IC_ERR_INVALID, "conv_h1: bad shape %dx%dx%d", N, H, W);
    H1Params p;
    memset(&p, 0, sizeof(p));
    p.x = x;
    p.weights = (const uint8_t*)weights;
    p.scale = scale;
    p.shift = shift;
    p.out = out;
    p.N = N;
    p.H = H;
    p.W = W;
    p.normalize = normalize;
    p.relu = relu;
    const char* gain_env = getenv("IC_TC_ACC_GAIN");
    p.acc_gain = (gain_env && atoi(gain_env) == 0) ? 1.0f : 1.0f + 1.606e-8f * (float)(H1_KCH / 2);
    if (x_is_u8) return exact ? launch_h1_t<uint8_t, true>(p, s) : launch_h1_t<uint8_t, false>(p, s);
    return exact ? launch_h1_t<float, true>(p, s) : launch_h1_t<float, false>(p, s);
}

// conv2d HWIO weights [5][5][3][64] -> [10 chunks of k][hi rows 0..63 | lo rows 64..127][8 k], k = (ky*5 + kx)*3 + c,
// scaled by a power of two (*inv_scale_out undoes it in the BN scale)
int pack_weights_h1_im2col(const float* w_hwio, int cin, int cout, std::vector<__half>& packed, float* inv_scale_out) {
    if (cin != 3 || cout != H1_NOUT) return IC_ERR_UNSUPPORTED;
    float mx = 0.f;
    for (int i = 0; i < 25 * cin * cout; ++i) mx = fmaxf(mx, fabsf(w_hwio[i]));
    int e = 0;
    if (mx > 0.f) {
        int ex;
        frexpf(mx, &ex);
        e = 8 - ex;
    }
    const float sc = ldexpf(1.f, e);
    *inv_scale_out = ldexpf(1.f, -e);
    packed.assign((size_t)H1_KCH * 2 * H1_NOUT * 8, __float2half(0.f));
    for (int k = 0; k < H1_K; ++k)
        for (int o = 0; o < cout; ++o) {
            const float v = w_hwio[(size_t)k * cout + o] * sc;       // HWIO flattened: ((ky*5 + kx)*3 + c)*cout + o
            const __half hi = __float2half_rn(v);
            const __half lo = __float2half_rn(v - __half2float(hi));
            const size_t idx = ((size_t)(k / 8) * 2 * H1_NOUT + o) * 8 + (k % 8);
            packed[idx] = hi;
            packed[idx + (size_t)H1_NOUT * 8] = lo;
        }
    return IC_OK;
}

}  // namespace tc
}  // namespace ic
