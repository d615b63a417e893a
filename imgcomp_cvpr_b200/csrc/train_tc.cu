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
                                                              float* __restrict__ shift, const float* __restrict__ kept = nullptr) {
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
        scale[threadIdx.x] = 1.f / sc[0];                  // powers of two: exact
        shift[threadIdx.x] = 0.f;
    }
}

// Same stage order as tc::pack_weights (k = 3, stride 1): stage = (cin quarter q, tap), within a stage
// [plane hi|lo][4 chunks][128 cout rows][8 cin]; values pre-scaled by sc[0] (a power of two).
__global__ void __launch_bounds__(256) pack3x3_kernel(const float* __restrict__ w, int data_grad, const float* __restrict__ sc,
                                                      __half* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;       // over (stage, chunk, cout row, cin-in-chunk)
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
    const size_t base = (size_t)s * 2 * kPlaneElems + (size_t)(ch * kC + co) * 8 + ei;
    packed[base] = hi;
    packed[base + kPlaneElems] = lo;
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

// dW = (sum over splits, fixed order) / (scale_x * scale_dy)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int n_splits, const float* __restrict__ params,
                                                           float* __restrict__ dw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kW) return;
    float s = 0.f;
    for (int k = 0; k < n_splits; ++k) s += partial[(size_t)k * kW + i];
    dw[i] = s / (params[0] * params[1]);
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

// conv over planes `in` (already split and pre-scaled) with weights d_w scaled by params[0] -> float32 NHWC, multiplied
// by unscale[0] (the inverse of the input's pre-scale) when the planes are merged back
int conv_planes(const __half* in, const float* d_w, int data_grad, const float* params, const float* unscale, const float* scale,
                const float* shift, __half* wp, __half* bo, int N, int H, int W, float* d_y, cudaStream_t s) {
    {
        ProfScope ps(IC_PROF_ELEMENTWISE, s);
        pack3x3_kernel<<<cdiv(kStages * kPlaneElems, 256), 256, 0, s>>>(d_w, data_grad, params, wp);
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
    return tc::launch_merge_to_nhwc(bo, N, H, W, kC, d_y, 1, s, unscale);
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

}  // extern "C"
