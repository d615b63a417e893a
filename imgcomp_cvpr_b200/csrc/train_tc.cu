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

namespace ic {

namespace {

constexpr int kC = 128, kTaps = 9, kW = kTaps * kC * kC;       // 147 456 weights
constexpr int kMaxBlocks = 64;
constexpr int kPlaneElems = 4 * kC * 8;                        // one plane of one stage: [4 chunks][128 rows][8 cin]
constexpr int kStages = (kC / 32) * kTaps;                     // 36

__global__ void __launch_bounds__(256) maxabs_partial_kernel(const float* __restrict__ w, int n, float* __restrict__ partial) {
    __shared__ float red[8];
    float m = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
        partial[blockIdx.x] = m;
    }
}

// Same stage order and scaling rule as tc::pack_weights (k = 3, stride 1): stage = (cin quarter q, tap), within a stage
// [plane hi|lo][4 chunks][128 cout rows][8 cin]; values scaled by 2^e with the scaled maximum in [2^7, 2^8) so that the lo
// part stays a normal fp16 number; the inverse scale goes into the epilogue's per-channel scale.
__global__ void __launch_bounds__(256) pack3x3_kernel(const float* __restrict__ w, int data_grad, const float* __restrict__ partial,
                                                      int nblocks, __half* __restrict__ packed, float* __restrict__ scale,
                                                      float* __restrict__ shift) {
    float mx = 0.f;
    for (int i = 0; i < nblocks; ++i) mx = fmaxf(mx, partial[i]);
    int e = 0;
    if (mx > 0.f) {
        int ex;
        frexpf(mx, &ex);
        e = 8 - ex;
    }
    const float sc = ldexpf(1.f, e);
    if (blockIdx.x == 0 && threadIdx.x < kC) {
        scale[threadIdx.x] = ldexpf(1.f, -e);
        shift[threadIdx.x] = 0.f;
    }
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
    v *= sc;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const size_t base = (size_t)s * 2 * kPlaneElems + (size_t)(ch * kC + co) * 8 + ei;
    packed[base] = hi;
    packed[base + kPlaneElems] = lo;
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

}  // namespace

}  // namespace ic

using namespace ic;

extern "C" {

size_t ic_nn_conv3x3_tc_workspace_bytes(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    const size_t planes = align_up((size_t)N * H * W * kC * 2 * sizeof(__half), 256);
    return 2 * planes + align_up((size_t)kStages * 2 * kPlaneElems * sizeof(__half), 256) + 4096;
}

int ic_nn_conv3x3_tc(const float* d_x, const float* d_w, int N, int H, int W, int data_grad, float* d_y, void* d_workspace,
                     size_t workspace_bytes, void* stream) {
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
    float* partial = ar.get<float>(kMaxBlocks);
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "ic_nn_conv3x3_tc: workspace too small");
    {
        ProfScope ps(IC_PROF_ELEMENTWISE, s, 2);
        maxabs_partial_kernel<<<kMaxBlocks, 256, 0, s>>>(d_w, kW, partial);
        IC_CHECK_LAUNCH();
        pack3x3_kernel<<<cdiv(kStages * kPlaneElems, 256), 256, 0, s>>>(d_w, data_grad, partial, kMaxBlocks, wp, scale, shift);
        IC_CHECK_LAUNCH();
    }
    rc = tc::launch_split_from_nhwc(d_x, N, H, W, kC, 0, bi, 1, s);
    if (rc != IC_OK) return rc;
    tc::ConvTcArgs a;
    memset(&a, 0, sizeof(a));
    a.in = bi;
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
    rc = tc::launch_conv_tc(a, s);
    if (rc != IC_OK) return rc;
    return tc::launch_merge_to_nhwc(bo, N, H, W, kC, d_y, 1, s);
}

}  // extern "C"
