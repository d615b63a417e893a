// Forward pieces of the training loss (code/train.py:303-336 get_loss, :352-431 Distortions):
// HBM-bound reductions with deterministic two-stage sums (double accumulation).
#include "common.cuh"

namespace ic {
namespace {

constexpr int RT = 256;

// block partial sums of bc and bc * heatmap
__global__ void __launch_bounds__(RT) masked_sums_kernel(const float* __restrict__ bc, const float* __restrict__ hm, int64_t n,
                                                         double* __restrict__ partial) {
    double a = 0.0, b = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * RT + threadIdx.x; i < n; i += (int64_t)gridDim.x * RT) {
        float v = bc[i];
        a += v;
        if (hm) b += __fmul_rn(v, hm[i]);
    }
    __shared__ double s0[RT], s1[RT];
    s0[threadIdx.x] = a;
    s1[threadIdx.x] = b;
    __syncthreads();
    for (int s = RT / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            s0[threadIdx.x] += s0[threadIdx.x + s];
            s1[threadIdx.x] += s1[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = s0[0];
        partial[2 * blockIdx.x + 1] = s1[0];
    }
}

__global__ void finish_sums_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int i = 0; i < nblocks; ++i) {
            a += partial[2 * i];
            b += partial[2 * i + 1];
        }
        out[0] = a;
        out[1] = b;
    }
}

// Distortions.get_mse_per_img (code/train.py:400-418): optional tf.cast(int32) (truncation), squared error,
// to_float, mean over (C,H,W).  One block per image.
__global__ void __launch_bounds__(RT) mse_per_image_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t per_img,
                                                           int cast_to_int, float* __restrict__ out) {
    const float* px = x + (int64_t)blockIdx.x * per_img;
    const float* py = y + (int64_t)blockIdx.x * per_img;
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < per_img; i += RT) {
        float d;
        if (cast_to_int) d = (float)((int)py[i] - (int)px[i]);
        else d = __fsub_rn(py[i], px[i]);
        a += (double)__fmul_rn(d, d);
    }
    __shared__ double s0[RT];
    s0[threadIdx.x] = a;
    __syncthreads();
    for (int s = RT / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) s0[threadIdx.x] += s0[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = (float)(s0[0] / (double)per_img);
}

// d/dx_out of the distortion the reference minimises when distortion_to_minimize is 'mse' or 'psnr' (code/train.py:381-397,
// float32 branch of get_mse_per_img):  mse: mean_n mse_n -> 2 (x_out - x) / (N CHW);
// psnr: K_psnr - mean_n 10 log10(255^2 / mse_n) -> (10 / (ln 10 N mse_n)) 2 (x_out - x) / CHW.   mse: per-image MSE (device).
__global__ void dist_bwd_kernel(const float* __restrict__ x, const float* __restrict__ x_out, const float* __restrict__ mse, int N,
                                int64_t per_img, int psnr, float* __restrict__ d) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * per_img) return;
    const int n = (int)(i / per_img);
    const double base = 2.0 / ((double)N * (double)per_img);
    const double c = psnr ? base * 10.0 / (2.302585092994046 * (double)mse[n]) : base;
    d[i] = (float)(c * ((double)x_out[i] - (double)x[i]));
}

}  // namespace
}  // namespace ic

using namespace ic;

extern "C" {

size_t ic_loss_workspace_bytes(void) { return 2 * sizeof(double) * 1024 + 256; }

int ic_masked_sums_fwd(const float* d_bc, const float* d_heatmap, int64_t n, double* d_out, void* d_workspace,
                       size_t workspace_bytes, void* stream) {
    IC_REQUIRE(d_bc && d_out && d_workspace && n > 0, IC_ERR_INVALID, "ic_masked_sums_fwd: bad argument");
    IC_REQUIRE(workspace_bytes >= ic_loss_workspace_bytes(), IC_ERR_WORKSPACE, "ic_masked_sums_fwd: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    int nb = cdiv(n, RT * 8);
    if (nb > 1024) nb = 1024;
    ProfScope ps(IC_PROF_ELEMENTWISE, s, 2);
    masked_sums_kernel<<<nb, RT, 0, s>>>(d_bc, d_heatmap, n, (double*)d_workspace);
    IC_CHECK_LAUNCH();
    finish_sums_kernel<<<1, 32, 0, s>>>((const double*)d_workspace, nb, d_out);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int ic_mse_per_image_fwd(const float* d_x, const float* d_x_out, int N, int64_t per_image, int cast_to_int, float* d_out,
                         void* stream) {
    IC_REQUIRE(d_x && d_x_out && d_out && N > 0 && per_image > 0, IC_ERR_INVALID, "ic_mse_per_image_fwd: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    mse_per_image_kernel<<<N, RT, 0, s>>>(d_x, d_x_out, per_image, cast_to_int, d_out);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

/* gradient of the 'mse' / 'psnr' distortion loss w.r.t. x_out (see dist_bwd_kernel); d_mse: ic_mse_per_image_fwd(..., 0, ...) */
int ic_nn_distortion_bwd(const float* d_x, const float* d_x_out, const float* d_mse, int N, int64_t per_image, int psnr,
                         float* d_dx_out, void* stream) {
    IC_REQUIRE(d_x && d_x_out && d_mse && d_dx_out && N > 0 && per_image > 0, IC_ERR_INVALID, "ic_nn_distortion_bwd: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    dist_bwd_kernel<<<cdiv((int64_t)N * per_image, 256), 256, 0, s>>>(d_x, d_x_out, d_mse, N, per_image, psnr, d_dx_out);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

}  // extern "C"
