// Multi-scale SSIM, both variants of the reference:
//   * ic_msssim_tf_fwd : code/ms_ssim.py:16-186 (float32, one scalar per batch, the training loss)
//   * ic_msssim_np_fwd : code/ms_ssim_np.py:51-200 (float64 on uint8, one value per image, val.py:93)
// Per level one fused kernel: separable Gaussian blur (VALID) of x, y, x*x, y*y, x*y
// in shared memory -> ssim / cs maps -> block partial sums (double); a 2x2 box
// downsample kernel feeds the next level.  All reductions are deterministic
// (fixed-order partial sums, no atomics).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace ic {

size_t msssim_workspace_bytes(int N, int H, int W, int is_double);

namespace {

constexpr int TW = 16, TH = 16, MAXK = 11;

template <typename T>
struct LevelParams {
    int H, W;          // plane size at this level
    int K;             // taps
    int p1, p2;        // REFLECT padding before/after on both axes (tf variant), 0 for np
    int Ho, Wo;        // VALID output size
    T taps[MAXK];
    T c1, c2;
};

__device__ __forceinline__ int reflect_idx(int u, int n) {   // tf.pad REFLECT (no edge repeat)
    if (u < 0) u = -u;
    if (u >= n) u = 2 * (n - 1) - u;
    return u;
}

template <typename T, typename TIn>
__global__ void __launch_bounds__(TW* TH) ssim_level_kernel(const TIn* __restrict__ a, const TIn* __restrict__ b,
                                                            LevelParams<T> p, double* __restrict__ partial) {
    __shared__ T sx[TH + MAXK - 1][TW + MAXK - 1];
    __shared__ T sy[TH + MAXK - 1][TW + MAXK - 1];
    __shared__ T hq[5][TH + MAXK - 1][TW];
    __shared__ double red[2][TW * TH / 32];

    const int plane = blockIdx.z;
    const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
    const int tid = threadIdx.y * TW + threadIdx.x;
    const TIn* pa = a + (int64_t)plane * p.H * p.W;
    const TIn* pb = b + (int64_t)plane * p.H * p.W;
    const int rows = TH + p.K - 1, cols = TW + p.K - 1;
    for (int i = tid; i < rows * cols; i += TW * TH) {
        int r = i / cols, c = i - r * cols;
        int uy = oy0 + r, ux = ox0 + c;             // padded coordinates
        T vx = 0, vy = 0;
        if (uy < p.H + p.p1 + p.p2 && ux < p.W + p.p1 + p.p2) {
            int yy = reflect_idx(uy - p.p1, p.H), xx = reflect_idx(ux - p.p1, p.W);
            vx = (T)pa[(int64_t)yy * p.W + xx];
            vy = (T)pb[(int64_t)yy * p.W + xx];
        }
        sx[r][c] = vx;
        sy[r][c] = vy;
    }
    __syncthreads();
    for (int i = tid; i < rows * TW; i += TW * TH) {
        int r = i / TW, c = i - r * TW;
        T h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0;
        for (int t = 0; t < p.K; ++t) {
            T x = sx[r][c + t], y = sy[r][c + t], k = p.taps[t];
            T xx = x * x, yy = y * y, xy = x * y;   // img1*img1 etc. are materialised in the reference
            h0 += x * k;
            h1 += y * k;
            h2 += xx * k;
            h3 += yy * k;
            h4 += xy * k;
        }
        hq[0][r][c] = h0;
        hq[1][r][c] = h1;
        hq[2][r][c] = h2;
        hq[3][r][c] = h3;
        hq[4][r][c] = h4;
    }
    __syncthreads();
    double s_ssim = 0.0, s_cs = 0.0;
    const int oy = oy0 + threadIdx.y, ox = ox0 + threadIdx.x;
    if (oy < p.Ho && ox < p.Wo) {
        T mu1 = 0, mu2 = 0, s11 = 0, s22 = 0, s12 = 0;
        for (int t = 0; t < p.K; ++t) {
            T k = p.taps[t];
            mu1 += hq[0][threadIdx.y + t][threadIdx.x] * k;
            mu2 += hq[1][threadIdx.y + t][threadIdx.x] * k;
            s11 += hq[2][threadIdx.y + t][threadIdx.x] * k;
            s22 += hq[3][threadIdx.y + t][threadIdx.x] * k;
            s12 += hq[4][threadIdx.y + t][threadIdx.x] * k;
        }
        T mu11 = mu1 * mu1, mu22 = mu2 * mu2, mu12 = mu1 * mu2;
        s11 -= mu11;
        s22 -= mu22;
        s12 -= mu12;
        T v1 = (T)2.0 * s12 + p.c2;
        T v2 = s11 + s22 + p.c2;
        T ssim = (((T)2.0 * mu12 + p.c1) * v1) / ((mu11 + mu22 + p.c1) * v2);
        T cs = v1 / v2;
        s_ssim = (double)ssim;
        s_cs = (double)cs;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s_ssim += __shfl_down_sync(0xffffffffu, s_ssim, o);
        s_cs += __shfl_down_sync(0xffffffffu, s_cs, o);
    }
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = s_ssim;
        red[1][tid >> 5] = s_cs;
    }
    __syncthreads();
    if (tid == 0) {
        double a0 = 0, a1 = 0;
        for (int i = 0; i < TW * TH / 32; ++i) {
            a0 += red[0][i];
            a1 += red[1][i];
        }
        int64_t blk = ((int64_t)plane * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        partial[2 * blk] = a0;
        partial[2 * blk + 1] = a1;
    }
}

// Row-streaming form of ssim_level_kernel for the common case (K taps, no REFLECT padding): a block owns SW output
// columns (one per thread) and marches down `rows_per_block` output rows.  Every input row is staged once in shared
// memory (coalesced loads, prefetched one row ahead), its horizontal blur of (x, y, xx, yy, xy) goes into a K-deep
// register ring, and the vertical blur reads the ring: per output pixel 2 K shared-memory reads and 10 K FMAs instead of
// the 16x16 tile's 2.6x halo re-reads and shared-memory round trip of the half-blurred maps.  Same per-pixel operation
// order as ssim_level_kernel (taps ascending, horizontal pass first), so the per-pixel values are identical.
constexpr int SW = 128;        // output columns (= threads) per block
constexpr int SROWS_MAX = 64;  // output rows per block (the launch halves it until the grid fills the GPU)

template <typename T, typename TIn, int K>
__global__ void __launch_bounds__(SW) ssim_level_stream_kernel(const TIn* __restrict__ a, const TIn* __restrict__ b,
                                                               LevelParams<T> p, int srows, double* __restrict__ partial) {
    __shared__ T sx[2][SW + K - 1];
    __shared__ T sy[2][SW + K - 1];
    __shared__ double red[2][SW / 32];
    const int plane = blockIdx.z, tid = threadIdx.x;
    const int ox0 = blockIdx.x * SW, oy0 = blockIdx.y * srows;
    const int nout = min(srows, p.Ho - oy0), nin = nout + K - 1;       // input rows oy0 .. oy0 + nin - 1 (< H)
    const TIn* pa = a + ((int64_t)plane * p.H + oy0) * p.W;
    const TIn* pb = b + ((int64_t)plane * p.H + oy0) * p.W;
    const int c0 = ox0 + tid, c1 = ox0 + SW + tid;
    const bool v0 = c0 < p.W, v1 = tid < K - 1 && c1 < p.W;
    const bool out_ok = c0 < p.Wo;
    T taps[K];
#pragma unroll
    for (int t = 0; t < K; ++t) taps[t] = p.taps[t];
    T ring[K][5];
    T nx0 = 0, ny0 = 0, nx1 = 0, ny1 = 0;
    if (v0) {
        nx0 = (T)pa[c0];
        ny0 = (T)pb[c0];
    }
    if (v1) {
        nx1 = (T)pa[c1];
        ny1 = (T)pb[c1];
    }
    double s_ssim = 0.0, s_cs = 0.0;
    for (int r0 = 0; r0 < nin; r0 += K) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int r = r0 + j;
            if (r < nin) {                      // block-uniform
                const int buf = r & 1;          // two row buffers: one barrier per row
                sx[buf][tid] = nx0;
                sy[buf][tid] = ny0;
                if (tid < K - 1) {
                    sx[buf][SW + tid] = nx1;
                    sy[buf][SW + tid] = ny1;
                }
                if (r + 1 < nin) {              // next row's loads fly during this row's arithmetic
                    const int64_t o = (int64_t)(r + 1) * p.W;
                    if (v0) {
                        nx0 = (T)pa[o + c0];
                        ny0 = (T)pb[o + c0];
                    }
                    if (v1) {
                        nx1 = (T)pa[o + c1];
                        ny1 = (T)pb[o + c1];
                    }
                }
                __syncthreads();
                T h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0;
#pragma unroll
                for (int t = 0; t < K; ++t) {
                    T x = sx[buf][tid + t], y = sy[buf][tid + t], k = taps[t];
                    T xx = x * x, yy = y * y, xy = x * y;
                    h0 += x * k;
                    h1 += y * k;
                    h2 += xx * k;
                    h3 += yy * k;
                    h4 += xy * k;
                }
                ring[j][0] = h0;
                ring[j][1] = h1;
                ring[j][2] = h2;
                ring[j][3] = h3;
                ring[j][4] = h4;
                if (r >= K - 1) {               // output row r - (K-1): input rows r-K+1 .. r = ring slots (j+1+t) % K
                    T mu1 = 0, mu2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
                    for (int t = 0; t < K; ++t) {
                        const int q = (j + 1 + t) % K;
                        T k = taps[t];
                        mu1 += ring[q][0] * k;
                        mu2 += ring[q][1] * k;
                        s11 += ring[q][2] * k;
                        s22 += ring[q][3] * k;
                        s12 += ring[q][4] * k;
                    }
                    T mu11 = mu1 * mu1, mu22 = mu2 * mu2, mu12 = mu1 * mu2;
                    s11 -= mu11;
                    s22 -= mu22;
                    s12 -= mu12;
                    T v1n = (T)2.0 * s12 + p.c2;
                    T v2n = s11 + s22 + p.c2;
                    T ssim = (((T)2.0 * mu12 + p.c1) * v1n) / ((mu11 + mu22 + p.c1) * v2n);
                    T cs = v1n / v2n;
                    if (out_ok) {
                        s_ssim += (double)ssim;
                        s_cs += (double)cs;
                    }
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s_ssim += __shfl_down_sync(0xffffffffu, s_ssim, o);
        s_cs += __shfl_down_sync(0xffffffffu, s_cs, o);
    }
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = s_ssim;
        red[1][tid >> 5] = s_cs;
    }
    __syncthreads();
    if (tid == 0) {
        double a0 = 0, a1 = 0;
        for (int i = 0; i < SW / 32; ++i) {
            a0 += red[0][i];
            a1 += red[1][i];
        }
        int64_t blk = ((int64_t)plane * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        partial[2 * blk] = a0;
        partial[2 * blk + 1] = a1;
    }
}

// sums the block partials of `planes_per_group` consecutive planes in fixed order
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int blocks_per_plane, int planes_per_group,
                                       double inv_count, double* __restrict__ out /* [groups][2] */) {
    __shared__ double r0[256], r1[256];
    const int g = blockIdx.x;
    const int64_t n = (int64_t)blocks_per_plane * planes_per_group;
    const double* p = partial + 2 * (int64_t)g * n;
    double a0 = 0, a1 = 0;
    for (int64_t i = threadIdx.x; i < n; i += 256) {
        a0 += p[2 * i];
        a1 += p[2 * i + 1];
    }
    r0[threadIdx.x] = a0;
    r1[threadIdx.x] = a1;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            r0[threadIdx.x] += r0[threadIdx.x + s];
            r1[threadIdx.x] += r1[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[2 * g] = r0[0] * inv_count;
        out[2 * g + 1] = r1[0] * inv_count;
    }
}

// 2x2 box downsample.
//   tf: kernel_blur(pad=True) REFLECT pad (0,1) + [.5,.5] separable + [::2, ::2]  (code/ms_ssim.py:46-64,179-181)
//   np: ndimage.convolve(ones(2,2)/4, mode='reflect')[::2, ::2]                    (code/ms_ssim_np.py:96,106-108)
template <typename T, typename TIn, bool TF>
__global__ void downsample_kernel(const TIn* __restrict__ in, int H, int W, int Hd, int Wd, int64_t total,
                                  T* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int x = (int)(i % Wd);
    int64_t r = i / Wd;
    int y = (int)(r % Hd);
    int64_t plane = r / Hd;
    const TIn* p = in + plane * H * W;
    int y0 = 2 * y, x0 = 2 * x;
    int y1 = y0 + 1, x1 = x0 + 1;
    if (TF) {
        if (y1 >= H) y1 = H - 2;    // REFLECT
        if (x1 >= W) x1 = W - 2;
    } else {
        if (y1 >= H) y1 = H - 1;    // scipy 'reflect' duplicates the edge sample
        if (x1 >= W) x1 = W - 1;
    }
    T a = (T)p[(int64_t)y0 * W + x0], b = (T)p[(int64_t)y0 * W + x1];
    T c = (T)p[(int64_t)y1 * W + x0], d = (T)p[(int64_t)y1 * W + x1];
    if (TF) {
        T h0 = (T)0.5 * a + (T)0.5 * b, h1 = (T)0.5 * c + (T)0.5 * d;
        out[i] = (T)0.5 * h0 + (T)0.5 * h1;
    } else {
        out[i] = (a + c + b + d) / (T)4.0;
    }
}

__global__ void combine_tf_kernel(const double* __restrict__ lv /* [5][2] */, float* out, float* levels) {
    if (threadIdx.x != 0) return;
    const float w[5] = {0.0448f, 0.2856f, 0.3001f, 0.2363f, 0.1333f};
    float prod = 1.f;
    for (int l = 0; l < 4; ++l) prod *= powf((float)lv[2 * l + 1], w[l]);
    out[0] = prod * powf((float)lv[2 * 4], w[4]);
    if (levels)
        for (int l = 0; l < 5; ++l) {
            levels[l] = (float)lv[2 * l];
            levels[5 + l] = (float)lv[2 * l + 1];
        }
}

__global__ void combine_np_kernel(const double* __restrict__ lv /* [5][N][2] */, int N, double* out) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double w[5] = {0.0448, 0.2856, 0.3001, 0.2363, 0.1333};
    double prod = 1.0;
    for (int l = 0; l < 4; ++l) prod *= pow(lv[((int64_t)l * N + n) * 2 + 1], w[l]);
    out[n] = prod * pow(lv[((int64_t)4 * N + n) * 2], w[4]);
}

template <typename T>
int fill_level(LevelParams<T>& p, int H, int W, bool tf) {
    p.H = H;
    p.W = W;
    const int size = std::min(11, std::min(H, W));
    const double sigma = size * 1.5 / 11;
    double g[MAXK];
    double sum = 0;
    if (tf) {   // ms_ssim.gauss_kernel: 2*(size//2)+1 taps (code/ms_ssim.py:5-13)
        const int n = size / 2;
        p.K = 2 * n + 1;
        for (int i = 0; i < p.K; ++i) {
            double x = i - n;
            g[i] = exp(-x * x / (2 * sigma * sigma));
            sum += fabs(g[i]);
        }
        // gaussian_blur pads by the W-derived amount on BOTH axes; "total_pad + 1 // 2" (code/ms_ssim.py:24-29)
        int total_pad = std::max(p.K - W, 0);
        p.p1 = total_pad + 1 / 2;
        p.p2 = total_pad / 2;
    } else {    // _FSpecialGauss, `size` taps, half-sample offset when even (code/ms_ssim_np.py:113-124)
        p.K = size;
        const double off = (size % 2 == 0) ? 0.5 : 0.0;
        for (int i = 0; i < p.K; ++i) {
            double x = off - size / 2 + i;
            g[i] = exp(-(x * x) / (2.0 * sigma * sigma));
            sum += g[i];
        }
        p.p1 = p.p2 = 0;
    }
    for (int i = 0; i < p.K; ++i) p.taps[i] = (T)(g[i] / sum);
    p.Ho = H + p.p1 + p.p2 - p.K + 1;
    p.Wo = W + p.p1 + p.p2 - p.K + 1;
    p.c1 = (T)((0.01 * 255) * (0.01 * 255));
    p.c2 = (T)((0.03 * 255) * (0.03 * 255));
    if (p.Ho <= 0 || p.Wo <= 0 || p.p1 >= H || p.p1 >= W) return IC_ERR_INVALID;
    return IC_OK;
}

template <typename T>
struct LevelBufs {
    T* a[5];
    T* b[5];
    int h[5], w[5];
};

template <typename T, typename TIn, bool TF>
int run_msssim(const TIn* img1, const TIn* img2, int N, int H, int W, void* ws, size_t ws_bytes, double* lv_out,
               cudaStream_t s, LevelBufs<T>* bufs_out = nullptr) {
    const int P = N * 3;
    Arena ar(ws, ws_bytes);
    // level buffers (levels 1..4), both images
    T* bufA[5] = {nullptr};
    T* bufB[5] = {nullptr};
    int hs[5], wsz[5];
    hs[0] = H;
    wsz[0] = W;
    for (int l = 1; l < 5; ++l) {
        hs[l] = (hs[l - 1] + 1) / 2;
        wsz[l] = (wsz[l - 1] + 1) / 2;
        bufA[l] = ar.get<T>((size_t)P * hs[l] * wsz[l]);
        bufB[l] = ar.get<T>((size_t)P * hs[l] * wsz[l]);
    }
    size_t max_blocks = (size_t)P * cdiv(H, TH) * cdiv(W, TW);
    double* partial = ar.get<double>(2 * max_blocks);
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "ms-ssim workspace too small: need %zu, have %zu", ar.off, ws_bytes);
    const int groups = TF ? 1 : N;
    ProfScope ps(IC_PROF_MSSSIM, s, 5 * 2 + 4 * 2 + 1);
    for (int l = 0; l < 5; ++l) {
        LevelParams<T> p;
        int rc = fill_level<T>(p, hs[l], wsz[l], TF);
        IC_REQUIRE(rc == IC_OK, IC_ERR_INVALID,
                   "ms-ssim: level %d is %dx%d, too small for the %d-tap blur (the reference raises here)", l, hs[l],
                   wsz[l], p.K);
        // large full 11-tap levels without REFLECT padding: row-streaming kernel;
        // IC_MSSSIM_TILED=1 forces the 16x16-tile kernel everywhere (A/B in tests)
        const bool force_tiled = getenv("IC_MSSSIM_TILED") && atoi(getenv("IC_MSSSIM_TILED"));
        // (below ~0.4 M output pixels per level the 16x16 tiles win: the row march is a serial chain of Ho barriers.
        //  B200, 72 planes, float: 192x128 59 -> 35 us, 96x64 19 -> 19 us, 48x32 8 -> 18 us; profiles/r2_summary.md)
        const bool stream = p.K == MAXK && p.p1 == 0 && p.p2 == 0 && !force_tiled && (int64_t)P * p.Ho * p.Wo >= 400000;
        dim3 grid(cdiv(p.Wo, TW), cdiv(p.Ho, TH), P), block(TW, TH);
        if (stream) {
            // rows per block: as many as keep >= 8 blocks per SM in flight (each extra block re-reads K-1 halo rows)
            int srows = SROWS_MAX;
            while (srows > 16 && (int64_t)cdiv(p.Wo, SW) * cdiv(p.Ho, srows) * P < 8 * 148) srows >>= 1;
            grid = dim3(cdiv(p.Wo, SW), cdiv(p.Ho, srows), P);
            if (l == 0)
                ssim_level_stream_kernel<T, TIn, MAXK><<<grid, SW, 0, s>>>(img1, img2, p, srows, partial);
            else
                ssim_level_stream_kernel<T, T, MAXK><<<grid, SW, 0, s>>>(bufA[l], bufB[l], p, srows, partial);
        } else if (l == 0) {
            ssim_level_kernel<T, TIn><<<grid, block, 0, s>>>(img1, img2, p, partial);
        } else {
            ssim_level_kernel<T, T><<<grid, block, 0, s>>>(bufA[l], bufB[l], p, partial);
        }
        IC_CHECK_LAUNCH();
        int bpp = grid.x * grid.y;
        double inv = 1.0 / ((double)p.Ho * p.Wo * (TF ? P : 3));
        reduce_partials_kernel<<<groups, 256, 0, s>>>(partial, bpp, TF ? P : 3, inv, lv_out + (size_t)l * groups * 2);
        IC_CHECK_LAUNCH();
        if (l < 4) {
            int64_t total = (int64_t)P * hs[l + 1] * wsz[l + 1];
            if (TF) IC_REQUIRE(hs[l] >= 2 && wsz[l] >= 2, IC_ERR_INVALID, "ms-ssim: cannot REFLECT-pad a %dx%d level", hs[l], wsz[l]);
            if (l == 0) {
                downsample_kernel<T, TIn, TF><<<cdiv(total, 256), 256, 0, s>>>(img1, hs[l], wsz[l], hs[l + 1], wsz[l + 1], total, bufA[l + 1]);
                downsample_kernel<T, TIn, TF><<<cdiv(total, 256), 256, 0, s>>>(img2, hs[l], wsz[l], hs[l + 1], wsz[l + 1], total, bufB[l + 1]);
            } else {
                downsample_kernel<T, T, TF><<<cdiv(total, 256), 256, 0, s>>>(bufA[l], hs[l], wsz[l], hs[l + 1], wsz[l + 1], total, bufA[l + 1]);
                downsample_kernel<T, T, TF><<<cdiv(total, 256), 256, 0, s>>>(bufB[l], hs[l], wsz[l], hs[l + 1], wsz[l + 1], total, bufB[l + 1]);
            }
            IC_CHECK_LAUNCH();
        }
    }
    if (bufs_out)
        for (int l = 0; l < 5; ++l) {
            bufs_out->a[l] = bufA[l];
            bufs_out->b[l] = bufB[l];
            bufs_out->h[l] = hs[l];
            bufs_out->w[l] = wsz[l];
        }
    return IC_OK;
}

// ------------------------------------------------------------------ backward of the tf variant (w.r.t. img2)
// value = ssim_4^w4 * prod_{l<4} cs_l^wl with ssim_l / cs_l the batch means of the per-pixel maps
// (code/ms_ssim.py:110-111,183-186): d value / d cs_l = w_l value / cs_l, spread evenly over the level's
// count_l map entries.  coef[2l] multiplies the per-pixel ssim map of level l, coef[2l+1] its cs map.
struct BwdCounts {
    double count[5];
};

__global__ void bwd_coef_kernel(const double* __restrict__ lv /* [5][2] */, BwdCounts cnt, float grad_out, float* __restrict__ coef,
                                float* __restrict__ value_out) {
    if (threadIdx.x != 0) return;
    const float w[5] = {0.0448f, 0.2856f, 0.3001f, 0.2363f, 0.1333f};
    double prod = 1.0;
    for (int l = 0; l < 4; ++l) prod *= pow(lv[2 * l + 1], (double)w[l]);
    const double val = prod * pow(lv[2 * 4], (double)w[4]);
    for (int l = 0; l < 5; ++l) {
        coef[2 * l] = (l == 4) ? (float)((double)grad_out * w[4] * val / lv[2 * 4] / cnt.count[4]) : 0.f;
        coef[2 * l + 1] = (l < 4) ? (float)((double)grad_out * w[l] * val / lv[2 * l + 1] / cnt.count[l]) : 0.f;
    }
    if (value_out) value_out[0] = (float)val;
}

// Recomputes the blurred statistics of one level exactly like ssim_level_kernel and writes, per map entry, the
// gradient w.r.t. mu2 = G*b, e22 = G*(b*b), e12 = G*(a*b)   (code/ms_ssim.py:86-109).
__global__ void __launch_bounds__(TW* TH) ssim_level_bwd_maps_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                     LevelParams<float> p, const float* __restrict__ coef, int level,
                                                                     float* __restrict__ g_mu, float* __restrict__ g_22,
                                                                     float* __restrict__ g_12) {
    __shared__ float sx[TH + MAXK - 1][TW + MAXK - 1];
    __shared__ float sy[TH + MAXK - 1][TW + MAXK - 1];
    __shared__ float hq[5][TH + MAXK - 1][TW];
    const int plane = blockIdx.z;
    const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
    const int tid = threadIdx.y * TW + threadIdx.x;
    const float* pa = a + (int64_t)plane * p.H * p.W;
    const float* pb = b + (int64_t)plane * p.H * p.W;
    const int rows = TH + p.K - 1, cols = TW + p.K - 1;
    for (int i = tid; i < rows * cols; i += TW * TH) {
        int r = i / cols, c = i - r * cols;
        int uy = oy0 + r, ux = ox0 + c;
        float vx = 0, vy = 0;
        if (uy < p.H + p.p1 + p.p2 && ux < p.W + p.p1 + p.p2) {
            int yy = reflect_idx(uy - p.p1, p.H), xx = reflect_idx(ux - p.p1, p.W);
            vx = pa[(int64_t)yy * p.W + xx];
            vy = pb[(int64_t)yy * p.W + xx];
        }
        sx[r][c] = vx;
        sy[r][c] = vy;
    }
    __syncthreads();
    for (int i = tid; i < rows * TW; i += TW * TH) {
        int r = i / TW, c = i - r * TW;
        float h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0;
        for (int t = 0; t < p.K; ++t) {
            float x = sx[r][c + t], y = sy[r][c + t], k = p.taps[t];
            h0 += x * k;
            h1 += y * k;
            h2 += (x * x) * k;
            h3 += (y * y) * k;
            h4 += (x * y) * k;
        }
        hq[0][r][c] = h0;
        hq[1][r][c] = h1;
        hq[2][r][c] = h2;
        hq[3][r][c] = h3;
        hq[4][r][c] = h4;
    }
    __syncthreads();
    const int oy = oy0 + threadIdx.y, ox = ox0 + threadIdx.x;
    if (oy >= p.Ho || ox >= p.Wo) return;
    float mu1 = 0, mu2 = 0, e11 = 0, e22 = 0, e12 = 0;
    for (int t = 0; t < p.K; ++t) {
        float k = p.taps[t];
        mu1 += hq[0][threadIdx.y + t][threadIdx.x] * k;
        mu2 += hq[1][threadIdx.y + t][threadIdx.x] * k;
        e11 += hq[2][threadIdx.y + t][threadIdx.x] * k;
        e22 += hq[3][threadIdx.y + t][threadIdx.x] * k;
        e12 += hq[4][threadIdx.y + t][threadIdx.x] * k;
    }
    const float gs = coef[2 * level], gcs = coef[2 * level + 1];
    const float v1 = 2.f * (e12 - mu1 * mu2) + p.c2;
    const float v2 = (e11 - mu1 * mu1) + (e22 - mu2 * mu2) + p.c2;
    const float An = 2.f * mu1 * mu2 + p.c1, Bn = mu1 * mu1 + mu2 * mu2 + p.c1;
    const float lum = An / Bn, cs = v1 / v2;
    const float wcs = gcs + gs * lum;      // weight on the cs factor
    const float wl = gs * cs;              // weight on the luminance factor
    const float gv1 = wcs / v2, gv2 = -wcs * cs / v2;
    const float dlum = (2.f * mu1 * Bn - An * 2.f * mu2) / (Bn * Bn);
    const int64_t o = ((int64_t)plane * p.Ho + oy) * p.Wo + ox;
    g_mu[o] = wl * dlum - 2.f * mu1 * gv1 - 2.f * mu2 * gv2;
    g_22[o] = gv2;
    g_12[o] = 2.f * gv1;
}

// rows of the REFLECT-padded plane that read source row y: y + p1, and the mirrored copies
__device__ __forceinline__ int pad_preimages(int y, int n, int p1, int p2, int* u) {
    int c = 0;
    u[c++] = y + p1;
    if (y >= 1 && y <= p1) u[c++] = p1 - y;
    if (y <= n - 2 && y >= n - 1 - p2) u[c++] = p1 + 2 * (n - 1) - y;
    return c;
}

// d loss / d b of one level: adjoint of the VALID separable blur of (b, b*b, a*b) (full correlation with the
// same taps), folded through the REFLECT padding, plus the adjoint of the 2x2 box downsample that produced the
// next level (code/ms_ssim.py:46-64,179-181).  One thread per source pixel, gather form: no atomics.
__global__ void __launch_bounds__(256) ssim_level_bwd_input_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                    LevelParams<float> p, const float* __restrict__ g_mu,
                                                                    const float* __restrict__ g_22, const float* __restrict__ g_12,
                                                                    const float* __restrict__ d_next, int Hn, int Wn, int64_t total,
                                                                    float* __restrict__ d_b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % p.W);
    const int64_t r = i / p.W;
    const int y = (int)(r % p.H);
    const int64_t plane = r / p.H;
    const float av = a[i], bv = b[i];
    const float* pm = g_mu + plane * p.Ho * p.Wo;
    const float* p22 = g_22 + plane * p.Ho * p.Wo;
    const float* p12 = g_12 + plane * p.Ho * p.Wo;
    int uys[3], uxs[3];
    const int ny = pad_preimages(y, p.H, p.p1, p.p2, uys), nx = pad_preimages(x, p.W, p.p1, p.p2, uxs);
    float s_mu = 0.f, s_22 = 0.f, s_12 = 0.f;
    for (int iy = 0; iy < ny; ++iy)
        for (int ix = 0; ix < nx; ++ix) {
            const int uy = uys[iy], ux = uxs[ix];
            const int t0 = max(0, uy - (p.Ho - 1)), t1 = min(p.K - 1, uy);
            const int q0 = max(0, ux - (p.Wo - 1)), q1 = min(p.K - 1, ux);
            for (int t = t0; t <= t1; ++t) {
                const int64_t row = (int64_t)(uy - t) * p.Wo;
                float r_mu = 0.f, r_22 = 0.f, r_12 = 0.f;
                for (int q = q0; q <= q1; ++q) {
                    const float k = p.taps[q];
                    const int64_t o = row + (ux - q);
                    r_mu += k * pm[o];
                    r_22 += k * p22[o];
                    r_12 += k * p12[o];
                }
                s_mu += p.taps[t] * r_mu;
                s_22 += p.taps[t] * r_22;
                s_12 += p.taps[t] * r_12;
            }
        }
    float acc = s_mu + 2.f * bv * s_22 + av * s_12;
    if (d_next) {
        // output (i, j) of the downsample read rows 2i and 2i+1 (REFLECT: row H -> H-2)
        const float* dn = d_next + plane * Hn * Wn;
        int is[2], js[2], ni = 0, nj = 0;
        is[ni++] = y >> 1;
        if ((p.H & 1) && y == p.H - 2) is[ni++] = (p.H - 1) >> 1;
        js[nj++] = x >> 1;
        if ((p.W & 1) && x == p.W - 2) js[nj++] = (p.W - 1) >> 1;
        float s = 0.f;
        for (int ii = 0; ii < ni; ++ii)
            for (int jj = 0; jj < nj; ++jj) s += dn[(int64_t)is[ii] * Wn + js[jj]];
        acc += 0.25f * s;
    }
    d_b[i] = acc;
}

}  // namespace

size_t msssim_bwd_workspace_bytes(int N, int H, int W) {
    const size_t P = (size_t)N * 3;
    const size_t h1 = (H + 1) / 2, w1 = (W + 1) / 2;
    size_t b = 256 + 256;                                        // level means, coefficients
    b += 3 * (align_up(P * H * W * 4, 256) + 256);               // g maps of the largest level
    b += 2 * (align_up(P * h1 * w1 * 4, 256) + 256);             // gradient ping-pong of levels >= 1
    return b + msssim_workspace_bytes(N, H, W, 0) + 1024;
}

int msssim_tf_bwd(const float* a, const float* b, int N, int H, int W, float grad_out, float* d_b, float* value_out, void* ws,
                  size_t ws_bytes, cudaStream_t s) {
    IC_REQUIRE(ws_bytes >= msssim_bwd_workspace_bytes(N, H, W), IC_ERR_WORKSPACE, "ms-ssim backward workspace too small: need %zu, have %zu",
               msssim_bwd_workspace_bytes(N, H, W), ws_bytes);
    const int P = N * 3;
    Arena ar(ws, ws_bytes);
    double* lv = ar.get<double>(32);
    float* coef = ar.get<float>(64);
    float* g[3];
    for (int i = 0; i < 3; ++i) g[i] = ar.get<float>((size_t)P * H * W);
    const size_t n1 = (size_t)P * ((H + 1) / 2) * ((W + 1) / 2);
    float* dbuf[2] = {ar.get<float>(n1), ar.get<float>(n1)};
    char* fws = (char*)ar.get<char>(0);
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "ms-ssim backward workspace too small");
    LevelBufs<float> lb;
    int rc = run_msssim<float, float, true>(a, b, N, H, W, fws, ws_bytes - ar.off, lv, s, &lb);
    if (rc != IC_OK) return rc;
    lb.a[0] = const_cast<float*>(a);
    lb.b[0] = const_cast<float*>(b);
    LevelParams<float> lp[5];
    BwdCounts cnt;
    for (int l = 0; l < 5; ++l) {
        rc = fill_level<float>(lp[l], lb.h[l], lb.w[l], true);
        IC_REQUIRE(rc == IC_OK, IC_ERR_INVALID, "ms-ssim backward: level %d too small", l);
        cnt.count[l] = (double)lp[l].Ho * lp[l].Wo * P;
    }
    ProfScope ps(IC_PROF_MSSSIM, s, 1 + 5 * 2);
    bwd_coef_kernel<<<1, 32, 0, s>>>(lv, cnt, grad_out, coef, value_out);
    IC_CHECK_LAUNCH();
    const float* d_next = nullptr;
    for (int l = 4; l >= 0; --l) {
        const LevelParams<float>& p = lp[l];
        dim3 grid(cdiv(p.Wo, TW), cdiv(p.Ho, TH), P), block(TW, TH);
        ssim_level_bwd_maps_kernel<<<grid, block, 0, s>>>(lb.a[l], lb.b[l], p, coef, l, g[0], g[1], g[2]);
        IC_CHECK_LAUNCH();
        float* out = (l == 0) ? d_b : dbuf[l & 1];
        const int64_t total = (int64_t)P * p.H * p.W;
        ssim_level_bwd_input_kernel<<<cdiv(total, 256), 256, 0, s>>>(lb.a[l], lb.b[l], p, g[0], g[1], g[2], d_next,
                                                                     l < 4 ? lb.h[l + 1] : 0, l < 4 ? lb.w[l + 1] : 0, total, out);
        IC_CHECK_LAUNCH();
        d_next = out;
    }
    return IC_OK;
}

size_t msssim_workspace_bytes(int N, int H, int W, int is_double) {
    size_t e = is_double ? 8 : 4;
    size_t b = 0;
    int h = H, w = W;
    for (int l = 1; l < 5; ++l) {
        h = (h + 1) / 2;
        w = (w + 1) / 2;
        b += 2 * (align_up((size_t)N * 3 * h * w * e, 256) + 256);
    }
    b += 2 * sizeof(double) * (size_t)N * 3 * cdiv(H, TH) * cdiv(W, TW) + 256;
    b += sizeof(double) * 2 * 5 * (size_t)std::max(N, 1) + 256;
    return b + 1024;
}

int msssim_tf(const float* a, const float* b, int N, int H, int W, float* out, float* levels, void* ws,
              size_t ws_bytes, cudaStream_t s) {
    IC_REQUIRE(ws_bytes >= 256, IC_ERR_WORKSPACE, "ms-ssim workspace too small");
    double* lv = (double*)ws;    // first 256 B: level means
    int rc = run_msssim<float, float, true>(a, b, N, H, W, (char*)ws + 256, ws_bytes - 256, lv, s);
    if (rc != IC_OK) return rc;
    combine_tf_kernel<<<1, 32, 0, s>>>(lv, out, levels);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int msssim_np(const uint8_t* a, const uint8_t* b, int N, int H, int W, double* out, void* ws, size_t ws_bytes,
              cudaStream_t s) {
    size_t lvb = align_up(sizeof(double) * 10 * (size_t)N, 256);
    IC_REQUIRE(ws_bytes >= lvb, IC_ERR_WORKSPACE, "ms-ssim workspace too small");
    double* lv = (double*)ws;
    int rc = run_msssim<double, uint8_t, false>(a, b, N, H, W, (char*)ws + lvb, ws_bytes - lvb, lv, s);
    if (rc != IC_OK) return rc;
    combine_np_kernel<<<cdiv(N, 128), 128, 0, s>>>(lv, N, out);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

}  // namespace ic
