// Training-step primitives (float32 FFMA): what tf.gradients + the two Adam optimisers of
// code/train.py:339-349 execute for the graph of code/train.py:86-132, restated as explicit
// forward / backward kernels.  Activations are NHWC float32 (channel counts padded to a
// multiple of 4), convolution weights [KH][KW][Cin][Cout] in the orientation of the op.
//
//   convolution      forward: conv_simt.cu (scale 1, shift 0); backward-data: the same kernel in
//                    the other mode over per-tap transposed weights; backward-filter: implicit
//                    GEMM dW = A^T dY with a split reduction over the output pixels
//   batch norm       slim.batch_norm(is_training=True, fused) (code/autoencoder.py:115-125, A.1):
//                    batch mean / biased variance, moving averages with the unbiased variance
//   heatmap+quantize backward of _get_heatmap3D, _mask_with_heatmap, qsoft (autoencoder.py:171-200,
//                    quantizer.py:60-100); qbar = qsoft + stop_gradient(qhard - qsoft)
//   Adam             tf.train.AdamOptimizer._apply_dense
// All reductions are fixed-order (partials + a finalize kernel), no float atomics.
#include <math.h>
#include <string.h>

#include <algorithm>

#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace ic {

namespace {

// ------------------------------------------------------------------ small helpers
float* g_ones = nullptr;
float* g_zeros = nullptr;
constexpr int kConstLen = 1024;

int ensure_consts() {
    if (g_ones) return IC_OK;
    float h[kConstLen];
    for (int i = 0; i < kConstLen; ++i) h[i] = 1.f;
    IC_CHECK_CUDA(cudaMalloc((void**)&g_ones, sizeof(h)));
    IC_CHECK_CUDA(cudaMemcpy(g_ones, h, sizeof(h), cudaMemcpyHostToDevice));
    IC_CHECK_CUDA(cudaMalloc((void**)&g_zeros, sizeof(h)));
    IC_CHECK_CUDA(cudaMemset(g_zeros, 0, sizeof(h)));
    return IC_OK;
}

struct Geo {
    int N, Hi, Wi, Cin, KH, KW, stride, Cout, transposed, valid;
    int Ho, Wo, pad_t, pad_l;
};

// SAME (TF, A.2) or VALID geometry of the op
int make_geo(Geo& g) {
    IC_REQUIRE(g.N > 0 && g.Hi > 0 && g.Wi > 0 && g.Cin > 0 && g.Cout > 0 && g.KH > 0 && g.KW > 0 && g.stride > 0,
               IC_ERR_INVALID, "nn conv: bad geometry");
    IC_REQUIRE(g.Cin % 4 == 0 && g.Cout % 4 == 0, IC_ERR_INVALID, "nn conv: Cin (%d) and Cout (%d) must be multiples of 4", g.Cin,
               g.Cout);
    IC_REQUIRE(g.Cout <= kConstLen && g.Cin <= kConstLen, IC_ERR_UNSUPPORTED, "nn conv: more than %d channels", kConstLen);
    if (g.valid) {
        IC_REQUIRE(!g.transposed && g.Hi >= g.KH && g.Wi >= g.KW, IC_ERR_INVALID, "nn conv: VALID needs input >= kernel");
        g.Ho = (g.Hi - g.KH) / g.stride + 1;
        g.Wo = (g.Wi - g.KW) / g.stride + 1;
        g.pad_t = g.pad_l = 0;
    } else if (!g.transposed) {
        g.Ho = (g.Hi + g.stride - 1) / g.stride;
        g.Wo = (g.Wi + g.stride - 1) / g.stride;
        g.pad_t = same_pad_before(g.Hi, g.KH, g.stride);
        g.pad_l = same_pad_before(g.Wi, g.KW, g.stride);
    } else {
        g.Ho = g.Hi * g.stride;
        g.Wo = g.Wi * g.stride;
        g.pad_t = same_pad_before(g.Ho, g.KH, g.stride);     // of the forward conv this op is the gradient of
        g.pad_l = same_pad_before(g.Wo, g.KW, g.stride);
    }
    return IC_OK;
}

__global__ void transpose_taps_kernel(const float* __restrict__ w, float* __restrict__ wT, int taps, int Cin, int Cout) {
    const int64_t total = (int64_t)taps * Cin * Cout;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int tap = (int)(i / ((int64_t)Cin * Cout));
        const int r = (int)(i - (int64_t)tap * Cin * Cout);
        const int ci = r / Cout, co = r - ci * Cout;
        wT[((int64_t)tap * Cout + co) * Cin + ci] = w[i];
    }
}

// ------------------------------------------------------------------ backward filter
// dW[k][co] = sum_m A[m][k] * dY[m][co]; A = the im2col gather of the forward op (same index
// arithmetic as conv_simt.cu's load_a), m over the N*Ho*Wo output pixels.
struct WgradDesc {
    const float* x;
    const float* dy;
    float* part;           // [splits][K][Cout]
    Geo g;
    int64_t M;
    int64_t m_per_split;
};

constexpr int WM = 16;

// WTN = 64: 64 k x 64 output channels per block; WTN = 32: 128 k x 32 channels (layers with <= 32 output channels)
template <int WTN>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(WgradDesc d) {
    constexpr int WTK = 64 * 64 / WTN;          // k (tap, cin) entries per block
    constexpr int TXN = WTN / 4;                // threads across the channel dimension
    constexpr int KQ = WTK / 4;                 // float4 per A row
    constexpr int AL = WM * KQ / 256;           // A float4 loads per thread (1 or 2)
    __shared__ __align__(16) float As[WM][WTK + 4];
    __shared__ __align__(16) float Bs[WM][WTN + 4];
    const Geo& g = d.g;
    const int t = threadIdx.x;
    const int K = g.KH * g.KW * g.Cin;
    const int k0 = blockIdx.x * WTK, n0 = blockIdx.y * WTN;
    const int64_t m_begin = (int64_t)blockIdx.z * d.m_per_split;
    const int64_t m_end = min(d.M, m_begin + d.m_per_split);
    // A-load assignment: AL x (row lm_a, 4 consecutive k: same tap since Cin % 4 == 0)
    int lm_a[AL], l4_a[AL], ky[AL], kx[AL], ci[AL];
    bool kvalid[AL];
#pragma unroll
    for (int j = 0; j < AL; ++j) {
        const int idx = t + 256 * j;
        lm_a[j] = idx / KQ;
        l4_a[j] = (idx % KQ) * 4;
        const int k = k0 + l4_a[j];
        kvalid[j] = k < K;
        ky[j] = kx[j] = ci[j] = 0;
        if (kvalid[j]) {
            const int tap = k / g.Cin;
            ci[j] = k - tap * g.Cin;
            ky[j] = tap / g.KW;
            kx[j] = tap - ky[j] * g.KW;
        }
    }
    // B-load assignment: the first WM * TXN threads
    const int lm_b = t / TXN, l4_b = (t % TXN) * 4;
    const bool bload = t < WM * TXN;
    const bool nvalid = bload && n0 + l4_b < g.Cout;
    const int tx = t % TXN, ty = t / TXN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int64_t hw = (int64_t)g.Ho * g.Wo;
    // gathers this thread's float4s of the A (im2col) and B (output gradient) tiles of the 16 rows starting at mb
    auto load = [&](int64_t mb, float4 (&a)[AL], float4& b) {
#pragma unroll
        for (int j = 0; j < AL; ++j) {
            a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int64_t m = mb + lm_a[j];
            if (m < m_end && kvalid[j]) {
                const int pn = (int)(m / hw);
                const int r = (int)(m - (int64_t)pn * hw);
                const int oy = r / g.Wo, ox = r - oy * g.Wo;
                int iy, ix;
                bool ok;
                if (!g.transposed) {
                    iy = oy * g.stride - g.pad_t + ky[j];
                    ix = ox * g.stride - g.pad_l + kx[j];
                    ok = iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi;
                } else {
                    const int ny = oy + g.pad_t - ky[j], nx = ox + g.pad_l - kx[j];
                    ok = ny >= 0 && nx >= 0 && (ny % g.stride) == 0 && (nx % g.stride) == 0;
                    iy = ny / g.stride;
                    ix = nx / g.stride;
                    ok = ok && iy < g.Hi && ix < g.Wi;
                }
                if (ok) a[j] = *reinterpret_cast<const float4*>(d.x + (((int64_t)pn * g.Hi + iy) * g.Wi + ix) * g.Cin + ci[j]);
            }
        }
        b = make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t mbr = mb + lm_b;
        if (nvalid && mbr < m_end) b = *reinterpret_cast<const float4*>(d.dy + mbr * g.Cout + n0 + l4_b);
    };
    float4 a[AL], b;
    load(m_begin, a, b);
    for (int64_t mb = m_begin; mb < m_end; mb += WM) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < AL; ++j) *reinterpret_cast<float4*>(&As[lm_a[j]][l4_a[j]]) = a[j];
        if (bload) *reinterpret_cast<float4*>(&Bs[lm_b][l4_b]) = b;
        __syncthreads();
        if (mb + WM < m_end) load(mb + WM, a, b);        // in flight while this tile is consumed
#pragma unroll
        for (int mm = 0; mm < WM; ++mm) {
            const float4 av = *reinterpret_cast<const float4*>(&As[mm][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[mm][tx * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
    float* out = d.part + (int64_t)blockIdx.z * K * g.Cout;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int kk = k0 + ty * 4 + i;
        if (kk >= K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < g.Cout) out[(int64_t)kk * g.Cout + n] = acc[i][j];
        }
    }
}

__global__ void sum_splits_kernel(const float* __restrict__ part, int splits, int64_t count, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += part[(int64_t)k * count + i];
        out[i] = s;
    }
}

int wgrad_splits(const Geo& g) {
    const int64_t M = (int64_t)g.N * g.Ho * g.Wo;
    const int K = g.KH * g.KW * g.Cin;
    const int wtn = g.Cout <= 32 ? 32 : 64;
    const int tiles = cdiv(K, 64 * 64 / wtn) * cdiv(g.Cout, wtn);
    int64_t s = (148 * 4 + tiles - 1) / tiles;
    const int64_t smax = std::max<int64_t>(1, M / 256);
    if (s > smax) s = smax;
    if (s > 512) s = 512;
    if (s < 1) s = 1;
    return (int)s;
}

// ------------------------------------------------------------------ column reductions (batch norm)
// rows x C matrix, two per-column sums in double:
//   mode 0: (x, x^2)              -> mean / variance
//   mode 1: (dyr, dyr * xhat)     -> dbeta / dgamma, dyr = dy masked by the ReLU of the forward output
struct ColArgs {
    const float* x;
    const float* dy;
    const float *mean, *invstd, *gamma, *beta;
    int64_t M;
    int C, relu, mode;
    float* maxout;      // mode 1, C == 128, optional: [blocks][128][2] = max |dyr|, max |xhat| per block and channel
};

__device__ __forceinline__ void col_terms(const ColArgs& a, int64_t row, int c, float& q1, float& q2) {
    const float x = a.x[row * a.C + c];
    if (a.mode == 0) {
        // shifted sums: x - pivot with pivot = the channel's value in row 0.  E[x^2] - mean^2 on raw float32 sums cancels
        // catastrophically when |mean| >> std; around a data point of the channel the sums stay O(std)
        q1 = x - a.x[c];
        q2 = q1 * q1;
    } else {
        const float xh = (x - a.mean[c]) * a.invstd[c];
        float dy = a.dy[row * a.C + c];
        if (a.relu && fmaf(xh, a.gamma[c], a.beta[c]) <= 0.f) dy = 0.f;
        q1 = dy;
        q2 = dy * xh;
    }
}

constexpr int CR_ROWS = 64;       // rows per block (8 per warp): enough blocks to fill 148 SMs at 51 200 rows

// A thread adds its 8 rows in float32 (8 terms: ~1e-7 relative), everything above that -- the 8 warps of the block,
// the blocks -- is summed in double in a fixed order.
__global__ void __launch_bounds__(256, 3) col_partial_kernel(ColArgs a, double* __restrict__ partial /* [blocks][C][2] */) {
    __shared__ float red[8][128][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * CR_ROWS;
    const int64_t r1 = min(a.M, r0 + CR_ROWS);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.C == 128) {
        // the 128-channel trunk layers: a lane owns channels 4 lane .. 4 lane + 3 and reads them as one 16-byte load per
        // row (a warp = one 512-byte row per instruction, all of a thread's rows in flight at once).  Every channel is still
        // summed by ONE thread over the same rows in the same order as in the scalar loop below: same bits.
        const int c0 = 4 * lane;
        float pv[4] = {0.f, 0.f, 0.f, 0.f}, mu[4], is[4], ga[4], be[4];
        float m1[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (a.mode == 0) {
                pv[u] = a.x[c0 + u];
            } else {
                mu[u] = a.mean[c0 + u];
                is[u] = a.invstd[c0 + u];
                ga[u] = a.gamma[c0 + u];
                be[u] = a.beta[c0 + u];
            }
        }
        // two batches of 4 rows: 8 (mode 0) / 16 (mode 1) 16-byte loads in flight per thread at 64 registers, 4 blocks per SM
#pragma unroll
        for (int kb = 0; kb < CR_ROWS / 8; kb += 4) {
        float4 xv[4], dv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t r = r0 + warp + 8 * (kb + k);
            dv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < r1) {
                xv[k] = *reinterpret_cast<const float4*>(a.x + r * 128 + c0);
                if (a.mode != 0) dv[k] = *reinterpret_cast<const float4*>(a.dy + r * 128 + c0);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t r = r0 + warp + 8 * (kb + k);
            if (r < r1) {
                const float xe[4] = {xv[k].x, xv[k].y, xv[k].z, xv[k].w};
                const float de[4] = {dv[k].x, dv[k].y, dv[k].z, dv[k].w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float q1, q2;
                    if (a.mode == 0) {
                        q1 = xe[u] - pv[u];
                        q2 = q1 * q1;
                    } else {
                        const float xh = (xe[u] - mu[u]) * is[u];
                        float dy = de[u];
                        if (a.relu && fmaf(xh, ga[u], be[u]) <= 0.f) dy = 0.f;
                        q1 = dy;
                        q2 = dy * xh;
                        m1[u] = fmaxf(m1[u], fabsf(dy));
                        m2[u] = fmaxf(m2[u], fabsf(xh));
                    }
                    s1[u] += q1;
                    s2[u] += q2;
                }
            }
        }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            red[warp][c0 + u][0] = s1[u];
            red[warp][c0 + u][1] = s2[u];
        }
        if (a.maxout) {          // block maxima through the same buffer, before the sums are consumed
            __shared__ float redm[8][128][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                redm[warp][c0 + u][0] = m1[u];
                redm[warp][c0 + u][1] = m2[u];
            }
            __syncthreads();
            if (threadIdx.x < 128) {
                float a1 = 0.f, a2 = 0.f;
                for (int w = 0; w < 8; ++w) {
                    a1 = fmaxf(a1, redm[w][threadIdx.x][0]);
                    a2 = fmaxf(a2, redm[w][threadIdx.x][1]);
                }
                a.maxout[((int64_t)blockIdx.x * 128 + threadIdx.x) * 2 + 0] = a1;
                a.maxout[((int64_t)blockIdx.x * 128 + threadIdx.x) * 2 + 1] = a2;
            }
        }
    } else {
        for (int64_t r = r0 + warp; r < r1; r += 8) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = lane + 32 * u;
                if (c < a.C) {
                    float q1, q2;
                    col_terms(a, r, c, q1, q2);
                    s1[u] += q1;
                    s2[u] += q2;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            red[warp][lane + 32 * u][0] = s1[u];
            red[warp][lane + 32 * u][1] = s2[u];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < a.C; c += 256) {
        double t1 = 0, t2 = 0;
        for (int w = 0; w < 8; ++w) {
            t1 += (double)red[w][c][0];
            t2 += (double)red[w][c][1];
        }
        partial[((int64_t)blockIdx.x * a.C + c) * 2 + 0] = t1;
        partial[((int64_t)blockIdx.x * a.C + c) * 2 + 1] = t2;
    }
}

// mode 0: mean, invstd (+ moving averages); mode 1: dbeta, dgamma.  One block per channel, fixed-order tree.
__global__ void __launch_bounds__(128) col_finalize_kernel(const double* __restrict__ partial, int blocks, int C, int64_t M, int mode,
                                                           float eps, float* __restrict__ o1, float* __restrict__ o2,
                                                           float* __restrict__ mov_mean, float* __restrict__ mov_var, float decay,
                                                           const float* __restrict__ pivot_row /* mode 0: x row 0 */,
                                                           const float* __restrict__ maxin = nullptr, const float* __restrict__ gamma = nullptr,
                                                           const float* __restrict__ invstd = nullptr, float* __restrict__ bound = nullptr) {
    __shared__ double r1[128], r2[128];
    __shared__ float rm[128][2];
    const int c = blockIdx.x;
    if (maxin) {        // mode 1: channel maxima of |dyr| and |xhat| over the blocks (order independent)
        float a1 = 0.f, a2 = 0.f;
        for (int b = threadIdx.x; b < blocks; b += 128) {
            a1 = fmaxf(a1, maxin[((int64_t)b * C + c) * 2 + 0]);
            a2 = fmaxf(a2, maxin[((int64_t)b * C + c) * 2 + 1]);
        }
        rm[threadIdx.x][0] = a1;
        rm[threadIdx.x][1] = a2;
    }
    double s1 = 0, s2 = 0;
    for (int b = threadIdx.x; b < blocks; b += 128) {
        s1 += partial[((int64_t)b * C + c) * 2 + 0];
        s2 += partial[((int64_t)b * C + c) * 2 + 1];
    }
    r1[threadIdx.x] = s1;
    r2[threadIdx.x] = s2;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            r1[threadIdx.x] += r1[threadIdx.x + o];
            r2[threadIdx.x] += r2[threadIdx.x + o];
            if (maxin) {
                rm[threadIdx.x][0] = fmaxf(rm[threadIdx.x][0], rm[threadIdx.x + o][0]);
                rm[threadIdx.x][1] = fmaxf(rm[threadIdx.x][1], rm[threadIdx.x + o][1]);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    s1 = r1[0];
    s2 = r2[0];
    if (maxin) {
        // |dx| = |gamma invstd (dyr - dbeta / M - xhat dgamma / M)| <= this, for every element of the channel
        const float inv_m = 1.f / (float)M;
        bound[c] = fabsf(gamma[c] * invstd[c]) * (rm[0][0] + fabsf((float)s1) * inv_m + rm[0][1] * fabsf((float)s2) * inv_m);
    }
    if (mode == 0) {
        const double dm = s1 / (double)M;                       // mean of (x - pivot)
        const double mean = (double)pivot_row[c] + dm;
        double var = s2 / (double)M - dm * dm;
        if (var < 0) var = 0;
        o1[c] = (float)mean;
        o2[c] = (float)(1.0 / sqrt(var + (double)eps));
        if (mov_mean) {
            // fused batch norm feeds the unbiased variance to the moving average (A.1)
            const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
            mov_mean[c] = decay * mov_mean[c] + (1.f - decay) * (float)mean;
            mov_var[c] = decay * mov_var[c] + (1.f - decay) * (float)unb;
        }
    } else {
        o1[c] = (float)s1;
        o2[c] = (float)s2;
    }
}

__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                                const float* __restrict__ res1, const float* __restrict__ res2, int64_t total, int C,
                                float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float v = fmaf((x[i] - mean[c]) * invstd[c], gamma[c], beta[c]);
        if (relu) v = fmaxf(v, 0.f);
        if (res1) v += res1[i];
        if (res2) v += res2[i];
        out[i] = v;
    }
}

// C % 4 == 0 (every layer of the step): four channels per thread, 16-byte accesses; per element the same operations as above
__global__ void __launch_bounds__(256) bn_apply4_kernel(const float4* __restrict__ x, const float* __restrict__ mean,
                                                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int relu, const float4* __restrict__ res1,
                                                        const float4* __restrict__ res2, int64_t total4, int C4,
                                                        float4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = 4 * (int)(i % C4);
        const float4 xv = x[i];
        const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u] = fmaf((xe[u] - mean[c + u]) * invstd[c + u], gamma[c + u], beta[c + u]);
            if (relu) v[u] = fmaxf(v[u], 0.f);
        }
        if (res1) {
            const float4 r = res1[i];
            v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        }
        if (res2) {
            const float4 r = res2[i];
            v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        }
        out[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// C % 8 == 0: eight channels per thread; besides the float32 output the kernel writes its fp16 hi/lo planes
// [plane][n][C/8][hw][8] -- the input format of the next tensor-core conv (conv_tc.cu), which then needs no split pass
__global__ void __launch_bounds__(256) bn_apply8_planes_kernel(const float4* __restrict__ x, const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, int relu, const float4* __restrict__ res1,
                                                               const float4* __restrict__ res2, int64_t total8, int CH, int64_t hw,
                                                               float4* __restrict__ out, __half* __restrict__ planes, int64_t plane) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
        const int chunk = (int)(i % CH);
        const int64_t r = i / CH;
        const int c = 8 * chunk;
        const float4 xa = x[2 * i], xb = x[2 * i + 1];
        const float xe[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            v[u] = fmaf((xe[u] - mean[c + u]) * invstd[c + u], gamma[c + u], beta[c + u]);
            if (relu) v[u] = fmaxf(v[u], 0.f);
        }
        if (res1) {
            const float4 a = res1[2 * i], b = res1[2 * i + 1];
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
        }
        if (res2) {
            const float4 a = res2[2 * i], b = res2[2 * i + 1];
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
        }
        if (out) {          // NULL: the only consumer is the next tensor-core conv, which reads the planes
            out[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
            out[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
        }
        float4 hi4, lo4;
        __half2* hi = reinterpret_cast<__half2*>(&hi4);
        __half2* lo = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            hi[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
            const float2 hf = __half22float2(hi[e]);
            lo[e] = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
        }
        const int64_t n = r / hw, rr = r - n * hw;
        const size_t off = (((size_t)n * CH + chunk) * hw + rr) * 8;
        *reinterpret_cast<float4*>(planes + off) = hi4;
        *reinterpret_cast<float4*>(planes + plane + off) = lo4;
    }
}

__global__ void __launch_bounds__(256) bn_bwd_apply4_kernel(const float4* __restrict__ x, const float4* __restrict__ dy,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ dbeta, const float* __restrict__ dgamma, int relu,
                                                            int affine_only, int64_t total4, int C4, float inv_m,
                                                            float4* __restrict__ dx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = 4 * (int)(i % C4);
        const float4 xv = x[i], dv = dy[i];
        const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
        const float de[4] = {dv.x, dv.y, dv.z, dv.w};
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float xh = (xe[u] - mean[c + u]) * invstd[c + u];
            float d = de[u];
            if (relu && fmaf(xh, gamma[c + u], beta[c + u]) <= 0.f) d = 0.f;
            if (!affine_only) d = d - dbeta[c + u] * inv_m - xh * dgamma[c + u] * inv_m;
            o[u] = gamma[c + u] * invstd[c + u] * d;
        }
        dx[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// The backward of a trunk layer's batch norm feeds exactly one consumer, the 3x3 conv's backward on tensor cores, which reads
// fp16 hi/lo planes pre-scaled by a power of two: this variant writes THOSE (no float32 dx, no maximum search, no split
// pass).  The scale comes from the per-channel bounds of col_finalize_kernel (>= the true maximum, typically within a few
// percent: the same binade or the next one) and is stored in scale_out = {s, 1/s} for the conv.
__global__ void __launch_bounds__(256) bn_bwd_apply8_planes_kernel(const float4* __restrict__ x, const float4* __restrict__ dy,
                                                                   const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   const float* __restrict__ dbeta, const float* __restrict__ dgamma,
                                                                   const float* __restrict__ bound, int relu, int64_t total8, int64_t hw,
                                                                   float inv_m, __half* __restrict__ planes, int64_t plane,
                                                                   float* __restrict__ scale_out) {
    __shared__ float sred[4];
    __shared__ float s_scale;
    if (threadIdx.x < 128) {
        float m = bound[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float mx = fmaxf(fmaxf(sred[0], sred[1]), fmaxf(sred[2], sred[3]));
        float sc = 1.f;
        if (mx > 0.f && isfinite(mx)) {         // largest magnitude into [2^5, 2^6): the rule of train_tc.cu for gradients
            int ex;
            frexpf(mx, &ex);
            sc = ldexpf(1.f, 6 - ex);
        }
        s_scale = sc;
        if (blockIdx.x == 0) {
            scale_out[0] = sc;
            scale_out[1] = 1.f / sc;
        }
    }
    __syncthreads();
    const float g = s_scale;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
        const int chunk = (int)(i & 15);
        const int64_t r = i >> 4;
        const int c = 8 * chunk;
        const float4 xa = x[2 * i], xb = x[2 * i + 1], da = dy[2 * i], db = dy[2 * i + 1];
        const float xe[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        const float de[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float xh = (xe[u] - mean[c + u]) * invstd[c + u];
            float d = de[u];
            if (relu && fmaf(xh, gamma[c + u], beta[c + u]) <= 0.f) d = 0.f;
            d = d - dbeta[c + u] * inv_m - xh * dgamma[c + u] * inv_m;
            v[u] = gamma[c + u] * invstd[c + u] * d * g;
        }
        float4 hi4, lo4;
        __half2* hi = reinterpret_cast<__half2*>(&hi4);
        __half2* lo = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            hi[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
            const float2 hf = __half22float2(hi[e]);
            lo[e] = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
        }
        const int64_t n = r / hw, rr = r - n * hw;
        const size_t off = (((size_t)n * 16 + chunk) * hw + rr) * 8;
        *reinterpret_cast<float4*>(planes + off) = hi4;
        *reinterpret_cast<float4*>(planes + plane + off) = lo4;
    }
}


// ---- tile-transposing forms of the two plane-writing kernels (C = 128).  The float32 side is NHWC (a pixel = 512 contiguous
// bytes), the planes are [chunk][pixel][8 ch]: with one thread per (pixel, chunk) one of the two sides is always accessed in
// 16- or 32-byte pieces at a 512-byte (or plane-sized) stride -- ncu: l1tex throughput 72 %, DRAM 11-23 % for a 28 us kernel.
// Here a block moves a tile of 32 pixels x 128 channels through shared memory: phase 1 reads / writes the float32 side with a
// warp per pixel (512 contiguous bytes per instruction), phase 2 writes the planes with a warp per chunk (32 pixels = 512
// contiguous bytes).  Shared layout [chunk][pixel][8] floats with a chunk pitch of 264 floats: both phases are conflict free.
constexpr int TT_PIX = 32, TT_PITCH = 264, TT_SMEM_FLOATS = 16 * TT_PITCH;

__device__ __forceinline__ void tt_store_planes(const float* __restrict__ tile, int64_t r0, int64_t M, int64_t hw,
                                                __half* __restrict__ planes, int64_t plane, float g) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int item = threadIdx.x + 256 * j, c = item >> 5, p = item & 31;
        const int64_t r = r0 + p;
        if (r >= M) continue;
        const float4 a = *reinterpret_cast<const float4*>(tile + c * TT_PITCH + p * 8);
        const float4 b = *reinterpret_cast<const float4*>(tile + c * TT_PITCH + p * 8 + 4);
        const float v[8] = {a.x * g, a.y * g, a.z * g, a.w * g, b.x * g, b.y * g, b.z * g, b.w * g};
        float4 hi4, lo4;
        __half2* hi = reinterpret_cast<__half2*>(&hi4);
        __half2* lo = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            hi[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
            const float2 hf = __half22float2(hi[e]);
            lo[e] = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
        }
        const int64_t n = r / hw, rr = r - n * hw;
        const size_t off = (((size_t)n * 16 + c) * hw + rr) * 8;
        *reinterpret_cast<float4*>(planes + off) = hi4;
        *reinterpret_cast<float4*>(planes + plane + off) = lo4;
    }
}

__global__ void __launch_bounds__(256) bn_apply_planes_tt_kernel(const float4* __restrict__ x, const float* __restrict__ mean,
                                                                 const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, int relu, const float4* __restrict__ res1,
                                                                 const float4* __restrict__ res2, int64_t M, int64_t hw,
                                                                 float4* __restrict__ out, __half* __restrict__ planes, int64_t plane) {
    // persistent over tiles; the loads of tile i + 1 are issued before the plane stores of tile i (register prefetch): a
    // one-tile-per-block version spent most of its 16 us in exposed load latency (ncu: DRAM 20 %, l1tex 46 %)
    __shared__ __align__(16) float tile[TT_SMEM_FLOATS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t ntiles = (M + TT_PIX - 1) / TT_PIX;
    float mu[4], is[4], ga[4], be[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        mu[u] = mean[4 * lane + u];
        is[u] = invstd[4 * lane + u];
        ga[u] = gamma[4 * lane + u];
        be[u] = beta[4 * lane + u];
    }
    float4 xr[4], q1[4], q2[4];
    auto fetch = [&](int64_t t) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t r = t * TT_PIX + w * 4 + j;
            if (r < M) {
                const int64_t i4 = r * 32 + lane;
                xr[j] = x[i4];
                if (res1) q1[j] = res1[i4];
                if (res2) q2[j] = res2[i4];
            }
        }
    };
    int64_t t = blockIdx.x;
    if (t < ntiles) fetch(t);
    for (; t < ntiles; t += gridDim.x) {
        const int64_t r0 = t * TT_PIX;
        float4 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float xe[4] = {xr[j].x, xr[j].y, xr[j].z, xr[j].w};
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = fmaf((xe[u] - mu[u]) * is[u], ga[u], be[u]);
                if (relu) v[u] = fmaxf(v[u], 0.f);
            }
            if (res1) {
                v[0] += q1[j].x; v[1] += q1[j].y; v[2] += q1[j].z; v[3] += q1[j].w;
            }
            if (res2) {
                v[0] += q2[j].x; v[1] += q2[j].y; v[2] += q2[j].z; v[3] += q2[j].w;
            }
            o[j] = make_float4(v[0], v[1], v[2], v[3]);
        }
        if (t + gridDim.x < ntiles) fetch(t + gridDim.x);         // next tile's loads fly during this tile's stores
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = w * 4 + j;
            const int64_t r = r0 + p;
            if (r < M) {
                if (out) out[r * 32 + lane] = o[j];
                *reinterpret_cast<float4*>(tile + (lane >> 1) * TT_PITCH + p * 8 + (lane & 1) * 4) = o[j];
            }
        }
        __syncthreads();
        tt_store_planes(tile, r0, M, hw, planes, plane, 1.f);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_planes_tt_kernel(const float4* __restrict__ x, const float4* __restrict__ dy,
                                                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                     const float* __restrict__ dbeta, const float* __restrict__ dgamma,
                                                                     const float* __restrict__ bound, int relu, int64_t M, int64_t hw,
                                                                     float inv_m, __half* __restrict__ planes, int64_t plane,
                                                                     float* __restrict__ scale_out) {
    __shared__ __align__(16) float tile[TT_SMEM_FLOATS];
    __shared__ float sred[4];
    __shared__ float s_scale;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x < 128) {
        float m = bound[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
        if (lane == 0) sred[w] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float mx = fmaxf(fmaxf(sred[0], sred[1]), fmaxf(sred[2], sred[3]));
        float sc = 1.f;
        if (mx > 0.f && isfinite(mx)) {         // largest magnitude into [2^5, 2^6): the rule of train_tc.cu for gradients
            int ex;
            frexpf(mx, &ex);
            sc = ldexpf(1.f, 6 - ex);
        }
        s_scale = sc;
        if (blockIdx.x == 0) {
            scale_out[0] = sc;
            scale_out[1] = 1.f / sc;
        }
    }
    const int64_t ntiles = (M + TT_PIX - 1) / TT_PIX;
    float mu[4], is[4], ga[4], be[4], db[4], dg[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        mu[u] = mean[4 * lane + u];
        is[u] = invstd[4 * lane + u];
        ga[u] = gamma[4 * lane + u];
        be[u] = beta[4 * lane + u];
        db[u] = dbeta[4 * lane + u];
        dg[u] = dgamma[4 * lane + u];
    }
    // persistent over tiles with a register prefetch of the next tile (see bn_apply_planes_tt_kernel)
    float4 xr[4], dr[4];
    auto fetch = [&](int64_t t) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t r = t * TT_PIX + w * 4 + j;
            if (r < M) {
                xr[j] = x[r * 32 + lane];
                dr[j] = dy[r * 32 + lane];
            }
        }
    };
    int64_t t = blockIdx.x;
    if (t < ntiles) fetch(t);
    for (; t < ntiles; t += gridDim.x) {
        const int64_t r0 = t * TT_PIX;
        float4 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float xe[4] = {xr[j].x, xr[j].y, xr[j].z, xr[j].w};
            const float de[4] = {dr[j].x, dr[j].y, dr[j].z, dr[j].w};
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float xh = (xe[u] - mu[u]) * is[u];
                float d = de[u];
                if (relu && fmaf(xh, ga[u], be[u]) <= 0.f) d = 0.f;
                d = d - db[u] * inv_m - xh * dg[u] * inv_m;
                v[u] = ga[u] * is[u] * d;
            }
            o[j] = make_float4(v[0], v[1], v[2], v[3]);
        }
        if (t + gridDim.x < ntiles) fetch(t + gridDim.x);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = w * 4 + j;
            if (r0 + p < M) *reinterpret_cast<float4*>(tile + (lane >> 1) * TT_PITCH + p * 8 + (lane & 1) * 4) = o[j];
        }
        __syncthreads();          // also orders s_scale (written by thread 0 above) before its first use
        tt_store_planes(tile, r0, M, hw, planes, plane, s_scale);
        __syncthreads();
    }
}

// dx = gamma * invstd * (dyr - dbeta / M - xhat * dgamma / M); affine_only: dx = gamma * invstd * dyr
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ dbeta,
                                    const float* __restrict__ dgamma, int relu, int affine_only, int64_t total, int C,
                                    float inv_m, float* __restrict__ dx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float xh = (x[i] - mean[c]) * invstd[c];
        float d = dy[i];
        if (relu && fmaf(xh, gamma[c], beta[c]) <= 0.f) d = 0.f;
        if (!affine_only) d = d - dbeta[c] * inv_m - xh * dgamma[c] * inv_m;
        dx[i] = gamma[c] * invstd[c] * d;
    }
}

// ------------------------------------------------------------------ heatmap + soft quantizer, backward
// per latent pixel (n, y, x): bn[0] heatmap logit, bn[1 + c] features.
//   hm2d = sigmoid(bn0) * C ; hm_c = max(min(hm2d - c, 1), 0) ; z_c = hm_c * bn[1 + c]
//   qsoft_c = sum_j softmax_j(-(z_c - ctr_j)^2) ctr_j ; qbar = qsoft + stop_gradient(qhard - qsoft)
// in: dq (NHWC, C) gradient w.r.t. qbar, dhm (NCHW, optional) extra gradient w.r.t. hm (from H_mask)
// out: dbn (NHWC, Cb >= C + 1 channels, padding channels zeroed), per-block partial sums of dcenters
constexpr int HQ_THREADS = 128;

__global__ void __launch_bounds__(HQ_THREADS) hq_bwd_kernel(const float* __restrict__ bn, int Cb, int C, int heatmap,
                                                           const float* __restrict__ centers, int L,
                                                           const float* __restrict__ dq, const float* __restrict__ dhm, int h,
                                                           int w, int64_t npix, float* __restrict__ dbn,
                                                           double* __restrict__ dcent_partial /* [blocks][8] */) {
    __shared__ double red[HQ_THREADS / 32][8];
    float ctr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ctr[j] = j < L ? centers[j] : 0.f;
    double dc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npix) {
        const float* b = bn + p * Cb;
        float* db = dbn + p * Cb;
        const int64_t n = p / ((int64_t)h * w), yx = p - n * (int64_t)h * w;
        float sg = 0.f, hm2d = 0.f;
        if (heatmap) {
            sg = 1.f / (1.f + expf(-b[0]));
            hm2d = sg * (float)C;
        }
        float dhm2d = 0.f;
        const int f0 = heatmap ? 1 : 0;
        for (int c = 0; c < C; ++c) {
            float hmc = 1.f;
            const float t = hm2d - (float)c;
            if (heatmap) hmc = fmaxf(fminf(t, 1.f), 0.f);
            const float feat = b[f0 + c];
            const float z = hmc * feat;
            // softmax over -(z - ctr_j)^2
            float e[8], m = -3.4e38f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < L) {
                    const float dd = z - ctr[j];
                    e[j] = -dd * dd;
                    m = fmaxf(m, e[j]);
                }
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < L) {
                    e[j] = expf(e[j] - m);
                    s += e[j];
                }
            float qs = 0.f, gbar = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < L) {
                    e[j] /= s;                               // p_j
                    qs += e[j] * ctr[j];
                    gbar += e[j] * (-2.f * (z - ctr[j]));    // sum_k p_k g_k, g_k = d(-(z - c_k)^2)/dz
                }
            const float g = dq[p * C + c];
            float dz = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < L) {
                    const float gj = -2.f * (z - ctr[j]);
                    dz += ctr[j] * e[j] * (gj - gbar);
                    // d qsoft / d ctr_j = p_j + 2 (z - ctr_j) p_j (ctr_j - qsoft)
                    dc[j] += (double)(g * (e[j] + 2.f * (z - ctr[j]) * e[j] * (ctr[j] - qs)));
                }
            dz *= g;
            db[f0 + c] = dz * hmc;
            if (heatmap) {
                float dh = dz * feat;
                if (dhm) dh += dhm[(n * C + c) * (int64_t)h * w + yx];
                if (t >= 0.f && t <= 1.f) dhm2d += dh;      // gradient of maximum(minimum(t, 1), 0)
            }
        }
        if (heatmap) db[0] = dhm2d * (float)C * sg * (1.f - sg);
        for (int c = f0 + C; c < Cb; ++c) db[c] = 0.f;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        double v = dc[j];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double v = 0;
        for (int wv = 0; wv < HQ_THREADS / 32; ++wv) v += red[wv][threadIdx.x];
        dcent_partial[(int64_t)blockIdx.x * 8 + threadIdx.x] = v;
    }
}

__global__ void sum_double_partials_kernel(const double* __restrict__ partial, int blocks, int width, int n_out,
                                           float* __restrict__ out) {
    const int j = threadIdx.x;
    if (j >= n_out) return;
    double s = 0;
    for (int b = 0; b < blocks; ++b) s += partial[(int64_t)b * width + j];
    out[j] = (float)s;
}

// ------------------------------------------------------------------ image range: denormalise + clip, and back
__constant__ float t_norm_mean[4] = {121.853699f, 113.588608f, 100.637154f, 0.f};
__constant__ float t_norm_std[4] = {68.8939514f, 66.7393417f, 69.3702698f, 0.f};

// v NHWC (4 channels, last is padding) -> x_out NCHW float [0,255] (code/autoencoder.py:146-158)
__global__ void denorm_clip_fwd_kernel(const float* __restrict__ v, int64_t npix_img, int64_t total_pix, float* __restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total_pix; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = p / npix_img, r = p - n * npix_img;
        const float4 q = *reinterpret_cast<const float4*>(v + p * 4);
        const float qq[3] = {q.x, q.y, q.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float t = __fadd_rn(__fmul_rn(qq[c], t_norm_std[c]), t_norm_mean[c]);
            out[(n * 3 + c) * npix_img + r] = fminf(fmaxf(t, 0.f), 255.f);
        }
    }
}

// dx_out NCHW -> dv NHWC4: passes where the un-clipped value lies in [0,255] (gradient of clip_by_value)
__global__ void denorm_clip_bwd_kernel(const float* __restrict__ v, const float* __restrict__ dout, int64_t npix_img,
                                       int64_t total_pix, float* __restrict__ dv) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total_pix; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = p / npix_img, r = p - n * npix_img;
        const float4 q = *reinterpret_cast<const float4*>(v + p * 4);
        const float qq[3] = {q.x, q.y, q.z};
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float t = __fadd_rn(__fmul_rn(qq[c], t_norm_std[c]), t_norm_mean[c]);
            if (t >= 0.f && t <= 255.f) o[c] = dout[(n * 3 + c) * npix_img + r] * t_norm_std[c];
        }
        *reinterpret_cast<float4*>(dv + p * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ------------------------------------------------------------------ layout + arithmetic helpers
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, int C, int Cs, int64_t hw, int64_t total, float* __restrict__ out) {
    // in N,hw,Cs (first C channels used) -> out N,C,hw
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / (C * hw), r = i - n * C * hw;
        const int c = (int)(r / hw);
        const int64_t p = r - c * hw;
        out[i] = in[(n * hw + p) * Cs + c];
    }
}

__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ in, int C, int Cs, int64_t hw, int64_t total, float* __restrict__ out) {
    // in N,C,hw -> out N,hw,Cs (channels >= C zero)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / (Cs * hw), r = i - n * Cs * hw;
        const int64_t p = r / Cs;
        const int c = (int)(r - p * Cs);
        out[i] = c < C ? in[(n * C + c) * hw + p] : 0.f;
    }
}

__global__ void axpby_kernel(float a, const float* __restrict__ x, float b, const float* __restrict__ y, int64_t n,
                             float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a * x[i] + (y ? b * y[i] : 0.f);
}

// out = a[0] * x
__global__ void scale_dev_kernel(const float* __restrict__ a, const float* __restrict__ x, int64_t n, float* __restrict__ out) {
    const float s = a[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = s * x[i];
}

// d pc_loss / d bc coefficient of code/train.py:309-316: beta * 0.5 / n when 0.5 * (H_mask + H_real) > H_target, else 0
__global__ void rate_coef_kernel(const double* __restrict__ sums, double n, float beta, float h_target, int has_heatmap,
                                 float* __restrict__ coef) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double h_real = sums[0] / n, h_mask = has_heatmap ? sums[1] / n : h_real;
    coef[0] = (0.5 * (h_mask + h_real) > (double)h_target) ? (float)((double)beta * 0.5 / n) : 0.f;
}

__global__ void mul_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = x[i] * y[i];
}

// tf.train.AdamOptimizer._apply_dense; grad_scale folds the regulariser: g = grad + l2 * w
__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, float lr_t, float beta1, float beta2, float eps, float l2, const float* __restrict__ mask,
                            const float* __restrict__ lr_ptr) {
    if (lr_ptr) lr_t = lr_ptr[0];      // bias-corrected step size written by the host before a CUDA-graph replay
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] + l2 * w[i];
        if (mask) gi *= mask[i];
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        w[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

inline int ew_grid(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, 148 * 16); }


// ------------------------------------------------------------------ context model in training mode
// The latent volume is laid out depth-major: (D, N, H, W, C) -- one "image" per (depth slice, batch element) --
// so that the (2,3,3) VALID conv3d of code/probclass.py:227-261 is two VALID conv2d passes over the contiguous
// slice ranges [0, D-1) and [1, D) (filter depth 0 and 1), accumulated.
// pad_for_probclass3d (code/probclass.py:268-292): q NCHW -> (C+4, N, h+8, w+8, 4), channel 0 = value, 1..3 = 0
__global__ void pc_pad_kernel(const float* __restrict__ q, int N, int C, int h, int w, float pad_value,
                              const float* __restrict__ pad_ptr, int64_t total, float4* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    if (pad_ptr) pad_value = pad_ptr[0];
    const int Wp = w + 8, Hp = h + 8;
    const int x = (int)(i % Wp);
    int64_t r = i / Wp;
    const int y = (int)(r % Hp);
    r /= Hp;
    const int n = (int)(r % N);
    const int d = (int)(r / N);
    float v = pad_value;
    if (d >= 4 && y >= 4 && y < h + 4 && x >= 4 && x < w + 4)
        v = q[(((int64_t)n * C + (d - 4)) * h + (y - 4)) * w + (x - 4)];
    out[i] = make_float4(v, 0.f, 0.f, 0.f);
}

// softmax_cross_entropy_with_logits * log2(e) (code/probclass.py:99-104).  logits rows in (C, N, h, w) order with
// Cs floats per row (first L valid); symbols / heatmap / bc in the reference's NCHW order.
template <bool BWD>
__global__ void pc_xent_kernel(const float* __restrict__ logits, int Cs, int L, const int64_t* __restrict__ symbols,
                               const float* __restrict__ heatmap, int N, int C, int h, int w, float coef_real, float coef_mask,
                               const float* __restrict__ coef_ptr, int64_t total, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // row in (C, N, h, w) order
    if (i >= total) return;
    const int64_t hw = (int64_t)h * w;
    const int64_t yx = i % hw;
    int64_t r = i / hw;
    const int n = (int)(r % N);
    const int c = (int)(r / N);
    const int64_t j = ((int64_t)n * C + c) * hw + yx;                     // NCHW index
    const float* lg = logits + i * Cs;
    float m = lg[0];
    for (int k = 1; k < L; ++k) m = fmaxf(m, lg[k]);
    float s = 0.f;
    for (int k = 0; k < L; ++k) s += expf(lg[k] - m);
    const int sym = (int)symbols[j];
    const float log2e = 1.4426950408889634f;
    if (!BWD) {
        out[j] = (logf(s) - (lg[sym] - m)) * log2e;
    } else {
        if (coef_ptr) coef_real = coef_mask = coef_ptr[0];
        const float g = (coef_real + (heatmap ? coef_mask * heatmap[j] : coef_mask)) * log2e;
        float* o = out + i * Cs;
        for (int k = 0; k < Cs; ++k) o[k] = k < L ? g * (expf(lg[k] - m) / s - (k == sym ? 1.f : 0.f)) : 0.f;
    }
}

// out[a][y][x][:] = in[a][y + crop][x + crop][:]   /   dx[a][y + crop][x + crop][:] += dy[a][y][x][:]
template <bool BWD>
__global__ void crop_kernel(const float4* __restrict__ src, int H, int W, int C4, int crop, int64_t total, float4* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // over the cropped tensor (float4 units)
    if (i >= total) return;
    const int Ho = H - 2 * crop, Wo = W - 2 * crop;
    const int c = (int)(i % C4);
    int64_t r = i / C4;
    const int x = (int)(r % Wo);
    r /= Wo;
    const int y = (int)(r % Ho);
    const int64_t a = r / Ho;
    const int64_t big = ((a * H + y + crop) * W + x + crop) * C4 + c;
    if (!BWD) {
        dst[i] = src[big];
    } else {
        float4 v = dst[big];
        const float4 d = src[i];
        v.x += d.x;
        v.y += d.y;
        v.z += d.z;
        v.w += d.w;
        dst[big] = v;
    }
}

}  // namespace
}  // namespace ic

using namespace ic;

extern "C" {

size_t ic_nn_conv2d_workspace_bytes(int N, int Hi, int Wi, int Cin, int KH, int KW, int stride, int Cout, int transposed, int valid) {
    Geo g{N, Hi, Wi, Cin, KH, KW, stride, Cout, transposed, valid, 0, 0, 0, 0};
    if (make_geo(g) != IC_OK) return 0;
    const size_t wbytes = (size_t)KH * KW * Cin * Cout * sizeof(float);
    int splits = wgrad_splits(g);
    if (transposed) {       // the filter gradient runs on the swapped (regular strided conv) geometry
        Geo g2{N, g.Ho, g.Wo, Cout, KH, KW, stride, Cin, 0, 0, 0, 0, 0, 0};
        if (make_geo(g2) == IC_OK) splits = std::max(splits, wgrad_splits(g2));
    }
    return 2 * align_up(wbytes, 256) + align_up((size_t)splits * wbytes, 256) + 256;
}

int ic_nn_conv2d_fwd(const float* d_x, const float* d_w, int N, int Hi, int Wi, int Cin, int KH, int KW, int stride, int Cout,
                     int transposed, int valid, float* d_y, void* stream) {
    IC_REQUIRE(d_x && d_w && d_y, IC_ERR_INVALID, "ic_nn_conv2d_fwd: NULL argument");
    Geo g{N, Hi, Wi, Cin, KH, KW, stride, Cout, transposed, valid, 0, 0, 0, 0};
    int rc = make_geo(g);
    if (rc != IC_OK) return rc;
    rc = ensure_consts();
    if (rc != IC_OK) return rc;
    ConvDesc d;
    memset(&d, 0, sizeof(d));
    d.in = d_x; d.w = d_w; d.scale = g_ones; d.shift = g_zeros; d.out = d_y;
    d.N = N; d.Hi = Hi; d.Wi = Wi; d.Cin = Cin; d.Ho = g.Ho; d.Wo = g.Wo; d.Cout = Cout; d.ldw = Cout;
    d.KH = KH; d.KW = KW; d.stride = stride; d.pad_t = g.pad_t; d.pad_l = g.pad_l; d.transposed = transposed;
    return launch_conv_simt(d, (cudaStream_t)stream);
}

int ic_nn_conv2d_bwd_data(const float* d_dy, const float* d_w, int N, int Hi, int Wi, int Cin, int KH, int KW, int stride,
                          int Cout, int transposed, int valid, float* d_dx, void* d_workspace, size_t workspace_bytes,
                          void* stream) {
    IC_REQUIRE(d_dy && d_w && d_dx && d_workspace, IC_ERR_INVALID, "ic_nn_conv2d_bwd_data: NULL argument");
    Geo g{N, Hi, Wi, Cin, KH, KW, stride, Cout, transposed, valid, 0, 0, 0, 0};
    int rc = make_geo(g);
    if (rc != IC_OK) return rc;
    rc = ensure_consts();
    if (rc != IC_OK) return rc;
    const size_t wcount = (size_t)KH * KW * Cin * Cout;
    IC_REQUIRE(workspace_bytes >= wcount * sizeof(float), IC_ERR_WORKSPACE, "ic_nn_conv2d_bwd_data: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    float* wT = (float*)d_workspace;
    transpose_taps_kernel<<<ew_grid((int64_t)wcount), 256, 0, s>>>(d_w, wT, KH * KW, Cin, Cout);
    IC_CHECK_LAUNCH();
    ConvDesc d;
    memset(&d, 0, sizeof(d));
    d.in = d_dy; d.w = wT; d.scale = g_ones; d.shift = g_zeros; d.out = d_dx;
    d.N = N; d.Hi = g.Ho; d.Wi = g.Wo; d.Cin = Cout; d.Ho = Hi; d.Wo = Wi; d.Cout = Cin; d.ldw = Cin;
    d.KH = KH; d.KW = KW; d.stride = stride; d.pad_t = g.pad_t; d.pad_l = g.pad_l;
    d.transposed = transposed ? 0 : 1;       // the gradient of a conv is the other kind of conv
    return launch_conv_simt(d, s);
}

int ic_nn_conv2d_bwd_filter(const float* d_x, const float* d_dy, int N, int Hi, int Wi, int Cin, int KH, int KW, int stride,
                            int Cout, int transposed, int valid, float* d_dw, void* d_workspace, size_t workspace_bytes,
                            void* stream) {
    IC_REQUIRE(d_x && d_dy && d_dw && d_workspace, IC_ERR_INVALID, "ic_nn_conv2d_bwd_filter: NULL argument");
    Geo g{N, Hi, Wi, Cin, KH, KW, stride, Cout, transposed, valid, 0, 0, 0, 0};
    int rc = make_geo(g);
    if (rc != IC_OK) return rc;
    const int64_t count = (int64_t)KH * KW * Cin * Cout;
    cudaStream_t s = (cudaStream_t)stream;
    // conv2d_transpose: y[i*s - pad + tap] += x[i] w[tap], so dW[tap][ci][co] = sum_i x[i][ci] dy[i*s - pad + tap][co]:
    // the filter gradient of the REGULAR strided conv dy-shaped -> x-shaped with the roles of the tensors swapped and
    // the two channel axes transposed.  Iterating over the (4x fewer) input pixels i visits only taps that exist,
    // where the output-pixel form would multiply 3 zeros out of 4.
    Geo gr = g;
    const float *a_src = d_x, *b_src = d_dy;
    if (transposed) {
        gr = Geo{N, g.Ho, g.Wo, Cout, KH, KW, stride, Cin, 0, 0, 0, 0, 0, 0};
        rc = make_geo(gr);
        if (rc != IC_OK) return rc;
        IC_REQUIRE(gr.Ho == Hi && gr.Wo == Wi, IC_ERR_STATE, "ic_nn_conv2d_bwd_filter: transposed geometry");
        a_src = d_dy;
        b_src = d_x;
    }
    const int splits = wgrad_splits(gr);
    const size_t wb = align_up((size_t)count * sizeof(float), 256);
    IC_REQUIRE(workspace_bytes >= wb + (size_t)splits * count * sizeof(float), IC_ERR_WORKSPACE,
               "ic_nn_conv2d_bwd_filter: workspace too small (use ic_nn_conv2d_workspace_bytes)");
    float* tmp = (float*)d_workspace;                               // [tap][Cout][Cin] before the channel transpose
    float* part = (float*)((char*)d_workspace + wb);
    WgradDesc d;
    d.x = a_src; d.dy = b_src; d.part = part; d.g = gr;
    d.M = (int64_t)N * gr.Ho * gr.Wo;
    d.m_per_split = (d.M + splits - 1) / splits;
    d.m_per_split = (d.m_per_split + WM - 1) / WM * WM;
    const int K = KH * KW * gr.Cin;
    const int wtn = gr.Cout <= 32 ? 32 : 64;
    dim3 grid(cdiv(K, 64 * 64 / wtn), cdiv(gr.Cout, wtn), splits);
    {
        ProfScope ps(IC_PROF_CONV_OTHER, s, transposed ? 3 : 2);
        if (wtn == 32) conv_wgrad_kernel<32><<<grid, 256, 0, s>>>(d);
        else conv_wgrad_kernel<64><<<grid, 256, 0, s>>>(d);
        IC_CHECK_LAUNCH();
        sum_splits_kernel<<<ew_grid(count), 256, 0, s>>>(part, splits, count, transposed ? tmp : d_dw);
        IC_CHECK_LAUNCH();
        if (transposed) {
            transpose_taps_kernel<<<ew_grid(count), 256, 0, s>>>(tmp, d_dw, KH * KW, Cout, Cin);
            IC_CHECK_LAUNCH();
        }
    }
    return IC_OK;
}

size_t ic_nn_bn_workspace_bytes(int64_t M, int C) {
    if (M <= 0 || C <= 0 || C > 128) return 0;
    const int64_t blocks = (M + CR_ROWS - 1) / CR_ROWS;
    // partial sums (double) + the block maxima and channel bounds of ic_nn_bn_train_bwd_ex (float, C = 128)
    return (size_t)blocks * C * 2 * sizeof(double) + (size_t)blocks * 128 * 2 * sizeof(float) + 128 * sizeof(float) + 256;
}

/* slim.batch_norm(is_training=True): statistics of x (M rows, C channels), then
 * out = relu?((x - mean) * invstd * gamma + beta) (+ res1) (+ res2).  d_mean / d_invstd are saved
 * for the backward pass; d_mov_mean / d_mov_var (optional) get the decay-0.9 moving-average update.
 * gamma == NULL / beta == NULL: identity statistics are not computed -- pass use_stats = 0 for a
 * plain affine layer y = relu?(x * gamma + beta) with d_mean = 0, d_invstd = 1 supplied by the caller. */
int ic_nn_bn_train_fwd(const float* d_x, int64_t M, int C, const float* d_gamma, const float* d_beta, float eps, int relu,
                       int use_stats, const float* d_res1, const float* d_res2, float* d_mean, float* d_invstd,
                       float* d_mov_mean, float* d_mov_var, float* d_out, void* d_workspace, size_t workspace_bytes, void* stream) {
    return ic_nn_bn_train_fwd_ex(d_x, M, C, d_gamma, d_beta, eps, relu, use_stats, d_res1, d_res2, d_mean, d_invstd, d_mov_mean,
                                 d_mov_var, d_out, nullptr, nullptr, 0, d_workspace, workspace_bytes, stream);
}

/* ic_nn_bn_train_fwd with the two fusions of the tensor-core trunk: d_partial_in (optional) = the statistics partial sums
 * ic_nn_conv3x3_tc_fused accumulated while it wrote d_x (skips the statistics pass over d_x); d_planes_out (optional, C % 8 == 0,
 * 2 M C fp16 elements) receives the fp16 hi/lo planes [plane][M / hw][C / 8][hw][8] of d_out for the next conv; d_out may
 * then be NULL (no float32 copy is written: for an output whose only consumer is that conv). */
int ic_nn_bn_train_fwd_ex(const float* d_x, int64_t M, int C, const float* d_gamma, const float* d_beta, float eps, int relu,
                          int use_stats, const float* d_res1, const float* d_res2, float* d_mean, float* d_invstd,
                          float* d_mov_mean, float* d_mov_var, float* d_out, const double* d_partial_in, void* d_planes_out,
                          int64_t hw, void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(d_x && d_gamma && d_beta && d_mean && d_invstd && (d_out || d_planes_out), IC_ERR_INVALID, "ic_nn_bn_train_fwd: NULL argument");
    IC_REQUIRE(M > 0 && C > 0 && C <= 128, IC_ERR_INVALID, "ic_nn_bn_train_fwd: bad shape (C <= 128)");
    IC_REQUIRE(!d_partial_in || C == 128, IC_ERR_INVALID, "ic_nn_bn_train_fwd_ex: partial sums are those of a 128-channel conv");
    IC_REQUIRE(!d_planes_out || (C % 8 == 0 && hw > 0 && M % hw == 0), IC_ERR_INVALID, "ic_nn_bn_train_fwd_ex: planes need C % 8 == 0 and M = n hw");
    cudaStream_t s = (cudaStream_t)stream;
    if (use_stats) {
        const int blocks = (int)((M + CR_ROWS - 1) / CR_ROWS);
        const double* partial = d_partial_in;
        if (!partial) {
            IC_REQUIRE(d_workspace && workspace_bytes >= ic_nn_bn_workspace_bytes(M, C), IC_ERR_WORKSPACE, "ic_nn_bn_train_fwd: workspace");
            ColArgs a;
            memset(&a, 0, sizeof(a));
            a.x = d_x; a.M = M; a.C = C; a.mode = 0;
            col_partial_kernel<<<blocks, 256, 0, s>>>(a, (double*)d_workspace);
            IC_CHECK_LAUNCH();
            partial = (const double*)d_workspace;
        }
        col_finalize_kernel<<<C, 128, 0, s>>>(partial, blocks, C, M, 0, eps, d_mean, d_invstd, d_mov_mean, d_mov_var, 0.9f, d_x);
        IC_CHECK_LAUNCH();
    }
    if (d_planes_out) {
        IC_REQUIRE((((uintptr_t)d_x | (uintptr_t)d_out | (uintptr_t)d_res1 | (uintptr_t)d_res2 | (uintptr_t)d_planes_out) & 15) == 0,
                   IC_ERR_INVALID, "ic_nn_bn_train_fwd_ex: unaligned tensor");
        static const bool tt = !(getenv("IC_TRAIN_TT") && atoi(getenv("IC_TRAIN_TT")) == 0);      // 0: one thread per (pixel, chunk)
        if (C == 128 && tt)
            bn_apply_planes_tt_kernel<<<(unsigned)std::min<int64_t>((M + TT_PIX - 1) / TT_PIX, 148 * 2), 256, 0, s>>>((const float4*)d_x, d_mean, d_invstd, d_gamma, d_beta,
                                                                                      relu, (const float4*)d_res1, (const float4*)d_res2, M, hw,
                                                                                      (float4*)d_out, (__half*)d_planes_out, M * C);
        else
            bn_apply8_planes_kernel<<<ew_grid(M * C / 8), 256, 0, s>>>((const float4*)d_x, d_mean, d_invstd, d_gamma, d_beta, relu,
                                                                       (const float4*)d_res1, (const float4*)d_res2, M * C / 8, C / 8, hw,
                                                                       (float4*)d_out, (__half*)d_planes_out, M * C);
        IC_CHECK_LAUNCH();
        return IC_OK;
    }
    const bool vec4 = C % 4 == 0 && (((uintptr_t)d_x | (uintptr_t)d_out | (uintptr_t)d_res1 | (uintptr_t)d_res2) & 15) == 0;
    if (vec4)
        bn_apply4_kernel<<<ew_grid(M * C / 4), 256, 0, s>>>((const float4*)d_x, d_mean, d_invstd, d_gamma, d_beta, relu,
                                                            (const float4*)d_res1, (const float4*)d_res2, M * C / 4, C / 4, (float4*)d_out);
    else
        bn_apply_kernel<<<ew_grid(M * C), 256, 0, s>>>(d_x, d_mean, d_invstd, d_gamma, d_beta, relu, d_res1, d_res2, M * C, C, d_out);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

/* backward of ic_nn_bn_train_fwd w.r.t. x, gamma, beta (the residual inputs receive d_dy unchanged). */
int ic_nn_bn_train_bwd(const float* d_x, const float* d_dy, int64_t M, int C, const float* d_gamma, const float* d_beta, int relu,
                       int use_stats, const float* d_mean, const float* d_invstd, float* d_dx, float* d_dgamma, float* d_dbeta,
                       void* d_workspace, size_t workspace_bytes, void* stream) {
    return ic_nn_bn_train_bwd_ex(d_x, d_dy, M, C, d_gamma, d_beta, relu, use_stats, d_mean, d_invstd, d_dx, d_dgamma, d_dbeta, nullptr,
                                 nullptr, 0, d_workspace, workspace_bytes, stream);
}

/* ic_nn_bn_train_bwd whose dx goes to a tensor-core conv backward: d_dx_planes (C = 128, batch statistics; 2 M C fp16
 * elements) receives the fp16 hi/lo planes [plane][M / hw][16][hw][8] of dx * s, d_scale_out = {s, 1/s} with s the power of
 * two that brings an upper bound of max |dx| into [2^5, 2^6); d_dx may then be NULL (no float32 copy is written). */
int ic_nn_bn_train_bwd_ex(const float* d_x, const float* d_dy, int64_t M, int C, const float* d_gamma, const float* d_beta, int relu,
                          int use_stats, const float* d_mean, const float* d_invstd, float* d_dx, float* d_dgamma, float* d_dbeta,
                          void* d_dx_planes, float* d_scale_out, int64_t hw, void* d_workspace, size_t workspace_bytes,
                          void* stream) {
    IC_REQUIRE(d_x && d_dy && d_gamma && d_beta && d_mean && d_invstd && (d_dx || d_dx_planes) && d_dgamma && d_dbeta && d_workspace,
               IC_ERR_INVALID, "ic_nn_bn_train_bwd: NULL argument");
    IC_REQUIRE(M > 0 && C > 0 && C <= 128, IC_ERR_INVALID, "ic_nn_bn_train_bwd: bad shape (C <= 128)");
    IC_REQUIRE(workspace_bytes >= ic_nn_bn_workspace_bytes(M, C), IC_ERR_WORKSPACE, "ic_nn_bn_train_bwd: workspace");
    IC_REQUIRE(!d_dx_planes || (C == 128 && use_stats && d_scale_out && hw > 0 && M % hw == 0 && !d_dx), IC_ERR_INVALID,
               "ic_nn_bn_train_bwd_ex: planes need C = 128, batch statistics, M = n hw and no float32 dx");
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = (int)((M + CR_ROWS - 1) / CR_ROWS);
    // workspace: [blocks][C][2] double partial sums | (planes) [blocks][128][2] float maxima | [128] float bounds
    float* maxbuf = d_dx_planes ? (float*)((char*)d_workspace + (size_t)blocks * C * 2 * sizeof(double)) : nullptr;
    float* bound = d_dx_planes ? maxbuf + (size_t)blocks * 128 * 2 : nullptr;
    ColArgs a;
    a.x = d_x; a.dy = d_dy; a.mean = d_mean; a.invstd = d_invstd; a.gamma = d_gamma; a.beta = d_beta;
    a.M = M; a.C = C; a.relu = relu; a.mode = 1; a.maxout = maxbuf;
    col_partial_kernel<<<blocks, 256, 0, s>>>(a, (double*)d_workspace);
    IC_CHECK_LAUNCH();
    col_finalize_kernel<<<C, 128, 0, s>>>((const double*)d_workspace, blocks, C, M, 1, 0.f, d_dbeta, d_dgamma, nullptr,
                                                     nullptr, 0.f, nullptr, maxbuf, d_gamma, d_invstd, bound);
    IC_CHECK_LAUNCH();
    if (d_dx_planes) {
        IC_REQUIRE((((uintptr_t)d_x | (uintptr_t)d_dy | (uintptr_t)d_dx_planes) & 15) == 0, IC_ERR_INVALID, "ic_nn_bn_train_bwd_ex: unaligned tensor");
        static const bool tt = !(getenv("IC_TRAIN_TT") && atoi(getenv("IC_TRAIN_TT")) == 0);
        if (tt)
            bn_bwd_apply_planes_tt_kernel<<<(unsigned)std::min<int64_t>((M + TT_PIX - 1) / TT_PIX, 148 * 2), 256, 0, s>>>((const float4*)d_x, (const float4*)d_dy, d_mean, d_invstd,
                                                                                          d_gamma, d_beta, d_dbeta, d_dgamma, bound, relu, M, hw,
                                                                                          1.f / (float)M, (__half*)d_dx_planes, M * 128, d_scale_out);
        else
            bn_bwd_apply8_planes_kernel<<<ew_grid(M * 16), 256, 0, s>>>((const float4*)d_x, (const float4*)d_dy, d_mean, d_invstd, d_gamma, d_beta,
                                                                        d_dbeta, d_dgamma, bound, relu, M * 16, hw, 1.f / (float)M,
                                                                        (__half*)d_dx_planes, M * 128, d_scale_out);
        IC_CHECK_LAUNCH();
        return IC_OK;
    }
    const bool vec4 = C % 4 == 0 && (((uintptr_t)d_x | (uintptr_t)d_dy | (uintptr_t)d_dx) & 15) == 0;
    if (vec4)
        bn_bwd_apply4_kernel<<<ew_grid(M * C / 4), 256, 0, s>>>((const float4*)d_x, (const float4*)d_dy, d_mean, d_invstd, d_gamma, d_beta,
                                                                d_dbeta, d_dgamma, relu, use_stats ? 0 : 1, M * C / 4, C / 4,
                                                                1.f / (float)M, (float4*)d_dx);
    else
        bn_bwd_apply_kernel<<<ew_grid(M * C), 256, 0, s>>>(d_x, d_dy, d_mean, d_invstd, d_gamma, d_beta, d_dbeta, d_dgamma, relu,
                                                           use_stats ? 0 : 1, M * C, C, 1.f / (float)M, d_dx);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

size_t ic_nn_hq_workspace_bytes(int64_t npix) { return (size_t)((npix + HQ_THREADS - 1) / HQ_THREADS) * 8 * sizeof(double) + 256; }

/* backward of heatmap + mask + soft quantizer (code/autoencoder.py:127-134,171-200; quantizer.py:60-100).
 * d_bn N,h,w,Cb (channel 0 heatmap logit, 1..C features); d_dq N,h,w,C gradient w.r.t. qbar;
 * d_dhm N,C,h,w optional gradient w.r.t. heatmap3D -> d_dbn N,h,w,Cb and d_dcenters (L). */
int ic_nn_hq_bwd(const float* d_bn, int N, int h, int w, int C, int Cb, int heatmap, const float* d_centers, int L,
                 const float* d_dq, const float* d_dhm, float* d_dbn, float* d_dcenters, void* d_workspace,
                 size_t workspace_bytes, void* stream) {
    IC_REQUIRE(d_bn && d_centers && d_dq && d_dbn && d_dcenters && d_workspace, IC_ERR_INVALID, "ic_nn_hq_bwd: NULL argument");
    IC_REQUIRE(L >= 1 && L <= 8 && Cb >= C + (heatmap ? 1 : 0), IC_ERR_INVALID, "ic_nn_hq_bwd: bad shape");
    const int64_t npix = (int64_t)N * h * w;
    IC_REQUIRE(workspace_bytes >= ic_nn_hq_workspace_bytes(npix), IC_ERR_WORKSPACE, "ic_nn_hq_bwd: workspace");
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = (int)((npix + HQ_THREADS - 1) / HQ_THREADS);
    hq_bwd_kernel<<<blocks, HQ_THREADS, 0, s>>>(d_bn, Cb, C, heatmap, d_centers, L, d_dq, d_dhm, h, w, npix, d_dbn,
                                                (double*)d_workspace);
    IC_CHECK_LAUNCH();
    sum_double_partials_kernel<<<1, 32, 0, s>>>((const double*)d_workspace, blocks, 8, L, d_dcenters);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int ic_nn_denorm_clip_fwd(const float* d_v_nhwc4, int N, int H, int W, float* d_x_out_nchw, void* stream) {
    IC_REQUIRE(d_v_nhwc4 && d_x_out_nchw, IC_ERR_INVALID, "ic_nn_denorm_clip_fwd: NULL argument");
    const int64_t hw = (int64_t)H * W;
    denorm_clip_fwd_kernel<<<ew_grid(N * hw), 256, 0, (cudaStream_t)stream>>>(d_v_nhwc4, hw, N * hw, d_x_out_nchw);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int ic_nn_denorm_clip_bwd(const float* d_v_nhwc4, const float* d_dx_out_nchw, int N, int H, int W, float* d_dv_nhwc4, void* stream) {
    IC_REQUIRE(d_v_nhwc4 && d_dx_out_nchw && d_dv_nhwc4, IC_ERR_INVALID, "ic_nn_denorm_clip_bwd: NULL argument");
    const int64_t hw = (int64_t)H * W;
    denorm_clip_bwd_kernel<<<ew_grid(N * hw), 256, 0, (cudaStream_t)stream>>>(d_v_nhwc4, d_dx_out_nchw, hw, N * hw, d_dv_nhwc4);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int ic_nn_nhwc_to_nchw(const float* d_in, int N, int C, int Cs, int64_t hw, float* d_out, void* stream) {
    IC_REQUIRE(d_in && d_out && Cs >= C, IC_ERR_INVALID, "ic_nn_nhwc_to_nchw: bad argument");
    nhwc_to_nchw_kernel<<<ew_grid(N * C * hw), 256, 0, (cudaStream_t)stream>>>(d_in, C, Cs, hw, N * C * hw, d_out);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int ic_nn_nchw_to_nhwc(const float* d_in, int N, int C, int Cs, int64_t hw, float* d_out, void* stream) {
    IC_REQUIRE(d_in && d_out && Cs >= C, IC_ERR_INVALID, "ic_nn_nchw_to_nhwc: bad argument");
    nchw_to_nhwc_pad_kernel<<<ew_grid(N * Cs * hw), 256, 0, (cudaStream_t)stream>>>(d_in, C, Cs, hw, N * Cs * hw, d_out);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

/* out = a * x + b * y (y optional); out may alias x or y */
int ic_nn_axpby(float a, const float* d_x, float b, const float* d_y, int64_t n, float* d_out, void* stream) {
    IC_REQUIRE(d_x && d_out && n >= 0, IC_ERR_INVALID, "ic_nn_axpby: bad argument");
    if (n == 0) return IC_OK;
    axpby_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(a, d_x, b, d_y, n, d_out);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int ic_nn_mul(const float* d_x, const float* d_y, int64_t n, float* d_out, void* stream) {
    IC_REQUIRE(d_x && d_y && d_out && n >= 0, IC_ERR_INVALID, "ic_nn_mul: bad argument");
    if (n == 0) return IC_OK;
    mul_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_x, d_y, n, d_out);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

/* one tf.train.AdamOptimizer step on a flat tensor: g <- grad + l2 * w (slim l2 regulariser), optional 0/1 mask,
 * lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), step t counts from 1. */
int ic_nn_adam_step(float* d_w, const float* d_grad, float* d_m, float* d_v, int64_t n, float lr, float beta1, float beta2,
                    float eps, int64_t step, float l2, const float* d_mask, void* stream) {
    IC_REQUIRE(d_w && d_grad && d_m && d_v && n >= 0 && step >= 1, IC_ERR_INVALID, "ic_nn_adam_step: bad argument");
    if (n == 0) return IC_OK;
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
    adam_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_w, d_grad, d_m, d_v, n, (float)lr_t, beta1, beta2, eps, l2, d_mask,
                                                              nullptr);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

/* the same update with the bias-corrected step size lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) read from device
 * memory (d_lr_t[0]): the launch arguments do not change from step to step, so the step can live in a CUDA graph */
int ic_nn_adam_step_dev(float* d_w, const float* d_grad, float* d_m, float* d_v, int64_t n, const float* d_lr_t, float beta1,
                        float beta2, float eps, float l2, const float* d_mask, void* stream) {
    IC_REQUIRE(d_w && d_grad && d_m && d_v && d_lr_t && n >= 0, IC_ERR_INVALID, "ic_nn_adam_step_dev: bad argument");
    if (n == 0) return IC_OK;
    adam_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_w, d_grad, d_m, d_v, n, 0.f, beta1, beta2, eps, l2, d_mask, d_lr_t);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

/* out = d_a[0] * x (scalar in device memory) */
int ic_nn_scale_dev(const float* d_a, const float* d_x, int64_t n, float* d_out, void* stream) {
    IC_REQUIRE(d_a && d_x && d_out && n >= 0, IC_ERR_INVALID, "ic_nn_scale_dev: bad argument");
    if (n == 0) return IC_OK;
    scale_dev_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d_a, d_x, n, d_out);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

/* d pc_loss / d bitcost coefficient of code/train.py:309-316 from the sums of ic_masked_sums_fwd, on the device:
 * d_coef[0] = beta * 0.5 / n if 0.5 * (mean(bc * heatmap) + mean(bc)) > H_target else 0 */
int ic_nn_rate_coef(const double* d_sums, int64_t n, float beta, float h_target, int has_heatmap, float* d_coef, void* stream) {
    IC_REQUIRE(d_sums && d_coef && n > 0, IC_ERR_INVALID, "ic_nn_rate_coef: bad argument");
    rate_coef_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_sums, (double)n, beta, h_target, has_heatmap, d_coef);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

/* _normalize (code/autoencoder.py:136-144): x NCHW (uint8 or float32, [0,255]) -> normalised NHWC, 4 channels (last 0) */
int ic_nn_normalize_fwd(const void* d_x_nchw, int x_is_u8, int N, int H, int W, float* d_out_nhwc4, void* stream) {
    IC_REQUIRE(d_x_nchw && d_out_nhwc4 && N > 0 && H > 0 && W > 0, IC_ERR_INVALID, "ic_nn_normalize_fwd: bad argument");
    return launch_prep_input(d_x_nchw, x_is_u8, N, H, W, 1, d_out_nhwc4, (cudaStream_t)stream);
}

/* forward twin of ic_nn_hq_bwd: d_bn N,h,w,Cb -> NCHW z, heatmap3D, qbar, qhard, qsoft (each optional) and int64 symbols */
int ic_nn_hq_fwd(const float* d_bn, int N, int h, int w, int C, int Cb, int heatmap, const float* d_centers, int L, float* d_z,
                 float* d_heatmap, float* d_qbar, float* d_qhard, float* d_qsoft, int64_t* d_symbols, void* stream) {
    IC_REQUIRE(d_bn && d_centers && L >= 1 && L <= 8 && Cb >= C + (heatmap ? 1 : 0), IC_ERR_INVALID, "ic_nn_hq_fwd: bad argument");
    return launch_heatmap_quantize(d_bn, N, h, w, C, heatmap, d_centers, L, d_z, d_heatmap, d_qbar, d_qhard, d_symbols, nullptr,
                                   d_qsoft, (cudaStream_t)stream, Cb);
}

int ic_nn_pc_pad_fwd(const float* d_q_nchw, int N, int C, int h, int w, float pad_value, const float* d_pad_value, float* d_out,
                     void* stream) {
    IC_REQUIRE(d_q_nchw && d_out && N > 0 && C > 0 && h > 0 && w > 0, IC_ERR_INVALID, "ic_nn_pc_pad_fwd: bad argument");
    const int64_t total = (int64_t)(C + 4) * N * (h + 8) * (w + 8);
    pc_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_q_nchw, N, C, h, w, pad_value, d_pad_value, total,
                                                                                      (float4*)d_out);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int ic_nn_pc_xent_fwd(const float* d_logits, int Cs, int L, const int64_t* d_symbols, int N, int C, int h, int w, float* d_bc_nchw,
                      void* stream) {
    IC_REQUIRE(d_logits && d_symbols && d_bc_nchw && L >= 1 && Cs >= L, IC_ERR_INVALID, "ic_nn_pc_xent_fwd: bad argument");
    const int64_t total = (int64_t)N * C * h * w;
    pc_xent_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_logits, Cs, L, d_symbols, nullptr, N, C, h, w, 0.f, 0.f,
                                                                            nullptr, total, d_bc_nchw);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int ic_nn_pc_xent_bwd(const float* d_logits, int Cs, int L, const int64_t* d_symbols, const float* d_heatmap, int N, int C, int h,
                      int w, float coef_real, float coef_mask, const float* d_coef, float* d_dlogits, void* stream) {
    IC_REQUIRE(d_logits && d_symbols && d_dlogits && L >= 1 && Cs >= L, IC_ERR_INVALID, "ic_nn_pc_xent_bwd: bad argument");
    const int64_t total = (int64_t)N * C * h * w;
    pc_xent_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_logits, Cs, L, d_symbols, d_heatmap, N, C, h, w, coef_real,
                                                                           coef_mask, d_coef, total, d_dlogits);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int ic_nn_crop_fwd(const float* d_in, int64_t A, int H, int W, int C, int crop, float* d_out, void* stream) {
    IC_REQUIRE(d_in && d_out && C % 4 == 0 && crop >= 0 && H > 2 * crop && W > 2 * crop, IC_ERR_INVALID, "ic_nn_crop_fwd: bad argument");
    const int64_t total = A * (H - 2 * crop) * (W - 2 * crop) * (C / 4);
    crop_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_in, H, W, C / 4, crop, total, (float4*)d_out);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int ic_nn_crop_bwd_add(const float* d_dy, int64_t A, int H, int W, int C, int crop, float* d_dx, void* stream) {
    IC_REQUIRE(d_dy && d_dx && C % 4 == 0 && crop >= 0 && H > 2 * crop && W > 2 * crop, IC_ERR_INVALID, "ic_nn_crop_bwd_add: bad argument");
    const int64_t total = A * (H - 2 * crop) * (W - 2 * crop) * (C / 4);
    crop_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_dy, H, W, C / 4, crop, total, (float4*)d_dx);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

size_t ic_msssim_bwd_workspace_bytes(int N, int H, int W) { return msssim_bwd_workspace_bytes(N, H, W); }

int ic_msssim_tf_bwd(const float* d_img1, const float* d_img2, int N, int H, int W, float grad_out, float* d_dimg2, float* d_value,
                     void* d_workspace, size_t workspace_bytes, void* stream) {
    IC_REQUIRE(d_img1 && d_img2 && d_dimg2 && d_workspace, IC_ERR_INVALID, "ic_msssim_tf_bwd: NULL argument");
    IC_REQUIRE(N > 0 && H > 0 && W > 0, IC_ERR_INVALID, "ic_msssim_tf_bwd: bad shape");
    return msssim_tf_bwd(d_img1, d_img2, N, H, W, grad_out, d_dimg2, d_value, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
