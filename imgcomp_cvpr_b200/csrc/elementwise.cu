// HBM-bound elementwise stages of the encoder: input normalisation / layout
// change, importance-map (heatmap) masking and the nearest-centre quantizer.
#include <cuda_fp16.h>

#include "common.cuh"

namespace ic {

namespace {

__constant__ float c_mean[3] = {121.853699f, 113.588608f, 100.637154f};
__constant__ float c_std[3] = {68.8939514f, 66.7393417f, 69.3702698f};   // float32 sqrt(var + 1e-10)

// _Network._normalize (code/autoencoder.py:136-144) fused with tf.to_float
// (val.py:83) and the NCHW -> NHWC(4) layout change.  One thread per pixel.
template <typename TIn>
__global__ void prep_input_kernel(const TIn* __restrict__ x, int64_t npix_per_img, int64_t total, int normalize,
                                  float4* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t n = i / npix_per_img, r = i - n * npix_per_img;
    const TIn* p = x + n * 3 * npix_per_img + r;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float f = (float)p[c * npix_per_img];
        v[c] = normalize ? __fdiv_rn(__fsub_rn(f, c_mean[c]), c_std[c]) : f;
    }
    out[i] = make_float4(v[0], v[1], v[2], 0.f);
}

// Same normalisation, but written as the input of the tensor-core h1: fp16 hi/lo planes in
// space-to-depth form [plane][N][4 phases][H/2][W/2][8] (channels 0..2 = RGB, 3..7 = 0), so that the
// 5x5 stride-2 conv becomes 9 stride-1 taps over ONE 32-channel group.
template <typename TIn>
__global__ void prep_input_s2d_kernel(const TIn* __restrict__ x, int H, int W, int64_t total, int normalize,
                                      __half* __restrict__ out, int write_lo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per input pixel
    if (i >= total) return;
    const int64_t hw = (int64_t)H * W;
    const int64_t n = i / hw, r = i - n * hw;
    const int y = (int)(r / W), xx = (int)(r - (int64_t)y * W);
    const TIn* p = x + n * 3 * hw + r;
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float v = 0.f;
        if (c < 3) {
            float f = (float)p[c * hw];
            v = normalize ? __fdiv_rn(__fsub_rn(f, c_mean[c]), c_std[c]) : f;
        }
        hi[c] = __float2half_rn(v);
        lo[c] = __float2half_rn(v - __half2float(hi[c]));
    }
    const int ph = (y & 1) * 2 + (xx & 1);
    const size_t off = ((((size_t)n * 4 + ph) * (H / 2) + (y >> 1)) * (W / 2) + (xx >> 1)) * 8;
    const size_t plane = (size_t)(total / hw) * hw * 8;              // N * 4 * (H/2) * (W/2) * 8
    *reinterpret_cast<float4*>(out + off) = *reinterpret_cast<const float4*>(hi);
    if (write_lo) *reinterpret_cast<float4*>(out + plane + off) = *reinterpret_cast<const float4*>(lo);
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, int C, int64_t hw, int64_t total,
                                    float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // index in NHWC order
    if (i >= total) return;
    int c = (int)(i % C);
    int64_t p = i / C;
    int64_t n = p / hw, r = p - n * hw;
    out[i] = in[(n * C + c) * hw + r];
}

// quantizer._quantize1d for one value (code/quantizer.py:72-95).
//   dist_j  = square(abs(x - c_j))                       two roundings, no FMA
//   symbol  = argmax_j softmax(-1e7 * dist)_j            first index among maxima; the
//             softmax is monotone in fl(-1e7*dist_j) and exp(l_j - max) < 1 whenever
//             l_j < max, so this is the first j minimising fl(1e7 * dist_j)
//   qhard   = sum_j onehot_j * c_j = c[symbol] exactly
//   qsoft   = sum_j softmax(-sigma*dist)_j * c_j         summed in index order
template <int MAXL>
__device__ __forceinline__ void quantize_one(float x, const float* __restrict__ c, int L, float sigma,
                                             float& qsoft, float& qhard, int& sym) {
    float dist[MAXL];
    float best = 0.f;
    int bi = 0;
    float lmax = 0.f;
#pragma unroll
    for (int j = 0; j < MAXL; ++j) {
        if (j < L) {
            float df = fabsf(__fsub_rn(x, c[j]));
            dist[j] = __fmul_rn(df, df);
            float lh = __fmul_rn(1e7f, dist[j]);
            float ls = __fmul_rn(-sigma, dist[j]);
            if (j == 0 || lh < best) { best = lh; bi = j; }
            if (j == 0 || ls > lmax) lmax = ls;
        }
    }
    float e[MAXL];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXL; ++j)
        if (j < L) {
            e[j] = expf(__fsub_rn(__fmul_rn(-sigma, dist[j]), lmax));
            s = __fadd_rn(s, e[j]);
        }
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < MAXL; ++j)
        if (j < L) {
            float term = __fmul_rn(__fdiv_rn(e[j], s), c[j]);
            acc = (j == 0) ? term : __fadd_rn(acc, term);
        }
    qsoft = acc;
    qhard = c[bi];
    sym = bi;
}

constexpr int kMaxL = 8;

// _get_heatmap3D + _mask_with_heatmap + _quantize (code/autoencoder.py:127-134,171-200)
// in: to_bn output NHWC (N,h,w,C+1); out: NCHW tensors.  One thread per (n,c,y,x).
__global__ void heatmap_quantize_kernel(const float* __restrict__ bn, int h, int w, int C, int heatmap,
                                        const float* __restrict__ centers, int L, int64_t total,
                                        float* __restrict__ z_out, float* __restrict__ hm_out,
                                        float* __restrict__ qbar_out, float* __restrict__ qhard_out,
                                        int64_t* __restrict__ sym_out, uint8_t* __restrict__ sym8_out,
                                        float* __restrict__ qsoft_out, int cb_stride) {
    __shared__ float sc[kMaxL];
    if (threadIdx.x < L) sc[threadIdx.x] = centers[threadIdx.x];
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // NCHW index
    if (i >= total) return;
    int64_t hw = (int64_t)h * w;
    int64_t r = i % hw;
    int64_t nc = i / hw;
    int c = (int)(nc % C);
    int64_t n = nc / C;
    int CB = cb_stride > 0 ? cb_stride : (heatmap ? C + 1 : C);
    const float* px = bn + (n * hw + r) * CB;
    float z, hm = 1.f;
    if (heatmap) {
        float h0 = px[0];
        // tf.nn.sigmoid(x) * C ; heatmap3D = max(min(hm2D - c, 1), 0)
        float sg = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-h0)));
        float hm2d = __fmul_rn(sg, (float)C);
        hm = fmaxf(fminf(__fsub_rn(hm2d, (float)c), 1.f), 0.f);
        z = __fmul_rn(hm, px[1 + c]);
    } else {
        z = px[c];
    }
    float qsoft, qhard;
    int sym;
    quantize_one<kMaxL>(z, sc, L, 1.f, qsoft, qhard, sym);
    if (z_out) z_out[i] = z;
    if (hm_out) hm_out[i] = hm;
    if (qsoft_out) qsoft_out[i] = qsoft;
    if (qhard_out) qhard_out[i] = qhard;
    // qbar = qsoft + stop_gradient(qhard - qsoft)   (code/autoencoder.py:133)
    if (qbar_out) qbar_out[i] = __fadd_rn(qsoft, __fsub_rn(qhard, qsoft));
    if (sym_out) sym_out[i] = sym;
    if (sym8_out) sym8_out[i] = (uint8_t)sym;
}

// The same, tiled for the memory system: a block stages the to_bn rows of 64 consecutive latent pixels (64 x CB floats,
// contiguous in the NHWC input: coalesced 16-byte loads) in shared memory, then every thread produces FOUR consecutive
// pixels of one channel, so each NCHW output row is written as 16-byte (32-byte for the int64 symbols) stores, 256
// contiguous bytes per channel and tile.  29 B written per symbol against ~2 B read: the kernel is write-bound.
// Requires h*w % 4 == 0 (a quad may not straddle two images); arithmetic identical to heatmap_quantize_kernel.
constexpr int kHqTile = 64;

__global__ void __launch_bounds__(256) heatmap_quantize_tiled_kernel(
    const float* __restrict__ bn, int64_t hw, int C, int heatmap, const float* __restrict__ centers, int L, int64_t npix,
    float* __restrict__ z_out, float* __restrict__ hm_out, float* __restrict__ qbar_out, float* __restrict__ qhard_out,
    int64_t* __restrict__ sym_out, uint8_t* __restrict__ sym8_out, float* __restrict__ qsoft_out, int CB) {
    extern __shared__ __align__(16) float tile[];          // [kHqTile][CB + 1] (odd pitch: conflict-free column reads)
    __shared__ float sc[kMaxL];
    __shared__ float s_hm2d[kHqTile];
    if (threadIdx.x < L) sc[threadIdx.x] = centers[threadIdx.x];
    const int64_t p0 = (int64_t)blockIdx.x * kHqTile;
    const int np = (int)min((int64_t)kHqTile, npix - p0);
    const int pitch = CB | 1;
    const float* src = bn + p0 * CB;
    for (int i = threadIdx.x; i < np * CB; i += blockDim.x) tile[(i / CB) * pitch + i % CB] = src[i];
    __syncthreads();
    if (heatmap && threadIdx.x < np) {
        // tf.nn.sigmoid(x) * C   (code/autoencoder.py:183-186)
        const float sg = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-tile[threadIdx.x * pitch])));
        s_hm2d[threadIdx.x] = __fmul_rn(sg, (float)C);
    }
    __syncthreads();
    const int quads = kHqTile / 4;
    for (int item = threadIdx.x; item < C * quads; item += blockDim.x) {
        const int c = item / quads, qd = item - c * quads;
        const int px = qd * 4;
        if (px >= np) continue;
        const int64_t p = p0 + px;                  // first pixel of the quad; hw % 4 == 0: all four in one image
        const int64_t n = p / hw, r = p - n * hw;
        const int64_t o = (n * C + c) * hw + r;
        float z[4], hm[4], qs[4], qh[4], qb[4];
        int sy[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float* px_row = tile + (px + k) * pitch;
            hm[k] = 1.f;
            if (heatmap) {
                hm[k] = fmaxf(fminf(__fsub_rn(s_hm2d[px + k], (float)c), 1.f), 0.f);    // heatmap3D = max(min(hm2D - c, 1), 0)
                z[k] = __fmul_rn(hm[k], px_row[1 + c]);
            } else {
                z[k] = px_row[c];
            }
            quantize_one<kMaxL>(z[k], sc, L, 1.f, qs[k], qh[k], sy[k]);
            qb[k] = __fadd_rn(qs[k], __fsub_rn(qh[k], qs[k]));            // qbar (code/autoencoder.py:133)
        }
        if (z_out) *reinterpret_cast<float4*>(z_out + o) = make_float4(z[0], z[1], z[2], z[3]);
        if (hm_out) *reinterpret_cast<float4*>(hm_out + o) = make_float4(hm[0], hm[1], hm[2], hm[3]);
        if (qsoft_out) *reinterpret_cast<float4*>(qsoft_out + o) = make_float4(qs[0], qs[1], qs[2], qs[3]);
        if (qhard_out) *reinterpret_cast<float4*>(qhard_out + o) = make_float4(qh[0], qh[1], qh[2], qh[3]);
        if (qbar_out) *reinterpret_cast<float4*>(qbar_out + o) = make_float4(qb[0], qb[1], qb[2], qb[3]);
        if (sym_out) {
            *reinterpret_cast<longlong2*>(sym_out + o) = make_longlong2(sy[0], sy[1]);
            *reinterpret_cast<longlong2*>(sym_out + o + 2) = make_longlong2(sy[2], sy[3]);
        }
        if (sym8_out) *reinterpret_cast<uchar4*>(sym8_out + o) = make_uchar4((uint8_t)sy[0], (uint8_t)sy[1], (uint8_t)sy[2], (uint8_t)sy[3]);
    }
}

__global__ void quantize_kernel(const float* __restrict__ x, const float* __restrict__ centers, int L, float sigma,
                                int64_t n, float* __restrict__ qsoft_out, float* __restrict__ qhard_out,
                                int64_t* __restrict__ sym_out) {
    __shared__ float sc[kMaxL];
    if (threadIdx.x < L) sc[threadIdx.x] = centers[threadIdx.x];
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float qsoft, qhard;
    int sym;
    quantize_one<kMaxL>(x[i], sc, L, sigma, qsoft, qhard, sym);
    if (qsoft_out) qsoft_out[i] = qsoft;
    if (qhard_out) qhard_out[i] = qhard;
    if (sym_out) sym_out[i] = sym;
}

}  // namespace

int launch_prep_input(const void* x, int is_u8, int N, int H, int W, int normalize, float* out_nhwc4,
                      cudaStream_t s) {
    int64_t hw = (int64_t)H * W, total = hw * N;
    int nb = cdiv(total, 256);
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    if (is_u8)
        prep_input_kernel<uint8_t><<<nb, 256, 0, s>>>((const uint8_t*)x, hw, total, normalize, (float4*)out_nhwc4);
    else
        prep_input_kernel<float><<<nb, 256, 0, s>>>((const float*)x, hw, total, normalize, (float4*)out_nhwc4);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int launch_prep_input_s2d(const void* x, int is_u8, int N, int H, int W, int normalize, __half* out, int write_lo,
                          cudaStream_t s) {
    int64_t total = (int64_t)N * H * W;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    if (is_u8)
        prep_input_s2d_kernel<uint8_t><<<cdiv(total, 256), 256, 0, s>>>((const uint8_t*)x, H, W, total, normalize, out, write_lo);
    else
        prep_input_s2d_kernel<float><<<cdiv(total, 256), 256, 0, s>>>((const float*)x, H, W, total, normalize, out, write_lo);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int launch_nchw_to_nhwc(const float* in, int N, int C, int H, int W, float* out, cudaStream_t s) {
    int64_t hw = (int64_t)H * W, total = hw * N * C;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    nchw_to_nhwc_kernel<<<cdiv(total, 256), 256, 0, s>>>(in, C, hw, total, out);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int launch_heatmap_quantize(const float* bn_nhwc, int N, int h, int w, int C, int heatmap, const float* centers,
                            int L, float* z, float* hm, float* qbar, float* qhard, int64_t* sym, uint8_t* sym8,
                            float* qsoft, cudaStream_t s, int cb_stride) {
    IC_REQUIRE(L <= kMaxL, IC_ERR_UNSUPPORTED, "num_centers %d > %d", L, kMaxL);
    int64_t total = (int64_t)N * C * h * w;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    const int CB = cb_stride > 0 ? cb_stride : (heatmap ? C + 1 : C);
    const int64_t hw = (int64_t)h * w, npix = (int64_t)N * hw;
    const size_t tile_bytes = (size_t)kHqTile * (CB | 1) * sizeof(float);
    auto al = [](const void* q, size_t a) { return q == nullptr || ((uintptr_t)q & (a - 1)) == 0; };
    if (hw % 4 == 0 && tile_bytes <= 40 * 1024 && al(z, 16) && al(hm, 16) && al(qbar, 16) && al(qhard, 16) && al(qsoft, 16) &&
        al(sym, 16) && al(sym8, 4)) {
        heatmap_quantize_tiled_kernel<<<cdiv(npix, kHqTile), 256, tile_bytes, s>>>(bn_nhwc, hw, C, heatmap, centers, L, npix, z, hm, qbar,
                                                                                  qhard, sym, sym8, qsoft, CB);
        IC_CHECK_LAUNCH();
        return IC_OK;
    }
    heatmap_quantize_kernel<<<cdiv(total, 256), 256, 0, s>>>(bn_nhwc, h, w, C, heatmap, centers, L, total, z, hm,
                                                             qbar, qhard, sym, sym8, qsoft, cb_stride);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int launch_quantize(const float* x, const float* centers, int L, float sigma, int64_t n, float* qsoft,
                    float* qhard, int64_t* sym, cudaStream_t s) {
    IC_REQUIRE(L <= kMaxL, IC_ERR_UNSUPPORTED, "num_centers %d > %d", L, kMaxL);
    if (n == 0) return IC_OK;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    quantize_kernel<<<cdiv(n, 256), 256, 0, s>>>(x, centers, L, sigma, n, qsoft, qhard, sym);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

}  // namespace ic
