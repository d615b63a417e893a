// Masked causal 3-D-conv context model ("res_shallow", code/probclass.py:199-221),
// float32 FFMA version: four VALID conv3d layers over the latent volume
// (depth = latent channel), each restricted to the non-masked filter taps
// (create_first_mask / create_other_mask, code/probclass.py:150-176).
//
// The padding of pad_for_probclass3d (code/probclass.py:268-292: depth front 4,
// H/W 4 each side, constant) is applied on the fly by the first layer, so the
// padded volume is never materialised.
//
// Reduction order of every output value is fixed (taps in (fd,fy,fx) raster
// order, input channels ascending, bias last) and does not depend on the
// position in the volume: the batched pass over a whole latent and the
// per-context evaluation of code/probclass.py:441-476 give bit-identical logits.
#include <cuda_fp16.h>
#include <string.h>

#include "common.cuh"
#include "probclass.cuh"

namespace ic {

namespace {

constexpr float kLog2e = 1.44269502f;   // float32(np.log2(np.e)), code/probclass.py:101

// ---------------------------------------------------------------- layer 0
// source of the (virtually padded) input volume
struct SrcFloat {
    const float* q;            // N,D,H,W
    float pad_value;
    __device__ float at(int64_t idx) const { return q[idx]; }
};
struct SrcSymbols {
    const int64_t* sym;        // N,D,H,W
    float centers[8];
    float pad_value;           // = centers[0]: symbol volume is padded with symbol 0 (probclass.py:449-451)
    __device__ float at(int64_t idx) const { return centers[(int)sym[idx]]; }
};

// SPLIT = false: float32 N,D0,H0,W0,KC.  SPLIT = true: fp16 hi/lo planes [plane][N*D0][4][H0][W0][8]
// (channels padded to 32), the input layout of the tensor-core layers.
template <int KC, typename Src, bool SPLIT>
__global__ void __launch_bounds__(128) pc_conv0_kernel(Src src, int D, int H, int W, int pd, int ph,
                                                       const float* __restrict__ wgt,   // [13][KC]
                                                       const float* __restrict__ bias,  // [KC]
                                                       int D0, int H0, int W0, int64_t total,
                                                       float* __restrict__ out, __half* __restrict__ out_split) {
    __shared__ __align__(16) float sw[13 * KC + KC];
    for (int i = threadIdx.x; i < 13 * KC; i += blockDim.x) sw[i] = wgt[i];
    for (int i = threadIdx.x; i < KC; i += blockDim.x) sw[13 * KC + i] = bias[i];
    __syncthreads();
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= total) return;
    int x = (int)(v % W0);
    int64_t r = v / W0;
    int y = (int)(r % H0);
    r /= H0;
    int d = (int)(r % D0);
    int64_t n = r / D0;
    float in[13];
    int t = 0;
#pragma unroll
    for (int fd = 0; fd < 2; ++fd)
#pragma unroll
        for (int fy = 0; fy < 3; ++fy)
#pragma unroll
            for (int fx = 0; fx < 3; ++fx) {
                if (fd == 1 && (fy > 1 || (fy == 1 && fx >= 1))) continue;    // first mask
                int zd = d + fd - pd, zy = y + fy - ph, zx = x + fx - ph;     // un-padded coords
                float val = src.pad_value;
                if (zd >= 0 && zy >= 0 && zy < H && zx >= 0 && zx < W)        // no back padding in depth
                    val = src.at(((n * D + zd) * H + zy) * (int64_t)W + zx);
                in[t++] = val;
            }
    // Weights are read as float4 (4 output channels per LDS.128): with one LDS.32 per FMA the kernel was bound by the
    // shared-memory pipe (312 broadcast loads per voxel), not by its 128 B / voxel store stream.  Every output still
    // sums its 13 taps in raster order with fmaf, then adds the bias: bit-identical to the scalar form.
    static_assert(KC % 8 == 0, "8 output channels per step");
    if (!SPLIT) {
        float* o = out + v * KC;
#pragma unroll 1
        for (int c = 0; c < KC / 8; ++c) {
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 13; ++k) {
                const float4 w0 = *reinterpret_cast<const float4*>(sw + k * KC + c * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(sw + k * KC + c * 8 + 4);
                acc[0] = fmaf(in[k], w0.x, acc[0]); acc[1] = fmaf(in[k], w0.y, acc[1]);
                acc[2] = fmaf(in[k], w0.z, acc[2]); acc[3] = fmaf(in[k], w0.w, acc[3]);
                acc[4] = fmaf(in[k], w1.x, acc[4]); acc[5] = fmaf(in[k], w1.y, acc[5]);
                acc[6] = fmaf(in[k], w1.z, acc[6]); acc[7] = fmaf(in[k], w1.w, acc[7]);
            }
            float r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = fmaxf(acc[j] + sw[13 * KC + c * 8 + j], 0.f);
            *reinterpret_cast<float4*>(o + c * 8) = make_float4(r[0], r[1], r[2], r[3]);
            *reinterpret_cast<float4*>(o + c * 8 + 4) = make_float4(r[4], r[5], r[6], r[7]);
        }
    } else {
        const size_t cs = (size_t)H0 * W0 * 8;                         // chunk stride (elements)
        const size_t plane = (size_t)(total / ((int64_t)H0 * W0)) * 4 * cs;
        const size_t base = (((size_t)(n * D0 + d) * 4) * H0 + y) * W0 * 8 + (size_t)x * 8;
#pragma unroll 1
        for (int c = 0; c < (KC + 7) / 8; ++c) {          // the padding chunk (channels 24..31) is never read: not written
            float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (c * 8 < KC) {
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 13; ++k) {
                    const float4 w0 = *reinterpret_cast<const float4*>(sw + k * KC + c * 8);
                    const float4 w1 = *reinterpret_cast<const float4*>(sw + k * KC + c * 8 + 4);
                    acc[0] = fmaf(in[k], w0.x, acc[0]); acc[1] = fmaf(in[k], w0.y, acc[1]);
                    acc[2] = fmaf(in[k], w0.z, acc[2]); acc[3] = fmaf(in[k], w0.w, acc[3]);
                    acc[4] = fmaf(in[k], w1.x, acc[4]); acc[5] = fmaf(in[k], w1.y, acc[5]);
                    acc[6] = fmaf(in[k], w1.z, acc[6]); acc[7] = fmaf(in[k], w1.w, acc[7]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = fmaxf(acc[j] + sw[13 * KC + c * 8 + j], 0.f);
            }
            __align__(16) __half2 hi[4];
            __align__(16) __half2 lo[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
                hi[e2] = __floats2half2_rn(r[2 * e2], r[2 * e2 + 1]);
                float2 hf = __half22float2(hi[e2]);
                lo[e2] = __floats2half2_rn(r[2 * e2] - hf.x, r[2 * e2 + 1] - hf.y);
            }
            *reinterpret_cast<float4*>(out_split + base + c * cs) = *reinterpret_cast<const float4*>(hi);
            *reinterpret_cast<float4*>(out_split + plane + base + c * cs) = *reinterpret_cast<const float4*>(lo);
        }
    }
}

// ---------------------------------------------------------------- middle layers
// in N,Di,Hi,Wi,KC -> out N,Di-1,Hi-2,Wi-2,KC ; 14 taps ("other" mask).
template <int KC, bool RELU, bool RESIDUAL>
__global__ void __launch_bounds__(128) pc_conv_mid_kernel(const float* __restrict__ in, int Di, int Hi, int Wi,
                                                          const float* __restrict__ wgt,   // [14][KC][KC]
                                                          const float* __restrict__ bias,  // [KC]
                                                          const float* __restrict__ res,   // N,Dr,Hr,Wr,KC (layer-0 out)
                                                          int Dr, int Hr, int Wr, int64_t total,
                                                          float* __restrict__ out) {
    __shared__ __align__(16) float sw[KC * KC];
    const int Do = Di - 1, Ho = Hi - 2, Wo = Wi - 2;
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = v < total;
    int x = 0, y = 0, d = 0;
    int64_t n = 0;
    if (valid) {
        x = (int)(v % Wo);
        int64_t r = v / Wo;
        y = (int)(r % Ho);
        r /= Ho;
        d = (int)(r % Do);
        n = r / Do;
    }
    float acc[KC];
#pragma unroll
    for (int i = 0; i < KC; ++i) acc[i] = 0.f;
    int tap = 0;
    for (int fd = 0; fd < 2; ++fd)
        for (int fy = 0; fy < 3; ++fy)
            for (int fx = 0; fx < 3; ++fx) {
                if (fd == 1 && (fy > 1 || (fy == 1 && fx > 1))) continue;     // other mask
                __syncthreads();
                for (int i = threadIdx.x; i < KC * KC / 4; i += blockDim.x)
                    reinterpret_cast<float4*>(sw)[i] = reinterpret_cast<const float4*>(wgt + (int64_t)tap * KC * KC)[i];
                __syncthreads();
                ++tap;
                if (!valid) continue;
                const float4* p = reinterpret_cast<const float4*>(
                    in + ((((n * Di + d + fd) * Hi + y + fy) * (int64_t)Wi) + x + fx) * KC);
#pragma unroll 2
                for (int c4 = 0; c4 < KC / 4; ++c4) {
                    float4 a4 = p[c4];
                    float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4* wr = reinterpret_cast<const float4*>(sw + (c4 * 4 + u) * KC);
#pragma unroll
                        for (int o4 = 0; o4 < KC / 4; ++o4) {
                            float4 w4 = wr[o4];
                            acc[o4 * 4 + 0] = fmaf(a[u], w4.x, acc[o4 * 4 + 0]);
                            acc[o4 * 4 + 1] = fmaf(a[u], w4.y, acc[o4 * 4 + 1]);
                            acc[o4 * 4 + 2] = fmaf(a[u], w4.z, acc[o4 * 4 + 2]);
                            acc[o4 * 4 + 3] = fmaf(a[u], w4.w, acc[o4 * 4 + 3]);
                        }
                    }
                }
            }
    if (!valid) return;
    float* o = out + v * KC;
    const float* rp = nullptr;
    if (RESIDUAL)   // x + residual_input[..., 2:, 2:-2, 2:-2, :]   (code/probclass.py:196)
        rp = res + ((((n * Dr + d + 2) * Hr + y + 2) * (int64_t)Wr) + x + 2) * KC;
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        float r = acc[c] + bias[c];
        if (RELU) r = fmaxf(r, 0.f);
        if (RESIDUAL) r += rp[c];
        o[c] = r;
    }
}

// ---------------------------------------------------------------- final layer + head
enum { HEAD_LOGITS = 0, HEAD_BITCOST = 1, HEAD_FREQS = 2 };

template <int KC, int HEAD>
__global__ void __launch_bounds__(128) pc_final_kernel(const float* __restrict__ in, int Di, int Hi, int Wi,
                                                       const float* __restrict__ wgt,   // [14][KC][L]
                                                       const float* __restrict__ bias, int L,
                                                       const int64_t* __restrict__ symbols,   // N,Do,Ho,Wo
                                                       int64_t total, float* __restrict__ out_f,
                                                       int64_t* __restrict__ out_freqs, double* __restrict__ bits_sum,
                                                       uint32_t* __restrict__ out_freqs32 = nullptr) {
    extern __shared__ float swf[];      // 14*KC*L + L
    const int nw = 14 * KC * L;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) swf[i] = wgt[i];
    for (int i = threadIdx.x; i < L; i += blockDim.x) swf[nw + i] = bias[i];
    __syncthreads();
    const int Do = Di - 1, Ho = Hi - 2, Wo = Wi - 2;
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = v < total;
    float bits = 0.f;
    int64_t n = 0;
    if (valid) {
        int x = (int)(v % Wo);
        int64_t r = v / Wo;
        int y = (int)(r % Ho);
        r /= Ho;
        int d = (int)(r % Do);
        n = r / Do;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        int tap = 0;
        for (int fd = 0; fd < 2; ++fd)
            for (int fy = 0; fy < 3; ++fy)
                for (int fx = 0; fx < 3; ++fx) {
                    if (fd == 1 && (fy > 1 || (fy == 1 && fx > 1))) continue;
                    const float4* p = reinterpret_cast<const float4*>(
                        in + ((((n * Di + d + fd) * Hi + y + fy) * (int64_t)Wi) + x + fx) * KC);
                    const float* wt = swf + tap * KC * L;
                    ++tap;
#pragma unroll 2
                    for (int c4 = 0; c4 < KC / 4; ++c4) {
                        float4 a4 = p[c4];
                        float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int l = 0; l < 8; ++l)
                                if (l < L) acc[l] = fmaf(a[u], wt[(c4 * 4 + u) * L + l], acc[l]);
                    }
                }
        // bias + ReLU: the last conv3d keeps the default activation (code/probclass.py:220,233)
        float lg[8];
        float m = 0.f;
#pragma unroll
        for (int l = 0; l < 8; ++l)
            if (l < L) {
                lg[l] = fmaxf(acc[l] + swf[nw + l], 0.f);
                m = (l == 0) ? lg[l] : fmaxf(m, lg[l]);
            }
        if (HEAD == HEAD_LOGITS) {
#pragma unroll
            for (int l = 0; l < 8; ++l)
                if (l < L) out_f[v * L + l] = lg[l];
        } else {
            float e[8], s = 0.f;
#pragma unroll
            for (int l = 0; l < 8; ++l)
                if (l < L) {
                    e[l] = expf(__fsub_rn(lg[l], m));
                    s = __fadd_rn(s, e[l]);
                }
            int sym = symbols ? (int)symbols[v] : 0;
            float lsel = 0.f;
#pragma unroll
            for (int l = 0; l < 8; ++l)
                if (l == sym) lsel = lg[l];
            // softmax_cross_entropy_with_logits * log2(e)   (code/probclass.py:100-104)
            bits = __fmul_rn(__fsub_rn(logf(s), __fsub_rn(lsel, m)), kLog2e);
            if (HEAD == HEAD_BITCOST) {
                out_f[v] = bits;
            } else {
                // PredictionNetwork: int64(softmax * 1e9), max(.,1)   (code/probclass.py:443-444,473)
#pragma unroll
                for (int l = 0; l < 8; ++l)
                    if (l < L) {
                        float pr = __fdiv_rn(e[l], s);
                        long long f = (long long)__fmul_rn(pr, 1e9f);
                        f = f < 1 ? 1 : f;
                        if (out_freqs32) out_freqs32[v * L + l] = (uint32_t)f;        // <= 1e9 < 2^30: fits
                        else out_freqs[v * L + l] = f;
                    }
            }
        }
    }
    if (HEAD != HEAD_LOGITS && bits_sum) {
        // per-image sum of bits, double: warp shuffle then one atomic per warp
        // (the whole warp belongs to one image except at image boundaries -> handle generally)
        double b = valid ? (double)bits : 0.0;
        unsigned mask = 0xffffffffu;
        int64_t n0 = __shfl_sync(mask, n, 0);
        bool uniform = __all_sync(mask, (!valid) || n == n0);
        if (uniform) {
            for (int o = 16; o > 0; o >>= 1) b += __shfl_down_sync(mask, b, o);
            if ((threadIdx.x & 31) == 0 && __shfl_sync(mask, (int)valid, 0)) atomicAdd(bits_sum + n0, b);
        } else if (valid) {
            atomicAdd(bits_sum + n, b);
        }
    }
}

template <int KC>
int run_pc(const PcWeights& w, const PcInput& in, int head, float* out_f, int64_t* out_freqs, double* bits_sum,
           void* ws, size_t ws_bytes, cudaStream_t s, uint32_t* out_freqs32 = nullptr) {
    const int N = in.N, D = in.D, H = in.H, W = in.W, pd = in.pad_d, ph = in.pad_hw, L = w.L;
    const int Dp = D + pd, Hp = H + 2 * ph, Wp = W + 2 * ph;
    IC_REQUIRE(Dp >= 5 && Hp >= 9 && Wp >= 9, IC_ERR_INVALID, "probclass: volume %dx%dx%d smaller than the 5x9x9 context",
               Dp, Hp, Wp);
    const int D0 = Dp - 1, H0 = Hp - 2, W0 = Wp - 2;
    const int D1 = D0 - 1, H1 = H0 - 2, W1 = W0 - 2;
    const int D2 = D1 - 1, H2 = H1 - 2, W2 = W1 - 2;
    const int D3 = D2 - 1, H3 = H2 - 2, W3 = W2 - 2;
    Arena ar(ws, ws_bytes);
    float* a0 = ar.get<float>((size_t)N * D0 * H0 * W0 * KC);
    float* a1 = ar.get<float>((size_t)N * D1 * H1 * W1 * KC);
    float* a2 = ar.get<float>((size_t)N * D2 * H2 * W2 * KC);
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "probclass workspace too small: need %zu, have %zu", ar.off, ws_bytes);

    int64_t t0 = (int64_t)N * D0 * H0 * W0;
    ProfScope ps(IC_PROF_PROBCLASS, s, 4);
    if (in.symbols) {
        SrcSymbols src;
        src.sym = in.symbols;
        for (int i = 0; i < 8; ++i) src.centers[i] = i < L ? in.centers_host[i] : 0.f;
        src.pad_value = in.centers_host[0];
        pc_conv0_kernel<KC, SrcSymbols, false><<<cdiv(t0, 128), 128, 0, s>>>(src, D, H, W, pd, ph, w.w0, w.b0, D0, H0, W0, t0, a0, nullptr);
    } else {
        SrcFloat src{in.q, in.pad_value};
        pc_conv0_kernel<KC, SrcFloat, false><<<cdiv(t0, 128), 128, 0, s>>>(src, D, H, W, pd, ph, w.w0, w.b0, D0, H0, W0, t0, a0, nullptr);
    }
    IC_CHECK_LAUNCH();
    int64_t t1 = (int64_t)N * D1 * H1 * W1;
    pc_conv_mid_kernel<KC, true, false><<<cdiv(t1, 128), 128, 0, s>>>(a0, D0, H0, W0, w.w1, w.b1, nullptr, 0, 0, 0, t1, a1);
    IC_CHECK_LAUNCH();
    int64_t t2 = (int64_t)N * D2 * H2 * W2;
    pc_conv_mid_kernel<KC, false, true><<<cdiv(t2, 128), 128, 0, s>>>(a1, D1, H1, W1, w.w2, w.b2, a0, D0, H0, W0, t2, a2);
    IC_CHECK_LAUNCH();
    int64_t t3 = (int64_t)N * D3 * H3 * W3;
    size_t smem = (size_t)(14 * KC * L + L) * sizeof(float);
    if (bits_sum && head != HEAD_LOGITS) IC_CHECK_CUDA(cudaMemsetAsync(bits_sum, 0, sizeof(double) * N, s));
    if (head == HEAD_LOGITS)
        pc_final_kernel<KC, HEAD_LOGITS><<<cdiv(t3, 128), 128, smem, s>>>(a2, D2, H2, W2, w.w3, w.b3, L, nullptr, t3, out_f, nullptr, nullptr);
    else if (head == HEAD_BITCOST)
        pc_final_kernel<KC, HEAD_BITCOST><<<cdiv(t3, 128), 128, smem, s>>>(a2, D2, H2, W2, w.w3, w.b3, L, in.target_symbols, t3, out_f, nullptr, bits_sum);
    else
        pc_final_kernel<KC, HEAD_FREQS><<<cdiv(t3, 128), 128, smem, s>>>(a2, D2, H2, W2, w.w3, w.b3, L, in.target_symbols, t3, nullptr, out_freqs, bits_sum,
                                                                         out_freqs32);
    IC_CHECK_LAUNCH();
    return IC_OK;
}


// ---------------------------------------------------------------- tensor-core path (K = 24)
// conv0 stays FFMA (Cin = 1, 13 taps) but writes the hi/lo plane layout; layers 1..3 run on the
// grouped-tap tcgen05 kernel (conv_tc.cu): "image" = one (n, depth) slice, group = filter depth,
// VALID taps, 24 -> 32 padded channels, float32-class hi/lo arithmetic.
int run_pc_tc(const PcWeights& w, const PcInput& in, int head, float* out_f, int64_t* out_freqs, double* bits_sum,
              void* ws, size_t ws_bytes, cudaStream_t s) {
    const int N = in.N, D = in.D, H = in.H, W = in.W, pd = in.pad_d, ph = in.pad_hw, L = w.L;
    const int Dp = D + pd, Hp = H + 2 * ph, Wp = W + 2 * ph;
    IC_REQUIRE(Dp >= 5 && Hp >= 9 && Wp >= 9, IC_ERR_INVALID, "probclass: volume %dx%dx%d smaller than the 5x9x9 context",
               Dp, Hp, Wp);
    int Dl[4], Hl[4], Wl[4];
    Dl[0] = Dp - 1; Hl[0] = Hp - 2; Wl[0] = Wp - 2;
    for (int l = 1; l < 4; ++l) {
        Dl[l] = Dl[l - 1] - 1; Hl[l] = Hl[l - 1] - 2; Wl[l] = Wl[l - 1] - 2;
    }
    Arena ar(ws, ws_bytes);
    __half* a[3];
    size_t plane[3];
    for (int l = 0; l < 3; ++l) {
        plane[l] = (size_t)N * Dl[l] * 4 * Hl[l] * Wl[l] * 8;
        a[l] = ar.get<__half>(2 * plane[l]);
    }
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "probclass workspace too small: need %zu, have %zu", ar.off, ws_bytes);
    const int64_t t0 = (int64_t)N * Dl[0] * Hl[0] * Wl[0];
    {
        ProfScope ps(IC_PROF_PROBCLASS, s, 1);
        if (in.symbols) {
            SrcSymbols src;
            src.sym = in.symbols;
            for (int i = 0; i < 8; ++i) src.centers[i] = i < L ? in.centers_host[i] : 0.f;
            src.pad_value = in.centers_host[0];
            pc_conv0_kernel<24, SrcSymbols, true><<<cdiv(t0, 128), 128, 0, s>>>(src, D, H, W, pd, ph, w.w0, w.b0, Dl[0], Hl[0],
                                                                                Wl[0], t0, nullptr, a[0]);
        } else {
            SrcFloat src{in.q, in.pad_value};
            pc_conv0_kernel<24, SrcFloat, true><<<cdiv(t0, 128), 128, 0, s>>>(src, D, H, W, pd, ph, w.w0, w.b0, Dl[0], Hl[0],
                                                                              Wl[0], t0, nullptr, a[0]);
        }
        IC_CHECK_LAUNCH();
    }
    if (bits_sum && head != PC_HEAD_LOGITS) IC_CHECK_CUDA(cudaMemsetAsync(bits_sum, 0, sizeof(double) * N, s));
    for (int l = 1; l <= 3; ++l) {
        tc::ConvTcArgs c;
        memset(&c, 0, sizeof(c));
        c.in = a[l - 1];
        c.Nimg = N * Dl[l - 1];
        c.in_chunks = 4;                        // pitch of the planes; only 3 chunks (24 channels) are ever written or read:
        c.in_chunks_valid = 3;                  // the 4th chunk of a group's box is TMA zero fill
        c.store_chunks = 3;
        c.Hin = Hl[l - 1];
        c.Win = Wl[l - 1];
        c.weights = w.wt[l - 1];
        c.groups = &w.gt[l - 1];
        c.scale = w.scale_t[l - 1];
        c.shift = w.shift_t[l - 1];
        c.N = N * Dl[l];
        c.H = Hl[l];
        c.W = Wl[l];
        c.halo0 = 0;
        c.img_mul = 1;
        c.img_div = Dl[l];
        c.img_div_mul = 1;                      // D_in - D_out
        c.cpg = 4;
        c.exact = 1;
        c.head = -1;
        c.pair_c2 = 1;                          // 24 channels: chunk 3 of every input voxel is zero
        c.prof_class = IC_PROF_PROBCLASS;
        if (l < 3) {
            c.out = a[l];
            c.nout = 32;
            c.cout = 24;
            c.relu = (l == 1);
            if (l == 2) {                       // + residual_input[..., 2:, 2:-2, 2:-2, :]  (code/probclass.py:196)
                c.res1 = a[0];
                c.res_H = Hl[0];
                c.res_W = Wl[0];
                c.res_dy = 2;
                c.res_dx = 2;
                c.res_div_mul = 2;              // D_res - D_out
                c.res_img_off = 2;
                c.res_plane = plane[0];
            }
        } else {
            c.nout = 16;
            c.cout = L;
            c.relu = 1;
            c.head = head;
            c.symbols = in.target_symbols;
            c.out_f32 = out_f;
            c.out_freqs = out_freqs;
            c.bits_sum = bits_sum;
        }
        int rc = tc::launch_conv_tc(c, s);
        if (rc != IC_OK) return rc;
    }
    return IC_OK;
}

}  // namespace

size_t pc_workspace_bytes(int KC, int N, int D, int H, int W, int pad_d, int pad_hw) {
    const int64_t Dp = D + pad_d, Hp = H + 2 * pad_hw, Wp = W + 2 * pad_hw;
    size_t b = 0;
    int64_t d = Dp, h = Hp, w = Wp;
    const size_t per_voxel = KC == 24 ? 32 * 4 : KC * sizeof(float);     // tensor-core path: 32 ch x (hi+lo)
    for (int l = 0; l < 3; ++l) {
        d -= 1; h -= 2; w -= 2;
        if (d <= 0 || h <= 0 || w <= 0) return 0;
        b = align_up(b, 256) + (size_t)N * d * h * w * per_voxel;
    }
    return b + 1024;
}

int pc_forward(const PcWeights& w, const PcInput& in, int head, float* out_f, int64_t* out_freqs,
               double* bits_sum, void* ws, size_t ws_bytes, cudaStream_t s, bool canonical, uint32_t* out_freqs32) {
    IC_REQUIRE(!out_freqs32 || canonical, IC_ERR_INVALID, "probclass: 32-bit tables are the codec (float32 chain) tables");
    if (w.K == 24 && w.tc && !canonical) return run_pc_tc(w, in, head, out_f, out_freqs, bits_sum, ws, ws_bytes, s);
    if (w.K == 24) return run_pc<24>(w, in, head, out_f, out_freqs, bits_sum, ws, ws_bytes, s, out_freqs32);
    if (w.K == 64) return run_pc<64>(w, in, head, out_f, out_freqs, bits_sum, ws, ws_bytes, s, out_freqs32);
    set_error("probclass: arch_param__k = %d not built (24 and 64 are)", w.K);
    return IC_ERR_UNSUPPORTED;
}

}  // namespace ic
