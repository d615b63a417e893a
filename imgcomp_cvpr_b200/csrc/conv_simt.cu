// float32 FFMA implicit-GEMM convolution (IC_MODE_FP32).
//
// Computes what slim.conv2d / slim.conv2d_transpose + fused batch norm + ReLU
// compute in the reference (code/autoencoder.py:106-125,222-237,251,264-265),
// plus the residual additions of residual_block (:274-287) and the skip
// connections of _CVPR._encode/_decode (:231,234,259,262), and for h13 the
// _denormalize + _clip_to_image_range epilogue (:146-158).
//
// GEMM view: M = N*Ho*Wo output pixels, N = Cout, K = KH*KW*Cin.
// 64x64 block tile, BK = 16, 256 threads, 4x4 register tile, double buffered.
// Stride-2 transposed convs: an output pixel of phase (oy & 1, ox & 1) only meets the taps of matching parity (9, 6, 6
// or 4 of 25; 1, 2, 2 or 4 of 9), so blocks are formed over pixels of ONE phase and walk only those taps, in ascending
// tap order: the same non-zero terms in the same order as the plain loop (bit-identical), a quarter of the work.
#include "common.cuh"

namespace ic {

namespace {

constexpr int BK = 16, NT = 256;

__constant__ float c_norm_mean[3] = {121.853699f, 113.588608f, 100.637154f};
// np.sqrt(var + 1e-10) evaluated in float32 (code/autoencoder.py:143,153,160-169)
__constant__ float c_norm_std[3] = {68.8939514f, 66.7393417f, 69.3702698f};

// BN_ = 64: 64 pixels x 64 channels per block; BN_ = 32: 128 pixels x 32 channels (layers with <= 32 output channels:
// the context model's 24, h13's 3, the data gradient into 4-channel inputs).  Per-output arithmetic (one fmaf chain over
// k in ascending order) does not depend on the tile shape.
template <int BN_>
__global__ void __launch_bounds__(NT) conv_simt_kernel(ConvDesc d) {
    constexpr int BM = 64 * 64 / BN_;            // pixels per block
    constexpr int TXN = BN_ / 4;                 // threads across the channel dimension
    constexpr int AR = BM / 64;                  // A rows loaded per thread
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN_ + 4];

    const int t = threadIdx.x;
    const bool phased = d.transposed && d.stride == 2;
    // phase (pa, pb) = (oy & 1, ox & 1) owns ceil((Ho - pa) / 2) x ceil((Wo - pb) / 2) pixels per image; the grid is the
    // concatenation of the four phases' block ranges (launch_conv_simt computes the same counts)
    int phase = 0, blk = blockIdx.x;
    if (phased) {
        for (; phase < 3; ++phase) {
            const int64_t mp = (int64_t)d.N * ((d.Ho - (phase >> 1) + 1) / 2) * ((d.Wo - (phase & 1) + 1) / 2);
            const int nb = (int)((mp + BM - 1) / BM);
            if (blk < nb) break;
            blk -= nb;
        }
    }
    const int pa = phase >> 1, pb = phase & 1;
    const int Hq = (d.Ho - pa + 1) / 2, Wq = (d.Wo - pb + 1) / 2;
    const int64_t M = phased ? (int64_t)d.N * Hq * Wq : (int64_t)d.N * d.Ho * d.Wo;      // pixels (of this phase)
    const int64_t m0 = (int64_t)blk * BM;
    const int n0 = blockIdx.y * BN_;
    // taps this block walks: all, or those of the phase's parity
    const int ky0 = phased ? ((pa + d.pad_t) & 1) : 0, kx0 = phased ? ((pb + d.pad_l) & 1) : 0;
    const int kstep = phased ? 2 : 1;
    const int nky = phased ? (d.KH - ky0 + 1) / 2 : d.KH, nkx = phased ? (d.KW - kx0 + 1) / 2 : d.KW;
    const int K = nky * nkx * d.Cin;
    // row of the tile -> (image, output y, output x); returns false past the end
    auto row_pixel = [&](int64_t mrow, int& pn, int& oy, int& ox) -> bool {
        if (mrow >= M) return false;
        if (!phased) {
            pn = (int)(mrow / ((int64_t)d.Ho * d.Wo));
            int r = (int)(mrow - (int64_t)pn * d.Ho * d.Wo);
            oy = r / d.Wo;
            ox = r - oy * d.Wo;
        } else {
            pn = (int)(mrow / ((int64_t)Hq * Wq));
            int r = (int)(mrow - (int64_t)pn * Hq * Wq);
            int yq = r / Wq;
            oy = 2 * yq + pa;
            ox = 2 * (r - yq * Wq) + pb;
        }
        return true;
    };

    // A-load assignment: AR output pixel rows (64 apart), 4 consecutive k
    const int am = t >> 2, ak = (t & 3) * 4;
    int pn[AR], oy[AR], ox[AR];
    bool mvalid[AR];
#pragma unroll
    for (int j = 0; j < AR; ++j) {
        pn[j] = oy[j] = ox[j] = 0;
        mvalid[j] = row_pixel(m0 + am + 64 * j, pn[j], oy[j], ox[j]);
    }
    // B-load assignment (the first 4 * BN_ threads)
    const int bk = t / TXN, bn = (t % TXN) * 4;
    const bool bload = t < BK * TXN;

    auto load_a = [&](int k0, int j) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int k = k0 + ak;
        if (mvalid[j] && k < K) {
            int tap = k / d.Cin, ci = k - tap * d.Cin;
            int jy = tap / nkx;
            int ky = ky0 + kstep * jy, kx = kx0 + kstep * (tap - jy * nkx);
            int iy, ix;
            bool ok;
            if (!d.transposed) {
                iy = oy[j] * d.stride - d.pad_t + ky;
                ix = ox[j] * d.stride - d.pad_l + kx;
                ok = iy >= 0 && iy < d.Hi && ix >= 0 && ix < d.Wi;
            } else {
                int ny = oy[j] + d.pad_t - ky, nx = ox[j] + d.pad_l - kx;
                ok = ny >= 0 && nx >= 0 && (ny % d.stride) == 0 && (nx % d.stride) == 0;
                iy = ny / d.stride;
                ix = nx / d.stride;
                ok = ok && iy < d.Hi && ix < d.Wi;
            }
            if (ok) v = *reinterpret_cast<const float4*>(d.in + (((int64_t)pn[j] * d.Hi + iy) * d.Wi + ix) * d.Cin + ci);
        }
        return v;
    };
    auto load_b = [&](int k0) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int k = k0 + bk;
        if (bload && k < K && n0 + bn < d.ldw) {
            int tap = k / d.Cin, ci = k - tap * d.Cin;
            int jy = tap / nkx;
            int64_t row = (int64_t)((ky0 + kstep * jy) * d.KW + kx0 + kstep * (tap - jy * nkx)) * d.Cin + ci;
            v = *reinterpret_cast<const float4*>(d.w + row * d.ldw + n0 + bn);
        }
        return v;
    };
    auto store = [&](int buf, const float4 (&a)[AR], float4 b) {
#pragma unroll
        for (int j = 0; j < AR; ++j) {
            As[buf][ak + 0][am + 64 * j] = a[j].x;
            As[buf][ak + 1][am + 64 * j] = a[j].y;
            As[buf][ak + 2][am + 64 * j] = a[j].z;
            As[buf][ak + 3][am + 64 * j] = a[j].w;
        }
        if (bload) *reinterpret_cast<float4*>(&Bs[buf][bk][bn]) = b;
    };

    const int tx = t % TXN, ty = t / TXN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float4 ra[AR], rb;
#pragma unroll
    for (int j = 0; j < AR; ++j) ra[j] = load_a(0, j);
    rb = load_b(0);
    store(0, ra, rb);
    __syncthreads();
    const int nk = (K + BK - 1) / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) {
#pragma unroll
            for (int j = 0; j < AR; ++j) ra[j] = load_a((kt + 1) * BK, j);
            rb = load_b((kt + 1) * BK);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store(buf ^ 1, ra, rb);
            __syncthreads();
        }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int en, ey, ex;
        if (!row_pixel(m0 + ty * 4 + i, en, ey, ex)) continue;
        const int64_t m = ((int64_t)en * d.Ho + ey) * d.Wo + ex;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= d.Cout) continue;
            float v = fmaf(acc[i][j], d.scale[n], d.shift[n]);
            if (d.relu) v = fmaxf(v, 0.f);
            int64_t o = m * d.Cout + n;
            if (d.res1) v += d.res1[o];
            if (d.res2) v += d.res2[o];
            if (!d.out_nchw) {
                d.out[o] = v;
            } else {
                if (d.denorm) {
                    v = __fadd_rn(__fmul_rn(v, c_norm_std[n]), c_norm_mean[n]);
                    v = fminf(fmaxf(v, 0.f), 255.f);
                }
                int64_t hw = (int64_t)d.Ho * d.Wo;
                int64_t img = m / hw, r = m - img * hw;
                int64_t oo = (img * d.Cout + n) * hw + r;
                d.out[oo] = v;
                if (d.out_u8) d.out_u8[oo] = (uint8_t)v;   // tf.cast truncation (val.py:91); v in [0,255]
            }
        }
    }
}

}  // namespace

int launch_conv_simt(const ConvDesc& d, cudaStream_t stream) {
    IC_REQUIRE(d.Cin % 4 == 0 && d.ldw % 4 == 0, IC_ERR_INVALID, "conv_simt: Cin (%d) and ldw (%d) must be multiples of 4",
               d.Cin, d.ldw);
    const bool narrow = d.Cout <= 32;
    const int BN = narrow ? 32 : 64, BM = 64 * 64 / BN;
    int64_t M = (int64_t)d.N * d.Ho * d.Wo;
    dim3 grid(cdiv(M, BM), cdiv(d.Cout, BN));
    if (d.transposed && d.stride == 2) {        // one block range per output phase
        grid.x = 0;
        for (int ph = 0; ph < 4; ++ph)
            grid.x += cdiv((int64_t)d.N * ((d.Ho - (ph >> 1) + 1) / 2) * ((d.Wo - (ph & 1) + 1) / 2), BM);
        if (grid.x == 0) return IC_OK;
    }
    const bool res_conv = d.KH == 3 && !d.transposed && d.Cin == 128 && d.Cout == 128;
    ProfScope ps(res_conv ? IC_PROF_CONV3X3 : IC_PROF_CONV_OTHER, stream);
    if (narrow) conv_simt_kernel<32><<<grid, NT, 0, stream>>>(d);
    else conv_simt_kernel<64><<<grid, NT, 0, stream>>>(d);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

}  // namespace ic
