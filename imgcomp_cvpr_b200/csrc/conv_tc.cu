// 3x3 128->128 convolution on tcgen05 tensor cores (sm_100a): the 32 residual
// convs of _CVPR._encode / _decode (code/autoencoder.py:225-233,253-261,274-287),
// i.e. 95 % of the FLOPs of the hot path, with the fused epilogue
//     out = [relu]( acc * bn_scale + bn_shift ) + res1 + res2
// (fused batch norm inference, residual_block's "+ x", and the group skip).
//
// Arithmetic.  Operands are fp16, accumulation is fp32 in TMEM.  In EXACT mode
// every fp32 value v is carried as a pair (hi, lo) = (fp16(v), fp16(v - hi)) and
// the product a*w is evaluated as a_hi*w_hi + a_hi*w_lo + a_lo*w_hi (three MMAs
// into the same accumulator): each fp16 x fp16 product is exact in fp32, the
// dropped a_lo*w_lo term is < 2^-22 relative, so results are float32-class.  In
// FAST mode only the hi planes are used (one MMA).  Weights are pre-scaled by a
// power of two per layer so that w_lo stays in fp16's normal range.
//
// Data layout (HBM).  Activations: fp16, [plane(hi,lo)][N][C/8][H][W][8]
// ("NC/8HW8"): one TMA box {(TW*T+2)*8, TH+2, 8 chunks} lands in shared memory
// as [chunk][halo pixel][8 ch] = the UMMA K-major *no-swizzle* canonical layout
// (core matrix = 8 pixels x 16 B), in which a filter tap (dy,dx) is just a
// 16-byte-granular shift of the descriptor start address: the halo tile is
// loaded ONCE per 64-channel half and reused by all 9 taps.  TMA zero-fills
// outside the image = TF 'SAME' padding for free.
// Weights: fp16, pre-arranged on the host in exactly the shared-memory order of
// one pipeline stage (32 input channels of one tap, both planes) -> plain
// cp.async.bulk copies.
//
// Kernel structure (persistent, one CTA per SM, 7 warps):
//   warp 0  activation producer (TMA tensor loads, 2 half buffers)
//   warp 1  weight producer     (bulk copies, 3 stages)
//   warp 2  MMA issuer          (one elected lane; TMEM alloc/dealloc)
//   warps 3-6 epilogue          (tcgen05.ld -> BN/ReLU/residual -> hi/lo split -> st.global)
// A CTA iteration computes a 16 x (8*T) pixel super tile = T MMA tiles of 128
// pixels (M=128, N=128, K=16 UMMA) sharing each weight stage; accumulators are
// double buffered in TMEM (2 sets x T tiles x 128 columns) so the epilogue of
// super tile i overlaps the MMAs of i+1.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"
#include "conv_tc.cuh"

namespace ic {
namespace tc {

namespace {

constexpr int TH = 16, TW = 8;             // one UMMA tile: 16 rows x 8 pixels = M 128
constexpr int WSTAGES = 3;
constexpr int NTHREADS = 7 * 32;
constexpr int W_STAGE_PLANE_BYTES = 4 * 128 * 16;   // 4 chunks x 128 cout x 8 cin fp16
constexpr uint32_t kIdesc = (1u << 4) /* D fp32 */ | (0u << 7) /* A f16 */ | (0u << 10) /* B f16 */ |
                            ((128u >> 3) << 17) /* N */ | ((128u >> 4) << 24) /* M */;

template <int T>
struct Cfg {
    static constexpr int HALO_W = TW * T + 2, HALO_H = TH + 2, HALO_PIX = HALO_W * HALO_H;
    static constexpr int A_PLANE_BYTES = 8 * HALO_PIX * 16;       // 8 chunks of one 64-channel half
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

struct __align__(8) Barriers {
    uint64_t a_full[2], a_empty[2], w_full[WSTAGES], w_empty[WSTAGES], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float4 ld16(const void* p) { return *reinterpret_cast<const float4*>(p); }

// fp16 hi/lo chunk (8 channels) -> 8 floats
__device__ __forceinline__ void add_pair(const __half* hi_plane, const __half* lo_plane, size_t off, bool has_lo,
                                         float (&v)[8]) {
    float4 h4 = ld16(hi_plane + off);
    const __half2* h2 = reinterpret_cast<const __half2*>(&h4);
    if (has_lo) {
        float4 l4 = ld16(lo_plane + off);
        const __half2* l2 = reinterpret_cast<const __half2*>(&l4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 a = __half22float2(h2[i]), b = __half22float2(l2[i]);
            v[2 * i] += a.x + b.x;
            v[2 * i + 1] += a.y + b.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 a = __half22float2(h2[i]);
            v[2 * i] += a.x;
            v[2 * i + 1] += a.y;
        }
    }
}

template <int T, int NPL>
__global__ void __launch_bounds__(NTHREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap in_map, ConvTcParams p) {
    using C = Cfg<T>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_buf = smem;                                              // [2 halves][NPL][A_PLANE_BYTES]
    uint8_t* w_buf = a_buf + 2 * NPL * C::A_PLANE_BYTES;                // [WSTAGES][NPL][W_STAGE_PLANE_BYTES]
    float* s_scale = reinterpret_cast<float*>(w_buf + WSTAGES * NPL * W_STAGE_PLANE_BYTES);
    float* s_shift = s_scale + 128;
    Barriers* bars = reinterpret_cast<Barriers*>(s_shift + 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (p.W + TW * T - 1) / (TW * T), tiles_y = (p.H + TH - 1) / TH;
    const int n_super = p.N * tiles_y * tiles_x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars->a_full[i]), 1);
            mbar_init(smem_u32(&bars->a_empty[i]), 1);
            mbar_init(smem_u32(&bars->acc_full[i]), 1);
            mbar_init(smem_u32(&bars->acc_empty[i]), 128);
        }
        for (int i = 0; i < WSTAGES; ++i) {
            mbar_init(smem_u32(&bars->w_full[i]), 1);
            mbar_init(smem_u32(&bars->w_empty[i]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 128) {
        s_scale[threadIdx.x] = p.scale[threadIdx.x];
        s_shift[threadIdx.x] = p.shift[threadIdx.x];
    }
    if (warp == 2) {   // TMEM: 2 accumulator sets x T tiles x 128 fp32 columns (power of two >= 32)
        constexpr uint32_t ncols = (2 * T * 128 <= 256) ? 256 : 512;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                     "n"(ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================== activation producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int st = blockIdx.x; st < n_super; st += gridDim.x, ++it) {
                const int n = st / (tiles_y * tiles_x), r = st - n * tiles_y * tiles_x;
                const int y0 = (r / tiles_x) * TH, x0 = (r % tiles_x) * TW * T;
                for (int h = 0; h < 2; ++h) {
                    mbar_wait(smem_u32(&bars->a_empty[h]), (it & 1) ^ 1);
                    const uint32_t full = smem_u32(&bars->a_full[h]);
                    mbar_expect_tx(full, NPL * C::A_PLANE_BYTES);
                    for (int pl = 0; pl < NPL; ++pl)
                        tma_load_5d(smem_u32(a_buf + (h * NPL + pl) * C::A_PLANE_BYTES), &in_map, full, (x0 - 1) * 8,
                                    y0 - 1, h * 8, n, pl);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== weight producer =====================
        if (lane == 0) {
            uint32_t ws = 0;     // running stage counter
            for (int st = blockIdx.x; st < n_super; st += gridDim.x) {
                for (int s = 0; s < 36; ++s, ++ws) {      // (half, tap, cin32 block) in MMA order
                    const uint32_t slot = ws % WSTAGES, ph = (ws / WSTAGES) & 1;
                    mbar_wait(smem_u32(&bars->w_empty[slot]), ph ^ 1);
                    const uint32_t full = smem_u32(&bars->w_full[slot]);
                    mbar_expect_tx(full, NPL * W_STAGE_PLANE_BYTES);
                    // global stage = [2 planes][8 KB]; FAST mode copies the hi plane only
                    bulk_load(smem_u32(w_buf + slot * NPL * W_STAGE_PLANE_BYTES),
                              p.weights + (size_t)s * 2 * W_STAGE_PLANE_BYTES, NPL * W_STAGE_PLANE_BYTES, full);
                }
            }
        }
    } else if (warp == 2) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t it = 0, ws = 0;
            for (int st = blockIdx.x; st < n_super; st += gridDim.x, ++it) {
                const uint32_t set = it & 1;
                mbar_wait(smem_u32(&bars->acc_empty[set]), ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int h = 0; h < 2; ++h) {
                    mbar_wait(smem_u32(&bars->a_full[h]), it & 1);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(a_buf + h * NPL * C::A_PLANE_BYTES);
                    for (int tap = 0; tap < 9; ++tap) {
                        const int dy = tap / 3, dx = tap - dy * 3;
                        for (int j = 0; j < 2; ++j, ++ws) {
                            const uint32_t slot = ws % WSTAGES, ph = (ws / WSTAGES) & 1;
                            mbar_wait(smem_u32(&bars->w_full[slot]), ph);
                            tc_fence_after();
                            const uint32_t w_base = smem_u32(w_buf + slot * NPL * W_STAGE_PLANE_BYTES);
#pragma unroll
                            for (int t = 0; t < T; ++t) {
                                const uint32_t d_tmem = tmem_base + (set * T + t) * 128;
#pragma unroll
                                for (int ks = 0; ks < 2; ++ks) {
                                    const uint32_t a_off = (uint32_t)(j * 4 + ks * 2) * (C::HALO_PIX * 16) +
                                                           (uint32_t)(dy * C::HALO_W + dx + t * TW) * 16;
                                    const uint32_t w_off = (uint32_t)(ks * 2) * (128 * 16);
                                    const uint64_t a_hi = make_desc(a_base + a_off, C::HALO_PIX * 16, C::HALO_W * 16);
                                    const uint64_t w_hi = make_desc(w_base + w_off, 128 * 16, 128);
                                    const uint32_t first = (h | tap | j | ks) == 0 ? 0u : 1u;
                                    umma_f16(d_tmem, a_hi, w_hi, kIdesc, first);
                                    if (NPL == 2) {
                                        const uint64_t a_lo = make_desc(a_base + C::A_PLANE_BYTES + a_off, C::HALO_PIX * 16,
                                                                        C::HALO_W * 16);
                                        const uint64_t w_lo = make_desc(w_base + W_STAGE_PLANE_BYTES + w_off, 128 * 16, 128);
                                        umma_f16(d_tmem, a_hi, w_lo, kIdesc, 1u);
                                        umma_f16(d_tmem, a_lo, w_hi, kIdesc, 1u);
                                    }
                                }
                            }
                            umma_commit(smem_u32(&bars->w_empty[slot]));    // stage free once these MMAs retire
                        }
                    }
                    umma_commit(smem_u32(&bars->a_empty[h]));
                }
                umma_commit(smem_u32(&bars->acc_full[set]));
            }
        }
    } else {
        // ===================== epilogue (warps 3..6) =====================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;                  // accumulator row = pixel within the tile
        const int ty = m >> 3, tx = m & 7;
        const size_t plane = (size_t)p.N * 16 * p.H * p.W * 8;      // elements per hi/lo plane
        uint32_t it = 0;
        for (int st = blockIdx.x; st < n_super; st += gridDim.x, ++it) {
            const uint32_t set = it & 1;
            const int n = st / (tiles_y * tiles_x), r = st - n * tiles_y * tiles_x;
            const int y = (r / tiles_x) * TH + ty;
            mbar_wait(smem_u32(&bars->acc_full[set]), (it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int t = 0; t < T; ++t) {
                const int x = (r % tiles_x) * TW * T + t * TW + tx;
                const bool inside = y < p.H && x < p.W;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (set * T + t) * 128;
#pragma unroll 1
                for (int cc = 0; cc < 8; ++cc) {
                    uint32_t rr[16];
                    tmem_ld16(taddr + cc * 16, rr);
                    tmem_ld_wait();
                    if (inside) {
#pragma unroll
                        for (int hc = 0; hc < 2; ++hc) {
                            const int chunk = cc * 2 + hc;
                            float v[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                float a = __uint_as_float(rr[hc * 8 + e]);
                                a = fmaf(a, s_scale[chunk * 8 + e], s_shift[chunk * 8 + e]);
                                v[e] = p.relu ? fmaxf(a, 0.f) : a;
                            }
                            const size_t off = ((((size_t)n * 16 + chunk) * p.H + y) * p.W + x) * 8;
                            if (p.res1) add_pair(p.res1, p.res1 + plane, off, NPL == 2, v);
                            if (p.res2) add_pair(p.res2, p.res2 + plane, off, NPL == 2, v);
                            __align__(16) __half2 hi[4];
                            __align__(16) __half2 lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                hi[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                                float2 hf = __half22float2(hi[e]);
                                lo[e] = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
                            }
                            *reinterpret_cast<float4*>(p.out + off) = *reinterpret_cast<const float4*>(hi);
                            if (NPL == 2) *reinterpret_cast<float4*>(p.out + plane + off) = *reinterpret_cast<const float4*>(lo);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->acc_empty[set]));
        }
    }
    // teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        constexpr uint32_t ncols = (2 * T * 128 <= 256) ? 256 : 512;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ncols));
    }
}

// ------------------------------------------------------------- layout changes
// fp32 NHWC (N,H,W,128) <-> fp16 hi/lo planes [2][N][16][H][W][8]
__global__ void split_from_nhwc_kernel(const float* __restrict__ in, int H, int W, int64_t total_chunks, int64_t plane,
                                       __half* __restrict__ out, int write_lo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // one thread per (pixel, chunk), chunk fastest
    if (i >= total_chunks) return;
    int chunk = (int)(i & 15);
    int64_t pix = i >> 4;                 // n*H*W + y*W + x
    int64_t hw = (int64_t)H * W;
    int64_t n = pix / hw, r = pix - n * hw;
    const float4* src = reinterpret_cast<const float4*>(in + pix * 128 + chunk * 8);
    float4 a = src[0], b = src[1];
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    __align__(16) __half2 hi[4];
    __align__(16) __half2 lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        hi[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        float2 hf = __half22float2(hi[e]);
        lo[e] = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
    }
    size_t off = (((size_t)n * 16 + chunk) * hw + r) * 8;
    *reinterpret_cast<float4*>(out + off) = *reinterpret_cast<const float4*>(hi);
    if (write_lo) *reinterpret_cast<float4*>(out + plane + off) = *reinterpret_cast<const float4*>(lo);
}

__global__ void merge_to_nhwc_kernel(const __half* __restrict__ in, int H, int W, int64_t total_chunks, int64_t plane,
                                     float* __restrict__ out, int has_lo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_chunks) return;
    int chunk = (int)(i & 15);
    int64_t pix = i >> 4;
    int64_t hw = (int64_t)H * W;
    int64_t n = pix / hw, r = pix - n * hw;
    size_t off = (((size_t)n * 16 + chunk) * hw + r) * 8;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    add_pair(in, in + plane, off, has_lo != 0, v);
    float4* dst = reinterpret_cast<float4*>(out + pix * 128 + chunk * 8);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

template <int T, int NPL>
int launch_t(const ConvTcArgs& a, cudaStream_t s) {
    using C = Cfg<T>;
    EncodeTiledFn enc = get_encode_fn();
    IC_REQUIRE(enc, IC_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    CUtensorMap map;
    const cuuint64_t dims[5] = {(cuuint64_t)a.W * 8, (cuuint64_t)a.H, 16, (cuuint64_t)a.N, (cuuint64_t)NPL};
    const cuuint64_t strides[4] = {(cuuint64_t)a.W * 16, (cuuint64_t)a.H * a.W * 16, (cuuint64_t)16 * a.H * a.W * 16,
                                   (cuuint64_t)a.N * 16 * a.H * a.W * 16};
    const cuuint32_t box[5] = {(cuuint32_t)C::HALO_W * 8, (cuuint32_t)C::HALO_H, 8, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, (void*)a.in, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IC_REQUIRE(r == CUDA_SUCCESS, IC_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d (N=%d H=%d W=%d)", (int)r, a.N, a.H, a.W);
    ConvTcParams p;
    p.weights = (const uint8_t*)a.weights;
    p.scale = a.scale;
    p.shift = a.shift;
    p.res1 = a.res1;
    p.res2 = a.res2;
    p.out = a.out;
    p.N = a.N;
    p.H = a.H;
    p.W = a.W;
    p.relu = a.relu;
    const size_t smem = 2 * NPL * C::A_PLANE_BYTES + WSTAGES * NPL * W_STAGE_PLANE_BYTES + 1024 + sizeof(Barriers) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        IC_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<T, NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int tiles_x = (a.W + TW * T - 1) / (TW * T), tiles_y = (a.H + TH - 1) / TH;
    const int n_super = a.N * tiles_y * tiles_x;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n_super < sms ? n_super : sms;
    ProfScope ps(IC_PROF_CONV3X3, s);
    conv3x3_tc_kernel<T, NPL><<<grid, NTHREADS, smem, s>>>(map, p);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

}  // namespace

int launch_conv3x3_tc(const ConvTcArgs& a, cudaStream_t s) {
    IC_REQUIRE(a.N > 0 && a.H > 0 && a.W > 0, IC_ERR_INVALID, "conv3x3_tc: bad shape");
    IC_REQUIRE(((uintptr_t)a.in & 15) == 0 && ((uintptr_t)a.out & 15) == 0, IC_ERR_INVALID, "conv3x3_tc: unaligned buffers");
    const bool wide = a.W > 8;
    if (a.exact) return wide ? launch_t<2, 2>(a, s) : launch_t<1, 2>(a, s);
    return wide ? launch_t<2, 1>(a, s) : launch_t<1, 1>(a, s);
}

int launch_split_from_nhwc(const float* in, int N, int H, int W, __half* out, int write_lo, cudaStream_t s) {
    int64_t total = (int64_t)N * H * W * 16, plane = (int64_t)N * H * W * 128;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    split_from_nhwc_kernel<<<cdiv(total, 256), 256, 0, s>>>(in, H, W, total, plane, out, write_lo);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int launch_merge_to_nhwc(const __half* in, int N, int H, int W, float* out, int has_lo, cudaStream_t s) {
    int64_t total = (int64_t)N * H * W * 16, plane = (int64_t)N * H * W * 128;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    merge_to_nhwc_kernel<<<cdiv(total, 256), 256, 0, s>>>(in, H, W, total, plane, out, has_lo);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

// Host: HWIO float weights (3,3,128,128) -> the kernel's stage order, fp16 hi/lo, scaled by 2^e.
// Stage s = (half h, tap, cin32 block j) in MMA order; within a stage [plane][4 chunks][128 cout][8 cin].
void pack_weights_3x3(const float* w_hwio, std::vector<__half>& packed, float* inv_scale_out) {
    float mx = 0.f;
    for (int i = 0; i < 9 * 128 * 128; ++i) mx = fmaxf(mx, fabsf(w_hwio[i]));
    int e = 0;
    if (mx > 0.f) {
        int ex;
        frexpf(mx, &ex);        // mx = f * 2^ex, f in [0.5,1)
        e = 8 - ex;             // scaled max in [2^7, 2^8): w_lo stays a normal fp16 number
    }
    const float sc = ldexpf(1.f, e);
    *inv_scale_out = ldexpf(1.f, -e);
    packed.assign((size_t)36 * 2 * 4 * 128 * 8, __float2half(0.f));
    for (int h = 0; h < 2; ++h)
        for (int tap = 0; tap < 9; ++tap)
            for (int j = 0; j < 2; ++j) {
                const int s = (h * 9 + tap) * 2 + j;
                for (int ch = 0; ch < 4; ++ch)
                    for (int co = 0; co < 128; ++co)
                        for (int ei = 0; ei < 8; ++ei) {
                            const int ci = h * 64 + j * 32 + ch * 8 + ei;
                            const float v = w_hwio[((size_t)tap * 128 + ci) * 128 + co] * sc;
                            const __half hi = __float2half_rn(v);
                            const __half lo = __float2half_rn(v - __half2float(hi));
                            const size_t base = (size_t)s * 2 * 4 * 128 * 8;
                            const size_t idx = ((size_t)ch * 128 + co) * 8 + ei;
                            packed[base + idx] = hi;
                            packed[base + 4 * 128 * 8 + idx] = lo;
                        }
            }
}

}  // namespace tc
}  // namespace ic
