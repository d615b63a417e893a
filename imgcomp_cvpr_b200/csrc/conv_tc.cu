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
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace ic {
namespace tc {

namespace {

constexpr int TH = 16, TW = 8;             // one UMMA tile: 16 rows x 8 pixels = M 128
constexpr int MAX_WSTAGES = 16;
static_assert(MAX_WSTAGES >= 8, "Cfg::WSTAGES must fit the barrier arrays");
constexpr int MAX_RES_STAGES = 14;          // resident-weight mode: the context model has 14 non-masked taps
// activation-tile slots: 2 for the streamed-weight kernels (a group's MMAs take longer than a load), 6 for the context
// model: its tiles have two short groups (9 + 5 narrow MMAs, ~1k tensor cycles per tile) and a load from L2 / HBM takes
// 2-3k cycles, so three tiles must be in flight.  ncu (profiles/r2x_pc_raw.csv) with 2 slots and two CTAs per SM: tensor pipe
// 32 %, operand pipe 57 %, DRAM 26 %, i.e. nothing busy -- the issuer waited for a_full.
constexpr int MAX_ASLOTS = 6;
// epilogue warps: 4 (one per TMEM lane quarter), or 8 for the 128-column plane-output kernels (two per quarter, half of the
// output channels each): with the issue loop and the pair's operand traffic out of the way the EPILOGUE paced the 3x3
// convs (IC_TC_DBG: the issuer waited 20 % of its time on acc_empty) -- its residual loads are latency bound, and twice
// the warps keep twice the loads in flight
// (also the context model's plane-output layers: layer 2 reads a residual, IC_TC_DBG=2 showed its issuer waiting 24 % on acc_empty)
// and the decoder's depth-to-space convs (256 columns: their issuer waited 75-83 % on acc_empty, profiles/r2z5_issuer_all.txt)
constexpr int epi_warps(int nout, int outmode) { return ((nout == 128 || nout == 32 || nout == 256) && outmode == 0) ? 8 : 4; }
constexpr int nthreads(int nout, int outmode) { return (3 + epi_warps(nout, outmode)) * 32; }

template <int T, int NOUT, int CPG>
struct Cfg {
    static constexpr int HALO_W = TW * T + 2, HALO_H = TH + 2, HALO_PIX = HALO_W * HALO_H;
    static constexpr int A_PLANE_BYTES = CPG * HALO_PIX * 16;     // one group = CPG chunks of 8 channels
    // weight pipeline depth: a refill takes ~1 us (commit -> producer -> L2 -> smem) while a stage is consumed in
    // 0.3-0.5 us, so >= 4 stages must be in flight; 8 x 16 KB fits beside two 41 KB activation groups
    static constexpr int WSTAGES = NOUT == 256 ? 4 : (NOUT == 128 ? 8 : 6);
    static constexpr int W_PLANE_BYTES = 4 * NOUT * 16;           // one stage = 32 channels of one tap
    static constexpr int NCOL = NOUT > 128 ? 256 : (NOUT > 64 ? 128 : 64);   // TMEM columns per accumulator tile
    static constexpr uint32_t IDESC = (1u << 4) /* D fp32 */ | (0u << 7) /* A f16 */ | (0u << 10) /* B f16 */ |
                                      ((uint32_t)(NOUT >> 3) << 17) /* N */ | ((128u >> 4) << 24) /* M */;
};

__constant__ float c_img_mean[3] = {121.853699f, 113.588608f, 100.637154f};
__constant__ float c_img_std[3] = {68.8939514f, 66.7393417f, 69.3702698f};   // float32 sqrt(var + 1e-10)

struct __align__(8) Barriers {
    uint64_t a_full[MAX_ASLOTS], a_empty[MAX_ASLOTS], w_full[MAX_WSTAGES], w_empty[MAX_WSTAGES], acc_full[4], acc_empty[4];
    uint32_t tmem_base;      // (4 accumulator barriers: the context model's depth walk keeps a ring of 4 tiles, everything else 2)
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

__device__ __forceinline__ float4 ldg16(const __half* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// residual operands of one 16-column step (2 chunks): res1 hi/lo, res2 hi/lo
struct ResRegs {
    float4 v[2][4];
};

template <int NPL, bool RES2 = true>
__device__ __forceinline__ void load_res(ResRegs& r, const ConvTcParams& p, size_t off1, size_t stride1, size_t plane1,
                                         size_t off2, size_t stride2, size_t plane2) {
#pragma unroll
    for (int hc = 0; hc < 2; ++hc) {
        if (p.res1) {
            r.v[hc][0] = ldg16(p.res1 + off1 + hc * stride1);
            if (NPL == 2) r.v[hc][1] = ldg16(p.res1 + plane1 + off1 + hc * stride1);
        }
        if (RES2 && p.res2) {
            r.v[hc][2] = ldg16(p.res2 + off2 + hc * stride2);
            if (NPL == 2) r.v[hc][3] = ldg16(p.res2 + plane2 + off2 + hc * stride2);
        }
    }
}

__device__ __forceinline__ void add_h8(const float4& q, float (&v)[8]) {
    const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 a = __half22float2(h2[i]);
        v[2 * i] += a.x;
        v[2 * i + 1] += a.y;
    }
}

// pair (hi, lo) -> float, added as (hi + lo) like a float32 residual read
__device__ __forceinline__ void add_h8_pair(const float4& qh, const float4& ql, float (&v)[8]) {
    const __half2* h2 = reinterpret_cast<const __half2*>(&qh);
    const __half2* l2 = reinterpret_cast<const __half2*>(&ql);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 a = __half22float2(h2[i]), b = __half22float2(l2[i]);
        v[2 * i] += a.x + b.x;
        v[2 * i + 1] += a.y + b.y;
    }
}

// depth walk of the context model (ConvTcParams::walk): work item w = (image, tile, depth segment); the segment produces
// output slices [da, db) from input slices da .. db
struct WalkItem {
    int img, r, da, db;
};
__device__ __forceinline__ WalkItem walk_item(int w, int tiles, const ConvTcParams& p) {
    const int seg = w % p.walk_nseg, col = w / p.walk_nseg;
    WalkItem it;
    it.r = col % tiles;
    it.img = col / tiles;
    it.da = seg * p.walk;
    it.db = min(p.img_div, it.da + p.walk);
    return it;
}

// OUTMODE 0: fp16 hi/lo planes [plane][N][NOUT/8][H][W][8]; 1: float32 NHWC [N][H][W][cout];
// 2: context-model head (ReLU logits -> bit cost / coder frequencies / logits), code/probclass.py:100-104,443-444
// WRES: all weight stages of the layer stay resident in shared memory (loaded once per CTA; context model)
// PAIR: 2-CTA clusters, cta_group::2 MMAs (M = 256 over both CTAs, each CTA holds half of B); w_map is the
//       tensor map of the pair-packed weights [stage][half][plane][4][64][8] (unused otherwise).
template <int T, int NPL, int NOUT, int OUTMODE, int CPG, bool WRES, bool PAIR>
__global__ void __launch_bounds__(nthreads(NOUT, OUTMODE), (NOUT <= 32 && T == 1) ? 2 : 1)      // context model: two CTAs per SM
conv_tc_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap w_map, const ConvTcParams p,
               const GroupTable gt) {
    using C = Cfg<T, NOUT, CPG>;
    static_assert(!PAIR || (NOUT == 128 && OUTMODE == 0 && !WRES), "pair mode is built for the 128->128 convs");
    constexpr int W_ROWS = PAIR ? NOUT / 2 : NOUT;                 // B rows held by this CTA
    constexpr int W_PLANE = 4 * W_ROWS * 16;                       // bytes of one plane of one stage in this CTA
    constexpr uint32_t IDESC = PAIR ? ((C::IDESC & ~(0x1Fu << 24)) | ((256u >> 4) << 24)) : C::IDESC;
    // Resident-weight instantiations (context model, NOUT = 32 / 16) use B-concatenation: a stage is stored
    // [4 chunks][hi rows | lo rows][8 cin], a_hi * [w_hi | w_lo] is ONE MMA of N = 2 NOUT into columns [hh | x] of the tile
    // and a_lo * w_hi a second one into x: two A-operand fetches per k-step instead of three (these narrow MMAs are bound
    // by the 4 KB A fetch, not by the tensor pipe), and the cross terms get their own accumulator (see conv_cat_kernel).
    constexpr bool CAT = WRES;
    static_assert(!CAT || (2 * NOUT <= C::NCOL && NPL == 2 && !PAIR), "B-concatenation needs 2 NOUT accumulator columns per tile");
    constexpr uint32_t IDESC_CAT = (C::IDESC & ~(0x3Fu << 17)) | ((uint32_t)((2 * NOUT) >> 3) << 17);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    const int cta_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // work-loop start
    const int cta_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_buf = smem;                                              // [2 slots][NPL][A_PLANE_BYTES]
    // pair mode: a stage is half the size per CTA and its hand-over crosses the cluster, so twice the stages are in flight
    constexpr int WSTAGES = WRES ? MAX_RES_STAGES : (PAIR ? 2 * C::WSTAGES : C::WSTAGES);
    const uint32_t ASLOTS = (uint32_t)p.aslots;
    uint8_t* w_buf = a_buf + ASLOTS * NPL * C::A_PLANE_BYTES;           // [WSTAGES][NPL][W_PLANE]
    float* s_scale = reinterpret_cast<float*>(w_buf + WSTAGES * NPL * W_PLANE);
    float* s_shift = s_scale + 128;
    Barriers* bars = reinterpret_cast<Barriers*>(s_shift + 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (p.W + TW * T - 1) / (TW * T), tiles_y = (p.H + TH - 1) / TH;
    const int n_super = p.N * tiles_y * tiles_x;
    int n_work = PAIR ? (n_super + 1) / 2 : n_super;              // pair mode: one work item = two super tiles
    const int n_walk = (CAT && T == 1 && p.walk) ? (p.N / p.img_div) * tiles_y * tiles_x * p.walk_nseg : 0;
    // (context model: room for the depth walk's ring of 4 accumulators of 2 NOUT columns; two CTAs per SM still fit 512)
    constexpr uint32_t kAccCols = (WRES && T == 1 && 8 * NOUT > 2 * T * C::NCOL) ? 8 * NOUT : 2 * T * C::NCOL;
    constexpr uint32_t kTmemCols = (kAccCols <= 32) ? 32 : (kAccCols <= 64) ? 64 : (kAccCols <= 128) ? 128 : (kAccCols <= 256) ? 256 : 512;

    if (threadIdx.x == 0) {
        for (int i = 0; i < (int)ASLOTS; ++i) {
            // pair: ONE arrival (the leader's producer, which announces the bytes of BOTH CTAs); the peer's TMA only adds
            // its complete_tx.  A remote arrive.expect_tx per stage from the peer's producer thread (release.cluster) made
            // that thread the bottleneck: IC_TC_DBG showed the issuer waiting on w_full 74 % of the time.
            mbar_init(smem_u32(&bars->a_full[i]), 1);
            mbar_init(smem_u32(&bars->a_empty[i]), 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(smem_u32(&bars->acc_full[i]), 1);
            mbar_init(smem_u32(&bars->acc_empty[i]), (PAIR ? 2 : 1) * 32 * epi_warps(NOUT, OUTMODE));
        }
        for (int i = 0; i < (WRES ? 1 : WSTAGES); ++i) {      // resident mode uses w_full[0] only
            mbar_init(smem_u32(&bars->w_full[i]), 1);
            mbar_init(smem_u32(&bars->w_empty[i]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 128) {
        s_scale[threadIdx.x] = threadIdx.x < NOUT ? p.scale[threadIdx.x] * (CAT ? 1.f : p.acc_gain) : 0.f;
        s_shift[threadIdx.x] = threadIdx.x < NOUT ? p.shift[threadIdx.x] : 0.f;
    }
    if (warp == 2) {   // TMEM: 2 accumulator sets x T tiles x NCOL fp32 columns
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                         "n"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                         "n"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();       // peer barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================== activation producer =====================
        if (lane == 0) {
            uint32_t gi = 0;     // running group counter -> A slot / phase
            if constexpr (CAT && T == 1) {
                if (p.walk) {       // depth walk: one load per input slice of the segment
                    for (int wi = cta_id; wi < n_walk; wi += cta_stride) {
                        const WalkItem w = walk_item(wi, tiles_y * tiles_x, p);
                        const int y0 = (w.r / tiles_x) * TH, x0 = (w.r % tiles_x) * TW;
                        for (int j = w.da; j <= w.db; ++j, ++gi) {
                            const uint32_t slot = gi % ASLOTS, ph = (gi / ASLOTS) & 1;
                            mbar_wait(smem_u32(&bars->a_empty[slot]), ph ^ 1);
                            const int img = w.img * (p.img_div + p.img_div_mul) + j + p.img_base;
                            const uint32_t full = smem_u32(&bars->a_full[slot]);
                            mbar_expect_tx(full, NPL * C::A_PLANE_BYTES);
                            for (int pl = 0; pl < NPL; ++pl)
                                tma_load_5d(smem_u32(a_buf + (slot * NPL + pl) * C::A_PLANE_BYTES), &in_map, full,
                                            (x0 + p.halo0) * 8, y0 + p.halo0, 0, img, pl);
                        }
                    }
                    n_work = 0;
                }
            }
            for (int wi = cta_id; wi < n_work; wi += cta_stride) {
                const int st = PAIR ? 2 * wi + (int)rank : wi;      // past-the-end super tiles load zeros (TMA OOB)
                const int n = st / (tiles_y * tiles_x), r = st - n * tiles_y * tiles_x;
                const int y0 = (r / tiles_x) * TH, x0 = (r % tiles_x) * TW * T;
                for (int g = 0; g < gt.ngroups; ++g, ++gi) {
                    const uint32_t slot = gi % ASLOTS, ph = (gi / ASLOTS) & 1;
                    mbar_wait(smem_u32(&bars->a_empty[slot]), ph ^ 1);
                    const int img = n * p.img_mul + (p.img_div > 0 ? (n / p.img_div) * p.img_div_mul : 0) + gt.img_off[g] * p.img_off_mul + p.img_base;
                    if (PAIR) {
                        const uint32_t full = mapa_rank(smem_u32(&bars->a_full[slot]), 0);     // the leader's barrier
                        if (rank == 0) mbar_expect_tx(smem_u32(&bars->a_full[slot]), 2 * NPL * C::A_PLANE_BYTES);   // both tiles
                        for (int pl = 0; pl < NPL; ++pl)
                            tma_load_5d_2sm(smem_u32(a_buf + (slot * NPL + pl) * C::A_PLANE_BYTES), &in_map, full,
                                            (x0 + p.halo0) * 8, y0 + p.halo0, gt.chunk0[g], img, pl);
                    } else {
                        const uint32_t full = smem_u32(&bars->a_full[slot]);
                        mbar_expect_tx(full, NPL * C::A_PLANE_BYTES);
                        for (int pl = 0; pl < NPL; ++pl)
                            tma_load_5d(smem_u32(a_buf + (slot * NPL + pl) * C::A_PLANE_BYTES), &in_map, full,
                                        (x0 + p.halo0) * 8, y0 + p.halo0, gt.chunk0[g], img, pl);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== weight producer =====================
        if (lane == 0) {
            if (WRES) {
                // every stage of the layer fits: one load, one barrier (w_full[0]), never refilled
                const uint32_t full = smem_u32(&bars->w_full[0]);
                mbar_expect_tx(full, gt.nstages * NPL * C::W_PLANE_BYTES);
                bool walk_layout = false;
                if constexpr (CAT && T == 1) walk_layout = p.walk != 0;
                if (walk_layout) {
                    // depth walk: taps 0..4 are stored as ONE operand per tap for both filter depths,
                    //   [4 chunks][W1 lo | W1 hi | W0 hi | W0 lo rows][8 cin]       (W0 = filter depth 0 = global stage t,
                    //                                                                W1 = filter depth 1 = global stage 9 + t)
                    // so that a_hi * (all 4 NOUT rows) and a_lo * (the middle 2 NOUT rows) are one MMA each; taps 5..8 exist at
                    // depth 0 only and keep the [4 chunks][hi | lo][8] stage.  Built from the global stages by bulk copies.
                    constexpr uint32_t kStage1 = NPL * C::W_PLANE_BYTES, kRows = NOUT * 16;      // bytes: one stage, NOUT rows of a chunk
                    for (int t = 0; t < 5; ++t)
                        for (int c = 0; c < 4; ++c) {
                            const uint32_t dst = smem_u32(w_buf + t * 2 * kStage1 + c * 4 * kRows);
                            const uint8_t* w0 = p.weights + (size_t)t * kStage1 + c * 2 * kRows;
                            const uint8_t* w1 = p.weights + (size_t)(9 + t) * kStage1 + c * 2 * kRows;
                            bulk_load(dst, w1 + kRows, kRows, full);
                            bulk_load(dst + kRows, w1, kRows, full);
                            bulk_load(dst + 2 * kRows, w0, 2 * kRows, full);
                        }
                    for (int t = 5; t < 9; ++t)
                        bulk_load(smem_u32(w_buf + 10 * kStage1 + (t - 5) * kStage1), p.weights + (size_t)t * kStage1, kStage1, full);
                } else
                for (int s = 0; s < gt.nstages; ++s)
                    bulk_load(smem_u32(w_buf + s * NPL * C::W_PLANE_BYTES), p.weights + (size_t)s * 2 * C::W_PLANE_BYTES,
                              NPL * C::W_PLANE_BYTES, full);
            } else {
                uint32_t ws = 0;     // running stage counter
                for (int wi = cta_id; wi < n_work; wi += cta_stride) {
                    for (int s = 0; s < gt.nstages; ++s, ++ws) {      // (group, tap) in MMA order
                        const uint32_t slot = ws % WSTAGES, ph = (ws / WSTAGES) & 1;
                        mbar_wait(smem_u32(&bars->w_empty[slot]), ph ^ 1);
                        if (PAIR) {
                            // this CTA's half of the stage (64 of the 128 B rows), signalled on the leader's barrier;
                            // pair-packed global layout [stage][half][plane][4][64][8] seen as [stage*2+half][16][256]
                            const uint32_t full = mapa_rank(smem_u32(&bars->w_full[slot]), 0);
                            if (rank == 0) mbar_expect_tx(smem_u32(&bars->w_full[slot]), 2 * NPL * W_PLANE);      // both halves
                            tma_load_3d_2sm(smem_u32(w_buf + slot * NPL * W_PLANE), &w_map, full, 0, 0, s * 2 + (int)rank);
                        } else {
                            const uint32_t full = smem_u32(&bars->w_full[slot]);
                            mbar_expect_tx(full, NPL * C::W_PLANE_BYTES);
                            // global stage = [2 planes][W_PLANE_BYTES]; FAST mode copies the hi plane only
                            bulk_load(smem_u32(w_buf + slot * NPL * C::W_PLANE_BYTES),
                                      p.weights + (size_t)s * 2 * C::W_PLANE_BYTES, NPL * C::W_PLANE_BYTES, full);
                        }
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ===================== MMA issuer =====================
        // The WHOLE warp walks the loops (waits included) and one elected lane issues: with warp-uniform control flow the
        // descriptors stay in uniform registers and a tcgen05.mma costs ~4 issue slots.  Under `if (lane == 0)` every
        // operand lived in a vector register and each MMA paid an ELECT / R2UR.BROADCAST x6 / BRA.U.ANY waterfall (~17
        // dependent instructions, ~60-100 cycles): the single issuing thread, not the tensor pipe, was the bottleneck
        // (profiles/r2_issue_loop.md: the issuer showed no barrier-wait samples while the tensor pipe sat at 69 %).
        if (rank == 0) {      // pair mode: only the leader CTA issues MMAs (for both CTAs)
            // Descriptors are 64-bit words whose low 14 bits hold (address >> 4): every tap / k-step / tile /
            // plane variant is the base descriptor plus a small constant, so the issue loop is adds + MMAs only.
            constexpr uint32_t kALbo = C::HALO_PIX * 16, kASbo = C::HALO_W * 16;
            constexpr uint64_t kAPlane = C::A_PLANE_BYTES >> 4, kWPlane = W_PLANE >> 4;
            constexpr uint64_t kAKs = (2 * C::HALO_PIX * 16) >> 4, kWKs = (2 * (CAT ? 2 : 1) * W_ROWS * 16) >> 4;
            const uint64_t a_desc0 = make_desc(smem_u32(a_buf), kALbo, kASbo);
            const uint64_t w_desc0 = make_desc(smem_u32(w_buf), (CAT ? 2 : 1) * W_ROWS * 16, 128);
            uint32_t it = 0, ws = 0, gi = 0;
            if (WRES) {
                mbar_wait(smem_u32(&bars->w_full[0]), 0);
                tc_fence_after();
            }
            long long dbg_t[3] = {0, 0, 0}, dbg_c = 0;
            const long long dbg_start = p.dbg ? clock64() : 0;
            if constexpr (CAT && T == 1) {
                // Context-model layers ("other" mask): the schedule of a tile is the same for every tile -- group 0 = the 9
                // taps of filter depth 0, group 1 = 5 consecutive taps (0..4 forward, 4..8 data gradient) -- so it is spelled
                // out at compile time.  The table-driven loop below cost ~500 cycles PER TAP in this instantiation (tap table in
                // vector registers: LDC / IMAD / a dozen R2UR per tap; IC_TC_DBG=2: the issuer never waited, it was busy
                // 7.6 k cycles per tile for ~1 k tensor cycles of narrow MMAs): profiles/r2z_issuer_pc.txt.
                if (p.walk) {
                    // ---- depth walk.  Step j of a segment holds input slice j: it FINISHES output slice j - 1 (filter depth 1,
                    // taps 0..4) and STARTS output slice j (filter depth 0, taps 5..8 first, then 0..4).  Accumulators live in a
                    // ring of 4 tiles of [X | Y] = 2 NOUT columns; an output is X + Y (added in the epilogue) with
                    //     start   X += a_hi W0hi + a_lo W0hi     Y += a_hi W0lo
                    //     finish  X += a_hi W1lo                 Y += a_hi W1hi + a_lo W1hi
                    // so when the finishing tile sits right below the starting one, taps 0..4 are ONE a_hi MMA of N = 4 NOUT over
                    // [X_f | Y_f | X_s | Y_s] with B rows [W1lo | W1hi | W0hi | W0lo] and ONE a_lo MMA of N = 2 NOUT over the middle
                    // [Y_f | X_s] with rows [W1hi | W0hi]: 14 A fetches per plane and slice instead of 22.  Where the ring wraps
                    // (and at the ends of a segment) the two halves are issued separately on the same operands, in the same
                    // order per column -- every output sees the same sequence of accumulations whatever its ring position.
                    constexpr uint64_t kS1 = (uint64_t)(NPL * kWPlane);                  // single stage, 16-byte units
                    constexpr uint32_t IDESC_N4 = (C::IDESC & ~(0x3Fu << 17)) | ((uint32_t)((4 * NOUT) >> 3) << 17);
                    const uint64_t w2_desc0 = make_desc(smem_u32(w_buf), 4 * NOUT * 16, 128);                          // taps 0..4
                    const uint64_t w1_desc0 = make_desc(smem_u32(w_buf) + 10 * NPL * W_PLANE, 2 * NOUT * 16, 128);     // taps 5..8
                    auto issue_taps = [&](auto nt_c, auto tap0_c, auto cat2_c, uint64_t a_g, uint64_t w_g, uint32_t d_hi, uint32_t idesc_hi,
                                          uint32_t d_lo, uint64_t w_lo_off, uint32_t idesc_lo, bool first) {
                        constexpr int NT = decltype(nt_c)::value, TAP0 = decltype(tap0_c)::value;
                        constexpr bool CAT2 = decltype(cat2_c)::value;
                        constexpr uint64_t kStage = CAT2 ? 2 * kS1 : kS1;
                        constexpr uint64_t kWLboF = (uint64_t)(((CAT2 ? 4 : 2) * NOUT * 16) >> 4), kWKs2 = 2 * kWLboF;
                        constexpr uint64_t kALboF = (uint64_t)(kALbo >> 4);
#pragma unroll
                        for (int ti = 0; ti < NT; ++ti) {
                            const int tap = TAP0 + ti;
                            const uint64_t a_t = a_g + (uint64_t)((tap / 3) * C::HALO_W + (tap % 3));
                            const uint64_t w_t = w_g + (uint64_t)ti * kStage;
                            if (p.pair_c2 && (ti & 1)) {          // chunk 2 of taps ti-1 and ti as one k-step (see the generic loop)
                                const int tp = tap - 1;
                                const uint64_t a_prev = a_g + (uint64_t)((tp / 3) * C::HALO_W + (tp % 3));
                                const uint64_t a_pr = a_prev + kAKs - (kALboF << 16) + ((a_t - a_prev) << 16);
                                const uint64_t w_pr = (w_t - kStage) + kWKs2 - (kWLboF << 16) + (kStage << 16);
                                umma_f16(d_hi, a_pr, w_pr, idesc_hi, 1u);
                                umma_f16(d_lo, a_pr + kAPlane, w_pr + w_lo_off, idesc_lo, 1u);
                            }
                            umma_f16(d_hi, a_t, w_t, idesc_hi, (first && ti == 0) ? 0u : 1u);
                            umma_f16(d_lo, a_t + kAPlane, w_t + w_lo_off, idesc_lo, 1u);
                            if (!p.pair_c2 || ((ti & 1) == 0 && ti == NT - 1)) {
                                umma_f16(d_hi, a_t + kAKs, w_t + kWKs2, idesc_hi, 1u);
                                umma_f16(d_lo, a_t + kAKs + kAPlane, w_t + kWKs2 + w_lo_off, idesc_lo, 1u);
                            }
                        }
                    };
                    const std::integral_constant<int, 4> c4{};
                    const std::integral_constant<int, 5> c5{};
                    const std::integral_constant<int, 0> c0{};
                    const std::true_type cat2{};
                    const std::false_type cat1{};
                    uint32_t nacc = 0;          // accumulators started so far: ring position = nacc & 3, use count = nacc >> 2
                    for (int wi = cta_id; wi < n_walk; wi += cta_stride) {
                        const WalkItem w = walk_item(wi, tiles_y * tiles_x, p);
                        for (int j = w.da; j <= w.db; ++j, ++gi) {
                            const bool start = j < w.db, finish = j > w.da;
                            const uint32_t ss = nacc & 3, sf = (nacc - 1) & 3;
                            if (start) {
                                if (p.dbg) dbg_c = clock64();
                                mbar_wait(smem_u32(&bars->acc_empty[ss]), ((nacc >> 2) & 1) ^ 1);
                                if (p.dbg) dbg_t[0] += clock64() - dbg_c;
                            }
                            const uint32_t aslot = gi % ASLOTS;
                            if (p.dbg) dbg_c = clock64();
                            mbar_wait(smem_u32(&bars->a_full[aslot]), (gi / ASLOTS) & 1);
                            if (p.dbg) dbg_t[1] += clock64() - dbg_c;
                            tc_fence_after();
                            const uint64_t a_g = a_desc0 + (uint64_t)aslot * (NPL * kAPlane);
                            const uint32_t d_s = tmem_base + ss * (2 * NOUT), d_f = tmem_base + sf * (2 * NOUT);
                            if (elect_one()) {
                                if (start) issue_taps(c4, c5, cat1, a_g, w1_desc0, d_s, IDESC_CAT, d_s, 0ull, IDESC, true);
                                if (start && finish && ss != 0) {
                                    issue_taps(c5, c0, cat2, a_g, w2_desc0, d_f, IDESC_N4, d_f + NOUT, (uint64_t)NOUT, IDESC_CAT, false);
                                } else {
                                    if (start)
                                        issue_taps(c5, c0, cat2, a_g, w2_desc0 + (uint64_t)(2 * NOUT), d_s, IDESC_CAT, d_s, 0ull, IDESC, false);
                                    if (finish)
                                        issue_taps(c5, c0, cat2, a_g, w2_desc0, d_f, IDESC_CAT, d_f + NOUT, (uint64_t)NOUT, IDESC, false);
                                }
                                umma_commit(smem_u32(&bars->a_empty[aslot]));
                                if (finish) umma_commit(smem_u32(&bars->acc_full[sf]));
                            }
                            __syncwarp();
                            if (start) ++nacc;
                        }
                    }
                    n_work = 0;       // the loops below have nothing left to do
                } else if (p.pc_static) {
                    auto issue_group = [&](auto nt_c, auto tap0_c, uint64_t a_g, uint64_t w_g, uint32_t d_tmem, bool first_group) {
                        constexpr int NT = decltype(nt_c)::value, TAP0 = decltype(tap0_c)::value;
                        constexpr uint64_t kStage = (uint64_t)(NPL * kWPlane);
                        constexpr uint64_t kALboF = (uint64_t)(kALbo >> 4), kWLboF = (uint64_t)((2 * W_ROWS * 16) >> 4);
#pragma unroll
                        for (int ti = 0; ti < NT; ++ti) {
                            const int tap = TAP0 + ti;
                            const uint64_t a_t = a_g + (uint64_t)((tap / 3) * C::HALO_W + (tap % 3));
                            const uint64_t w_t = w_g + (uint64_t)ti * kStage;
                            if (p.pair_c2 && (ti & 1)) {          // chunk 2 of taps ti-1 and ti as one k-step (see below)
                                const int tp = tap - 1;
                                const uint64_t a_prev = a_g + (uint64_t)((tp / 3) * C::HALO_W + (tp % 3));
                                const uint64_t a_pr = a_prev + kAKs - (kALboF << 16) + ((a_t - a_prev) << 16);
                                const uint64_t w_pr = (w_t - kStage) + kWKs - (kWLboF << 16) + (kStage << 16);
                                umma_f16(d_tmem, a_pr, w_pr, IDESC_CAT, 1u);
                                umma_f16(d_tmem + NOUT, a_pr + kAPlane, w_pr, IDESC, 1u);
                            }
                            umma_f16(d_tmem, a_t, w_t, IDESC_CAT, (first_group && ti == 0) ? 0u : 1u);
                            umma_f16(d_tmem + NOUT, a_t + kAPlane, w_t, IDESC, 1u);
                            if (!p.pair_c2 || ((ti & 1) == 0 && ti == NT - 1)) {
                                umma_f16(d_tmem, a_t + kAKs, w_t + kWKs, IDESC_CAT, 1u);
                                umma_f16(d_tmem + NOUT, a_t + kAKs + kAPlane, w_t + kWKs, IDESC, 1u);
                            }
                        }
                    };
                    for (int wi = cta_id; wi < n_work; wi += cta_stride, ++it) {
                        const uint32_t set = it & 1;
                        if (p.dbg) dbg_c = clock64();
                        mbar_wait(smem_u32(&bars->acc_empty[set]), ((it >> 1) & 1) ^ 1);
                        if (p.dbg) dbg_t[0] += clock64() - dbg_c;
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + set * C::NCOL;
                        for (int g = 0; g < 2; ++g, ++gi) {
                            const uint32_t aslot = gi % ASLOTS;
                            if (p.dbg) dbg_c = clock64();
                            mbar_wait(smem_u32(&bars->a_full[aslot]), (gi / ASLOTS) & 1);
                            if (p.dbg) dbg_t[1] += clock64() - dbg_c;
                            tc_fence_after();
                            const uint64_t a_g = a_desc0 + (uint64_t)aslot * (NPL * kAPlane);
                            if (elect_one()) {
                                if (g == 0) {
                                    issue_group(std::integral_constant<int, 9>{}, std::integral_constant<int, 0>{}, a_g, w_desc0, d_tmem, true);
                                } else {
                                    const uint64_t w_g = w_desc0 + 9ull * (uint64_t)(NPL * kWPlane);
                                    if (p.pc_static == 1) issue_group(std::integral_constant<int, 5>{}, std::integral_constant<int, 0>{}, a_g, w_g, d_tmem, false);
                                    else issue_group(std::integral_constant<int, 5>{}, std::integral_constant<int, 4>{}, a_g, w_g, d_tmem, false);
                                }
                                umma_commit(smem_u32(&bars->a_empty[aslot]));
                                if (g == 1) umma_commit(smem_u32(&bars->acc_full[set]));
                            }
                            __syncwarp();
                        }
                    }
                    n_work = 0;       // the generic loop below has nothing left to do
                }
            }
            for (int wi = cta_id; wi < n_work; wi += cta_stride, ++it) {
                const uint32_t set = it & 1;
                if (p.dbg) dbg_c = clock64();
                mbar_wait(smem_u32(&bars->acc_empty[set]), ((it >> 1) & 1) ^ 1);
                if (p.dbg) dbg_t[0] += clock64() - dbg_c;
                tc_fence_after();
                uint32_t sidx = 0;      // stage index within the layer (resident mode)
                for (int g = 0; g < gt.ngroups; ++g, ++gi) {
                    const uint32_t aslot = gi % ASLOTS;
                    if (p.dbg) dbg_c = clock64();
                    mbar_wait(smem_u32(&bars->a_full[aslot]), (gi / ASLOTS) & 1);
                    if (p.dbg) dbg_t[1] += clock64() - dbg_c;
                    tc_fence_after();
                    const uint64_t a_g = a_desc0 + (uint64_t)aslot * (NPL * kAPlane);
                    const int nt = gt.ntaps[g];
                    for (int ti = 0; ti < nt; ++ti, ++ws, ++sidx) {
                        const int tap = gt.taps[g][ti];
                        const int dy = tap / 3, dx = tap - dy * 3;
                        uint32_t slot;
                        if (WRES) {
                            slot = sidx;
                        } else {
                            slot = ws % WSTAGES;
                            if (p.dbg) dbg_c = clock64();
                            mbar_wait(smem_u32(&bars->w_full[slot]), (ws / WSTAGES) & 1);
                            if (p.dbg) dbg_t[2] += clock64() - dbg_c;
                            tc_fence_after();
                        }
                        const uint64_t a_t = a_g + (uint64_t)(dy * C::HALO_W + dx);
                        const uint64_t w_t = w_desc0 + (uint64_t)slot * (NPL * kWPlane);
                        const uint32_t first0 = (g | ti) == 0 ? 0u : 1u;
                        // context model, <= 24 input channels (3 chunks): k-step 1 of a tap holds one real chunk and one
                        // all-zero chunk.  With pair_c2 the real chunks of taps 2j and 2j+1 share one k-step instead: A's two
                        // K core matrices are then (tap 2j, chunk 2) and (tap 2j+1, chunk 2), i.e. LBO = the tap shift, and
                        // B's are chunk 2 of the two taps' resident stages, LBO = the stage pitch -- descriptor fields only,
                        // no other layout.  22 instead of 28 k-steps per tile for these A-fetch-bound MMAs.
                        const bool pairing = CAT && p.pair_c2;
                        const bool do_ks1 = !pairing || ((ti & 1) == 0 && ti == nt - 1);      // unpaired last tap of an odd group
                        uint64_t a_pr = 0, w_pr = 0;
                        if (pairing && (ti & 1)) {
                            const int tp = gt.taps[g][ti - 1];
                            const uint64_t a_prev = a_g + (uint64_t)((tp / 3) * C::HALO_W + (tp % 3));
                            const uint64_t w_prev = w_t - (uint64_t)(NPL * kWPlane);
                            constexpr uint64_t kALboF = (uint64_t)(kALbo >> 4), kWLboF = (uint64_t)(((CAT ? 2 : 1) * W_ROWS * 16) >> 4);
                            a_pr = a_prev + kAKs - (kALboF << 16) + ((a_t - a_prev) << 16);
                            w_pr = w_prev + kWKs - (kWLboF << 16) + ((uint64_t)(NPL * kWPlane) << 16);
                        }
                        if (elect_one()) {
#pragma unroll
                            for (int t = 0; t < T; ++t) {
                                const uint32_t d_tmem = tmem_base + (set * T + t) * C::NCOL;
                                if (CAT && pairing && (ti & 1)) {
                                    umma_f16(d_tmem, a_pr + (uint64_t)(t * TW), w_pr, IDESC_CAT, 1u);
                                    umma_f16(d_tmem + NOUT, a_pr + (uint64_t)(t * TW) + kAPlane, w_pr, IDESC, 1u);
                                }
#pragma unroll
                                for (int ks = 0; ks < 2; ++ks) {
                                    if (ks == 1 && !do_ks1) continue;
                                    const uint64_t a_hi = a_t + (uint64_t)(t * TW) + ks * kAKs;
                                    const uint64_t w_hi = w_t + ks * kWKs;
                                    const uint32_t first = ks == 0 ? first0 : 1u;
                                    if (PAIR) {
                                        umma_f16_2sm(d_tmem, a_hi, w_hi, IDESC, first);
                                        if (NPL == 2) {
                                            umma_f16_2sm(d_tmem, a_hi, w_hi + kWPlane, IDESC, 1u);
                                            umma_f16_2sm(d_tmem, a_hi + kAPlane, w_hi, IDESC, 1u);
                                        }
                                    } else if (CAT) {
                                        umma_f16(d_tmem, a_hi, w_hi, IDESC_CAT, first);                // [hh | x]
                                        umma_f16(d_tmem + NOUT, a_hi + kAPlane, w_hi, IDESC, 1u);       // x += a_lo * w_hi
                                    } else {
                                        umma_f16(d_tmem, a_hi, w_hi, IDESC, first);
                                        if (NPL == 2) {
                                            umma_f16(d_tmem, a_hi, w_hi + kWPlane, IDESC, 1u);
                                            umma_f16(d_tmem, a_hi + kAPlane, w_hi, IDESC, 1u);
                                        }
                                    }
                                }
                            }
                            if (!WRES) {                          // stage free (in both CTAs) once these MMAs retire
                                if (PAIR) umma_commit_2sm(smem_u32(&bars->w_empty[slot]));
                                else umma_commit(smem_u32(&bars->w_empty[slot]));
                            }
                            if (ti == nt - 1) {
                                if (PAIR) umma_commit_2sm(smem_u32(&bars->a_empty[aslot]));
                                else umma_commit(smem_u32(&bars->a_empty[aslot]));
                                if (g == gt.ngroups - 1) {
                                    if (PAIR) umma_commit_2sm(smem_u32(&bars->acc_full[set]));
                                    else umma_commit(smem_u32(&bars->acc_full[set]));
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            if (p.dbg && lane == 0) {
                unsigned long long* d = p.dbg + (size_t)blockIdx.x * 4;
                d[0] = (unsigned long long)dbg_t[0];
                d[1] = (unsigned long long)dbg_t[1];
                d[2] = (unsigned long long)dbg_t[2];
                d[3] = (unsigned long long)(clock64() - dbg_start);
            }
        }
    } else {
        // ===================== epilogue (warps 3..6) =====================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;                  // accumulator row = pixel within the tile
        const int ty = m >> 3, tx = m & 7;
        constexpr int NCH = NOUT / 8;                 // output chunks (OUTMODE 0)
        constexpr int NCC = NOUT / 16 / (epi_warps(NOUT, OUTMODE) / 4);      // 16-column steps per warp
        const int cc0 = ((warp - 3) >> 2) * NCC;      // first step of this warp (8 epilogue warps: second half of the channels)
        const size_t plane = (size_t)p.N * NCH * p.H * p.W * 8;     // elements per hi/lo plane
        // one accumulator set (T tiles from TMEM column tcol) -> output pixels of super tile r of image n
        // (walk: the depth walk's [X | Y] accumulators, output = (X + Y) gain; otherwise [main | cross] = main gain + cross)
        // (resident-weight kernels have no second residual operand: its registers are compiled out)
        constexpr bool RES2 = !WRES;
        auto drain = [&](const int n, const int r, const uint32_t tcol, const bool walk) {
            const int y = (r / tiles_x) * TH + ty;
#pragma unroll 1
            for (int t = 0; t < T; ++t) {
                const int x = (r % tiles_x) * TW * T + t * TW + tx;
                const bool inside = y < p.H && x < p.W && n < p.N;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + tcol + t * C::NCOL;
                if (OUTMODE == 0) {
                    // chunk-0 offset of this pixel; optionally written in space-to-depth form for the next stride-2 conv
                    size_t pix_off = ((size_t)n * NCH * p.H + y) * p.W * 8 + (size_t)x * 8;
                    size_t chunk_stride = (size_t)p.H * p.W * 8;
                    if (p.out_s2d) {
                        const int ph = (y & 1) * 2 + (x & 1);
                        chunk_stride = (size_t)(p.H >> 1) * (p.W >> 1) * 8;
                        pix_off = (((size_t)n * 4 * NCH + ph * NCH) * (p.H >> 1) + (y >> 1)) * (p.W >> 1) * 8 + (size_t)(x >> 1) * 8;
                    }
                    const bool has_res = (p.res1 != nullptr) || (RES2 && p.res2 != nullptr);
                    // res1 may live in a larger tensor (context model: conv0 output cropped [2:, 2:-2, 2:-2])
                    const int rimg = n + (p.img_div > 0 ? (n / p.img_div) * p.res_div_mul : 0) + p.res_img_off;
                    const size_t rstride = (size_t)p.res_H * p.res_W * 8;
                    const size_t roff = ((size_t)rimg * NCH * p.res_H + y + p.res_dy) * p.res_W * 8 + (size_t)(x + p.res_dx) * 8;
                    ResRegs cur, nxt;
                    const size_t r2off = ((size_t)n * NCH * p.H + y) * p.W * 8 + (size_t)x * 8;     // res2: output geometry
                    const size_t r2stride = (size_t)p.H * p.W * 8;
                    if (inside && has_res)
                        load_res<NPL, RES2>(cur, p, roff + (size_t)cc0 * 2 * rstride, rstride, p.res_plane, r2off + (size_t)cc0 * 2 * r2stride,
                                            r2stride, plane);
#pragma unroll 1
                    for (int cc = cc0; cc < cc0 + NCC; ++cc) {
                        uint32_t rr[16];
                        tmem_ld16(taddr + cc * 16, rr);
                        if (CAT) {          // main sum (gain: see launch_t) + cross sum
                            uint32_t rx[16];
                            tmem_ld16(taddr + NOUT + cc * 16, rx);
                            tmem_ld_wait();
#pragma unroll
                            for (int e = 0; e < 16; ++e)
                                rr[e] = __float_as_uint(walk ? (__uint_as_float(rr[e]) + __uint_as_float(rx[e])) * p.acc_gain
                                                             : fmaf(__uint_as_float(rr[e]), p.acc_gain, __uint_as_float(rx[e])));
                        }
                        if (inside && has_res && cc + 1 < cc0 + NCC)
                            load_res<NPL, RES2>(nxt, p, roff + (size_t)(cc + 1) * 2 * rstride, rstride, p.res_plane,
                                                r2off + (size_t)(cc + 1) * 2 * r2stride, r2stride, plane);
                        tmem_ld_wait();
                        if (inside) {
#pragma unroll
                            for (int hc = 0; hc < 2; ++hc) {
                                const int chunk = cc * 2 + hc;
                                if (p.store_chunks && chunk >= p.store_chunks) continue;      // all-zero padding channels
                                // depth-to-space output (transposed convs): column block = (output phase, channel chunk)
                                const int sch = p.d2s_cch ? chunk % p.d2s_cch : chunk;
                                float v[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    float a = __uint_as_float(rr[hc * 8 + e]);
                                    a = fmaf(a, s_scale[sch * 8 + e], s_shift[sch * 8 + e]);
                                    v[e] = p.relu ? fmaxf(a, 0.f) : a;
                                }
                                if (p.res1) {
                                    if (NPL == 2) add_h8_pair(cur.v[hc][0], cur.v[hc][1], v);
                                    else add_h8(cur.v[hc][0], v);
                                }
                                if (RES2 && p.res2) {
                                    if (NPL == 2) add_h8_pair(cur.v[hc][2], cur.v[hc][3], v);
                                    else add_h8(cur.v[hc][2], v);
                                }
                                __align__(16) __half2 hi[4];
                                __align__(16) __half2 lo[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    hi[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                                    float2 hf = __half22float2(hi[e]);
                                    lo[e] = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
                                }
                                size_t off = pix_off + (size_t)chunk * chunk_stride, oplane = plane;
                                if (p.d2s_cch) {
                                    const int phs = p.d2s_ph0 + chunk / p.d2s_cch;
                                    const int yo = 2 * y + (phs >> 1), xo = 2 * x + (phs & 1);
                                    off = ((((size_t)n * p.d2s_cch + sch) * (2 * p.H) + yo) * (2 * p.W) + xo) * 8;
                                    oplane = (size_t)p.N * p.d2s_cch * (2 * p.H) * (2 * p.W) * 8;
                                }
                                *reinterpret_cast<float4*>(p.out + off) = *reinterpret_cast<const float4*>(hi);
                                if (NPL == 2)
                                    *reinterpret_cast<float4*>(p.out + oplane + off) = *reinterpret_cast<const float4*>(lo);
                            }
                        }
                        cur = nxt;
                    }
                } else if (OUTMODE == 2) {
                    uint32_t rr[16];
                    tmem_ld16(taddr, rr);
                    if (CAT) {
                        uint32_t rx[16];
                        tmem_ld16(taddr + NOUT, rx);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            rr[e] = __float_as_uint(walk ? (__uint_as_float(rr[e]) + __uint_as_float(rx[e])) * p.acc_gain
                                                         : fmaf(__uint_as_float(rr[e]), p.acc_gain, __uint_as_float(rx[e])));
                    }
                    tmem_ld_wait();
                    const int L = p.cout;
                    const size_t v = ((size_t)n * p.H + y) * p.W + x;         // (N, D, h, w) linear index
                    float bits = 0.f;
                    if (inside) {
                        float lg[8], m = 0.f;
#pragma unroll
                        for (int l = 0; l < 8; ++l)
                            if (l < L) {    // bias + ReLU: the last conv3d keeps the default activation (probclass.py:220,233)
                                lg[l] = fmaxf(fmaf(__uint_as_float(rr[l]), s_scale[l], s_shift[l]), 0.f);
                                m = (l == 0) ? lg[l] : fmaxf(m, lg[l]);
                            }
                        if (p.head == 0) {
#pragma unroll
                            for (int l = 0; l < 8; ++l)
                                if (l < L) p.out_f32[v * L + l] = lg[l];
                        } else {
                            float e[8], ssum = 0.f;
#pragma unroll
                            for (int l = 0; l < 8; ++l)
                                if (l < L) {
                                    e[l] = expf(__fsub_rn(lg[l], m));
                                    ssum = __fadd_rn(ssum, e[l]);
                                }
                            const int sym = p.symbols ? (int)p.symbols[v] : 0;
                            float lsel = 0.f;
#pragma unroll
                            for (int l = 0; l < 8; ++l)
                                if (l == sym) lsel = lg[l];
                            bits = __fmul_rn(__fsub_rn(logf(ssum), __fsub_rn(lsel, m)), 1.44269502f);
                            if (p.head == 1) {
                                p.out_f32[v] = bits;
                            } else {
#pragma unroll
                                for (int l = 0; l < 8; ++l)
                                    if (l < L) {
                                        long long f = (long long)__fmul_rn(__fdiv_rn(e[l], ssum), 1e9f);
                                        p.out_freqs[v * L + l] = f < 1 ? 1 : f;
                                    }
                            }
                        }
                    }
                    if (p.head != 0 && p.bits_sum) {     // per-image sum: the whole warp belongs to one (n, d) slice
                        double b = inside && p.symbols ? (double)bits : 0.0;
                        for (int o = 16; o > 0; o >>= 1) b += __shfl_down_sync(0xffffffffu, b, o);
                        if (lane == 0 && b != 0.0) atomicAdd(p.bits_sum + (p.img_div > 0 ? n / p.img_div : n), b);
                    }
                } else if (OUTMODE == 3) {
                    // h13: 4 output phases x 3 channels -> BN, _denormalize, clip (code/autoencoder.py:146-158), NCHW image
                    uint32_t rr[16];
                    tmem_ld16(taddr, rr);
                    tmem_ld_wait();
                    if (inside) {
                        const size_t H2 = 2 * (size_t)p.H, W2 = 2 * (size_t)p.W;
#pragma unroll
                        for (int phs = 0; phs < 4; ++phs)
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                float v = fmaf(__uint_as_float(rr[phs * 3 + c]), s_scale[c], s_shift[c]);
                                if (p.denorm) {
                                    v = __fadd_rn(__fmul_rn(v, c_img_std[c]), c_img_mean[c]);
                                    v = fminf(fmaxf(v, 0.f), 255.f);
                                }
                                const size_t o = (((size_t)n * 3 + c) * H2 + 2 * y + (phs >> 1)) * W2 + 2 * x + (phs & 1);
                                p.out_f32[o] = v;
                                if (p.out_u8) p.out_u8[o] = (uint8_t)v;     // tf.cast truncation (val.py:91)
                            }
                    }
                } else {
                    float* o = p.out_f32 + (((size_t)n * p.H + y) * p.W + x) * p.cout;
#pragma unroll 1
                    for (int cc = 0; cc < NOUT / 16; ++cc) {
                        uint32_t rr[16];
                        tmem_ld16(taddr + cc * 16, rr);
                        if (CAT) {          // main sum (gain: see launch_t) + cross sum
                            uint32_t rx[16];
                            tmem_ld16(taddr + NOUT + cc * 16, rx);
                            tmem_ld_wait();
#pragma unroll
                            for (int e = 0; e < 16; ++e)
                                rr[e] = __float_as_uint(walk ? (__uint_as_float(rr[e]) + __uint_as_float(rx[e])) * p.acc_gain
                                                             : fmaf(__uint_as_float(rr[e]), p.acc_gain, __uint_as_float(rx[e])));
                        }
                        tmem_ld_wait();
                        if (inside) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const int c = cc * 16 + e;
                                if (c < p.cout) {
                                    float a = fmaf(__uint_as_float(rr[e]), s_scale[c], s_shift[c]);
                                    o[c] = p.relu ? fmaxf(a, 0.f) : a;
                                }
                            }
                        }
                    }
                }
            }
        };
        if constexpr (CAT && T == 1) {
            if (p.walk) {       // depth walk: outputs arrive in the order the issuer starts them, ring of 4 accumulators
                const int tiles = tiles_y * tiles_x;
                uint32_t nacc = 0;
                int wi = cta_id, d = 0;
                bool more = wi < n_walk, ins = false;
                WalkItem w;
                // element offsets of this thread's pixel in the output / residual tensors (first chunk of this warp), advanced by
                // one (n, depth) image per output: the divisions run once per work item, not once per accumulator
                const size_t chunk_stride = (size_t)p.H * p.W * 8, rchunk = (size_t)p.res_H * p.res_W * 8;
                const int wc0 = (OUTMODE == 0 && NCC == 1) ? 2 * cc0 : 0;
                size_t ooff = 0, roff = 0;
                auto enter = [&]() {        // first output of work item wi
                    w = walk_item(wi, tiles, p);
                    d = w.da;
                    const int y = (w.r / tiles_x) * TH + ty, x = (w.r % tiles_x) * TW + tx;
                    ins = y < p.H && x < p.W;
                    ooff = ((size_t)(w.img * p.img_div + d) * NCH * p.H + y) * p.W * 8 + (size_t)x * 8 + (size_t)wc0 * chunk_stride;
                    const int rimg = w.img * (p.img_div + p.res_div_mul) + d + p.res_img_off;
                    roff = ((size_t)rimg * NCH * p.res_H + y + p.res_dy) * p.res_W * 8 + (size_t)(x + p.res_dx) * 8 + (size_t)wc0 * rchunk;
                };
                auto advance = [&]() {      // -> the output after the current one
                    if (++d < w.db) {
                        ooff += NCH * chunk_stride;
                        roff += NCH * rchunk;
                    } else {
                        wi += cta_stride;
                        more = wi < n_walk;
                        if (more) enter();
                    }
                };
                if (more) enter();
                if constexpr (OUTMODE == 0 && NCC == 1) {
                    // Plane-output layers: a lean drain.  ncu (profiles/r2u3_pc_walk_raw.csv): these kernels run at the issue rate
                    // of their epilogue warps -- 4.7 k (layer 1) / 6.6 k (layer 2) warp instructions per accumulator through the
                    // generic drain, 0.46 IPC per scheduler, time proportional to the instruction count -- not at the tensor pipe's.
                    // Here: this warp's two channel chunks one at a time (8 + 8 TMEM columns live instead of 32), chunks past
                    // store_chunks never read, incremental addresses, and the residual of the NEXT output requested into the
                    // registers the current one has just consumed.  Same operations per value as the generic drain.
                    const int nch = p.store_chunks ? max(0, min(2, p.store_chunks - 2 * cc0)) : 2;
                    const bool has_res = p.res1 != nullptr;
                    const float gain = p.acc_gain;
                    float4 rh[2], rl[2];
                    if (more && has_res && ins) {
                        const __half* rp = p.res1 + roff;
#pragma unroll
                        for (int hc = 0; hc < 2; ++hc)
                            if (hc < nch) {
                                rh[hc] = ldg16(rp + hc * rchunk);
                                rl[hc] = ldg16(rp + p.res_plane + hc * rchunk);
                            }
                    }
                    while (more) {
                        const bool inside = ins;
                        __half* op = p.out + ooff;
                        advance();
                        const __half* rnext = (more && has_res && ins) ? p.res1 + roff : nullptr;
                        const uint32_t set = nacc & 3;
                        mbar_wait(smem_u32(&bars->acc_full[set]), (nacc >> 2) & 1);
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + set * (2 * NOUT) + cc0 * 16;
#pragma unroll
                        for (int hc = 0; hc < 2; ++hc) {
                            if (hc < nch) {
                                uint32_t ax[8], ay[8];
                                tmem_ld8(taddr + hc * 8, ax);
                                tmem_ld8(taddr + NOUT + hc * 8, ay);
                                tmem_ld_wait();
                                if (inside) {
                                    const int c0 = (2 * cc0 + hc) * 8;
                                    const float4 s0 = *reinterpret_cast<const float4*>(s_scale + c0), s1 = *reinterpret_cast<const float4*>(s_scale + c0 + 4);
                                    const float4 h0 = *reinterpret_cast<const float4*>(s_shift + c0), h1 = *reinterpret_cast<const float4*>(s_shift + c0 + 4);
                                    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                                    const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                                    float v[8];
#pragma unroll
                                    for (int e = 0; e < 8; ++e) {
                                        float a = (__uint_as_float(ax[e]) + __uint_as_float(ay[e])) * gain;
                                        a = fmaf(a, sc[e], sh[e]);
                                        v[e] = p.relu ? fmaxf(a, 0.f) : a;
                                    }
                                    if (has_res) add_h8_pair(rh[hc], rl[hc], v);
                                    __align__(16) __half2 hi[4];
                                    __align__(16) __half2 lo[4];
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        hi[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                                        const float2 hf = __half22float2(hi[e]);
                                        lo[e] = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
                                    }
                                    *reinterpret_cast<float4*>(op + hc * chunk_stride) = *reinterpret_cast<const float4*>(hi);
                                    *reinterpret_cast<float4*>(op + plane + hc * chunk_stride) = *reinterpret_cast<const float4*>(lo);
                                }
                                if (rnext) {
                                    rh[hc] = ldg16(rnext + hc * rchunk);
                                    rl[hc] = ldg16(rnext + p.res_plane + hc * rchunk);
                                }
                            }
                        }
                        tc_fence_before();
                        mbar_arrive(smem_u32(&bars->acc_empty[set]));
                        ++nacc;
                    }
                } else {
                    while (more) {
                        const int n = w.img * p.img_div + d, r = w.r;
                        advance();
                        const uint32_t set = nacc & 3;
                        mbar_wait(smem_u32(&bars->acc_full[set]), (nacc >> 2) & 1);
                        tc_fence_after();
                        drain(n, r, set * (2 * NOUT), true);
                        tc_fence_before();
                        mbar_arrive(smem_u32(&bars->acc_empty[set]));
                        ++nacc;
                    }
                }
                n_work = 0;
            }
        }
        uint32_t it = 0;
        for (int wi = cta_id; wi < n_work; wi += cta_stride, ++it) {
            const uint32_t set = it & 1;
            const int st = PAIR ? 2 * wi + (int)rank : wi;
            const int n = st / (tiles_y * tiles_x), r = st - n * tiles_y * tiles_x;
            mbar_wait(smem_u32(&bars->acc_full[set]), (it >> 1) & 1);
            tc_fence_after();
            drain(n, r, set * T * C::NCOL, false);
            tc_fence_before();
            if (PAIR) mbar_arrive_cluster(mapa_rank(smem_u32(&bars->acc_empty[set]), 0));    // the leader waits for both epilogues
            else mbar_arrive(smem_u32(&bars->acc_empty[set]));
        }
    }
    // teardown
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();       // neither CTA may exit (or free TMEM) while the peer can still touch it
    if (warp == 2) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
    }
}

// ------------------------------------------------------------- B-concatenated kernel (EXACT, 128 output channels)
// The residual 3x3 convs and h2 in EXACT mode.  An SS-mode UMMA fetches its operands from shared memory at ~64 B/clk
// per SM (tools/ubench/mma_shapes.cu), so M128 x N128 x K16 (4 KB of A + 4 KB of B per 64 tensor cycles) runs at half
// rate and what counts is operand BYTES per product.  Per k-step (16 input channels of one tap):
//     conv_tc_kernel      3 x N128:  a_hi*w_hi, a_hi*w_lo, a_lo*w_hi                       24 KB  (one accumulator)
//     here, one CTA       N256 + N128: a_hi*[w_hi | w_lo] -> [hh | x],  a_lo*w_hi -> x     20 KB
//     here, CTA pair      the same with cta_group::2 (M = 256 over two CTAs, each CTA supplies half of B's rows)  14 KB / CTA
// The weights of a stage are stored [4 chunks][hi rows 0..127 | lo rows 128..255][8 cin] so that the N = 256 descriptor
// walks both planes and the N = 128 descriptor the hi rows alone.  hi*hi and the two cross terms land in SEPARATE fp32
// accumulators (TMEM columns [hh | x] of a tile) and are added in the fp32 epilogue: the tensor core's truncating
// accumulate then acts 72 times on the main sum instead of 216 (and on the 2^-11 times smaller cross sum, where it does
// not matter), a third of conv_tc_kernel's truncation error.
// A tile needs 256 TMEM columns, a 16 x 16 super tile (2 tiles sharing every weight stage) all 512: there is no second
// accumulator set to hide the epilogue behind.  Instead the two tiles are SKEWED: tile 0 runs SK stages ahead of tile 1
//     phase A   tile 0: stages 0 .. SK-1                    (tile 1's accumulators of the previous item are drained)
//     phase B   tile 0: stage s, tile 1: stage s - SK       (s = SK .. S-1)
//     phase C   tile 1: stages S-SK .. S-1                  (tile 0's accumulators are drained)
// so the tensor pipe always has SK stages (~1.5k cycles) of the other tile's MMAs while 8 epilogue warps (two per TMEM
// lane quarter, 64 output channels each) drain one tile.  Every tile still sums its K dimension in the same order.
// Pair mode: the leader CTA issues for both, every barrier it waits on is signalled by both CTAs (as in conv_tc_kernel);
// CTA r holds, per stage, block P = 128 rows (r = 0: w_hi, r = 1: w_lo: the N = 256 operand) and block Q = w_hi rows
// 64r .. 64r+63 (the N = 128 operand), 12 KB, loaded as one TMA box of the pair-packed tensor.
constexpr int CAT_SK = 4;                     // stages tile 0 runs ahead of tile 1
constexpr int CAT_WSTAGES = 8;
constexpr int CAT_WSTAGES_PAIR = 10;          // 12 KB stages
constexpr int CAT_EPI_WARPS = 8;
constexpr int CAT_NTHREADS = (3 + CAT_EPI_WARPS) * 32;
constexpr int CAT_MAX_STAGES = 80;            // stages per super tile (3x3: 36, h2: 50)

// per stage of a super tile, flattened over (group, tap): kernel parameter, so the issuer reads it with uniform loads
struct CatStageTable {
    uint16_t aoff[CAT_MAX_STAGES];      // tap offset in the halo tile (16-byte units)
    uint8_t flag[CAT_MAX_STAGES];       // bit 0: first stage of its group, bit 1: last
};

struct __align__(8) CatBarriers {
    uint64_t a_full[2], a_empty[2], w_full[CAT_WSTAGES_PAIR], w_empty[CAT_WSTAGES_PAIR], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

template <bool PAIR>
__global__ void __launch_bounds__(CAT_NTHREADS, 1)
conv_cat_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap w_map, const ConvTcParams p,
                const GroupTable gt, const __grid_constant__ CatStageTable tab) {
    using C = Cfg<2, 128, 4>;
    constexpr int NPL = 2, T = 2;
    constexpr int WSTAGES = PAIR ? CAT_WSTAGES_PAIR : CAT_WSTAGES;
    constexpr int W_STAGE_BYTES = PAIR ? 12288 : 16384;            // per CTA
    constexpr uint32_t IDESC256 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
    constexpr uint32_t IDESC128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    const int cta_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int cta_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_buf = smem;                                              // [2 slots][2 planes][A_PLANE_BYTES]
    uint8_t* w_buf = a_buf + 2 * NPL * C::A_PLANE_BYTES;                // [WSTAGES][W_STAGE_BYTES]
    float* s_scale = reinterpret_cast<float*>(w_buf + WSTAGES * W_STAGE_BYTES);
    float* s_shift = s_scale + 128;
    CatBarriers* bars = reinterpret_cast<CatBarriers*>(s_shift + 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (p.W + TW * T - 1) / (TW * T), tiles_y = (p.H + TH - 1) / TH;
    const int n_super = p.N * tiles_y * tiles_x;
    const int n_work = PAIR ? (n_super + 1) / 2 : n_super;
    const int S = gt.nstages;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars->a_full[i]), 1);          // pair: the leader's producer announces both CTAs' bytes
            mbar_init(smem_u32(&bars->a_empty[i]), 1);
            mbar_init(smem_u32(&bars->acc_full[i]), 1);
            mbar_init(smem_u32(&bars->acc_empty[i]), (PAIR ? 2 : 1) * CAT_EPI_WARPS * 32);
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
    if (warp == 2) {   // TMEM: 2 tiles x [hh | x] x 128 fp32 columns
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================== activation producer =====================
        if (lane == 0) {
            uint32_t gi = 0;
            for (int wi = cta_id; wi < n_work; wi += cta_stride) {
                const int st = PAIR ? 2 * wi + (int)rank : wi;      // past-the-end super tiles load zeros (TMA OOB)
                const int n = st / (tiles_y * tiles_x), r = st - n * tiles_y * tiles_x;
                const int y0 = (r / tiles_x) * TH, x0 = (r % tiles_x) * TW * T;
                for (int g = 0; g < gt.ngroups; ++g, ++gi) {
                    const uint32_t slot = gi & 1, ph = (gi >> 1) & 1;
                    mbar_wait(smem_u32(&bars->a_empty[slot]), ph ^ 1);
                    if (PAIR) {
                        const uint32_t full = mapa_rank(smem_u32(&bars->a_full[slot]), 0);
                        if (rank == 0) mbar_expect_tx(smem_u32(&bars->a_full[slot]), 2 * NPL * C::A_PLANE_BYTES);
                        for (int pl = 0; pl < NPL; ++pl)
                            tma_load_5d_2sm(smem_u32(a_buf + (slot * NPL + pl) * C::A_PLANE_BYTES), &in_map, full,
                                            (x0 + p.halo0) * 8, y0 + p.halo0, gt.chunk0[g], n, pl);
                    } else {
                        const uint32_t full = smem_u32(&bars->a_full[slot]);
                        mbar_expect_tx(full, NPL * C::A_PLANE_BYTES);
                        for (int pl = 0; pl < NPL; ++pl)
                            tma_load_5d(smem_u32(a_buf + (slot * NPL + pl) * C::A_PLANE_BYTES), &in_map, full,
                                        (x0 + p.halo0) * 8, y0 + p.halo0, gt.chunk0[g], n, pl);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== weight producer =====================
        if (lane == 0) {
            uint32_t ws = 0;
            for (int wi = cta_id; wi < n_work; wi += cta_stride) {
                for (int s = 0; s < S; ++s, ++ws) {
                    const uint32_t slot = ws % WSTAGES, ph = (ws / WSTAGES) & 1;
                    mbar_wait(smem_u32(&bars->w_empty[slot]), ph ^ 1);
                    if (PAIR) {
                        const uint32_t full = mapa_rank(smem_u32(&bars->w_full[slot]), 0);
                        if (rank == 0) mbar_expect_tx(smem_u32(&bars->w_full[slot]), 2 * W_STAGE_BYTES);
                        tma_load_3d_2sm(smem_u32(w_buf + slot * W_STAGE_BYTES), &w_map, full, 0, 0, s * 2 + (int)rank);
                    } else {
                        const uint32_t full = smem_u32(&bars->w_full[slot]);
                        mbar_expect_tx(full, W_STAGE_BYTES);
                        bulk_load(smem_u32(w_buf + slot * W_STAGE_BYTES), p.weights + (size_t)s * W_STAGE_BYTES, W_STAGE_BYTES, full);
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ===================== MMA issuer (whole warp walks the schedule, one elected lane issues) =====================
        if (rank == 0) {
            constexpr uint32_t kALbo = C::HALO_PIX * 16, kASbo = C::HALO_W * 16;
            constexpr uint64_t kAPlane = C::A_PLANE_BYTES >> 4;
            constexpr uint64_t kAKs = (2 * C::HALO_PIX * 16) >> 4;
            // single CTA: one block [4 chunks][256 rows][8]; pair: P = [4][128][8] then Q = [4][64][8]
            constexpr uint32_t kPRows = PAIR ? 128 : 256, kQRows = PAIR ? 64 : 256;
            constexpr uint64_t kPKs = (2 * kPRows * 16) >> 4, kQKs = (2 * kQRows * 16) >> 4;
            constexpr uint64_t kWStage = W_STAGE_BYTES >> 4;
            const uint64_t a_desc0 = make_desc(smem_u32(a_buf), kALbo, kASbo);
            const uint64_t p_desc0 = make_desc(smem_u32(w_buf), kPRows * 16, 128);
            const uint64_t q_desc0 = make_desc(smem_u32(w_buf) + (PAIR ? 8192 : 0), kQRows * 16, 128);
            uint32_t J0 = 0, J1 = 0, G0 = 0, G1 = 0;       // running stage / group counters of tile 0 and tile 1
            // the MMAs of one stage of one tile (+ the commits that follow them), by the elected lane
            auto mma_stage = [&](int t, int s, uint32_t aslot, uint32_t wslot, bool rel_w, bool rel_a, bool acc_done) {
                const uint64_t a_t = a_desc0 + (uint64_t)aslot * (NPL * kAPlane) + (uint64_t)tab.aoff[s] + (uint64_t)(t * TW);
                const uint64_t wp = p_desc0 + (uint64_t)wslot * kWStage, wq = q_desc0 + (uint64_t)wslot * kWStage;
                const uint32_t d = tmem_base + t * 256;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint64_t a_hi = a_t + ks * kAKs;
                        const uint32_t acc = (s | ks) == 0 ? 0u : 1u;
                        if (PAIR) {
                            umma_f16_2sm(d, a_hi, wp + ks * kPKs, IDESC256, acc);
                            umma_f16_2sm(d + 128, a_hi + kAPlane, wq + ks * kQKs, IDESC128, 1u);
                        } else {
                            umma_f16(d, a_hi, wp + ks * kPKs, IDESC256, acc);
                            umma_f16(d + 128, a_hi + kAPlane, wq + ks * kQKs, IDESC128, 1u);
                        }
                    }
                    if (PAIR) {
                        if (rel_w) umma_commit_2sm(smem_u32(&bars->w_empty[wslot]));     // both tiles are done with the stage
                        if (rel_a) umma_commit_2sm(smem_u32(&bars->a_empty[aslot]));
                        if (acc_done) umma_commit_2sm(smem_u32(&bars->acc_full[t]));
                    } else {
                        if (rel_w) umma_commit(smem_u32(&bars->w_empty[wslot]));
                        if (rel_a) umma_commit(smem_u32(&bars->a_empty[aslot]));
                        if (acc_done) umma_commit(smem_u32(&bars->acc_full[t]));
                    }
                }
                __syncwarp();
            };
            auto issue0 = [&](int s) {
                const uint32_t fl = tab.flag[s];
                if (fl & 1) {
                    mbar_wait(smem_u32(&bars->a_full[G0 & 1]), (G0 >> 1) & 1);
                    tc_fence_after();
                }
                const uint32_t wslot = J0 % WSTAGES;
                mbar_wait(smem_u32(&bars->w_full[wslot]), (J0 / WSTAGES) & 1);
                tc_fence_after();
                mma_stage(0, s, G0 & 1, wslot, false, false, s == S - 1);
                if (fl & 2) ++G0;
                ++J0;
            };
            auto issue1 = [&](int s) {
                const uint32_t fl = tab.flag[s];
                mma_stage(1, s, G1 & 1, J1 % WSTAGES, true, (fl & 2) != 0, s == S - 1);
                if (fl & 2) ++G1;
                ++J1;
            };
            uint32_t it = 0;
            for (int wi = cta_id; wi < n_work; wi += cta_stride, ++it) {
                const uint32_t par = (it & 1) ^ 1;
                mbar_wait(smem_u32(&bars->acc_empty[0]), par);      // tile 0 of the previous item has been drained
                tc_fence_after();
                for (int s = 0; s < CAT_SK; ++s) issue0(s);
                mbar_wait(smem_u32(&bars->acc_empty[1]), par);
                tc_fence_after();
                for (int s = CAT_SK; s < S; ++s) {
                    issue0(s);
                    issue1(s - CAT_SK);
                }
                for (int s = S - CAT_SK; s < S; ++s) issue1(s);
            }
        }
    } else {
        // ===================== epilogue (warps 3..10): two warps per TMEM lane quarter, 64 output channels each ====
        const int q = warp & 3;
        const int half = (warp - 3) >> 2;
        const int m = q * 32 + lane;
        const int ty = m >> 3, tx = m & 7;
        constexpr int NCH = 16;
        const size_t plane = (size_t)p.N * NCH * p.H * p.W * 8;
        const float gain = p.acc_gain;
        uint32_t it = 0;
        for (int wi = cta_id; wi < n_work; wi += cta_stride, ++it) {
            const int st = PAIR ? 2 * wi + (int)rank : wi;
            const int n = st / (tiles_y * tiles_x), r = st - n * tiles_y * tiles_x;
            const int y = (r / tiles_x) * TH + ty;
#pragma unroll 1
            for (int t = 0; t < T; ++t) {
                mbar_wait(smem_u32(&bars->acc_full[t]), it & 1);
                tc_fence_after();
                const int x = (r % tiles_x) * TW * T + t * TW + tx;
                const bool inside = y < p.H && x < p.W && n < p.N;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + t * 256 + half * 64;
                size_t pix_off = ((size_t)n * NCH * p.H + y) * p.W * 8 + (size_t)x * 8;
                size_t chunk_stride = (size_t)p.H * p.W * 8;
                if (p.out_s2d) {
                    const int ph = (y & 1) * 2 + (x & 1);
                    chunk_stride = (size_t)(p.H >> 1) * (p.W >> 1) * 8;
                    pix_off = (((size_t)n * 4 * NCH + ph * NCH) * (p.H >> 1) + (y >> 1)) * (p.W >> 1) * 8 + (size_t)(x >> 1) * 8;
                }
                const bool has_res = (p.res1 != nullptr) || (p.res2 != nullptr);
                const size_t rstride = (size_t)p.H * p.W * 8;
                const size_t roff = ((size_t)n * NCH * p.H + y) * p.W * 8 + (size_t)x * 8 + (size_t)(half * 8) * rstride;
                ResRegs cur, nxt;
                if (inside && has_res) load_res<NPL>(cur, p, roff, rstride, plane, roff, rstride, plane);
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    uint32_t rh[16], rx[16];
                    tmem_ld16(taddr + cc * 16, rh);
                    tmem_ld16(taddr + 128 + cc * 16, rx);
                    if (inside && has_res && cc + 1 < 4)
                        load_res<NPL>(nxt, p, roff + (size_t)(cc + 1) * 2 * rstride, rstride, plane,
                                      roff + (size_t)(cc + 1) * 2 * rstride, rstride, plane);
                    tmem_ld_wait();
                    if (inside) {
#pragma unroll
                        for (int hc = 0; hc < 2; ++hc) {
                            const int chunk = half * 8 + cc * 2 + hc;
                            float v[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                // main sum (gain undoes the mean of its truncation, see launch_cat) + cross sum, then BN
                                float a = fmaf(__uint_as_float(rh[hc * 8 + e]), gain, __uint_as_float(rx[hc * 8 + e]));
                                a = fmaf(a, s_scale[chunk * 8 + e], s_shift[chunk * 8 + e]);
                                v[e] = p.relu ? fmaxf(a, 0.f) : a;
                            }
                            if (p.res1) add_h8_pair(cur.v[hc][0], cur.v[hc][1], v);
                            if (p.res2) add_h8_pair(cur.v[hc][2], cur.v[hc][3], v);
                            __align__(16) __half2 hi[4];
                            __align__(16) __half2 lo[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                hi[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                                float2 hf = __half22float2(hi[e]);
                                lo[e] = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
                            }
                            const size_t off = pix_off + (size_t)chunk * chunk_stride;
                            *reinterpret_cast<float4*>(p.out + off) = *reinterpret_cast<const float4*>(hi);
                            *reinterpret_cast<float4*>(p.out + plane + off) = *reinterpret_cast<const float4*>(lo);
                        }
                    }
                    cur = nxt;
                }
                tc_fence_before();
                if (PAIR) mbar_arrive_cluster(mapa_rank(smem_u32(&bars->acc_empty[t]), 0));
                else mbar_arrive(smem_u32(&bars->acc_empty[t]));
            }
        }
    }
    // teardown
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    if (warp == 2) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// ------------------------------------------------------------- layout changes
// All activations of the tensor-core path live as fp16 hi/lo planes [plane][N][CH][H][W][8]
// (CH = channels/8).  "s2d" = space-to-depth by 2: [plane][N][4*CH][H/2][W/2][8] with
// chunk' = ((y&1)*2 + (x&1))*CH + chunk, which turns a stride-2 5x5 conv into a stride-1 conv
// with <= 3x3 taps per phase (the tap -> (phase, offset) table is built on the host).
__device__ __forceinline__ void split8(const float (&v)[8], float4& hi4, float4& lo4) {
    __half2* hi = reinterpret_cast<__half2*>(&hi4);
    __half2* lo = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        hi[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        float2 hf = __half22float2(hi[e]);
        lo[e] = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
    }
}

// fp32 NHWC (N,H,W,C) -> planes (optionally space-to-depth).  One thread per (chunk, pixel), pixel fastest.
__global__ void split_from_nhwc_kernel(const float* __restrict__ in, int H, int W, int CH, int s2d, int64_t total,
                                       int64_t plane, __half* __restrict__ out, int write_lo, const float* __restrict__ mul) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float g = mul ? mul[0] : 1.f;          // power-of-two pre-scale (training gradients), exact
    const int64_t hw = (int64_t)H * W;
    const int64_t r = i % hw;
    const int64_t nc = i / hw;
    const int chunk = (int)(nc % CH);
    const int64_t n = nc / CH;
    const float4* src = reinterpret_cast<const float4*>(in + ((n * hw + r) * CH + chunk) * 8);
    float4 a = src[0], b = src[1];
    float v[8] = {a.x * g, a.y * g, a.z * g, a.w * g, b.x * g, b.y * g, b.z * g, b.w * g};
    float4 hi, lo;
    split8(v, hi, lo);
    size_t off;
    if (!s2d) {
        off = (((size_t)n * CH + chunk) * hw + r) * 8;
    } else {
        const int y = (int)(r / W), x = (int)(r - (int64_t)y * W);
        const int ph = (y & 1) * 2 + (x & 1);
        off = ((((size_t)n * 4 * CH + ph * CH + chunk) * (H / 2) + (y >> 1)) * (W / 2) + (x >> 1)) * 8;
    }
    *reinterpret_cast<float4*>(out + off) = hi;
    if (write_lo) *reinterpret_cast<float4*>(out + plane + off) = lo;
}

// fp32 NCHW (N,C,H,W) -> planes [pl][N][C/8][H][W][8].  One thread per (n, chunk, pixel).
__global__ void split_from_nchw_kernel(const float* __restrict__ in, int CH, int64_t hw, int64_t total, int64_t plane,
                                       __half* __restrict__ out, int write_lo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t r = i % hw;
    const int64_t nc = i / hw;          // n * CH + chunk
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = in[(nc * 8 + e) * hw + r];
    float4 hi, lo;
    split8(v, hi, lo);
    *reinterpret_cast<float4*>(out + i * 8) = hi;
    if (write_lo) *reinterpret_cast<float4*>(out + plane + i * 8) = lo;
}

// planes [pl][N][CH][H][W][8] -> fp32 NHWC.  One thread per (pixel, chunk), chunk fastest (coalesced writes).
__global__ void merge_to_nhwc_kernel(const __half* __restrict__ in, int H, int W, int CH, int64_t total, int64_t plane,
                                     float* __restrict__ out, int has_lo, const float* __restrict__ mul,
                                     const float* __restrict__ add = nullptr) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int chunk = (int)(i % CH);
    const int64_t pix = i / CH;
    const int64_t hw = (int64_t)H * W;
    const int64_t n = pix / hw, r = pix - n * hw;
    const size_t off = (((size_t)n * CH + chunk) * hw + r) * 8;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    add_pair(in, in + plane, off, has_lo != 0, v);
    if (mul) {                                   // undo a power-of-two pre-scale (training gradients), exact
        const float g = mul[0];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= g;
    }
    if (add) {          // gradient accumulation fused into the merge: out = merged + add (same sum as a separate axpby pass)
        const float4 a0 = reinterpret_cast<const float4*>(add + (pix * CH + chunk) * 8)[0];
        const float4 a1 = reinterpret_cast<const float4*>(add + (pix * CH + chunk) * 8)[1];
        v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w; v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
    }
    float4* dst = reinterpret_cast<float4*>(out + (pix * CH + chunk) * 8);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// planes [pl][N][CH][H][W][8] -> space-to-depth planes [pl][N][4CH][H/2][W/2][8].  Thread per (plane, n, chunk, y, x).
__global__ void s2d_planes_kernel(const float4* __restrict__ in, int H, int W, int CH, int64_t total, float4* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // index into the INPUT (16-byte units)
    if (i >= total) return;
    const int x = (int)(i % W);
    int64_t r = i / W;
    const int y = (int)(r % H);
    r /= H;
    const int chunk = (int)(r % CH);
    const int64_t pn = r / CH;            // plane * N + n
    const int ph = (y & 1) * 2 + (x & 1);
    out[(((pn * 4 * CH + ph * CH + chunk) * (H / 2)) + (y >> 1)) * (W / 2) + (x >> 1)] = in[i];
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

}  // namespace

// 5-D tensor map of an fp16 plane tensor [planes][Nimg][chunks][H][W][8] with a box of box_w pixels x box_h rows x
// box_chunks chunks (one image, one plane); out-of-bounds elements read as zero (TF SAME padding, ragged tiles)
int encode_planes_map(CUtensorMap* map, const __half* base, int planes, int Nimg, int chunks, int H, int W, int box_w, int box_h,
                      int box_chunks) {
    EncodeTiledFn enc = get_encode_fn();
    IC_REQUIRE(enc, IC_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t hw16 = (cuuint64_t)H * W * 16;
    const cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)chunks, (cuuint64_t)Nimg, (cuuint64_t)planes};
    const cuuint64_t strides[4] = {(cuuint64_t)W * 16, hw16, hw16 * chunks, hw16 * chunks * Nimg};
    const cuuint32_t box[5] = {(cuuint32_t)box_w * 8, (cuuint32_t)box_h, (cuuint32_t)box_chunks, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IC_REQUIRE(r == CUDA_SUCCESS, IC_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d (N=%d H=%d W=%d chunks=%d)", (int)r, Nimg, H, W, chunks);
    return IC_OK;
}

namespace {

template <int T, int NPL, int NOUT, int OUTMODE, int CPG, bool WRES, bool PAIR = false>
int launch_t(const ConvTcArgs& a, cudaStream_t s) {
    using C = Cfg<T, NOUT, CPG>;
    IC_REQUIRE(!WRES || a.groups->nstages <= MAX_RES_STAGES, IC_ERR_UNSUPPORTED, "conv_tc: too many stages for resident weights");
    EncodeTiledFn enc = get_encode_fn();
    IC_REQUIRE(enc, IC_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    CUtensorMap map;
    const cuuint64_t hw16 = (cuuint64_t)a.Hin * a.Win * 16;
    const cuuint64_t dims[5] = {(cuuint64_t)a.Win * 8, (cuuint64_t)a.Hin, (cuuint64_t)(a.in_chunks_valid ? a.in_chunks_valid : a.in_chunks),
                                (cuuint64_t)a.Nimg, (cuuint64_t)NPL};
    const cuuint64_t strides[4] = {(cuuint64_t)a.Win * 16, hw16, hw16 * a.in_chunks, hw16 * a.in_chunks * a.Nimg};
    const cuuint32_t box[5] = {(cuuint32_t)C::HALO_W * 8, (cuuint32_t)C::HALO_H, (cuuint32_t)CPG, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, (void*)a.in, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IC_REQUIRE(r == CUDA_SUCCESS, IC_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d (N=%d H=%d W=%d chunks=%d)", (int)r, a.Nimg,
               a.Hin, a.Win, a.in_chunks);
    ConvTcParams p;
    p.weights = (const uint8_t*)a.weights;
    p.scale = a.scale;
    p.shift = a.shift;
    p.res1 = a.res1;
    p.res2 = a.res2;
    p.out = a.out;
    p.out_f32 = a.out_f32;
    p.N = a.N;
    p.H = a.H;
    p.W = a.W;
    p.relu = a.relu;
    p.cout = a.cout;
    p.halo0 = a.halo0;
    p.img_mul = a.img_mul;
    p.img_div = a.img_div;
    p.img_div_mul = a.img_div_mul;
    p.img_off_mul = a.img_off_mul ? a.img_off_mul : 1;
    p.img_base = a.img_base;
    {
        const char* pe = getenv("IC_PC_PAIR_CHUNKS");
        p.pair_c2 = (WRES && a.pair_c2 && !(pe && atoi(pe) == 0)) ? 1 : 0;
        // compile-time schedule when the group table is the context model's: 9 taps 0..8, then 5 consecutive taps from 0 or 4
        p.pc_static = 0;
        const GroupTable& g = *a.groups;
        const char* se = getenv("IC_PC_STATIC");
        if (WRES && T == 1 && !(se && atoi(se) == 0) && g.ngroups == 2 && g.ntaps[0] == 9 && g.ntaps[1] == 5 && (g.taps[1][0] == 0 || g.taps[1][0] == 4)) {
            bool ok = true;
            for (int i = 0; i < 9; ++i) ok = ok && g.taps[0][i] == i;
            for (int i = 0; i < 5; ++i) ok = ok && g.taps[1][i] == g.taps[1][0] + i;
            if (ok) p.pc_static = g.taps[1][0] == 0 ? 1 : 2;
        }
        p.walk = 0;
        p.walk_nseg = 1;
    }
    p.res_H = a.res_H ? a.res_H : a.H;
    p.res_W = a.res_W ? a.res_W : a.W;
    p.res_dy = a.res_dy;
    p.res_dx = a.res_dx;
    p.res_div_mul = a.res_div_mul;
    p.res_img_off = a.res_img_off;
    p.res_plane = a.res_plane ? a.res_plane : (size_t)a.N * (NOUT / 8) * a.H * a.W * 8;
    p.head = a.head;
    // the tensor core's fp32 accumulate rounds toward zero: measured bias -1.606e-8 relative per accumulate step
    // (tools/tc_bias.py: -3.47e-6 +- 0.05e-6 for 216 steps, independent of layer and input distribution); undo its mean
    // IC_TC_ACC_GAIN=0 switches the compensation off (tests/test_gpu_conv_tc.py measures both against float64)
    const char* gain_env = getenv("IC_TC_ACC_GAIN");
    // (B-concatenated instantiations apply it to the main accumulator alone: eff_ksteps accumulate steps)
    int acc_steps = a.groups->eff_ksteps * ((NPL == 2 && !WRES) ? 3 : 1);
    if (WRES && a.pair_c2 && a.groups->eff_ksteps == 2 * a.groups->nstages) {       // paired third chunks: fewer accumulate steps
        acc_steps = a.groups->nstages;
        for (int g = 0; g < a.groups->ngroups; ++g) acc_steps += (a.groups->ntaps[g] + 1) / 2;
    }
    p.acc_gain = (gain_env && atoi(gain_env) == 0) ? 1.0f : 1.0f + 1.606e-8f * (float)acc_steps;
    p.out_s2d = a.out_s2d;
    p.store_chunks = a.store_chunks;
    p.d2s_cch = a.d2s_cch;
    p.d2s_ph0 = a.d2s_ph0;
    p.denorm = a.denorm;
    p.out_u8 = a.out_u8;
    p.symbols = a.symbols;
    p.out_freqs = a.out_freqs;
    p.bits_sum = a.bits_sum;
    p.dbg = nullptr;
    static unsigned long long* dbg_buf = nullptr;
    const char* dbg_env = getenv("IC_TC_DBG");
    const bool dbg_on = dbg_env && atoi(dbg_env) && ((NOUT == 128 && OUTMODE == 0) || (WRES && atoi(dbg_env) == 2) || atoi(dbg_env) == 3);
    if (dbg_on) {
        if (!dbg_buf) IC_CHECK_CUDA(cudaMalloc((void**)&dbg_buf, 4096 * 4 * sizeof(unsigned long long)));
        IC_CHECK_CUDA(cudaMemsetAsync(dbg_buf, 0, 4096 * 4 * sizeof(unsigned long long), s));
        p.dbg = dbg_buf;
    }
    // activation-tile slots (2 ... MAX_ASLOTS; IC_PC_ASLOTS overrides the context-model kernels' count)
    p.aslots = 2;
    if (WRES) {
        const char* se = getenv("IC_PC_ASLOTS");
        const int v = se ? atoi(se) : 2;
        p.aslots = v < 2 ? 2 : (v > MAX_ASLOTS ? MAX_ASLOTS : v);
    }
    const size_t smem = p.aslots * NPL * C::A_PLANE_BYTES +
                        (WRES ? MAX_RES_STAGES : (PAIR ? 2 * C::WSTAGES : C::WSTAGES)) * NPL * (C::W_PLANE_BYTES / (PAIR ? 2 : 1)) + 1024 +
                        sizeof(Barriers) + 64;
    CUtensorMap wmap = map;        // placeholder unless PAIR
    if (PAIR) {
        // pair-packed weights [stage][half][plane][4][64][8] as a 3-D tensor [stage*2+half][16][256] of fp16;
        // a CTA's half stage = NPL planes = 8*NPL rows of 256 elements
        const cuuint64_t wd[3] = {256, 16, (cuuint64_t)a.groups->nstages * 2};
        const cuuint64_t wst[2] = {512, 8192};
        const cuuint32_t wbox[3] = {256, (cuuint32_t)(8 * NPL), 1};
        const cuuint32_t we[3] = {1, 1, 1};
        CUresult wr = enc(&wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)a.weights_pair, wd, wst, wbox, we,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IC_REQUIRE(wr == CUDA_SUCCESS, IC_ERR_CUDA, "cuTensorMapEncodeTiled (weights) failed: %d", (int)wr);
    }
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        IC_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<T, NPL, NOUT, OUTMODE, CPG, WRES, PAIR>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    const int tiles_x = (a.W + TW * T - 1) / (TW * T), tiles_y = (a.H + TH - 1) / TH;
    const int n_super = a.N * tiles_y * tiles_x;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // two CTAs per SM only if both shared memory and TMEM (512 columns per SM) allow it
    const int tmem_cols = 2 * T * C::NCOL;
    const int ctas = (2 * (smem + 1024) <= 232448 && tmem_cols <= 256) ? 2 * sms : sms;
    int grid = n_super < ctas ? n_super : ctas;
    if (WRES && T == 1 && OUTMODE != 1 && p.pc_static == 1 && a.img_div >= 1 && a.img_mul == 1 && a.img_div_mul == 1 && p.img_off_mul == 1 &&
        a.img_base == 0 && a.N % a.img_div == 0 && a.groups->img_off[0] == 0 && a.groups->img_off[1] == 1) {
        // depth walk (inference layers of the context model): (image, tile) columns, cut into depth segments only when there
        // are too few columns to fill the machine (a segment of s outputs loads s + 1 input slices)
        const char* we = getenv("IC_PC_WALK");
        if (!(we && atoi(we) == 0)) {
            const int dout = a.img_div, ncols = (a.N / dout) * tiles_y * tiles_x;
            int nseg = (2 * ctas + ncols - 1) / ncols;
            nseg = nseg < 1 ? 1 : (nseg > dout ? dout : nseg);
            p.walk = (dout + nseg - 1) / nseg;
            p.walk_nseg = (dout + p.walk - 1) / p.walk;
            const int n_items = ncols * p.walk_nseg;
            grid = n_items < ctas ? n_items : ctas;
        }
    }
    ProfScope ps(a.prof_class, s);
    if (PAIR) {
        const int n_work = (n_super + 1) / 2;
        grid = 2 * (n_work < sms / 2 ? n_work : sms / 2);
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(nthreads(NOUT, OUTMODE));
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        IC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<T, NPL, NOUT, OUTMODE, CPG, WRES, PAIR>, map, wmap, p, *a.groups));
    } else {
        conv_tc_kernel<T, NPL, NOUT, OUTMODE, CPG, WRES, PAIR><<<grid, nthreads(NOUT, OUTMODE), smem, s>>>(map, wmap, p, *a.groups);
    }
    IC_CHECK_LAUNCH();
    if (dbg_on) {       // diagnostic mode only: synchronises and prints where the MMA issuers waited
        static int printed = 0;
        std::vector<unsigned long long> h(4096 * 4);
        IC_CHECK_CUDA(cudaStreamSynchronize(s));
        IC_CHECK_CUDA(cudaMemcpy(h.data(), dbg_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        double w[4] = {0, 0, 0, 0};
        int n = 0;
        for (int b = 0; b < grid; ++b)
            if (h[b * 4 + 3]) {
                for (int k = 0; k < 4; ++k) w[k] += (double)h[b * 4 + k];
                ++n;
            }
        if (n && printed++ < 40)
            fprintf(stderr, "[IC_TC_DBG] T=%d nout=%d outmode=%d pair=%d issuers=%d  wait acc_empty %.1f%%  a_full %.1f%%  w_full %.1f%%  of %.0f cycles\n", T, NOUT, OUTMODE, (int)PAIR, n,
                    100 * w[0] / w[3], 100 * w[1] / w[3], 100 * w[2] / w[3], w[3] / n);
    }
    return IC_OK;
}

// EXACT 128-output-channel convs on the B-concatenated kernel (single CTA or CTA pair)
template <bool PAIR>
int launch_cat(const ConvTcArgs& a, cudaStream_t s) {
    using C = Cfg<2, 128, 4>;
    EncodeTiledFn enc = get_encode_fn();
    IC_REQUIRE(enc, IC_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const GroupTable& gt = *a.groups;
    IC_REQUIRE(gt.nstages <= CAT_MAX_STAGES && gt.nstages >= 2 * CAT_SK, IC_ERR_UNSUPPORTED, "conv_cat: %d stages", gt.nstages);
    CatStageTable tab;
    memset(&tab, 0, sizeof(tab));
    int si = 0;
    for (int g = 0; g < gt.ngroups; ++g) {
        // tile 1 trails tile 0 by CAT_SK stages: an activation slot is released CAT_SK stages into the next group
        IC_REQUIRE(gt.ntaps[g] >= CAT_SK, IC_ERR_UNSUPPORTED, "conv_cat: group %d has %d taps", g, gt.ntaps[g]);
        IC_REQUIRE(gt.img_off[g] == 0, IC_ERR_UNSUPPORTED, "conv_cat: 2-D convs only");
        for (int ti = 0; ti < gt.ntaps[g]; ++ti, ++si) {
            const int tap = gt.taps[g][ti];
            tab.aoff[si] = (uint16_t)((tap / 3) * C::HALO_W + tap % 3);
            tab.flag[si] = (uint8_t)((ti == 0 ? 1 : 0) | (ti == gt.ntaps[g] - 1 ? 2 : 0));
        }
    }
    CUtensorMap map;
    const cuuint64_t hw16 = (cuuint64_t)a.Hin * a.Win * 16;
    const cuuint64_t dims[5] = {(cuuint64_t)a.Win * 8, (cuuint64_t)a.Hin, (cuuint64_t)a.in_chunks, (cuuint64_t)a.Nimg, 2};
    const cuuint64_t strides[4] = {(cuuint64_t)a.Win * 16, hw16, hw16 * a.in_chunks, hw16 * a.in_chunks * a.Nimg};
    const cuuint32_t box[5] = {(cuuint32_t)C::HALO_W * 8, (cuuint32_t)C::HALO_H, 4, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, (void*)a.in, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IC_REQUIRE(r == CUDA_SUCCESS, IC_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d (N=%d H=%d W=%d chunks=%d)", (int)r, a.Nimg,
               a.Hin, a.Win, a.in_chunks);
    CUtensorMap wmap = map;        // placeholder unless PAIR
    if (PAIR) {
        // pair-packed weights [stage][rank][12 KB] as a 3-D fp16 tensor [stage*2+rank][24][256]
        const cuuint64_t wd[3] = {256, 24, (cuuint64_t)gt.nstages * 2};
        const cuuint64_t wst[2] = {512, 12288};
        const cuuint32_t wbox[3] = {256, 24, 1};
        const cuuint32_t we[3] = {1, 1, 1};
        CUresult wr = enc(&wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)a.weights_cat_pair, wd, wst, wbox, we,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IC_REQUIRE(wr == CUDA_SUCCESS, IC_ERR_CUDA, "cuTensorMapEncodeTiled (weights) failed: %d", (int)wr);
    }
    ConvTcParams p;
    memset(&p, 0, sizeof(p));
    p.weights = (const uint8_t*)a.weights_cat;
    p.scale = a.scale;
    p.shift = a.shift;
    p.res1 = a.res1;
    p.res2 = a.res2;
    p.out = a.out;
    p.N = a.N;
    p.H = a.H;
    p.W = a.W;
    p.relu = a.relu;
    p.cout = a.cout;
    p.halo0 = a.halo0;
    p.out_s2d = a.out_s2d;
    // The tensor core's fp32 accumulate rounds toward zero: a relative bias of -1.606e-8 per accumulate step on the main
    // (hi*hi) sum (tools/tc_bias.py; the cross sum is 2^-11 times smaller, its truncation does not matter).  Undo the mean;
    // IC_TC_ACC_GAIN=0 switches the compensation off (tests/test_gpu_conv_tc.py measures both on sign-mixed inputs).
    const char* gain_env = getenv("IC_TC_ACC_GAIN");
    const bool gain_on = !(gain_env && atoi(gain_env) == 0);
    p.acc_gain = gain_on ? 1.0f + 1.606e-8f * (float)gt.eff_ksteps : 1.0f;
    const size_t smem = 2 * 2 * C::A_PLANE_BYTES + (PAIR ? CAT_WSTAGES_PAIR * 12288 : CAT_WSTAGES * 16384) + 1024 + sizeof(CatBarriers) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        IC_CHECK_CUDA(cudaFuncSetAttribute(conv_cat_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int tiles_x = (a.W + TW * 2 - 1) / (TW * 2), tiles_y = (a.H + TH - 1) / TH;
    const int n_super = a.N * tiles_y * tiles_x;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    ProfScope ps(a.prof_class, s);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(CAT_NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    if (PAIR) {
        const int n_work = (n_super + 1) / 2;
        cfg.gridDim = dim3(2 * (n_work < sms / 2 ? n_work : sms / 2));
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    } else {
        cfg.gridDim = dim3(n_super < sms ? n_super : sms);
    }
    IC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_cat_kernel<PAIR>, map, wmap, p, gt, tab));
    IC_CHECK_LAUNCH();
    return IC_OK;
}

template <int NOUT, int OUTMODE>
int launch_n(const ConvTcArgs& a, cudaStream_t s) {
    const bool wide = a.W > 8;
    if (a.exact) return wide ? launch_t<2, 2, NOUT, OUTMODE, 4, false>(a, s) : launch_t<1, 2, NOUT, OUTMODE, 4, false>(a, s);
    return wide ? launch_t<2, 1, NOUT, OUTMODE, 4, false>(a, s) : launch_t<1, 1, NOUT, OUTMODE, 4, false>(a, s);
}

// context model: 32-channel (4 chunk) groups, always hi/lo planes
template <int NOUT, int OUTMODE>
int launch_pc(const ConvTcArgs& a, cudaStream_t s) {
    // 16x8 tiles (T = 1): 102 KB of shared memory and 128 TMEM columns per CTA, so two CTAs share an SM and
    // overlap each other's pipeline bubbles (these layers have only 84 short MMAs per tile)
    return launch_t<1, 2, NOUT, OUTMODE, 4, true>(a, s);
}

}  // namespace

int count_eff_ksteps(const std::vector<__half>& packed, int nstages, int nout);

int launch_conv_tc(const ConvTcArgs& a, cudaStream_t s) {
    IC_REQUIRE(a.N > 0 && a.H > 0 && a.W > 0 && a.groups, IC_ERR_INVALID, "conv_tc: bad shape");
    IC_REQUIRE(((uintptr_t)a.in & 15) == 0, IC_ERR_INVALID, "conv_tc: unaligned input");
    IC_REQUIRE(a.cpg == 4, IC_ERR_UNSUPPORTED, "conv_tc: groups are 32 channels (4 chunks)");
    if (a.nout == 128 && a.out && a.exact && a.W > 8 && !a.d2s_cch && (a.weights_cat || a.weights_cat_pair)) {
        // B-concatenated kernel: needs >= CAT_SK taps per group (3x3 convs and h2 qualify)
        bool ok = a.groups->nstages <= CAT_MAX_STAGES && a.groups->nstages >= 2 * CAT_SK;
        for (int g = 0; g < a.groups->ngroups; ++g) ok = ok && a.groups->ntaps[g] >= CAT_SK && a.groups->img_off[g] == 0;
        if (ok) return a.weights_cat_pair ? launch_cat<true>(a, s) : launch_cat<false>(a, s);
    }
    if (a.nout == 128 && a.out && a.weights_pair && a.W > 8) {       // 2-CTA pairs (cta_group::2)
        return a.exact ? launch_t<2, 2, 128, 0, 4, false, true>(a, s) : launch_t<2, 1, 128, 0, 4, false, true>(a, s);
    }
    if (a.nout == 128 && a.out) return launch_n<128, 0>(a, s);
    if (a.nout == 64 && a.out) return launch_n<64, 0>(a, s);
    if (a.nout == 256 && a.out && a.d2s_cch)      // depth-to-space transposed convs: one 256-column tile, T = 1
        return a.exact ? launch_t<1, 2, 256, 0, 4, false>(a, s) : launch_t<1, 1, 256, 0, 4, false>(a, s);
    if (a.pc_f32) {        // training: context-model layers with float32 NHWC output (forward and data gradient)
        IC_REQUIRE(a.exact && a.out_f32 && (a.nout == 32 || a.nout == 16), IC_ERR_UNSUPPORTED, "conv_tc: pc_f32 needs nout 32 / 16");
        return a.nout == 32 ? launch_pc<32, 1>(a, s) : launch_pc<16, 1>(a, s);
    }
    if (a.nout == 16 && a.head < 0 && a.out_f32) return launch_n<16, 3>(a, s);
    if (a.nout == 48 && a.out_f32) return launch_n<48, 1>(a, s);
    if (a.nout == 80 && a.out_f32) return launch_n<80, 1>(a, s);
    if (a.nout == 32 || a.nout == 16) IC_REQUIRE(a.exact, IC_ERR_UNSUPPORTED, "conv_tc: the context model runs in hi/lo precision only");
    if (a.nout == 32 && a.out) return launch_pc<32, 0>(a, s);
    if (a.nout == 16 && a.head >= 0) return launch_pc<16, 2>(a, s);
    set_error("conv_tc: unsupported output configuration (nout=%d)", a.nout);
    return IC_ERR_UNSUPPORTED;
}

int launch_split_from_nhwc(const float* in, int N, int H, int W, int C, int s2d, __half* out, int write_lo, cudaStream_t s,
                           const float* d_mul) {
    IC_REQUIRE(C % 8 == 0 && (!s2d || (H % 2 == 0 && W % 2 == 0)), IC_ERR_INVALID, "split_from_nhwc: bad shape");
    int64_t total = (int64_t)N * H * W * (C / 8), plane = (int64_t)N * H * W * C;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    split_from_nhwc_kernel<<<cdiv(total, 256), 256, 0, s>>>(in, H, W, C / 8, s2d, total, plane, out, write_lo, d_mul);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int launch_split_from_nchw(const float* in, int N, int C, int H, int W, __half* out, int write_lo, cudaStream_t s) {
    IC_REQUIRE(C % 8 == 0, IC_ERR_INVALID, "split_from_nchw: C must be a multiple of 8");
    int64_t hw = (int64_t)H * W, total = (int64_t)N * (C / 8) * hw, plane = total * 8;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    split_from_nchw_kernel<<<cdiv(total, 256), 256, 0, s>>>(in, C / 8, hw, total, plane, out, write_lo);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int launch_merge_to_nhwc(const __half* in, int N, int H, int W, int C, float* out, int has_lo, cudaStream_t s, const float* d_mul,
                         const float* d_add) {
    int64_t total = (int64_t)N * H * W * (C / 8), plane = (int64_t)N * H * W * C;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    merge_to_nhwc_kernel<<<cdiv(total, 256), 256, 0, s>>>(in, H, W, C / 8, total, plane, out, has_lo, d_mul, d_add);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

int launch_s2d_planes(const __half* in, int N, int H, int W, int C, int planes, __half* out, cudaStream_t s) {
    IC_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, IC_ERR_INVALID, "s2d_planes: bad shape");
    int64_t total = (int64_t)planes * N * (C / 8) * H * W;
    ProfScope ps(IC_PROF_ELEMENTWISE, s);
    s2d_planes_kernel<<<cdiv(total, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(in), H, W, C / 8, total,
                                                       reinterpret_cast<float4*>(out));
    IC_CHECK_LAUNCH();
    return IC_OK;
}

// Host: conv2d HWIO float weights (k,k,cin,cout) -> the kernel's stage order (fp16 hi/lo, scaled by
// 2^e) plus the group/tap table.
//   k = 3, stride 1: groups = cin/64 halves, 9 taps each.
//   k = 5, stride 2 (input in space-to-depth layout): groups = 4 phases x cin/64; tap ky belongs to
//   phase (ky-1)&1 with halo row floor((ky-1)/2)+1 (TF SAME on an even size: pad_before = 1).
// Stage = (group, tap, cin32 block j); within a stage [plane][4 chunks][nout rows][8 cin].
int pack_weights(const float* w_hwio, int k, int stride, int cin, int cout, int nout, std::vector<__half>& packed,
                 GroupTable& gt, float* inv_scale_out) {
    if (!((k == 3 && stride == 1) || (k == 5 && stride == 2)) || cin % 32 != 0 || cout > nout) return IC_ERR_UNSUPPORTED;
    float mx = 0.f;
    for (size_t i = 0; i < (size_t)k * k * cin * cout; ++i) mx = fmaxf(mx, fabsf(w_hwio[i]));
    int e = 0;
    if (mx > 0.f) {
        int ex;
        frexpf(mx, &ex);        // mx = f * 2^ex, f in [0.5,1)
        e = 8 - ex;             // scaled max in [2^7, 2^8): w_lo stays a normal fp16 number
    }
    const float sc = ldexpf(1.f, e);
    *inv_scale_out = ldexpf(1.f, -e);
    const int quarters = cin / 32;            // one group = 32 input channels (4 chunks)
    memset(&gt, 0, sizeof(gt));
    struct Tap { int ky, kx; };
    std::vector<std::vector<Tap>> gtaps;
    std::vector<int> gq;
    if (stride == 1) {
        for (int q = 0; q < quarters; ++q) {
            std::vector<Tap> t;
            for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx) t.push_back({ky, kx});
            gtaps.push_back(t);
            gq.push_back(q);
        }
    } else {
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px)
                for (int q = 0; q < quarters; ++q) {
                    std::vector<Tap> t;
                    for (int ky = 0; ky < 5; ++ky)
                        for (int kx = 0; kx < 5; ++kx)
                            if ((((ky - 1) & 1) == py) && (((kx - 1) & 1) == px)) t.push_back({ky, kx});
                    gtaps.push_back(t);
                    gq.push_back(q);
                }
    }
    if (gtaps.size() > 16) return IC_ERR_UNSUPPORTED;
    gt.ngroups = (int)gtaps.size();
    for (int g = 0; g < gt.ngroups; ++g) gt.chunk0[g] = (uint8_t)(g * 4);
    const size_t plane_elems = (size_t)4 * nout * 8;
    packed.clear();
    int nst = 0;
    for (size_t g = 0; g < gtaps.size(); ++g) {
        gt.ntaps[g] = (uint8_t)gtaps[g].size();
        for (size_t ti = 0; ti < gtaps[g].size(); ++ti, ++nst) {
            const Tap tp = gtaps[g][ti];
            int dy, dx;
            if (stride == 1) {
                dy = tp.ky;
                dx = tp.kx;
            } else {   // floor((k-1)/2) + 1 for k-1 in {-1,0,1,2,3}
                dy = ((tp.ky - 1) >> 1) + 1;
                dx = ((tp.kx - 1) >> 1) + 1;
            }
            gt.taps[g][ti] = (uint8_t)(dy * 3 + dx);
            const size_t base = packed.size();
            packed.resize(base + 2 * plane_elems, __float2half(0.f));
            for (int ch = 0; ch < 4; ++ch)
                for (int co = 0; co < cout; ++co)
                    for (int ei = 0; ei < 8; ++ei) {
                        const int ci = gq[g] * 32 + ch * 8 + ei;
                        const float v = w_hwio[(((size_t)tp.ky * k + tp.kx) * cin + ci) * cout + co] * sc;
                        const __half hi = __float2half_rn(v);
                        const __half lo = __float2half_rn(v - __half2float(hi));
                        const size_t idx = ((size_t)ch * nout + co) * 8 + ei;
                        packed[base + idx] = hi;
                        packed[base + plane_elems + idx] = lo;
                    }
        }
    }
    gt.nstages = nst;
    gt.eff_ksteps = count_eff_ksteps(packed, nst, nout);
    return IC_OK;
}

// conv2d_transpose weights [k][k][Cout][Cin] (code/autoencoder.py:251,264-265), stride 2, as a stride-1 conv over the
// INPUT grid whose output columns are (output phase, channel): y[2m+r] = sum_t x[m + (r+pb-t)/2] w[t] over taps with
// (r+pb-t) even (pb = 0 for k = 3, 1 for k = 5).  `phases` lists the output phases (ry*2+rx) this
// launch produces; halo offset of a tap = (r+pb-t)/2 + 1 in {0,1,2}.
int pack_weights_tconv(const float* w, int k, int cin, int cout, const int* phases, int nphases, int nout,
                       std::vector<__half>& packed, GroupTable& gt, float* inv_scale_out) {
    if (!(k == 3 || k == 5) || cin % 32 != 0 || nphases * cout > nout) return IC_ERR_UNSUPPORTED;
    const int pb = k == 3 ? 0 : 1;
    float mx = 0.f;
    for (size_t i = 0; i < (size_t)k * k * cin * cout; ++i) mx = fmaxf(mx, fabsf(w[i]));
    int e = 0;
    if (mx > 0.f) {
        int ex;
        frexpf(mx, &ex);
        e = 8 - ex;
    }
    const float sc = ldexpf(1.f, e);
    *inv_scale_out = ldexpf(1.f, -e);
    memset(&gt, 0, sizeof(gt));
    const int quarters = cin / 32;
    if (quarters > 16) return IC_ERR_UNSUPPORTED;
    // tap positions (dy,dx) used by at least one phase of this launch
    bool used[9] = {false};
    for (int pi = 0; pi < nphases; ++pi) {
        const int ry = phases[pi] >> 1, rx = phases[pi] & 1;
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) {
                const int ty = ry + pb - 2 * (dy - 1), tx = rx + pb - 2 * (dx - 1);
                if (ty >= 0 && ty < k && tx >= 0 && tx < k) used[dy * 3 + dx] = true;
            }
    }
    gt.ngroups = quarters;
    const size_t plane_elems = (size_t)4 * nout * 8;
    packed.clear();
    int nst = 0;
    for (int q = 0; q < quarters; ++q) {
        gt.chunk0[q] = (uint8_t)(q * 4);
        int nt = 0;
        for (int tap = 0; tap < 9; ++tap) {
            if (!used[tap]) continue;
            const int dy = tap / 3, dx = tap % 3;
            gt.taps[q][nt++] = (uint8_t)tap;
            const size_t base = packed.size();
            packed.resize(base + 2 * plane_elems, __float2half(0.f));
            for (int pi = 0; pi < nphases; ++pi) {
                const int ry = phases[pi] >> 1, rx = phases[pi] & 1;
                const int ty = ry + pb - 2 * (dy - 1), tx = rx + pb - 2 * (dx - 1);
                if (ty < 0 || ty >= k || tx < 0 || tx >= k) continue;
                for (int ch = 0; ch < 4; ++ch)
                    for (int co = 0; co < cout; ++co)
                        for (int ei = 0; ei < 8; ++ei) {
                            const int ci = q * 32 + ch * 8 + ei;
                            const float v = w[(((size_t)ty * k + tx) * cout + co) * cin + ci] * sc;
                            const __half hi = __float2half_rn(v);
                            const __half lo = __float2half_rn(v - __half2float(hi));
                            const size_t idx = ((size_t)ch * nout + pi * cout + co) * 8 + ei;
                            packed[base + idx] = hi;
                            packed[base + plane_elems + idx] = lo;
                        }
            }
            ++nst;
        }
        gt.ntaps[q] = (uint8_t)nt;
    }
    gt.nstages = nst;
    gt.eff_ksteps = count_eff_ksteps(packed, nst, nout);
    return IC_OK;
}

// number of (stage, k-step) pairs whose hi-plane weight block is not all zero = MMAs that really add something
int count_eff_ksteps(const std::vector<__half>& packed, int nstages, int nout) {
    const size_t plane_elems = (size_t)4 * nout * 8;
    int n = 0;
    for (int s = 0; s < nstages; ++s)
        for (int ks = 0; ks < 2; ++ks) {
            bool nz = false;
            const __half* b = packed.data() + (size_t)s * 2 * plane_elems + (size_t)ks * 2 * nout * 8;
            for (size_t i = 0; i < (size_t)2 * nout * 8 && !nz; ++i) nz = __half2float(b[i]) != 0.f;
            n += nz ? 1 : 0;
        }
    return n;
}

// standard stage layout [plane][4][128][8] -> pair layout [half][plane][4][64][8] (each CTA of a pair holds 64 B rows)
void repack_pair(const std::vector<__half>& packed, int nstages, std::vector<__half>& out) {
    out.resize(packed.size());
    const size_t stage = 2 * 4 * 128 * 8;
    for (int s = 0; s < nstages; ++s)
        for (int half = 0; half < 2; ++half)
            for (int pl = 0; pl < 2; ++pl)
                for (int ch = 0; ch < 4; ++ch)
                    for (int r = 0; r < 64; ++r)
                        for (int e = 0; e < 8; ++e)
                            out[s * stage + ((((size_t)half * 2 + pl) * 4 + ch) * 64 + r) * 8 + e] =
                                packed[s * stage + (((size_t)pl * 4 + ch) * 128 + half * 64 + r) * 8 + e];
}

// standard stage layout [plane][4][128][8] -> B-concatenated [4][hi rows 0..127 | lo rows 128..255][8] (conv_cat_kernel)
void repack_cat(const std::vector<__half>& packed, int nstages, std::vector<__half>& out) {
    out.resize(packed.size());
    const size_t stage = 2 * 4 * 128 * 8;
    for (int s = 0; s < nstages; ++s)
        for (int pl = 0; pl < 2; ++pl)
            for (int ch = 0; ch < 4; ++ch)
                for (int r = 0; r < 128; ++r)
                    for (int e = 0; e < 8; ++e)
                        out[s * stage + (((size_t)ch * 2 + pl) * 128 + r) * 8 + e] =
                            packed[s * stage + (((size_t)pl * 4 + ch) * 128 + r) * 8 + e];
}

// ... -> pair layout [stage][rank][P: [4][128][8] (rank 0: hi, rank 1: lo) | Q: [4][64][8] = hi rows 64*rank ..]  (12 KB per CTA)
void repack_cat_pair(const std::vector<__half>& packed, int nstages, std::vector<__half>& out) {
    const size_t stage = 2 * 4 * 128 * 8, half_stage = 6144;
    out.assign((size_t)nstages * 2 * half_stage, __float2half(0.f));
    for (int s = 0; s < nstages; ++s)
        for (int rk = 0; rk < 2; ++rk) {
            __half* dst = out.data() + ((size_t)s * 2 + rk) * half_stage;
            for (int ch = 0; ch < 4; ++ch) {
                for (int r = 0; r < 128; ++r)
                    for (int e = 0; e < 8; ++e)
                        dst[((size_t)ch * 128 + r) * 8 + e] = packed[s * stage + (((size_t)rk * 4 + ch) * 128 + r) * 8 + e];
                for (int r = 0; r < 64; ++r)
                    for (int e = 0; e < 8; ++e)
                        dst[4096 + ((size_t)ch * 64 + r) * 8 + e] = packed[s * stage + (((size_t)0 * 4 + ch) * 128 + rk * 64 + r) * 8 + e];
            }
        }
}

// h1: conv2d 5x5 stride 2 with cin <= 8 (RGB) on a space-to-depth input whose 4 phases x 8 padded channels form
// ONE 32-channel group: 9 stride-1 taps; tap (dy,dx) carries, for phase (py,px), the weights of
// ky = 2(dy-1)+py+1, kx = 2(dx-1)+px+1 when those are inside the 5x5 window (zero otherwise).
int pack_weights_h1(const float* w_hwio, int cin, int cout, int nout, std::vector<__half>& packed, GroupTable& gt,
                    float* inv_scale_out) {
    if (cin > 8 || cout > nout) return IC_ERR_UNSUPPORTED;
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
    memset(&gt, 0, sizeof(gt));
    gt.ngroups = 1;
    gt.ntaps[0] = 9;
    gt.nstages = 9;
    gt.eff_ksteps = 0;       // filled below
    const size_t plane_elems = (size_t)4 * nout * 8;
    packed.assign((size_t)9 * 2 * plane_elems, __float2half(0.f));
    for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx) {
            const int tap = dy * 3 + dx;
            gt.taps[0][tap] = (uint8_t)tap;
            const size_t base = (size_t)tap * 2 * plane_elems;
            for (int py = 0; py < 2; ++py)
                for (int px = 0; px < 2; ++px) {
                    const int ky = 2 * (dy - 1) + py + 1, kx = 2 * (dx - 1) + px + 1;
                    if (ky < 0 || ky > 4 || kx < 0 || kx > 4) continue;
                    const int chunk = py * 2 + px;                  // phase = chunk of the s2d input
                    for (int c = 0; c < cin; ++c)
                        for (int o = 0; o < cout; ++o) {
                            const float v = w_hwio[(((size_t)ky * 5 + kx) * cin + c) * cout + o] * sc;
                            const __half hi = __float2half_rn(v);
                            const __half lo = __float2half_rn(v - __half2float(hi));
                            const size_t idx = ((size_t)chunk * nout + o) * 8 + c;
                            packed[base + idx] = hi;
                            packed[base + plane_elems + idx] = lo;
                        }
                }
        }
    gt.eff_ksteps = count_eff_ksteps(packed, 9, nout);
    return IC_OK;
}

// Context-model layer: conv3d weights [2][3][3][ci][co] (code/probclass.py:249-257) with the "other" mask
// (code/probclass.py:164-176) -> 2 groups (filter depth 0: 9 taps, depth 1: 5 taps) of one 32-channel stage
// per tap; ci <= 32 and co <= nout are zero padded.
int pack_weights_pc(const float* w, int ci, int co, int nout, std::vector<__half>& packed, GroupTable& gt,
                    float* inv_scale_out) {
    if (ci > 32 || co > nout) return IC_ERR_UNSUPPORTED;
    float mx = 0.f;
    for (int i = 0; i < 18 * ci * co; ++i) mx = fmaxf(mx, fabsf(w[i]));
    int e = 0;
    if (mx > 0.f) {
        int ex;
        frexpf(mx, &ex);
        e = 8 - ex;
    }
    const float sc = ldexpf(1.f, e);
    *inv_scale_out = ldexpf(1.f, -e);
    memset(&gt, 0, sizeof(gt));
    gt.ngroups = 2;
    const size_t plane_elems = (size_t)4 * nout * 8;
    packed.clear();
    int nst = 0;
    for (int fd = 0; fd < 2; ++fd) {
        int nt = 0;
        gt.img_off[fd] = (uint8_t)fd;
        gt.chunk0[fd] = 0;
        for (int fy = 0; fy < 3; ++fy)
            for (int fx = 0; fx < 3; ++fx) {
                if (fd == 1 && (fy > 1 || (fy == 1 && fx > 1))) continue;
                gt.taps[fd][nt++] = (uint8_t)(fy * 3 + fx);
                const size_t base = packed.size();
                packed.resize(base + 2 * plane_elems, __float2half(0.f));
                for (int c = 0; c < ci; ++c)
                    for (int o = 0; o < co; ++o) {
                        const float v = w[((((size_t)fd * 3 + fy) * 3 + fx) * ci + c) * co + o] * sc;
                        const __half hi = __float2half_rn(v);
                        const __half lo = __float2half_rn(v - __half2float(hi));
                        // B-concatenated stage layout [4 chunks][hi rows | lo rows][8 cin] (conv_tc_kernel, WRES)
                        const size_t idx = ((size_t)(c / 8) * 2 * nout + o) * 8 + (c % 8);
                        packed[base + idx] = hi;
                        packed[base + idx + (size_t)nout * 8] = lo;
                    }
                ++nst;
            }
        gt.ntaps[fd] = (uint8_t)nt;
    }
    gt.nstages = nst;
    gt.eff_ksteps = nst * ((ci + 15) / 16);      // k-steps (16 channels) that hold weights: accumulate steps of the main sum
    return IC_OK;
}

}  // namespace tc
}  // namespace ic
