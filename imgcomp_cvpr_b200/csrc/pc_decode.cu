// Decoder-side context model: the sequential half of --real_bpp
// (code/bit_counter.py:137-163 `_decode`: for every symbol, in raster C -> H -> W
// order, evaluate the context model on the 5x9x9 block of already decoded
// symbols, hand the frequency table to the range decoder, write the symbol back).
//
// The reference pays one sess.run over a full context per symbol (README.md:65:
// ~200 s per Kodak image).  Here one CTA owns one image and keeps the layer
// activations of the causal network (code/probclass.py:214-221) cached, as
// README.md:72-73 suggests: decoding a symbol only adds the terms that depend on
// it.  The range decoder (code/arithmetic_coding.py:163-222) runs on the device
// too, so nothing crosses PCIe per symbol.
//
// Bit consistency with the encoder: every layer output is ONE fmaf chain in the
// order of the float32 batched kernels of probclass.cu (taps in (fd,fy,fx) raster
// order, input channels ascending, then + bias).  That order is causal in the
// decoder's own raster, so the chain of a position can be advanced as its
// inputs appear:
//   taps 0..8   previous latent channel          -> helper warps, one row ahead
//   taps 9..11  row above, same channel          -> all warps, at row start
//   tap  12     left neighbour                   -> decode warp, one step ahead
//   tap  13     the position itself (layers 1-3) -> decode warp, critical path
// and the table it ends in is bit-identical to ic_pc_codec_freqs_fwd's.
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "probclass.cuh"

namespace ic {

namespace {

constexpr int KC = 24;            // arch_param__k
constexpr int PROW = 80;          // floats per position in a partial-sum row: 24 (L0) + 24 (L1) + 24 (L2) + 8 (L3)
constexpr int NT = 256;

struct DecArgs {
    const float *w0, *b0, *w1, *b1, *w2, *b2, *w3, *b3;
    int C, h, w, L;
    float centers[8];
    const uint8_t* stream;          // concatenated bitstreams
    const int64_t* stream_off;      // [N + 1] byte offsets
    const int32_t* first_sym;       // [N] side information (code/bit_counter.py:118-121)
    uint8_t* sym_out;               // N,C,h,w
    float* act;                     // per image: [3 layers][2 channel slots][h+6][w+6][24]
    float* p9;                      // per image: [w+6][PROW]
    uint32_t* words;                // per image: [words_img] zero padded big-endian copy of the stream
    int words_img;
    const uint8_t* force_sym;       // debug: teacher forcing, no range decoding
    int64_t* freqs_out;             // debug: N,C,h,w,L tables as the decoder saw them
    long long* prof;                // debug (IC_PC_DECODE_PROF=1): cycle counters of image 0
};


struct Smem {
    float* w1;     // [14][24][24]
    float* w2;     // [14][24][24]
    float* w3;     // [14][24][8]   outputs zero padded to 8
    float* w0;     // [13][24]
    float* b;      // b0[24] b1[24] b2[24] b3[8]
    float* cent;   // [8]
    float* bc;     // [120] broadcast scratch of the decode warp, [128..139] CoderShared
    float* P;      // [w+7][PROW]  chains of the current row after taps 0..11 (+ one spare row)
};
constexpr int SMEM_FIXED_FLOATS = 14 * 24 * 24 * 2 + 14 * 24 * 8 + 13 * 24 + PROW + 8 + 144;

struct Geom {
    int C, h, w, H6, W6;
    size_t slot, layer;            // strides (floats) of one channel slot / one layer in `act`
};

__device__ __forceinline__ float* act_row(float* act_img, const Geom& g, int layer, int c, int y) {
    return act_img + layer * g.layer + (size_t)((c + 4) & 1) * g.slot + (size_t)(y + 3) * g.W6 * KC;
}

// value of the (virtually padded) input volume, code/probclass.py:268-292 + :449-451
__device__ __forceinline__ float in_val(const Geom& g, const float* cent, const uint8_t* syms, int c, int y, int x) {
    if (c < 0 || y < 0 || y >= g.h || x < 0 || x >= g.w) return cent[0];
    return cent[syms[((size_t)c * g.h + y) * g.w + x]];
}

// ------------------------------------------------------------------ chains, taps 0..11
// FIRST: taps 0..8 (previous channel) from zero into the global row p9;
// !FIRST: taps 9..11 (row above) continue p9 into the shared row P.
template <bool FIRST, int LAYER>
__device__ __forceinline__ void chain_mid(const Geom& g, const Smem& sm, int tid, int nthr, int c, int y, float* act_img,
                                          float* p9) {
    constexpr int NJG = LAYER == 3 ? 2 : 6;
    constexpr int WOUT = LAYER == 3 ? 8 : 24;
    const float* sW = LAYER == 1 ? sm.w1 : (LAYER == 2 ? sm.w2 : sm.w3);
    const int xlo = -3 + LAYER, nx = g.w + 6 - 2 * LAYER;
    float* dst = FIRST ? p9 : sm.P;
    for (int it = tid; it < nx * NJG; it += nthr) {
        const int xo = it / NJG, jg = it - xo * NJG;
        const int xi = xlo + xo + 3;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (!FIRST) {
            const float4 v = *reinterpret_cast<const float4*>(p9 + (size_t)xi * PROW + LAYER * 24 + jg * 4);
            acc[0] = v.x; acc[1] = v.y; acc[2] = v.z; acc[3] = v.w;
        }
#pragma unroll 1
        for (int t = FIRST ? 0 : 9; t < (FIRST ? 9 : 12); ++t) {
            const int dy = FIRST ? t / 3 - 1 : -1;
            const int dx = FIRST ? t % 3 - 1 : t - 10;
            const float4* ap = reinterpret_cast<const float4*>(act_row(act_img, g, LAYER - 1, FIRST ? c - 1 : c, y + dy) +
                                                               (size_t)(xi + dx) * KC);
            const float* wt = sW + (size_t)t * 24 * WOUT + jg * 4;
#pragma unroll
            for (int i4 = 0; i4 < 6; ++i4) {
                const float4 av = ap[i4];
                const float a[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 wv = *reinterpret_cast<const float4*>(wt + (i4 * 4 + u) * WOUT);
                    acc[0] = fmaf(a[u], wv.x, acc[0]);
                    acc[1] = fmaf(a[u], wv.y, acc[1]);
                    acc[2] = fmaf(a[u], wv.z, acc[2]);
                    acc[3] = fmaf(a[u], wv.w, acc[3]);
                }
            }
        }
        *reinterpret_cast<float4*>(dst + (size_t)xi * PROW + LAYER * 24 + jg * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
}

template <bool FIRST>
__device__ __forceinline__ void chain_l0(const Geom& g, const Smem& sm, int tid, int nthr, int c, int y,
                                         const uint8_t* syms, float* p9) {
    float* dst = FIRST ? p9 : sm.P;
    for (int it = tid; it < g.W6 * 6; it += nthr) {
        const int xi = it / 6, jg = it - xi * 6;
        const int x = xi - 3;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (!FIRST) {
            const float4 v = *reinterpret_cast<const float4*>(p9 + (size_t)xi * PROW + jg * 4);
            acc[0] = v.x; acc[1] = v.y; acc[2] = v.z; acc[3] = v.w;
        }
#pragma unroll 1
        for (int t = FIRST ? 0 : 9; t < (FIRST ? 9 : 12); ++t) {
            const int dy = FIRST ? t / 3 - 1 : -1;
            const int dx = FIRST ? t % 3 - 1 : t - 10;
            const float v = in_val(g, sm.cent, syms, FIRST ? c - 1 : c, y + dy, x + dx);
            const float4 wv = *reinterpret_cast<const float4*>(sm.w0 + t * 24 + jg * 4);
            acc[0] = fmaf(v, wv.x, acc[0]);
            acc[1] = fmaf(v, wv.y, acc[1]);
            acc[2] = fmaf(v, wv.z, acc[2]);
            acc[3] = fmaf(v, wv.w, acc[3]);
        }
        *reinterpret_cast<float4*>(dst + (size_t)xi * PROW + jg * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
}

// which layers have an output in row (c, y): layer l lives on c >= l-3, y in [l-3, h+2-l]
__device__ __forceinline__ bool row_has(const Geom& g, int layer, int c, int y) {
    return c >= layer - 3 && y >= layer - 3 && y <= g.h + 2 - layer;
}

template <bool FIRST>
__device__ __forceinline__ void chains(const Geom& g, const Smem& sm, int tid, int nthr, int c, int y, const uint8_t* syms,
                                       float* act_img, float* p9) {
    if (row_has(g, 1, c, y)) chain_mid<FIRST, 1>(g, sm, tid, nthr, c, y, act_img, p9);
    if (row_has(g, 2, c, y)) chain_mid<FIRST, 2>(g, sm, tid, nthr, c, y, act_img, p9);
    if (row_has(g, 3, c, y)) chain_mid<FIRST, 3>(g, sm, tid, nthr, c, y, act_img, p9);
    chain_l0<FIRST>(g, sm, tid, nthr, c, y, syms, p9);
}

// ------------------------------------------------------------------ range decoder
// code/arithmetic_coding.py:163-222 with 32-bit state.  Decoding a symbol has two halves and only
// the first one sits between two symbols:
//   find   (decode warp): value = ((code - low + 1) * total - 1) / range, symbol = last k with
//          cum_k <= value (ArithmeticDecoder.read :181-197) -- one multiplication by 1/range;
//   settle (coder warp):  ArithmeticCoderBase.update (:80-115) for that symbol -- two divisions by
//          `total`, the shifts, the bit reads, the next 1/range -- while the decode warp already
//          runs the network for the next position.
// The two warps hand over through shared memory and two named barriers.
struct CoderShared {
    uint64_t range;          // high - low + 1
    double inv_range;
    uint32_t low, code;
    uint32_t p_lo, p_hi, p_total;     // cumulative bounds of the symbol just found
    uint32_t pad_;
};

__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
constexpr int BAR_FOUND = 1;      // decode warp -> coder warp: p_lo / p_hi / p_total are written
constexpr int BAR_STATE = 2;      // coder warp -> decode warp: low / code / range / inv_range are written

// floor(num / den) from a double estimate, |estimate - exact| < 1; den <= 2^32, num < 2^63
__device__ __forceinline__ uint64_t div_fix(uint64_t num, uint64_t den, double inv_den) {
    uint64_t q = __double2ull_rz(__ull2double_rz(num) * inv_den);
    const int64_t r = (int64_t)(num - q * den);
    q = r < 0 ? q - 1 : (r >= (int64_t)den ? q + 1 : q);
    return q;
}

// The stream has been copied to a zero padded, word aligned scratch (big-endian words), so "bits
// past the end read as zero" (:217-222) needs no test.
struct Coder {
    uint32_t low, high, code;
    const uint32_t* words;   // padded stream as big-endian 32-bit words
    int nwords, pos;         // pos: next word to load into nxt
    uint64_t buf;            // next bits of the stream, MSB first
    int avail;               // valid bits in buf
    uint32_t nxt;            // prefetched word after buf

    __device__ __forceinline__ void init(const uint32_t* w, int n) {
        words = w;
        nwords = n;
        low = 0;
        high = 0xffffffffu;
        code = w[0];                    // ArithmeticDecoder.__init__ reads STATE_SIZE bits (:176-178)
        buf = (uint64_t)w[1] << 32;
        avail = 32;
        nxt = w[2];
        pos = 3;
    }
    // next k bits, 0 <= k <= 32
    __device__ __forceinline__ uint32_t bits(int k) {
        if (avail < 32) {
            buf |= (uint64_t)nxt << (32 - avail);
            avail += 32;
            nxt = words[pos];
            pos = min(pos + 1, nwords - 1);       // the last word of the scratch is zero
        }
        const uint32_t v = (uint32_t)((buf >> 32) >> (32 - k));
        buf <<= k;
        avail -= k;
        return v;
    }
    // The reference shifts one bit per loop iteration (:99-115, :204-213); both loops in closed form:
    //   loop 1 runs while the top bits of low and high agree   -> clz(low ^ high) iterations
    //   loop 2 runs while low = 01..., high = 10...            -> leading ones of ((low & ~high) << 1)
    __device__ __forceinline__ void settle(uint32_t p_lo, uint32_t p_hi, uint32_t p_total) {
        const uint64_t range = (uint64_t)high - low + 1;
        const double inv_t = __drcp_rn((double)p_total);
        const uint64_t lo_b = div_fix((uint64_t)p_lo * range, p_total, inv_t);
        const uint64_t hi_b = p_hi == p_total ? range : div_fix((uint64_t)p_hi * range, p_total, inv_t);
        uint32_t nl = low + (uint32_t)lo_b;
        uint32_t nh = (uint32_t)((uint64_t)low + hi_b - 1);
        const int k = __clz((int)(nl ^ nh));                               // 0..32
        code = (uint32_t)((uint64_t)code << k) | bits(k);
        nl = (uint32_t)((uint64_t)nl << k);
        nh = (uint32_t)(((uint64_t)nh << k) | ((1ull << k) - 1));
        const int m = __clz((int)~((nl & ~nh) << 1));                      // 0..31
        code = (code & 0x80000000u) | ((code << m) & 0x7fffffffu) | bits(m);
        low = (nl << m) & 0x7fffffffu;
        high = ((nh << m) & 0x7fffffffu) | 0x80000000u | ((1u << m) - 1);
    }
    __device__ __forceinline__ void publish(CoderShared* cs, int lane) const {
        if (lane == 0) {
            const uint64_t range = (uint64_t)high - low + 1;
            cs->range = range;
            cs->inv_range = __drcp_rn((double)range);
            cs->low = low;
            cs->code = code;
        }
    }
};

// coder warp: one settle per decoded symbol of row (c, y)
__device__ __forceinline__ void coder_row(const Geom& g, int lane, int c, int y, Coder& cd, CoderShared* cs) {
    if (!row_has(g, 3, c, y)) return;
    const size_t vol = (size_t)g.C * g.h * g.w;
    for (int x = 0; x < g.w; ++x) {
        const size_t at = ((size_t)c * g.h + y) * g.w + x;
        if (at == 0) continue;                    // side information, never coded
        bar_sync(BAR_FOUND);
        cd.settle(cs->p_lo, cs->p_hi, cs->p_total);
        if (at + 1 < vol) {
            cd.publish(cs, lane);
            bar_arrive(BAR_STATE);
        }
    }
}

// ------------------------------------------------------------------ one row, symbol by symbol (warp 0)
// Broadcasts go through shared memory: a lone warp issues one SHFL per 8 cycles (tools/ubench), a
// 24-vector by shuffles costs twice the fmaf chain it feeds.
template <int L>
__device__ __forceinline__ void decode_row(const Geom& g, const Smem& sm, const DecArgs& a, int lane, int c, int y,
                                           CoderShared* cs, uint8_t* sym_img, const uint8_t* force_img, float* act_img, int64_t* freqs_img,
                                           int first_sym, long long* lp) {
    const int j = lane < KC ? lane : KC - 1;      // lanes 24..31 shadow lane 23; their stores are masked
    const int j3 = lane & 7;
    float W1l[KC], W1c[KC], W2l[KC], W2c[KC], W3l[KC], W3c[KC];
#pragma unroll
    for (int i = 0; i < KC; ++i) {
        W1l[i] = sm.w1[(12 * 24 + i) * 24 + j];
        W1c[i] = sm.w1[(13 * 24 + i) * 24 + j];
        W2l[i] = sm.w2[(12 * 24 + i) * 24 + j];
        W2c[i] = sm.w2[(13 * 24 + i) * 24 + j];
        W3l[i] = sm.w3[(12 * 24 + i) * 8 + j3];
        W3c[i] = sm.w3[(13 * 24 + i) * 8 + j3];
    }
    const float w0l = sm.w0[12 * 24 + j];
    const float b0 = sm.b[j], b1 = sm.b[24 + j], b2 = sm.b[48 + j], b3 = sm.b[72 + j3];
    float cent[L];
#pragma unroll
    for (int l = 0; l < L; ++l) cent[l] = sm.cent[l];
    const float pad = cent[0];
    const bool r1 = row_has(g, 1, c, y), r2 = row_has(g, 2, c, y), r3 = row_has(g, 3, c, y);
    float* a0row = act_row(act_img, g, 0, c, y) + lane;
    float* a1row = act_row(act_img, g, 1, c, y) + lane;
    float* a2row = act_row(act_img, g, 2, c, y) + lane;
    const float* P = sm.P;
    float* bc = sm.bc;                            // [3][32] layer outputs, [3][8] head rounds (then CoderShared)
    const bool prof = lp != nullptr;

    float vleft = pad;                            // input at (c, y, -4)
    float acc1 = P[24 + j], acc2 = P[48 + j], acc3 = P[72 + j3];
    float p0 = P[j];
    for (int xi = 0; xi < g.W6; ++xi) {
        const int x = xi - 3;
        const long long t0 = prof ? clock64() : 0;
        // chains of the next position after taps 0..11 (the last iteration reads the spare row of P)
        const float* Pn = P + (size_t)(xi + 1) * PROW;
        float n1 = Pn[24 + j], n2 = Pn[48 + j], n3 = Pn[72 + j3];
        const float p0n = Pn[j];
        // layer 0: tap 12 (left input), bias, ReLU
        const float a0 = fmaxf(__fadd_rn(fmaf(vleft, w0l, p0), b0), 0.f);
        bc[lane] = a0;
        __syncwarp();
        // layer 1: tap 13 of this position, tap 12 of the next one
#pragma unroll
        for (int i4 = 0; i4 < KC / 4; ++i4) {
            const float4 v = reinterpret_cast<const float4*>(bc)[i4];
            const float ai[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc1 = fmaf(ai[u], W1c[i4 * 4 + u], acc1);
                n1 = fmaf(ai[u], W1l[i4 * 4 + u], n1);
            }
        }
        const float a1 = fmaxf(__fadd_rn(acc1, b1), 0.f);
        bc[32 + lane] = a1;
        __syncwarp();
#pragma unroll
        for (int i4 = 0; i4 < KC / 4; ++i4) {
            const float4 v = reinterpret_cast<const float4*>(bc + 32)[i4];
            const float ai[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc2 = fmaf(ai[u], W2c[i4 * 4 + u], acc2);
                n2 = fmaf(ai[u], W2l[i4 * 4 + u], n2);
            }
        }
        const float a2 = __fadd_rn(__fadd_rn(acc2, b2), a0);      // + residual (code/probclass.py:196)
        bc[64 + lane] = a2;
        __syncwarp();
#pragma unroll
        for (int i4 = 0; i4 < KC / 4; ++i4) {
            const float4 v = reinterpret_cast<const float4*>(bc + 64)[i4];
            const float ai[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc3 = fmaf(ai[u], W3c[i4 * 4 + u], acc3);
                n3 = fmaf(ai[u], W3l[i4 * 4 + u], n3);
            }
        }
        const float lg = fmaxf(__fadd_rn(acc3, b3), 0.f);         // ReLU on the logits (code/probclass.py:220,233)
        if (lane < KC) {
            a0row[(size_t)xi * KC] = a0;
            if (r1) a1row[(size_t)xi * KC] = a1;
            if (r2) a2row[(size_t)xi * KC] = a2;
        }
        const long long t1 = prof ? clock64() : 0;
        float vnew = pad;
        if (r3 && x >= 0 && x < g.w) {
            // softmax -> int64(pr * 1e9) -> max(., 1): same operation order as pc_final_kernel<HEAD_FREQS>.
            // lane l < 8 owns logit l (lanes 8..31 repeat them); three broadcast rounds
            float* hb = bc + 96;
            if (lane < 8) hb[lane] = lg;
            __syncwarp();
            float lv[8];
            *reinterpret_cast<float4*>(lv) = reinterpret_cast<const float4*>(hb)[0];
            *reinterpret_cast<float4*>(lv + 4) = reinterpret_cast<const float4*>(hb)[1];
            float m = lv[0];
#pragma unroll
            for (int l = 1; l < L; ++l) m = fmaxf(m, lv[l]);
            const float e = expf(__fsub_rn(lg, m));
            if (lane < 8) hb[8 + lane] = e;
            __syncwarp();
            float ev[8];
            *reinterpret_cast<float4*>(ev) = reinterpret_cast<const float4*>(hb + 8)[0];
            *reinterpret_cast<float4*>(ev + 4) = reinterpret_cast<const float4*>(hb + 8)[1];
            float s = 0.f;
#pragma unroll
            for (int l = 0; l < L; ++l) s = __fadd_rn(s, ev[l]);
            long long f = (long long)__fmul_rn(__fdiv_rn(e, s), 1e9f);
            if (f < 1) f = 1;
            const uint32_t fown = (uint32_t)f;
            if (lane < 8) reinterpret_cast<uint32_t*>(hb)[16 + lane] = fown;
            __syncwarp();
            uint32_t fv[8];
            *reinterpret_cast<uint4*>(fv) = reinterpret_cast<const uint4*>(hb + 16)[0];
            *reinterpret_cast<uint4*>(fv + 4) = reinterpret_cast<const uint4*>(hb + 16)[1];
            uint32_t cums[L + 1];             // cums[k] = sum of f_l, l < k
            cums[0] = 0;
#pragma unroll
            for (int l = 0; l < L; ++l) cums[l + 1] = cums[l] + fv[l];
            const uint32_t total = cums[L];
            const long long t2 = prof ? clock64() : 0;
            const size_t at = ((size_t)c * g.h + y) * g.w + x;
            if (freqs_img && lane < L) freqs_img[at * L + lane] = (int64_t)fown;
            int sym;
            if (force_img) {
                sym = force_img[at];
            } else if (at == 0) {
                sym = first_sym;
            } else {
                // ArithmeticDecoder.read (:181-197); the coder warp has applied the previous symbol
                bar_sync(BAR_STATE);
                const uint64_t range = cs->range;
                const uint64_t num = ((uint64_t)(cs->code - cs->low) + 1) * total - 1;
                const uint64_t value = div_fix(num, range, cs->inv_range);
                sym = 0;
#pragma unroll
                for (int l = 1; l < L; ++l) sym += (uint64_t)cums[l] <= value ? 1 : 0;
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    lo = l == sym ? cums[l] : lo;
                    hi = l == sym ? cums[l + 1] : hi;
                }
                if (lane == 0) {
                    cs->p_lo = lo;
                    cs->p_hi = hi;
                    cs->p_total = total;
                }
                bar_arrive(BAR_FOUND);
            }
            if (lane == 0) sym_img[at] = (uint8_t)sym;
#pragma unroll
            for (int l = 0; l < L; ++l) vnew = l == sym ? cent[l] : vnew;
            if (prof) {
                lp[1] += t2 - t1;
                lp[2] += clock64() - t2;
            }
        }
        if (prof) lp[0] += t1 - t0;
        vleft = vnew;
        acc1 = n1;
        acc2 = n2;
        acc3 = n3;
        p0 = p0n;
    }
}

template <int L>
__global__ void __launch_bounds__(NT, 1) pc_seq_decode_kernel(const DecArgs a) {
    extern __shared__ __align__(16) float smem[];
    Smem sm;
    sm.w1 = smem;
    sm.w2 = sm.w1 + 14 * 24 * 24;
    sm.w3 = sm.w2 + 14 * 24 * 24;
    sm.w0 = sm.w3 + 14 * 24 * 8;
    sm.b = sm.w0 + 13 * 24;
    sm.cent = sm.b + PROW;
    sm.bc = sm.cent + 8;
    sm.P = sm.bc + 144;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 14 * 24 * 24; i += NT) {
        sm.w1[i] = a.w1[i];
        sm.w2[i] = a.w2[i];
    }
    for (int i = tid; i < 14 * 24 * 8; i += NT) {
        const int o = i & 7;
        sm.w3[i] = o < L ? a.w3[(i >> 3) * L + o] : 0.f;
    }
    for (int i = tid; i < 13 * 24; i += NT) sm.w0[i] = a.w0[i];
    if (tid < 24) {
        sm.b[tid] = a.b0[tid];
        sm.b[24 + tid] = a.b1[tid];
        sm.b[48 + tid] = a.b2[tid];
    }
    if (tid < 8) {
        sm.b[72 + tid] = tid < L ? a.b3[tid] : 0.f;
        sm.cent[tid] = a.centers[tid];
    }

    Geom g;
    g.C = a.C; g.h = a.h; g.w = a.w; g.H6 = a.h + 6; g.W6 = a.w + 6;
    g.slot = (size_t)g.H6 * g.W6 * KC;
    g.layer = 2 * g.slot;
    for (int i = tid; i < PROW; i += NT) sm.P[(size_t)g.W6 * PROW + i] = 0.f;     // spare row read by the last step
    const int n = blockIdx.x;
    const size_t vol = (size_t)a.C * a.h * a.w;
    uint8_t* sym_img = a.sym_out + n * vol;
    const uint8_t* force_img = a.force_sym ? a.force_sym + n * vol : nullptr;
    const uint8_t* syms = force_img ? force_img : sym_img;      // what the chains read back
    float* act_img = a.act + (size_t)n * 3 * g.layer;
    float* p9 = a.p9 + (size_t)n * g.W6 * PROW;
    int64_t* freqs_img = a.freqs_out ? a.freqs_out + n * vol * L : nullptr;
    // stream -> word aligned, zero padded scratch (the coder never tests for the end of the stream)
    const int64_t nbytes = a.stream_off[n + 1] - a.stream_off[n];
    const int nwords = (int)min((int64_t)a.words_img, (nbytes + 3) / 4 + 8);
    {
        const uint8_t* sp = a.stream + a.stream_off[n];
        uint32_t* wd = a.words + (size_t)n * a.words_img;
        for (int i = tid; i < nwords; i += NT) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) v = (v << 8) | ((int64_t)i * 4 + k < nbytes ? (uint32_t)sp[(int64_t)i * 4 + k] : 0u);
            wd[i] = v;
        }
    }
    __syncthreads();
    CoderShared* cs = reinterpret_cast<CoderShared*>(sm.bc + 128);
    Coder cd;
    cd.init(a.words + (size_t)n * a.words_img, nwords);
    const bool coding = !a.force_sym && vol > 1;       // at least one symbol goes through the range decoder
    if (warp == 1 && coding) {
        cd.publish(cs, lane);
        bar_arrive(BAR_STATE);
    }
    const int first_sym = a.first_sym[n];

    const bool prof = a.prof && n == 0 && tid == 0;
    long long lpv[3] = {0, 0, 0}, rows[3] = {0, 0, 0};
    long long* lp = (a.prof && n == 0) ? lpv : nullptr;
    const long long tk0 = prof ? clock64() : 0;
    const int R = (a.C + 3) * g.H6;
    for (int r = 0; r < R; ++r) {
        const int c = r / g.H6 - 3, y = r % g.H6 - 3;
        const long long ta = prof ? clock64() : 0;
        __syncthreads();            // row r-1 is decoded and stored; helper chains of row r are in p9
        const long long tb = prof ? clock64() : 0;
        if (y == -3) {              // first row of a channel: nobody could run ahead
            chains<true>(g, sm, tid, NT, c, y, syms, act_img, p9);
            __syncthreads();
        }
        chains<false>(g, sm, tid, NT, c, y, syms, act_img, p9);
        __syncthreads();
        const long long tc = prof ? clock64() : 0;
        if (warp == 0)
            decode_row<L>(g, sm, a, lane, c, y, cs, sym_img, force_img, act_img, freqs_img, first_sym, lp);
        else if (warp == 1) {
            if (coding) coder_row(g, lane, c, y, cd, cs);
        } else if (y + 1 <= a.h + 2)
            chains<true>(g, sm, tid - 64, NT - 64, c, y + 1, syms, act_img, p9);
        if (prof) {
            rows[0] += tb - ta;               // waiting for the helper warps
            rows[1] += tc - tb;               // chains at row start
            rows[2] += clock64() - tc;        // decode_row
        }
    }
    if (prof) {
        a.prof[0] = lpv[0]; a.prof[1] = lpv[1]; a.prof[2] = lpv[2];
        a.prof[3] = rows[0]; a.prof[4] = rows[1]; a.prof[5] = rows[2];
        a.prof[6] = clock64() - tk0;
    }
}

template <int L>
int launch_decode(const DecArgs& a, int N, size_t smem, cudaStream_t s) {
    IC_CHECK_CUDA(cudaFuncSetAttribute(pc_seq_decode_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pc_seq_decode_kernel<L><<<N, NT, smem, s>>>(a);
    IC_CHECK_LAUNCH();
    return IC_OK;
}

}  // namespace

// a range coder with total <= 2^30 + 2 spends at most ~30 bits on a symbol: 4 bytes per symbol + slack
static int stream_words(int C, int h, int w) { return C * h * w + 16; }

size_t pc_decode_workspace_bytes(int N, int C, int h, int w) {
    const size_t act = (size_t)3 * 2 * (h + 6) * (w + 6) * KC * sizeof(float);
    const size_t p9 = (size_t)(w + 6) * PROW * sizeof(float);
    const size_t words = (size_t)stream_words(C, h, w) * sizeof(uint32_t);
    return (size_t)N * (align_up(act, 256) + align_up(p9, 256) + align_up(words, 256)) + 1024;
}

int pc_decode(const PcWeights& w, const PcDecodeInput& in, void* ws, size_t ws_bytes, cudaStream_t s) {
    IC_REQUIRE(w.K == KC, IC_ERR_UNSUPPORTED, "pc_decode: arch_param__k = %d not built (24 is)", w.K);
    IC_REQUIRE(w.L >= 2 && w.L <= 8, IC_ERR_UNSUPPORTED, "pc_decode: num_centers = %d not built (2..8)", w.L);
    const size_t smem = ((size_t)SMEM_FIXED_FLOATS + (size_t)(in.w + 7) * PROW) * sizeof(float);
    IC_REQUIRE(smem <= 227 * 1024, IC_ERR_UNSUPPORTED, "pc_decode: latent width %d does not fit one CTA's shared memory (max %d)",
               in.w, (int)((227 * 1024 / sizeof(float) - SMEM_FIXED_FLOATS) / PROW) - 7);
    Arena ar(ws, ws_bytes);
    const size_t act_img = (size_t)3 * 2 * (in.h + 6) * (in.w + 6) * KC;
    float* act = ar.get<float>((size_t)in.N * act_img);
    float* p9 = ar.get<float>((size_t)in.N * (in.w + 6) * PROW);
    const int words_img = stream_words(in.C, in.h, in.w);
    uint32_t* words = ar.get<uint32_t>((size_t)in.N * words_img);
    IC_REQUIRE(ar.ok(), IC_ERR_WORKSPACE, "pc_decode workspace too small: need %zu, have %zu", ar.off, ws_bytes);
    DecArgs a;
    a.w0 = w.w0; a.b0 = w.b0; a.w1 = w.w1; a.b1 = w.b1; a.w2 = w.w2; a.b2 = w.b2; a.w3 = w.w3; a.b3 = w.b3;
    a.C = in.C; a.h = in.h; a.w = in.w; a.L = w.L;
    for (int i = 0; i < 8; ++i) a.centers[i] = i < w.L ? in.centers_host[i] : 0.f;
    a.stream = in.stream;
    a.stream_off = in.stream_off;
    a.first_sym = in.first_sym;
    a.sym_out = in.sym_out;
    a.act = act;
    a.p9 = p9;
    a.words = words;
    a.words_img = words_img;
    a.force_sym = in.force_sym;
    a.freqs_out = in.freqs_out;
    a.prof = nullptr;
    if (getenv("IC_PC_DECODE_PROF")) {
        IC_CHECK_CUDA(cudaMalloc((void**)&a.prof, 8 * sizeof(long long)));
        IC_CHECK_CUDA(cudaMemsetAsync(a.prof, 0, 8 * sizeof(long long), s));
    }
    int rc;
    {
        ProfScope ps(IC_PROF_PROBCLASS, s, 1);
        switch (w.L) {
            case 2: rc = launch_decode<2>(a, in.N, smem, s); break;
            case 3: rc = launch_decode<3>(a, in.N, smem, s); break;
            case 4: rc = launch_decode<4>(a, in.N, smem, s); break;
            case 5: rc = launch_decode<5>(a, in.N, smem, s); break;
            case 6: rc = launch_decode<6>(a, in.N, smem, s); break;
            case 7: rc = launch_decode<7>(a, in.N, smem, s); break;
            default: rc = launch_decode<8>(a, in.N, smem, s); break;
        }
    }
    if (a.prof) {
        long long h[8];
        IC_CHECK_CUDA(cudaStreamSynchronize(s));
        IC_CHECK_CUDA(cudaMemcpy(h, a.prof, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(a.prof);
        fprintf(stderr, "pc_decode cycles (image 0): total %lld | rows: helper wait %lld, row-start chains %lld, decode_row %lld | "
                        "steps: layers %lld, head %lld, range decode %lld\n", h[6], h[3], h[4], h[5], h[0], h[1], h[2]);
    }
    return rc;
}

}  // namespace ic
