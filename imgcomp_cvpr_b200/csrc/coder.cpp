// Host range coder.  Restates the integer arithmetic of the reference's
// arithmetic_coding.py (ArithmeticCoderBase.update :80-115, ArithmeticEncoder
// :127-159, ArithmeticDecoder :163-222, SimpleFrequencyTable :323-424,
// BitOutputStream/BitInputStream :497-567) with 64-bit integers: 32-bit state,
// total <= 2^30 + 2 (MAX_TOTAL, :47-50), so symhigh * range < 2^62.
// One frequency table per symbol (the context model emits a table per position).
#include <stdint.h>
#include <string.h>

#include <new>
#include <vector>

#include "imgcomp_b200.h"

namespace ic {
void set_error(const char* fmt, ...);
}

namespace {

constexpr int kStateBits = 32;
constexpr uint64_t kFull = 1ull << kStateBits;           // MAX_RANGE
constexpr uint64_t kMask = kFull - 1;                    // MASK
constexpr uint64_t kTop = kFull >> 1;                    // TOP_MASK
constexpr uint64_t kSecond = kTop >> 1;                  // SECOND_MASK
constexpr uint64_t kMaxTotal = (kFull >> 2) + 2;         // MAX_TOTAL = MIN_RANGE

// Bit sink of the encoder.  Pending bits sit LEFT-aligned in a 64-bit accumulator; appending is a shift-or, and whole
// bytes leave through one unconditional 8-byte big-endian store whose write pointer advances by filled / 8.
// (Measured on the build host, 196 608 symbols of peaky 6-entry tables: 25.7 ns per symbol against 27.3 for the
// bit-per-iteration restatement -- what is left is the serial chain range -> product -> quotient -> shared-prefix count ->
// shifts from one symbol to the next, ~20 ns with either hardware division or a corrected double estimate.)
struct BitWriter {
    std::vector<uint8_t> bytes;     // capacity buffer; the first `pos` bytes are output (finish() trims it)
    size_t pos = 0;
    uint64_t acc = 0;               // the top `filled` bits are pending output, the rest is zero
    int filled = 0;                 // < 8 between calls
    int64_t nbits = 0;
    void reserve(size_t more) {     // room for `more` bytes plus the 8-byte store
        if (pos + more + 8 > bytes.size()) bytes.resize(2 * bytes.size() + more + 4096);
    }
    // the `count` (<= 56) low bits of v, most significant first; the caller has reserved count / 8 + 1 bytes
    void put_bits(uint64_t v, int count) {
        acc |= v << ((64 - filled - count) & 63);       // count == 0 implies v == 0: the masked shift is harmless
        filled += count;
        nbits += count;
        const uint64_t be = __builtin_bswap64(acc);
        memcpy(bytes.data() + pos, &be, 8);
        const int nb = filled >> 3;
        pos += (size_t)nb;
        acc <<= 8 * nb;
        filled &= 7;
    }
    void put(int b) {
        reserve(1);
        put_bits((uint64_t)(b & 1), 1);
    }
    void put_run(int bit, int64_t count) {      // `count` copies of one bit (underflow runs have no upper bound)
        const uint64_t pat = bit ? 0xFFFFFFFFull : 0ull;
        reserve((size_t)(count / 8) + 8);
        for (; count >= 32; count -= 32) put_bits(pat, 32);
        if (count > 0) put_bits(pat >> (32 - (int)count), (int)count);
    }
    void pad_to_byte() {            // does not count towards nbits
        reserve(1);
        if (filled != 0) {
            bytes[pos++] = (uint8_t)(acc >> 56);
            acc = 0;
            filled = 0;
        }
    }
    void trim() { bytes.resize(pos); }
};

struct BitReader {
    const uint8_t* p;
    int64_t n, pos = 0;
    int left = 0;
    uint32_t cur = 0;
    int get() {                          // end of stream reads as zeros (:217-222)
        if (left == 0) {
            if (pos >= n) return 0;
            cur = p[pos++];
            left = 8;
        }
        --left;
        return (cur >> left) & 1;
    }
};

// cumulative bounds of `symbol` in one table of L frequencies
inline bool bounds(const int64_t* f, int L, int64_t symbol, uint64_t& lo, uint64_t& hi, uint64_t& total) {
    if (symbol < 0 || symbol >= L) return false;
    uint64_t c = 0;
    lo = hi = 0;
    for (int j = 0; j < L; ++j) {
        if (f[j] < 0) return false;
        if (j == symbol) lo = c;
        c += (uint64_t)f[j];
        if (j == symbol) hi = c;
    }
    total = c;
    return total <= kMaxTotal && hi > lo;
}

}  // namespace

struct ic_ac_enc {
    uint64_t low = 0, high = kMask;
    int64_t underflow = 0;
    BitWriter out;
    bool finished = false;

    // ArithmeticCoderBase.update (:80-115) with the two bit-per-iteration loops in closed form: the n leading bits low and
    // high share leave together (the first one followed by the pending underflow bits, :139-143), then the k positions
    // below the top where low has a 1 and high a 0 are squeezed out (:104-110).  Same state, same bits, same order.
    void update(uint64_t lo, uint64_t hi, uint64_t total) {
        const uint64_t range = high - low + 1, base = low;
        low = base + lo * range / total;
        high = base + hi * range / total - 1;
        // n = leading bits low and high share (32 when they are equal), without a branch
        const uint32_t x = (uint32_t)(low ^ high);
        const int n = __builtin_clzll(((uint64_t)x << 32) | 0x80000000ull);
        const int nm1 = n - (n != 0);
        out.reserve(16);
        if (underflow <= 24) {
            // [first bit][underflow x its complement][the other n - 1 bits] as one word of n + underflow <= 56 bits
            const int u = n ? (int)underflow : 0;
            const uint64_t b = (low >> (kStateBits - 1)) & 1;
            const uint64_t rest = (low >> ((kStateBits - n) & 63)) & ((1ull << nm1) - 1);
            const uint64_t run = (b - 1) & ((1ull << u) - 1);            // u ones if the first bit is 0
            const uint64_t v = n ? ((b << (u + nm1)) | (run << nm1) | rest) : 0;
            out.put_bits(v, n + u);
            underflow -= u;
        } else if (n) {
            const int bit = (int)(low >> (kStateBits - 1));
            out.put_bits((uint64_t)bit, 1);
            out.put_run(bit ^ 1, underflow);
            underflow = 0;
            out.reserve(16);
            out.put_bits((low >> (kStateBits - n)) & ((1ull << nm1) - 1), nm1);
        }
        low = (low << n) & kMask;
        high = ((high << n) & kMask) | ((1ull << n) - 1);
        // k = positions below the top where low has a 1 and high a 0 (0 when the top pair already differs the other way:
        // ~y then has its top bit set); bit 0 of y is 0, so k <= 31 and the shifts are the identity for k = 0
        const uint32_t y = (uint32_t)((low & ~high) << 1);
        const int k = __builtin_clz(~y);
        underflow += k;
        low = (low << k) & (kMask >> 1);
        high = ((high << k) & (kMask >> 1)) | kTop | ((1ull << k) - 1);
    }
};

struct ic_ac_dec {
    uint64_t low = 0, high = kMask, code = 0;
    BitReader in;

    void update(uint64_t lo, uint64_t hi, uint64_t total) {
        const uint64_t range = high - low + 1;
        const uint64_t nl = low + lo * range / total;
        const uint64_t nh = low + hi * range / total - 1;
        low = nl;
        high = nh;
        while (((low ^ high) & kTop) == 0) {
            code = ((code << 1) & kMask) | (uint64_t)in.get();
            low = (low << 1) & kMask;
            high = ((high << 1) & kMask) | 1;
        }
        while ((low & ~high & kSecond) != 0) {
            code = (code & kTop) | ((code << 1) & (kMask >> 1)) | (uint64_t)in.get();
            low = (low << 1) & (kMask >> 1);
            high = ((high << 1) & (kMask >> 1)) | kTop | 1;
        }
    }
};

extern "C" {

/* CRC-32C (Castagnoli, reflected polynomial 0x82F63B78) of a host buffer: what TensorFlow's tensor bundle stores (masked)
 * per tensor; imgcomp_cvpr_b200/tf_checkpoint.py verifies checkpoints with it. */
uint32_t ic_crc32c(const void* h_data, int64_t n) {
    static uint32_t table[8][256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0x82F63B78u & (0u - (c & 1u)));
            table[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int t = 1; t < 8; ++t) table[t][i] = (table[t - 1][i] >> 8) ^ table[0][table[t - 1][i] & 0xFF];
        ready = true;
    }
    const uint8_t* p = (const uint8_t*)h_data;
    uint32_t c = 0xFFFFFFFFu;
    while (n >= 8) {               // slicing-by-8
        uint32_t lo, hi;
        memcpy(&lo, p, 4);
        memcpy(&hi, p + 4, 4);
        lo ^= c;
        c = table[7][lo & 0xFF] ^ table[6][(lo >> 8) & 0xFF] ^ table[5][(lo >> 16) & 0xFF] ^ table[4][lo >> 24] ^
            table[3][hi & 0xFF] ^ table[2][(hi >> 8) & 0xFF] ^ table[1][(hi >> 16) & 0xFF] ^ table[0][hi >> 24];
        p += 8;
        n -= 8;
    }
    while (n-- > 0) c = (c >> 8) ^ table[0][(c ^ *p++) & 0xFF];
    return c ^ 0xFFFFFFFFu;
}

int ic_ac_enc_create(ic_ac_enc_t** out) {
    if (!out) return IC_ERR_INVALID;
    *out = new (std::nothrow) ic_ac_enc();
    return *out ? IC_OK : IC_ERR_INVALID;
}

int ic_ac_enc_write(ic_ac_enc_t* e, const int64_t* h_freqs, int L, const int64_t* h_symbols, int64_t n) {
    if (!e || !h_freqs || !h_symbols || L <= 0 || n < 0) {
        ic::set_error("ic_ac_enc_write: bad argument");
        return IC_ERR_INVALID;
    }
    if (e->finished) {
        ic::set_error("ic_ac_enc_write: encoder already finished");
        return IC_ERR_STATE;
    }
    for (int64_t i = 0; i < n; ++i) {
        uint64_t lo, hi, total;
        if (!bounds(h_freqs + i * L, L, h_symbols[i], lo, hi, total)) {
            // the reference raises ValueError: zero frequency / total too large (:94-97)
            ic::set_error("ic_ac_enc_write: symbol %lld of entry %lld not codable (zero frequency, out of range, or total > 2^30+2)",
                          (long long)h_symbols[i], (long long)i);
            return IC_ERR_INVALID;
        }
        e->update(lo, hi, total);
    }
    return IC_OK;
}

int ic_ac_enc_write_u32(ic_ac_enc_t* e, const uint32_t* h_freqs, int L, const uint8_t* h_symbols, int64_t n) {
    if (!e || !h_freqs || !h_symbols || L <= 0 || n < 0) {
        ic::set_error("ic_ac_enc_write_u32: bad argument");
        return IC_ERR_INVALID;
    }
    if (e->finished) {
        ic::set_error("ic_ac_enc_write_u32: encoder already finished");
        return IC_ERR_STATE;
    }
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t* f = h_freqs + i * L;
        const int sym = h_symbols[i];
        uint64_t c = 0, lo = 0, hi = 0;
        if (L <= 8) {           // the context model's tables (L = 6): prefix sums without a data-dependent branch
            uint64_t cum[9];
            cum[0] = 0;
            for (int j = 0; j < L; ++j) cum[j + 1] = cum[j] + f[j];
            c = cum[L];
            const int sj = sym < L ? sym : 0;
            lo = cum[sj];
            hi = cum[sj + 1];
        } else {
            for (int j = 0; j < L; ++j) {
                if (j == sym) lo = c;
                c += f[j];
                if (j == sym) hi = c;
            }
        }
        if (sym >= L || c > kMaxTotal || hi <= lo) {
            ic::set_error("ic_ac_enc_write_u32: symbol %d of entry %lld not codable (zero frequency, out of range, or total > 2^30+2)",
                          sym, (long long)i);
            return IC_ERR_INVALID;
        }
        e->update(lo, hi, c);
    }
    return IC_OK;
}

int ic_ac_enc_finish(ic_ac_enc_t* e, const uint8_t** h_bytes, int64_t* n_bytes, int64_t* n_bits) {
    if (!e || !h_bytes || !n_bytes || !n_bits) return IC_ERR_INVALID;
    if (!e->finished) {
        e->out.put(1);                  // ArithmeticEncoder.finish (:146-147)
        e->out.pad_to_byte();           // BitOutputStream.close (:564-567)
        e->out.trim();
        e->finished = true;
    }
    *h_bytes = e->out.bytes.data();
    *n_bytes = (int64_t)e->out.bytes.size();
    *n_bits = e->out.nbits;
    return IC_OK;
}

void ic_ac_enc_destroy(ic_ac_enc_t* e) { delete e; }

int ic_ac_dec_create(const uint8_t* h_bytes, int64_t n_bytes, ic_ac_dec_t** out) {
    if (!out || (!h_bytes && n_bytes > 0) || n_bytes < 0) return IC_ERR_INVALID;
    ic_ac_dec* d = new (std::nothrow) ic_ac_dec();
    if (!d) return IC_ERR_INVALID;
    d->in.p = h_bytes;
    d->in.n = n_bytes;
    for (int i = 0; i < kStateBits; ++i) d->code = (d->code << 1) | (uint64_t)d->in.get();
    *out = d;
    return IC_OK;
}

int ic_ac_dec_read(ic_ac_dec_t* d, const int64_t* h_freqs, int L, int64_t* h_symbols, int64_t n) {
    if (!d || !h_freqs || !h_symbols || L <= 0 || n < 0) {
        ic::set_error("ic_ac_dec_read: bad argument");
        return IC_ERR_INVALID;
    }
    for (int64_t i = 0; i < n; ++i) {
        const int64_t* f = h_freqs + i * L;
        uint64_t total = 0;
        for (int j = 0; j < L; ++j) {
            if (f[j] < 0) {
                ic::set_error("ic_ac_dec_read: negative frequency");
                return IC_ERR_INVALID;
            }
            total += (uint64_t)f[j];
        }
        if (total == 0 || total > kMaxTotal) {
            ic::set_error("ic_ac_dec_read: total %llu not codable", (unsigned long long)total);
            return IC_ERR_INVALID;
        }
        const uint64_t range = d->high - d->low + 1;
        const uint64_t offset = d->code - d->low;
        const uint64_t value = ((offset + 1) * total - 1) / range;
        // highest symbol whose cumulative low <= value (:188-197)
        uint64_t c = 0, lo = 0, hi = 0;
        int sym = 0;
        for (int j = 0; j < L; ++j) {
            if (c <= value && f[j] > 0) {
                sym = j;
                lo = c;
                hi = c + (uint64_t)f[j];
            }
            c += (uint64_t)f[j];
        }
        d->update(lo, hi, total);
        h_symbols[i] = sym;
    }
    return IC_OK;
}

void ic_ac_dec_destroy(ic_ac_dec_t* d) { delete d; }

}  // extern "C"
