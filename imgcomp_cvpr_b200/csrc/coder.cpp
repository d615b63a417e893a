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

struct BitWriter {
    std::vector<uint8_t> bytes;
    uint32_t cur = 0;
    int filled = 0;
    int64_t nbits = 0;
    void put(int b) {
        cur = (cur << 1) | (uint32_t)b;
        ++nbits;
        if (++filled == 8) {
            bytes.push_back((uint8_t)cur);
            cur = 0;
            filled = 0;
        }
    }
    void pad_to_byte() {
        while (filled != 0) {           // does not count towards nbits
            cur <<= 1;
            if (++filled == 8) {
                bytes.push_back((uint8_t)cur);
                cur = 0;
                filled = 0;
            }
        }
    }
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

    void update(uint64_t lo, uint64_t hi, uint64_t total) {
        const uint64_t range = high - low + 1;
        const uint64_t nl = low + lo * range / total;
        const uint64_t nh = low + hi * range / total - 1;
        low = nl;
        high = nh;
        while (((low ^ high) & kTop) == 0) {
            const int bit = (int)(low >> (kStateBits - 1));
            out.put(bit);
            for (; underflow > 0; --underflow) out.put(bit ^ 1);
            low = (low << 1) & kMask;
            high = ((high << 1) & kMask) | 1;
        }
        while ((low & ~high & kSecond) != 0) {
            ++underflow;
            low = (low << 1) & (kMask >> 1);
            high = ((high << 1) & (kMask >> 1)) | kTop | 1;
        }
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
        for (int j = 0; j < L; ++j) {
            if (j == sym) lo = c;
            c += f[j];
            if (j == sym) hi = c;
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
