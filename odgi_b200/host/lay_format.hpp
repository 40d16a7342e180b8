// lay_format.hpp — odgi's binary 2D layout container (`.lay`), written and read without sdsl.
//
// The reference stores a layout as (src/algorithms/layout.cpp:43-66, layout.hpp:27-30)
//     double min_value                                   the smallest coordinate of X and Y
//     sdsl::enc_vector<> xy                              over the bit patterns of (X[i]-min_value), (Y[i]-min_value), interleaved
// and sdsl::enc_vector<coder::elias_delta, 128, 0> serialises as (deps/sdsl-lite/include/sdsl/enc_vector.hpp:216-283, 341-350)
//     u64 size                                           number of values
//     int_vector<0> z   {u64 bits, u8 width = 1, words}  Elias-delta codes of v[i] - v[i-1] (mod 2^64), none for sampled i
//     int_vector<0> sp  {u64 bits, u8 width = W, words}  per 128 values: (absolute value, bit offset into z); then (0, |z|+1)
// with bits packed LSB-first into 64-bit words (bits.hpp:478-499) and a delta of 0 coded as 2^64
// (coder_elias_delta.hpp:151-155, 198-212).  This header restates that layout so that `pgsgd layout -o g.lay` produces
// files `odgi draw` reads, byte-identical to what `odgi layout -o` writes for the same coordinates
// (tests/test_host_cpu.py compares against files written by the reference itself).
#pragma once

#include <cstdint>
#include <cstring>
#include <istream>
#include <limits>
#include <ostream>
#include <stdexcept>
#include <vector>

namespace pgsgd {
namespace lay {

inline unsigned hi_bit(uint64_t x) { return 63u - (unsigned) __builtin_clzll(x); }  // position of the highest set bit, x > 0

// LSB-first bit sink over 64-bit words
class BitWriter {
public:
    explicit BitWriter(uint64_t n_bits) : words_((n_bits + 63) / 64, 0) {}
    void put(uint64_t x, unsigned len) {  // the low `len` bits of x, 0 <= len <= 64
        if (len == 0) return;
        if (len < 64) x &= (1ULL << len) - 1;
        const uint64_t w = pos_ >> 6;
        const unsigned off = (unsigned) (pos_ & 63);
        words_[w] |= x << off;
        if (off + len > 64) words_[w + 1] |= x >> (64 - off);
        pos_ += len;
    }
    uint64_t bits() const { return pos_; }
    const std::vector<uint64_t>& words() const { return words_; }
private:
    std::vector<uint64_t> words_;
    uint64_t pos_ = 0;
};

inline unsigned elias_delta_length(uint64_t w) {
    const unsigned len_1 = w ? hi_bit(w) : 64;
    return len_1 + (hi_bit(len_1 + 1) << 1) + 1;
}

inline void elias_delta_put(BitWriter& out, uint64_t x) {
    const unsigned len = x ? hi_bit(x) + 1 : 65;  // 0 stands for 2^64
    const unsigned len_1_len = hi_bit(len);
    out.put(1ULL << len_1_len, len_1_len + 1);     // unary length of the length
    if (len_1_len) {
        out.put(len, len_1_len);                    // the length without its top bit
        out.put(x, len - 1);                        // the value without its top bit
    }
}

inline void write_u64(std::ostream& out, uint64_t v) { out.write(reinterpret_cast<const char*>(&v), 8); }

inline void write_int_vector(std::ostream& out, uint64_t n_bits, uint8_t width, const std::vector<uint64_t>& words) {
    write_u64(out, n_bits);
    out.write(reinterpret_cast<const char*>(&width), 1);
    out.write(reinterpret_cast<const char*>(words.data()), (std::streamsize) (((n_bits + 63) / 64) * 8));
}

constexpr uint64_t SAMPLE_DENS = 128;

// sdsl::enc_vector<>(vals).serialize(out)
inline void write_enc_vector(std::ostream& out, const std::vector<uint64_t>& vals) {
    if (vals.empty()) throw std::runtime_error("lay: empty layout");
    uint64_t z_bits = 0, max_sample = 0, samples = 0;
    for (uint64_t i = 0; i < vals.size(); ++i) {
        if (i % SAMPLE_DENS == 0) { if (vals[i] > max_sample) max_sample = vals[i]; ++samples; }
        else z_bits += elias_delta_length(vals[i] - vals[i - 1]);
    }
    const uint8_t width = (uint8_t) (hi_bit(max_sample > z_bits + 1 ? max_sample : z_bits + 1) + 1);
    BitWriter z(z_bits), sp((2 * samples + 2) * width);
    for (uint64_t i = 0; i < vals.size(); ++i) {
        if (i % SAMPLE_DENS == 0) { sp.put(vals[i], width); sp.put(z.bits(), width); }
        else elias_delta_put(z, vals[i] - vals[i - 1]);
    }
    sp.put(0, width);
    sp.put(z_bits + 1, width);
    write_u64(out, vals.size());
    write_int_vector(out, z_bits, 1, z.words());
    write_int_vector(out, (2 * samples + 2) * width, width, sp.words());
}

// algorithms::layout::Layout(X, Y).serialize(out)
inline void write_lay(std::ostream& out, const std::vector<double>& X, const std::vector<double>& Y) {
    if (X.size() != Y.size()) throw std::runtime_error("lay: X and Y differ in length");
    double min_value = std::numeric_limits<double>::max();
    for (double v : X) min_value = v < min_value ? v : min_value;
    for (double v : Y) min_value = v < min_value ? v : min_value;
    std::vector<uint64_t> vals;
    vals.reserve(2 * X.size());
    for (uint64_t i = 0; i < X.size(); ++i) {
        const double x = X[i] - min_value, y = Y[i] - min_value;
        uint64_t bx, by;
        std::memcpy(&bx, &x, 8);
        std::memcpy(&by, &y, 8);
        vals.push_back(bx);
        vals.push_back(by);
    }
    out.write(reinterpret_cast<const char*>(&min_value), 8);
    write_enc_vector(out, vals);
}

// ---- reading (Layout::load + get_X / get_Y) ----------------------------------------------------------------------
class BitReader {
public:
    BitReader(const std::vector<uint64_t>& words, uint64_t n_bits) : w_(words), n_(n_bits) {}
    uint64_t get(uint64_t pos, unsigned len) const {  // len <= 64
        if (len == 0) return 0;
        if (pos + len > n_) throw std::runtime_error("lay: bit stream truncated");
        const uint64_t w = pos >> 6;
        const unsigned off = (unsigned) (pos & 63);
        uint64_t x = w_[w] >> off;
        if (off + len > 64) x |= w_[w + 1] << (64 - off);
        return len < 64 ? x & ((1ULL << len) - 1) : x;
    }
private:
    const std::vector<uint64_t>& w_;
    uint64_t n_;
};

inline uint64_t read_u64(std::istream& in) {
    uint64_t v = 0;
    in.read(reinterpret_cast<char*>(&v), 8);
    if (!in) throw std::runtime_error("lay: file truncated");
    return v;
}

inline void read_int_vector(std::istream& in, uint64_t& n_bits, uint8_t& width, std::vector<uint64_t>& words) {
    n_bits = read_u64(in);
    in.read(reinterpret_cast<char*>(&width), 1);
    if (n_bits > (1ULL << 46)) throw std::runtime_error("lay: implausible vector size");
    words.assign((n_bits + 63) / 64 + 1, 0);  // one spare word: get() may touch w + 1
    in.read(reinterpret_cast<char*>(words.data()), (std::streamsize) (((n_bits + 63) / 64) * 8));
    if (!in) throw std::runtime_error("lay: file truncated");
}

inline std::vector<uint64_t> read_enc_vector(std::istream& in) {
    const uint64_t n = read_u64(in);
    uint64_t z_bits, sp_bits;
    uint8_t z_width, width;
    std::vector<uint64_t> zw, spw;
    read_int_vector(in, z_bits, z_width, zw);
    read_int_vector(in, sp_bits, width, spw);
    if (n == 0) return {};
    if (z_width != 1 || width == 0 || width > 64) throw std::runtime_error("lay: not an enc_vector<elias_delta,128>");
    const uint64_t samples = (n + SAMPLE_DENS - 1) / SAMPLE_DENS;
    if (sp_bits != (2 * samples + 2) * width) throw std::runtime_error("lay: sample table does not match the value count");
    BitReader z(zw, z_bits), sp(spw, sp_bits);
    std::vector<uint64_t> vals(n);
    uint64_t pos = 0, v = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (i % SAMPLE_DENS == 0) {
            v = sp.get((2 * (i / SAMPLE_DENS)) * width, width);
            pos = sp.get((2 * (i / SAMPLE_DENS) + 1) * width, width);
        } else {
            unsigned len_1_len = 0;
            while (z.get(pos, 1) == 0) { ++pos; ++len_1_len; if (len_1_len > 6) throw std::runtime_error("lay: corrupt Elias-delta code"); }
            ++pos;
            uint64_t delta = 1;
            if (len_1_len) {
                const unsigned len = (unsigned) (z.get(pos, len_1_len) | (1ULL << len_1_len));
                pos += len_1_len;
                if (len > 65) throw std::runtime_error("lay: corrupt Elias-delta code");
                const uint64_t low = z.get(pos, len - 1);
                pos += len - 1;
                delta = len == 65 ? 0 : (low | (1ULL << (len - 1)));  // 2^64 == 0 (mod 2^64)
            }
            v += delta;
        }
        vals[i] = v;
    }
    return vals;
}

inline void read_lay(std::istream& in, std::vector<double>& X, std::vector<double>& Y) {
    double min_value = 0;
    in.read(reinterpret_cast<char*>(&min_value), 8);
    if (!in) throw std::runtime_error("lay: file truncated");
    const std::vector<uint64_t> vals = read_enc_vector(in);
    if (vals.size() % 2) throw std::runtime_error("lay: odd number of values");
    X.resize(vals.size() / 2);
    Y.resize(vals.size() / 2);
    for (uint64_t i = 0; i < X.size(); ++i) {
        double x, y;
        std::memcpy(&x, &vals[2 * i], 8);
        std::memcpy(&y, &vals[2 * i + 1], 8);
        X[i] = x + min_value;  // Layout::get_x (layout.cpp:86-90)
        Y[i] = y + min_value;
    }
}

}  // namespace lay
}  // namespace pgsgd
