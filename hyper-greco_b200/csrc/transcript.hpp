// Host-side Fiat-Shamir transcript, byte-exact with the reference's `Keccak256Transcript`
// (/root/reference/bfv-gkr/src/transcript.rs:117-203). Stays on the host (north star).
//
// Reference behaviour reproduced on purpose:
//   * squeeze_challenge: hash = finalize_fixed_reset(); update(hash); fe_mod_from_le_bytes(hash)   (transcript.rs:199-203)
//     -> the challenge stream is the chain keccak^{i+1}("") and never depends on what the prover wrote;
//   * common_felt is a no-op                                                                      (transcript.rs:156)
//   * write_felt appends to_repr() REVERSED (big-endian) to the proof stream only                 (transcript.rs:183-189)
//   * an extension element is written/squeezed as its DEGREE base coordinates in order            (transcript.rs:149-154,191-195)
// Keccak256 is original Keccak (0x01 padding, rate 136), as in plonkish_backend::util::hash::Keccak256.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace hg {

class Keccak256 {
  public:
    Keccak256() { reset(); }
    void reset() { memset(a_, 0, sizeof a_); fill_ = 0; }
    void update(const uint8_t* p, size_t n) {
        uint8_t* s = reinterpret_cast<uint8_t*>(a_);
        while (n) {
            size_t take = kRate - fill_ < n ? kRate - fill_ : n;
            for (size_t i = 0; i < take; i++) s[fill_ + i] ^= p[i];
            fill_ += take; p += take; n -= take;
            if (fill_ == kRate) { permute(); fill_ = 0; }
        }
    }
    // digest of everything absorbed so far, then reset (digest::FixedOutputReset)
    void finalize_reset(uint8_t out[32]) {
        uint8_t* s = reinterpret_cast<uint8_t*>(a_);
        s[fill_] ^= 0x01;
        s[kRate - 1] ^= 0x80;
        permute();
        memcpy(out, s, 32);
        reset();
    }

  private:
    static constexpr size_t kRate = 136;
    uint64_t a_[25];
    size_t fill_;
    static inline uint64_t rol(uint64_t v, unsigned s) { return (v << s) | (v >> (64 - s)); }
    // round constants from the degree-8 LFSR x^8+x^6+x^5+x^4+1, rho offsets from the (x,y) -> (y, 2x+3y) orbit; computed once
    struct Tables {
        uint64_t rc[24];
        unsigned rot[24];
        int lane[24];
        Tables() {
            uint8_t lfsr = 1;
            for (int round = 0; round < 24; round++) {
                uint64_t c = 0;
                for (int j = 0; j < 7; j++) {
                    if (lfsr & 1) c ^= 1ULL << ((1 << j) - 1);
                    lfsr = (lfsr & 0x80) ? (uint8_t)((lfsr << 1) ^ 0x71) : (uint8_t)(lfsr << 1);
                }
                rc[round] = c;
            }
            int x = 1, y = 0;
            for (int t = 0; t < 24; t++) {
                rot[t] = ((t + 1) * (t + 2) / 2) % 64;
                int ny = (2 * x + 3 * y) % 5;
                x = y; y = ny;
                lane[t] = x + 5 * y;
            }
        }
    };
    static const Tables& tables() { static const Tables t; return t; }
    void permute() {
        const Tables& T = tables();
        uint64_t* a = a_;
        for (int round = 0; round < 24; round++) {
            uint64_t c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20], c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21],
                     c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22], c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23],
                     c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
            uint64_t d0 = c4 ^ rol(c1, 1), d1 = c0 ^ rol(c2, 1), d2 = c1 ^ rol(c3, 1), d3 = c2 ^ rol(c4, 1), d4 = c3 ^ rol(c0, 1);
            for (int y = 0; y < 25; y += 5) { a[y] ^= d0; a[y + 1] ^= d1; a[y + 2] ^= d2; a[y + 3] ^= d3; a[y + 4] ^= d4; }
            uint64_t cur = a[1];
            for (int t = 0; t < 24; t++) {
                uint64_t nxt = a[T.lane[t]];
                a[T.lane[t]] = rol(cur, T.rot[t]);
                cur = nxt;
            }
            for (int y = 0; y < 25; y += 5) {
                uint64_t r0 = a[y], r1 = a[y + 1], r2 = a[y + 2], r3 = a[y + 3], r4 = a[y + 4];
                a[y] = r0 ^ (~r1 & r2); a[y + 1] = r1 ^ (~r2 & r3); a[y + 2] = r2 ^ (~r3 & r4); a[y + 3] = r3 ^ (~r4 & r0); a[y + 4] = r4 ^ (~r0 & r1);
            }
            a[0] ^= T.rc[round];
        }
    }
};

struct TranscriptError : std::runtime_error { using std::runtime_error::runtime_error; };

// HF: host field traits (see host_field.hpp): Base, Ext, base_from_le_bytes_mod, base_to_repr_le, base_from_repr_le ...
template <class HF> class Keccak256Transcript {
  public:
    typedef typename HF::Base Base;
    typedef typename HF::Ext Ext;
    Keccak256Transcript() {}                                                  // Keccak256Transcript::<Vec<u8>>::default()
    Keccak256Transcript(const uint8_t* proof, size_t n) : rd_(proof, proof + n), reading_(true) {}  // from_proof

    Base squeeze_base() {
        uint8_t h[32];
        state_.finalize_reset(h);
        state_.update(h, 32);
        n_squeezed_++;
        return HF::base_from_le_bytes_mod(h, 32);
    }
    Ext squeeze_challenge() {
        Base b[HF::DEGREE];
        for (int i = 0; i < HF::DEGREE; i++) b[i] = squeeze_base();
        return HF::ext_from_bases(b);
    }
    void common_felt(const Base&) {}
    void write_felt(const Base& f) {
        uint8_t b[HF::REPR_BYTES];
        HF::base_to_repr_le(f, b);
        for (int i = HF::REPR_BYTES - 1; i >= 0; i--) stream_.push_back(b[i]);
    }
    void write_felt_ext(const Ext& e) {
        Base b[HF::DEGREE];
        HF::ext_as_bases(e, b);
        for (int i = 0; i < HF::DEGREE; i++) write_felt(b[i]);
    }
    Base read_felt() {
        if (pos_ + HF::REPR_BYTES > rd_.size()) throw TranscriptError("failed to fill whole buffer");
        uint8_t b[HF::REPR_BYTES];
        for (int i = 0; i < HF::REPR_BYTES; i++) b[HF::REPR_BYTES - 1 - i] = rd_[pos_ + i];
        pos_ += HF::REPR_BYTES;
        Base f;
        if (!HF::base_from_repr_le(b, &f)) throw TranscriptError("Invalid field element read from stream");
        return f;
    }
    Ext read_felt_ext() {
        Base b[HF::DEGREE];
        for (int i = 0; i < HF::DEGREE; i++) b[i] = read_felt();
        return HF::ext_from_bases(b);
    }
    const std::vector<uint8_t>& proof() const { return stream_; }  // into_proof
    size_t num_base_squeezed() const { return n_squeezed_; }
    size_t read_pos() const { return pos_; }

  private:
    Keccak256 state_;
    std::vector<uint8_t> stream_, rd_;
    size_t pos_ = 0, n_squeezed_ = 0;
    bool reading_ = false;
};

}  // namespace hg
