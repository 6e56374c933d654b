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
        while (n--) {
            reinterpret_cast<uint8_t*>(a_)[fill_++] ^= *p++;
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
    static uint64_t rol(uint64_t v, unsigned s) { return (v << s) | (v >> ((64 - s) & 63)); }
    void permute() {
        uint64_t rc = 1;  // round constants from the degree-8 LFSR x^8+x^6+x^5+x^4+1
        uint8_t lfsr = 1;
        for (int round = 0; round < 24; round++) {
            uint64_t c[5];
            for (int x = 0; x < 5; x++) c[x] = a_[x] ^ a_[x + 5] ^ a_[x + 10] ^ a_[x + 15] ^ a_[x + 20];
            for (int x = 0; x < 5; x++) {
                uint64_t d = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
                for (int y = 0; y < 25; y += 5) a_[y + x] ^= d;
            }
            // rho + pi, walking the (x,y) -> (y, 2x+3y) orbit
            int x = 1, y = 0;
            uint64_t cur = a_[1];
            for (int t = 0; t < 24; t++) {
                unsigned r = ((t + 1) * (t + 2) / 2) % 64;
                int ny = (2 * x + 3 * y) % 5;
                x = y; y = ny;
                uint64_t nxt = a_[x + 5 * y];
                a_[x + 5 * y] = r ? rol(cur, r) : cur;
                cur = nxt;
            }
            for (int yy = 0; yy < 25; yy += 5) {
                uint64_t row[5];
                for (int xx = 0; xx < 5; xx++) row[xx] = a_[yy + xx];
                for (int xx = 0; xx < 5; xx++) a_[yy + xx] = row[xx] ^ (~row[(xx + 1) % 5] & row[(xx + 2) % 5]);
            }
            rc = 0;
            for (int j = 0; j < 7; j++) {
                if (lfsr & 1) rc ^= 1ULL << ((1 << j) - 1);
                lfsr = (lfsr & 0x80) ? (uint8_t)((lfsr << 1) ^ 0x71) : (uint8_t)(lfsr << 1);
            }
            a_[0] ^= rc;
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
