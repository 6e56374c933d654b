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
#include <mutex>
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
    // Keccak-f[1600], fully unrolled (constant lane indices keep the state in registers). Round constants come from the
    // degree-8 LFSR x^8+x^6+x^5+x^4+1 and the rho/pi schedule from the (x,y) -> (y, 2x+3y) orbit; both are expanded by the
    // generator in this file's history and checked against keccak256("") in the tests.
    void permute() {
        static const uint64_t RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
        uint64_t a[25];
        memcpy(a, a_, sizeof a);
        for (int round = 0; round < 24; round++) {
            const uint64_t c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20], c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21],
                           c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22], c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23],
                           c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
            const uint64_t d0 = c4 ^ rol(c1, 1), d1 = c0 ^ rol(c2, 1), d2 = c1 ^ rol(c3, 1), d3 = c2 ^ rol(c4, 1), d4 = c3 ^ rol(c0, 1);
            a[0] ^= d0; a[5] ^= d0; a[10] ^= d0; a[15] ^= d0; a[20] ^= d0;
            a[1] ^= d1; a[6] ^= d1; a[11] ^= d1; a[16] ^= d1; a[21] ^= d1;
            a[2] ^= d2; a[7] ^= d2; a[12] ^= d2; a[17] ^= d2; a[22] ^= d2;
            a[3] ^= d3; a[8] ^= d3; a[13] ^= d3; a[18] ^= d3; a[23] ^= d3;
            a[4] ^= d4; a[9] ^= d4; a[14] ^= d4; a[19] ^= d4; a[24] ^= d4;
            uint64_t cur = a[1], nxt;
            nxt = a[10]; a[10] = rol(cur, 1); cur = nxt;
            nxt = a[7]; a[7] = rol(cur, 3); cur = nxt;
            nxt = a[11]; a[11] = rol(cur, 6); cur = nxt;
            nxt = a[17]; a[17] = rol(cur, 10); cur = nxt;
            nxt = a[18]; a[18] = rol(cur, 15); cur = nxt;
            nxt = a[3]; a[3] = rol(cur, 21); cur = nxt;
            nxt = a[5]; a[5] = rol(cur, 28); cur = nxt;
            nxt = a[16]; a[16] = rol(cur, 36); cur = nxt;
            nxt = a[8]; a[8] = rol(cur, 45); cur = nxt;
            nxt = a[21]; a[21] = rol(cur, 55); cur = nxt;
            nxt = a[24]; a[24] = rol(cur, 2); cur = nxt;
            nxt = a[4]; a[4] = rol(cur, 14); cur = nxt;
            nxt = a[15]; a[15] = rol(cur, 27); cur = nxt;
            nxt = a[23]; a[23] = rol(cur, 41); cur = nxt;
            nxt = a[19]; a[19] = rol(cur, 56); cur = nxt;
            nxt = a[13]; a[13] = rol(cur, 8); cur = nxt;
            nxt = a[12]; a[12] = rol(cur, 25); cur = nxt;
            nxt = a[2]; a[2] = rol(cur, 43); cur = nxt;
            nxt = a[20]; a[20] = rol(cur, 62); cur = nxt;
            nxt = a[14]; a[14] = rol(cur, 18); cur = nxt;
            nxt = a[22]; a[22] = rol(cur, 39); cur = nxt;
            nxt = a[9]; a[9] = rol(cur, 61); cur = nxt;
            nxt = a[6]; a[6] = rol(cur, 20); cur = nxt;
            nxt = a[1]; a[1] = rol(cur, 44); cur = nxt;
#define HG_CHI(o) { const uint64_t r0 = a[o], r1 = a[o + 1], r2 = a[o + 2], r3 = a[o + 3], r4 = a[o + 4]; \
                    a[o] = r0 ^ (~r1 & r2); a[o + 1] = r1 ^ (~r2 & r3); a[o + 2] = r2 ^ (~r3 & r4); a[o + 3] = r3 ^ (~r4 & r0); a[o + 4] = r4 ^ (~r0 & r1); }
            HG_CHI(0) HG_CHI(5) HG_CHI(10) HG_CHI(15) HG_CHI(20)
#undef HG_CHI
            a[0] ^= RC[round];
        }
        memcpy(a_, a, sizeof a);
    }
};

struct TranscriptError : std::runtime_error { using std::runtime_error::runtime_error; };

// The challenge stream of the reference transcript is the fixed chain c_i = fe_mod_from_le_bytes(keccak^{i+1}("")): squeeze
// re-hashes the hasher's own previous output and nothing on the prove path is ever absorbed (transcript.rs:156,183-203, SURVEY
// F3). A fresh transcript therefore always yields the same sequence, so the chain is computed once per process and extended on
// demand; squeezing the i-th challenge is a table lookup. Bit-identical to hashing every time (tests/test_abi.py).
template <class HF> class ChallengeChain {
  public:
    static typename HF::Base get(size_t i) {
        static ChallengeChain chain;
        std::lock_guard<std::mutex> lock(chain.mu_);
        while (chain.vals_.size() <= i) chain.extend();
        return chain.vals_[i];
    }

  private:
    void extend() {
        Keccak256 k;
        if (!vals_.empty()) k.update(last_, 32);
        k.finalize_reset(last_);
        vals_.push_back(HF::base_from_le_bytes_mod(last_, 32));
    }
    std::mutex mu_;
    std::vector<typename HF::Base> vals_;
    uint8_t last_[32];
};

// A transcript owned by the CALLER: the three primitives the prove / verify path uses (squeeze_challenge, write_felt_ext,
// read_felt_ext: transcript.rs:146-157, :172-177, :191-195) as C callbacks over canonical little-endian limbs. This is what a
// `&mut dyn TranscriptWrite<F, E>` handed to Node::prove_claim_reduction (lasso.rs:58-63) is wrapped into.
// Non-zero return = the callback failed (mapped to TranscriptError).
struct TranscriptHooks {
    void* user = nullptr;
    int (*squeeze)(void* user, uint64_t* out_ext) = nullptr;
    int (*write)(void* user, const uint64_t* ext) = nullptr;
    int (*read)(void* user, uint64_t* out_ext) = nullptr;
    bool message_independent = false;  // the caller guarantees challenges do not depend on written messages (prefetch mode allowed)
};

// HF: host field traits (see host_field.hpp): Base, Ext, base_from_le_bytes_mod, base_to_repr_le, base_from_repr_le ...
template <class HF> class Keccak256Transcript {
  public:
    typedef typename HF::Base Base;
    typedef typename HF::Ext Ext;
    Keccak256Transcript() {}                                                  // Keccak256Transcript::<Vec<u8>>::default()
    Keccak256Transcript(const uint8_t* proof, size_t n) : rd_(proof, proof + n), reading_(true) {}  // from_proof
    explicit Keccak256Transcript(const TranscriptHooks& h) : hooks_(h), hooked_(true) {}             // caller-owned transcript behind callbacks

    bool hooked() const { return hooked_; }
    // may every challenge of a proof be squeezed before any message is written? True for the reference transcript (it never
    // absorbs, transcript.rs:156,183-203); for a caller-owned one only if the caller says so
    bool prefetch_legal() const { return !hooked_ || hooks_.message_independent; }
    Base squeeze_base() {
        if (hooked_) throw TranscriptError("squeeze_base on a callback transcript");
        return ChallengeChain<HF>::get(n_squeezed_++);
    }
    Ext squeeze_challenge() {
        if (hooked_) {
            uint64_t limbs[8] = {0};
            if (!hooks_.squeeze || hooks_.squeeze(hooks_.user, limbs)) throw TranscriptError("transcript callback squeeze_challenge failed");
            n_squeezed_ += HF::DEGREE;
            return HF::x_from_limbs(limbs);
        }
        Base b[HF::DEGREE];
        for (int i = 0; i < HF::DEGREE; i++) b[i] = squeeze_base();
        return HF::ext_from_bases(b);
    }
    void common_felt(const Base&) {}
    void write_felt(const Base& f) {
        if (discard_) return;
        uint8_t b[HF::REPR_BYTES];
        HF::base_to_repr_le(f, b);
        const size_t o = stream_.size();
        stream_.resize(o + HF::REPR_BYTES);
        uint8_t* dst = stream_.data() + o;
        for (int i = 0; i < HF::REPR_BYTES; i++) dst[i] = b[HF::REPR_BYTES - 1 - i];
    }
    // a transcript that swallows what is written to it (Channel: closures replayed for their bookkeeping only)
    void set_discard(bool d) { discard_ = d; }
    void append_raw(const uint8_t* b, size_t n) {  // bytes serialised elsewhere (a sharded proof assembled from per-rank parts)
        if (hooked_) throw TranscriptError("append_raw on a callback transcript");
        stream_.insert(stream_.end(), b, b + n);
    }
    void write_felt_ext(const Ext& e) {
        if (discard_) return;
        if (hooked_) {
            uint64_t limbs[8] = {0};
            HF::x_to_limbs(e, limbs);
            if (!hooks_.write || hooks_.write(hooks_.user, limbs)) throw TranscriptError("transcript callback write_felt_ext failed");
            return;
        }
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
        if (hooked_) {
            uint64_t limbs[8] = {0};
            if (!hooks_.read || hooks_.read(hooks_.user, limbs)) throw TranscriptError("transcript callback read_felt_ext failed");
            return HF::x_from_limbs(limbs);
        }
        Base b[HF::DEGREE];
        for (int i = 0; i < HF::DEGREE; i++) b[i] = read_felt();
        return HF::ext_from_bases(b);
    }
    const std::vector<uint8_t>& proof() const { return stream_; }  // into_proof
    void append_bytes(const std::vector<uint8_t>& b) {  // messages serialised elsewhere (never for a callback transcript: Channel::begin_side)
        if (hooked_) throw TranscriptError("append_bytes on a callback transcript");
        stream_.insert(stream_.end(), b.begin(), b.end());
    }
    size_t num_base_squeezed() const { return n_squeezed_; }
    size_t read_pos() const { return pos_; }

  private:
    std::vector<uint8_t> stream_, rd_;
    size_t pos_ = 0, n_squeezed_ = 0;
    bool reading_ = false, discard_ = false;
    TranscriptHooks hooks_;
    bool hooked_ = false;
};

}  // namespace hg
