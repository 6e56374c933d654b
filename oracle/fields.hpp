// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement used as the parity checker.
// Nothing under hyper-greco_b200/ may include, link or call this file.
//
// Field arithmetic for the oracle: Goldilocks, GoldilocksExt2, BN254 Fr, Keccak-256.
//
// The reference takes these from un-vendored crates:
//   goldilocks  = github.com/han0110/goldilocks branch feature/qe_op_b, patched to
//                 github.com/nulltea/goldilocks branch to_canonical_repr   (/root/reference/Cargo.toml:28,67-68)
//   halo2curves = 0.7.0 (bn256::Fr)                                        (/root/reference/Cargo.toml:29)
//   plonkish_backend (Keccak256, fe_mod_from_le_bytes), no rev             (/root/reference/Cargo.toml:17)
// so what follows restates the published definitions:
//   Goldilocks p = 2^64 - 2^32 + 1 (/root/reference/scripts/circuit_sk.py:49)
//   GoldilocksExt2 = F[X]/(X^2 - 7)                       
//   BN254 Fr modulus r (/root/reference/scripts/circuit_sk.py:50)
//   Keccak-256 = original Keccak, pad 0x01, rate 136      
// Parity status: pinned only by the Appendix E chain values and keccak256("").
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace hgo {

typedef unsigned __int128 u128;

// ---------------------------------------------------------------- Goldilocks
struct Gl {
    static constexpr uint64_t P = 0xFFFFFFFF00000001ULL;
    static constexpr uint64_t EPS = 0xFFFFFFFFULL;
    uint64_t v;  // canonical, < P
    Gl() : v(0) {}
    explicit Gl(uint64_t x) : v(x >= P ? x - P : x) {}
    static Gl zero() { return Gl(); }
    static Gl one() { return Gl(1); }
    static Gl from_u64(uint64_t x) { return Gl(x); }
    static inline uint64_t reduce128(u128 x) {
        uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
        uint64_t hi_hi = hi >> 32, hi_lo = hi & EPS;
        uint64_t t0;
        if (__builtin_sub_overflow(lo, hi_hi, &t0)) t0 -= EPS;
        uint64_t t1 = hi_lo * EPS;
        uint64_t res;
        if (__builtin_add_overflow(t0, t1, &res)) res += EPS;
        if (res >= P) res -= P;
        return res;
    }
    friend inline Gl operator+(Gl a, Gl b) {
        uint64_t s;
        bool c = __builtin_add_overflow(a.v, b.v, &s);
        if (c || s >= P) s -= P;
        Gl r; r.v = s; return r;
    }
    friend inline Gl operator-(Gl a, Gl b) {
        Gl r; r.v = a.v >= b.v ? a.v - b.v : a.v + (P - b.v); return r;
    }
    friend inline Gl operator*(Gl a, Gl b) { Gl r; r.v = reduce128((u128)a.v * b.v); return r; }
    inline Gl operator-() const { Gl r; r.v = v ? P - v : 0; return r; }
    inline Gl& operator+=(Gl b) { *this = *this + b; return *this; }
    inline Gl& operator-=(Gl b) { *this = *this - b; return *this; }
    inline Gl& operator*=(Gl b) { *this = *this * b; return *this; }
    inline bool operator==(Gl b) const { return v == b.v; }
    inline bool operator!=(Gl b) const { return v != b.v; }
    Gl dbl() const { return *this + *this; }
    Gl square() const { return *this * *this; }
    Gl pow(uint64_t e) const {
        Gl r = one(), b = *this;
        while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
        return r;
    }
    Gl inv() const { return pow(P - 2); }
    // multiplicative generator 7, two-adicity 32: ROOT_OF_UNITY = 7^((p-1)/2^32)
    static constexpr int TWO_ADICITY = 32;
    static Gl root_of_unity() { return Gl(7).pow((P - 1) >> 32); }
    // canonical little-endian repr (A1)
    static constexpr int REPR_BYTES = 8;
    void to_repr_le(uint8_t* out) const { for (int i = 0; i < 8; i++) out[i] = (uint8_t)(v >> (8 * i)); }
    static bool from_repr_le(const uint8_t* in, Gl* out) {
        uint64_t x = 0; for (int i = 0; i < 8; i++) x |= (uint64_t)in[i] << (8 * i);
        if (x >= P) return false;
        out->v = x;
        return true;
    }
    // fe_mod_from_le_bytes: 256-bit little-endian integer mod p (A11)
    static Gl from_le_bytes_mod(const uint8_t* h, size_t n) {
        Gl acc, base = Gl(256);
        for (size_t i = n; i-- > 0;) acc = acc * base + Gl(h[i]);
        return acc;
    }
    uint64_t low_u64() const { return v; }
    static constexpr int LIMBS = 1;
    void to_limbs(uint64_t* out) const { out[0] = v; }
    static Gl from_limbs(const uint64_t* in) { return Gl(in[0] % P); }
};

struct Gl2 {
    Gl c0, c1;
    Gl2() {}
    Gl2(Gl a, Gl b) : c0(a), c1(b) {}
    static Gl2 zero() { return Gl2(); }
    static Gl2 one() { return Gl2(Gl::one(), Gl::zero()); }
    static constexpr int DEGREE = 2;
    static Gl2 from_base(Gl a) { return Gl2(a, Gl::zero()); }
    static Gl2 from_bases(const Gl* b) { return Gl2(b[0], b[1]); }
    void as_bases(Gl* b) const { b[0] = c0; b[1] = c1; }
    Gl base0() const { return c0; }
    friend inline Gl2 operator+(Gl2 a, Gl2 b) { return Gl2(a.c0 + b.c0, a.c1 + b.c1); }
    friend inline Gl2 operator-(Gl2 a, Gl2 b) { return Gl2(a.c0 - b.c0, a.c1 - b.c1); }
    friend inline Gl2 operator*(Gl2 a, Gl2 b) {
        // (a0 + a1 X)(b0 + b1 X) mod X^2 - 7
        Gl t = a.c1 * b.c1;
        return Gl2(a.c0 * b.c0 + Gl(7) * t, a.c0 * b.c1 + a.c1 * b.c0);
    }
    friend inline Gl2 operator*(Gl2 a, Gl b) { return Gl2(a.c0 * b, a.c1 * b); }
    inline Gl2 operator-() const { return Gl2(-c0, -c1); }
    inline Gl2& operator+=(Gl2 b) { *this = *this + b; return *this; }
    inline Gl2& operator-=(Gl2 b) { *this = *this - b; return *this; }
    inline Gl2& operator*=(Gl2 b) { *this = *this * b; return *this; }
    inline bool operator==(Gl2 b) const { return c0 == b.c0 && c1 == b.c1; }
    inline bool operator!=(Gl2 b) const { return !(*this == b); }
    Gl2 dbl() const { return *this + *this; }
    Gl2 square() const { return *this * *this; }
    Gl2 inv() const {
        // 1/(a0 + a1 X) = (a0 - a1 X)/(a0^2 - 7 a1^2)
        Gl n = c0 * c0 - Gl(7) * c1 * c1;
        Gl ni = n.inv();
        return Gl2(c0 * ni, (-c1) * ni);
    }
};

// ---------------------------------------------------------------- BN254 Fr (4x64 Montgomery)
struct Fr {
    uint64_t l[4];  // Montgomery form, canonical (< r)
    static constexpr uint64_t MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL,
                                        0x30644e72e131a029ULL};
    static constexpr uint64_t INV = 0xc2e1f593efffffffULL;  // -r^{-1} mod 2^64
    static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL,
                                       0x0216d0b17f4e44a5ULL};  // 2^512 mod r
    Fr() { l[0] = l[1] = l[2] = l[3] = 0; }
    static Fr zero() { return Fr(); }
    static Fr from_raw(const uint64_t* x) {  // x canonical integer < r  -> Montgomery
        Fr a; memcpy(a.l, x, 32);
        Fr r2; memcpy(r2.l, R2, 32);
        return mont_mul(a, r2);
    }
    static Fr from_u64(uint64_t x) { uint64_t t[4] = {x, 0, 0, 0}; return from_raw(t); }
    static Fr one() { return from_u64(1); }
    static inline bool geq_mod(const uint64_t* a) {
        for (int i = 3; i >= 0; i--) { if (a[i] > MOD[i]) return true; if (a[i] < MOD[i]) return false; }
        return true;
    }
    static inline void sub_mod_inplace(uint64_t* a) {
        u128 b = 0;
        for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - MOD[i] - (uint64_t)b; a[i] = (uint64_t)t; b = (t >> 64) & 1; }
    }
    static Fr mont_mul(const Fr& a, const Fr& b) {
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            u128 c = 0;
            for (int j = 0; j < 4; j++) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * INV;
            c = (u128)m * MOD[0] + t[0]; c >>= 64;
            for (int j = 1; j < 4; j++) { c += (u128)m * MOD[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
        }
        Fr r; memcpy(r.l, t, 32);
        if (t[4] || geq_mod(r.l)) sub_mod_inplace(r.l);
        return r;
    }
    void to_raw(uint64_t* out) const {  // Montgomery -> canonical integer
        Fr one_raw; one_raw.l[0] = 1;
        Fr r = mont_mul(*this, one_raw);
        memcpy(out, r.l, 32);
    }
    friend inline Fr operator+(const Fr& a, const Fr& b) {
        Fr r; u128 c = 0;
        for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
        if (c || geq_mod(r.l)) sub_mod_inplace(r.l);
        return r;
    }
    friend inline Fr operator-(const Fr& a, const Fr& b) {
        Fr r; u128 bw = 0;
        for (int i = 0; i < 4; i++) { u128 t = (u128)a.l[i] - b.l[i] - (uint64_t)bw; r.l[i] = (uint64_t)t; bw = (t >> 64) & 1; }
        if (bw) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.l[i] + MOD[i]; r.l[i] = (uint64_t)c; c >>= 64; } }
        return r;
    }
    friend inline Fr operator*(const Fr& a, const Fr& b) { return mont_mul(a, b); }
    inline Fr operator-() const { return zero() - *this; }
    inline Fr& operator+=(const Fr& b) { *this = *this + b; return *this; }
    inline Fr& operator-=(const Fr& b) { *this = *this - b; return *this; }
    inline Fr& operator*=(const Fr& b) { *this = *this * b; return *this; }
    inline bool operator==(const Fr& b) const { return !memcmp(l, b.l, 32); }
    inline bool operator!=(const Fr& b) const { return !(*this == b); }
    Fr dbl() const { return *this + *this; }
    Fr square() const { return *this * *this; }
    Fr pow(const uint64_t* e, int n) const {
        Fr r = one(), b = *this;
        for (int i = 0; i < n; i++) for (int k = 0; k < 64; k++) { if ((e[i] >> k) & 1) r *= b; b *= b; }
        return r;
    }
    Fr pow(uint64_t e) const { return pow(&e, 1); }
    Fr inv() const {
        uint64_t e[4]; memcpy(e, MOD, 32); e[0] -= 2;
        return pow(e, 4);
    }
    static constexpr int TWO_ADICITY = 28;
    static Fr root_of_unity() {  // 7^((r-1)/2^28) = halo2curves bn256::Fr::ROOT_OF_UNITY 0x03ddb9f5..0c37c9c (A9)
        uint64_t e[4]; memcpy(e, MOD, 32); e[0] -= 1;
        // shift right by 28
        for (int i = 0; i < 4; i++) e[i] = (e[i] >> 28) | (i < 3 ? (e[i + 1] << 36) : 0);
        return from_u64(7).pow(e, 4);
    }
    static constexpr int REPR_BYTES = 32;
    void to_repr_le(uint8_t* out) const { uint64_t t[4]; to_raw(t); memcpy(out, t, 32); }
    static bool from_repr_le(const uint8_t* in, Fr* out) {
        uint64_t t[4]; memcpy(t, in, 32);
        if (geq_mod(t)) return false;
        *out = from_raw(t); return true;
    }
    static Fr from_le_bytes_mod(const uint8_t* h, size_t n) {
        Fr acc, base = from_u64(256);
        for (size_t i = n; i-- > 0;) acc = acc * base + from_u64(h[i]);
        return acc;
    }
    uint64_t low_u64() const { uint64_t t[4]; to_raw(t); return t[0]; }
    static constexpr int LIMBS = 4;
    void to_limbs(uint64_t* out) const { to_raw(out); }
    static Fr from_limbs(const uint64_t* in) { return from_raw(in); }
    // extension-degree-1 interface (for BN254 the reference uses E = F, sk_encryption_circuit.rs:616)
    static constexpr int DEGREE = 1;
    static Fr from_base(const Fr& a) { return a; }
    static Fr from_bases(const Fr* b) { return b[0]; }
    void as_bases(Fr* b) const { b[0] = *this; }
    Fr base0() const { return *this; }
};

// ---------------------------------------------------------------- Keccak-256 (pad 0x01)
static inline uint64_t rotl64(uint64_t x, int s) { return s ? (x << s) | (x >> (64 - s)) : x; }
static inline void keccak_f1600(uint64_t st[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int round = 0; round < 24; round++) {
        uint64_t C[5], D[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) st[i] ^= D[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(st[x + 5 * y], ROT[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) st[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        st[0] ^= RC[round];
    }
}
static inline void keccak256(const uint8_t* in, size_t n, uint8_t out[32]) {
    const size_t rate = 136;
    uint64_t st[25]; memset(st, 0, sizeof st);
    while (n >= rate) {
        for (size_t i = 0; i < rate / 8; i++) { uint64_t w; memcpy(&w, in + 8 * i, 8); st[i] ^= w; }
        keccak_f1600(st); in += rate; n -= rate;
    }
    uint8_t blk[136]; memset(blk, 0, rate); memcpy(blk, in, n);
    blk[n] ^= 0x01; blk[rate - 1] ^= 0x80;
    for (size_t i = 0; i < rate / 8; i++) { uint64_t w; memcpy(&w, blk + 8 * i, 8); st[i] ^= w; }
    keccak_f1600(st);
    memcpy(out, st, 32);
}

// field traits glue: base field F, "extension" E
template <class F> struct ExtOf;
template <> struct ExtOf<Gl> { typedef Gl2 type; };
template <> struct ExtOf<Fr> { typedef Fr type; };

inline Gl2 operator*(Gl a, Gl2 b) { return b * a; }

}  // namespace hgo
