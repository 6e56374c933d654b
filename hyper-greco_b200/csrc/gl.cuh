// Goldilocks (p = 2^64 - 2^32 + 1) and GoldilocksExt2 = F[X]/(X^2 - 7) arithmetic, host + device.
//
// Replaces the `goldilocks` crate the reference links (/root/reference/Cargo.toml:28,67-68; used at
// /root/reference/bfv-gkr/src/sk_encryption_circuit.rs:539). Elements are canonical u64 (< p) in memory; the
// reduction is the 64x64->128 mul.lo/mul.hi + "2^64 = 2^32 - 1, 2^96 = -1" fold (IMAD pipe, no tensor cores).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define HG_HD __host__ __device__ __forceinline__
#else
#define HG_HD inline
#endif

namespace hg {

typedef uint64_t u64;
typedef uint32_t u32;

constexpr u64 GL_P = 0xFFFFFFFF00000001ULL;
constexpr u64 GL_EPS = 0xFFFFFFFFULL;  // 2^64 mod p

HG_HD u64 gl_add(u64 a, u64 b) {
    u64 s = a + b;
    if (s < a) s += GL_EPS;       // wrapped: + 2^64 = + EPS (result < p)
    else if (s >= GL_P) s -= GL_P;
    return s;
}
HG_HD u64 gl_sub(u64 a, u64 b) {
    u64 d = a - b;
    if (a < b) d -= GL_EPS;       // wrapped: - 2^64 = - EPS
    return d;
}
HG_HD u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }
HG_HD u64 gl_dbl(u64 a) { return gl_add(a, a); }

HG_HD void mul_wide(u64 a, u64 b, u64& lo, u64& hi) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    lo = (u64)p;
    hi = (u64)(p >> 64);
#endif
}
// (lo + 2^64 hi) mod p, canonical
HG_HD u64 gl_reduce128(u64 lo, u64 hi) {
    u64 hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    u64 t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS;
    u64 t1 = (hi_lo << 32) - hi_lo;  // hi_lo * (2^32 - 1)
    u64 r = t0 + t1;
    if (r < t1) r += GL_EPS;
    if (r >= GL_P) r -= GL_P;
    return r;
}
HG_HD u64 gl_mul(u64 a, u64 b) {
    u64 lo, hi;
    mul_wide(a, b, lo, hi);
    return gl_reduce128(lo, hi);
}
HG_HD u64 gl_from_u64(u64 x) { return x >= GL_P ? x - GL_P : x; }
HG_HD u64 gl_pow(u64 b, u64 e) {
    u64 r = 1;
    while (e) { if (e & 1) r = gl_mul(r, b); b = gl_mul(b, b); e >>= 1; }
    return r;
}
HG_HD u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }

// ------------------------------------------------------------------ Ext2
struct __align__(16) gl2 {
    u64 c0, c1;
};
HG_HD gl2 gl2_make(u64 a, u64 b) { gl2 r; r.c0 = a; r.c1 = b; return r; }
HG_HD gl2 gl2_zero() { return gl2_make(0, 0); }
HG_HD gl2 gl2_one() { return gl2_make(1, 0); }
HG_HD gl2 gl2_lift(u64 a) { return gl2_make(a, 0); }
HG_HD gl2 gl2_add(gl2 a, gl2 b) { return gl2_make(gl_add(a.c0, b.c0), gl_add(a.c1, b.c1)); }
HG_HD gl2 gl2_sub(gl2 a, gl2 b) { return gl2_make(gl_sub(a.c0, b.c0), gl_sub(a.c1, b.c1)); }
HG_HD gl2 gl2_neg(gl2 a) { return gl2_make(gl_neg(a.c0), gl_neg(a.c1)); }
HG_HD u64 gl_mul7(u64 a) {  // 7a = 8a - a
    u64 lo = a << 3, hi = a >> 61;
    u64 r = gl_reduce128(lo, hi);
    return gl_sub(r, a);
}
HG_HD gl2 gl2_mul(gl2 a, gl2 b) {
    // Karatsuba: c0 = a0 b0 + 7 a1 b1, c1 = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1
    u64 v0 = gl_mul(a.c0, b.c0), v1 = gl_mul(a.c1, b.c1);
    u64 m = gl_mul(gl_add(a.c0, a.c1), gl_add(b.c0, b.c1));
    return gl2_make(gl_add(v0, gl_mul7(v1)), gl_sub(gl_sub(m, v0), v1));
}
HG_HD gl2 gl2_mul_base(gl2 a, u64 b) { return gl2_make(gl_mul(a.c0, b), gl_mul(a.c1, b)); }
HG_HD gl2 gl2_dbl(gl2 a) { return gl2_add(a, a); }
HG_HD bool gl2_eq(gl2 a, gl2 b) { return a.c0 == b.c0 && a.c1 == b.c1; }
HG_HD gl2 gl2_inv(gl2 a) {
    u64 n = gl_sub(gl_mul(a.c0, a.c0), gl_mul7(gl_mul(a.c1, a.c1)));
    u64 ni = gl_inv(n);
    return gl2_make(gl_mul(a.c0, ni), gl_mul(gl_neg(a.c1), ni));
}

}  // namespace hg
