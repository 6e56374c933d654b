// Grand-product layer sumcheck kernels (K8; /root/reference/lasso/src/memory_checking/prover.rs:223-279), the dominant
// cost of the Lasso node. One round of
//        g = t_0 * sum_i c_i * l_i * r_i          (t_0 = l_0, F4/Q2; c_i = gamma^i under A5)
// per launch, evaluation fused with the fold by the previous challenge. Differences from the generic k_sc_round:
//   * work is split over (pair-index tiles) x (term groups) so that small rounds still fill the GPU and every thread has
//     several independent 256-bit loads in flight;
//   * because h(X) = sum_i c_i sum_b t_0 l_i r_i is linear in the terms, the c_i are folded into the TABLES once
//     (round 1 writes l'_i = c_i * l_i; folding is linear, so later rounds never multiply by c_i again) and the final
//     evaluations are unscaled on the host with c_i^{-1}. Field arithmetic is exact, so the messages are bit-identical;
//   * round 0 runs entirely in the base field and multiplies by c_i once per (thread, term).
#pragma once
#include "kernels.cuh"

namespace hg {

// ---- 256-bit global loads / stores (LDG.E.256 / STG.E.256 on sm_100a), streaming (no L1 allocation)
__device__ __forceinline__ void ldg256(const void* p, u64& a, u64& b, u64& c, u64& d) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, u64 a, u64 b, u64 c, u64 d) {
    asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ void ldg128(const void* p, u64& a, u64& b) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}
// four consecutive elements starting at p (32-byte aligned)
__device__ __forceinline__ void load4(const u64* p, u64 (&v)[4]) { ldg256(p, v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void load4(const gl2* p, gl2 (&v)[4]) {
    ldg256(p, v[0].c0, v[0].c1, v[1].c0, v[1].c1);
    ldg256(p + 2, v[2].c0, v[2].c1, v[3].c0, v[3].c1);
}
__device__ __forceinline__ void load2(const u64* p, u64 (&v)[2]) { ldg128(p, v[0], v[1]); }
__device__ __forceinline__ void store2(gl2* p, gl2 a, gl2 b) { stg256(p, a.c0, a.c1, b.c0, b.c1); }

// values of the line through (0, lo), (1, hi) at the message points: slot 0 -> X=0, 1 -> X=2, 2 -> X=3 [, 3 -> X=1]
template <class FP, class T, int NP> __device__ __forceinline__ void line_points(T lo, T hi, T (&v)[NP]) {
    T df = FP::sub(hi, lo);
    v[0] = lo;
    v[1] = FP::add(hi, df);
    v[2] = FP::add(v[1], df);
    if constexpr (NP == 4) v[3] = hi;
}

// ---- round 0: base-field tables [nvec][2n] (l_i = first half, r_i = second half of vector i), n = 2^nv
template <class FP, int U, bool WITH_H1>
__global__ void __launch_bounds__(HG_BLOCK)
k_gp_r0(const typename FP::B* __restrict__ tables, size_t n, int nvec, int tpg, const typename FP::X* __restrict__ coeffs,
        typename FP::X* partials, unsigned* counter, typename FP::X* msg) {
    typedef typename FP::B B;
    typedef typename FP::X X;
    constexpr int NP = WITH_H1 ? 4 : 3;
    const size_t npairs = n / 2, stride = (size_t)gridDim.x * blockDim.x;
    const int i0 = blockIdx.y * tpg, i1 = min(nvec, i0 + tpg);
    X acc[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) acc[p] = FP::x_zero();
    for (size_t b0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b0 < npairs; b0 += stride * U) {
        B t0[U][NP];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t b = b0 + u * stride;
            B p[2] = {FP::b_zero(), FP::b_zero()};
            if (b < npairs) load2(tables + 2 * b, p);
            line_points<FP, B, NP>(p[0], p[1], t0[u]);
        }
        for (int i = i0; i < i1; i++) {
            const B* li = tables + (size_t)i * 2 * n;
            const B* ri = li + n;
            B l[U][2], r[U][2];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const size_t b = b0 + u * stride;
                l[u][0] = l[u][1] = r[u][0] = r[u][1] = FP::b_zero();
                if (b < npairs) { load2(li + 2 * b, l[u]); load2(ri + 2 * b, r[u]); }
            }
            B s[NP];
#pragma unroll
            for (int p = 0; p < NP; p++) s[p] = FP::b_zero();
#pragma unroll
            for (int u = 0; u < U; u++) {
                B vl[NP], vr[NP];
                line_points<FP, B, NP>(l[u][0], l[u][1], vl);
                line_points<FP, B, NP>(r[u][0], r[u][1], vr);
#pragma unroll
                for (int p = 0; p < NP; p++) s[p] = FP::b_add(s[p], FP::b_mul(t0[u][p], FP::b_mul(vl[p], vr[p])));
            }
            const X c = coeffs[i];
#pragma unroll
            for (int p = 0; p < NP; p++) acc[p] = FP::x_add(acc[p], FP::x_mul_b(c, s[p]));
        }
    }
    block_reduce_finalize<FP, NP, true>(acc, partials, counter, msg);
}

// ---- rounds >= 1: fold by r_prev, write the folded tables, evaluate the next round polynomial on them
//   in : [2*nvec][n_in] (TIN = B in round 1, X later);  out: [2*nvec][n_in/2]
//   SCALE (round 1): l_i (i > 0) is multiplied by c_i when it is written; later rounds see pre-scaled tables
template <class FP, class TIN, bool SCALE, bool WITH_H1>
__global__ void __launch_bounds__(HG_BLOCK)
k_gp_fold(const TIN* __restrict__ in, typename FP::X* __restrict__ out, size_t n_in, int nvec, int tpg,
          const typename FP::X* __restrict__ coeffs, const typename FP::X* __restrict__ r_prev, typename FP::X* partials, unsigned* counter,
          typename FP::X* msg) {
    typedef typename FP::X X;
    constexpr int NP = WITH_H1 ? 4 : 3;
    const size_t npairs = n_in / 4, n_out = n_in / 2;
    const int i0 = blockIdx.y * tpg, i1 = min(nvec, i0 + tpg);
    const X r = *r_prev;
    const X c0 = coeffs[0];
    X acc[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) acc[p] = FP::x_zero();
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < npairs; b += (size_t)gridDim.x * blockDim.x) {
        X t0[NP];
        {
            TIN a[4];
            load4(in + 4 * b, a);
            X lo = FP::x_add(FP::as_x(a[0]), FP::mul(r, FP::sub(a[1], a[0])));
            X hi = FP::x_add(FP::as_x(a[2]), FP::mul(r, FP::sub(a[3], a[2])));
            line_points<FP, X, NP>(lo, hi, t0);
        }
        X P[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) P[p] = FP::x_zero();
#pragma unroll 2
        for (int i = i0; i < i1; i++) {
            TIN a[4], c[4];
            load4(in + (size_t)(2 * i) * n_in + 4 * b, a);
            load4(in + (size_t)(2 * i + 1) * n_in + 4 * b, c);
            X l_lo = FP::x_add(FP::as_x(a[0]), FP::mul(r, FP::sub(a[1], a[0])));
            X l_hi = FP::x_add(FP::as_x(a[2]), FP::mul(r, FP::sub(a[3], a[2])));
            X r_lo = FP::x_add(FP::as_x(c[0]), FP::mul(r, FP::sub(c[1], c[0])));
            X r_hi = FP::x_add(FP::as_x(c[2]), FP::mul(r, FP::sub(c[3], c[2])));
            if (SCALE && i > 0) {
                const X ci = coeffs[i];
                l_lo = FP::x_mul(l_lo, ci);
                l_hi = FP::x_mul(l_hi, ci);
            }
            store2(out + (size_t)(2 * i) * n_out + 2 * b, l_lo, l_hi);
            store2(out + (size_t)(2 * i + 1) * n_out + 2 * b, r_lo, r_hi);
            X vl[NP], vr[NP];
            line_points<FP, X, NP>(l_lo, l_hi, vl);
            line_points<FP, X, NP>(r_lo, r_hi, vr);
#pragma unroll
            for (int p = 0; p < NP; p++) {
                X pr = FP::x_mul(vl[p], vr[p]);
                if (i == 0) pr = FP::x_mul(pr, c0);
                P[p] = FP::x_add(P[p], pr);
            }
        }
#pragma unroll
        for (int p = 0; p < NP; p++) acc[p] = FP::x_add(acc[p], FP::x_mul(t0[p], P[p]));
    }
    block_reduce_finalize<FP, NP, true>(acc, partials, counter, msg);
}

}  // namespace hg
