// Grand-product layer sumcheck kernels (K8; /root/reference/lasso/src/memory_checking/prover.rs:223-279), the dominant
// cost of the Lasso node. One round of
//        g = t_0 * sum_i c_i * l_i * r_i          (t_0 = l_0, F4/Q2; c_i = gamma^i under A5)
// per launch, evaluation fused with the fold by the previous challenge.
//   * work is split over (pair-index tiles) x (term groups) so that small rounds still fill the GPU and every thread has
//     several independent 256-bit loads in flight;
//   * h(X) = sum_i c_i sum_b t_0 l_i r_i is linear in the terms, so the c_i are folded into the TABLES once (round 1
//     writes l'_i = c_i * l_i; folding is linear, later rounds never multiply by c_i again) and the final evaluations are
//     unscaled on the host with c_i^{-1}. Field arithmetic is exact, so the messages are bit-identical;
//   * the round polynomial is sampled at X = 0, infinity (leading coefficient) and -1: those line values cost one
//     subtraction each. Together with the running TRUE sum h(0) + h(1) (tracked on the host, h(1) is sampled only in
//     round 0) they determine the cubic exactly; the host derives whatever the wire format needs (A3);
//   * products are accumulated unreduced (gl.cuh acc192 / xacc) and reduced once per accumulator: the kernels are bound by
//     the integer ALU pipe, not by HBM, so instruction count is what matters (profiles/).
// Message slots: [h(0), h(inf), h(-1)] and, in round 0 only, [.., h(1)].
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

// ---- round 0: base-field tables [nvec][2n] (l_i = first half, r_i = second half of vector i), n = 2^nv
//      msg: h(0), h(inf), h(-1), h(1)
template <class FP, int U>
__global__ void __launch_bounds__(HG_BLOCK)
k_gp_r0(const typename FP::B* __restrict__ tables, size_t n, int nvec, int tpg, const typename FP::X* __restrict__ coeffs,
        typename FP::X* partials, unsigned* counter, typename FP::X* msg) {
    typedef typename FP::B B;
    typedef typename FP::X X;
    constexpr int NP = 4;
    const size_t npairs = n / 2, stride = (size_t)gridDim.x * blockDim.x;
    const int i0 = blockIdx.y * tpg, i1 = min(nvec, i0 + tpg);
    typename FP::XAcc accx[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) accx[p] = FP::xacc_zero_();
    for (size_t b0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b0 < npairs; b0 += stride * U) {
        B t0[U][NP];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t b = b0 + u * stride;
            B p[2] = {FP::b_zero(), FP::b_zero()};
            if (b < npairs) load2(tables + 2 * b, p);
            t0[u][0] = p[0]; t0[u][1] = FP::slope(p[0], p[1]); t0[u][2] = FP::at_m1(p[0], p[1]); t0[u][3] = p[1];
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
            typename FP::BAcc s[NP];
#pragma unroll
            for (int p = 0; p < NP; p++) s[p] = FP::bacc_zero();
#pragma unroll
            for (int u = 0; u < U; u++) {
                FP::bacc_mad(s[0], t0[u][0], FP::fmul(l[u][0], r[u][0]));
                FP::bacc_mad(s[1], t0[u][1], FP::fmul(FP::slope(l[u][0], l[u][1]), FP::slope(r[u][0], r[u][1])));
                FP::bacc_mad(s[2], t0[u][2], FP::fmul(FP::at_m1(l[u][0], l[u][1]), FP::at_m1(r[u][0], r[u][1])));
                FP::bacc_mad(s[3], t0[u][3], FP::fmul(l[u][1], r[u][1]));
            }
            const X c = coeffs[i];
#pragma unroll
            for (int p = 0; p < NP; p++) FP::xacc_mad_b(accx[p], c, FP::bacc_reduce(s[p]));
        }
    }
    X acc[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) acc[p] = FP::xacc_reduce_(accx[p]);
    block_reduce_finalize<FP, NP, true>(acc, partials, counter, msg);
}

// ---- rounds >= 1: fold by r_prev, write the folded tables, evaluate the next round polynomial on them
//   in : [2*nvec][n_in] (TIN = B in round 1, X later);  out: [2*nvec][n_in/2]
//   SCALE (round 1, TIN = B): l_i (i > 0) is written as c_i * l_i; cr[i] = c_i * r_prev is precomputed (k_gp_coeffs)
//   msg: h(0), h(inf), h(-1)
template <class FP, class TIN, bool SCALE>
__global__ void __launch_bounds__(HG_BLOCK)
k_gp_fold(const TIN* __restrict__ in, typename FP::X* __restrict__ out, size_t n_in, int nvec, int tpg,
          const typename FP::X* __restrict__ coeffs, const typename FP::X* __restrict__ cr, const typename FP::X* __restrict__ r_prev,
          typename FP::X* partials, unsigned* counter, typename FP::X* msg) {
    typedef typename FP::X X;
    constexpr int NP = 3;
    const size_t npairs = n_in / 4, n_out = n_in / 2;
    const int i0 = blockIdx.y * tpg, i1 = min(nvec, i0 + tpg);
    const X r = *r_prev;
    const typename FP::FoldAux aux = FP::fold_aux(r);
    X acc[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) acc[p] = FP::x_zero();
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < npairs; b += (size_t)gridDim.x * blockDim.x) {
        X t0[NP];
        {
            TIN a[4];
            load4(in + 4 * b, a);
            X lo = FP::fold(a[0], a[1], r, aux), hi = FP::fold(a[2], a[3], r, aux);
            t0[0] = lo; t0[1] = FP::slope(lo, hi); t0[2] = FP::at_m1(lo, hi);
        }
        typename FP::XAcc P[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) P[p] = FP::xacc_zero_();
        for (int i = i0; i < i1; i++) {
            TIN a[4], c[4];
            load4(in + (size_t)(2 * i) * n_in + 4 * b, a);
            load4(in + (size_t)(2 * i + 1) * n_in + 4 * b, c);
            X l_lo, l_hi;
            if constexpr (SCALE) {
                if (i > 0) {
                    const X ci = coeffs[i], cri = cr[i];
                    l_lo = FP::fold_scaled(a[0], a[1], ci, cri);
                    l_hi = FP::fold_scaled(a[2], a[3], ci, cri);
                } else {
                    l_lo = FP::fold(a[0], a[1], r, aux);
                    l_hi = FP::fold(a[2], a[3], r, aux);
                }
            } else {
                l_lo = FP::fold(a[0], a[1], r, aux);
                l_hi = FP::fold(a[2], a[3], r, aux);
            }
            const X r_lo = FP::fold(c[0], c[1], r, aux), r_hi = FP::fold(c[2], c[3], r, aux);
            store2(out + (size_t)(2 * i) * n_out + 2 * b, l_lo, l_hi);
            store2(out + (size_t)(2 * i + 1) * n_out + 2 * b, r_lo, r_hi);
            if (i == 0) {
                // term 0 keeps l_0 unscaled (it is also t_0), so its products take c_0 explicitly
                const X c0 = coeffs[0];
                FP::xacc_mad_(P[0], FP::fmul(c0, l_lo), r_lo);
                FP::xacc_mad_(P[1], FP::fmul(c0, FP::slope(l_lo, l_hi)), FP::slope(r_lo, r_hi));
                FP::xacc_mad_(P[2], FP::fmul(c0, FP::at_m1(l_lo, l_hi)), FP::at_m1(r_lo, r_hi));
            } else {
                FP::xacc_mad_(P[0], l_lo, r_lo);
                FP::xacc_mad_(P[1], FP::slope(l_lo, l_hi), FP::slope(r_lo, r_hi));
                FP::xacc_mad_(P[2], FP::at_m1(l_lo, l_hi), FP::at_m1(r_lo, r_hi));
            }
        }
#pragma unroll
        for (int p = 0; p < NP; p++) acc[p] = FP::x_add(acc[p], FP::fmul(t0[p], FP::xacc_reduce_(P[p])));
    }
    block_reduce_finalize<FP, NP, true>(acc, partials, counter, msg);
}

// per-layer coefficient tables: c[i] = gamma^i (A5: or reversed), cr[i] = c[i] * r_0 when r0 != nullptr
template <class FP>
__global__ void k_gp_coeffs(const typename FP::X* __restrict__ gamma, const typename FP::X* __restrict__ r0, int n, int ascending,
                            typename FP::X* __restrict__ c, typename FP::X* __restrict__ cr) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    typename FP::X p = FP::x_one(), g = *gamma;
    for (int i = 0; i < n; i++) {
        const int k = ascending ? i : n - 1 - i;
        c[k] = p;
        if (r0) cr[k] = FP::x_mul(p, *r0);
        p = FP::x_mul(p, g);
    }
}

}  // namespace hg
