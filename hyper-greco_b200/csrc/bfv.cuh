// Forward evaluation of the BFV secret-key-encryption circuit on the device (`circuit.evaluate`,
// /root/reference/bfv-gkr/src/sk_encryption_circuit.rs:442 over the topology of :86-293), restricted to the layers whose
// values are consumed downstream: the Lasso node's input (`lasso_inputs_batched`, :163-181) and the `sum` output (:280-285,
// which must equal ct0is). Vector layout is the reference's get_inputs (:365-415): length 2^L, position k <-> degree 2n-2-k.
#pragma once
#include "ntt.cuh"

namespace hg {

struct BfvShape {
    int log2_size;  // L = N_LOG2 + 1
    int K;          // num_reps
    int n_chunks;   // r2is chunks of 2^L (sk_encryption_circuit.rs:150-161)
};

// lasso_inputs = [r1is[i] + R1B[i] | r2 chunks + R2B[0] (Q7) | s + SB | e + EB | k1 + K1B]
template <class FP>
__global__ void k_bfv_lasso_inputs(BfvShape sh, const typename FP::B* __restrict__ s, const typename FP::B* __restrict__ e,
                                   const typename FP::B* __restrict__ k1, const typename FP::B* __restrict__ r1is,
                                   const typename FP::B* __restrict__ r2is, size_t r2_len, const typename FP::B* __restrict__ r1_bounds,
                                   typename FP::B r2_bound0, typename FP::B s_bound, typename FP::B e_bound, typename FP::B k1_bound,
                                   typename FP::B* __restrict__ out) {
    typedef typename FP::B B;
    const size_t N2 = (size_t)1 << sh.log2_size;
    const size_t total = (size_t)(sh.K + sh.n_chunks + 3) * N2;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const size_t seg = t / N2, j = t % N2;
    B v, b;
    if (seg < (size_t)sh.K) { v = r1is[seg * N2 + j]; b = r1_bounds[seg]; }
    else if (seg < (size_t)(sh.K + sh.n_chunks)) {
        const size_t q = (seg - sh.K) * N2 + j;
        v = q < r2_len ? r2is[q] : FP::b_zero();
        b = r2_bound0;
    } else {
        const size_t w = seg - sh.K - sh.n_chunks;
        v = w == 0 ? s[j] : (w == 1 ? e[j] : k1[j]);
        b = w == 0 ? s_bound : (w == 1 ? e_bound : k1_bound);
    }
    out[t] = FP::b_add(v, b);
}

// sum[i][j] = sai[i][j] + e[j] + k1[j]*k0[i] + r1is[i][j]*q[i] + r2i_cyclo[i][j]      (:97-141, :262-290)
template <class FP>
__global__ void k_bfv_sum(BfvShape sh, const typename FP::B* __restrict__ sai, const typename FP::B* __restrict__ e,
                          const typename FP::B* __restrict__ k1, const typename FP::B* __restrict__ r1is, const typename FP::B* __restrict__ r2is,
                          const typename FP::B* __restrict__ qis, const typename FP::B* __restrict__ k0is, typename FP::B* __restrict__ out) {
    typedef typename FP::B B;
    const size_t N2 = (size_t)1 << sh.log2_size, n = N2 / 2;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= N2) return;
    B v = FP::b_add(sai[(size_t)i * N2 + j], e[j]);
    v = FP::b_add(v, FP::b_mul(k1[j], k0is[i]));
    v = FP::b_add(v, FP::b_mul(r1is[(size_t)i * N2 + j], qis[i]));
    const size_t jj = j % n;
    if (jj < n - 1) v = FP::b_add(v, r2is[(size_t)i * n + jj]);
    out[(size_t)i * N2 + j] = v;
}

}  // namespace hg
