// K10: radix-2 Goldilocks NTT for the FFT -> dot-product -> IFFT poly-mult layers
// (call sites /root/reference/bfv-gkr/src/sk_encryption_circuit.rs:224,249,251; the transform itself lives in the
// un-vendored gkr crate, assumption A9: w = ROOT_OF_UNITY^(2^(32 - log n)), natural order in and out, inverse scaled by 1/n).
//
// Four-step decomposition N = N1 * N2 so that every butterfly runs in shared memory and every global access is a
// coalesced 128-byte segment:
//   pass 1 (k_ntt_cols): for a tile of TILE columns j2, N1-point NTT over j1 (x[j1*N2 + j2]), times w_N^(j2*k1) -> Y[k1*N2 + j2]
//   pass 2 (k_ntt_rows): for a tile of TILE rows k1, N2-point NTT over j2 -> X[k1 + N1*k2]
// Batched over independent transforms (blockIdx.y). In-shared-memory transforms are decimation-in-frequency with the
// bit-reversed output undone when the tile is written back.
#pragma once
#include "field_policy.cuh"

namespace hg {

constexpr int HG_NTT_TILE = 16;
constexpr int HG_NTT_THREADS = 256;

// table[k] = w^k for k < n, w = primitive n-th root (or its inverse)
template <class FP> __global__ void k_ntt_twiddles(int log_n, int inverse, typename FP::B root, typename FP::B* __restrict__ table) {
    typedef typename FP::B B;
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)1 << log_n;
    if (k >= n) return;
    B w = root;  // primitive 2^TWO_ADICITY-th root of unity
    for (int i = log_n; i < FP::TWO_ADICITY; i++) w = FP::b_mul(w, w);
    if (inverse) w = FP::b_inv(w);
    B r = FP::b_one(), b = w;
    size_t e = k;
    while (e) { if (e & 1) r = FP::b_mul(r, b); b = FP::b_mul(b, b); e >>= 1; }
    table[k] = r;
}

__device__ __forceinline__ unsigned bitrev(unsigned x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0; }

// DIF NTT of `tile` interleaved transforms of length n = 2^logn held as s[j*ts + c] (ts = padded tile stride);
// root table: tw[k*tw_stride] = w_n^k
// All sizes are powers of two: index arithmetic is shifts and masks (a runtime division costs more than the butterfly).
template <class FP>
__device__ __forceinline__ void smem_ntt_dif(typename FP::B* s, int logn, int log_tile, int ts, const typename FP::B* __restrict__ tw_global, size_t tw_stride) {
    typedef typename FP::B B;
    const int n = 1 << logn, tmask = (1 << log_tile) - 1;
    const int nbf = (n >> 1) << log_tile;
    // the n/2 twiddles of this transform size, once per CTA, behind the data tile (the launch reserves the space)
    B* tw = s + (size_t)n * ts;
    for (int k = threadIdx.x; k < (n >> 1); k += blockDim.x) tw[k] = tw_global[(size_t)k * tw_stride];
    __syncthreads();
    for (int loglen = logn - 1; loglen >= 0; loglen--) {
        const int len = 1 << loglen;
        const int logstep = logn - 1 - loglen;  // exponent step of this stage = (n/2) / len
        for (int q = threadIdx.x; q < nbf; q += blockDim.x) {
            const int c = q & tmask, bf = q >> log_tile;
            const int j = bf & (len - 1), i = ((bf >> loglen) << (loglen + 1)) | j;
            const B u = s[i * ts + c], v = s[(i + len) * ts + c];
            s[i * ts + c] = FP::b_add(u, v);
            const B d = FP::b_sub(u, v);
            s[(i + len) * ts + c] = j ? FP::fmul(d, tw[j << logstep]) : d;
        }
        __syncthreads();
    }
}

// pass 1. x, y: [batch][N]; tw: w_N^k, k < N
template <class FP>
__global__ void __launch_bounds__(HG_NTT_THREADS) k_ntt_cols(const typename FP::B* __restrict__ x, typename FP::B* __restrict__ y, int log_n1,
                                                             int log_n2, int tile, const typename FP::B* __restrict__ tw) {
    typedef typename FP::B B;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    B* s = reinterpret_cast<B*>(smem_raw);
    const int n1 = 1 << log_n1, n2 = 1 << log_n2;
    const size_t N = (size_t)n1 * n2;
    const int j2_0 = blockIdx.x * tile, ts = tile | 1;
    const int log_tile = 31 - __clz(tile), tmask = tile - 1;
    const B* xb = x + (size_t)blockIdx.y * N;
    B* yb = y + (size_t)blockIdx.y * N;
    for (int q = threadIdx.x; q < n1 * tile; q += blockDim.x) {
        const int c = q & tmask, j1 = q >> log_tile;
        s[j1 * ts + c] = xb[((size_t)j1 << log_n2) + j2_0 + c];
    }
    __syncthreads();
    smem_ntt_dif<FP>(s, log_n1, log_tile, ts, tw, (size_t)n2);  // w_{N1} = w_N^{N2}
    for (int q = threadIdx.x; q < n1 * tile; q += blockDim.x) {
        const int c = q & tmask, pos = q >> log_tile;
        const int k1 = (int)bitrev((unsigned)pos, log_n1);
        const int j2 = j2_0 + c;
        B v = s[pos * ts + c];
        const size_t e = ((size_t)j2 * k1) & (N - 1);
        if (e) v = FP::fmul(v, tw[e]);
        yb[((size_t)k1 << log_n2) + j2] = v;
    }
}

// pass 2. y: [batch][N] (rows k1 of length N2) -> out[k1 + N1*k2]; scale = 1 or N^{-1}
template <class FP>
__global__ void __launch_bounds__(HG_NTT_THREADS) k_ntt_rows(const typename FP::B* __restrict__ y, typename FP::B* __restrict__ out, int log_n1,
                                                             int log_n2, int tile, const typename FP::B* __restrict__ tw, typename FP::B scale,
                                                             int do_scale) {
    typedef typename FP::B B;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    B* s = reinterpret_cast<B*>(smem_raw);
    const int n1 = 1 << log_n1, n2 = 1 << log_n2;
    const size_t N = (size_t)n1 * n2;
    const int k1_0 = blockIdx.x * tile, ts = tile | 1;
    const int log_tile = 31 - __clz(tile), tmask = tile - 1;
    const B* yb = y + (size_t)blockIdx.y * N;
    B* ob = out + (size_t)blockIdx.y * N;
    for (int q = threadIdx.x; q < n2 * tile; q += blockDim.x) {
        const int j2 = q & (n2 - 1), r = q >> log_n2;  // contiguous rows
        s[j2 * ts + r] = yb[((size_t)(k1_0 + r) << log_n2) + j2];
    }
    __syncthreads();
    smem_ntt_dif<FP>(s, log_n2, log_tile, ts, tw, (size_t)n1);  // w_{N2} = w_N^{N1}
    for (int q = threadIdx.x; q < n2 * tile; q += blockDim.x) {
        const int r = q & tmask, pos = q >> log_tile;
        const int k2 = (int)bitrev((unsigned)pos, log_n2);
        B v = s[pos * ts + r];
        if (do_scale) v = FP::fmul(v, scale);
        ob[(size_t)(k1_0 + r) + ((size_t)k2 << log_n1)] = v;
    }
}

// out[i][j] = a[i][j] * b[j]  (broadcast over the batch): the dot-product layer between the FFTs (sk_encryption_circuit.rs:245-250)
template <class FP>
__global__ void k_pointwise_mul_bcast(typename FP::B* __restrict__ a, const typename FP::B* __restrict__ b, size_t n) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    typename FP::B* row = a + (size_t)blockIdx.y * n;
    row[j] = FP::b_mul(row[j], b[j]);
}

}  // namespace hg
