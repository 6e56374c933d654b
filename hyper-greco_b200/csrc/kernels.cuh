// sm_100a kernels for the Lasso node hot path. Templated on a field policy FP
// (field_policy.cuh). All are HBM-streaming integer kernels: coalesced 64/128-bit loads, 64x64->128 IMAD
// reduction, warp-shuffle + block-tree reduction of round-polynomial evaluations, no tensor cores.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "field_policy.cuh"

namespace hg {

typedef unsigned short u16;
typedef unsigned char u8;

constexpr int HG_MAX_LOOKUPS = 64;
constexpr int HG_MAX_MEMORIES = 128;
constexpr int HG_MAX_C = 8;
constexpr int HG_BLOCK = 256;

// Flattened LassoPreprocessing maps (lasso.rs:513-523) for device use.
struct NodeMeta {
    int C, log2M, num_lookups, num_memories;
    u8 total_bits[HG_MAX_LOOKUPS];            // sum(chunk_bits) per lookup type (lasso.rs:389, range.rs:234-250)
    u8 lookup_nmem[HG_MAX_LOOKUPS];
    u8 lookup_mem[HG_MAX_LOOKUPS][HG_MAX_C];  // lookup_to_memory_indices (lasso.rs:590-602)
    u8 mem_sub[HG_MAX_MEMORIES];              // memory_to_subtable_index
    u8 mem_dim[HG_MAX_MEMORIES];              // memory_to_dimension_index
    u64 mem_used[HG_MAX_MEMORIES];            // bit l = lookup type l touches this memory (lasso.rs:182-183)
};

// ---------------------------------------------------------------------------------------------------------
// block reduction of NP extension values + cross-block finalisation by the last block to arrive.
// partials: [gridDim.y][gridDim.x][NP]; counter: [gridDim.y]; out: [gridDim.y][NP]
// FLAT: a single output summed over the whole (x, y) grid (partials: [gridDim.x*gridDim.y][NP], one counter).
template <class FP, int NP, bool FLAT = false>
__device__ __forceinline__ void block_reduce_finalize(typename FP::X (&acc)[NP], typename FP::X* partials, unsigned* counter,
                                                      typename FP::X* out) {
    typedef typename FP::X X;
    __shared__ X sm[32][NP];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    const unsigned nblk = FLAT ? gridDim.x * gridDim.y : gridDim.x;
    const unsigned bid = FLAT ? blockIdx.y * gridDim.x + blockIdx.x : blockIdx.x;
    const unsigned slot = FLAT ? 0 : blockIdx.y;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        X v = acc[p];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = FP::x_add(v, FP::x_shfl_down(v, off));
        if (lane == 0) sm[warp][p] = v;
    }
    __syncthreads();
    X* my_part = partials + ((size_t)slot * nblk + bid) * NP;
    if (warp == 0) {
#pragma unroll
        for (int p = 0; p < NP; p++) {
            X v = lane < nwarps ? sm[lane][p] : FP::x_zero();
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v = FP::x_add(v, FP::x_shfl_down(v, off));
            if (lane == 0) my_part[p] = v;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicAdd(counter + slot, 1u);
        is_last = (t == nblk - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const X* all = partials + (size_t)slot * nblk * NP;
    X s[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) s[p] = FP::x_zero();
    for (unsigned b = threadIdx.x; b < nblk; b += blockDim.x)
#pragma unroll
        for (int p = 0; p < NP; p++) s[p] = FP::x_add(s[p], FP::x_ldcg(all + (size_t)b * NP + p));
    __syncthreads();
#pragma unroll
    for (int p = 0; p < NP; p++) {
        X v = s[p];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = FP::x_add(v, FP::x_shfl_down(v, off));
        if (lane == 0) sm[warp][p] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int p = 0; p < NP; p++) {
            X v = lane < nwarps ? sm[lane][p] : FP::x_zero();
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v = FP::x_add(v, FP::x_shfl_down(v, off));
            if (lane == 0) out[(size_t)slot * NP + p] = v;
        }
        if (lane == 0) counter[slot] = 0;  // re-arm for the next launch that uses this slot
    }
}

// ---------------------------------------------------------------------------------------------------------
// K1 + K2(E) + collation pre-combination.  One thread per lookup row (lasso.rs:157-250, :381-414).
//   dims[c][j]   16-bit limbs of the (truncated) input        -> u16 [C][R]
//   E[mi][j]     T_sub(mi)[dims[dim(mi)][j]] or 0              -> B   [m][R]
//   S[j]         sum_i coeff[i] * E[i][j]   (collation inner sum, lasso.rs:457-475, is multilinear in the E tables)
//   out[j]       combine_lookups of the row's lookup type      (lasso.rs:438-449, range.rs:184-195)
template <class FP>
__global__ void k_polynomialize(const typename FP::B* __restrict__ inputs, size_t n_rows, const u8* __restrict__ row_lookup,
                                const NodeMeta* __restrict__ meta, const typename FP::B* __restrict__ subtables,
                                const typename FP::B* __restrict__ coll_coeff, const typename FP::B* __restrict__ wpow, size_t R,
                                u16* __restrict__ dims, typename FP::B* __restrict__ E, typename FP::B* __restrict__ S,
                                typename FP::B* __restrict__ out) {
    typedef typename FP::B B;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= R) return;
    const int C = meta->C, log2M = meta->log2M, m = meta->num_memories;
    const size_t M = (size_t)1 << log2M;
    int l = 0xFF;
    u64 x = 0;
    if (j < n_rows) {
        l = row_lookup[j];
        B in = inputs[j];
        x = FP::b_low_u64(in);
        int tb = meta->total_bits[l];
        if (tb < 64) x &= (((u64)1) << tb) - 1;
    }
    u16 d[HG_MAX_C];
    for (int c = 0; c < C; c++) {
        u64 v = (c * log2M < 64) ? ((x >> (c * log2M)) & (M - 1)) : 0;
        d[c] = (l == 0xFF) ? 0 : (u16)v;
        dims[(size_t)c * R + j] = d[c];
    }
    // sums of products are accumulated unreduced (field_policy.cuh): one reduction for S, one for the lookup output
    typename FP::BAcc s = FP::bacc_zero();
    for (int mi = 0; mi < m; mi++) {
        B e = FP::b_zero();
        if (l != 0xFF && ((meta->mem_used[mi] >> l) & 1)) {
            e = subtables[(size_t)meta->mem_sub[mi] * M + d[meta->mem_dim[mi]]];
            FP::bacc_mad(s, coll_coeff[mi], e);
        }
        E[(size_t)mi * R + j] = e;
    }
    S[j] = FP::bacc_reduce(s);
    typename FP::BAcc o = FP::bacc_zero();
    if (l != 0xFF) {
        int nm = meta->lookup_nmem[l];
        for (int t = 0; t < nm; t++) {
            int mi = meta->lookup_mem[l][t];
            FP::bacc_mad(o, wpow[l * HG_MAX_C + t], subtables[(size_t)meta->mem_sub[mi] * M + d[meta->mem_dim[mi]]]);  // combine_lookups of the row's own lookup type
        }
    }
    out[j] = FP::bacc_reduce(o);
}

// ---------------------------------------------------------------------------------------------------------
// K2 counters (lasso.rs:177-196): read_cts[j] = number of earlier rows that touch the same address of memory `mem`,
// final_cts[a] = total. Order-dependent: a stable two-pass radix sort of (address, row) groups the rows of an address in row
// order; the counters are read off the run boundaries (kernels.cu).
struct CntSlots { const u16* addr[HG_MAX_C]; u64 used[HG_MAX_C]; };  // per chunk slot: its address column and the lookup types that use it
struct SlotMap { int s[HG_MAX_C]; };  // blockIdx.y (.x in k_cnt_digit_starts) -> chunk slot: a device of a sharded proof counts only the slots it needs
__global__ void k_cnt_digit_hist(SlotMap sm, int pass, CntSlots sl, const u8* __restrict__ row_lookup, size_t n_rows, size_t cap, const u64* __restrict__ src,
                                 const u32* __restrict__ n_valid, int nblk, u32* __restrict__ blk_hist /*[slot][nblk][256]*/);
__global__ void k_cnt_digit_scan(SlotMap sm, int nblk, const u32* __restrict__ blk_hist, u32* __restrict__ blk_base, u32* __restrict__ digit_total /*[slot][256]*/);
__global__ void k_cnt_digit_starts(SlotMap sm, const u32* __restrict__ digit_total, u32* __restrict__ digit_start /*[slot][256]*/, u32* __restrict__ n_valid);
__global__ void k_cnt_digit_scatter(SlotMap sm, int pass, CntSlots sl, const u8* __restrict__ row_lookup, size_t n_rows, size_t cap, const u64* __restrict__ src,
                                    const u32* __restrict__ n_valid, int nblk, const u32* __restrict__ blk_base, const u32* __restrict__ digit_start,
                                    u64* __restrict__ dst);
__global__ void k_cnt_heads(SlotMap sm, size_t cap, const u64* __restrict__ sorted, const u32* __restrict__ n_valid, size_t M, u32* __restrict__ start, u32* __restrict__ end);
__global__ void k_cnt_finish(SlotMap sm, size_t cap, const u64* __restrict__ sorted, const u32* __restrict__ n_valid, size_t M, const u32* __restrict__ start,
                             const u32* __restrict__ end, size_t R, u32* __restrict__ read_cts, u32* __restrict__ final_cts);

// ---------------------------------------------------------------------------------------------------------
// eq(point, k) = prod_i (k_i ? r_i : 1 - r_i), k_0 = LSB (plonkish MultilinearPolynomial::eq_xy, lasso.rs:432), kept as
// two small factor tables: eq(point, k) = eq_lo[k mod 2^lo_bits] * eq_hi[k >> lo_bits]
template <class FP>
__global__ void k_eq_split(const typename FP::X* __restrict__ point, int nv, int lo_bits, typename FP::X* __restrict__ eq_lo,
                           typename FP::X* __restrict__ eq_hi) {
    typedef typename FP::X X;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nlo = (size_t)1 << lo_bits, nhi = (size_t)1 << (nv - lo_bits);
    if (t >= nlo + nhi) return;
    const bool hi = t >= nlo;
    const size_t k = hi ? t - nlo : t;
    const int first = hi ? lo_bits : 0, last = hi ? nv : lo_bits;
    X acc = FP::x_one();
    for (int i = first; i < last; i++) {
        X r = point[i];
        X f = ((k >> (i - first)) & 1) ? r : FP::x_sub(FP::x_one(), r);
        acc = FP::x_mul(acc, f);
    }
    (hi ? eq_hi : eq_lo)[k] = acc;
}

template <class FP, class T> struct ToBase { __device__ __forceinline__ static typename FP::B f(T v) { return FP::to_base(v); } };

// Batched MLE evaluation (mod.rs:80-93, lasso.rs:422-454): out[y] = sum_k eq(point, k) * tables[y][k], one streaming pass
// per table. Row kh of 2^lo_bits elements is reduced against eq_lo with unreduced accumulation, then scaled by eq_hi[kh].
// grid = (blocks, ntables)
// four consecutive table entries with one load where the type allows it
template <class FP, class T> struct Load4 {
    __device__ __forceinline__ static void f(const T* p, typename FP::B (&v)[4]) {
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = FP::to_base(p[e]);
    }
};
template <class FP> struct Load4<FP, unsigned short> {
    __device__ __forceinline__ static void f(const unsigned short* p, typename FP::B (&v)[4]) {
        const ushort4 t = __ldg(reinterpret_cast<const ushort4*>(p));
        v[0] = FP::to_base(t.x); v[1] = FP::to_base(t.y); v[2] = FP::to_base(t.z); v[3] = FP::to_base(t.w);
    }
};
template <class FP> struct Load4<FP, unsigned int> {
    __device__ __forceinline__ static void f(const unsigned int* p, typename FP::B (&v)[4]) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
        v[0] = FP::to_base(t.x); v[1] = FP::to_base(t.y); v[2] = FP::to_base(t.z); v[3] = FP::to_base(t.w);
    }
};
template <class FP> struct Load4<FP, u64> {
    __device__ __forceinline__ static void f(const u64* p, typename FP::B (&v)[4]) {
        const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2*>(p)), b = __ldg(reinterpret_cast<const ulonglong2*>(p) + 1);
        v[0] = FP::to_base((u64)a.x); v[1] = FP::to_base((u64)a.y); v[2] = FP::to_base((u64)b.x); v[3] = FP::to_base((u64)b.y);
    }
};
template <class FP, class T>
__global__ void __launch_bounds__(HG_BLOCK) k_dot_eq(const T* __restrict__ tables, size_t stride, size_t n, int lo_bits, const typename FP::X* __restrict__ eq_lo,
                         const typename FP::X* __restrict__ eq_hi, typename FP::X* partials, unsigned* counter, typename FP::X* out) {
    typedef typename FP::X X;
    typedef typename FP::B B;
    const T* t = tables + (size_t)blockIdx.y * stride;
    const size_t nlo = (size_t)1 << lo_bits, nhi = n >> lo_bits;
    X acc[1] = {FP::x_zero()};
    if (sizeof(X) <= 16 && nlo == 16 * (size_t)blockDim.x) {  // (64 registers of cached eq_lo for Goldilocks; too many for a 32-byte field)
        // the usual shape (2^12 low entries, 256 threads): a thread always meets the same 16 entries of eq_lo, so they live in
        // registers for the whole launch and a row costs only its own 4 loads of 4 table entries, all issued up front
        X el[4][4];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int e = 0; e < 4; e++) el[u][e] = eq_lo[4 * (size_t)threadIdx.x + (size_t)u * 4 * blockDim.x + e];
        for (size_t kh = blockIdx.x; kh < nhi; kh += gridDim.x) {
            const T* row = t + kh * nlo;
            B v[4][4];
#pragma unroll
            for (int u = 0; u < 4; u++) Load4<FP, T>::f(row + 4 * (size_t)threadIdx.x + (size_t)u * 4 * blockDim.x, v[u]);
            typename FP::XAcc a = FP::xacc_zero_();
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int e = 0; e < 4; e++) FP::xacc_mad_b(a, el[u][e], v[u][e]);
            acc[0] = FP::x_add(acc[0], FP::fmul(FP::xacc_reduce_(a), eq_hi[kh]));
        }
    } else {
        for (size_t kh = blockIdx.x; kh < nhi; kh += gridDim.x) {
            typename FP::XAcc a = FP::xacc_zero_();
            const T* row = t + kh * nlo;
            if (nlo >= 4) {
                for (size_t kl = 4 * (size_t)threadIdx.x; kl < nlo; kl += 4 * (size_t)blockDim.x) {
                    B v[4];
                    Load4<FP, T>::f(row + kl, v);
#pragma unroll
                    for (int e = 0; e < 4; e++) FP::xacc_mad_b(a, eq_lo[kl + e], v[e]);
                }
            } else {
                for (size_t kl = threadIdx.x; kl < nlo; kl += blockDim.x) FP::xacc_mad_b(a, eq_lo[kl], ToBase<FP, T>::f(row[kl]));
            }
            acc[0] = FP::x_add(acc[0], FP::fmul(FP::xacc_reduce_(a), eq_hi[kh]));
        }
    }
    block_reduce_finalize<FP, 1>(acc, partials, counter, out);
}

// Vectors a device needs when one proof is split over several devices (LassoNodeDev::prove_shard): the ones it owns and
// vector 0, whose first half is the common factor t_0 of every term. All of them when the proof is not split.
struct VecRange { int lo, hi; };
__device__ __forceinline__ bool vec_needed(VecRange r, int v) { return v == 0 || (v >= r.lo && v < r.hi); }

// two consecutive base elements with one store where the type allows it (p 16-byte aligned)
__device__ __forceinline__ void store_pair(u64* p, u64 a, u64 b) { *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a, b); }
template <class T> __device__ __forceinline__ void store_pair(T* p, const T& a, const T& b) { p[0] = a; p[1] = b; }

// ---------------------------------------------------------------------------------------------------------
// K6 multiset hashes (prover.rs:35-89): h(a,v,t) = a + v*gamma + t*gamma^2 - tau, gamma/tau truncated to the base
// field (prover.rs:38-39). V = [reads (m, chunk-major) | writes (m)] x R.
template <class FP>
__global__ void k_hash_rw(const u16* __restrict__ dims, const u32* __restrict__ read_cts, const typename FP::B* __restrict__ E,
                          const int* __restrict__ pos_mem, const int* __restrict__ pos_dim, const int* __restrict__ pos_slot,
                          const typename FP::X* __restrict__ gamma_tau, size_t R, int m, typename FP::B* __restrict__ V) {
    typedef typename FP::B B;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int pos = blockIdx.y;
    if (j >= R) return;
    const B gamma = FP::x_base0(gamma_tau[0]), tau = FP::x_base0(gamma_tau[1]), gamma2 = FP::b_mul(gamma, gamma);
    B a = FP::b_from_u64(dims[(size_t)pos_dim[pos] * R + j]);
    B t = FP::b_from_u64(read_cts[(size_t)pos_slot[pos] * R + j]);
    B e = E[(size_t)pos_mem[pos] * R + j];
    B rd = FP::b_sub(FP::b_add(FP::b_add(a, FP::b_mul(e, gamma)), FP::b_mul(t, gamma2)), tau);
    V[(size_t)pos * R + j] = rd;
    V[(size_t)(m + pos) * R + j] = FP::b_add(rd, gamma2);
}
// Same, fused with the first product-tree level (prover.rs:332-354): each thread hashes rows j and j + R/2, so the bottom
// layer is written once and never re-read to build layer 1.  V = layer 0 ([2m][R]), up = layer 1 ([2m][R/2]).
template <class FP>
__global__ void k_hash_rw_up(const u16* __restrict__ dims, const u32* __restrict__ read_cts, const typename FP::B* __restrict__ E,
                             const int* __restrict__ pos_mem, const int* __restrict__ pos_dim, const int* __restrict__ pos_slot,
                             const typename FP::X* __restrict__ gamma_tau, size_t R, int m, typename FP::B* __restrict__ V,
                             typename FP::B* __restrict__ up, VecRange own) {
    typedef typename FP::B B;
    const size_t h = R / 2;
    const int pos = blockIdx.y;
    const bool need_r = vec_needed(own, pos), need_w = vec_needed(own, m + pos);  // vector pos = reads, m + pos = writes of this memory
    if (!need_r && !need_w) return;
    const B gamma = FP::x_base0(gamma_tau[0]), tau = FP::x_base0(gamma_tau[1]), gamma2 = FP::b_mul(gamma, gamma);
    const u16* dm = dims + (size_t)pos_dim[pos] * R;
    const u32* ts = read_cts + (size_t)pos_slot[pos] * R;
    const B* e = E + (size_t)pos_mem[pos] * R;
    // a + e*gamma + t*gamma^2 - tau with one reduction (unreduced accumulator, field_policy.cuh)
    auto hash = [&](size_t q) {
        typename FP::BAcc acc = FP::bacc_zero();
        FP::bacc_mad(acc, e[q], gamma);
        FP::bacc_mad(acc, FP::b_from_u64(ts[q]), gamma2);
        return FP::b_sub(FP::b_add(FP::bacc_reduce(acc), FP::b_from_u64(dm[q])), tau);
    };
    if constexpr (sizeof(B) > 8) {  // wide fields: one row per thread (register pressure)
        for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < h; j += (size_t)gridDim.x * blockDim.x) {
            B r0 = hash(j), r1 = hash(j + h);
            B w0 = FP::b_add(r0, gamma2), w1 = FP::b_add(r1, gamma2);
            if (need_r) { V[(size_t)pos * R + j] = r0; V[(size_t)pos * R + j + h] = r1; up[(size_t)pos * h + j] = FP::fmul(r0, r1); }
            if (need_w) { V[(size_t)(m + pos) * R + j] = w0; V[(size_t)(m + pos) * R + j + h] = w1; up[(size_t)(m + pos) * h + j] = FP::fmul(w0, w1); }
        }
        return;
    }
    // two neighbouring rows per thread: 16-byte stores of the bottom layer and of layer 1
    const size_t j0 = 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (j0 >= h) return;
    B rd[2][2], wr[2][2];  // [row j0 + u][half s]
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const size_t q = j0 + u + s * h;
            if (j0 + u < h) { rd[u][s] = hash(q); wr[u][s] = FP::b_add(rd[u][s], gamma2); }
            else { rd[u][s] = FP::b_zero(); wr[u][s] = FP::b_zero(); }
        }
    B* Vr = V + (size_t)pos * R;
    B* Vw = V + (size_t)(m + pos) * R;
    B* ur = up + (size_t)pos * h;
    B* uw = up + (size_t)(m + pos) * h;
    if (j0 + 1 < h) {
        if (need_r) {
#pragma unroll
            for (int s = 0; s < 2; s++) store_pair(Vr + j0 + s * h, rd[0][s], rd[1][s]);
            store_pair(ur + j0, FP::fmul(rd[0][0], rd[0][1]), FP::fmul(rd[1][0], rd[1][1]));
        }
        if (need_w) {
#pragma unroll
            for (int s = 0; s < 2; s++) store_pair(Vw + j0 + s * h, wr[0][s], wr[1][s]);
            store_pair(uw + j0, FP::fmul(wr[0][0], wr[0][1]), FP::fmul(wr[1][0], wr[1][1]));
        }
    } else {  // odd tail (h is a power of two >= 2 in practice, so this is h == 1 only)
        for (int s = 0; s < 2; s++) { Vr[j0 + s * h] = rd[0][s]; Vw[j0 + s * h] = wr[0][s]; }
        ur[j0] = FP::fmul(rd[0][0], rd[0][1]);
        uw[j0] = FP::fmul(wr[0][0], wr[0][1]);
    }
}
// V2 = [inits (m) | final_reads (m)] x M
template <class FP>
__global__ void k_hash_if(const typename FP::B* __restrict__ subtables, const u32* __restrict__ final_cts, const int* __restrict__ pos_sub,
                          const int* __restrict__ pos_slot, const typename FP::X* __restrict__ gamma_tau, size_t M, int m,
                          typename FP::B* __restrict__ V, VecRange own) {
    typedef typename FP::B B;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int pos = blockIdx.y;
    if (i >= M || !(vec_needed(own, pos) || vec_needed(own, m + pos))) return;
    const B gamma = FP::x_base0(gamma_tau[0]), tau = FP::x_base0(gamma_tau[1]), gamma2 = FP::b_mul(gamma, gamma);
    B v = subtables[(size_t)pos_sub[pos] * M + i];
    B in = FP::b_sub(FP::b_add(FP::b_from_u64(i), FP::b_mul(v, gamma)), tau);
    B fc = FP::b_from_u64(final_cts[(size_t)pos_slot[pos] * M + i]);
    V[(size_t)pos * M + i] = in;
    V[(size_t)(m + pos) * M + i] = FP::b_add(in, FP::b_mul(fc, gamma2));
}

// K7 product-tree layer (prover.rs:332-354): out[i][k] = in[i][k] * in[i][k + h]; halves = top index bit
template <class FP>
__global__ void k_tree_up(const typename FP::B* __restrict__ in, typename FP::B* __restrict__ out, size_t h, VecRange own) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= h || !vec_needed(own, blockIdx.y)) return;
    const typename FP::B* v = in + (size_t)blockIdx.y * 2 * h;
    out[(size_t)blockIdx.y * h + k] = FP::fmul(v[k], v[k + h]);
}
// two levels per launch: in [nvec][4q] -> out1 [nvec][2q] -> out2 [nvec][q]; the intermediate layer is written but not re-read
__device__ __forceinline__ void load_pair(const u64* p, u64& a, u64& b) { const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(p); a = t.x; b = t.y; }
template <class T> __device__ __forceinline__ void load_pair(const T* p, T& a, T& b) { a = p[0]; b = p[1]; }
template <class FP>
__global__ void k_tree_up2(const typename FP::B* __restrict__ in, typename FP::B* __restrict__ out1, typename FP::B* __restrict__ out2, size_t q,
                           VecRange own) {
    typedef typename FP::B B;
    if (!vec_needed(own, blockIdx.y)) return;
    const B* v = in + (size_t)blockIdx.y * 4 * q;
    B* o1 = out1 + (size_t)blockIdx.y * 2 * q;
    B* o2 = out2 + (size_t)blockIdx.y * q;
    if (q >= 2) {  // two neighbouring entries per thread: 16-byte accesses
        const size_t k = 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
        if (k >= q) return;
        B x[4][2];
#pragma unroll
        for (int s = 0; s < 4; s++) load_pair(v + k + s * q, x[s][0], x[s][1]);
        B a[2], b[2];
#pragma unroll
        for (int u = 0; u < 2; u++) { a[u] = FP::fmul(x[0][u], x[2][u]); b[u] = FP::fmul(x[1][u], x[3][u]); }
        store_pair(o1 + k, a[0], a[1]);
        store_pair(o1 + k + q, b[0], b[1]);
        store_pair(o2 + k, FP::fmul(a[0], b[0]), FP::fmul(a[1], b[1]));
    } else {
        const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (k >= q) return;
        const B a = FP::fmul(v[k], v[k + 2 * q]), b = FP::fmul(v[k + q], v[k + 3 * q]);
        o1[k] = a; o1[k + q] = b;
        o2[k] = FP::fmul(a, b);
    }
}
// every remaining level of a tree whose vectors have at most HG_TREE_TAIL elements, one CTA per vector, levels in shared
// memory: layer k (vector length len) -> layers k+1, k+2, ... down to vector length 2. Layers are stored back to back:
// layer[k+1] = layer[k] + nvec * len(k).
constexpr int HG_TREE_TAIL = 2048;
template <class FP>
__global__ void __launch_bounds__(256) k_tree_tail(typename FP::B* __restrict__ layer, int nvec, int len, VecRange own) {
    typedef typename FP::B B;
    extern __shared__ __align__(16) unsigned char tree_smem[];
    if (!vec_needed(own, blockIdx.x)) return;
    B* cur = reinterpret_cast<B*>(tree_smem);  // [len]
    const int vec = blockIdx.x;
    for (int e = threadIdx.x; e < len; e += blockDim.x) cur[e] = layer[(size_t)vec * len + e];
    __syncthreads();
    B* base = layer;
    while (len > 2) {
        const int h = len / 2;
        B* next = base + (size_t)nvec * len;
        B keep[(HG_TREE_TAIL / 2 + 255) / 256];
        int cnt = 0;
        for (int e = threadIdx.x; e < h; e += blockDim.x) keep[cnt++] = FP::fmul(cur[e], cur[e + h]);
        __syncthreads();
        cnt = 0;
        for (int e = threadIdx.x; e < h; e += blockDim.x) { const B v = keep[cnt++]; cur[e] = v; next[(size_t)vec * h + e] = v; }
        __syncthreads();
        base = next;
        len = h;
    }
}
// roots (prover.rs:197-203) and the nv = 0 layer's evaluations (prover.rs:232-236) from the top layer [nvec][2]
template <class FP>
__global__ void k_tree_top(const typename FP::B* __restrict__ top, int i_begin, int i_end, typename FP::X* __restrict__ roots,
                           typename FP::X* __restrict__ evals) {
    const int i = i_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i_end) return;
    typename FP::B l = top[2 * i], r = top[2 * i + 1];
    roots[i] = FP::lift(FP::b_mul(l, r));
    evals[2 * i] = FP::lift(l);
    evals[2 * i + 1] = FP::lift(r);
}

// ---------------------------------------------------------------------------------------------------------
// K5 (and the generic hg_sumcheck_prove): one sumcheck round of g = t_0 * sum_i coeff[i] * prod_{k<ARITY} t_{ARITY*i+k}
// (lasso.rs:457-475, prover.rs:268-279), evaluation FUSED with the fold by the previous round's challenge, so every table
// is read once and its folded image written once per round.
//   in : ntab tables of n_in elements (TIN = B in rounds 0/1, X later), table t at in + t*n_in
//   FOLD: out[t][q] = in[t][2q] + r_prev*(in[t][2q+1] - in[t][2q]) is written (n_in/2 per table) and the round
//         polynomial is sampled on the folded values; !FOLD (= round 0): sampled on `in` directly.
//   msg: h(0), h(inf) [, h(-1) when ARITY == 2] and, in round 0 only, h(1) last. X = lowest remaining variable (A4).
//   The host tracks the true running sum h(0) + h(1) and reconstructs the polynomial (prover.cuh emit_round).
template <class FP, class TIN, int ARITY, bool FOLD>
__global__ void __launch_bounds__(HG_BLOCK)
k_sc_round(const TIN* __restrict__ in, typename FP::X* __restrict__ out, size_t n_in, int nterm,
           const typename FP::X* __restrict__ coeffs, const typename FP::X* __restrict__ r_prev, typename FP::X* partials,
           unsigned* counter, typename FP::X* msg) {
    typedef typename FP::X X;
    constexpr int D = ARITY + 1;
    constexpr int NP = FOLD ? D : D + 1;
    typedef typename std::conditional<FOLD, X, TIN>::type EL;  // element type the round polynomial is sampled on
    const size_t npairs = FOLD ? n_in / 4 : n_in / 2;
    const size_t n_out = n_in / 2;
    X acc[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) acc[p] = FP::x_zero();
    X r = FP::x_zero();
    if (FOLD) r = *r_prev;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < npairs; b += (size_t)gridDim.x * blockDim.x) {
        X inner[NP];
        EL t0[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) inner[p] = FP::x_zero();
        for (int i = 0; i < nterm; i++) {
            EL prod[NP];
#pragma unroll
            for (int k = 0; k < ARITY; k++) {
                const int t = ARITY * i + k;
                EL lo, hi;
                if constexpr (FOLD) {
                    const TIN* src = in + (size_t)t * n_in + 4 * b;
                    TIN a0 = src[0], a1 = src[1], a2 = src[2], a3 = src[3];
                    lo = FP::x_add(FP::as_x(a0), FP::mul(r, FP::sub(a1, a0)));
                    hi = FP::x_add(FP::as_x(a2), FP::mul(r, FP::sub(a3, a2)));
                    X* dst = out + (size_t)t * n_out + 2 * b;
                    dst[0] = lo;
                    dst[1] = hi;
                } else {
                    const TIN* src = in + (size_t)t * n_in + 2 * b;
                    lo = src[0];
                    hi = src[1];
                }
                EL v[NP];
                EL df = FP::sub(hi, lo);
                v[0] = lo;
                v[1] = df;
                if constexpr (D == 3) v[2] = FP::sub(lo, df);
                if constexpr (!FOLD) v[D] = hi;
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    if (k == 0) prod[p] = v[p]; else prod[p] = FP::mul(prod[p], v[p]);
                    if (i == 0 && k == 0) t0[p] = v[p];
                }
            }
            X c = coeffs[i];
#pragma unroll
            for (int p = 0; p < NP; p++) inner[p] = FP::x_add(inner[p], FP::mul(c, prod[p]));
        }
#pragma unroll
        for (int p = 0; p < NP; p++) acc[p] = FP::x_add(acc[p], FP::mul(inner[p], t0[p]));
    }
    block_reduce_finalize<FP, NP>(acc, partials, counter, msg);
}

// final fold of 2-element tables: evals[t] = in[t][0] + r*(in[t][1] - in[t][0])   (prove_sum_check's returned evals)
template <class FP, class TIN>
__global__ void k_fold_final(const TIN* __restrict__ in, int ntab, const typename FP::X* __restrict__ r, typename FP::X* __restrict__ evals) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntab) return;
    TIN a0 = in[2 * t], a1 = in[2 * t + 1];
    evals[t] = FP::x_add(FP::as_x(a0), FP::mul(*r, FP::sub(a1, a0)));
}
// plain fold (used by standalone MLE evaluation / tests): out[q] = in[2q] + r*(in[2q+1]-in[2q])
template <class FP, class TIN>
__global__ void k_fold(const TIN* __restrict__ in, size_t n_out_total, const typename FP::X* __restrict__ r, typename FP::X* __restrict__ out) {
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_out_total) return;
    TIN a0 = in[2 * q], a1 = in[2 * q + 1];
    out[q] = FP::x_add(FP::as_x(a0), FP::mul(*r, FP::sub(a1, a0)));
}

template <class T> __global__ void k_fill(T* p, size_t n, T v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace hg
