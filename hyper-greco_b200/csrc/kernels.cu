// Non-templated kernels: the order-dependent access counters of LassoNode::polynomialize
// (/root/reference/lasso/src/lasso.rs:177-196). See kernels.cuh for the scheme.
#include "kernels.cuh"

namespace hg {

// ---------------------------------------------------------------------------------------------------------
// Access counters by a stable two-pass radix sort of (address, row) -- scratch and time proportional to the number of rows,
// independent of how the addresses are distributed (range-checked limbs are heavily skewed towards address 0):
//   pass 0: stable partition by the low address byte, pass 1: by the high byte  =>  sorted by address, rows in order inside
//   every address run;  read_cts[row] = position - start of the run,  final_cts[address] = length of the run.
// Element = address << 32 | row. Every kernel handles all chunk slots (blockIdx.y). Tiles of HG_CNT_TILE consecutive
// positions, 1024 threads: thread t of chunk c holds position tile*TILE + c*1024 + t, so (warp, lane) order = row order.
constexpr int HG_CNT_TILE = 4096;
__device__ __forceinline__ bool cnt_row_valid(const CntSlots& sl, int slot, const u8* row_lookup, size_t j, size_t n_rows) {
    if (j >= n_rows) return false;
    const u8 l = row_lookup[j];
    return l != 0xFF && ((sl.used[slot] >> l) & 1);
}
// element at position `pos` of pass `pass` (pass 0 reads the rows themselves), digit = low / high address byte
__device__ __forceinline__ bool cnt_fetch(int pass, const CntSlots& sl, int slot, const u8* row_lookup, size_t n_rows, const u64* src, const u32* n_valid,
                                          size_t pos, u64* elem, u32* digit) {
    if (pass == 0) {
        if (!cnt_row_valid(sl, slot, row_lookup, pos, n_rows)) return false;
        const u32 a = sl.addr[slot][pos];
        *elem = ((u64)a << 32) | (u64)pos;
        *digit = a & 255u;
        return true;
    }
    if (pos >= n_valid[slot]) return false;
    *elem = src[pos];
    *digit = (u32)(*elem >> 40) & 255u;
    return true;
}
__global__ void __launch_bounds__(1024) k_cnt_digit_hist(SlotMap sm, int pass, CntSlots sl, const u8* __restrict__ row_lookup, size_t n_rows, size_t cap, const u64* __restrict__ src_all,
                                                        const u32* __restrict__ n_valid, int nblk, u32* __restrict__ blk_hist_all /*[slot][nblk][256]*/) {
    __shared__ u32 h[256];
    const int slot = sm.s[blockIdx.y];
    if (threadIdx.x < 256) h[threadIdx.x] = 0;
    __syncthreads();
    const u64* src = src_all + (size_t)slot * cap;
    for (int c = 0; c < HG_CNT_TILE / 1024; c++) {
        const size_t pos = (size_t)blockIdx.x * HG_CNT_TILE + c * 1024 + threadIdx.x;
        u64 e; u32 d;
        const bool ok = cnt_fetch(pass, sl, slot, row_lookup, n_rows, src, n_valid, pos, &e, &d);
        // one shared atomic per distinct digit of a warp (address 0 dominates: per-element atomics would serialise)
        const unsigned peers = __match_any_sync(0xffffffffu, ok ? d : 256u);
        if (ok && (peers & ((1u << (threadIdx.x & 31)) - 1)) == 0) atomicAdd(&h[d], (u32)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < 256) blk_hist_all[((size_t)slot * nblk + blockIdx.x) * 256 + threadIdx.x] = h[threadIdx.x];
}
// per (slot, digit): exclusive prefix over tiles. grid = (8, slots): a CTA owns 32 digits, its 1024 threads are 32 groups of
// tiles x 32 digits (separate input and output arrays: the loads do not wait for the stores). Digit totals -> digit_total.
__global__ void __launch_bounds__(1024) k_cnt_digit_scan(SlotMap sm, int nblk, const u32* __restrict__ blk_hist_all, u32* __restrict__ blk_base_all,
                                                        u32* __restrict__ digit_total_all /*[slot][256]*/) {
    __shared__ u32 part[32][33];
    const int slot = sm.s[blockIdx.y], dl = threadIdx.x & 31, grp = threadIdx.x >> 5, d = blockIdx.x * 32 + dl;
    const u32* bh = blk_hist_all + (size_t)slot * nblk * 256;
    u32* bb = blk_base_all + (size_t)slot * nblk * 256;
    const int per = (nblk + 31) / 32, b0 = min(nblk, grp * per), b1 = min(nblk, b0 + per);
    u32 sum = 0;
#pragma unroll 4
    for (int b = b0; b < b1; b++) sum += bh[(size_t)b * 256 + d];
    part[grp][dl] = sum;
    __syncthreads();
    u32 run = 0;
    for (int g = 0; g < grp; g++) run += part[g][dl];
#pragma unroll 4
    for (int b = b0; b < b1; b++) { const u32 v = bh[(size_t)b * 256 + d]; bb[(size_t)b * 256 + d] = run; run += v; }
    if (grp == 31) digit_total_all[slot * 256 + d] = run;  // the last group ends with the grand total (empty groups pass the sum through)
}
// digit totals -> exclusive digit starts, n_valid. One CTA of 256 threads per slot.
__global__ void __launch_bounds__(256) k_cnt_digit_starts(SlotMap sm, const u32* __restrict__ digit_total_all, u32* __restrict__ digit_start_all, u32* __restrict__ n_valid) {
    __shared__ u32 tot[256];
    const int slot = sm.s[blockIdx.x], d = threadIdx.x;
    tot[d] = digit_total_all[slot * 256 + d];
    __syncthreads();
    // inclusive scan in shared memory (Hillis-Steele, 8 steps)
    for (int off = 1; off < 256; off <<= 1) {
        const u32 v = d >= off ? tot[d - off] : 0;
        __syncthreads();
        tot[d] += v;
        __syncthreads();
    }
    digit_start_all[slot * 256 + d] = d ? tot[d - 1] : 0;
    if (d == 255) n_valid[slot] = tot[255];
}
__global__ void __launch_bounds__(1024) k_cnt_digit_scatter(SlotMap sm, int pass, CntSlots sl, const u8* __restrict__ row_lookup, size_t n_rows, size_t cap,
                                                           const u64* __restrict__ src_all, const u32* __restrict__ n_valid, int nblk,
                                                           const u32* __restrict__ blk_base_all, const u32* __restrict__ digit_start_all,
                                                           u64* __restrict__ dst_all) {
    __shared__ unsigned short cnt[32][257];  // per warp and digit: elements of this chunk, then their exclusive prefix over warps
    __shared__ u32 run[256];                 // global offset of (tile, digit) plus the elements of earlier chunks of this tile
    __shared__ u32 tot[256];
    const int slot = sm.s[blockIdx.y], lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64* src = src_all + (size_t)slot * cap;
    u64* dst = dst_all + (size_t)slot * cap;
    if (threadIdx.x < 256)
        run[threadIdx.x] = digit_start_all[slot * 256 + threadIdx.x] + blk_base_all[((size_t)slot * nblk + blockIdx.x) * 256 + threadIdx.x];
    for (int c = 0; c < HG_CNT_TILE / 1024; c++) {
        for (int q = threadIdx.x; q < 32 * 257; q += 1024) (&cnt[0][0])[q] = 0;
        __syncthreads();
        const size_t pos = (size_t)blockIdx.x * HG_CNT_TILE + c * 1024 + threadIdx.x;
        u64 e = 0;
        u32 d = 256;
        const bool ok = cnt_fetch(pass, sl, slot, row_lookup, n_rows, src, n_valid, pos, &e, &d);
        if (!ok) d = 256;
        // rows of a warp are consecutive: the rank among equal digits inside the warp is a popcount of the lower peers
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const u32 lane_rank = (u32)__popc(peers & ((1u << lane) - 1));
        if (ok && lane_rank == 0) cnt[warp][d] = (unsigned short)__popc(peers);
        __syncthreads();
        if (threadIdx.x < 256) {
            u32 s = 0;
            for (int w = 0; w < 32; w++) { const u32 v = cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = (unsigned short)s; s += v; }
            tot[threadIdx.x] = s;
        }
        __syncthreads();
        if (ok) dst[run[d] + cnt[warp][d] + lane_rank] = e;
        __syncthreads();
        if (threadIdx.x < 256) run[threadIdx.x] += tot[threadIdx.x];
    }
}
// sorted: run boundaries of every address
__global__ void k_cnt_heads(SlotMap sm, size_t cap, const u64* __restrict__ sorted_all, const u32* __restrict__ n_valid, size_t M, u32* __restrict__ start_all,
                            u32* __restrict__ end_all) {
    const int slot = sm.s[blockIdx.y];
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 n = n_valid[slot];
    if (p >= n) return;
    const u64* s = sorted_all + (size_t)slot * cap;
    const u32 a = (u32)(s[p] >> 32);
    if (p == 0 || (u32)(s[p - 1] >> 32) != a) start_all[(size_t)slot * M + a] = (u32)p;
    if (p + 1 == n || (u32)(s[p + 1] >> 32) != a) end_all[(size_t)slot * M + a] = (u32)(p + 1);
}
// read_cts[row] = position inside the run of its address; final_cts[address] = run length (0 for addresses never read)
__global__ void k_cnt_finish(SlotMap sm, size_t cap, const u64* __restrict__ sorted_all, const u32* __restrict__ n_valid, size_t M, const u32* __restrict__ start_all,
                             const u32* __restrict__ end_all, size_t R, u32* __restrict__ read_cts_all, u32* __restrict__ final_cts_all) {
    const int slot = sm.s[blockIdx.y];
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < M) final_cts_all[(size_t)slot * M + p] = end_all[(size_t)slot * M + p] - start_all[(size_t)slot * M + p];
    if (p >= n_valid[slot]) return;
    const u64 e = sorted_all[(size_t)slot * cap + p];
    read_cts_all[(size_t)slot * R + (u32)e] = (u32)p - start_all[(size_t)slot * M + (u32)(e >> 32)];
}

}  // namespace hg
