// Non-templated kernels: the order-dependent access counters of LassoNode::polynomialize
// (/root/reference/lasso/src/lasso.rs:177-196). See kernels.cuh for the scheme.
#include "kernels.cuh"

namespace hg {

// All kernels below handle every chunk slot in one launch (blockIdx.y = slot): the slots are independent and each of them
// alone does not fill the GPU.
__global__ void k_cnt_hist(CntSlots sl, const u8* __restrict__ row_lookup, size_t n_rows, int rows_per_block, u16* __restrict__ blk_hist_all,
                           int nblk, int log2M) {
    extern __shared__ u32 sh[];  // M/2 words, two 16-bit counters per word
    const size_t M = (size_t)1 << log2M, words = M >> 1;
    const u16* __restrict__ addr = sl.addr[blockIdx.y];
    const u64 used_mask = sl.used[blockIdx.y];
    u16* __restrict__ blk_hist = blk_hist_all + (size_t)blockIdx.y * nblk * M;
    for (size_t i = threadIdx.x; i < words; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const size_t row0 = (size_t)blockIdx.x * rows_per_block;
    const size_t row1 = min(row0 + (size_t)rows_per_block, n_rows);
    for (size_t j = row0 + threadIdx.x; j < row1; j += blockDim.x) {
        u8 l = row_lookup[j];
        if (l != 0xFF && ((used_mask >> l) & 1)) {
            u32 a = addr[j];
            atomicAdd(&sh[a >> 1], 1u << ((a & 1) * 16));
        }
    }
    __syncthreads();
    u32* dst = reinterpret_cast<u32*>(blk_hist + (size_t)blockIdx.x * M);
    for (size_t i = threadIdx.x; i < words; i += blockDim.x) dst[i] = sh[i];
}

__global__ void k_cnt_scan(const u16* __restrict__ blk_hist_all, int nblk, int log2M, u32* __restrict__ blk_base_all, u32* __restrict__ final_cts_all) {
    const size_t M = (size_t)1 << log2M;
    const size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= M) return;
    const u16* __restrict__ blk_hist = blk_hist_all + (size_t)blockIdx.y * nblk * M;
    u32* __restrict__ blk_base = blk_base_all + (size_t)blockIdx.y * nblk * M;
    u32* __restrict__ final_cts = final_cts_all + (size_t)blockIdx.y * M;
    u32 run = 0;
    for (int b = 0; b < nblk; b++) {
        u32 c = blk_hist[(size_t)b * M + a];
        blk_base[(size_t)b * M + a] = run;
        run += c;
    }
    final_cts[a] = run;
}

// Ordered rank inside a block of rows_per_block (<= 4096) rows, fully parallel: sort the keys (address << 12 | local row)
// with a bitonic network in shared memory, find the start of every equal-address run with a max-scan, and the rank of a
// row is its distance from the run start. Cross-block order comes from blk_base (k_cnt_scan). blockDim.x = 1024.
__global__ void __launch_bounds__(1024) k_cnt_rank(CntSlots sl, const u8* __restrict__ row_lookup, size_t n_rows, size_t R, int rows_per_block,
                                                   const u32* __restrict__ blk_base_all, int nblk, int log2M, u32* __restrict__ read_cts_all) {
    constexpr int N = 4096;
    const u16* __restrict__ addr = sl.addr[blockIdx.y];
    const u64 used_mask = sl.used[blockIdx.y];
    const u32* __restrict__ blk_base = blk_base_all + (size_t)blockIdx.y * nblk * ((size_t)1 << log2M);
    u32* __restrict__ read_cts = read_cts_all + (size_t)blockIdx.y * R;
    __shared__ u32 key[N];
    __shared__ u32 runstart[N];
    __shared__ u32 warp_max[32];
    const size_t M = (size_t)1 << log2M;
    const size_t row0 = (size_t)blockIdx.x * rows_per_block;
    const size_t row1 = min(row0 + (size_t)rows_per_block, n_rows);
    const int nrows = (int)(row1 - row0);
    const int tid = threadIdx.x;
    for (int i = tid; i < N; i += blockDim.x) {
        u32 k = 0xFFFFFFFFu;
        if (i < nrows) {
            const u8 l = row_lookup[row0 + i];
            if ((l != 0xFF) && ((used_mask >> l) & 1)) k = ((u32)addr[row0 + i] << 12) | (u32)i;
        }
        key[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < N / 2; t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool up = (i & k) == 0;
                const u32 x = key[i], y = key[p];
                if ((x > y) == up) { key[i] = y; key[p] = x; }
            }
            __syncthreads();
        }
    }
    // run starts: inclusive max-scan of (position if the address differs from the previous one else 0); 4 elements / thread
    u32 loc[4];
    u32 run = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int p = tid * 4 + e;
        const u32 a = key[p] >> 12;
        const bool head = (p == 0) || ((key[p - 1] >> 12) != a);
        run = head ? (u32)p : run;
        loc[e] = head ? (u32)p : 0xFFFFFFFFu;  // marks "inherit"
    }
    // thread-level: last run start in this thread's 4 elements (0 if none started here and none inherited yet)
    u32 tmax = 0;
    bool any = false;
#pragma unroll
    for (int e = 0; e < 4; e++) if (loc[e] != 0xFFFFFFFFu) { tmax = loc[e]; any = true; }
    u32 v = any ? tmax : 0;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, v, off); if (lane >= off) v = max(v, o); }
    if (lane == 31) warp_max[warp] = v;
    __syncthreads();
    if (warp == 0) {
        u32 w = warp_max[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, w, off); if (lane >= off) w = max(w, o); }
        warp_max[lane] = w;
    }
    __syncthreads();
    u32 excl = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) excl = 0;
    if (warp > 0) excl = max(excl, warp_max[warp - 1]);
    u32 cur = excl;  // run start inherited from earlier threads
#pragma unroll
    for (int e = 0; e < 4; e++) {
        if (loc[e] != 0xFFFFFFFFu) cur = loc[e];
        runstart[tid * 4 + e] = cur;
    }
    __syncthreads();
    const u32* base = blk_base + (size_t)blockIdx.x * M;
    for (int p = tid; p < N; p += blockDim.x) {
        const u32 k = key[p];
        if (k != 0xFFFFFFFFu) read_cts[row0 + (k & 0xFFFu)] = base[k >> 12] + ((u32)p - runstart[p]);
    }
}

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
__global__ void __launch_bounds__(1024) k_cnt_digit_hist(int pass, CntSlots sl, const u8* __restrict__ row_lookup, size_t n_rows, size_t cap, const u64* __restrict__ src_all,
                                                        const u32* __restrict__ n_valid, int nblk, u32* __restrict__ blk_hist_all /*[slot][nblk][256]*/) {
    __shared__ u32 h[256];
    const int slot = blockIdx.y;
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
// per (slot, digit): exclusive prefix over tiles, then digit starts; n_valid = total. One CTA of 1024 threads per slot: four
// groups of 256 digits, each group owns a quarter of the tiles (separate input and output arrays: the loads do not wait for
// the stores).
__global__ void __launch_bounds__(1024) k_cnt_digit_scan(int nblk, const u32* __restrict__ blk_hist_all, u32* __restrict__ blk_base_all,
                                                        u32* __restrict__ digit_start_all /*[slot][256]*/, u32* __restrict__ n_valid) {
    __shared__ u32 part[4][256];
    __shared__ u32 tot[256];
    const int slot = blockIdx.x, d = threadIdx.x & 255, grp = threadIdx.x >> 8;
    const u32* bh = blk_hist_all + (size_t)slot * nblk * 256;
    u32* bb = blk_base_all + (size_t)slot * nblk * 256;
    const int per = (nblk + 3) / 4, b0 = grp * per, b1 = min(nblk, b0 + per);
    u32 sum = 0;
#pragma unroll 8
    for (int b = b0; b < b1; b++) sum += bh[(size_t)b * 256 + d];
    part[grp][d] = sum;
    __syncthreads();
    u32 run = 0;
    for (int g = 0; g < grp; g++) run += part[g][d];
#pragma unroll 8
    for (int b = b0; b < b1; b++) { const u32 v = bh[(size_t)b * 256 + d]; bb[(size_t)b * 256 + d] = run; run += v; }
    if (grp == 3) tot[d] = run;  // the last group ends with the grand total of the digit (empty groups pass the sum through)
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 s = 0;
        for (int k = 0; k < 256; k++) { const u32 v = tot[k]; tot[k] = s; s += v; }
        n_valid[slot] = s;
    }
    __syncthreads();
    if (grp == 0) digit_start_all[slot * 256 + d] = tot[d];
}
__global__ void __launch_bounds__(1024) k_cnt_digit_scatter(int pass, CntSlots sl, const u8* __restrict__ row_lookup, size_t n_rows, size_t cap,
                                                           const u64* __restrict__ src_all, const u32* __restrict__ n_valid, int nblk,
                                                           const u32* __restrict__ blk_base_all, const u32* __restrict__ digit_start_all,
                                                           u64* __restrict__ dst_all) {
    __shared__ unsigned short cnt[32][257];  // per warp and digit: elements of this chunk, then their exclusive prefix over warps
    __shared__ u32 run[256];                 // global offset of (tile, digit) plus the elements of earlier chunks of this tile
    __shared__ u32 tot[256];
    const int slot = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
__global__ void k_cnt_heads(size_t cap, const u64* __restrict__ sorted_all, const u32* __restrict__ n_valid, size_t M, u32* __restrict__ start_all,
                            u32* __restrict__ end_all) {
    const int slot = blockIdx.y;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 n = n_valid[slot];
    if (p >= n) return;
    const u64* s = sorted_all + (size_t)slot * cap;
    const u32 a = (u32)(s[p] >> 32);
    if (p == 0 || (u32)(s[p - 1] >> 32) != a) start_all[(size_t)slot * M + a] = (u32)p;
    if (p + 1 == n || (u32)(s[p + 1] >> 32) != a) end_all[(size_t)slot * M + a] = (u32)(p + 1);
}
// read_cts[row] = position inside the run of its address; final_cts[address] = run length (0 for addresses never read)
__global__ void k_cnt_finish(size_t cap, const u64* __restrict__ sorted_all, const u32* __restrict__ n_valid, size_t M, const u32* __restrict__ start_all,
                             const u32* __restrict__ end_all, size_t R, u32* __restrict__ read_cts_all, u32* __restrict__ final_cts_all) {
    const int slot = blockIdx.y;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < M) final_cts_all[(size_t)slot * M + p] = end_all[(size_t)slot * M + p] - start_all[(size_t)slot * M + p];
    if (p >= n_valid[slot]) return;
    const u64 e = sorted_all[(size_t)slot * cap + p];
    read_cts_all[(size_t)slot * R + (u32)e] = (u32)p - start_all[(size_t)slot * M + (u32)(e >> 32)];
}

}  // namespace hg
