// Non-templated kernels: the order-dependent access counters of LassoNode::polynomialize
// (/root/reference/lasso/src/lasso.rs:177-196). See kernels.cuh for the scheme.
#include "kernels.cuh"

namespace hg {

__global__ void k_cnt_hist(const u16* __restrict__ addr, const u8* __restrict__ row_lookup, u64 used_mask, size_t n_rows, int rows_per_block,
                           u16* __restrict__ blk_hist, int log2M) {
    extern __shared__ u32 sh[];  // M/2 words, two 16-bit counters per word
    const size_t M = (size_t)1 << log2M, words = M >> 1;
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

__global__ void k_cnt_scan(const u16* __restrict__ blk_hist, int nblk, int log2M, u32* __restrict__ blk_base, u32* __restrict__ final_cts) {
    const size_t M = (size_t)1 << log2M;
    const size_t a = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= M) return;
    u32 run = 0;
    for (int b = 0; b < nblk; b++) {
        u32 c = blk_hist[(size_t)b * M + a];
        blk_base[(size_t)b * M + a] = run;
        run += c;
    }
    final_cts[a] = run;
}

// Ordered rank. The block first stages its rows in shared memory (coalesced, all warps), then ONE warp walks them in
// order with everything on chip (equal addresses inside a 32-row step are ranked with match_any), then all warps add the
// cross-block base offsets and write the result. blockDim.x must be a multiple of 32; rows_per_block <= HG_CNT_ROWS.
__global__ void k_cnt_rank(const u16* __restrict__ addr, const u8* __restrict__ row_lookup, u64 used_mask, size_t n_rows, size_t R,
                           int rows_per_block, const u32* __restrict__ blk_base, int log2M, u32* __restrict__ read_cts) {
    extern __shared__ u32 sh[];
    const size_t M = (size_t)1 << log2M, words = M >> 1;
    u32* cnt = sh;                                   // M/2 words: two 16-bit counters per word
    u32* key = sh + words;                           // rows_per_block: address, or 0x10000 | lane for unused rows
    u16* rank16 = reinterpret_cast<u16*>(key + rows_per_block);  // rows_per_block: rank inside this block
    const size_t row0 = (size_t)blockIdx.x * rows_per_block;
    const size_t row1 = min(row0 + (size_t)rows_per_block, n_rows);
    const int nrows = (int)(row1 - row0);
    for (size_t i = threadIdx.x; i < words; i += blockDim.x) cnt[i] = 0;
    for (int i = threadIdx.x; i < nrows; i += blockDim.x) {
        const u8 l = row_lookup[row0 + i];
        const bool valid = (l != 0xFF) && ((used_mask >> l) & 1);
        key[i] = valid ? (u32)addr[row0 + i] : (0x10000u | (i & 31));
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int start = 0; start < nrows; start += 32) {
            const int i = start + lane;
            const u32 k = i < nrows ? key[i] : (0x10000u | lane);
            const bool valid = k < 0x10000u;
            const unsigned peers = __match_any_sync(0xffffffffu, k);
            const unsigned rank = __popc(peers & ((1u << lane) - 1));
            const unsigned shift = (k & 1) * 16;
            const u32 local = valid ? ((cnt[k >> 1] >> shift) & 0xFFFFu) : 0;
            __syncwarp();
            if (valid && (31 - __clz(peers)) == lane) atomicAdd(&cnt[k >> 1], (u32)__popc(peers) << shift);
            __syncwarp();
            if (i < nrows) rank16[i] = (u16)(local + rank);
        }
    }
    __syncthreads();
    const u32* base = blk_base + (size_t)blockIdx.x * M;
    for (int i = threadIdx.x; i < nrows; i += blockDim.x) {
        const u32 k = key[i];
        if (k < 0x10000u) read_cts[row0 + i] = base[k] + rank16[i];
    }
}

}  // namespace hg
