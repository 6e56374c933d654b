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

}  // namespace hg
