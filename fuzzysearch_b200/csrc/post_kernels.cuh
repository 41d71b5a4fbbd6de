// post_kernels.cuh -- on-device ordering and consolidation of the raw match list (small lists).
//
// Two kernels run behind the verify kernel, on the same stream, and replace the host's sorts for
// lists of up to kPostMax records (the common case: matches are sparse); nothing waits for the host:
//   k_rank_scatter   orders the records by all-pairs ranking (rank = number of records with a smaller
//                    key; O(n^2) compares spread over the whole GPU -- 64 CTAs, shared-memory tiles --
//                    which for n <= 16K beats any multi-pass sort that has to synchronise):
//                    * the raw stream in the reference's generation order (n-gram ordinal, hit index)
//                      -- or canonical (start, end, dist) order for the unanchored routes;
//                    * the packed canonical keys in (start, end, dist) order for the consolidation.
//   k_consolidate    consolidate_overlapping_matches (common.py:145-189) on the sorted keys: running
//                    maximum of `end` as the hull, a record opens a new group iff start >= hull
//                    (connected components of interval overlap), winner per group =
//                    min (dist, -(end-start)), ties -> first in (start, end) order.
// Larger lists fall back to the host implementation (consolidate_recs in api.cu), which computes
// exactly the same thing.
#pragma once
#include "common.cuh"

namespace fzb {

constexpr int kPostMax = 16384;
constexpr int kRankThreads = 256;
constexpr int kConsThreads = 1024;
enum { CNT_POST_DONE = 5, CNT_NFINAL = 6 };
constexpr size_t kConsSmem = (size_t)kPostMax * 4;  // segbest
constexpr int kFinCols = 5;  // start, end, dist, hull_start, hull_end of the group

// canonical order (start, end, dist): start < 2^46, end-start < 2^10, dist < 2^8
__device__ __forceinline__ uint64_t canonical_key(const RawRec &r) {
    return ((uint64_t)r.start << 18) | ((uint64_t)(r.end - r.start) << 8) | (uint64_t)r.dist;
}
// generation order of the n-gram search: (n-gram ordinal, hit index)
__device__ __forceinline__ uint64_t generation_key(const RawRec &r) {
    return ((uint64_t)(uint32_t)r.ngram << 48) | (uint64_t)r.idx;
}

// All-pairs ranking.  Grid (i-tiles, j-chunks): block (bi, bj) ranks records i in [256 bi, 256 bi+256)
// against records j in [kRankChunk bj, kRankChunk (bj+1)) and adds its partial counts to ranks[] (zeroed
// by the host's counter memset).  ranks[i] = generation-order rank, ranks[kPostMax + i] = canonical rank.
// raw_mode: 0 -> raw order (ngram, idx); 1 -> raw order (start, end, dist); 2 -> the raw stream is left in
// arrival order (the host orders it lazily, only if somebody asks for it) and only canonical ranks
// are computed.
constexpr int kRankChunk = 1024;

__global__ void __launch_bounds__(kRankThreads)
k_rank(const RawRec *recs, uint32_t cap, int raw_mode, uint32_t *ranks, const uint32_t *counters) {
    const bool two_keys = raw_mode == 0;
    __shared__ uint64_t s1[kRankThreads], s2[kRankThreads];
    const uint32_t n = counters[CNT_OUT];
    if (n > (uint32_t)kPostMax || n > cap) return;
    if (blockIdx.x * kRankThreads >= n || blockIdx.y * kRankChunk >= n) return;
    const uint32_t i = blockIdx.x * kRankThreads + threadIdx.x;
    uint64_t k1 = ~0ull, k2 = ~0ull;
    if (i < n) {
        const RawRec me = recs[i];
        k2 = canonical_key(me);
        k1 = two_keys ? generation_key(me) : k2;
    }
    uint32_t r1 = 0, r2 = 0;
    const uint32_t t0 = blockIdx.y * (kRankChunk / kRankThreads);
    for (uint32_t t = t0; t < t0 + kRankChunk / kRankThreads && t * kRankThreads < n; t++) {
        const uint32_t j = t * kRankThreads + threadIdx.x;
        uint64_t a = ~0ull, b = ~0ull;
        if (j < n) {
            const RawRec o = recs[j];
            b = canonical_key(o);
            a = two_keys ? generation_key(o) : b;
        }
        __syncthreads();
        s1[threadIdx.x] = a;
        s2[threadIdx.x] = b;
        __syncthreads();
        if (t < blockIdx.x) {  // every j of the tile is < i: ties rank before me
#pragma unroll 8
            for (int jj = 0; jj < kRankThreads; jj++) r2 += (s2[jj] <= k2);
            if (two_keys) {
#pragma unroll 8
                for (int jj = 0; jj < kRankThreads; jj++) r1 += (s1[jj] <= k1);
            }
        } else if (t > blockIdx.x) {
#pragma unroll 8
            for (int jj = 0; jj < kRankThreads; jj++) r2 += (s2[jj] < k2);
            if (two_keys) {
#pragma unroll 8
                for (int jj = 0; jj < kRankThreads; jj++) r1 += (s1[jj] < k1);
            }
        } else {
#pragma unroll 8
            for (int jj = 0; jj < kRankThreads; jj++) {
                const bool before = jj < (int)threadIdx.x;
                r1 += (s1[jj] < k1) || (before && s1[jj] == k1);
                r2 += (s2[jj] < k2) || (before && s2[jj] == k2);
            }
        }
    }
    if (i < n) {
        atomicAdd(&ranks[i], two_keys ? r1 : r2);
        atomicAdd(&ranks[kPostMax + i], r2);
    }
}

// block-wide exclusive scans over one value per thread (1024 threads)
__device__ __forceinline__ unsigned long long block_excl_scan_max(unsigned long long v, unsigned long long *warp_tot) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc = max(inc, t);
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    unsigned long long pre = 0;
    for (int i = 0; i < w; i++) pre = max(pre, warp_tot[i]);
    unsigned long long exc = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if (lane == 0) exc = 0;
    __syncthreads();
    return max(pre, exc);
}

__device__ __forceinline__ uint32_t block_excl_scan_sum(uint32_t v, uint32_t *warp_tot, uint32_t *total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    uint32_t pre = 0, tot = 0;
    for (int i = 0; i < 32; i++) {
        if (i < w) pre += warp_tot[i];
        tot += warp_tot[i];
    }
    *total = tot;
    uint32_t exc = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if (lane == 0) exc = 0;
    __syncthreads();
    return pre + exc;
}

// One CTA: scatter the records into rank order, then the consolidation sweep.
__global__ void __launch_bounds__(kConsThreads)
k_consolidate(const RawRec *recs, const uint32_t *ranks, RawRec *raw_sorted, uint64_t *keys, uint32_t cap,
              int raw_mode, int do_consolidate, int64_t *fin, uint32_t *counters) {
    extern __shared__ __align__(16) uint8_t post_smem[];
    uint32_t *segbest = reinterpret_cast<uint32_t *>(post_smem);
    __shared__ unsigned long long warp_max[32];
    __shared__ uint32_t warp_sum[32];
    const uint32_t n = counters[CNT_OUT];
    if (n > (uint32_t)kPostMax || n > cap) {
        if (threadIdx.x == 0) counters[CNT_POST_DONE] = 0;
        return;
    }
    for (uint32_t i = threadIdx.x; i < n; i += kConsThreads) {
        const RawRec me = recs[i];
        if (raw_mode != 2) raw_sorted[ranks[i]] = me;
        keys[ranks[kPostMax + i]] = canonical_key(me);
    }
    __syncthreads();  // keys[] was written by this CTA: visible after the barrier
    uint32_t nfinal = 0;
    if (do_consolidate && n > 0) {
        const int chunk = (int)((n + kConsThreads - 1) / kConsThreads);  // <= 16
        const int lo = threadIdx.x * chunk;  // thread t owns sorted elements [t*chunk, (t+1)*chunk)
        uint64_t mykeys[kPostMax / kConsThreads];
        unsigned long long local_max = 0;
#pragma unroll
        for (int c = 0; c < kPostMax / kConsThreads; c++) {
            const int i = lo + c;
            mykeys[c] = (c < chunk && i < (int)n) ? keys[i] : ~0ull;
            if (c < chunk && i < (int)n)
                local_max = max(local_max, (unsigned long long)((mykeys[c] >> 18) + ((mykeys[c] >> 8) & 1023u)));
        }
        const unsigned long long hull0 = block_excl_scan_max(local_max, warp_max);
        unsigned long long hull = hull0;
        uint32_t nflags = 0;
#pragma unroll
        for (int c = 0; c < kPostMax / kConsThreads; c++) {  // group heads in my chunk
            const int i = lo + c;
            if (c < chunk && i < (int)n) {
                const unsigned long long s = mykeys[c] >> 18, e = s + ((mykeys[c] >> 8) & 1023u);
                if (i == 0 || s >= hull) nflags++;
                hull = max(hull, e);
            }
        }
        uint32_t total = 0;
        uint32_t seg = block_excl_scan_sum(nflags, warp_sum, &total);  // groups opened before my chunk
        nfinal = total;
        for (int t = threadIdx.x; t < (int)nfinal; t += kConsThreads) segbest[t] = 0xFFFFFFFFu;
        __syncthreads();
        hull = hull0;
#pragma unroll
        for (int c = 0; c < kPostMax / kConsThreads; c++) {
            const int i = lo + c;
            if (c < chunk && i < (int)n) {
                const uint64_t kx = mykeys[c];
                const unsigned long long s = kx >> 18, len = (kx >> 8) & 1023u, e = s + len;
                if (i == 0 || s >= hull) {  // group head: hull_start of this group, hull_end of the previous
                    seg++;
                    fin[kFinCols * (seg - 1) + 3] = (int64_t)s;
                    if (seg >= 2) fin[kFinCols * (seg - 2) + 4] = (int64_t)hull;
                }
                hull = max(hull, e);
                if (i == (int)n - 1) fin[kFinCols * (seg - 1) + 4] = (int64_t)hull;  // last group
                const uint32_t score = ((uint32_t)(kx & 255u) << 24) | ((uint32_t)(1023u - len) << 14) | (uint32_t)i;
                atomicMin(&segbest[seg - 1], score);
            }
        }
        __syncthreads();
        for (int g = threadIdx.x; g < (int)nfinal; g += kConsThreads) {
            const uint64_t kx = keys[segbest[g] & 16383u];
            const int64_t s = (int64_t)(kx >> 18);
            fin[kFinCols * g + 0] = s;
            fin[kFinCols * g + 1] = s + (int64_t)((kx >> 8) & 1023u);
            fin[kFinCols * g + 2] = (int64_t)(kx & 255u);
        }
    }
    if (threadIdx.x == 0) {
        counters[CNT_NFINAL] = nfinal;
        counters[CNT_POST_DONE] = 1;
    }
}

// Pack this shard's groups for the all-gather: slot = header row {count, valid, 0, 0, 0} + up to `cap`
// group rows.  valid = the device-side consolidation ran and the groups fit the slot.
__global__ void __launch_bounds__(256)
k_pack_groups(const int64_t *fin, const uint32_t *counters, int64_t *slot, uint32_t cap) {
    const uint32_t nf = counters[CNT_NFINAL];
    const bool valid = counters[CNT_POST_DONE] != 0 && nf <= cap && counters[CNT_OVERFLOW] == 0;
    if (blockIdx.x == 0 && threadIdx.x < kFinCols)
        slot[threadIdx.x] = threadIdx.x == 0 ? (int64_t)nf : (threadIdx.x == 1 ? (int64_t)valid : 0);
    if (!valid) return;
    const uint32_t total = nf * kFinCols;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
        slot[kFinCols + i] = fin[i];
}

}  // namespace fzb
