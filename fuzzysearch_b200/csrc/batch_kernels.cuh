// batch_kernels.cuh -- many patterns over one haystack in ONE pass (BASELINE.json configs[4]).
//
// The reference has no batch API: a caller loops find_near_matches over the patterns
// (src/fuzzysearch/__init__.py:35-57), reading the haystack (k+1) times per pattern.  Here every pattern for
// which the q-sample lemma of k_filter_sampled holds (floor((m-k-3)/4) >= k+1, m <= 64) shares one scan:
//   k_filter_multi  streams the haystack once.  The 4-grams of ALL those patterns (~30 K for 800 patterns) set bits
//        in a 2^20-bit table in shared memory (128 KiB, one CTA of 1024 threads per SM); each 4-byte-aligned
//        haystack word costs one hash + one bit test.  Table hits (~3 %: a Bloom filter at 3 % load) are confirmed by
//        the flagged lane itself against an L2-resident open-addressing table gram -> postings (pattern, offset);
//        a real hit of pattern p at offset o marks the granules that can hold an n-gram anchor of an occurrence
//        containing that word, anchors [g-o-k, g-o+k+m-L] (the word is aligned with pattern offset o up to k
//        insertions/deletions).  (pattern, granule) pairs are de-duplicated by a device hash set (CAS insert)
//        and the inserting thread appends the pair to the work list.
//   k_verify_multi  one warp per (pattern, granule): loads that pattern, builds its bit-parallel match table
//        in shared memory and runs the same verify_granule_lev as the single-pattern search; records are tagged
//        with the pattern number.  The set slot is cleared on the way out (the set is empty between batches).
// The host splits the records by pattern and consolidates each list (api.cu).
#pragma once
#include "kernels.cuh"

namespace fzb {

constexpr int kMultiThreads = 1024;
constexpr int kMultiTblBits = 20;
constexpr int kMultiTblWords = 1 << (kMultiTblBits - 5);  // 32768 words = 128 KiB
constexpr size_t kMultiSmem = (size_t)kMultiTblWords * 4;
constexpr int kMultiUnroll = 4;                            // uint4 loads in flight per thread
constexpr int kMultiTileVecs = kMultiThreads * kMultiUnroll;
constexpr uint32_t kGramMul = 0x85EBCA77u;
constexpr int kBatchMaxM = 64;

struct BatchPat {  // device copy of one pattern
    uint8_t P[kBatchMaxM];
    int32_t m, k, L, n_ngrams;
};

struct WorkItem {
    uint32_t pid, granule, slot, pad;
};

struct MultiParams {
    const uint8_t *H;
    int64_t buf_lo, buf_len, N, own_lo, own_hi;
    const uint32_t *bits;      // global copy of the bit table (kMultiTblWords words)
    const uint2 *gtab;         // open addressing: .x gram, .y = first posting | count << 24 (0 = empty slot)
    uint32_t gtab_mask;
    const uint32_t *postings;  // pid << 8 | offset
    const uint32_t *pinfo;     // per pattern: m | k << 8 | L << 16
    unsigned long long *set;   // de-duplication set of (pid << 32 | granule) + 1, 0 = empty
    uint32_t set_mask;
    WorkItem *work;
    uint32_t work_cap;
    uint32_t *counters;        // CNT_GRAN = work items appended, CNT_OVERFLOW
};

__device__ __forceinline__ uint32_t multi_hash(uint32_t w) { return (w * kHashMul) >> (32 - kMultiTblBits); }

// one (pattern, granule) pair: insert into the set; the inserter appends the work item
__device__ __forceinline__ void multi_mark(const MultiParams &p, uint32_t pid, uint32_t granule) {
    const unsigned long long key = (((unsigned long long)pid << 32) | granule) + 1ull;
    uint32_t slot = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & p.set_mask;
    for (uint32_t probe = 0; probe < 256; probe++) {
        const unsigned long long old = atomicCAS(&p.set[slot], 0ull, key);
        if (old == key) return;  // somebody marked it already
        if (old == 0ull) {
            const uint32_t w = atomicAdd(&p.counters[CNT_GRAN], 1u);
            if (w < p.work_cap) {
                WorkItem it;
                it.pid = pid;
                it.granule = granule;
                it.slot = slot;
                it.pad = 0;
                p.work[w] = it;
            } else {
                atomicExch(&p.counters[CNT_OVERFLOW], 1u);
            }
            return;
        }
        slot = (slot + 1) & p.set_mask;
    }
    atomicExch(&p.counters[CNT_OVERFLOW], 1u);  // set too full: the host redoes the batch pattern by pattern
}

// a haystack word whose table bit is set: look the gram up, mark the granules of every pattern that has it
__device__ __noinline__ void multi_confirm(const MultiParams &p, uint32_t w, int64_t word_off) {
    uint32_t slot = (w * kGramMul) & p.gtab_mask;
    for (;;) {
        const uint2 e = __ldg(p.gtab + slot);
        if (e.y == 0u) return;  // not a pattern gram: a false positive of the bit table
        if (e.x == w) {
            const uint32_t first = e.y & 0xFFFFFFu, cnt = e.y >> 24;
            const int64_t g = p.buf_lo + word_off;
            for (uint32_t i = 0; i < cnt; i++) {
                const uint32_t post = __ldg(p.postings + first + i);
                const uint32_t pid = post >> 8;
                const int o = (int)(post & 0xFFu);
                const uint32_t info = __ldg(p.pinfo + pid);
                const int m = (int)(info & 0xFFu), k = (int)((info >> 8) & 0xFFu), L = (int)((info >> 16) & 0xFFu);
                int64_t lo = g - o - k, hi = g - o + k + (m - L);  // anchors of occurrences aligning this word with offset o
                if (lo < p.own_lo) lo = p.own_lo;
                if (hi > p.own_hi - 1) hi = p.own_hi - 1;
                if (lo > hi) continue;
                const int64_t g0 = (lo - p.buf_lo) >> kGranuleShift, g1 = (hi - p.buf_lo) >> kGranuleShift;
                for (int64_t gr = g0; gr <= g1; gr++) multi_mark(p, pid, (uint32_t)gr);
            }
            // keep probing: a gram with more than 255 postings occupies several slots
        }
        slot = (slot + 1) & p.gtab_mask;
    }
}

__global__ void __launch_bounds__(kMultiThreads, 1)
k_filter_multi(const __grid_constant__ MultiParams p, int64_t nvec, int64_t ntiles) {
    extern __shared__ __align__(16) uint32_t mtbl[];
    for (int i = threadIdx.x; i < kMultiTblWords / 4; i += kMultiThreads)
        reinterpret_cast<uint4 *>(mtbl)[i] = __ldg(reinterpret_cast<const uint4 *>(p.bits) + i);
    __syncthreads();
    const uint4 *base = reinterpret_cast<const uint4 *>(p.H);
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t v0 = t * kMultiTileVecs + threadIdx.x;
        uint4 d[kMultiUnroll];
#pragma unroll
        for (int u = 0; u < kMultiUnroll; u++) {
            const int64_t v = v0 + (int64_t)u * kMultiThreads;
            d[u] = (v < nvec) ? ldg_stream(base + v) : make_uint4(0, 0, 0, 0);
        }
        uint32_t acc = 0;  // bit (15 - 4u - i) <-> word i of load u
#pragma unroll
        for (int u = 0; u < kMultiUnroll; u++) {
            const uint32_t ws[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t h = multi_hash(ws[i]);
                acc = acc * 2u + ((mtbl[h >> 5] >> (h & 31u)) & 1u);
            }
        }
        if (acc) {  // ~3 % of the words: the flagged lane confirms its own words (independent L2 probes)
#pragma unroll
            for (int u = 0; u < kMultiUnroll; u++) {
                const uint32_t nib = (acc >> (4 * (kMultiUnroll - 1 - u))) & 0xFu;
                if (!nib) continue;
                const int64_t off = (v0 + (int64_t)u * kMultiThreads) * 16;
                if (off >= p.buf_len) continue;  // zero padding behind the buffer
                if (nib & 8u) multi_confirm(p, d[u].x, off);
                if (nib & 4u) multi_confirm(p, d[u].y, off + 4);
                if (nib & 2u) multi_confirm(p, d[u].z, off + 8);
                if (nib & 1u) multi_confirm(p, d[u].w, off + 12);
            }
        }
    }
}

constexpr int kVmThreads = 128;

__global__ void __launch_bounds__(kVmThreads)
k_verify_multi(const __grid_constant__ MultiParams p, const BatchPat *pats, RawRec *out, uint32_t cap, uint32_t *counters) {
    __shared__ __align__(8) uint8_t sPall[kVmThreads / 32][kBatchMaxM];
    __shared__ unsigned long long sPMall[kVmThreads / 32][256];
    __shared__ uint32_t sWinAll[kVmThreads / 32][kWinWords];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *sP = sPall[warp];
    unsigned long long *sPM = sPMall[warp];
    uint32_t *sWin = sWinAll[warp];
    const uint32_t nitems = min(counters[CNT_GRAN], p.work_cap);
    uint32_t cur_pid = 0xFFFFFFFFu;
    VerifyCtx c;
    c.H = p.H;
    c.buf_lo = p.buf_lo;
    c.buf_len = p.buf_len;
    c.N = p.N;
    c.own_lo = p.own_lo;
    c.own_hi = p.own_hi;
    c.m = c.k = c.L = c.n_ngrams = 0;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(&counters[CNT_WORK], 1u);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= nitems) break;
        const WorkItem it = p.work[item];
        if (it.pid != cur_pid) {  // load the pattern and build its match table (PM[c] = {i : P[i] == c})
            const BatchPat *bp = pats + it.pid;
            __syncwarp();
            if (lane < kBatchMaxM / 4) reinterpret_cast<uint32_t *>(sP)[lane] = reinterpret_cast<const uint32_t *>(bp->P)[lane];
            for (int ch = lane; ch < 256; ch += 32) sPM[ch] = 0ull;
            c.m = bp->m;
            c.k = bp->k;
            c.L = bp->L;
            c.n_ngrams = bp->n_ngrams;
            __syncwarp();
            for (int i = lane; i < c.m; i += 32) atomicOr(&sPM[sP[i]], 1ull << i);
            __syncwarp();
            cur_pid = it.pid;
        }
        verify_granule_lev<1>(c, sP, sPM, sWin, (int64_t)it.granule, lane, nullptr, out, cap, counters, (int)(it.pid << 8));
        if (lane == 0) p.set[it.slot] = 0ull;  // the set is empty again when the kernel ends
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counters[CNT_CAND], nitems);
}

}  // namespace fzb
