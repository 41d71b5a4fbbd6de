// p2p_kernels.cuh -- the multi-GPU reduction of one search over NVLink peer memory (FZB_F_GLOBAL).
//
// shards are searched independently; the only exchange is the match list.  Every rank's
// k_post leaves its consolidated groups (winner + hull, sorted, disjoint) in device memory.  Then, on the
// same stream, with no host involvement and no NCCL call in the step:
//   k_push   writes this rank's rows straight into a slot of EVERY rank's receive area (peer pointers
//            obtained once through CUDA IPC / peer access: plain st.global over NVLink through the NVSwitch),
//            fences at system scope and raises one flag per peer (the search's epoch number);
//   k_merge  (one CTA per source rank) spins on its own flags until all slots of this epoch have landed
//            (bounded: a timeout reports an error instead of hanging the GPU), then builds the GLOBAL
//            consolidate_overlapping_matches (common.py:185-189) of all shards without sorting anything:
//            runs are sorted and internally disjoint, so a row is the head of a global group iff no row of
//            ANOTHER run that precedes it reaches its hull start -- one comparison against each run's boundary
//            values for all rows but the few next to a seam, which binary-search the neighbouring run.  The
//            global position of a row is its index plus the sizes of the preceding runs (seam rows: plus / minus
//            the interleaved ones), its group number is that position minus the (tiny) number of non-head rows
//            before it, and the winner of a group is an atomicMin over a packed (dist, -length, start) score.
//            The global final list goes to mapped pinned host memory in 16-byte rows.
// Double buffering by epoch parity makes the slots safe to reuse: a rank can only be one search ahead of
// a peer (it needs the peer's flag of search i to finish search i).
// If any rank could not contribute (its list overflowed a buffer), its header says so, every rank sees every
// header and all of them take the staged host path (NCCL all-gather of the rows, api.cu).
#pragma once
#include "post_kernels.cuh"

namespace fzb {

constexpr int kMaxWorld = 16;
constexpr int kPushCtas = 16;
constexpr int kPushThreads = 256;
constexpr int kHdrWords = 8;                 // int64 words of a slot header: count, valid, epoch, mode
constexpr uint32_t kMaxNonHeads = 2048;
enum { MS_OK = 1, MS_FALLBACK = 2, MS_TIMEOUT = 3, MS_OVERFLOW = 4 };

struct WorldArgs {
    int world, rank;
    uint32_t epoch;       // this search's number (>= 1); parity = epoch & 1
    uint32_t cap;         // rows per slot
    uint64_t slot_bytes;  // (kHdrWords + cap * kFinCols) * 8
    uint64_t flags_off;   // byte offset of flags[2][kMaxWorld] (uint32) inside a receive area
    uint8_t *peer[kMaxWorld];  // receive area of every rank (peer[rank] = mine)
};

__device__ __forceinline__ uint8_t *slot_ptr(const WorldArgs &w, int owner, int src) {
    return w.peer[owner] + ((uint64_t)(w.epoch & 1u) * (uint64_t)w.world + (uint64_t)src) * w.slot_bytes;
}

struct MergeScratch {  // device memory, per handle
    uint32_t ticket;      // k_push: CTAs done
    uint32_t barrier;     // k_merge: grid barrier arrivals
    uint32_t nh_count;    // non-head rows found
    uint32_t status;
    uint32_t nh_pos[kMaxNonHeads];
};

// ---- push ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPushThreads)
k_push(const WorldArgs w, const int64_t *fin, const uint32_t *counters, int mode, MergeScratch *ms) {
    __shared__ uint32_t s_ticket;
    const uint32_t nf = counters[CNT_NFINAL];
    const bool valid = counters[CNT_POST_DONE] != 0 && nf <= w.cap && counters[CNT_OVERFLOW] == 0 &&
                       counters[CNT_OUT] <= (uint32_t)kPostMax;
    const uint32_t nvec = valid ? (uint32_t)(((uint64_t)nf * kFinCols * 8 + 15) / 16) : 0;
    const uint4 *src = reinterpret_cast<const uint4 *>(fin);
    const uint32_t gtid = blockIdx.x * kPushThreads + threadIdx.x, gsize = gridDim.x * kPushThreads;
    for (int r = 0; r < w.world; r++) {
        uint8_t *slot = slot_ptr(w, r, w.rank);
        if (gtid < kHdrWords) {
            int64_t v = 0;
            if (gtid == 0) v = (int64_t)nf;
            if (gtid == 1) v = valid ? 1 : 0;
            if (gtid == 2) v = (int64_t)w.epoch;
            if (gtid == 3) v = (int64_t)mode;
            reinterpret_cast<int64_t *>(slot)[gtid] = v;
        }
        uint4 *dst = reinterpret_cast<uint4 *>(slot + kHdrWords * 8);
        for (uint32_t i = gtid; i < nvec; i += gsize) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(&ms->ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    // every CTA's stores are ordered before this point: raise the flags
    __threadfence_system();
    if (threadIdx.x < (uint32_t)w.world) {
        volatile uint32_t *flag = reinterpret_cast<volatile uint32_t *>(w.peer[threadIdx.x] + w.flags_off) +
                                  (w.epoch & 1u) * kMaxWorld + w.rank;
        *flag = w.epoch;
    }
    if (threadIdx.x == 0) {  // reset for k_merge (runs behind this kernel on the same stream)
        ms->ticket = 0;
        ms->barrier = 0;
        ms->nh_count = 0;
        ms->status = 0;
    }
    // every CTA of this kernel has read the search's counters (before its ticket): leave them zeroed for the next search
    if (valid && threadIdx.x < CNT_COUNT) const_cast<uint32_t *>(counters)[threadIdx.x] = 0;
}

// ---- merge -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
#ifdef FZB_EMU
    return *reinterpret_cast<const volatile uint32_t *>(p);
#else
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#endif
}

// all CTAs of the grid arrive; returns false on timeout
__device__ __forceinline__ bool grid_barrier(MergeScratch *ms, uint32_t target) {
    __shared__ uint32_t s_ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&ms->barrier, 1u);
        uint32_t ok = 1;
        const long long t0 = clock64();
        while (ld_volatile_u32(&ms->barrier) < target) {
            if (clock64() - t0 > 4000000000ll) {  // ~2 s
                ok = 0;
                break;
            }
        }
        __threadfence();
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}

struct MergeOut {
    unsigned long long *score;  // device scratch [world * cap]
    uint32_t *pos;              // device scratch [world * cap]
    int64_t *h_rows;            // mapped host: global final rows, 2 x int64 each: start, (end-start) << 32 | dist
    uint32_t *h_hdr;            // mapped host: [0] status, [1] number of global final rows, [2] epoch, [7] seq (LAST)
    uint32_t seq;
};

// CTA 0 / thread 0 only: publish the outcome; the sequence word goes last (the host polls it)
__device__ __forceinline__ void merge_finish(const MergeOut &o, uint32_t status, uint32_t ng, uint32_t epoch) {
    o.h_hdr[1] = ng;
    o.h_hdr[2] = epoch;
    o.h_hdr[0] = status;
    __threadfence_system();
    o.h_hdr[7] = o.seq;
}

__global__ void __launch_bounds__(kPostThreads, 1)
k_merge(const WorldArgs w, MergeScratch *ms, const MergeOut o) {
    __shared__ int64_t s_first[kMaxWorld], s_last[kMaxWorld], s_helast[kMaxWorld];
    __shared__ uint32_t s_cnt[kMaxWorld], s_off[kMaxWorld + 1];
    __shared__ uint32_t s_state;
    __shared__ uint32_t s_nh[kMaxNonHeads];
    const uint32_t tid = threadIdx.x;
    const int q = blockIdx.x;  // my run
    const int W = w.world;
    const uint32_t t_start = gtimer_lo();
    // ---- wait until every rank's slot of this epoch has landed in MY receive area ----------------------
    if (tid == 0) {
        const uint32_t *flags = reinterpret_cast<const uint32_t *>(w.peer[w.rank] + w.flags_off) + (w.epoch & 1u) * kMaxWorld;
        uint32_t state = MS_OK;
        const long long t0 = clock64();
        for (int r = 0; r < W && state == MS_OK; r++) {
            while (ld_volatile_u32(flags + r) != w.epoch) {
                if (clock64() - t0 > 8000000000ll) {  // ~4 s: a peer never arrived
                    state = MS_TIMEOUT;
                    break;
                }
            }
        }
        __threadfence_system();
        s_state = state;
    }
    __syncthreads();
    if (s_state == MS_OK && tid < (uint32_t)W) {
        const int64_t *hdr = reinterpret_cast<const int64_t *>(slot_ptr(w, w.rank, tid));
        const int64_t cnt = __ldcg(hdr + 0), valid = __ldcg(hdr + 1), ep = __ldcg(hdr + 2);
        if (!valid || ep != (int64_t)w.epoch) atomicExch(&s_state, (uint32_t)MS_FALLBACK);
        s_cnt[tid] = valid ? (uint32_t)cnt : 0;
    }
    __syncthreads();
    if (s_state != MS_OK) {
        if (q == 0 && tid == 0) merge_finish(o, s_state, 0, w.epoch);
        return;
    }
    const uint32_t t_landed = gtimer_lo();
    const int mode = (int)__ldcg(reinterpret_cast<const int64_t *>(slot_ptr(w, w.rank, 0)) + 3);
    if (tid < (uint32_t)W) {
        const int64_t *rows = reinterpret_cast<const int64_t *>(slot_ptr(w, w.rank, tid)) + kHdrWords;
        const uint32_t c = s_cnt[tid];
        s_first[tid] = c ? __ldcg(rows + 3) : 0;
        s_last[tid] = c ? __ldcg(rows + (size_t)kFinCols * (c - 1) + 3) : 0;
        s_helast[tid] = c ? __ldcg(rows + (size_t)kFinCols * (c - 1) + 4) : 0;
    }
    if (tid == 0) {
        uint32_t acc = 0;
        for (int r = 0; r < W; r++) {
            s_off[r] = acc;
            acc += s_cnt[r];
        }
        s_off[W] = acc;
    }
    __syncthreads();
    const uint32_t T = s_off[W];
    const int64_t *myrows = reinterpret_cast<const int64_t *>(slot_ptr(w, w.rank, q)) + kHdrWords;
    const uint32_t mycnt = s_cnt[q];
    // ---- phase 1: global position and head flag of each of my rows ----------------------------------------
    for (uint32_t i = tid; i < mycnt; i += kPostThreads) {
        const int64_t hs = __ldcg(myrows + (size_t)kFinCols * i + 3);
        uint32_t pos = i;
        bool head = true;
        for (int r = 0; r < W; r++) {
            const uint32_t c = s_cnt[r];
            if (r == q || c == 0) continue;
            // rows of run r that precede me: hull start smaller, or equal and from a lower rank
            const bool low = r < q;
            uint32_t nb;
            if (s_last[r] < hs || (low && s_last[r] == hs)) {
                nb = c;
            } else if (!(s_first[r] < hs || (low && s_first[r] == hs))) {
                nb = 0;
            } else {  // next to a seam: binary search in run r
                const int64_t *rr = reinterpret_cast<const int64_t *>(slot_ptr(w, w.rank, r)) + kHdrWords;
                uint32_t lo = 0, hi = c;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    const int64_t v = __ldcg(rr + (size_t)kFinCols * mid + 3);
                    if (v < hs || (low && v == hs)) lo = mid + 1; else hi = mid;
                }
                nb = lo;
            }
            pos += nb;
            if (nb && mode == 1) {  // hull ends grow along a run (its groups are disjoint): the last one is the max
                const int64_t *rr = reinterpret_cast<const int64_t *>(slot_ptr(w, w.rank, r)) + kHdrWords;
                const int64_t he = nb == c ? s_helast[r] : __ldcg(rr + (size_t)kFinCols * (nb - 1) + 4);
                if (he > hs) head = false;
            }
        }
        o.pos[s_off[q] + i] = pos;
        o.score[pos] = ~0ull;
        if (!head) {
            const uint32_t slot = atomicAdd(&ms->nh_count, 1u);
            if (slot < kMaxNonHeads) ms->nh_pos[slot] = pos;
        }
    }
    if (!grid_barrier(ms, (uint32_t)W)) {
        if (q == 0 && tid == 0) merge_finish(o, MS_TIMEOUT, 0, w.epoch);
        return;
    }
    // ---- phase 2: group number = position - non-heads at or before it; winners by atomicMin ----------------
    const uint32_t nh = ld_volatile_u32(&ms->nh_count);
    if (nh > kMaxNonHeads) {  // pathological chaining across seams: the host does it
        if (q == 0 && tid == 0) merge_finish(o, MS_OVERFLOW, 0, w.epoch);
        return;
    }
    for (uint32_t i = tid; i < nh; i += kPostThreads) s_nh[i] = __ldcg(ms->nh_pos + i);
    __syncthreads();
    for (uint32_t i = tid; i < mycnt; i += kPostThreads) {
        const uint32_t pos = o.pos[s_off[q] + i];
        uint32_t below = 0;
        for (uint32_t j = 0; j < nh; j++) below += (s_nh[j] <= pos);
        const int64_t s = __ldcg(myrows + (size_t)kFinCols * i), e = __ldcg(myrows + (size_t)kFinCols * i + 1),
                      d = __ldcg(myrows + (size_t)kFinCols * i + 2);
        const unsigned long long score = ((unsigned long long)d << 56) | ((unsigned long long)(1023 - (e - s)) << 46) |
                                         (unsigned long long)s;
        atomicMin(&o.score[pos - below], score);
    }
    if (!grid_barrier(ms, 2u * (uint32_t)W)) {
        if (q == 0 && tid == 0) merge_finish(o, MS_TIMEOUT, 0, w.epoch);
        return;
    }
    // ---- phase 3: decode the winners into the host rows (coalesced 16-byte stores) --------------------------
    const uint32_t ng = T - nh;
    for (uint32_t g = (uint32_t)q * kPostThreads + tid; g < ng; g += (uint32_t)W * kPostThreads) {
        const unsigned long long sc = __ldcg(o.score + g);
        const long long s = (long long)(sc & ((1ull << 46) - 1));
        const unsigned long long len = 1023ull - ((sc >> 46) & 1023ull), d = sc >> 56;
        longlong2 row;
        row.x = s;
        row.y = (long long)((len << 32) | d);
        reinterpret_cast<longlong2 *>(o.h_rows)[g] = row;
    }
    __threadfence_system();  // my rows are on their way to the host before the barrier says so
    if (!grid_barrier(ms, 3u * (uint32_t)W)) {
        if (q == 0 && tid == 0) merge_finish(o, MS_TIMEOUT, 0, w.epoch);
        return;
    }
    if (q == 0 && tid == 0) {
        o.h_hdr[4] = t_landed - t_start;      // ns spent waiting for the peers' slots (fzb_debug_counters)
        o.h_hdr[5] = gtimer_lo() - t_landed;  // ns of the merge itself
        o.h_hdr[6] = nh;
        merge_finish(o, MS_OK, ng, w.epoch);
    }
}

}  // namespace fzb
