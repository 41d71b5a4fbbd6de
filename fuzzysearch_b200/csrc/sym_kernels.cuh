// sym_kernels.cuh -- sequences of WIDE symbols (general Unicode str: UTF-32 / UCS-2 code units; any sequence
// whose items were numbered by the caller) reduced to the byte alphabet the search kernels scan.
//
// Every algorithm on the path compares a PATTERN symbol with a SEQUENCE symbol and nothing else
// (levenshtein_ngram.py:49,113 `subseq[j] != char`, search_exact.py:53-56, levenshtein.py:83,
// generic_search.py:86, substitutions_only.py:93-99): two sequence symbols are never compared with each
// other.  So the result is unchanged when every pattern symbol is renamed to its rank 1..d among the
// pattern's distinct symbols (d <= m <= 255) and every sequence symbol that does not occur in the pattern
// is renamed to 0.  The sequence then has one BYTE per symbol and all the byte kernels apply -- a 4-byte
// symbol kernel family would scan four times the bytes for the same answer.
//
// k_reduce_symbols: one pass over a chunk of code units (4n or 2n bytes read, n bytes written; HBM-bound,
// but the chunk arrives over PCIe, so it is never the limiter).  The sorted alphabet (<= 255 entries) sits
// in shared memory; a lane reduces 4 symbols (one 16-/8-byte load) with a branch-free 8-step binary search
// each and stores one 32-bit word.
#pragma once
#include "common.cuh"

namespace fzb {

constexpr int kSymThreads = 256;

template <typename SYM>
__device__ __forceinline__ uint32_t reduce_symbol(const uint32_t *sA, uint32_t c) {
    // sA[0..255]: the alphabet ascending, padded with 0xFFFFFFFF ... and sA[256] = n_alpha
    uint32_t lo = 0;  // first index whose entry is >= c  (entries are strictly ascending below n_alpha)
#pragma unroll
    for (int step = 128; step >= 1; step >>= 1)
        if (sA[lo + step - 1] < c) lo += step;
    return (lo < sA[256] && sA[lo] == c) ? lo + 1 : 0u;
}

template <typename SYM>
__global__ void __launch_bounds__(kSymThreads)
k_reduce_symbols(const SYM *__restrict__ src, uint64_t n, const uint32_t *__restrict__ alphabet, uint32_t n_alpha,
                 uint8_t *__restrict__ dst) {
    __shared__ uint32_t sA[257];
    for (int i = threadIdx.x; i < 256; i += kSymThreads) sA[i] = i < (int)n_alpha ? alphabet[i] : 0xFFFFFFFFu;
    if (threadIdx.x == 0) sA[256] = n_alpha;
    __syncthreads();
    const uint64_t nquads = n / 4;
    const uint64_t stride = (uint64_t)gridDim.x * kSymThreads;
    for (uint64_t q = (uint64_t)blockIdx.x * kSymThreads + threadIdx.x; q < nquads; q += stride) {
        uint32_t c0, c1, c2, c3;
        if (sizeof(SYM) == 4) {
            const uint4 v = reinterpret_cast<const uint4 *>(src)[q];
            c0 = v.x, c1 = v.y, c2 = v.z, c3 = v.w;
        } else {
            const uint2 v = reinterpret_cast<const uint2 *>(src)[q];
            c0 = v.x & 0xFFFFu, c1 = v.x >> 16, c2 = v.y & 0xFFFFu, c3 = v.y >> 16;
        }
        const uint32_t w = reduce_symbol<SYM>(sA, c0) | (reduce_symbol<SYM>(sA, c1) << 8) |
                           (reduce_symbol<SYM>(sA, c2) << 16) | (reduce_symbol<SYM>(sA, c3) << 24);
        reinterpret_cast<uint32_t *>(dst)[q] = w;
    }
    // tail (n % 4 symbols): one thread
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (uint64_t i = nquads * 4; i < n; i++) dst[i] = (uint8_t)reduce_symbol<SYM>(sA, (uint32_t)src[i]);
}

}  // namespace fzb
