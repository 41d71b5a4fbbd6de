// tests/emu/include/cuda_runtime.h -- TEST INFRASTRUCTURE, never shipped, never loaded by the product.
//
// A small CUDA execution-model emulator for the CPU suite: tests/emu/build_emu.py compiles the PRODUCT sources
// (fuzzysearch_b200/csrc/api.cu + *.cuh, textually unchanged except for the launch syntax, see the script) with g++
// against this header instead of the CUDA toolkit's, into tests/emu/_build/libfuzzb200_emu.so.  The library exports
// the same C-ABI, so the `-m gpu` parity tests can be replayed here, without a GPU, against the kernels' real source:
// host logic (api.cu), kernel logic, work lists, overflow paths, the post-processing kernel, the in-process
// multi-shard worlds -- everything except timing, memory-model behaviour, CUDA IPC and NCCL.
//
// Model: one host thread inside the emulator at a time (global mutex); the CTAs of a grid run one after another;
// the threads of a CTA are fibers (own stacks, hand-written x86-64 context switch) scheduled round-robin; a fiber
// runs until it blocks in __syncthreads(), a warp collective (*_sync: rendezvous of the lanes named in the mask) or
// a spin-wait.  A CTA in which every runnable thread merely spins (a grid-wide barrier, a flag a peer GPU raises)
// is SET ASIDE -- its __shared__ variables (function-local statics, registered by build_emu.py) and its dynamic
// shared memory saved -- and the next CTA starts; a launch whose remaining CTAs all wait is left pending on its
// stream and the call returns to the host, like an asynchronous launch; any later runtime call (cudaStreamQuery from
// the product's completion poll, a launch by another host thread) advances it.  That is enough for the multi-GPU
// reduction kernels (k_push / k_merge) of an in-process world with one host thread per shard.
// Device memory is host memory filled with 0xCD on allocation (an uninitialised read shows); shared memory
// starts as 0xA5.
#pragma once
#if !defined(__x86_64__)
#error "the CUDA emulator's context switch is written for x86-64"
#endif
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/asan_interface.h>
#endif
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <functional>
#include <mutex>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

// ---- qualifiers -------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

// ---- vector types -----------------------------------------------------------------------------------------------
struct uint3 {
    unsigned int x, y, z;
};
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(8))) uint2 {
    unsigned int x, y;
};
struct __attribute__((aligned(16))) uint4 {
    unsigned int x, y, z, w;
};
struct __attribute__((aligned(16))) ulonglong2 {
    unsigned long long x, y;
};
struct __attribute__((aligned(16))) longlong2 {
    long long x, y;
};
static inline uint2 make_uint2(unsigned int x, unsigned int y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned int x, unsigned int y, unsigned int z, unsigned int w) { return uint4{x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
static inline longlong2 make_longlong2(long long x, long long y) { return longlong2{x, y}; }

// CUDA's min/max overload set follows the usual arithmetic conversions for mixed arguments
template <class A, class B, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
static inline typename std::common_type<A, B>::type min(A a, B b) {
    typedef typename std::common_type<A, B>::type T;
    return (T)a < (T)b ? (T)a : (T)b;
}
template <class A, class B, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
static inline typename std::common_type<A, B>::type max(A a, B b) {
    typedef typename std::common_type<A, B>::type T;
    return (T)a > (T)b ? (T)a : (T)b;
}

// ---- ThreadSanitizer (FZB_EMU_TSAN build): fibers are TSan fibers, the only happens-before edges are the ones CUDA
// gives -- launch -> threads, __syncthreads(), the *_sync warp primitives, fences, mbarrier phases, thread exit ->
// end of the launch -- so two
// conflicting accesses of different CUDA threads that nothing orders are REPORTED (a racecheck over shared AND
// global memory).  The emulator's own bookkeeping is not instrumented.
#if defined(__SANITIZE_THREAD__)
#include <sanitizer/tsan_interface.h>
extern "C" void AnnotateBenignRaceSized(const char *file, int line, const volatile void *mem, long size,
                                        const char *description);
#define EMU_NOTSAN __attribute__((no_sanitize("thread")))
#define EMU_TSAN(...) __VA_ARGS__
#define EMU_ATOMIC_ORDER __ATOMIC_RELAXED  // device atomics order nothing by themselves
#else
#define EMU_NOTSAN
#define EMU_TSAN(...)
#define EMU_ATOMIC_ORDER __ATOMIC_SEQ_CST
#endif

// ---- the fiber scheduler ----------------------------------------------------------------------------------------
extern "C" void fzb_emu_switch(void **save_sp, void *load_sp);
asm(R"(
        .text
        .type fzb_emu_switch,@function
fzb_emu_switch:
        pushq %rbp
        pushq %rbx
        pushq %r12
        pushq %r13
        pushq %r14
        pushq %r15
        movq %rsp, (%rdi)
        movq %rsi, %rsp
        popq %r15
        popq %r14
        popq %r13
        popq %r12
        popq %rbx
        popq %rbp
        ret
        .size fzb_emu_switch, .-fzb_emu_switch
)");

namespace emu {

constexpr unsigned kMaxThreads = 1024;
constexpr size_t kStackBytes = 128 * 1024;
constexpr size_t kDynSmemBytes = 232448;  // 227 KiB
constexpr int kRendezvous = 4;            // concurrent collectives per warp (disjoint masks)
constexpr size_t kMaxLiveCtas = 32;       // co-resident CTAs of one launch that wait for something

enum Op { OP_SYNCWARP, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_ANY, OP_ALL, OP_MATCH, OP_REDUCE };

struct Rendezvous {
    uint32_t mask = 0;       // 0 = free
    uint32_t arrived = 0;
    uint32_t departed = 0;
    uint32_t complete = 0;   // set once every live lane of the mask has arrived
    uint32_t gen = 0;        // bumps on completion (what the waiters watch)
    int op = 0;
    uint64_t val[32];
};

struct Warp {
    uint32_t alive = 0;
    Rendezvous rv[kRendezvous];
};

struct Fiber {
    void *sp = nullptr;
    bool done = true;
    bool spun = false;                             // left the CPU through yield(): a spin-wait, no progress
    const volatile uint32_t *wait_word = nullptr;  // runnable iff !wait_word || *wait_word != wait_val
    uint32_t wait_val = 0;
    uint3 tid{0, 0, 0};
    unsigned lin = 0;
    void *tsan = nullptr;  // TSan fiber context (FZB_EMU_TSAN build)
};

// One CTA: resumable.  A CTA whose runnable threads all spin (a grid barrier, a flag another GPU raises) is set
// aside -- its __shared__ variables and dynamic shared memory saved -- while other CTAs / launches / host threads run.
struct Cta {
    Fiber fib[kMaxThreads];
    Warp warp[kMaxThreads / 32];
    unsigned nthreads = 0, alive = 0;
    unsigned bar_arrived = 0;
    uint32_t bar_gen = 0;
    char bar_token = 0;
    int bar_or_acc = 0, bar_or_res[2] = {0, 0};  // __syncthreads_or
    uint3 bid{0, 0, 0};
    char *stacks = nullptr;
    size_t stack_bytes = 0;
    bool progressed = false;
    size_t dyn_bytes = 0;
    size_t saved_vars = 0;  // how many registry entries the snapshot holds
    bool has_snapshot = false;
    std::vector<uint8_t> snap_static, snap_dyn;
};

struct Launch {
    dim3 grid, block;
    size_t smem = 0;
    std::function<void()> body;
    unsigned long long next = 0, total = 0;
    std::vector<Cta *> live;
};

struct SharedVar {
    void *p;
    size_t n;
};

}  // namespace emu

struct emu_stream {
    emu::Launch *pending = nullptr;  // at most one suspended launch (every stream-ordered call drains the stream first)
};
typedef emu_stream *cudaStream_t;

namespace emu {

struct Global {
    std::mutex mu;  // ONE host thread inside the emulator at a time
    Launch *cur_launch = nullptr;
    Cta *cur_cta = nullptr;
    Fiber *cur = nullptr;
    void *sched_sp = nullptr;
    Cta *owner = nullptr;  // whose data the __shared__ statics / the dynamic buffer hold right now
    SharedVar registry[4096];  // (a plain array: the sanitizer builds must not see the emulator's own bookkeeping)
    size_t nregistry = 0;
    std::vector<emu_stream *> pending;
    std::vector<Cta *> free_ctas;
    std::vector<std::pair<char *, size_t>> free_stacks;
    emu_stream default_stream;
    void *sched_tsan = nullptr;
    char launch_token = 0, done_token = 0, fence_token = 0;  // addresses TSan hangs happens-before edges on
};
EMU_NOTSAN inline Global &g() {
    static Global *G = new Global();  // never destroyed: host threads may still be inside at exit
    return *G;
}
alignas(1024) static uint8_t g_dyn_smem[kDynSmemBytes];
inline uint8_t *smem_base() { return g_dyn_smem; }
template <class T>
inline T *dyn_smem() {
    return reinterpret_cast<T *>(g_dyn_smem);
}

inline bool tsan_grid_mode() {
    static const bool on = [] {
        const char *e = getenv("FZB_EMU_TSAN_GRID");
        const bool v = e && *e && *e != '0';
        EMU_TSAN(if (v) AnnotateBenignRaceSized(__FILE__, __LINE__, g_dyn_smem, (long)kDynSmemBytes, "emulated dynamic smem");)
        return v;
    }();
    return on;
}

// every `__shared__` declaration registers its storage the first time control passes it (build_emu.py adds the
// registration next to the declaration)
struct SharedReg {
    EMU_NOTSAN SharedReg(void *p, size_t n) {
        Global &G = g();
        if (G.nregistry == 4096) abort();  // too many __shared__ declarations
        G.registry[G.nregistry++] = SharedVar{p, n};
        EMU_TSAN(if (tsan_grid_mode()) AnnotateBenignRaceSized(__FILE__, __LINE__, p, (long)n, "emulated __shared__");)
    }
};

[[noreturn]] inline void die(const char *what) {
    Global &G = g();
    fprintf(stderr, "cuda-emu: %s (block %u,%u,%u thread %u)\n", what, G.cur_cta ? G.cur_cta->bid.x : 0u,
            G.cur_cta ? G.cur_cta->bid.y : 0u, G.cur_cta ? G.cur_cta->bid.z : 0u, G.cur ? G.cur->lin : 0u);
    abort();
}

EMU_NOTSAN inline void yield() {  // still runnable: spin-wait on memory some other CTA, launch or host thread writes
    Global &G = g();
    Fiber *f = G.cur;
    f->wait_word = nullptr;
    f->spun = true;
    EMU_TSAN(__tsan_release(&G.done_token);)  // (only the scheduler acquires it, at CTA boundaries)
    EMU_TSAN(__tsan_switch_to_fiber(G.sched_tsan, __tsan_switch_to_fiber_no_sync);)
    fzb_emu_switch(&f->sp, G.sched_sp);
    EMU_TSAN(__tsan_acquire(&g().launch_token);)
}
EMU_NOTSAN inline void block_on(const volatile uint32_t *word, uint32_t val) {
    Global &G = g();
    Fiber *f = G.cur;
    while (*word == val) {
        f->wait_word = word;
        f->wait_val = val;
        EMU_TSAN(__tsan_release(&G.done_token);)
        EMU_TSAN(__tsan_switch_to_fiber(G.sched_tsan, __tsan_switch_to_fiber_no_sync);)
        fzb_emu_switch(&f->sp, G.sched_sp);
        EMU_TSAN(__tsan_acquire(&g().launch_token);)
    }
    f->wait_word = nullptr;
}

EMU_NOTSAN inline void rv_check_complete(Warp &w, Rendezvous &r) {
    const uint32_t need = r.mask & w.alive;
    if (!r.complete && r.mask && (r.arrived & need) == need) {
        r.complete = 1;
        r.gen++;
    }
}

EMU_NOTSAN inline void fiber_exit() {
    Global &G = g();
    Cta &c = *G.cur_cta;
    Fiber *f = G.cur;
    f->done = true;
    c.alive--;
    Warp &w = c.warp[f->lin >> 5];
    w.alive &= ~(1u << (f->lin & 31));
    for (int i = 0; i < kRendezvous; i++) rv_check_complete(w, w.rv[i]);  // a lane others were waiting for has left
    if (c.bar_arrived && c.bar_arrived == c.alive) {                       // ... or the CTA barrier was waiting for it
        c.bar_or_res[c.bar_gen & 1u] = c.bar_or_acc;
        c.bar_or_acc = 0;
        c.bar_arrived = 0;
        c.bar_gen++;
    }
    void *dummy;
    EMU_TSAN(__tsan_release(&G.done_token);)  // ... and what this thread did is visible to the host after the launch
    EMU_TSAN(__tsan_switch_to_fiber(G.sched_tsan, __tsan_switch_to_fiber_no_sync);)
    fzb_emu_switch(&dummy, G.sched_sp);
    die("resumed a finished fiber");
}

extern "C" EMU_NOTSAN inline void fzb_emu_trampoline() {
    EMU_TSAN(__tsan_acquire(&g().launch_token);)  // what the host did before the launch is visible to every thread
    g().cur_launch->body();
    fiber_exit();
}

// (a plain loop: under the sanitizers memcpy is an intercepted call, and these copies are the emulator's own business)
EMU_NOTSAN inline void raw_copy(void *dst, const void *src, size_t n) {
    uint8_t *d = static_cast<uint8_t *>(dst);
    const uint8_t *s = static_cast<const uint8_t *>(src);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {  // (constant-size builtin: inlined moves, no call, any alignment)
        uint64_t w;
        __builtin_memcpy(&w, s + i, 8);
        __builtin_memcpy(d + i, &w, 8);
    }
    for (; i < n; i++) d[i] = s[i];
}
EMU_NOTSAN inline void save_owner() {
    Global &G = g();
    Cta *o = G.owner;
    if (!o) return;
    size_t total = 0;
    for (size_t i = 0; i < G.nregistry; i++) total += G.registry[i].n;
    o->snap_static.resize(total);
    size_t off = 0;
    for (size_t i = 0; i < G.nregistry; i++) {
        raw_copy(o->snap_static.data() + off, G.registry[i].p, G.registry[i].n);
        off += G.registry[i].n;
    }
    o->saved_vars = G.nregistry;
    o->snap_dyn.resize(o->dyn_bytes);
    raw_copy(o->snap_dyn.data(), g_dyn_smem, o->dyn_bytes);
    o->has_snapshot = true;
}
EMU_NOTSAN inline void make_owner(Cta *c) {  // before c runs: its shared memory must be the one in place
    Global &G = g();
    if (G.owner == c) return;
    save_owner();
    if (c->has_snapshot) {
        size_t off = 0;
        for (size_t i = 0; i < c->saved_vars; i++) {
            raw_copy(G.registry[i].p, c->snap_static.data() + off, G.registry[i].n);
            off += G.registry[i].n;
        }
        if (!c->snap_dyn.empty()) raw_copy(g_dyn_smem, c->snap_dyn.data(), c->snap_dyn.size());
    } else {
        memset(g_dyn_smem, 0xA5, c->dyn_bytes);  // shared memory starts as garbage, like on the device
    }
    G.owner = c;
}

EMU_NOTSAN inline Cta *cta_start(Launch *L, unsigned long long index) {
    Global &G = g();
    Cta *c;
    if (!G.free_ctas.empty()) {
        c = G.free_ctas.back();
        G.free_ctas.pop_back();
    } else {
        c = new Cta();
    }
    const unsigned n = L->block.x * L->block.y * L->block.z;
    if (n == 0 || n > kMaxThreads) die("bad block size");
    const size_t need = (size_t)n * kStackBytes;
    c->stacks = nullptr;
    for (size_t i = 0; i < G.free_stacks.size(); i++)
        if (G.free_stacks[i].second >= need) {
            c->stacks = G.free_stacks[i].first;
            c->stack_bytes = G.free_stacks[i].second;
            G.free_stacks.erase(G.free_stacks.begin() + i);
            break;
        }
    if (!c->stacks) {
        c->stack_bytes = std::max(need, (size_t)256 * kStackBytes);
        c->stacks = (char *)mmap(nullptr, c->stack_bytes, PROT_READ | PROT_WRITE,
                                 MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (c->stacks == (char *)MAP_FAILED) die("cannot map fiber stacks");
        EMU_TSAN(if (tsan_grid_mode())
                     AnnotateBenignRaceSized(__FILE__, __LINE__, c->stacks, (long)c->stack_bytes, "fiber stacks");)
    }
#if defined(__SANITIZE_ADDRESS__)
    // FZB_EMU_ASAN build: fibers of the previous CTA never return (they switch away for good), so the red zones of
    // their frames are still poisoned in the shadow of this stack region
    __asan_unpoison_memory_region(c->stacks, need);
#endif
    c->bid.x = (unsigned)(index % L->grid.x);
    c->bid.y = (unsigned)((index / L->grid.x) % L->grid.y);
    c->bid.z = (unsigned)(index / ((unsigned long long)L->grid.x * L->grid.y));
    c->nthreads = c->alive = n;
    c->bar_arrived = 0;
    c->bar_or_acc = 0;
    c->dyn_bytes = L->smem;
    c->has_snapshot = false;
    c->progressed = false;
    for (unsigned w = 0; w < (n + 31) / 32; w++) {
        c->warp[w] = Warp();
        const unsigned cnt = std::min(32u, n - 32 * w);
        c->warp[w].alive = cnt == 32 ? 0xFFFFFFFFu : ((1u << cnt) - 1u);
    }
    for (unsigned i = 0; i < n; i++) {
        Fiber &f = c->fib[i];
        f.done = false;
        f.spun = false;
        f.wait_word = nullptr;
        f.lin = i;
        f.tid.x = i % L->block.x;
        f.tid.y = (i / L->block.x) % L->block.y;
        f.tid.z = i / (L->block.x * L->block.y);
        uint64_t *top = reinterpret_cast<uint64_t *>(c->stacks + (size_t)(i + 1) * kStackBytes);  // 16-byte aligned
        top[-1] = 0;                                                                             // fake return address
        top[-2] = reinterpret_cast<uint64_t>(&fzb_emu_trampoline);
        for (int r = 3; r <= 8; r++) top[-r] = 0;  // rbp rbx r12 r13 r14 r15
        f.sp = top - 8;
        EMU_TSAN(f.tsan = __tsan_create_fiber(0);)
    }
    return c;
}
EMU_NOTSAN inline void cta_release(Cta *c) {
    Global &G = g();
    if (G.owner == c) G.owner = nullptr;
    EMU_TSAN(for (unsigned i = 0; i < c->nthreads; i++) __tsan_destroy_fiber(c->fib[i].tsan);)
    G.free_stacks.push_back(std::make_pair(c->stacks, c->stack_bytes));
    c->stacks = nullptr;
    G.free_ctas.push_back(c);
}

inline int sched_order() {
    const char *e = getenv("FZB_EMU_SCHED");
    if (!e) return 0;
    return strcmp(e, "reverse") == 0 ? 1 : (strcmp(e, "random") == 0 ? 2 : 0);
}
inline uint64_t sched_rand() {
    static uint64_t x = 0x9E3779B97F4A7C15ull;
    x ^= x << 13;
    x ^= x >> 7;
    x ^= x << 17;
    return x;
}

// run the CTA until it has finished (true) or every thread that can run merely spins (false)
EMU_NOTSAN inline bool cta_run(Launch *L, Cta *c) {
    Global &G = g();
    make_owner(c);
    G.cur_launch = L;
    G.cur_cta = c;
    EMU_TSAN(G.sched_tsan = __tsan_get_current_fiber();)
#if defined(__SANITIZE_THREAD__)
    // The emulator reuses its stacks and its shared-memory storage from CTA to CTA.  Default mode: each CTA is ordered
    // after the previous one (scope = compute-sanitizer racecheck's: the threads of ONE CTA, but over shared AND
    // global memory).  FZB_EMU_TSAN_GRID=1: the CTAs of a grid stay UNORDERED, as on the device, and the reused
    // storage is declared race-free instead -- that pass looks for races BETWEEN CTAs in global memory (and is blind
    // to shared memory).
    // (A CTA that was set aside and is resumed here gets the same treatment: whatever ran meanwhile is ordered before
    // the rest of it; the fibers pick the edge up when they are switched in again.)
    if (!tsan_grid_mode()) {
        __tsan_acquire(&G.done_token);
        __tsan_release(&G.launch_token);
    }
#endif
    c->progressed = false;
    bool finished = true;
    const int order = sched_order();
    while (c->alive) {
        bool ran = false, progress = false;
        const unsigned rot = order == 2 ? (unsigned)(sched_rand() % c->nthreads) : 0u;
        for (unsigned j = 0; j < c->nthreads; j++) {
            // thread order of a sweep: ascending (default), descending, or rotated by a random amount per sweep with a
            // random direction -- FZB_EMU_SCHED=reverse|random: results must not depend on it (a missing barrier does)
            unsigned i = j;
            if (order == 1) i = c->nthreads - 1 - j;
            if (order == 2) i = ((rot & 1u) ? c->nthreads - 1 - j : j), i = (i + rot) % c->nthreads;
            Fiber &f = c->fib[i];
            if (f.done) continue;
            if (f.wait_word && *f.wait_word == f.wait_val) continue;
            G.cur = &f;
            ran = true;
            f.spun = false;
            EMU_TSAN(__tsan_switch_to_fiber(f.tsan, __tsan_switch_to_fiber_no_sync);)
            fzb_emu_switch(&G.sched_sp, f.sp);
            if (!f.spun) progress = true;
        }
        if (!ran) die("deadlock: every live thread of the CTA is blocked (divergent barrier or collective?)");
        if (!progress) {
            finished = false;
            break;
        }
        c->progressed = true;
    }
    G.cur = nullptr;
    G.cur_cta = nullptr;
    G.cur_launch = nullptr;
    return finished;
}

// advance a launch: true = every CTA has finished; false = what is left of it waits for somebody else
EMU_NOTSAN inline bool launch_run(Launch *L) {
    for (;;) {
        bool any = false;
        for (size_t i = 0; i < L->live.size();) {
            Cta *c = L->live[i];
            const bool fin = cta_run(L, c);
            any |= c->progressed;
            if (fin) {
                cta_release(c);
                L->live.erase(L->live.begin() + i);
            } else {
                i++;
            }
        }
        while (L->next < L->total && L->live.size() < kMaxLiveCtas) {
            Cta *c = cta_start(L, L->next++);
            any = true;
            if (cta_run(L, c))
                cta_release(c);
            else
                L->live.push_back(c);
        }
        if (L->live.empty() && L->next == L->total) return true;
        if (!any) return false;
    }
}

// give every suspended launch another go (any host thread that enters the emulator does this)
EMU_NOTSAN inline void pump_locked() {
    Global &G = g();
    for (size_t i = 0; i < G.pending.size();) {
        emu_stream *S = G.pending[i];
        if (launch_run(S->pending)) {
            EMU_TSAN(__tsan_acquire(&G.done_token); __tsan_release(&G.done_token);)
            delete S->pending;
            S->pending = nullptr;
            G.pending.erase(G.pending.begin() + i);
        } else {
            i++;
        }
    }
}
// stream order: whatever comes next on S waits for the launch S still has in flight
inline void drain_locked(std::unique_lock<std::mutex> &lk, emu_stream *S) {
    for (;;) {
        pump_locked();
        if (!S->pending) return;
        lk.unlock();
        usleep(50);  // let the host thread whose kernels are being waited for take the lock
        lk.lock();
    }
}
inline void drain_all_locked(std::unique_lock<std::mutex> &lk) {
    for (;;) {
        pump_locked();
        if (g().pending.empty()) return;
        lk.unlock();
        usleep(50);
        lk.lock();
    }
}

// one kernel launch: CTAs in order (a CTA that waits is set aside), arguments evaluated once by the caller
template <class Body>
EMU_NOTSAN inline void launch(dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Body &&body) {
    Global &G = g();
    std::unique_lock<std::mutex> lk(G.mu);
    if (G.cur) die("nested launch");
    if (smem > kDynSmemBytes) die("dynamic shared memory request too large");
    emu_stream *S = stream ? stream : &G.default_stream;
    drain_locked(lk, S);
    Launch *L = new Launch();
    L->grid = grid;
    L->block = block;
    L->smem = smem;
    L->body = std::forward<Body>(body);
    L->total = (unsigned long long)grid.x * grid.y * grid.z;
    EMU_TSAN(__tsan_release(&G.launch_token);)
    if (launch_run(L)) {
        EMU_TSAN(__tsan_acquire(&G.done_token);)
        delete L;
    } else {
        S->pending = L;
        G.pending.push_back(S);
    }
}

// ---- warp collectives ---------------------------------------------------------------------------------------
EMU_NOTSAN inline Rendezvous &rv_arrive(uint32_t mask, int op, uint64_t v, int *lane_out) {
    Global &G = g();
    Fiber *f = G.cur;
    const int lane = f->lin & 31;
    *lane_out = lane;
    Warp &w = G.cur_cta->warp[f->lin >> 5];
    if (!(mask & (1u << lane))) die("collective: the calling lane is not in its own mask");
    Rendezvous *r = nullptr;
    for (int i = 0; i < kRendezvous && !r; i++)
        if (w.rv[i].mask == mask && !w.rv[i].complete && !(w.rv[i].arrived & (1u << lane))) r = &w.rv[i];
    if (!r) {
        for (int i = 0; i < kRendezvous && !r; i++)
            if (w.rv[i].mask == 0) r = &w.rv[i];
        if (!r) die("collective: too many concurrent rendezvous in one warp");
        r->mask = mask;
        r->arrived = r->departed = r->complete = 0;
        r->op = op;
    }
    if (r->op != op) die("collective: lanes of one mask met in different operations (divergence bug)");
    r->val[lane] = v;
    r->arrived |= 1u << lane;
    const uint32_t gen = r->gen;
    // every *_sync primitive is a convergence point: no named lane leaves before all have arrived, so what a lane
    // did before it cannot collide with what another does after it (TSan: release on arrival, acquire on leaving)
    EMU_TSAN(__tsan_release(&w.alive);)
    rv_check_complete(w, *r);
    if (!r->complete) block_on(&r->gen, gen);
    EMU_TSAN(__tsan_acquire(&w.alive);)
    return *r;
}
EMU_NOTSAN inline void rv_depart(Rendezvous &r, int lane) {
    r.departed |= 1u << lane;
    if (r.departed == r.arrived) r.mask = 0;  // free
}

template <class T>
inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "collective operand wider than 64 bits");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
template <class T>
EMU_NOTSAN inline T shfl_generic(uint32_t mask, T v, int op, int arg, int width) {
    int lane;
    Rendezvous &r = rv_arrive(mask, op, to_bits(v), &lane);
    const int seg = lane & ~(width - 1);
    int src;
    bool ok;
    if (op == OP_SHFL) {
        src = seg + (arg & (width - 1));
        ok = true;
    } else if (op == OP_SHFL_UP) {
        src = lane - arg;
        ok = src >= seg;
    } else if (op == OP_SHFL_DOWN) {
        src = lane + arg;
        ok = src < seg + width;
    } else {
        src = lane ^ arg;
        ok = src < seg + width;
    }
    T out = v;
    if (ok && (r.arrived & (1u << src))) out = from_bits<T>(r.val[src]);
    rv_depart(r, lane);
    return out;
}

}  // namespace emu

#define threadIdx (emu::g().cur->tid)
#define blockIdx (emu::g().cur_cta->bid)
#define blockDim (emu::g().cur_launch->block)
#define gridDim (emu::g().cur_launch->grid)
#define warpSize 32

EMU_NOTSAN static inline void __syncthreads() {
    emu::Cta &c = *emu::g().cur_cta;
    const uint32_t gen = c.bar_gen;
    EMU_TSAN(__tsan_release(&c.bar_token);)
    if (++c.bar_arrived == c.alive) {
        c.bar_arrived = 0;
        c.bar_gen++;
        EMU_TSAN(__tsan_acquire(&c.bar_token);)
        return;
    }
    emu::block_on(&c.bar_gen, gen);
    EMU_TSAN(__tsan_acquire(&c.bar_token);)
}
EMU_NOTSAN EMU_NOTSAN static inline int __syncthreads_or(int pred) {  // barrier + OR-reduction of the predicate (BAR.RED.OR)
    emu::Cta &c = *emu::g().cur_cta;
    const uint32_t gen = c.bar_gen;
    if (pred) c.bar_or_acc = 1;
    EMU_TSAN(__tsan_release(&c.bar_token);)
    if (++c.bar_arrived == c.alive) {
        c.bar_or_res[gen & 1u] = c.bar_or_acc;
        c.bar_or_acc = 0;
        c.bar_arrived = 0;
        c.bar_gen++;
    } else {
        emu::block_on(&c.bar_gen, gen);
    }
    EMU_TSAN(__tsan_acquire(&c.bar_token);)
    return c.bar_or_res[gen & 1u];
}
static inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) {
    int lane;
    emu::Rendezvous &r = emu::rv_arrive(mask, emu::OP_SYNCWARP, 0, &lane);
    emu::rv_depart(r, lane);
}
template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    return emu::shfl_generic(mask, v, emu::OP_SHFL, src, width);
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    return emu::shfl_generic(mask, v, emu::OP_SHFL_UP, (int)delta, width);
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    return emu::shfl_generic(mask, v, emu::OP_SHFL_DOWN, (int)delta, width);
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    return emu::shfl_generic(mask, v, emu::OP_SHFL_XOR, lanemask, width);
}
EMU_NOTSAN static inline unsigned __ballot_sync(unsigned mask, int pred) {
    int lane;
    emu::Rendezvous &r = emu::rv_arrive(mask, emu::OP_BALLOT, pred ? 1 : 0, &lane);
    unsigned out = 0;
    for (int i = 0; i < 32; i++)
        if ((r.arrived >> i) & 1u) out |= (unsigned)(r.val[i] & 1u) << i;
    emu::rv_depart(r, lane);
    return out;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
EMU_NOTSAN static inline int __all_sync(unsigned mask, int pred) {
    int lane;
    emu::Rendezvous &r = emu::rv_arrive(mask, emu::OP_ALL, pred ? 1 : 0, &lane);
    int out = 1;
    for (int i = 0; i < 32; i++)
        if (((r.arrived >> i) & 1u) && !(r.val[i] & 1u)) out = 0;
    emu::rv_depart(r, lane);
    return out;
}
template <class T>
EMU_NOTSAN static inline unsigned __match_any_sync(unsigned mask, T v) {
    int lane;
    const uint64_t mine = emu::to_bits(v);
    emu::Rendezvous &r = emu::rv_arrive(mask, emu::OP_MATCH, mine, &lane);
    unsigned out = 0;
    for (int i = 0; i < 32; i++)
        if (((r.arrived >> i) & 1u) && r.val[i] == mine) out |= 1u << i;
    emu::rv_depart(r, lane);
    return out;
}
template <class T>
EMU_NOTSAN static inline T __reduce_add_sync(unsigned mask, T v) {
    int lane;
    emu::Rendezvous &r = emu::rv_arrive(mask, emu::OP_REDUCE, emu::to_bits(v), &lane);
    T out = 0;
    for (int i = 0; i < 32; i++)
        if ((r.arrived >> i) & 1u) out += emu::from_bits<T>(r.val[i]);
    emu::rv_depart(r, lane);
    return out;
}

// ---- atomics (one OS thread executes kernels at a time; other host threads only read results after a launch) ----
template <class T, class U>
static inline T atomicAdd(T *p, U v) { return __atomic_fetch_add(p, (T)v, EMU_ATOMIC_ORDER); }
template <class T, class U>
static inline T atomicOr(T *p, U v) { return __atomic_fetch_or(p, (T)v, EMU_ATOMIC_ORDER); }
template <class T, class U>
static inline T atomicAnd(T *p, U v) { return __atomic_fetch_and(p, (T)v, EMU_ATOMIC_ORDER); }
template <class T, class U>
static inline T atomicExch(T *p, U v) { return __atomic_exchange_n(p, (T)v, EMU_ATOMIC_ORDER); }
template <class T, class U>
static inline T atomicMin(T *p, U v) {
    T old = __atomic_load_n(p, EMU_ATOMIC_ORDER);
    if ((T)v < old) __atomic_store_n(p, (T)v, EMU_ATOMIC_ORDER);  // (one host thread runs kernels at a time)
    return old;
}
template <class T, class U>
static inline T atomicMax(T *p, U v) {
    T old = __atomic_load_n(p, EMU_ATOMIC_ORDER);
    if ((T)v > old) __atomic_store_n(p, (T)v, EMU_ATOMIC_ORDER);
    return old;
}
template <class T, class U, class V>
static inline T atomicCAS(T *p, U cmp, V v) {
    T expected = (T)cmp;
    __atomic_compare_exchange_n(p, &expected, (T)v, false, EMU_ATOMIC_ORDER, EMU_ATOMIC_ORDER);
    return expected;
}

// ---- integer intrinsics -----------------------------------------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned long long __brevll(unsigned long long x) {
    unsigned long long r = 0;
    for (int i = 0; i < 64; i++) r |= ((x >> i) & 1ull) << (63 - i);
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)(v >> (shift & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)((v << (shift & 31)) >> 32);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    const unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
template <class T>
static inline T __ldg(const T *p) { return *p; }
template <class T>
static inline T __ldcg(const T *p) { return *reinterpret_cast<const volatile T *>(p); }
template <>
inline uint4 __ldcg<uint4>(const uint4 *p) { return *p; }
template <>
inline uint2 __ldcg<uint2>(const uint2 *p) { return *p; }
template <>
inline ulonglong2 __ldcg<ulonglong2>(const ulonglong2 *p) { return *p; }
template <>
inline longlong2 __ldcg<longlong2>(const longlong2 *p) { return *p; }
template <class T>
static inline T __ldcs(const T *p) { return *p; }
#if defined(__SANITIZE_THREAD__)
// (TSan does not model atomic_thread_fence: a fence publishes everything before it to whoever fences later)
EMU_NOTSAN static inline void emu_fence() {
    __tsan_release(&emu::g().fence_token);
    __tsan_acquire(&emu::g().fence_token);
}
static inline void __threadfence() { emu_fence(); }
static inline void __threadfence_block() { emu_fence(); }
static inline void __threadfence_system() { emu_fence(); }
#else
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
#endif
namespace emu {
inline long long cycles() {  // a slow clock (0.2 "GHz" of wall time by default): the kernels' bounded spin-waits for a
                             // peer allow for host threads that take turns inside the emulator on a busy box
                             // (FZB_EMU_CLOCK_DIV: ns per tick, e.g. 200 for the sanitizer builds)
    static const long long div = [] {
        const char *e = getenv("FZB_EMU_CLOCK_DIV");
        const long long v = e ? atoll(e) : 0;
        return v > 0 ? v : 5ll;
    }();
    return (long long)(std::chrono::duration_cast<std::chrono::nanoseconds>(
                           std::chrono::steady_clock::now().time_since_epoch()).count() / div);
}
}  // namespace emu
// the kernels read the clock in their bounded spin-waits (a flag another CTA / GPU raises): the natural place to
// hand the CPU to whoever is being waited for
static inline long long clock64() {
    if (emu::g().cur) emu::yield();
    return emu::cycles();
}
static inline size_t __cvta_generic_to_shared(const void *p) {
    return (size_t)(reinterpret_cast<const uint8_t *>(p) - emu::smem_base());
}
static inline void __nanosleep(unsigned) { emu::yield(); }

// ---- the runtime API (host memory stands in for device memory; everything is synchronous) -----------------------
typedef int cudaError_t;
enum {
    cudaSuccess = 0,
    cudaErrorInvalidValue = 1,
    cudaErrorMemoryAllocation = 2,
    cudaErrorNotReady = 600,
    cudaErrorPeerAccessAlreadyEnabled = 704,
    cudaErrorNotSupported = 801
};
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
struct emu_event {
    std::chrono::steady_clock::time_point t;
};
typedef emu_event *cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0 };
enum { cudaHostAllocDefault = 0, cudaHostAllocPortable = 1, cudaHostAllocMapped = 2, cudaHostRegisterDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaEnableDefault = 0 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes {
    cudaMemoryType type;
    int device;
    void *devicePointer;
    void *hostPointer;
};
struct cudaDeviceProp {
    char name[256];
    int multiProcessorCount;
    int major, minor;
    size_t totalGlobalMem;
    size_t sharedMemPerBlockOptin;
    int l2CacheSize;
};
struct cudaIpcMemHandle_t {
    char reserved[64];
};
enum { cudaIpcMemLazyEnablePeerAccess = 1 };

namespace emu {
struct HostRanges {  // what cudaHostAlloc / cudaMallocHost handed out (cudaPointerGetAttributes)
    std::mutex mu;
    std::vector<std::pair<uintptr_t, size_t>> v;
};
inline HostRanges &host_ranges() {
    static HostRanges h;
    return h;
}
inline int sm_count() {
    const char *e = getenv("FZB_EMU_SMS");
    const int n = e ? atoi(e) : 0;
    return n > 0 ? n : 4;
}
// fault injection for the host error paths: FZB_EMU_FAIL_ALLOC=N makes the N-th allocation from now on fail
// (device or pinned), once; the counter re-arms whenever the variable's value changes
inline bool alloc_should_fail() {
    static std::mutex mu;
    static std::string armed;
    static long left = 0;
    const char *e = getenv("FZB_EMU_FAIL_ALLOC");
    std::lock_guard<std::mutex> l(mu);
    if (!e || !*e) {
        armed.clear();
        return false;
    }
    if (armed != e) {
        armed = e;
        left = atol(e);
    }
    if (left <= 0) return false;
    return --left == 0;
}
inline long &live_allocations() {  // device + pinned allocations not yet freed (leak checks of the error paths)
    static long n = 0;
    return n;
}
inline void *alloc_bytes(size_t n, int fill) {
    if (alloc_should_fail()) return nullptr;
    void *p = nullptr;
    if (posix_memalign(&p, 1024, n ? n : 1)) return nullptr;
    __atomic_add_fetch(&live_allocations(), 1, __ATOMIC_SEQ_CST);
    memset(p, fill, std::min(n, (size_t)8 << 20));  // (the tail of a huge buffer stays untouched: lazily mapped pages)
    return p;
}
}  // namespace emu

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) {
    *n = getenv("FZB_EMU_NO_DEVICE") ? 0 : 1;
    return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "cuda-emu (CPU)");
    p->multiProcessorCount = emu::sm_count();
    p->major = 10;
    p->minor = 0;
    p->totalGlobalMem = (size_t)8 << 30;
    p->sharedMemPerBlockOptin = emu::kDynSmemBytes;
    p->l2CacheSize = 126 << 20;
    return cudaSuccess;
}
template <class T>
static inline cudaError_t cudaMalloc(T **p, size_t n) {
    *p = (T *)emu::alloc_bytes(n, 0xCD);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaDeviceSynchronize();
static inline cudaError_t cudaFree(void *p) {
    cudaDeviceSynchronize();  // as on the device: no kernel may still be using it
    if (p) __atomic_sub_fetch(&emu::live_allocations(), 1, __ATOMIC_SEQ_CST);
    free(p);
    return cudaSuccess;
}
template <class T>
static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) {
    *p = (T *)emu::alloc_bytes(n, 0xCD);
    if (!*p) return cudaErrorMemoryAllocation;
    emu::HostRanges &h = emu::host_ranges();
    std::lock_guard<std::mutex> l(h.mu);
    h.v.push_back(std::make_pair((uintptr_t)*p, n));
    return cudaSuccess;
}
template <class T>
static inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaHostAlloc(p, n, 0); }
static inline cudaError_t cudaFreeHost(void *p) {
    emu::HostRanges &h = emu::host_ranges();
    {
        std::lock_guard<std::mutex> l(h.mu);
        for (size_t i = 0; i < h.v.size(); i++)
            if (h.v[i].first == (uintptr_t)p) {
                h.v.erase(h.v.begin() + i);
                break;
            }
    }
    if (p) __atomic_sub_fetch(&emu::live_allocations(), 1, __ATOMIC_SEQ_CST);
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) {
    emu::HostRanges &h = emu::host_ranges();
    std::lock_guard<std::mutex> l(h.mu);
    a->type = cudaMemoryTypeUnregistered;
    a->device = 0;
    a->devicePointer = a->hostPointer = const_cast<void *>(p);
    for (size_t i = 0; i < h.v.size(); i++)
        if ((uintptr_t)p >= h.v[i].first && (uintptr_t)p < h.v[i].first + h.v[i].second) a->type = cudaMemoryTypeHost;
    return cudaSuccess;
}
namespace emu {
inline void stream_op(cudaStream_t stream, const std::function<void()> &fn) {  // stream-ordered host-side operation
    Global &G = g();
    std::unique_lock<std::mutex> lk(G.mu);
    drain_locked(lk, stream ? stream : &G.default_stream);
    fn();
}
}  // namespace emu
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
    emu::stream_op(nullptr, [&]() { memmove(d, s, n); });
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t st = nullptr) {
    emu::stream_op(st, [&]() { memmove(d, s, n); });
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void *d, int v, size_t n) {
    emu::stream_op(nullptr, [&]() { memset(d, v, n); });
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st = nullptr) {
    emu::stream_op(st, [&]() { memset(d, v, n); });
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
    *s = new emu_stream();
    return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t st) {
    emu::stream_op(st, []() {});
    return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
    cudaStreamSynchronize(s);
    delete s;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamQuery(cudaStream_t st) {  // also the host's chance to advance suspended launches
    emu::Global &G = emu::g();
    std::unique_lock<std::mutex> lk(G.mu);
    emu::pump_locked();
    if (!(st ? st : &G.default_stream)->pending) return cudaSuccess;
    lk.unlock();
    usleep(20);  // the caller polls: give the peers' host threads room
    return cudaErrorNotReady;
}
static inline cudaError_t cudaDeviceSynchronize() {
    emu::Global &G = emu::g();
    std::unique_lock<std::mutex> lk(G.mu);
    emu::drain_all_locked(lk);
    return cudaSuccess;
}
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) {
    *e = new emu_event();
    return cudaSuccess;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete e;
    return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st = nullptr) {
    emu::stream_op(st, [&]() { e->t = std::chrono::steady_clock::now(); });
    return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) {
    *n = 2;
    return cudaSuccess;
}
static inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) {
    *can = 0;
    return cudaSuccess;
}
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaErrorNotSupported; }
cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, unsigned long long flags,
                                    cudaDriverEntryPointQueryResult *q);  // cuda.h

// emulator-only export (not part of include/fuzzb200.h): live device + pinned allocations
extern "C" __attribute__((visibility("default"), used)) long fzb_emu_live_allocations() { return emu::live_allocations(); }
