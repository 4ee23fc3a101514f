// tg_host_expand.cpp -- dispatcher + worker pool of the compact host-buffer step (see tg_host_expand.h).
// The expansion loop (tg_host_expand_impl.inc) is compiled three times -- baseline x86-64, AVX2 + BMI2, AVX-512 -- into
// tg_host_expand_{base,avx2,avx512}.o; the variant is picked once from the CPU the library runs on (the library is built in
// one container and runs on another box, so nothing here may assume the build machine's instruction set).
#include "tg_host_expand.h"

#include <emmintrin.h>
#include <sched.h>
#include <stdlib.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace tgh {

void expand_range_base(const ExpandCfg&, const ExpandArgs&, int64_t, int64_t);
void expand_range_avx2(const ExpandCfg&, const ExpandArgs&, int64_t, int64_t);
void expand_range_avx512(const ExpandCfg&, const ExpandArgs&, int64_t, int64_t);

typedef void (*expand_fn)(const ExpandCfg&, const ExpandArgs&, int64_t, int64_t);
static const char* g_isa = "unset";
static expand_fn pick() {
    __builtin_cpu_init();
    const char* force = getenv("TG_HOST_ISA");
    const bool avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
                        __builtin_cpu_supports("avx512vbmi") && __builtin_cpu_supports("bmi2");
    const bool avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
    if (force && force[0] == 'b') { g_isa = "base"; return expand_range_base; }
    if (force && force[0] == 'a' && force[3] == '2' && avx2) { g_isa = "avx2"; return expand_range_avx2; }
    if (avx512) { g_isa = "avx512"; return expand_range_avx512; }
    if (avx2) { g_isa = "avx2"; return expand_range_avx2; }
    g_isa = "base";
    return expand_range_base;
}
static expand_fn impl() {
    static expand_fn f = pick();
    return f;
}
const char* isa_name() { impl(); return g_isa; }

void build_vector_tables(ExpandCfg& c) {
    const int W = c.W, H = c.H, Wp = W + 8, OB = (H + 4) * Wp;
    c.nvec = (OB + 63) / 64;
    c.vec_ok = c.nvec <= 32;
    c.vreach = 0;
    for (int v = 0; v < c.nvec && c.vec_ok; v++) {
        int lo = 1 << 30;
        for (int t = 0; t < 64; t++) {
            const int pos = 64 * v + t, row = pos / Wp, col = pos % Wp;
            if (pos < OB && row < H && col >= 4 && col < 4 + W) { const int b = (row * W + col - 4) >> 1; if (b < lo) lo = b; }
        }
        c.vbase[v] = lo == (1 << 30) ? 0 : lo;
        c.vodd[v] = c.vcell[v] = 0;
        for (int t = 0; t < 64; t++) {
            const int pos = 64 * v + t, row = pos / Wp, col = pos % Wp;
            c.vidx[v][t] = 0;
            if (pos < OB && row < H && col >= 4 && col < 4 + W) {
                const int ci = row * W + col - 4, d = (ci >> 1) - c.vbase[v];
                if (d > 63) { c.vec_ok = 0; break; }
                c.vidx[v][t] = (uint8_t)d;
                c.vcell[v] |= 1ull << t;
                if (ci & 1) c.vodd[v] |= 1ull << t;
            }
        }
        if (c.vcell[v] && c.vbase[v] + 64 > c.vreach) c.vreach = c.vbase[v] + 64;
    }
}

void expand_range(const ExpandCfg& c, const ExpandArgs& a, int64_t e0, int64_t e1) { impl()(c, a, e0, e1); }

void stream_fill(uint8_t* dst, size_t bytes, int value) {
    const __m128i x = _mm_set1_epi8((char)value);
    for (size_t i = 0; i + 64 <= bytes; i += 64) {
        _mm_stream_si128((__m128i*)(dst + i), x); _mm_stream_si128((__m128i*)(dst + i + 16), x);
        _mm_stream_si128((__m128i*)(dst + i + 32), x); _mm_stream_si128((__m128i*)(dst + i + 48), x);
    }
    _mm_sfence();
}

int default_threads() {
    if (const char* t = getenv("TG_HOST_THREADS")) { int v = atoi(t); if (v >= 1 && v <= 1024) return v; }
    cpu_set_t cs;
    CPU_ZERO(&cs);
    int cores = 1;
    if (sched_getaffinity(0, sizeof cs, &cs) == 0) cores = CPU_COUNT(&cs);
    else cores = (int)std::thread::hardware_concurrency();
    int local_world = 1;   // one process per GPU (torchrun): the ranks of a box share its cores
    if (const char* t = getenv("LOCAL_WORLD_SIZE")) { int v = atoi(t); if (v >= 1) local_world = v; }
    int n = cores / local_world;
    return n < 1 ? 1 : n;
}

// ---- pool: nthreads - 1 workers + the caller; items are handed out with an atomic counter.  A host step runs one parallel-for
// per chunk, a few hundred microseconds apart: workers spin on the generation counter for a short while before they go to sleep
// on the condition variable, so that back-to-back chunks do not pay a futex wake-up each -----------------------------------------
struct Pool::Impl {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_start;
    std::atomic<uint64_t> generation{0};
    std::atomic<int> active{0};
    std::atomic<int> sleepers{0};
    std::atomic<bool> stop{false};
    void (*fn)(void*, int64_t) = nullptr;
    void* ctx = nullptr;
    int64_t n_items = 0;
    std::atomic<int64_t> next{0};

    void work() {
        for (;;) {
            const int64_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= n_items) break;
            fn(ctx, i);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            int spins = 0;
            while (generation.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_relaxed)) {
                if (++spins < 20000) { _mm_pause(); continue; }        // ~ 1 ms of spinning, then sleep
                std::unique_lock<std::mutex> lk(mu);
                sleepers.fetch_add(1);
                cv_start.wait(lk, [&] { return stop.load() || generation.load() != seen; });
                sleepers.fetch_sub(1);
            }
            if (stop.load()) return;
            seen = generation.load(std::memory_order_acquire);
            work();
            active.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
};

Pool::Pool(int threads) : impl_(new Impl), nthreads_(threads < 1 ? 1 : threads) {
    for (int i = 1; i < nthreads_; i++) impl_->workers.emplace_back([this] { impl_->loop(); });
}
Pool::~Pool() {
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->stop.store(true);
    }
    impl_->cv_start.notify_all();
    for (auto& t : impl_->workers) t.join();
    delete impl_;
}
void Pool::run(int64_t n_items, void (*fn)(void*, int64_t), void* ctx) {
    if (n_items <= 0) return;
    Impl& p = *impl_;
    p.fn = fn; p.ctx = ctx; p.n_items = n_items;
    p.next.store(0, std::memory_order_relaxed);
    p.active.store((int)p.workers.size(), std::memory_order_relaxed);
    {
        std::lock_guard<std::mutex> lk(p.mu);          // (a sleeper checks the generation under this lock)
        p.generation.fetch_add(1, std::memory_order_release);
    }
    if (p.sleepers.load() > 0) p.cv_start.notify_all();
    p.work();
    int spins = 0;
    while (p.active.load(std::memory_order_acquire) != 0) {
        if (++spins < 4000) _mm_pause();
        else std::this_thread::yield();
    }
}

}  // namespace tgh
