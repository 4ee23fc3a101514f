// hostbw.cu -- what bounds the end-to-end (host-buffer) step: the PCIe link, the host memory system or the host cores?
// Measures, on the box it runs on:
//   (1) pinned D2H / H2D copy bandwidth (1..4 concurrent streams, chunked like tg_step_host),
//   (2) CPU store bandwidth into a pinned buffer (plain stores, SSE2 / AVX2 non-temporal stores) for 1..T threads,
//   (3) CPU read bandwidth,
//   (4) both at once (a D2H stream running while the host threads write), which is what a compact-transfer + host-expansion
//       design does.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o hostbw hostbw.cu -lpthread
#include <cuda_runtime.h>
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void fill_plain(uint8_t* p, size_t n, int v) { memset(p, v, n); }
static void fill_nt_sse2(uint8_t* p, size_t n, int v) {
    __m128i x = _mm_set1_epi8((char)v);
    for (size_t i = 0; i + 64 <= n; i += 64) {
        _mm_stream_si128((__m128i*)(p + i), x); _mm_stream_si128((__m128i*)(p + i + 16), x);
        _mm_stream_si128((__m128i*)(p + i + 32), x); _mm_stream_si128((__m128i*)(p + i + 48), x);
    }
    _mm_sfence();
}
__attribute__((target("avx2"))) static void fill_nt_avx2(uint8_t* p, size_t n, int v) {
    __m256i x = _mm256_set1_epi8((char)v);
    for (size_t i = 0; i + 64 <= n; i += 64) {
        _mm256_stream_si256((__m256i*)(p + i), x); _mm256_stream_si256((__m256i*)(p + i + 32), x);
    }
    _mm_sfence();
}
static uint64_t read_sum(const uint8_t* p, size_t n) {
    const uint64_t* q = (const uint64_t*)p;
    uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (size_t i = 0; i + 4 <= n / 8; i += 4) { s0 += q[i]; s1 += q[i + 1]; s2 += q[i + 2]; s3 += q[i + 3]; }
    return s0 + s1 + s2 + s3;
}

template <class F>
static double run_threads(int T, uint8_t* buf, size_t bytes, int reps, F f) {
    std::vector<std::thread> th;
    std::atomic<int> go{0};
    size_t per = (bytes / T) & ~(size_t)4095;
    double t0 = 0;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            while (!go.load()) {}
            for (int r = 0; r < reps; r++) f(buf + (size_t)t * per, per, r);
        });
    t0 = now();
    go.store(1);
    for (auto& x : th) x.join();
    double dt = now() - t0;
    return (double)per * T * reps / dt / 1e9;
}

int main(int argc, char** argv) {
    size_t GB = argc > 1 ? (size_t)atol(argv[1]) : 4;
    int dev = argc > 2 ? atoi(argv[2]) : 0;
    const size_t bytes = GB << 30;
    CK(cudaSetDevice(dev));
    int hw = (int)std::thread::hardware_concurrency();
    cpu_set_t cs;
    CPU_ZERO(&cs);
    sched_getaffinity(0, sizeof cs, &cs);
    int aff = CPU_COUNT(&cs);
    printf("{\"hardware_concurrency\": %d, \"affinity\": %d, \"avx2\": %d, \"avx512f\": %d, \"buffer_gb\": %zu}\n", hw, aff,
           __builtin_cpu_supports("avx2"), __builtin_cpu_supports("avx512f"), GB);
    uint8_t *h = nullptr, *d = nullptr;
    CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
    CK(cudaMalloc(&d, bytes));
    CK(cudaMemset(d, 1, bytes));
    memset(h, 0, bytes);
    cudaStream_t s[4];
    for (int i = 0; i < 4; i++) CK(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));

    // (1) link
    for (int dir = 0; dir < 2; dir++)
        for (int ns = 1; ns <= 4; ns *= 2)
            for (size_t chunk_mb : {64, 1024}) {
                size_t chunk = chunk_mb << 20;
                double best = 0;
                for (int rep = 0; rep < 3; rep++) {
                    double t0 = now();
                    int k = 0;
                    for (size_t o = 0; o < bytes; o += chunk, k++)
                        CK(cudaMemcpyAsync(dir ? (void*)(d + o) : (void*)(h + o), dir ? (void*)(h + o) : (void*)(d + o), chunk,
                                           dir ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, s[k % ns]));
                    for (int i = 0; i < 4; i++) CK(cudaStreamSynchronize(s[i]));
                    double g = bytes / (now() - t0) / 1e9;
                    if (g > best) best = g;
                }
                printf("{\"test\": \"%s\", \"streams\": %d, \"chunk_mb\": %zu, \"gb_s\": %.1f}\n", dir ? "h2d_pinned" : "d2h_pinned", ns, chunk_mb, best);
            }
    // both directions at once
    {
        double t0 = now();
        CK(cudaMemcpyAsync(h, d, bytes / 2, cudaMemcpyDeviceToHost, s[0]));
        CK(cudaMemcpyAsync(d + bytes / 2, h + bytes / 2, bytes / 2, cudaMemcpyHostToDevice, s[1]));
        CK(cudaStreamSynchronize(s[0])); CK(cudaStreamSynchronize(s[1]));
        printf("{\"test\": \"d2h+h2d concurrently\", \"gb_s_total\": %.1f}\n", bytes / (now() - t0) / 1e9);
    }
    // (2) CPU stores into the pinned buffer
    const bool avx2 = __builtin_cpu_supports("avx2");
    std::vector<int> Ts;
    for (int t = 1; t < aff; t *= 2) Ts.push_back(t);
    Ts.push_back(aff);
    for (int T : Ts) {
        double a = run_threads(T, h, bytes, 2, [](uint8_t* p, size_t n, int r) { fill_plain(p, n, r); });
        double b = run_threads(T, h, bytes, 2, [](uint8_t* p, size_t n, int r) { fill_nt_sse2(p, n, r); });
        double c = avx2 ? run_threads(T, h, bytes, 2, [](uint8_t* p, size_t n, int r) { fill_nt_avx2(p, n, r); }) : 0;
        std::atomic<uint64_t> sink{0};
        double r = run_threads(T, h, bytes, 2, [&](uint8_t* p, size_t n, int) { sink += read_sum(p, n); });
        printf("{\"test\": \"cpu pinned\", \"threads\": %d, \"memset_gb_s\": %.1f, \"nt_sse2_gb_s\": %.1f, \"nt_avx2_gb_s\": %.1f, \"read_gb_s\": %.1f}\n", T, a, b, c, r);
        fflush(stdout);
    }
    // pageable memory with transparent huge pages, for comparison (TLB reach)
    {
        uint8_t* m = (uint8_t*)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        madvise(m, bytes, MADV_HUGEPAGE);
        memset(m, 0, bytes);
        for (int T : {1, aff}) {
            double b = run_threads(T, m, bytes, 2, [](uint8_t* p, size_t n, int r) { fill_nt_sse2(p, n, r); });
            printf("{\"test\": \"cpu pageable+THP\", \"threads\": %d, \"nt_sse2_gb_s\": %.1f}\n", T, b);
        }
        cudaError_t e = cudaHostRegister(m, bytes, cudaHostRegisterDefault);
        if (e == cudaSuccess) {
            double t0 = now();
            CK(cudaMemcpyAsync(m, d, bytes, cudaMemcpyDeviceToHost, s[0]));
            CK(cudaStreamSynchronize(s[0]));
            printf("{\"test\": \"d2h into registered THP memory\", \"gb_s\": %.1f}\n", bytes / (now() - t0) / 1e9);
            double b = run_threads(aff, m, bytes, 2, [](uint8_t* p, size_t n, int r) { fill_nt_sse2(p, n, r); });
            printf("{\"test\": \"cpu registered THP\", \"threads\": %d, \"nt_sse2_gb_s\": %.1f}\n", aff, b);
            cudaHostUnregister(m);
        } else {
            printf("{\"test\": \"cudaHostRegister THP\", \"error\": \"%s\"}\n", cudaGetErrorString(e));
            cudaGetLastError();
        }
        munmap(m, bytes);
    }
    // (4) a compact D2H stream (1/8 of the bytes) while the host threads rewrite the whole buffer
    for (int T : {aff > 2 ? aff - 1 : aff, aff}) {
        double t0 = now();
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaMemcpyAsync(h, d, bytes / 8, cudaMemcpyDeviceToHost, s[0]));
        }
        double cpu = run_threads(T, h + bytes / 8, bytes - bytes / 8, 2, [&](uint8_t* p, size_t n, int r) { if (avx2) fill_nt_avx2(p, n, r); else fill_nt_sse2(p, n, r); });
        double t_cpu = now() - t0;
        CK(cudaStreamSynchronize(s[0]));
        double t_all = now() - t0;
        printf("{\"test\": \"d2h(1/8) || cpu nt stores\", \"threads\": %d, \"cpu_gb_s\": %.1f, \"cpu_s\": %.3f, \"all_s\": %.3f}\n", T, cpu, t_cpu, t_all);
    }
    // full-dict D2H while all host threads ALSO write (contention on host memory)
    {
        double t0 = now();
        CK(cudaMemcpyAsync(h, d, bytes / 2, cudaMemcpyDeviceToHost, s[0]));
        double cpu = run_threads(aff, h + bytes / 2, bytes / 2, 2, [&](uint8_t* p, size_t n, int r) { if (avx2) fill_nt_avx2(p, n, r); else fill_nt_sse2(p, n, r); });
        double t_cpu = now() - t0;
        CK(cudaStreamSynchronize(s[0]));
        double t_all = now() - t0;
        printf("{\"test\": \"d2h(half) || cpu nt stores(half x2)\", \"cpu_gb_s\": %.1f, \"d2h_gb_s\": %.1f, \"cpu_s\": %.3f, \"all_s\": %.3f}\n", cpu, bytes / 2 / t_all / 1e9, t_cpu, t_all);
    }
    return 0;
}
