// wbw.cu -- write-only HBM bandwidth calibration on B200: what can a store stream reach, by store flavour and chunk size?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/wbw tools/microbench/wbw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void* g, const void* s, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(s)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// classic grid-stride memset, 16 B per thread
__global__ void k_gridstride(uint4* out, size_t n16) {
    uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}
// each warp owns consecutive chunks of `chunk` bytes and writes them in pieces of `piece` bytes (16 B per lane)
__global__ void k_warpchunk(uint8_t* out, size_t nchunks, int chunk, int piece) {
    const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t c = (size_t)blockIdx.x * nw + (threadIdx.x >> 5); c < nchunks; c += (size_t)gridDim.x * nw) {
        uint8_t* g = out + c * chunk;
        for (int o = 0; o < chunk; o += piece)
            for (int q = lane * 16; q < piece && o + q < chunk; q += 512) *(uint4*)(g + o + q) = v;
    }
}
// same ownership, but each piece is one TMA bulk store from shared memory (issued by lane 0, or spread over lanes)
__global__ void k_warpbulk(uint8_t* out, size_t nchunks, int chunk, int piece, int spread) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int lane = threadIdx.x & 31, nw = blockDim.x >> 5, warp = threadIdx.x >> 5;
    uint8_t* buf = sm + (size_t)warp * ((piece + 127) / 128 * 128);
    for (int i = lane * 4; i < piece; i += 128) *(uint32_t*)(buf + i) = 0x01020304u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const int np = chunk / piece;
    for (size_t c = (size_t)blockIdx.x * nw + warp; c < nchunks; c += (size_t)gridDim.x * nw) {
        uint8_t* g = out + c * chunk;
        if (spread) { for (int p = lane; p < np; p += 32) bulk_s2g(g + (size_t)p * piece, buf, piece); }
        else if (lane == 0) { for (int p = 0; p < np; p++) bulk_s2g(g + (size_t)p * piece, buf, piece); }
        bulk_commit();
        bulk_wait_all();
        __syncwarp();
    }
}
// CTA-wide tile: all warps fill one tile of `piece` bytes, thread 0 stores it (the k_step pattern)
__global__ void k_ctabulk(uint8_t* out, size_t ntiles, int piece) {
    extern __shared__ __align__(128) uint8_t sm[];
    for (int i = threadIdx.x * 4; i < piece; i += blockDim.x * 4) *(uint32_t*)(sm + i) = 0x01020304u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) { bulk_s2g(out + t * piece, sm, piece); bulk_commit(); asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); }
        bulk_wait_all();
    }
}


// traffic of the per-call step kernel without any compute: per tile of 32 envs TMA-load hot (1 KB) + board records (4.6 KB) +
// rng (512 B), TMA-store board image + mask image (13.8 KB each) + queue (3.5 KB) + holder (512 B) + hot (1 KB), and for a
// fraction of the envs a 144-byte board record.  One thread per CTA drives everything; 2 load stages; `sets` image sets.
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__global__ void k_mimic(uint8_t* hot, uint8_t* brd, uint8_t* rng, uint8_t* ob, uint8_t* om, uint8_t* oq, uint8_t* oh, size_t ntiles, int sets, int commit_pct) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t* bar = (uint64_t*)sm;              // 2 barriers
    uint8_t* stage = sm + 128;                  // 2 x 6272
    uint8_t* img = stage + 2 * 6272;            // sets x 31744
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x != 0) return;
    auto load = [&](size_t t, int s) {
        mbar_expect_tx(bar + s, 1024 + 4608 + 512);
        bulk_g2s(stage + s * 6272, hot + t * 1024, 1024, bar + s);
        bulk_g2s(stage + s * 6272 + 1024, brd + t * 4608, 4608, bar + s);
        bulk_g2s(stage + s * 6272 + 5632, rng + t * 512, 512, bar + s);
    };
    if (blockIdx.x < ntiles) load(blockIdx.x, 0);
    size_t k = 0;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x, k++) {
        int s = k & 1;
        if (t + gridDim.x < ntiles) load(t + gridDim.x, s ^ 1);
        mbar_wait(bar + s, (k >> 1) & 1);
        uint8_t* im = img + (k % sets) * 31744;
        if (sets == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        else if (sets == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
        bulk_s2g(ob + t * 13824, im, 13824);
        bulk_s2g(om + t * 13824, im + 13824, 13824);
        bulk_s2g(oq + t * 3584, im + 27648, 3584);
        bulk_s2g(oh + t * 512, im + 31232, 512);
        bulk_s2g(hot + t * 1024, stage + s * 6272, 1024);
        if (commit_pct >= 0) {
            for (int e = 0; e < 32; e++)
                if ((int)((t * 32 + e) * 2654435761u % 100u) < commit_pct) bulk_s2g(brd + (t * 32 + e) * 144, stage + s * 6272 + 1024 + e * 144, 144);
        } else {   // pairs: both records of an even/odd pair (288 B = 9 full sectors) when either is dirty
            for (int e = 0; e < 32; e += 2)
                if ((int)((t * 32 + e) * 2654435761u % 100u) < -commit_pct || (int)((t * 32 + e + 1) * 2654435761u % 100u) < -commit_pct)
                    bulk_s2g(brd + (t * 32 + e) * 144, stage + s * 6272 + 1024 + e * 144, 288);
        }
        bulk_commit();
    }
    bulk_wait_all();
}

template <class F> static float timeit(F f, int iters = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < iters; i++) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
    return ms / iters;
}

int main() {
    const size_t bytes = (size_t)4 << 30;
    uint8_t* d; cudaMalloc(&d, bytes + (1 << 20));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float ms = timeit([&] { cudaMemsetAsync(d, 1, bytes); });
    printf("cudaMemset                      %7.0f GB/s\n", bytes / ms / 1e6);
    for (int bpsm : {2, 4, 8}) {
        ms = timeit([&] { k_gridstride<<<sms * bpsm, 256>>>((uint4*)d, bytes / 16); });
        printf("grid-stride STG.128 %d CTA/SM    %7.0f GB/s\n", bpsm, bytes / ms / 1e6);
    }
    for (int chunk : {432 * 40, 6336, 1 << 16}) for (int piece : {432, 512, 2048}) {
        if (chunk % 16 || piece > chunk) continue;
        size_t nch = bytes / chunk;
        ms = timeit([&] { k_warpchunk<<<sms * 4, 256>>>(d, nch, chunk, piece); });
        printf("warp chunk %6d B, STG pieces %5d B      %7.0f GB/s\n", chunk, piece, nch * (size_t)chunk / ms / 1e6);
    }
    cudaFuncSetAttribute(k_warpbulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_ctabulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct { int chunk, piece; } cs[] = {{17280, 432}, {17280, 3456}, {17280, 17280}, {6336, 6336}, {16384, 16384}, {16384, 2048}, {16384, 512}};
    for (auto c : cs) for (int spread : {0, 1}) for (int wpc : {4, 8}) {
        size_t nch = bytes / c.chunk;
        size_t smem = (size_t)wpc * ((c.piece + 127) / 128 * 128);
        int per_sm = (int)((200 * 1024) / (smem + 1024)); if (per_sm > 32 / wpc) per_sm = 32 / wpc; if (per_sm < 1) per_sm = 1;
        ms = timeit([&] { k_warpbulk<<<sms * per_sm, wpc * 32, smem>>>(d, nch, c.chunk, c.piece, spread); });
        printf("warp chunk %6d B, TMA pieces %5d B %s %d warps/CTA x %d CTA/SM   %7.0f GB/s\n", c.chunk, c.piece, spread ? "lanes " : "lane0 ", wpc, per_sm,
               nch * (size_t)c.chunk / ms / 1e6);
    }
    for (int piece : {13824, 32768, 65536}) for (int per_sm : {2, 4}) {
        size_t nt = bytes / piece;
        if ((size_t)piece * per_sm > 200 * 1024) continue;
        ms = timeit([&] { k_ctabulk<<<sms * per_sm, 128, piece>>>(d, nt, piece); });
        printf("CTA tile TMA %6d B, %d CTA/SM   %7.0f GB/s\n", piece, per_sm, nt * (size_t)piece / ms / 1e6);
    }
    {
        const size_t nenv = (size_t)1 << 22, nt = nenv / 32;
        uint8_t *hot, *brd, *rng, *ob, *om, *oq, *oh;
        cudaMalloc(&hot, nenv * 32); cudaMalloc(&brd, nenv * 144); cudaMalloc(&rng, nenv * 16);
        cudaMalloc(&ob, nenv * 432); cudaMalloc(&om, nenv * 432); cudaMalloc(&oq, nenv * 112); cudaMalloc(&oh, nenv * 16);
        cudaFuncSetAttribute(k_mimic, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        for (int sets : {1, 2}) for (int per_sm : {2, 3, 4}) for (int cp : {0, 21, -21}) {
            size_t smem = 128 + 2 * 6272 + (size_t)sets * 31744;
            if (smem * per_sm > 220 * 1024) continue;
            ms = timeit([&] { k_mimic<<<sms * per_sm, 32, smem>>>(hot, brd, rng, ob, om, oq, oh, nt, sets, cp); });
            double bytes_env = 32 + 144 + 16 + 432 * 2 + 112 + 16 + 32 + (cp >= 0 ? 1.44 * cp : 144 * (1 - 0.79 * 0.79));
            printf("step-traffic mimic: %d image set(s), %d CTA/SM, commit %3d%% (negative = record pairs):  %6.2f G env/s  %7.0f GB/s\n", sets, per_sm, cp, nenv / ms / 1e6, nenv * bytes_env / ms / 1e6);
        }
    }
    return 0;
}
