// Microbenchmark: how fast can one CTA per SM stream HBM into a shared-memory ring with 1-D bulk async
// copies (cp.async.bulk + mbarrier), as a function of tile size, ring depth and number of issuing lanes?
// Consumers do nothing but wait(full) -> arrive(empty), so this is the supply ceiling of the R2 sweep.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ bool mb_try(uint64_t* b, uint32_t ph) {
    uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory"); return ok; }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t ph) { while (!mb_try(b, ph)) {} }
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory"); }

// warp 0: producer lanes 0..NP-1 (lane l issues tiles l, l+NP, ...); warps 1..NC: consumers, tile i -> warp i % NC
__global__ void __launch_bounds__(512, 1) k_supply(const char* src, long long tiles_per_cta, int tile_bytes, int nst, int NP, int NC,
                                                   int passes, int work, float* sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* full = (uint64_t*)sm; uint64_t* empty = full + 64;
    unsigned char* ring = sm + 1024;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < nst; ++i) { mb_init(&full[i], 1); mb_init(&empty[i], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const char* base = src + (size_t)blockIdx.x * tiles_per_cta * tile_bytes;
    const long long total = tiles_per_cta * passes;
    if (warp == 0) {
        if (lane < NP) {
            for (long long it = lane; it < total; it += NP) {
                const int st = (int)(it % nst); const uint32_t ph = (uint32_t)((it / nst) & 1);
                if (it >= nst) mb_wait(&empty[st], ph ^ 1u);
                mb_expect(&full[st], tile_bytes);
                bulk(ring + (size_t)st * tile_bytes, base + (size_t)(it % tiles_per_cta) * tile_bytes, tile_bytes, &full[st]);
            }
        }
    } else if (warp <= NC) {
        float acc = 0.f;
        for (long long it = warp - 1; it < total; it += NC) {
            const int st = (int)(it % nst); const uint32_t ph = (uint32_t)((it / nst) & 1);
            mb_wait(&full[st], ph);
            const float* t = (const float*)(ring + (size_t)st * tile_bytes);
            for (int w = 0; w < work; ++w) acc += t[(lane * 4 + w * 128) % (tile_bytes / 4)];   // optional fake work
            __syncwarp();
            if (lane == 0) mb_arrive(&empty[st]);
        }
        if (acc == 123.456f) sink[0] = acc;
    }
}

// reference: plain vectorised loads
__global__ void k_ldg(const float4* src, long long n4, float* sink) {
    float4 a = {0, 0, 0, 0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldcs(src + i); a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    if (a.x + a.y + a.z + a.w == 123.456f) sink[0] = a.x;
}

int main() {
    const size_t bytes = 132ull << 20;
    char* d; cudaMalloc(&d, bytes + (1 << 20)); cudaMemset(d, 0, bytes + (1 << 20));
    float* sink; cudaMalloc(&sink, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(k_supply, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const int G = 148, passes = 6;
    struct Cfg { int tile, nst, NP, NC, work; };
    std::vector<Cfg> cfgs;
    for (int tile : {3648, 7296, 14592, 29184})
        for (int budget : {96, 176, 216})
            for (int NP : {1, 4})
                cfgs.push_back({tile, (budget * 1024) / tile > 64 ? 64 : (budget * 1024) / tile, NP, 14, 0});
    cfgs.push_back({7296, 24, 1, 14, 64}); cfgs.push_back({7296, 24, 1, 14, 256}); cfgs.push_back({7296, 29, 1, 14, 0});
    cfgs.push_back({7296, 24, 1, 4, 0}); cfgs.push_back({7296, 24, 1, 1, 0}); cfgs.push_back({7296, 24, 8, 14, 0});
    setvbuf(stdout, nullptr, _IONBF, 0);
    for (auto c : cfgs) {
        if (c.nst < 4) continue;
        if (c.NC > c.nst - 2) c.NC = c.nst - 2;      // a consumer may never be a full ring revolution ahead of the copies
        const long long tiles_per_cta = (long long)(bytes / c.tile) / G;
        const size_t smem = 1024 + (size_t)c.nst * c.tile;
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k_supply<<<G, 512, smem>>>(d, tiles_per_cta, c.tile, c.nst, c.NP, c.NC, passes, c.work, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        cudaError_t err = cudaGetLastError();
        const double gb = (double)tiles_per_cta * c.tile * G * passes / 1e9;
        printf("tile %6d B  stages %2d (%3zu KB)  producers %d consumers %2d work %3d : %7.1f us/pass  %7.1f GB/s  %s\n", c.tile, c.nst,
               smem >> 10, c.NP, c.NC, c.work, best * 1e3 / passes, gb / (best * 1e-3), err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
    for (int blocks : {148 * 2, 148 * 4, 148 * 8}) {
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0); k_ldg<<<blocks, 512>>>((const float4*)d, bytes / 16, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("ldg float4 %4d blocks: %7.1f us  %7.1f GB/s\n", blocks, best * 1e3, bytes / 1e9 / (best * 1e-3));
    }
    return 0;
}
