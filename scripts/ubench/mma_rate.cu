// Microbenchmark: issue rate of legacy mma.sync on B200 (cycles per instruction per SM sub-partition).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int KIND>
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else if (KIND == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else if (KIND == 2)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(b0));
    else
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(b0));
}
template <int KIND, int CHAINS>
__global__ void k(int iters, float* out, long long* cyc) {
    float c[CHAINS][4];
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 0x3f800000u, 0x3f000000u};
    for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) mma<KIND>(c[i], a, 0x3f800000u + i, 0x3f000000u);
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int KIND, int CHAINS>
void run(const char* name, int warps) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<KIND, CHAINS><<<148, warps * 32>>>(iters, out, cyc);
    k<KIND, CHAINS><<<148, warps * 32>>>(iters, out, cyc);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_smsp = (double)h / ((double)iters * CHAINS * warps / 4.0);
    printf("%-22s chains %d warps/SM %2d : %6.2f cycles per mma per SMSP (latency-bound single chain = cycles/mma)  %s\n", name, CHAINS, warps,
           per_smsp, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8, 16}) {
        run<0, 1>("tf32 m16n8k8", w); run<0, 4>("tf32 m16n8k8", w); run<0, 8>("tf32 m16n8k8", w);
        run<1, 4>("bf16 m16n8k16", w); run<1, 8>("bf16 m16n8k16", w);
        run<2, 8>("tf32 m16n8k4", w); run<3, 8>("bf16 m16n8k8", w);
    }
    return 0;
}
