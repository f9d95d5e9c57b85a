// Micro-test of the tcgen05 (UMMA) building blocks the streaming engine uses, against a CPU reference:
//   forward : D[128 rows x 8 chains]  = X[128 x 56] * B[56 x 8]      A K-major,  B K-major   (no swizzle)
//   backward: G[cols x 8 chains]      = X^T[cols x 128] * R[128 x 8]  A MN-major, B MN-major (same X tile in smem)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_test umma_test.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cstring>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ROWS = 128, COLS = 56, NCH = 8;
constexpr int JCH = COLS / 4;                       // 16-byte column chunks
// X tile, core-matrix layout: byte offset of (r, c)
__host__ __device__ inline int x_off(int r, int c) { return ((c >> 2) * 16 + (r >> 3)) * 128 + (r & 7) * 16 + (c & 3) * 4; }
// B (beta) K-major: (n = chain, k = col)
__host__ __device__ inline int b_off(int n, int k) { return (k >> 2) * 128 + n * 16 + (k & 3) * 4; }
// R MN-major: (row, chain)
__host__ __device__ inline int r_off(int row, int ch) { return (ch >> 2) * 2048 + (row >> 3) * 128 + (row & 7) * 16 + (ch & 3) * 4; }

__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                          // version = 1 (Blackwell)
    return d;                                        // layout_type = 0 (no swizzle), base_offset = 0
}
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ inline void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ inline void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ inline void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ inline void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// out: [3][128 lanes][8] = forward D, backward G (M = 128), backward G (M = 64)
__global__ void __launch_bounds__(128, 1) k_test(const float* X, const float* B, const float* R, float* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* xs = smem;                        // 64 KB window (A of the M = 128 backward MMA reads 32 column chunks)
    unsigned char* bs = smem + 65536;                // 14 * 128 B
    unsigned char* rs = bs + 2048;                   // 2 * 2048 B
    uint64_t* bar = (uint64_t*)(rs + 4096);
    uint32_t* tmem_slot = (uint32_t*)(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += 128) ((float*)xs)[i] = 0.0f;
    __syncthreads();
    for (int i = tid; i < ROWS * COLS; i += 128) { const int r = i / COLS, c = i % COLS; *(float*)(xs + x_off(r, c)) = X[i]; }
    for (int i = tid; i < COLS * NCH; i += 128) { const int k = i / NCH, n = i % NCH; *(float*)(bs + b_off(n, k)) = B[i]; }
    for (int i = tid; i < ROWS * NCH; i += 128) { const int r = i / NCH, n = i % NCH; *(float*)(rs + r_off(r, n)) = R[i]; }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        // forward: 7 k-steps of 8 columns
        const uint32_t id_f = make_idesc(128, 8, 0, 0);
        for (int ks = 0; ks < COLS / 8; ++ks)
            umma_tf32(tmem + 0, make_desc(smem_u32(xs) + ks * 2 * 2048, 2048, 128), make_desc(smem_u32(bs) + ks * 2 * 128, 128, 1024), id_f, ks > 0);
        // backward, M = 128: 16 k-steps of 8 rows
        const uint32_t id_b = make_idesc(128, 8, 1, 1);
        for (int ks = 0; ks < ROWS / 8; ++ks)
            umma_tf32(tmem + 8, make_desc(smem_u32(xs) + ks * 128, 128, 2048), make_desc(smem_u32(rs) + ks * 128, 128, 2048), id_b, ks > 0);
        const uint32_t id_b64 = make_idesc(64, 8, 1, 1);
        for (int ks = 0; ks < ROWS / 8; ++ks)
            umma_tf32(tmem + 16, make_desc(smem_u32(xs) + ks * 128, 128, 2048), make_desc(smem_u32(rs) + ks * 128, 128, 2048), id_b64, ks > 0);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int which = 0; which < 3; ++which) {
        uint32_t v[8];
        tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + which * 8, v);
        for (int j = 0; j < 8; ++j) out[(which * 128 + tid) * 8 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(32u) : "memory");
}

static float tf32_trunc(float x) { uint32_t b; memcpy(&b, &x, 4); b &= 0xFFFFE000u; memcpy(&x, &b, 4); return x; }
static float tf32_rn(float x) { uint32_t b; memcpy(&b, &x, 4); b += 0x1000u; b &= 0xFFFFE000u; memcpy(&x, &b, 4); return x; }

int main() {
    std::vector<float> X(ROWS * COLS), B(COLS * NCH), R(ROWS * NCH), out(3 * 128 * 8);
    srand(1);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.0f - 1.0f; };
    for (auto& v : X) v = rnd();
    for (auto& v : B) v = rnd();
    for (auto& v : R) v = rnd();
    float *dX, *dB, *dR, *dO;
    CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dR, R.size() * 4)); CK(cudaMalloc(&dO, out.size() * 4));
    CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dR, R.data(), R.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(dO, 0, out.size() * 4));
    const int smem = 65536 + 2048 + 4096 + 64;
    CK(cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_test<<<1, 128, smem>>>(dX, dB, dR, dO);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
    for (int mode = 0; mode < 2; ++mode) {           // 0: operands truncated to tf32, 1: rounded to nearest
        auto cv = mode ? tf32_rn : tf32_trunc;
        double ef = 0, eb = 0, eb64 = 0;
        for (int r = 0; r < ROWS; ++r) for (int n = 0; n < NCH; ++n) {
            double a = 0; for (int k = 0; k < COLS; ++k) a += (double)cv(X[r * COLS + k]) * cv(B[k * NCH + n]);
            ef = fmax(ef, fabs(a - out[(0 * 128 + r) * 8 + n]));
        }
        for (int c = 0; c < COLS; ++c) for (int n = 0; n < NCH; ++n) {
            double a = 0; for (int r = 0; r < ROWS; ++r) a += (double)cv(X[r * COLS + c]) * cv(R[r * NCH + n]);
            eb = fmax(eb, fabs(a - out[(1 * 128 + c) * 8 + n]));
            // M = 64: try the two candidate row -> lane maps
            const int lane_a = c, lane_b = (c / 16) * 32 + (c % 16);
            eb64 = fmax(eb64, fmin(fabs(a - out[(2 * 128 + lane_a) * 8 + n]), fabs(a - out[(2 * 128 + lane_b) * 8 + n])));
        }
        printf("%s: max abs err forward %.3e  backward(M=128) %.3e  backward(M=64, best map) %.3e\n", mode ? "rn   " : "trunc", ef, eb, eb64);
    }
    // where do the rows of the M = 64 result live?
    { int c = 17, n = 3; double a = 0; for (int r = 0; r < ROWS; ++r) a += (double)tf32_trunc(X[r * COLS + c]) * tf32_trunc(R[r * NCH + n]);
      printf("M=64 probe: G[17][3] = %.6f; lanes holding ~that value in column 3:", a);
      for (int l = 0; l < 128; ++l) if (fabs(out[(2 * 128 + l) * 8 + n] - a) < 1e-2) printf(" %d", l);
      printf("\n"); }
    printf("forward D[5][0..3] = %.6f %.6f %.6f %.6f\n", out[5 * 8], out[5 * 8 + 1], out[5 * 8 + 2], out[5 * 8 + 3]);
    return 0;
}
