// Layout probe for tcgen05.mma kind::tf32 smem descriptors: the host builds the raw byte images of A and B for a
// layout hypothesis, the kernel issues `nsteps` MMAs with the given descriptor fields, the host checks D.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cstring>
#include <functional>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int ABYTES = 65536, BBYTES = 8192;
struct P { uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_step, b_step, idesc, nsteps, a_lt, b_lt; };
__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    return ((uint64_t)lt << 61) | (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__global__ void __launch_bounds__(128, 1) k_probe(const unsigned char* A, const unsigned char* B, P p, float* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* as = smem; unsigned char* bs = smem + ABYTES;
    uint64_t* bar = (uint64_t*)(bs + BBYTES); uint32_t* slot = (uint32_t*)(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < ABYTES / 4; i += 128) ((uint32_t*)as)[i] = ((const uint32_t*)A)[i];
    for (int i = tid; i < BBYTES / 4; i += 128) ((uint32_t*)bs)[i] = ((const uint32_t*)B)[i];
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(1u) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (tid == 0) {
        for (uint32_t ks = 0; ks < p.nsteps; ++ks) {
            const uint64_t ad = make_desc(smem_u32(as) + ks * p.a_step, p.a_lbo, p.a_sbo, p.a_lt), bd = make_desc(smem_u32(bs) + ks * p.b_step, p.b_lbo, p.b_sbo, p.b_lt);
            const uint32_t acc = ks > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         :: "r"(tmem), "l"(ad), "l"(bd), "r"(p.idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
    }
    { uint32_t ok = 0; while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory"); }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(tmem + ((uint32_t)(warp * 32) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[tid * 8 + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(32u) : "memory");
}
static float tf(float x) { uint32_t b; memcpy(&b, &x, 4); b &= 0xFFFFE000u; memcpy(&x, &b, 4); return x; }
static uint32_t idesc(int M, int N, int a_mn, int b_mn) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

int main() {
    const int M = 128, N = 8, K = 64;               // D[M x N] = A[M x K] B[K x N], K = 8 MMAs of k = 8
    std::vector<float> A(M * K), B(K * N);
    srand(2);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2 - 1;
    for (auto& v : B) v = (float)rand() / RAND_MAX * 2 - 1;
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double a = 0; for (int k = 0; k < K; ++k) a += (double)tf(A[m * K + k]) * tf(B[k * N + n]); ref[m * N + n] = a; }
    unsigned char *dA, *dB; float* dO;
    CK(cudaMalloc(&dA, ABYTES)); CK(cudaMalloc(&dB, BBYTES)); CK(cudaMalloc(&dO, 128 * 8 * 4));
    const int smem = ABYTES + BBYTES + 64;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    using Off = std::function<int(int, int)>;       // byte offset of element (mn, k)
    auto run = [&](const char* name, Off aoff, Off boff, P p) {
        std::vector<unsigned char> ia(ABYTES, 0), ib(BBYTES, 0);
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { int o = aoff(m, k); if (o < 0 || o + 4 > ABYTES) { printf("%s: A offset out of range\n", name); return; } memcpy(&ia[o], &A[m * K + k], 4); }
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { int o = boff(n, k); if (o < 0 || o + 4 > BBYTES) { printf("%s: B offset out of range\n", name); return; } memcpy(&ib[o], &B[k * N + n], 4); }
        CK(cudaMemcpy(dA, ia.data(), ABYTES, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, ib.data(), BBYTES, cudaMemcpyHostToDevice));
        CK(cudaMemset(dO, 0, 128 * 8 * 4));
        k_probe<<<1, 128, smem>>>(dA, dB, p, dO);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
        std::vector<float> out(128 * 8);
        CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
        double err = 0; for (int i = 0; i < M * N; ++i) err = fmax(err, fabs(ref[i] - out[i]));
        printf("%-62s max abs err %.3e %s  out[0..2] %.4f %.4f %.4f ref %.4f %.4f %.4f\n", name, err, err < 1e-4 ? "OK" : "", out[0], out[1], out[2], ref[0], ref[1], ref[2]);
    };
    // K-major, no swizzle: (mn, k) -> (k/4)*LBO + (mn/8)*SBO + (mn%8)*16 + (k%4)*4
    Off a_k = [](int m, int k) { return (k / 4) * 2048 + (m / 8) * 128 + (m % 8) * 16 + (k % 4) * 4; };
    Off b_k = [](int n, int k) { return (k / 4) * 128 + n * 16 + (k % 4) * 4; };
    // MN-major hypothesis H1: (mn, k) -> (mn/4)*S + (k/8)*L + (k%8)*16 + (mn%4)*4, with S = MN-group stride, L = K-group stride
    Off a_mn = [](int m, int k) { return (m / 4) * 1024 + (k / 8) * 128 + (k % 8) * 16 + (m % 4) * 4; };      // 32 MN groups x 1024 B; 8 K groups x 128 B
    Off b_mn = [](int n, int k) { return (n / 4) * 1024 + (k / 8) * 128 + (k % 8) * 16 + (n % 4) * 4; };
    run("A K-major, B K-major", a_k, b_k, P{2048, 128, 128, 1024, 4096, 256, idesc(128, 8, 0, 0), 8});
    run("A MN (LBO=K-grp 128, SBO=MN-grp 1024), B K", a_mn, b_k, P{128, 1024, 128, 1024, 128, 256, idesc(128, 8, 1, 0), 8});
    run("A MN (LBO=MN-grp 1024, SBO=K-grp 128), B K", a_mn, b_k, P{1024, 128, 128, 1024, 128, 256, idesc(128, 8, 1, 0), 8});
    run("A K, B MN (LBO=K-grp 128, SBO=MN-grp 1024)", a_k, b_mn, P{2048, 128, 128, 1024, 4096, 128, idesc(128, 8, 0, 1), 8});
    run("A K, B MN (LBO=MN-grp 1024, SBO=K-grp 128)", a_k, b_mn, P{2048, 128, 1024, 128, 4096, 128, idesc(128, 8, 0, 1), 8});
    // MN-major hypothesis H2: 8 MN x 8 K core matrix? (mn, k) -> (mn/8)*S + (k/8)*L + (k%8)*32 + (mn%8)*4
    Off a_mn2 = [](int m, int k) { return (m / 8) * 2048 + (k / 8) * 256 + (k % 8) * 32 + (m % 8) * 4; };
    run("A MN H2 (8x8 core, LBO=K 256, SBO=MN 2048), B K", a_mn2, b_k, P{256, 2048, 128, 1024, 256, 256, idesc(128, 8, 1, 0), 8});
    run("A MN H2 (8x8 core, LBO=MN 2048, SBO=K 256), B K", a_mn2, b_k, P{2048, 256, 128, 1024, 256, 256, idesc(128, 8, 1, 0), 8});
    // H3: K pairs? (k%4 within 16 B is impossible for MN-major) -- MN-major with K-groups of 4: (mn/4)*S + (k/4)*L + (k%4)*16 + (mn%4)*4
    Off a_mn3 = [](int m, int k) { return (m / 4) * 1024 + (k / 4) * 64 + (k % 4) * 16 + (m % 4) * 4; };
    run("A MN H3 (4-K groups, LBO=K 64, SBO=MN 1024), B K", a_mn3, b_k, P{64, 1024, 128, 1024, 128, 256, idesc(128, 8, 1, 0), 8});
    run("A MN H3 (4-K groups, LBO=MN 1024, SBO=K 64), B K", a_mn3, b_k, P{1024, 64, 128, 1024, 128, 256, idesc(128, 8, 1, 0), 8});
    // ---- swizzled layouts.  sw(o, bits): XOR 16-byte-chunk bits [4, 4+bits) with address bits [7, 7+bits)
    auto sw = [](int o, int bits) { return o ^ (((o >> 7) & ((1 << bits) - 1)) << 4); };
    // SW32, 8-float blocks: (row-ish index x, 8-float block blk, element e) -> blk*BLK + x*32 + e*4
    Off a_k32 = [sw](int m, int k) { return sw((k / 8) * 4096 + m * 32 + (k % 8) * 4, 1); };              // K-major: x = m, blk = k/8
    Off b_k32 = [sw](int n, int k) { return sw((k / 8) * 256 + n * 32 + (k % 8) * 4, 1); };
    Off a_mn32 = [sw](int m, int k) { return sw((m / 8) * 4096 + k * 32 + (m % 8) * 4, 1); };            // MN-major: x = k, blk = m/8 (K <= 128)
    Off b_mn32 = [sw](int n, int k) { return sw(k * 32 + n * 4, 1); };
    run("SW32: A K (SBO=256), B K (SBO=256)", a_k32, b_k32, P{16, 256, 16, 256, 4096, 256, idesc(128, 8, 0, 0), 8, 6, 6});
    run("SW32: A MN (LBO=4096 MN-blk, SBO=256 K-grp), B K sw32", a_mn32, b_k32, P{4096, 256, 16, 256, 256, 256, idesc(128, 8, 1, 0), 8, 6, 6});
    run("SW32: A MN (LBO=256, SBO=4096), B K sw32", a_mn32, b_k32, P{256, 4096, 16, 256, 256, 256, idesc(128, 8, 1, 0), 8, 6, 6});
    run("SW32: A K sw32, B MN (LBO=x, SBO=256 K-grp)", a_k32, b_mn32, P{16, 256, 4096, 256, 4096, 256, idesc(128, 8, 0, 1), 8, 6, 6});
    run("SW32: A K sw32, B MN (LBO=256, SBO=x)", a_k32, b_mn32, P{16, 256, 256, 4096, 4096, 256, idesc(128, 8, 0, 1), 8, 6, 6});
    run("SW32: A MN (LBO=4096, SBO=256), B MN (SBO=256)", a_mn32, b_mn32, P{4096, 256, 4096, 256, 256, 256, idesc(128, 8, 1, 1), 8, 6, 6});
    // SW128, 32-float blocks
    Off a_k128 = [sw](int m, int k) { return sw((k / 32) * 16384 + m * 128 + (k % 32) * 4, 3); };
    Off a_mn128 = [sw](int m, int k) { return sw((m / 32) * 16384 + k * 128 + (m % 32) * 4, 3); };       // K <= 128 rows of 128 B
    Off b_k128 = [sw](int n, int k) { return sw((k / 32) * 1024 + n * 128 + (k % 32) * 4, 3); };
    run("SW128: A K (SBO=1024), B K (SBO=1024), step 32 B", a_k128, b_k128, P{16, 1024, 16, 1024, 32, 32, idesc(128, 8, 0, 0), 8, 2, 2});
    run("SW128: A MN (LBO=16384, SBO=1024), B K sw32", a_mn128, b_k32, P{16384, 1024, 16, 256, 1024, 256, idesc(128, 8, 1, 0), 8, 2, 6});
    run("SW128: A MN (LBO=1024, SBO=16384), B K sw32", a_mn128, b_k32, P{1024, 16384, 16, 256, 1024, 256, idesc(128, 8, 1, 0), 8, 2, 6});
    return 0;
}
