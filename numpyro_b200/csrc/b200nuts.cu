// C ABI of the B200 NUTS/HMC engine (declared in include/b200nuts.h) + the small kernels:
// chain begin/resume, regime R1 (one warp per chain, whole run in one launch), parity hooks.
// The streaming engine (regime R2) lives in stream_engine.cuh.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false --extended-lambda -lineinfo
// (-fmad=false: the tree/adaptation bookkeeping follows the det-f32 convention; hot loops issue
// their FMAs explicitly).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <mutex>
#include <chrono>
#include <thread>
#include <unistd.h>
#include <cmath>
#include "engine.cuh"
#include "stream_engine.cuh"
#include "gemm_engine.h"
#include "predict.cuh"

using namespace b2;

// ------------------------------------------------------------------------------------------------
struct B200Nuts {
    B200NutsConfig cfg; FamilySpec fam; SiteLayout sites; TickCfg tick;
    int C = 0, D = 0, Dp = 0, regime = 0, device = 0, num_sms = 0;
    bool inited = false;
    ChainCtl* ctl = nullptr; float* vecs = nullptr; float* gtmp = nullptr; float* scratch = nullptr;
    uint32_t* keys = nullptr;
    float* dense = nullptr;                    // dense_mass: [C][4][D][D] (ChainVecs::dense)
    int32_t* ecs_idx = nullptr; bool ecs_proxy_set = false, ecs_idx_set = false;     // HMCECS: [C][m] subsample rows
    int32_t* cond_map = nullptr; float* cond_val = nullptr; float* cond_scratch = nullptr; bool cond_set = false;   // HMCGibbs
    bool imm_given = false;                    // b200nuts_set_inverse_mass_matrix was called
    // R2
    float2* partial = nullptr; uint4* beta = nullptr; StreamSync* sync = nullptr;
    float* img = nullptr; long long n_tiles = 0; int pad_rows = 0, ks = 0;   // engine-owned tile image of (X, y)
    int grid = 0, stages = 4, vecs_in_smem = 0, num_groups = 1; size_t smem = 0;
    long long launches = 0, passes = 0;
    unsigned long long dbg[16] = {0};
    unsigned int* trace_host = nullptr; unsigned int* trace_dev = nullptr;     // B200NUTS_TRACE: host-mapped progress words
    // row sharding
    int shard_rank = 0, shard_count = 1; bool shards_connected = false; unsigned int epoch = 0;
    float2* mail = nullptr; float2* mail_peer[kMaxShards] = {nullptr}; bool mail_ipc[kMaxShards] = {false};
    long long n_rows_global = 0; float nll_local_const = 0.0f;
    // R3
    GemmRegime* gemm = nullptr;
    // enqueue-only launches: the launch whose outcome b200nuts_sync has not collected yet
    bool pending = false; cudaStream_t pending_stream = nullptr;
    int cur_iter = -1;                         // HMCState.i of every chain when known (b200nuts_transition), else -1
    std::string err;
    std::mutex mu;
};

static std::string g_create_err;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return B200NUTS_ECUDA;                                                                 \
        }                                                                                          \
    } while (0)

// ------------------------------------------------------------------------------------------------
// small kernels
B2_D float* dense_of(float* dense, int chain, int D) { return dense ? dense + (size_t)chain * 4 * D * D : nullptr; }

__global__ void k_chain_begin(TickCfg cfg, ChainCtl* ctl, float* vecs, float* dense, int C, int Dp, const uint32_t* keys,
                              const float* z0) {
    const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chain >= C) return;
    ChainCtl c; memset(&c, 0, sizeof(c));
    ChainVecs cv; cv.base = vecs + (size_t)chain * Dp; cv.field_stride = C * Dp; cv.dense = dense_of(dense, chain, cfg.D);
    OutBufs none; memset(&none, 0, sizeof(none));
    Tick t{cfg, c, cv, none, chain, C};
    Key k; k.a = keys[2 * chain]; k.b = keys[2 * chain + 1];
    t.begin(k, z0 ? z0 + (size_t)chain * cfg.D : nullptr);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) ctl[chain] = c;
}

__global__ void k_chain_resume(TickCfg cfg, ChainCtl* ctl, float* vecs, float* dense, int C, int Dp) {
    const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chain >= C) return;
    ChainCtl c = ctl[chain];
    __syncwarp();
    if (c.phase != PH_DONE || c.init_failed || c.i >= cfg.total_iters) return;
    ChainVecs cv; cv.base = vecs + (size_t)chain * Dp; cv.field_stride = C * Dp; cv.dense = dense_of(dense, chain, cfg.D);
    OutBufs none; memset(&none, 0, sizeof(none));
    Tick t{cfg, c, cv, none, chain, C};
    t.begin_transition();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) ctl[chain] = c;
}

// Regime R1: each warp runs its chain to completion; the potential is evaluated inside the warp.
__global__ void __launch_bounds__(128) k_warp_run(TickCfg cfg, FamilySpec fam, OutBufs out, ChainCtl* ctl, float* vecs, float* dense,
                                                  float* gtmp, float* scratch, long long scratch_stride, int C, int Dp) {
    const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chain >= C) return;
    ChainCtl c = ctl[chain];
    ChainVecs cv; cv.base = vecs + (size_t)chain * Dp; cv.field_stride = C * Dp; cv.dense = dense_of(dense, chain, cfg.D);
    Tick t{cfg, c, cv, out, chain, C};
    float* g = gtmp + (size_t)chain * Dp;
    float* scr = scratch ? scratch + (size_t)chain * scratch_stride : nullptr;
    if (fam.ecs_m > 0) fam.ecs_idx += (size_t)chain * fam.ecs_m;
    if (fam.cond_Dfree > 0) { fam.cond_val += (size_t)chain * (fam.D + 1); fam.cond_scratch += (size_t)chain * 2 * fam.D; }
    while (c.phase != PH_DONE) {
        __syncwarp();
        float u;
        potential_inwarp(fam, cv.v(V_ZS), scr, u, g);
        __syncwarp();
        t.advance(u, g);
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) ctl[chain] = c;
}

__global__ void k_potential_warp(FamilySpec fam, const float* z, float* U, float* g, float* scratch,
                                 long long scratch_stride, int C) {
    const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chain >= C) return;
    if (fam.ecs_m > 0) fam.ecs_idx += (size_t)chain * fam.ecs_m;
    const int Dz = fam.cond_Dfree > 0 ? fam.cond_Dfree : fam.D;
    if (fam.cond_Dfree > 0) { fam.cond_val += (size_t)chain * (fam.D + 1); fam.cond_scratch += (size_t)chain * 2 * fam.D; }
    float u;
    potential_inwarp(fam, z + (size_t)chain * Dz, scratch ? scratch + (size_t)chain * scratch_stride : nullptr, u,
                     g + (size_t)chain * Dz);
    if ((threadIdx.x & 31) == 0) U[chain] = u;
}

// dense handles: (re)derive M^1/2 and the Cholesky workspace from the M^-1 block of every chain (set_dense_state)
__global__ void k_dense_roots(TickCfg cfg, ChainCtl* ctl, float* vecs, float* dense, int C, int Dp) {
    const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chain >= C) return;
    ChainCtl c = ctl[chain];
    ChainVecs cv; cv.base = vecs + (size_t)chain * Dp; cv.field_stride = C * Dp; cv.dense = dense_of(dense, chain, cfg.D);
    OutBufs none; memset(&none, 0, sizeof(none));
    Tick t{cfg, c, cv, none, chain, C};
    t.dense_roots();
}

// velocity_verlet halves for the leapfrog parity hook (numpyro/infer/hmc_util.py:289-309)
__global__ void k_leap_pre(int C, int D, const float* eps, const float* imm, float* z, float* r, const float* g) {
    const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chain >= C) return;
    const size_t o = (size_t)chain * D;
    leap_begin(D, eps[chain], imm + o, z + o, r + o, g + o, z + o, r + o);
}
__global__ void k_leap_post(int C, int D, const float* eps, float* r, const float* g) {
    const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chain >= C) return;
    const size_t o = (size_t)chain * D;
    const float half = 0.5f * eps[chain];
    B2_FOR_D(d, D) r[o + d] = r[o + d] - half * g[o + d];
}

// postprocess_fn: constrained latent sites (flat order) followed by the deterministic block
__global__ void k_constrain(FamilySpec f, const float* z, long long n, float* out, int Dc) {
    const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* zr = z + row * f.D; float* o = out + row * Dc;
    if (f.family == FAM_DIAG_GAUSSIAN) { B2_FOR_D(d, f.D) o[d] = zr[d]; return; }
    if (f.family == FAM_EIGHT_SCHOOLS) {
        const int J = f.D - 2;
        const float tau = expf(zr[1]);
        B2_FOR_D(d, f.D) o[d] = (d == 1) ? tau : zr[d];
        B2_FOR_D(j, J) o[f.D + j] = zr[0] + tau * zr[2 + j];        // theta = mu + tau * theta_base
        return;
    }
    B2_FOR_D(d, f.D) {
        const bool pos = (f.off_lambda >= 0 && d >= f.off_lambda && d < f.off_lambda + f.Dx) ||
                         (d == f.off_tau) || (d == f.off_prec);
        o[d] = pos ? expf(zr[d]) : zr[d];
    }
    if (Dc > f.D) { B2_FOR_D(j, f.Dx) o[f.D + j] = glm_scale_at(f, zr, j) * zr[f.off_u + j]; }   // betas
}

// HMCState fields of every chain -> caller's device buffers (the outputs of the XLA-FFI transition call)
__global__ void k_state_export(const ChainCtl* ctl, const float* vecs, int C, int D, int Dp, float* z, float* z_grad, float* scalars) {
    const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chain >= C) return;
    const float* vz = vecs + ((size_t)V_Z * C + chain) * Dp; const float* vg = vecs + ((size_t)V_G * C + chain) * Dp;
    B2_FOR_D(d, D) { if (z) z[(size_t)chain * D + d] = vz[d]; if (z_grad) z_grad[(size_t)chain * D + d] = vg[d]; }
    if (scalars && (threadIdx.x & 31) == 0) {
        const ChainCtl& c = ctl[chain]; float* o = scalars + (size_t)chain * B200NUTS_STATE_SCALARS;
        o[0] = (float)c.i; o[1] = c.pe; o[2] = c.energy; o[3] = (float)c.num_steps; o[4] = c.accept_prob; o[5] = c.mean_accept_prob;
        o[6] = (float)c.diverging; o[7] = c.step_size;
    }
}

__global__ void k_lgamma_sum(const float* y, long long n, double* out) {
    double a = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        a += lgamma((double)y[i] + 1.0);
    for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, a);
}

__global__ void k_prng_split(const uint32_t* keys, long long n_keys, int num, uint32_t* out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_keys * num) return;
    const long long kidx = i / num; const uint32_t j = (uint32_t)(i % num);
    Key k; k.a = keys[2 * kidx]; k.b = keys[2 * kidx + 1];
    const Key o = split_at(k, j);
    out[2 * i] = o.a; out[2 * i + 1] = o.b;
}
__global__ void k_prng_draw(int kind, uint32_t ka, uint32_t kb, long long n, float lo, float hi, void* out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Key k; k.a = ka; k.b = kb;
    if (kind == 0) ((uint32_t*)out)[i] = bits_at(k, (uint32_t)i);
    else if (kind == 1) ((float*)out)[i] = uniform_at(k, (uint32_t)i, lo, hi);
    else ((float*)out)[i] = normal_at(k, (uint32_t)i);
}
__global__ void k_detmath(int op, const float* x, long long n, float* out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    out[i] = op == 0 ? d_exp(v) : op == 1 ? d_log(v) : op == 2 ? d_log1p(v) : op == 3 ? d_expit(v) : d_erfinv(v);
}

// ------------------------------------------------------------------------------------------------
#ifdef B2_SPLIT_BUILD
// the instances live in stream_instances.cu, one object per tile width
const void* b2_stream_kernel_ks1(int lik, bool mg); const void* b2_stream_kernel_ks2(int lik, bool mg);
const void* b2_stream_kernel_ks4(int lik, bool mg); const void* b2_stream_kernel_ks7(int lik, bool mg);
const void* b2_stream_kernel_ks8(int lik, bool mg);
static const void* stream_kernel_for(int ks, int lik, int num_groups) {
    const bool mg = num_groups > 1;
    switch (ks) {
    case 1: return b2_stream_kernel_ks1(lik, mg);
    case 2: return b2_stream_kernel_ks2(lik, mg);
    case 4: return b2_stream_kernel_ks4(lik, mg);
    case 7: return b2_stream_kernel_ks7(lik, mg);
    case 8: return b2_stream_kernel_ks8(lik, mg);
    default: return nullptr;
    }
}
#else
template <int KS, bool MG>
static const void* stream_kernel_lik(int lik) {
    return lik == LIK_BERNOULLI ? (const void*)stream_engine_kernel<KS, LIK_BERNOULLI, MG>
         : lik == LIK_POISSON   ? (const void*)stream_engine_kernel<KS, LIK_POISSON, MG>
                                : (const void*)stream_engine_kernel<KS, LIK_NORMAL, MG>;
}
template <bool MG>
static const void* stream_kernel_for_mg(int ks, int lik) {
#ifdef B2_STREAM_FAST_KS           // development builds: one instance only (B200NUTS_FAST_KS=<ks> python -m numpyro_b200.build)
    return (ks == B2_STREAM_FAST_KS && lik == LIK_BERNOULLI) ? (const void*)stream_engine_kernel<B2_STREAM_FAST_KS, LIK_BERNOULLI, MG> : nullptr;
#else
    switch (ks) {
    case 1: return stream_kernel_lik<1, MG>(lik);
    case 2: return stream_kernel_lik<2, MG>(lik);
    case 4: return stream_kernel_lik<4, MG>(lik);
    case 7: return stream_kernel_lik<7, MG>(lik);
    case 8: return stream_kernel_lik<8, MG>(lik);
    default: return nullptr;
    }
#endif
}
// one chain group (<= 8 chains): the group bookkeeping folds away; more: passes rotate over the groups
static const void* stream_kernel_for(int ks, int lik, int num_groups) {
    return num_groups > 1 ? stream_kernel_for_mg<true>(ks, lik) : stream_kernel_for_mg<false>(ks, lik);
}
#endif

static int stream_launch(B200Nuts* h, int mode, const OutBufs& out, const float* z_in, float* u_out, float* g_out,
                         cudaStream_t st, int max_passes = 0) {
    StreamParams p; memset(&p, 0, sizeof(p));
    p.cfg = h->tick; p.cfg.D = h->D; p.fam = h->fam; p.out = out; p.C = h->C; p.Dp = h->Dp; p.mode = mode;
    p.num_groups = h->num_groups; p.max_passes = (h->num_groups == 1) ? max_passes : 0;
    p.ctl = h->ctl; p.vecs = h->vecs; p.partial = h->partial; p.beta = h->beta; p.sync = h->sync;
    p.z_in = z_in; p.u_out = u_out; p.g_out = g_out; p.stages = h->stages;
    p.vecs_in_smem = h->vecs_in_smem; p.spin_limit = 2000000000LL;      // ~1 s: every wait in the engine is bounded
    { const char* e = getenv("B200NUTS_SPIN_LIMIT"); if (e) p.spin_limit = atoll(e); }
    { const char* e = getenv("B200NUTS_NO_PREFETCH"); p.no_prefetch = e ? atoi(e) : 0; }
    { const char* e = getenv("B200NUTS_NO_PEEK"); p.no_peek = e ? atoi(e) : 0; }
    if (getenv("B200NUTS_TRACE") && !h->trace_host) {
        if (cudaHostAlloc((void**)&h->trace_host, sizeof(unsigned int) * 32 * h->grid, cudaHostAllocMapped) == cudaSuccess)
            cudaHostGetDevicePointer((void**)&h->trace_dev, h->trace_host, 0);
    }
    if (h->trace_host) memset(h->trace_host, 0, sizeof(unsigned int) * 32 * h->grid);
    p.trace = h->trace_dev;
    p.shard_rank = h->shard_rank; p.shard_count = h->shard_count;
    if (h->shard_count > 1) {
        if (!h->shards_connected) { h->err = "row-sharded handle: call b200nuts_shard_connect first"; return B200NUTS_ESTATE; }
        h->epoch += 1u; p.epoch = h->epoch & 0xFFu;        // every rank makes the same sequence of launches
        for (int q = 0; q < h->shard_count; ++q) p.mail[q] = h->mail_peer[q];
        p.fam.N = h->n_rows_global;                         // the priors / Normal-likelihood constants see the whole dataset
        p.fam.nll_const = 0.0f; p.nll_local_const = h->nll_local_const;
    }
    p.img = h->img; p.n_tiles = h->n_tiles; p.pad_rows = h->pad_rows;
    { const char* e = getenv("B200NUTS_DEBUG_SWEEP"); p.dbg_sweep = e ? atoi(e) : 0; }
    { const char* e = getenv("B200NUTS_DEBUG_WARPS"); p.dbg_warps = e ? atoi(e) : 1000; }
    CK(cudaMemsetAsync(h->sync, 0, sizeof(StreamSync), st));
    // sequence tags ride inside the exchanged data: a launch must not see the previous launch's tags
    CK(cudaMemsetAsync(h->beta, 0, sizeof(uint4) * kBetaCopies * h->num_groups * kBetaWords, st));
    CK(cudaMemsetAsync(h->partial, 0, sizeof(float2) * (size_t)h->grid * h->num_groups * kStreamCT * kGStride, st));
    void* args[] = {&p};
    const void* fn = stream_kernel_for(h->ks, h->fam.likelihood, h->num_groups);
    if (!fn) { h->err = "stream regime: no kernel instance for this shape"; return B200NUTS_EINVAL; }
    {   // (the dynamic shared-memory limit of the kernel was raised once, in b200nuts_create)
        cudaError_t le = cudaLaunchCooperativeKernel(fn, dim3(h->grid), dim3(kStreamThreads), args, h->smem, st);
        if (le != cudaSuccess) {
            cudaFuncAttributes fa; memset(&fa, 0, sizeof(fa)); cudaFuncGetAttributes(&fa, fn);
            int occ = -1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kStreamThreads, h->smem);
            char buf[256];
            snprintf(buf, sizeof(buf), " [regs %d, maxThreadsPerBlock %d, static smem %zu, dynamic smem %zu, threads %d, blocks/SM %d]",
                     fa.numRegs, fa.maxThreadsPerBlock, fa.sharedSizeBytes, h->smem, kStreamThreads, occ);
            h->err = std::string("cudaLaunchCooperativeKernel: ") + cudaGetErrorString(le) + buf;
            return B200NUTS_ECUDA;
        }
    }
    h->launches += 1;
    return 0;
}

static int check_stream_abort(B200Nuts* h, cudaStream_t st) {
    StreamSync s;
    {   // watchdog: a launch that does not finish is reported (with the progress trace, if enabled), never waited on forever
        double limit_s = 600.0;
        if (const char* e = getenv("B200NUTS_WATCHDOG_S")) limit_s = atof(e);
        const auto t0 = std::chrono::steady_clock::now();
        cudaError_t q;
        while ((q = cudaStreamQuery(st)) == cudaErrorNotReady) {
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit_s) {
                if (h->trace_host) {
                    fprintf(stderr, "[b200nuts] watchdog: progress (pass<<8|stage) per CTA: consumer / tick\n");
                    for (int g = 0; g < h->grid; ++g)
                        fprintf(stderr, "  cta %3d: cons %u.%u  tick lane0 %u.%u lane16 %u.%u lane31 %u.%u  fetch ballot %08x polls %u\n", g, h->trace_host[32 * g] >> 8, h->trace_host[32 * g] & 255u,
                                h->trace_host[32 * g + 1] >> 8, h->trace_host[32 * g + 1] & 255u, h->trace_host[32 * g + 4] >> 8, h->trace_host[32 * g + 4] & 255u,
                                h->trace_host[32 * g + 5] >> 8, h->trace_host[32 * g + 5] & 255u, h->trace_host[32 * g + 2], h->trace_host[32 * g + 3]);
                }
                if (getenv("B200NUTS_TRACE")) {      // post-mortem over a side stream (the stuck kernel keeps the main one busy)
                    cudaStream_t side; cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
                    std::vector<uint4> hb((size_t)kBetaCopies * h->num_groups * kBetaWords);
                    std::vector<float2> hp((size_t)h->grid * h->num_groups * kStreamCT * kGStride);
                    cudaMemcpyAsync(&s, h->sync, sizeof(s), cudaMemcpyDeviceToHost, side);
                    cudaMemcpyAsync(hb.data(), h->beta, hb.size() * sizeof(uint4), cudaMemcpyDeviceToHost, side);
                    cudaMemcpyAsync(hp.data(), h->partial, hp.size() * sizeof(float2), cudaMemcpyDeviceToHost, side);
                    cudaError_t ce = cudaStreamSynchronize(side);
                    fprintf(stderr, "[b200nuts] post-mortem copy: %s; abort_flag %u\n", cudaGetErrorString(ce), s.abort_flag);
                    for (int r = 0; r < kBetaCopies; ++r)
                        for (int c = 0; c < h->C && c < kStreamCT; ++c) {      // (group 0 only)
                            fprintf(stderr, "  beta replica %d chain %d tags (k-step 0, t=0..3):", r, c);
                            for (int t = 0; t < 4; ++t) fprintf(stderr, " %08x/%08x", hb[(size_t)r * h->num_groups * kBetaWords + c * 4 + t].y, hb[(size_t)r * h->num_groups * kBetaWords + c * 4 + t].w);
                            fprintf(stderr, "\n");
                        }
                    for (int c = 0; c < h->C && c < kStreamCT; ++c) {
                        fprintf(stderr, "  partial tags chain %d (word 0) per CTA:", c);
                        for (int g = 0; g < h->grid; ++g) { unsigned int tg; memcpy(&tg, &hp[(((size_t)c * h->grid + g) * ((8 * h->ks + 3) / 3)) * 2 + 1].y, 4); fprintf(stderr, " %u", tg); }
                        fprintf(stderr, "\n");
                    }
                }
                h->err = "stream engine watchdog: the launch did not finish in time"; return B200NUTS_ECUDA;
            }
            std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
        if (q != cudaSuccess) { h->err = std::string("stream engine: ") + cudaGetErrorString(q); return B200NUTS_ECUDA; }
    }
    CK(cudaMemcpyAsync(&s, h->sync, sizeof(s), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    h->passes += (long long)s.passes;
    for (int i = 0; i < 16; ++i) h->dbg[i] = s.dbg[i];
    // [11..13]: positions published ahead of the tick (Tick::peek_next), fall-backs, mismatches -- summed over the chains
    h->dbg[11] = h->dbg[12] = h->dbg[13] = 0ull;
    for (int c = 0; c < h->C && c < kMaxStreamChains; ++c) { h->dbg[11] += s.peek_hit[c]; h->dbg[12] += s.peek_fallback[c]; h->dbg[13] += s.peek_mismatch[c]; }
    if (getenv("B200NUTS_DEBUG_TICK")) {
        for (int c = 0; c < h->C; ++c)
            fprintf(stderr, "[b200nuts] owner %d: passes %llu tick avg %.0f max %llu | per pass: finish %.0f advance %.0f publish %.0f gredsum %.0f | look-ahead hit/miss: leaf %u/%u doubling %u/%u transition %u/%u\n",
                    c, s.passes, s.passes ? (double)s.tick_sum[c] / (double)s.passes : 0.0, s.tick_max[c],
                    (double)s.tick_lap[c][0] / s.passes, (double)s.tick_lap[c][1] / s.passes, (double)s.tick_lap[c][2] / s.passes,
                    (double)s.tick_lap[c][3] / s.passes, s.pre_hit[c][0], s.pre_miss[c][0], s.pre_hit[c][1], s.pre_miss[c][1],
                    s.pre_hit[c][2], s.pre_miss[c][2]);
    }
    if (getenv("B200NUTS_DEBUG_CTA") && s.passes) {
        fprintf(stderr, "[b200nuts] owner 0 per pass: gather rounds %.2f | beta poll %.0f cycles in %.2f rounds\n",
                (double)s.dbg[3] / s.passes, (double)s.dbg[12] / s.passes, (double)s.dbg[13] / s.passes);
        fprintf(stderr, "[b200nuts] owner 0 per pass: hand-over consumers -> tick warp %.0f cycles, tick warp -> consumers %.0f cycles\n",
                (double)s.wake[0] / s.passes, (double)s.wake[1] / s.passes);
        for (int g = 0; g < h->grid && g < 160; ++g)
            fprintf(stderr, "[b200nuts] cta %3d per pass: wait_beta %.0f sweep %.0f reduce %.0f gather %.0f\n", g, (double)s.cta_lap[g][0] / s.passes,
                    (double)s.cta_lap[g][1] / s.passes, (double)s.cta_lap[g][2] / s.passes, (double)s.cta_lap[g][3] / s.passes);
    }
#ifdef B2_TICK_LAPS
    if (getenv("B200NUTS_DEBUG_TICK") && s.passes) {
        fprintf(stderr, "[b200nuts] chain 0 tick laps, cycles per occurrence (occurrences):");
        for (int k = 0; k < 16; ++k) if (s.laps[16 + k]) fprintf(stderr, " [%d] %.0f (%llu)", k, (double)s.laps[k] / (double)s.laps[16 + k], s.laps[16 + k]);
        fprintf(stderr, "\n");
    }
#endif
    if (s.abort_flag) {
        static const char* what[] = {"", "a tile copy never landed", "beta fetch timed out", "partial poll timed out",
                                     "a peer rank's likelihood sums never arrived (row-sharded exchange)"};
        h->err = std::string("stream engine aborted: ") + (s.abort_flag < 5 ? what[s.abort_flag] : "?"); return B200NUTS_ECUDA;
    }
    return 0;
}

static long long scratch_stride(const B200Nuts* h) {
    return h->fam.ecs_m > 0 ? 2ll * h->fam.ecs_m + h->fam.Dx : (long long)h->fam.N + h->fam.Dx;
}

// Collect the outcome of the last enqueued launch (caller holds h->mu).
static int sync_locked(B200Nuts* h) {
    if (!h->pending) return 0;
    h->pending = false;
    cudaStream_t st = h->pending_stream;
    if (h->regime == B200NUTS_REGIME_STREAM) return check_stream_abort(h, st);
    if (h->regime == B200NUTS_REGIME_GEMM) {
        GemmStatus gs; memset(&gs, 0, sizeof(gs));
        const std::string e = gemm_sync(h->gemm, st, &gs, &h->launches);
        h->passes = (long long)gs.passes_total;
        for (int i = 0; i < 8; ++i) h->dbg[i] = gs.dbg[i];
        if (!e.empty()) { h->err = e; return B200NUTS_ECUDA; }
        return 0;
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" {

const char* b200nuts_last_error(const B200Nuts* h) { return h ? h->err.c_str() : g_create_err.c_str(); }
int b200nuts_dim(const B200Nuts* h) { return h ? h->D : B200NUTS_EINVAL; }
int b200nuts_regime(const B200Nuts* h) { return h ? h->regime : B200NUTS_EINVAL; }
int64_t b200nuts_launch_count(const B200Nuts* h) { return h ? h->launches : 0; }
int64_t b200nuts_pass_count(const B200Nuts* h) { return h ? h->passes : 0; }
int b200nuts_gemm_info(const B200Nuts* h, int32_t* out8) {
    if (!h || !out8 || !h->gemm) return B200NUTS_EINVAL;
    int v[8]; gemm_describe(h->gemm, v);
    for (int i = 0; i < 8; ++i) out8[i] = v[i];
    return 0;
}
int b200nuts_debug_clocks(const B200Nuts* h, uint64_t* out16) {
    if (!h || !out16) return B200NUTS_EINVAL;
    for (int i = 0; i < 16; ++i) out16[i] = h->dbg[i];
    return 0;
}

int b200nuts_constrained_dim(const B200Nuts* h) {
    if (!h) return B200NUTS_EINVAL;
    const int Df = h->fam.D;                       // (the full model's dimension, also for a handle conditioned on Gibbs sites)
    if (h->fam.family == FAM_EIGHT_SCHOOLS) return Df + (Df - 2);
    if (h->fam.family == FAM_GLM && (h->fam.off_lambda >= 0 || h->fam.gscale != SCALE_NONE)) return Df + h->fam.Dx;
    return Df;
}

void b200nuts_destroy(B200Nuts* h) {
    if (!h) return;
    cudaFree(h->ctl); cudaFree(h->vecs); cudaFree(h->gtmp); cudaFree(h->scratch); cudaFree(h->keys); cudaFree(h->dense); cudaFree(h->ecs_idx); cudaFree(h->cond_map); cudaFree(h->cond_val); cudaFree(h->cond_scratch);
    cudaFree(h->partial); cudaFree(h->beta); cudaFree(h->sync); cudaFree(h->img);
    if (h->trace_host) cudaFreeHost(h->trace_host);
    for (int q = 0; q < kMaxShards; ++q) if (h->mail_ipc[q] && h->mail_peer[q]) cudaIpcCloseMemHandle(h->mail_peer[q]);
    cudaFree(h->mail);
    gemm_destroy(h->gemm);
    delete h;
}

int b200nuts_create(const B200NutsConfig* cfg, B200Nuts** out) {
    if (!cfg || !out) { g_create_err = "null argument"; return B200NUTS_EINVAL; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_create_err = "no CUDA device: the engine has no CPU fallback"; return B200NUTS_ECUDA;
    }
    B200Nuts* h = new B200Nuts();
    h->cfg = *cfg;
    std::string e = make_family(*cfg, h->fam, h->sites);
    if (e.empty() && (cfg->shard_count < 0 || cfg->shard_count > kMaxShards || (cfg->shard_count > 1 && (cfg->shard_rank < 0 || cfg->shard_rank >= cfg->shard_count))))
        e = "shard_rank / shard_count out of range";
    if (!e.empty()) { g_create_err = e; delete h; return B200NUTS_EINVAL; }
    h->C = cfg->num_chains; h->D = h->fam.D;
    std::vector<int32_t> cond_map_host;
    if (cfg->cond_fixed) {                       // HMCGibbs: drop the Gibbs sites from the chain's vector (whole sites only)
        const int Df = h->fam.D;
        cond_map_host.assign(Df, -1);
        SiteLayout red; memset(&red, 0, sizeof(red));
        // free coordinates keep their (sorted-site) order; the init sites keep their trace order
        int nfree = 0;
        for (int d = 0; d < Df; ++d) if (!cfg->cond_fixed[d]) cond_map_host[d] = nfree++;
        for (int s = 0; s < h->sites.n_sites; ++s) {
            int fixed = 0;
            for (int j = 0; j < h->sites.size[s]; ++j) fixed += cfg->cond_fixed[h->sites.off[s] + j] ? 1 : 0;
            if (fixed == h->sites.size[s]) continue;
            if (fixed != 0) { g_create_err = "cond_fixed must cover whole sites"; delete h; return B200NUTS_EINVAL; }
            red.off[red.n_sites] = cond_map_host[h->sites.off[s]]; red.size[red.n_sites] = h->sites.size[s]; red.n_sites += 1;
        }
        if (nfree == 0 || nfree == Df) { g_create_err = "cond_fixed: at least one free and one fixed coordinate"; delete h; return B200NUTS_EINVAL; }
        if (cfg->shard_count > 1 || cfg->ecs_subsample_size > 0) { g_create_err = "conditioning does not combine with row sharding / subsampling"; delete h; return B200NUTS_EINVAL; }
        h->sites = red; h->fam.cond_Dfree = nfree; h->D = nfree;
    }
    h->Dp = (h->D + 3) & ~3;
    make_tick_cfg(*cfg, h->fam, h->sites, 0, false, h->tick);
    cudaGetDevice(&h->device);
    cudaDeviceSetLimit(cudaLimitStackSize, 4096);      // per-chain state machine frames (tick.cuh)
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device);
    // regime
    int regime = cfg->regime;
    const bool glm = h->fam.family == FAM_GLM;
    const bool stream_ok = glm && h->C <= kMaxStreamChains && h->fam.Dx <= 64 && h->C <= h->num_sms;
    const bool gemm_ok = glm && h->fam.Dx <= 1024 && h->fam.N * (long long)h->fam.Dx >= 4096;
    if (regime == B200NUTS_REGIME_AUTO) {
        // many chains: the gradient is a GEMM (tcgen05); tall data with a handful of chains: one HBM sweep per pass; else in-warp
        if (gemm_ok && h->C >= 128 && (long long)h->fam.N * h->fam.Dx >= (1LL << 17)) regime = B200NUTS_REGIME_GEMM;
        else if (stream_ok && (long long)h->fam.N * h->fam.Dx >= (1LL << 20)) regime = cfg->dense_mass ? B200NUTS_REGIME_GEMM : B200NUTS_REGIME_STREAM;
        else regime = B200NUTS_REGIME_WARP;
    }
    if (h->fam.cond_Dfree > 0) {
        if (regime != B200NUTS_REGIME_WARP && cfg->regime != B200NUTS_REGIME_AUTO) {
            g_create_err = "a handle conditioned on Gibbs sites runs in the warp regime"; delete h; return B200NUTS_EINVAL;
        }
        regime = B200NUTS_REGIME_WARP;
    }
    if (h->fam.ecs_m > 0) {                       // HMCECS inner potential: O(m) rows per gradient, evaluated inside the chain's warp
        if (regime != B200NUTS_REGIME_WARP && cfg->regime != B200NUTS_REGIME_AUTO) {
            g_create_err = "energy-conserving subsampling runs in the warp regime"; delete h; return B200NUTS_EINVAL;
        }
        if (cfg->shard_count > 1) { g_create_err = "energy-conserving subsampling does not combine with row sharding"; delete h; return B200NUTS_EINVAL; }
        regime = B200NUTS_REGIME_WARP;
    }
    if (regime == B200NUTS_REGIME_STREAM && cfg->dense_mass) {
        g_create_err = "dense_mass is not available in the streaming regime (use the warp or the gemm regime)"; delete h; return B200NUTS_EINVAL;
    }
    if (regime == B200NUTS_REGIME_STREAM && !stream_ok) {
        g_create_err = "stream regime needs a GLM family with <= 64 columns and one SM per chain (<= 148 chains)"; delete h; return B200NUTS_EINVAL;
    }
    if (regime == B200NUTS_REGIME_GEMM && !gemm_ok) {
        g_create_err = "gemm regime needs a GLM family with <= 1024 columns"; delete h; return B200NUTS_EINVAL;
    }
    if (cfg->shard_count > 1) {
        // rows split over ranks: the streaming regime when it applies (a handful of chains, <= 64 columns), else the GEMM regime
        if (cfg->regime == B200NUTS_REGIME_GEMM || !stream_ok || (cfg->regime == B200NUTS_REGIME_AUTO && gemm_ok && h->C > 32)) {
            if (!gemm_ok || cfg->regime == B200NUTS_REGIME_STREAM || cfg->regime == B200NUTS_REGIME_WARP) {
                g_create_err = "row-sharded handles need the streaming regime (GLM, <= 64 columns, one SM per chain) or the gemm regime (GLM, <= 1024 columns)"; delete h; return B200NUTS_EINVAL;
            }
            regime = B200NUTS_REGIME_GEMM;
        } else regime = B200NUTS_REGIME_STREAM;
        h->shard_rank = cfg->shard_rank; h->shard_count = cfg->shard_count;
        h->n_rows_global = cfg->n_rows_global > 0 ? cfg->n_rows_global : cfg->n_rows;
    }
    h->regime = regime;
    auto fail = [&](const char* what, cudaError_t ce) {
        g_create_err = std::string(what) + ": " + cudaGetErrorString(ce); b200nuts_destroy(h); return B200NUTS_ECUDA;
    };
    cudaError_t ce;
    if ((ce = cudaMalloc(&h->ctl, sizeof(ChainCtl) * h->C)) != cudaSuccess) return fail("cudaMalloc ctl", ce);
    if ((ce = cudaMalloc(&h->vecs, sizeof(float) * (size_t)V_COUNT * h->C * h->Dp)) != cudaSuccess) return fail("cudaMalloc vecs", ce);
    if ((ce = cudaMalloc(&h->gtmp, sizeof(float) * (size_t)h->C * h->Dp)) != cudaSuccess) return fail("cudaMalloc gtmp", ce);
    if ((ce = cudaMalloc(&h->keys, sizeof(uint32_t) * 2 * h->C)) != cudaSuccess) return fail("cudaMalloc keys", ce);
    cudaMemset(h->vecs, 0, sizeof(float) * (size_t)V_COUNT * h->C * h->Dp);
    cudaMemset(h->ctl, 0, sizeof(ChainCtl) * h->C);
    if (h->fam.cond_Dfree > 0) {
        const int Df = h->fam.D;
        if ((ce = cudaMalloc(&h->cond_map, sizeof(int32_t) * Df)) != cudaSuccess) return fail("cudaMalloc cond_map", ce);
        if ((ce = cudaMalloc(&h->cond_val, sizeof(float) * (size_t)h->C * (Df + 1))) != cudaSuccess) return fail("cudaMalloc cond_val", ce);
        if ((ce = cudaMalloc(&h->cond_scratch, sizeof(float) * (size_t)h->C * 2 * Df)) != cudaSuccess) return fail("cudaMalloc cond_scratch", ce);
        cudaMemcpy(h->cond_map, cond_map_host.data(), sizeof(int32_t) * Df, cudaMemcpyHostToDevice);
        cudaMemset(h->cond_val, 0, sizeof(float) * (size_t)h->C * (Df + 1));
        h->fam.cond_map = h->cond_map; h->fam.cond_val = h->cond_val; h->fam.cond_scratch = h->cond_scratch;
    }
    if (cfg->dense_mass) {
        const size_t bytes = sizeof(float) * (size_t)h->C * 4 * h->D * h->D;
        if (bytes > ((size_t)8 << 30)) { g_create_err = "dense_mass: num_chains x 4 x D^2 floats exceed 8 GiB"; b200nuts_destroy(h); return B200NUTS_EINVAL; }
        if ((ce = cudaMalloc(&h->dense, bytes)) != cudaSuccess) return fail("cudaMalloc dense mass matrices", ce);
        cudaMemset(h->dense, 0, bytes);
    }
    if (h->fam.ecs_m > 0) {
        const size_t stride = (size_t)h->fam.N + h->fam.Dx;      // (>= 2 m + Dx whenever m <= N / 2; checked here)
        if ((size_t)2 * h->fam.ecs_m > (size_t)h->fam.N) { g_create_err = "ecs_subsample_size must be <= n_rows / 2"; b200nuts_destroy(h); return B200NUTS_EINVAL; }
        if ((ce = cudaMalloc(&h->scratch, sizeof(float) * (2 * (size_t)h->fam.ecs_m + h->fam.Dx) * h->C)) != cudaSuccess) return fail("cudaMalloc scratch", ce);
        if ((ce = cudaMalloc(&h->ecs_idx, sizeof(int32_t) * (size_t)h->fam.ecs_m * h->C)) != cudaSuccess) return fail("cudaMalloc subsample indices", ce);
        cudaMemset(h->ecs_idx, 0, sizeof(int32_t) * (size_t)h->fam.ecs_m * h->C);
        h->fam.ecs_idx = h->ecs_idx;
        (void)stride;
    } else if (glm && regime == B200NUTS_REGIME_WARP) {
        const size_t stride = (size_t)h->fam.N + h->fam.Dx;
        if ((ce = cudaMalloc(&h->scratch, sizeof(float) * stride * h->C)) != cudaSuccess) return fail("cudaMalloc scratch", ce);
    }
    if (glm && h->fam.likelihood == LIK_POISSON) {
        double* d_acc = nullptr; double acc = 0.0;
        if ((ce = cudaMalloc(&d_acc, 8)) != cudaSuccess) return fail("cudaMalloc", ce);
        cudaMemset(d_acc, 0, 8);
        k_lgamma_sum<<<296, 256>>>(h->fam.y, h->fam.N, d_acc);
        ce = cudaMemcpy(&acc, d_acc, 8, cudaMemcpyDeviceToHost);
        cudaFree(d_acc);
        if (ce != cudaSuccess) return fail("lgamma sum", ce);
        h->fam.nll_const = (float)acc;
        h->launches += 1;
    }
    if (regime == B200NUTS_REGIME_STREAM) {
        h->grid = h->num_sms;
        // testing aid: a smaller grid lets two row-sharded handles run side by side on ONE device (tests/test_gpu_rowshard.py)
        if (const char* e = getenv("B200NUTS_GRID")) { const int g = atoi(e); if (g >= h->C && g >= 1 && g < h->grid) h->grid = g; }
        int max_smem = 0;
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
        h->ks = stream_ks_for(h->fam.Dx);
        h->stages = 0;
        const bool mg = h->C > kStreamCT;
        for (int vs = 1; vs >= 0 && !h->stages; --vs)
            for (int stg = kMaxStages; stg >= 2; --stg)      // slots per consumer warp: one in use, the others in flight
                if (stream_smem_bytes(h->ks, h->Dp, stg, vs != 0, mg) <= (size_t)max_smem) {
                    h->stages = stg; h->vecs_in_smem = vs; break;
                }
        if (!h->stages) { g_create_err = "stream regime: shared memory budget exceeded"; b200nuts_destroy(h); return B200NUTS_EINVAL; }
        // experiments: force a ring depth / where the chain vectors live (the combination must fit)
        if (const char* e = getenv("B200NUTS_FORCE_STAGES")) {
            const int stg = atoi(e); const char* v = getenv("B200NUTS_FORCE_VECS_SMEM"); const int vs = v ? atoi(v) : h->vecs_in_smem;
            if (stg >= 2 && stg <= kMaxStages && stream_smem_bytes(h->ks, h->Dp, stg, vs != 0, mg) <= (size_t)max_smem) { h->stages = stg; h->vecs_in_smem = vs; }
        }
        h->smem = stream_smem_bytes(h->ks, h->Dp, h->stages, h->vecs_in_smem != 0, mg);
        // engine-owned tile image of (X, y): one contiguous block per 32 rows (stream_engine.cuh)
        h->n_tiles = (h->fam.N + kTileRows - 1) / kTileRows;
        h->pad_rows = (int)(h->n_tiles * kTileRows - h->fam.N);
        const size_t img_floats = (size_t)h->n_tiles * stream_tile_floats(h->ks);
        if ((ce = cudaMalloc(&h->img, sizeof(float) * img_floats)) != cudaSuccess) return fail("cudaMalloc tile image", ce);
        k_stream_repack<<<h->num_sms * 8, 256>>>(h->fam.X, h->fam.y, h->fam.N, h->fam.Dx, stream_pitch(h->ks), h->n_tiles, h->img);
        if ((ce = cudaGetLastError()) != cudaSuccess) return fail("repack", ce);
        h->launches += 1;
        h->num_groups = (h->C + kStreamCT - 1) / kStreamCT;
        if ((ce = cudaMalloc(&h->partial, sizeof(float2) * (size_t)h->grid * h->num_groups * kStreamCT * kGStride)) != cudaSuccess) return fail("cudaMalloc partial", ce);
        if ((ce = cudaMalloc(&h->beta, sizeof(uint4) * kBetaCopies * h->num_groups * kBetaWords)) != cudaSuccess) return fail("cudaMalloc beta", ce);
        if ((ce = cudaMalloc(&h->sync, sizeof(StreamSync))) != cudaSuccess) return fail("cudaMalloc sync", ce);
        if (h->shard_count > 1) {
            if ((ce = cudaMalloc(&h->mail, sizeof(float2) * kMailFloat2)) != cudaSuccess) return fail("cudaMalloc mailbox", ce);
            cudaMemset(h->mail, 0, sizeof(float2) * kMailFloat2);
            h->nll_local_const = h->fam.nll_const;
        }
    }
    if (regime == B200NUTS_REGIME_GEMM) {
        const std::string ge = gemm_create(&h->gemm, h->fam, h->C, h->Dp, h->num_sms, h->ctl, h->vecs, h->dense, &h->launches);
        if (!ge.empty()) { g_create_err = ge; b200nuts_destroy(h); return B200NUTS_ECUDA; }
        if (h->shard_count > 1) {
            const size_t mb = gemm_mail_bytes(h->gemm);
            if ((ce = cudaMalloc(&h->mail, mb)) != cudaSuccess) return fail("cudaMalloc mailbox", ce);
            cudaMemset(h->mail, 0, mb);
            h->nll_local_const = h->fam.nll_const;
        }
    }
    {   // Load every kernel this handle can launch NOW.  With lazy module loading the first launch of a function may need
        // a context-wide synchronisation; a persistent kernel of another (row-sharded) handle that is waiting for this
        // handle's contribution would then never finish.
        cudaFuncAttributes fa;
        const void* fns[] = {(const void*)k_chain_begin, (const void*)k_chain_resume, (const void*)k_warp_run, (const void*)k_potential_warp, (const void*)k_dense_roots,
                             (const void*)k_leap_pre, (const void*)k_leap_post, (const void*)k_constrain,
                             regime == B200NUTS_REGIME_STREAM ? stream_kernel_for(h->ks, h->fam.likelihood, h->num_groups) : nullptr};
        for (const void* fn : fns)
            if (fn && (ce = cudaFuncGetAttributes(&fa, fn)) != cudaSuccess) return fail("cudaFuncGetAttributes", ce);
        if (regime == B200NUTS_REGIME_STREAM) {
            const void* fn = stream_kernel_for(h->ks, h->fam.likelihood, h->num_groups);
            if (!fn) { g_create_err = "stream regime: no kernel instance for this shape"; b200nuts_destroy(h); return B200NUTS_EINVAL; }
            if ((ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem)) != cudaSuccess) return fail("cudaFuncSetAttribute", ce);
        }
    }
    if ((ce = cudaDeviceSynchronize()) != cudaSuccess) return fail("create", ce);
    *out = h;
    return 0;
}

int b200nuts_init(B200Nuts* h, const uint32_t* keys, const float* z0, int32_t num_warmup, void* stream) {
    if (!h || !keys || num_warmup < 0) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    if (int rc = sync_locked(h)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    std::string e = make_tick_cfg(h->cfg, h->fam, h->sites, num_warmup, z0 != nullptr, h->tick);
    if (!e.empty()) { h->err = e; return B200NUTS_EINVAL; }
    CK(cudaMemcpyAsync(h->keys, keys, sizeof(uint32_t) * 2 * h->C, cudaMemcpyHostToDevice, st));
    const int blocks = (h->C + 3) / 4;
    h->tick.imm_given = h->imm_given ? 1 : 0;
    k_chain_begin<<<blocks, 128, 0, st>>>(h->tick, h->ctl, h->vecs, h->dense, h->C, h->Dp, h->keys, z0);
    CK(cudaGetLastError());
    h->launches += 1;
    h->inited = true;
    h->cur_iter = 0;
    return 0;
}

// Enqueue only (an XLA-FFI handler must not block); b200nuts_sync collects the outcome.
static int run_locked(B200Nuts* h, const B200NutsRun* run, cudaStream_t st) {
    if (!h->inited) { h->err = "b200nuts_run called before b200nuts_init"; return B200NUTS_ESTATE; }
    if (int rc = sync_locked(h)) return rc;            // at most one launch of a handle is in flight
    h->tick.total_iters = run->upper;
    h->tick.collect_start = run->collect_start; h->tick.thinning = run->thinning; h->tick.S = run->collection_size;
    OutBufs out;
    out.z = run->z; out.diverging = run->diverging; out.num_steps = run->num_steps; out.accept_prob = run->accept_prob;
    out.mean_accept_prob = run->mean_accept_prob; out.pe = run->potential_energy; out.energy = run->energy;
    out.step_size = run->step_size;
    if (run->max_passes < 0 || (run->max_passes > 0 && (h->regime == B200NUTS_REGIME_WARP || (h->regime == B200NUTS_REGIME_STREAM && h->num_groups != 1)))) {
        h->err = "max_passes needs the streaming regime with <= 8 chains or the gemm regime"; return B200NUTS_EINVAL;
    }
    const int blocks = (h->C + 3) / 4;
    k_chain_resume<<<blocks, 128, 0, st>>>(h->tick, h->ctl, h->vecs, h->dense, h->C, h->Dp);
    CK(cudaGetLastError());
    h->launches += 1;
    if (h->regime == B200NUTS_REGIME_WARP) {
        if (h->fam.ecs_m > 0 && !(h->ecs_idx_set && (h->ecs_proxy_set || h->fam.ecs_degree == 0))) {
            h->err = "HMCECS handle: call b200nuts_ecs_set_proxy / b200nuts_ecs_set_indices first"; return B200NUTS_ESTATE;
        }
        k_warp_run<<<blocks, 128, 0, st>>>(h->tick, h->fam, out, h->ctl, h->vecs, h->dense, h->gtmp, h->scratch,
                                           scratch_stride(h), h->C, h->Dp);
        CK(cudaGetLastError());
        h->launches += 1;
    } else if (h->regime == B200NUTS_REGIME_GEMM) {
        if (h->shard_count > 1 && !h->shards_connected) { h->err = "row-sharded handle: call b200nuts_shard_connect first"; return B200NUTS_ESTATE; }
        const std::string e = gemm_run(h->gemm, h->tick, out, run->max_passes, st);
        if (!e.empty()) { h->err = e; return B200NUTS_ECUDA; }
        h->launches += 2;
    } else {
        int rc = stream_launch(h, 0, out, nullptr, nullptr, nullptr, st, run->max_passes);
        if (rc) return rc;
    }
    h->pending = true; h->pending_stream = st;
    if (run->max_passes > 0) h->cur_iter = -1;         // chains may have paused anywhere
    else if (h->cur_iter >= 0 && run->upper > h->cur_iter) h->cur_iter = run->upper;
    return 0;
}

int b200nuts_run(B200Nuts* h, const B200NutsRun* run, void* stream) {
    if (!h || !run || run->thinning < 1 || run->upper < 0) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    return run_locked(h, run, (cudaStream_t)stream);
}

// MCMCKernel.sample granularity (mcmc.py:110-124): advance every chain by n_iter transitions, collecting nothing.
int b200nuts_transition(B200Nuts* h, int32_t n_iter, void* stream) {
    if (!h || n_iter < 0) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->cur_iter < 0) { h->err = "b200nuts_transition: the chains' iteration is unknown (pass-bounded run or mixed state); use b200nuts_run"; return B200NUTS_ESTATE; }
    B200NutsRun run; memset(&run, 0, sizeof(run));
    run.upper = h->cur_iter + n_iter; run.collect_start = run.upper; run.thinning = 1; run.collection_size = 0;
    return run_locked(h, &run, (cudaStream_t)stream);
}

// Wait for the last enqueued launch and report its outcome (a timed-out exchange, a failed launch).
int b200nuts_sync(B200Nuts* h) {
    if (!h) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    return sync_locked(h);
}

int b200nuts_state_to_device(B200Nuts* h, float* z, float* z_grad, float* scalars, void* stream) {
    if (!h) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    k_state_export<<<(h->C + 3) / 4, 128, 0, (cudaStream_t)stream>>>(h->ctl, h->vecs, h->C, h->D, h->Dp, z, z_grad, scalars);
    CK(cudaGetLastError());
    h->launches += 1;
    return 0;
}

int b200nuts_get_state(B200Nuts* h, B200NutsChainState* states, float* z, float* z_grad, float* inv_mass,
                       float* mass_sqrt, float* wf_mean, float* wf_m2, void* stream) {
    if (!h || !states) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    if (int rc0 = sync_locked(h)) return rc0;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<ChainCtl> ctl(h->C);
    CK(cudaMemcpyAsync(ctl.data(), h->ctl, sizeof(ChainCtl) * h->C, cudaMemcpyDeviceToHost, st));
    float* dst[6] = {z, z_grad, inv_mass, mass_sqrt, wf_mean, wf_m2};
    const int fld[6] = {V_Z, V_G, V_IMM, V_SQRTM, V_WF_MEAN, V_WF_M2};
    for (int k = 0; k < 6; ++k)
        if (dst[k])
            CK(cudaMemcpy2DAsync(dst[k], sizeof(float) * h->D, h->vecs + (size_t)fld[k] * h->C * h->Dp, sizeof(float) * h->Dp,
                                 sizeof(float) * h->D, h->C, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int rc = 0;
    for (int c = 0; c < h->C; ++c) {
        ctl_to_public(ctl[c], states[c]);
        if (ctl[c].init_failed) rc = B200NUTS_EINIT;
    }
    if (rc) h->err = "Cannot find valid initial parameters. Please check your model again.";
    return rc;
}

int b200nuts_set_state(B200Nuts* h, const B200NutsChainState* states, const float* z, const float* z_grad,
                       const float* inv_mass, const float* wf_mean, const float* wf_m2, int32_t num_warmup,
                       void* stream) {
    if (!h || !states || !z || !z_grad || !inv_mass) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    if (int rc0 = sync_locked(h)) return rc0;
    cudaStream_t st = (cudaStream_t)stream;
    std::string e = make_tick_cfg(h->cfg, h->fam, h->sites, num_warmup, true, h->tick);
    if (!e.empty()) { h->err = e; return B200NUTS_EINVAL; }
    std::vector<ChainCtl> ctl(h->C);
    for (int c = 0; c < h->C; ++c) public_to_ctl(states[c], ctl[c]);
    std::vector<float> sqrtm((size_t)h->C * h->D);
    for (size_t i = 0; i < sqrtm.size(); ++i) {
        if (!(inv_mass[i] > 0.0f) || !std::isfinite(inv_mass[i])) { h->err = "b200nuts_set_state: inverse_mass_matrix must be positive and finite"; return B200NUTS_EINVAL; }
        sqrtm[i] = 1.0f / sqrtf(inv_mass[i]);
    }
    h->cur_iter = states[0].i;
    for (int c = 1; c < h->C; ++c) if (states[c].i != states[0].i) h->cur_iter = -1;
    CK(cudaMemcpyAsync(h->ctl, ctl.data(), sizeof(ChainCtl) * h->C, cudaMemcpyHostToDevice, st));
    const float* src[6] = {z, z_grad, inv_mass, sqrtm.data(), wf_mean, wf_m2};
    const int fld[6] = {V_Z, V_G, V_IMM, V_SQRTM, V_WF_MEAN, V_WF_M2};
    for (int k = 0; k < 6; ++k)
        if (src[k])
            CK(cudaMemcpy2DAsync(h->vecs + (size_t)fld[k] * h->C * h->Dp, sizeof(float) * h->Dp, src[k], sizeof(float) * h->D,
                                 sizeof(float) * h->D, h->C, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    h->inited = true;
    return 0;
}

// ---- HMCGibbs inner potential: values of the Gibbs sites --------------------------------------------------------------
int b200nuts_cond_set_values(B200Nuts* h, const float* values, void* stream) {
    if (!h || !values) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->fam.cond_Dfree <= 0) { h->err = "not a handle conditioned on Gibbs sites"; return B200NUTS_ESTATE; }
    if (int rc0 = sync_locked(h)) return rc0;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(h->cond_val, values, sizeof(float) * (size_t)h->C * (h->fam.D + 1), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    h->cond_set = true;
    return 0;
}
int b200nuts_full_dim(const B200Nuts* h) { return h ? h->fam.D : B200NUTS_EINVAL; }

// ---- HMCECS inner potential ---------------------------------------------------------------
int b200nuts_ecs_set_proxy(B200Nuts* h, const float* ref, const float* eta_ref, const float* G, const float* H, float L0) {
    if (!h) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->fam.ecs_m <= 0 || h->fam.ecs_degree == 0) { h->err = "not an HMCECS handle with a Taylor proxy"; return B200NUTS_ESTATE; }
    if (!ref || !eta_ref || !G || (h->fam.ecs_degree == 2 && !H)) return B200NUTS_EINVAL;
    if (int rc0 = sync_locked(h)) return rc0;
    h->fam.ecs_ref = ref; h->fam.ecs_eta_ref = eta_ref; h->fam.ecs_G = G; h->fam.ecs_H = H; h->fam.ecs_L0 = L0;
    h->ecs_proxy_set = true;
    return 0;
}

int b200nuts_ecs_set_indices(B200Nuts* h, const int32_t* idx, void* stream) {
    if (!h || !idx) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->fam.ecs_m <= 0) { h->err = "not an HMCECS handle"; return B200NUTS_ESTATE; }
    if (int rc0 = sync_locked(h)) return rc0;
    const size_t n = (size_t)h->fam.ecs_m * h->C;
    for (size_t i = 0; i < n; ++i)
        if (idx[i] < 0 || (long long)idx[i] >= h->fam.N) { h->err = "b200nuts_ecs_set_indices: row index out of range"; return B200NUTS_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(h->ecs_idx, idx, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    h->ecs_idx_set = true;
    return 0;
}

// ---- mass matrix structure -------------------------------------------------------------
int b200nuts_set_inverse_mass_matrix(B200Nuts* h, const float* imm, int32_t ndim, void* stream) {
    if (!h || !imm || (ndim != 1 && ndim != 2)) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    if (int rc0 = sync_locked(h)) return rc0;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = h->D;
    if (h->dense) {
        std::vector<float> blk((size_t)h->C * 4 * D * D, 0.0f);
        for (int c = 0; c < h->C; ++c)
            for (int i = 0; i < D; ++i)
                for (int j = 0; j < D; ++j)
                    blk[((size_t)c * 4) * D * D + (size_t)i * D + j] = ndim == 2 ? imm[(size_t)i * D + j] : (i == j ? imm[i] : 0.0f);
        CK(cudaMemcpyAsync(h->dense, blk.data(), sizeof(float) * blk.size(), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
    } else {
        std::vector<float> v((size_t)h->C * D);
        for (int c = 0; c < h->C; ++c)
            for (int d = 0; d < D; ++d) v[(size_t)c * D + d] = ndim == 2 ? imm[(size_t)d * D + d] : imm[d];
        CK(cudaMemcpy2DAsync(h->vecs + (size_t)V_IMM * h->C * h->Dp, sizeof(float) * h->Dp, v.data(), sizeof(float) * D,
                             sizeof(float) * D, h->C, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    h->imm_given = true;
    return 0;
}

int b200nuts_get_dense_state(B200Nuts* h, float* inverse_mass_matrix, float* mass_matrix_sqrt, float* mass_matrix_sqrt_inv,
                             float* wf_m2, void* stream) {
    if (!h) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->dense) { h->err = "not a dense_mass handle"; return B200NUTS_ESTATE; }
    if (int rc0 = sync_locked(h)) return rc0;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = h->D; const size_t DD = (size_t)D * D;
    std::vector<float> blk((size_t)h->C * 4 * DD);
    CK(cudaMemcpyAsync(blk.data(), h->dense, sizeof(float) * blk.size(), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int c = 0; c < h->C; ++c) {
        const float* b = blk.data() + (size_t)c * 4 * DD;
        if (inverse_mass_matrix) memcpy(inverse_mass_matrix + c * DD, b, sizeof(float) * DD);
        if (mass_matrix_sqrt) memcpy(mass_matrix_sqrt + c * DD, b + DD, sizeof(float) * DD);
        if (wf_m2) memcpy(wf_m2 + c * DD, b + 2 * DD, sizeof(float) * DD);
        if (mass_matrix_sqrt_inv)                           // tril_inv[i][j] = Lc[D-1-j][D-1-i] (hmc_util.py:229-231)
            for (int i = 0; i < D; ++i)
                for (int j = 0; j < D; ++j) mass_matrix_sqrt_inv[c * DD + (size_t)i * D + j] = b[3 * DD + (size_t)(D - 1 - j) * D + (D - 1 - i)];
    }
    return 0;
}

int b200nuts_set_dense_state(B200Nuts* h, const float* inverse_mass_matrix, const float* wf_m2, void* stream) {
    if (!h || !inverse_mass_matrix) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->dense) { h->err = "not a dense_mass handle"; return B200NUTS_ESTATE; }
    if (int rc0 = sync_locked(h)) return rc0;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = h->D; const size_t DD = (size_t)D * D;
    for (int c = 0; c < h->C; ++c) {
        float* b = h->dense + (size_t)c * 4 * DD;
        CK(cudaMemcpyAsync(b, inverse_mass_matrix + c * DD, sizeof(float) * DD, cudaMemcpyHostToDevice, st));
        if (wf_m2) CK(cudaMemcpyAsync(b + 2 * DD, wf_m2 + c * DD, sizeof(float) * DD, cudaMemcpyHostToDevice, st));
    }
    k_dense_roots<<<(h->C + 3) / 4, 128, 0, st>>>(h->tick, h->ctl, h->vecs, h->dense, h->C, h->Dp);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    h->launches += 1;
    return 0;
}

// blob = { pid, device, raw device pointer, cudaIpcMemHandle_t }
struct ShardBlob { long long pid; int device; int pad; unsigned long long ptr; cudaIpcMemHandle_t ipc; };
static_assert(sizeof(ShardBlob) <= B200NUTS_SHARD_HANDLE_BYTES, "shard blob too large");

int b200nuts_shard_export(B200Nuts* h, void* blob) {
    if (!h || !blob) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->shard_count <= 1 || !h->mail) { h->err = "not a row-sharded handle"; return B200NUTS_ESTATE; }
    ShardBlob b; memset(&b, 0, sizeof(b));
    b.pid = (long long)getpid(); b.device = h->device; b.ptr = (unsigned long long)(uintptr_t)h->mail;
    CK(cudaSetDevice(h->device));
    CK(cudaIpcGetMemHandle(&b.ipc, h->mail));
    memset(blob, 0, B200NUTS_SHARD_HANDLE_BYTES); memcpy(blob, &b, sizeof(b));
    return 0;
}

int b200nuts_shard_connect(B200Nuts* h, const void* blobs) {
    if (!h || !blobs) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->shard_count <= 1 || !h->mail) { h->err = "not a row-sharded handle"; return B200NUTS_ESTATE; }
    CK(cudaSetDevice(h->device));
    for (int q = 0; q < h->shard_count; ++q) {
        ShardBlob b; memcpy(&b, (const unsigned char*)blobs + (size_t)q * B200NUTS_SHARD_HANDLE_BYTES, sizeof(b));
        if (q == h->shard_rank) { h->mail_peer[q] = h->mail; continue; }
        if (b.pid == (long long)getpid()) {                 // ranks are threads of one process: plain peer access
            if (b.device != h->device) {
                cudaError_t pe = cudaDeviceEnablePeerAccess(b.device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { h->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe); return B200NUTS_ECUDA; }
                cudaGetLastError();
            }
            h->mail_peer[q] = (float2*)(uintptr_t)b.ptr;
        } else {                                            // ranks are processes: CUDA IPC
            void* ptr = nullptr;
            CK(cudaIpcOpenMemHandle(&ptr, b.ipc, cudaIpcMemLazyEnablePeerAccess));
            h->mail_peer[q] = (float2*)ptr; h->mail_ipc[q] = true;
        }
    }
    if (h->regime == B200NUTS_REGIME_GEMM) {
        void* mail[kMaxShards];
        for (int q = 0; q < h->shard_count; ++q) mail[q] = h->mail_peer[q];
        const std::string e = gemm_set_shards(h->gemm, h->shard_rank, h->shard_count, mail, h->n_rows_global, h->nll_local_const);
        if (!e.empty()) { h->err = e; return B200NUTS_ECUDA; }
    }
    h->shards_connected = true;
    return 0;
}

int b200nuts_potential_and_grad(B200Nuts* h, const float* z, float* U, float* g, void* stream) {
    if (!h || !z || !U || !g) return B200NUTS_EINVAL;
    std::lock_guard<std::mutex> lk(h->mu);
    if (int rc0 = sync_locked(h)) return rc0;
    cudaStream_t st = (cudaStream_t)stream;
    if (h->regime == B200NUTS_REGIME_WARP) {
        if (h->fam.ecs_m > 0 && !(h->ecs_idx_set && (h->ecs_proxy_set || h->fam.ecs_degree == 0))) {
            h->err = "HMCECS handle: call b200nuts_ecs_set_proxy / b200nuts_ecs_set_indices first"; return B200NUTS_ESTATE;
        }
        k_potential_warp<<<(h->C + 3) / 4, 128, 0, st>>>(h->fam, z, U, g, h->scratch, scratch_stride(h), h->C);
        CK(cudaGetLastError());
        h->launches += 1;
        return 0;
    }
    if (h->regime == B200NUTS_REGIME_GEMM) {
        if (h->shard_count > 1 && !h->shards_connected) { h->err = "row-sharded handle: call b200nuts_shard_connect first"; return B200NUTS_ESTATE; }
        const std::string e = gemm_potential(h->gemm, z, U, g, st, &h->launches);
        if (!e.empty()) { h->err = e; return B200NUTS_ECUDA; }
        h->pending = true; h->pending_stream = st;
        return sync_locked(h);               // a parity hook: synchronise and report an aborted pass
    }
    OutBufs none; memset(&none, 0, sizeof(none));
    int rc = stream_launch(h, 1, none, z, U, g, st);
    if (rc) return rc;
    return check_stream_abort(h, st);       // a parity hook: synchronise and report a timed-out exchange instead of returning stale sums
}

int b200nuts_leapfrog(B200Nuts* h, const float* eps, const float* inv_mass, float* z, float* r, float* U, float* g,
                      int32_t n_steps, void* stream) {
    if (!h || !eps || !inv_mass || !z || !r || !U || !g || n_steps < 0) return B200NUTS_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = b200nuts_potential_and_grad(h, z, U, g, stream);
    const int blocks = (h->C + 3) / 4;
    for (int s = 0; s < n_steps && !rc; ++s) {
        k_leap_pre<<<blocks, 128, 0, st>>>(h->C, h->D, eps, inv_mass, z, r, g);
        rc = b200nuts_potential_and_grad(h, z, U, g, stream);
        k_leap_post<<<blocks, 128, 0, st>>>(h->C, h->D, eps, r, g);
        h->launches += 2;
    }
    if (!rc) CK(cudaGetLastError());
    return rc;
}

int b200nuts_constrain(B200Nuts* h, const float* z, int64_t n, float* out, void* stream) {
    if (!h || !z || !out || n < 0) return B200NUTS_EINVAL;
    if (n == 0) return 0;
    const int Dc = b200nuts_constrained_dim(h);
    k_constrain<<<(unsigned)((n + 3) / 4), 128, 0, (cudaStream_t)stream>>>(h->fam, z, n, out, Dc);
    CK(cudaGetLastError());
    h->launches += 1;
    return 0;
}

// ---- after the path: log-likelihoods and posterior-predictive draws over collected samples (predict.cuh) -------------
static int rows_launch(B200Nuts* h, int mode, const float* z, const uint32_t* keys, int64_t n, float* out, cudaStream_t st) {
    if (!h || !z || !out || n < 0 || (mode == 1 && !keys)) return B200NUTS_EINVAL;
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->fam.family == FAM_EIGHT_SCHOOLS) {
        const long long tot = n * (h->D - 2);
        if (mode == 0) k_eight_rows<0><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(h->fam, z, keys, n, out);
        else k_eight_rows<1><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(h->fam, z, keys, n, out);
    } else if (h->fam.family == FAM_GLM) {
        if (mode == 1 && h->fam.likelihood == LIK_POISSON) { h->err = "predictive draws for the Poisson likelihood are not implemented"; return B200NUTS_EINVAL; }
        const dim3 grid((unsigned)((h->fam.N + kPrRows - 1) / kPrRows), (unsigned)((n + kPrSamples - 1) / kPrSamples));
        if (grid.y > 65535u) { h->err = "too many samples in one call (<= 524280)"; return B200NUTS_EINVAL; }
        if (mode == 0) k_glm_rows<0><<<grid, kPrRows, 0, st>>>(h->fam, z, keys, n, out);
        else k_glm_rows<1><<<grid, kPrRows, 0, st>>>(h->fam, z, keys, n, out);
    } else { h->err = "log_likelihood / predictive: unsupported family"; return B200NUTS_EINVAL; }
    CK(cudaGetLastError());
    h->launches += 1;
    return 0;
}
int b200nuts_log_likelihood(B200Nuts* h, const float* z, int64_t n, float* out, void* stream) {
    return rows_launch(h, 0, z, nullptr, n, out, (cudaStream_t)stream);
}
int b200nuts_predict(B200Nuts* h, const float* z, const uint32_t* keys, int64_t n, float* out, void* stream) {
    return rows_launch(h, 1, z, keys, n, out, (cudaStream_t)stream);
}
int64_t b200nuts_obs_count(const B200Nuts* h) {
    if (!h) return B200NUTS_EINVAL;
    return h->fam.family == FAM_EIGHT_SCHOOLS ? (int64_t)(h->D - 2) : (h->fam.family == FAM_GLM ? (int64_t)h->fam.N : 0);
}

// ---- PRNG / det-math parity hooks: host in, host out -------------------------------------------
static int hook_io(const void* in, size_t in_bytes, size_t out_bytes, void** d_in, void** d_out) {
    *d_in = nullptr; *d_out = nullptr;
    if (in_bytes && cudaMalloc(d_in, in_bytes) != cudaSuccess) return B200NUTS_ECUDA;
    if (cudaMalloc(d_out, out_bytes) != cudaSuccess) { cudaFree(*d_in); return B200NUTS_ECUDA; }
    if (in_bytes && cudaMemcpy(*d_in, in, in_bytes, cudaMemcpyHostToDevice) != cudaSuccess) return B200NUTS_ECUDA;
    return 0;
}
static int hook_finish(void* host_out, size_t out_bytes, void* d_in, void* d_out) {
    cudaError_t e = cudaMemcpy(host_out, d_out, out_bytes, cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) { g_create_err = cudaGetErrorString(e); return B200NUTS_ECUDA; }
    return 0;
}

int b200nuts_prng_split(const uint32_t* keys, int64_t n_keys, int32_t num, uint32_t* out) {
    if (!keys || !out || n_keys <= 0 || num <= 0) return B200NUTS_EINVAL;
    void *di, *dout; const size_t ob = (size_t)n_keys * num * 8;
    if (hook_io(keys, (size_t)n_keys * 8, ob, &di, &dout)) return B200NUTS_ECUDA;
    const long long tot = n_keys * num;
    k_prng_split<<<(unsigned)((tot + 255) / 256), 256>>>((const uint32_t*)di, n_keys, num, (uint32_t*)dout);
    return hook_finish(out, ob, di, dout);
}
static int prng_draw(int kind, const uint32_t* key, int64_t n, float lo, float hi, void* out) {
    if (!key || !out || n <= 0) return B200NUTS_EINVAL;
    void *di, *dout;
    if (hook_io(nullptr, 0, (size_t)n * 4, &di, &dout)) return B200NUTS_ECUDA;
    k_prng_draw<<<(unsigned)((n + 255) / 256), 256>>>(kind, key[0], key[1], n, lo, hi, dout);
    return hook_finish(out, (size_t)n * 4, di, dout);
}
int b200nuts_prng_bits(const uint32_t* key, int64_t n, uint32_t* out) { return prng_draw(0, key, n, 0, 1, out); }
int b200nuts_prng_uniform(const uint32_t* key, int64_t n, float lo, float hi, float* out) { return prng_draw(1, key, n, lo, hi, out); }
int b200nuts_prng_normal(const uint32_t* key, int64_t n, float* out) { return prng_draw(2, key, n, 0, 1, out); }
int b200nuts_detmath(int32_t op, const float* x, int64_t n, float* out) {
    if (!x || !out || n <= 0 || op < 0 || op > 4) return B200NUTS_EINVAL;
    void *di, *dout;
    if (hook_io(x, (size_t)n * 4, (size_t)n * 4, &di, &dout)) return B200NUTS_ECUDA;
    k_detmath<<<(unsigned)((n + 255) / 256), 256>>>(op, (const float*)di, n, (float*)dout);
    return hook_finish(out, (size_t)n * 4, di, dout);
}

}  // extern "C"
