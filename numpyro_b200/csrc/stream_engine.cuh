// Regime R2: HBM-streaming lock-step engine for tall-data GLMs with a handful of chains
// (BASELINE config 2: covtype-shaped logistic regression, N = 581012, D = 54, 8 chains).
//
// One persistent cooperative grid (one CTA per SM) owns the whole MCMC run.  A *pass* serves one
// gradient evaluation of every chain: the grid sweeps X exactly once, computing for all chains
//     eta = X beta_c,  r = dl/deta(eta, y),  gbeta_c += X^T r,  nll_c += l(eta, y)
// in a single fused sweep (the reference does X@beta and X^T@r as two XLA ops = two sweeps, and a
// third for the loss).  Between passes the chains advance their own NUTS trees
// (tick.cuh): chain c is owned by the tick warp of CTA c, with its ~10 KB of tree state resident in
// that CTA's shared memory for the whole launch.
//
// CTA = 16 warps: warp 0 producer, warps 1..14 consumers (7 pairs), warp 15 tick.
// Data movement: X is consumed in its natural [N, D] row-major fp32 layout.  The producer streams
// contiguous 64-row tiles (13.8 KB at D = 54) into a <=12-deep shared-memory ring with 1-D bulk
// async copies (cp.async.bulk + mbarrier complete_tx); it free-runs across passes (X is the same
// every pass), so the ring keeps refilling during the inter-pass exchange.
// Compute mapping (tile t belongs to consumer pair t mod 7, each warp takes 32 of its 64 rows):
//   forward : half-warp h owns chains 4h..4h+3, lane l owns rows l and l+16; x comes as 8-byte
//             loads down the lane's own rows (conflict-free when D/2 is odd), beta as 16-byte
//             broadcast loads from shared memory; 8 logits per lane, no cross-lane reduction;
//   link    : each lane evaluates its 8 (row, chain) losses/residuals, residuals go to a 1 KB
//             per-warp buffer;
//   backward: half-warp h again owns chains 4h..4h+3, lane q owns 4 columns; per row one pair of
//             8-byte x loads, one 16-byte residual load and 8 packed FMAs into register accumulators
//             that live for the whole pass.
// FMAs are issued as packed fma.rn.f32x2 (SASS FFMA2).
// Inter-pass exchange (deterministic, no float atomics): per-CTA partials -> global; chain c's
// CTA sums the partials in fixed order, its tick warp finishes the potential (priors, Jacobians),
// ticks the chain and publishes the next beta_c; two monotone counters replace grid.sync.
#pragma once
#include <cuda_runtime.h>
#include "tick.cuh"
#include "families.cuh"

namespace b2 {

constexpr int kStreamCT = 8;             // chains per pass
constexpr int kConsWarps = 14;           // consumer warps = 7 pairs
constexpr int kPairs = kConsWarps / 2;
constexpr int kStreamThreads = 32 * (kConsWarps + 2);
constexpr int kTileRows = 64;            // rows per ring slot (two 32-row units)
constexpr int kMaxStages = 12;
constexpr int kGStride = 72;             // floats per (cta, chain) partial: gbeta[<=64], nll, pad
constexpr int kBarTop = 1, kBarCons = 2, kBarTick = 3;
constexpr int kConsThreads = kConsWarps * 32, kTopThreads = (kConsWarps + 1) * 32;

// dbg: clock64 totals on CTA 0 -- [0] wait for betas, [1] sweep + publish partial, [2] wait for all partials,
// [3] cross-CTA reduction, [4] finish + tick + publish beta, [5] total loop
struct StreamSync { unsigned int arrive, ready, done, abort_flag; unsigned long long passes; unsigned long long dbg[8]; };

struct StreamParams {
    TickCfg cfg; FamilySpec fam; OutBufs out;
    int C, Dp, mode;                     // mode 0: run chains, 1: evaluate potential at z_in
    ChainCtl* ctl; float* vecs;          // [C], [V_COUNT][C][Dp]
    float* partial;                      // [grid][kStreamCT][kGStride]
    float* beta;                         // [kStreamCT][64]
    StreamSync* sync;
    const float* z_in; float* u_out; float* g_out;    // mode 1
    int stages;                          // depth of the shared-memory ring
    int vecs_in_smem;
    long long spin_limit;
};

// ---- PTX wrappers ---------------------------------------------------------------------------
B2_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
B2_D void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
B2_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
B2_D void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
B2_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
B2_D void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
B2_D void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
B2_D unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
B2_D void red_release_add(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
B2_D unsigned long long pack2(float lo, float hi) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
B2_D void unpack2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
B2_D void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
// tf32 split: hi keeps the top 19 bits (exactly what the tensor core consumes), lo = x - hi (exact in fp32;
// its own low bits are dropped by the MMA: relative error 2^-22 of x).
B2_D void tf32_split(uint32_t& hi, uint32_t& lo) {
    const uint32_t h = hi & 0xFFFFE000u;
    lo = __float_as_uint(__uint_as_float(hi) - __uint_as_float(h));
    hi = h;
}
B2_D void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int ID, int N> B2_D void bar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
template <int ID, int N> B2_D void bar_arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(N) : "memory"); }

// Spin until *ctr >= target.  Gives up (and raises the abort flag for everybody) after spin_limit
// clocks so that a protocol bug can never wedge the GPU.
B2_D bool spin_ge(const unsigned int* ctr, unsigned int target, StreamSync* sy, long long limit) {
    const long long t0 = clock64();
    while (ld_acquire(ctr) < target) {
        if (ld_acquire(&sy->abort_flag)) return false;
        if (clock64() - t0 > limit) { atomicExch(&sy->abort_flag, 1u); return false; }
    }
    return true;
}

B2_HD size_t stream_fixed_smem(int Dp, bool vecs_in_smem) {
    size_t b = 0;
    b += 2 * kMaxStages * 8;                                       // mbarriers
    b += 64 * kStreamCT * 4;                                       // beta, [d][chain]
    b += (size_t)kConsWarps * 32 * kStreamCT * 4;                  // residual buffers, [warp][row][chain]
    b += (size_t)kConsWarps * kStreamCT * kGStride * 4;            // cross-warp reduction + tick scratch
    b += 64 * 4 + 64 * 4 + 64 + 64;                                // gred(+nll), tail tile header, flags, timers
    b += 16 * 64 * 4 + 64;                                         // tail group (<= 3 valid rows of a zeroed 16 x 64 block) + y
    b += ((sizeof(ChainCtl) + 15) / 16) * 16;
    if (vecs_in_smem) b += (size_t)V_COUNT * Dp * 4;
    return b + 128;
}
B2_HD size_t stream_smem_bytes(int D, int Dp, int stages, bool vecs_in_smem) {
    return stream_fixed_smem(Dp, vecs_in_smem) + (size_t)stages * ((size_t)kTileRows * D * 4 + kTileRows * 4);
}

// KS = ceil(D / 8): k-steps of the forward MMA (compile time so every fragment stays in registers).
template <int KS>
__global__ void __launch_bounds__(kStreamThreads, 1) stream_engine_kernel(const StreamParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int D = p.fam.Dx;                        // columns of X (coefficients)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    const int nst = p.stages;

    // ---- carve shared memory
    unsigned char* q = smem_raw;
    float* tiles = (float*)q; q += (size_t)nst * kTileRows * D * 4;
    float* ytiles = (float*)q; q += (size_t)nst * kTileRows * 4;
    uint64_t* full = (uint64_t*)q; q += kMaxStages * 8;
    uint64_t* empty = (uint64_t*)q; q += kMaxStages * 8;
    float* bs = (float*)q; q += 64 * kStreamCT * 4;
    float* rbuf_all = (float*)q; q += (size_t)kConsWarps * 32 * kStreamCT * 4;
    float* red = (float*)q; q += (size_t)kConsWarps * kStreamCT * kGStride * 4;
    float* gred = (float*)q; q += 64 * 4 + 64 * 4;
    int* flags = (int*)q; q += 64;
    unsigned long long* tdbg = (unsigned long long*)q; q += 64;
    float* tail = (float*)q; q += 16 * 64 * 4 + 64;            // rows N - N%4 .. N-1 in a zeroed 16-row group, then their y
    ChainCtl* sctl = (ChainCtl*)q; q += ((sizeof(ChainCtl) + 15) / 16) * 16;
    float* cvecs = (float*)q;
    const int tile_floats = kTileRows * D;

    // ---- this CTA's slice of rows, in units of 4 rows so every tile start is 16-byte aligned
    const long long N = p.fam.N;
    const long long Q = N / 4;
    const long long q0 = Q * cta / G, q1 = Q * (cta + 1) / G;
    const long long row0 = 4 * q0, row1 = 4 * q1;
    const int n_tiles = (int)((row1 - row0 + kTileRows - 1) / kTileRows);
    const bool is_tick = cta < p.C;
    const int n_tail = (cta == G - 1) ? (int)(N % 4) : 0;

    // ---- one-time setup: zero the ring (stale words must be finite), init barriers
    for (int i = tid; i < nst * (tile_floats + kTileRows); i += blockDim.x) tiles[i] = 0.0f;
    for (int i = tid; i < kConsWarps * kStreamCT * kGStride; i += blockDim.x) red[i] = 0.0f;
    for (int i = tid; i < 16 * 64 + 16; i += blockDim.x) {
        float v = 0.0f;
        if (n_tail > 0) {
            if (i < 16 * 64) { const int r = i / D, d = i - r * D; if (i < n_tail * D) v = p.fam.X[(4 * Q + r) * D + d]; }
            else if (i - 16 * 64 < n_tail) v = p.fam.y[4 * Q + (i - 16 * 64)];
        }
        tail[i] = v;
    }
    if (tid == 0) {
        for (int i = 0; i < nst; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        flags[0] = 0;                                // stop flag for the producer
        for (int i = 0; i < 8; ++i) tdbg[i] = 0ull;
    }
    ChainVecs cv; cv.base = nullptr; cv.field_stride = 0;
    if (is_tick) {
        if (p.vecs_in_smem) {
            for (int i = tid; i < V_COUNT * p.Dp; i += blockDim.x) {
                const int f = i / p.Dp, d = i - f * p.Dp;
                cvecs[i] = p.vecs[((size_t)f * p.C + cta) * p.Dp + d];
            }
            cv.base = cvecs; cv.field_stride = p.Dp;
        } else { cv.base = p.vecs + (size_t)cta * p.Dp; cv.field_stride = p.C * p.Dp; }
        if (tid == 0) *sctl = p.ctl[cta];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    // =============================================================== producer warp
    if (warp == 0) {
        if (lane == 0 && n_tiles > 0) {
            uint32_t it = 0;
            while (true) {
                for (int t = 0; t < n_tiles; ++t, ++it) {
                    const int st = it % nst; const uint32_t ph = (it / nst) & 1u;
                    if (it >= (uint32_t)nst) {
                        // wait for both consumer warps of the slot's previous tile; leave early when told to stop
                        while (!mbar_try_wait(&empty[st], ph ^ 1u)) {
                            if (*(volatile int*)&flags[0]) goto producer_done;
                        }
                    }
                    if (*(volatile int*)&flags[0]) goto producer_done;
                    const long long r = row0 + (long long)t * kTileRows;
                    const int rows = (int)((row1 - r < kTileRows) ? (row1 - r) : kTileRows);
                    const uint32_t xb = (uint32_t)rows * (uint32_t)D * 4u, yb = (uint32_t)rows * 4u;
                    mbar_expect_tx(&full[st], xb + yb);
                    bulk_g2s(tiles + (size_t)st * tile_floats, p.fam.X + r * D, xb, &full[st]);
                    bulk_g2s(ytiles + (size_t)st * kTileRows, p.fam.y + r, yb, &full[st]);
                }
            }
        producer_done:
            {   // drain: every issued copy must have landed before the CTA may exit
                const uint32_t issued = it;
                const uint32_t first = issued > (uint32_t)nst ? issued - nst : 0u;
                for (uint32_t k = first; k < issued; ++k) mbar_wait(&full[k % nst], (k / nst) & 1u);
            }
        }
        return;
    }

    StreamSync* sy = p.sync;
    unsigned int pass = 0;

    // =============================================================== tick warp
    if (warp == kConsWarps + 1) {
        const bool dbg = (cta == 0 && lane == 0);
        long long t_prev = clock64();
        auto publish_beta = [&](const float* zsrc, bool active) {
            float* bout = p.beta + (size_t)cta * 64;
            for (int d = lane; d < 64; d += 32) {
                float b = 0.0f;
                if (active && d < D) b = glm_scale_at(p.fam, zsrc, d) * zsrc[p.fam.off_u + d];
                __stcg(bout + d, b);
            }
            __syncwarp();
        };
        if (is_tick) {                               // prologue: the first beta
            const float* zsrc = (p.mode == 0) ? cv.v(V_ZS) : (p.z_in + (size_t)cta * p.cfg.D);
            const bool active = (p.mode == 1) || (sctl->phase != PH_DONE);
            publish_beta(zsrc, active);
            if (lane == 0) {
                if (p.mode == 0 && sctl->phase == PH_DONE) atomicAdd(&sy->done, 1u);
                __threadfence();
                red_release_add(&sy->ready, 1u);
            }
        }
        while (true) {
            bar_sync<kBarTop, kTopThreads>();
            if (!flags[1] || flags[2] >= p.C) break;
            if (is_tick) {
                bar_sync<kBarTick, kTopThreads>();   // segment sums are in `red`
                if (dbg) t_prev = clock64();
                if (flags[3]) {
                    for (int d = lane; d < 65; d += 32)
                        gred[d] = (((red[d] + red[kGStride + d]) + red[2 * kGStride + d]) + red[3 * kGStride + d]) +
                                  red[4 * kGStride + d];
                    __syncwarp();
                    const float nll = gred[64];
                    float* gz = red + 8 * kGStride;  // scratch for the gradient wrt z (<= Dp floats)
                    float u;
                    bool finished = false;
                    if (p.mode == 1) {
                        const float* zin = p.z_in + (size_t)cta * p.cfg.D;
                        glm_finish(p.fam, zin, nll, gred, u, gz);
                        __syncwarp();
                        if (lane == 0) p.u_out[cta] = u;
                        for (int d = lane; d < p.cfg.D; d += 32) p.g_out[(size_t)cta * p.cfg.D + d] = gz[d];
                        finished = true;
                    } else if (sctl->phase != PH_DONE) {
                        ChainCtl c = *sctl;
                        __syncwarp();
                        glm_finish(p.fam, cv.v(V_ZS), nll, gred, u, gz);
                        __syncwarp();
                        Tick tk{p.cfg, c, cv, p.out, cta, p.C};
                        tk.advance(u, gz);
                        __syncwarp();
                        if (lane == 0) *sctl = c;
                        finished = (c.phase == PH_DONE);
                        publish_beta(cv.v(V_ZS), !finished);
                    }
                    __syncwarp();
                    if (lane == 0) {
                        if (finished) atomicAdd(&sy->done, 1u);
                        __threadfence();
                        red_release_add(&sy->ready, 1u);
                        if (dbg) tdbg[4] += (unsigned long long)(clock64() - t_prev);
                    }
                }
            }
            ++pass;
        }
        if (is_tick && lane == 0 && p.mode == 0) p.ctl[cta] = *sctl;
        return;
    }

    // =============================================================== consumer warps
    const int cw = warp - 1;                         // 0..13
    const int ctid = tid - 32;                       // 0..447
    const int pair = cw >> 1, half = cw & 1;         // tile owner pair, which 32 rows of the tile
    const int g = lane >> 2, t = lane & 3;           // mma.sync fragment coordinates (groupID, threadID_in_group)
    constexpr int MT = (KS + 1) / 2;                 // 16-column tiles of the backward product
    uint32_t bhi[KS][2], blo[KS][2];                 // beta as B fragments of the forward MMA, tf32 hi / lo parts
    float gacc[MT][4];                               // gbeta in C-fragment layout: d = 16mt + g (+8), chain 2t (+1)
    float nll[2];                                    // loss of chains 2t, 2t+1 over this lane's rows
    uint32_t tiles_done = 0;                         // running tile counter of this CTA (ring position)

    const bool dbg = (cta == 0 && ctid == 0);
    long long t_prev = clock64();
    const long long t_begin = t_prev;
#define B2_DBG_LAP(k) do { if (dbg) { const long long t_now = clock64(); tdbg[k] += (unsigned long long)(t_now - t_prev); t_prev = t_now; } } while (0)

    // One 32-row unit = two 16-row groups.  Per group: forward MMA (logits), link function in the
    // C-fragment layout (4 (row, chain) pairs per lane, no reduction), residuals shuffled into B-fragment
    // layout, backward MMA (gbeta).  All products are 3xTF32 (hi*hi + hi*lo + lo*hi, fp32 accumulate).
    // Row permutations inside a group are chosen so every LDS.32 is bank-conflict free at D = 54:
    // forward fragment rows (g, g+8) <-> data rows (2g, 2g+1); backward k index (t, t+4) <-> rows (4t+2ks, +1).
    auto unit = [&](const float* xt, const float* yt, int nvalid) {
#pragma unroll 1
        for (int grp = 0; grp < 2; ++grp) {
            if (16 * grp >= nvalid) break;
            const float* xg = xt + 16 * grp * D;
            const float* yg = yt + 16 * grp;
            const int nv = nvalid - 16 * grp;
            // ---- forward
            const float* xa = xg + (2 * g) * D + t;
            const float* xb = xa + D;
            float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                uint32_t a[4], al[4];
                a[0] = __float_as_uint(xa[8 * kk]); a[1] = __float_as_uint(xb[8 * kk]);
                a[2] = __float_as_uint(xa[8 * kk + 4]); a[3] = __float_as_uint(xb[8 * kk + 4]);
#pragma unroll
                for (int i = 0; i < 4; ++i) tf32_split(a[i], al[i]);
                mma_tf32(c, al, bhi[kk][0], bhi[kk][1]);
                mma_tf32(c, a, blo[kk][0], blo[kk][1]);
                mma_tf32(c, a, bhi[kk][0], bhi[kk][1]);
            }
            // ---- link: c0 (row 2g, chain 2t), c1 (2g, 2t+1), c2 (2g+1, 2t), c3 (2g+1, 2t+1)
            float dl[4];
            {
                const float ya = yg[2 * g], yb = yg[2 * g + 1];
                const bool va = (2 * g) < nv, vb = (2 * g + 1) < nv;
                float ls[4];
                glm_loss_fast(p.fam.likelihood, c[0], ya, ls[0], dl[0]);
                glm_loss_fast(p.fam.likelihood, c[1], ya, ls[1], dl[1]);
                glm_loss_fast(p.fam.likelihood, c[2], yb, ls[2], dl[2]);
                glm_loss_fast(p.fam.likelihood, c[3], yb, ls[3], dl[3]);
                if (!va) { ls[0] = 0.0f; ls[1] = 0.0f; dl[0] = 0.0f; dl[1] = 0.0f; }
                if (!vb) { ls[2] = 0.0f; ls[3] = 0.0f; dl[2] = 0.0f; dl[3] = 0.0f; }
                nll[0] += ls[0]; nll[1] += ls[1]; nll[0] += ls[2]; nll[1] += ls[3];
            }
            // ---- residuals -> B fragments of the backward MMA: b0 = r[row 4t+2ks][chain g], b1 = r[row 4t+2ks+1][chain g]
            uint32_t rh[2][2], rl[2][2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int src = ((2 * t + ks) << 2) | (g >> 1);
                const float e0 = __shfl_sync(0xFFFFFFFFu, dl[0], src), e1 = __shfl_sync(0xFFFFFFFFu, dl[1], src);
                const float o0 = __shfl_sync(0xFFFFFFFFu, dl[2], src), o1 = __shfl_sync(0xFFFFFFFFu, dl[3], src);
                rh[ks][0] = __float_as_uint((g & 1) ? e1 : e0);
                rh[ks][1] = __float_as_uint((g & 1) ? o1 : o0);
                tf32_split(rh[ks][0], rl[ks][0]);
                tf32_split(rh[ks][1], rl[ks][1]);
            }
            // ---- backward: gbeta[16mt + m][chain] += sum_rows x[row][16mt + m] * r[row][chain]
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                float gt[4] = {0.0f, 0.0f, 0.0f, 0.0f};       // fresh accumulator per group: few tensor-core adds, then fp32 FADD
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const float* x0 = xg + (4 * t + 2 * ks) * D + 16 * mt + g;
                    const float* x1 = x0 + D;
                    uint32_t a[4], al[4];
                    a[0] = __float_as_uint(x0[0]); a[1] = __float_as_uint(x0[8]);
                    a[2] = __float_as_uint(x1[0]); a[3] = __float_as_uint(x1[8]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) tf32_split(a[i], al[i]);
                    mma_tf32(gt, al, rh[ks][0], rh[ks][1]);
                    mma_tf32(gt, a, rl[ks][0], rl[ks][1]);
                    mma_tf32(gt, a, rh[ks][0], rh[ks][1]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) gacc[mt][i] += gt[i];
            }
        }
    };

    while (true) {
        // ---- wait for every chain's beta of this pass
        if (ctid == 0) {
            const bool ok = spin_ge(&sy->ready, (unsigned)p.C * (pass + 1u), sy, p.spin_limit);
            flags[1] = ok ? 1 : 0;
            flags[2] = (int)ld_acquire(&sy->done);
            B2_DBG_LAP(0);
        }
        bar_sync<kBarTop, kTopThreads>();
        if (!flags[1] || flags[2] >= p.C) break;
        for (int i = ctid; i < 64 * kStreamCT; i += kConsThreads) {        // beta -> shared, [d][chain]
            const int d = i >> 3, c = i & 7;
            bs[i] = (c < p.C) ? __ldcg(p.beta + c * 64 + d) : 0.0f;
        }
        bar_sync<kBarCons, kConsThreads>();
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {                                   // beta -> B fragments (k = d, n = chain)
            bhi[kk][0] = __float_as_uint(bs[(8 * kk + t) * kStreamCT + g]);
            bhi[kk][1] = __float_as_uint(bs[(8 * kk + t + 4) * kStreamCT + g]);
            tf32_split(bhi[kk][0], blo[kk][0]);
            tf32_split(bhi[kk][1], blo[kk][1]);
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) { gacc[mt][0] = 0.0f; gacc[mt][1] = 0.0f; gacc[mt][2] = 0.0f; gacc[mt][3] = 0.0f; }
        nll[0] = 0.0f; nll[1] = 0.0f;

        // ---- sweep: tile t of this pass belongs to pair t mod 7; this warp takes rows [32*half, 32*half+32)
        for (int t = pair; t < n_tiles; t += kPairs) {
            const uint32_t gi = tiles_done + (uint32_t)t;
            const int st = gi % nst; const uint32_t ph = (gi / nst) & 1u;
            mbar_wait(&full[st], ph);
            const long long r0 = row0 + (long long)t * kTileRows;
            const int rows = (int)((row1 - r0 < kTileRows) ? (row1 - r0) : kTileRows);
            const int nvalid = rows - 32 * half;
            if (nvalid > 0) unit(tiles + (size_t)st * tile_floats + 32 * half * D, ytiles + (size_t)st * kTileRows + 32 * half, nvalid);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
        tiles_done += (uint32_t)n_tiles;
        if (n_tail > 0 && cw == 0) unit(tail, tail + 16 * 64, n_tail);      // last N % 4 rows (staged at start)

        // ---- reduce: lanes -> warp -> CTA, then publish the partial
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int d = 16 * mt + g + ((i & 2) ? 8 : 0), chain = 2 * t + (i & 1);
                if (d < D) red[((size_t)cw * kStreamCT + chain) * kGStride + d] = gacc[mt][i];
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            nll[k] += __shfl_xor_sync(0xFFFFFFFFu, nll[k], 4);
            nll[k] += __shfl_xor_sync(0xFFFFFFFFu, nll[k], 8);
            nll[k] += __shfl_xor_sync(0xFFFFFFFFu, nll[k], 16);
        }
        if (g == 0) {
            red[((size_t)cw * kStreamCT + 2 * t) * kGStride + 64] = nll[0];
            red[((size_t)cw * kStreamCT + 2 * t + 1) * kGStride + 64] = nll[1];
        }
        bar_sync<kBarCons, kConsThreads>();
        for (int o = ctid; o < kStreamCT * 65; o += kConsThreads) {
            const int c = o / 65, d = o - c * 65;
            float a = 0.0f;
#pragma unroll
            for (int w = 0; w < kConsWarps; ++w) a += red[((size_t)w * kStreamCT + c) * kGStride + d];
            __stcg(p.partial + ((size_t)cta * kStreamCT + c) * kGStride + d, a);
        }
        __threadfence();
        bar_sync<kBarCons, kConsThreads>();
        if (ctid == 0) { red_release_add(&sy->arrive, 1u); B2_DBG_LAP(1); }

        // ---- chain owner: sum the partials of all CTAs in fixed order, hand over to the tick warp
        if (is_tick) {
            if (ctid == 0) {
                const bool ok = spin_ge(&sy->arrive, (unsigned)G * (pass + 1u), sy, p.spin_limit);
                flags[3] = ok ? 1 : 0;
                B2_DBG_LAP(2);
            }
            bar_sync<kBarCons, kConsThreads>();
            if (flags[3]) {
                // 5 segments x 65 outputs; each thread adds its segment's CTAs in ascending order.
                // Loads are issued 16 at a time (independent, L2 latency overlapped), adds stay ordered.
                const int o = ctid % 65, seg = ctid / 65;         // seg 0..6 (only 0..4 used)
                if (seg < 5) {
                    float a = 0.0f;
                    const int g0 = G * seg / 5, g1 = G * (seg + 1) / 5;
                    const float* src = p.partial + (size_t)cta * kGStride + o;
                    for (int g = g0; g < g1; g += 16) {
                        float v[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k)
                            v[k] = (g + k < g1) ? __ldcg(src + (size_t)(g + k) * (kStreamCT * kGStride)) : 0.0f;
#pragma unroll
                        for (int k = 0; k < 16; ++k) a += v[k];
                    }
                    red[seg * kGStride + o] = a;
                }
            }
            __threadfence_block();
            bar_arrive<kBarTick, kTopThreads>();     // tick warp takes over; consumers go wait for the next beta
            if (ctid == 0) B2_DBG_LAP(3);
        }
        ++pass;
    }

    // ---- shutdown: stop the producer, write the chain state back
    if (ctid == 0) { *(volatile int*)&flags[0] = 1; }
    bar_sync<kBarCons, kConsThreads>();
    if (is_tick && p.vecs_in_smem && p.mode == 0) {
        // the tick warp left the loop through the same top barrier, so the vectors are final
        for (int i = ctid; i < V_COUNT * p.Dp; i += kConsThreads) {
            const int f = i / p.Dp, d = i - f * p.Dp;
            p.vecs[((size_t)f * p.C + cta) * p.Dp + d] = cvecs[i];
        }
    }
    if (cta == 0 && ctid == 0) {
        sy->passes = pass;
        tdbg[5] = (unsigned long long)(clock64() - t_begin);
        for (int i = 0; i < 8; ++i) sy->dbg[i] = tdbg[i];
    }
#undef B2_DBG_LAP
}

}  // namespace b2
