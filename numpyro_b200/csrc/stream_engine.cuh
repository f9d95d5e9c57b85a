// Regime R2: HBM-streaming lock-step engine for tall-data GLMs with a handful of chains
// (BASELINE config 2: covtype-shaped logistic regression, N = 581012, D = 54, 8 chains).
//
// One persistent cooperative grid (one CTA per SM) owns the whole MCMC run.  A *pass* serves one
// gradient evaluation of every chain: the grid sweeps X exactly once, computing for all chains
//     eta = X beta_c,  r = dl/deta(eta, y),  gbeta_c += X^T r,  nll_c += l(eta, y)
// in a single fused sweep (the reference does X@beta and X^T@r as two XLA ops = two sweeps, and a
// third for the loss).  Between passes the chains advance their own NUTS trees
// (tick.cuh) -- chain c is owned by warp 1 of CTA c, with its ~10 KB of tree state resident in that
// CTA's shared memory for the whole launch.
//
// Data movement: X is consumed in its natural [N, D] row-major fp32 layout.  A producer warp streams
// contiguous 192-row tiles (41 KB at D = 54) into a 4-stage shared-memory ring with 1-D bulk async
// copies (cp.async.bulk + mbarrier complete_tx); it is decoupled from the pass structure (X is the
// same every pass) so the ring keeps refilling during the inter-pass exchange.
// Compute mapping: a warp step covers 4 rows; the 8 lanes of a row own the strided coefficient
// slices d = j + 8i (beta and the gbeta accumulators live in registers for the whole pass, so the
// inner loop loads only X), partial dot products are transpose-reduced across the 8 lanes so that
// lane j ends up with the logit of chain j, evaluates one link function, and the 8 residuals are
// re-broadcast for the rank-1 update.  FMAs are issued as packed fma.rn.f32x2 (FFMA2).
// Inter-pass exchange (deterministic, no float atomics): per-CTA partials -> global; chain c's
// CTA sums the partials in fixed order, finishes the potential (priors, Jacobians), ticks the
// chain, and publishes the next beta_c; two monotonically increasing counters replace grid.sync.
#pragma once
#include <cuda_runtime.h>
#include "tick.cuh"
#include "families.cuh"

namespace b2 {

constexpr int kStreamCT = 8;             // chains per pass (lanes per row)
constexpr int kStreamWarps = 11;         // consumer warps (+1 producer warp = 384 threads, 168 regs/thread)
constexpr int kStreamThreads = 32 * (kStreamWarps + 1);
constexpr int kMaxTileRows = 192;        // rows per ring slot (runtime tile_rows <= this, multiple of 4*rho)
constexpr int kMaxStages = 6;
constexpr int kTilePad = 64;              // zeroed floats after each X stage (strided over-reads stay finite)
constexpr int kGStride = 72;             // floats per (cta, chain) partial: gbeta[<=64], nll, pad

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
    int rho;                             // row spacing between the 4 row groups of a warp step
    int stages;                          // depth of the shared-memory ring
    int tile_rows;                       // rows per ring slot
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

struct StreamSmem {
    float* tiles; float* ytiles; uint64_t* full; uint64_t* empty;
    float* red; float* gred; float* coef; ChainCtl* ctl; float* cvecs; int* flags;
    int tile_floats;
};

B2_HD size_t stream_smem_bytes(int D, int Dp, int stages, int kTileRows, bool vecs_in_smem) {
    size_t b = 0;
    b += (size_t)stages * ((size_t)kTileRows * D + kTilePad) * 4;  // X ring (+ zero pad per stage)
    b += (size_t)stages * kTileRows * 4;                           // y ring
    b += 2 * kMaxStages * 8;                                       // mbarriers
    b += (size_t)12 * kStreamCT * kGStride * 4;                    // cross-warp reduction + tick scratch
    b += 2 * 64 * 4 + 64 + 64;                                     // gred, coef, flags, debug timers
    b += ((sizeof(ChainCtl) + 15) / 16) * 16;
    if (vecs_in_smem) b += (size_t)V_COUNT * Dp * 4;
    return b + 128;
}

template <int DPL>
__global__ void __launch_bounds__(kStreamThreads, 1) stream_engine_kernel(const StreamParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int D = p.fam.Dx;                        // columns of X (coefficients)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;

    // ---- carve shared memory
    StreamSmem s;
    const int kStages = p.stages, kTileRows = p.tile_rows;
    s.tile_floats = kTileRows * D + kTilePad;
    unsigned char* q = smem_raw;
    s.tiles = (float*)q; q += (size_t)kStages * s.tile_floats * 4;
    s.ytiles = (float*)q; q += (size_t)kStages * kTileRows * 4;
    s.full = (uint64_t*)q; q += kMaxStages * 8;
    s.empty = (uint64_t*)q; q += kMaxStages * 8;
    s.red = (float*)q; q += (size_t)12 * kStreamCT * kGStride * 4;
    s.gred = (float*)q; q += 64 * 4;
    s.coef = (float*)q; q += 64 * 4;
    s.flags = (int*)q; q += 64;
    unsigned long long* tdbg = (unsigned long long*)q; q += 64;
    s.ctl = (ChainCtl*)q; q += ((sizeof(ChainCtl) + 15) / 16) * 16;
    s.cvecs = (float*)q;

    // ---- this CTA's slice of rows, in units of 4 rows so every tile start is 16-byte aligned
    const long long N = p.fam.N;
    const long long Q = N / 4;
    const long long q0 = Q * cta / G, q1 = Q * (cta + 1) / G;
    const long long row0 = 4 * q0, row1 = 4 * q1;
    const int n_tiles = (int)((row1 - row0 + kTileRows - 1) / kTileRows);
    const bool is_tick = cta < p.C;
    const bool has_tail = (cta == G - 1) && (N % 4 != 0);

    // ---- one-time setup: zero the ring (stale/pad words must be finite), init barriers
    for (int i = tid; i < kStages * s.tile_floats + kStages * kTileRows; i += blockDim.x) s.tiles[i] = 0.0f;
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], kStreamWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s.flags[0] = 0;                              // stop flag for the producer
        for (int i = 0; i < 8; ++i) tdbg[i] = 0ull;
    }
    ChainVecs cv;
    if (is_tick) {
        if (p.vecs_in_smem) {
            for (int i = tid; i < V_COUNT * p.Dp; i += blockDim.x) {
                const int f = i / p.Dp, d = i - f * p.Dp;
                s.cvecs[i] = p.vecs[((size_t)f * p.C + cta) * p.Dp + d];
            }
            cv.base = s.cvecs; cv.field_stride = p.Dp;
        } else { cv.base = p.vecs + (size_t)cta * p.Dp; cv.field_stride = p.C * p.Dp; }
        if (tid == 0) *s.ctl = p.ctl[cta];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    // =============================================================== producer warp
    if (warp == 0) {
        if (lane == 0 && n_tiles > 0) {
            uint32_t it = 0;
            while (true) {
                for (int t = 0; t < n_tiles; ++t, ++it) {
                    const int st = it % kStages; const uint32_t ph = (it / kStages) & 1u;
                    if (it >= (uint32_t)kStages) {
                        // wait for the consumers to release the slot; leave early when told to stop
                        while (!mbar_try_wait(&s.empty[st], ph ^ 1u)) {
                            if (*(volatile int*)&s.flags[0]) goto producer_done;
                        }
                    }
                    if (*(volatile int*)&s.flags[0]) goto producer_done;
                    const long long r = row0 + (long long)t * kTileRows;
                    const int rows = (int)((row1 - r < kTileRows) ? (row1 - r) : kTileRows);
                    const uint32_t xb = (uint32_t)rows * (uint32_t)D * 4u, yb = (uint32_t)rows * 4u;
                    mbar_expect_tx(&s.full[st], xb + yb);
                    bulk_g2s(s.tiles + (size_t)st * s.tile_floats, p.fam.X + r * D, xb, &s.full[st]);
                    bulk_g2s(s.ytiles + (size_t)st * kTileRows, p.fam.y + r, yb, &s.full[st]);
                }
            }
        producer_done:
            // drain: every issued copy must have landed before the CTA may exit
            {
                const uint32_t issued = it;
                const uint32_t first = issued > (uint32_t)kStages ? issued - kStages : 0u;
                for (uint32_t k = first; k < issued; ++k) {
                    // a slot whose full-phase was already consumed completes immediately
                    mbar_wait(&s.full[k % kStages], (k / kStages) & 1u);
                }
            }
        }
        return;
    }

    // =============================================================== consumer warps
    const int cw = warp - 1;                         // 0..11
    const int j = lane & 7, rsub = lane >> 3;        // coefficient slice / row group
    const int ctid = tid - 32;                       // 0..383
    const int rho = p.rho;
    const unsigned rho_magic = (unsigned)((0x100000000ULL + (unsigned)rho - 1) / (unsigned)rho);
    unsigned long long bet[DPL][kStreamCT / 2];      // beta[d = j + 8i][chain pair], packed for FFMA2
    unsigned long long acc[DPL][kStreamCT / 2];      // gbeta accumulators, same layout
    uint32_t cons_it = 0;
    unsigned int pass = 0;
    bool ok = true;

    // ---- prologue on the tick CTAs: publish the first beta
    if (is_tick && cw == 0) {
        const float* zsrc = (p.mode == 0) ? cv.v(V_ZS) : (p.z_in + (size_t)cta * p.cfg.D);
        const bool active = (p.mode == 1) || (s.ctl->phase != PH_DONE);
        float* bout = p.beta + (size_t)cta * 64;
        for (int d = lane; d < 64; d += 32) {
            float b = 0.0f;
            if (active && d < D) b = glm_scale_at(p.fam, zsrc, d) * zsrc[p.fam.off_u + d];
            __stcg(bout + d, b);
        }
        __syncwarp();
        if (lane == 0) {
            if (p.mode == 0 && s.ctl->phase == PH_DONE) atomicAdd(&p.sync->done, 1u);
            __threadfence();
            red_release_add(&p.sync->ready, 1u);
        }
    }

    const bool dbg = (cta == 0 && ctid == 0);
    long long t_prev = clock64(), t_red = 0;
    const long long t_begin = t_prev;
#define B2_DBG_LAP(k) do { if (dbg) { const long long t_now = clock64(); tdbg[k] += (unsigned long long)(t_now - t_prev); t_prev = t_now; } } while (0)
    while (true) {
        // ---- wait for every chain's beta of this pass
        if (ctid == 0) {
            ok = spin_ge(&p.sync->ready, (unsigned)p.C * (pass + 1u), p.sync, p.spin_limit);
            s.flags[1] = ok ? 1 : 0;
            s.flags[2] = (int)ld_acquire(&p.sync->done);
            B2_DBG_LAP(0);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kStreamWarps * 32) : "memory");
        if (!s.flags[1] || s.flags[2] >= p.C) break;
#pragma unroll
        for (int i = 0; i < DPL; ++i) {
#pragma unroll
            for (int c = 0; c < kStreamCT / 2; ++c) {
                const int d = j + 8 * i;
                const float b0 = (2 * c < p.C && d < 64) ? __ldcg(p.beta + (2 * c) * 64 + d) : 0.0f;
                const float b1 = (2 * c + 1 < p.C && d < 64) ? __ldcg(p.beta + (2 * c + 1) * 64 + d) : 0.0f;
                bet[i][c] = pack2(b0, b1);
                acc[i][c] = 0ull;
            }
        }
        float nll_acc = 0.0f;                        // lane j accumulates the loss of chain j

        auto step = [&](const float* xrow, float yv, bool valid, bool guard) {
            float x[DPL];
#pragma unroll
            for (int i = 0; i < DPL; ++i) x[i] = (!guard || (j + 8 * i < D)) ? xrow[8 * i] : 0.0f;
            unsigned long long L2[kStreamCT / 2];
#pragma unroll
            for (int c = 0; c < kStreamCT / 2; ++c) L2[c] = 0ull;
#pragma unroll
            for (int i = 0; i < DPL; ++i) {
                const unsigned long long xx = pack2(x[i], x[i]);
#pragma unroll
                for (int c = 0; c < kStreamCT / 2; ++c) ffma2(L2[c], xx, bet[i][c]);
            }
            float L[kStreamCT];
#pragma unroll
            for (int c = 0; c < kStreamCT / 2; ++c) unpack2(L2[c], L[2 * c], L[2 * c + 1]);
            // transpose-reduce over the 8 lanes of the row: lane j keeps chain j
            float M4[4], M2[2], eta;
            {
                const bool hi = (j & 4) != 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float send = hi ? L[k] : L[k + 4];
                    const float keep = hi ? L[k + 4] : L[k];
                    M4[k] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 4);
                }
                const bool mid = (j & 2) != 0;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float send = mid ? M4[k] : M4[k + 2];
                    const float keep = mid ? M4[k + 2] : M4[k];
                    M2[k] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 2);
                }
                const bool lo = (j & 1) != 0;
                const float send = lo ? M2[0] : M2[1];
                const float keep = lo ? M2[1] : M2[0];
                eta = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 1);
            }
            float loss, dl;
            glm_loss_fast(p.fam.likelihood, eta, yv, loss, dl);
            if (!valid) { loss = 0.0f; dl = 0.0f; }
            nll_acc += loss;
            float r[kStreamCT];
#pragma unroll
            for (int c = 0; c < kStreamCT; ++c) r[c] = __shfl_sync(0xFFFFFFFFu, dl, (lane & 24) | c);
#pragma unroll
            for (int i = 0; i < DPL; ++i) {
                const unsigned long long xx = pack2(x[i], x[i]);
#pragma unroll
                for (int c = 0; c < kStreamCT / 2; ++c) ffma2(acc[i][c], xx, pack2(r[2 * c], r[2 * c + 1]));
            }
        };

        // Two independent warp steps written side by side so the scheduler can overlap the
        // shuffle / MUFU latency chain of one with the FFMA2 blocks of the other.
        auto step2 = [&](const float* xa, float ya, bool va, const float* xb, float yb, bool vb) {
            float x[2][DPL];
#pragma unroll
            for (int i = 0; i < DPL; ++i) { x[0][i] = xa[8 * i]; x[1][i] = xb[8 * i]; }
            unsigned long long L2[2][kStreamCT / 2];
#pragma unroll
            for (int c = 0; c < kStreamCT / 2; ++c) { L2[0][c] = 0ull; L2[1][c] = 0ull; }
#pragma unroll
            for (int i = 0; i < DPL; ++i) {
                const unsigned long long xx0 = pack2(x[0][i], x[0][i]), xx1 = pack2(x[1][i], x[1][i]);
#pragma unroll
                for (int c = 0; c < kStreamCT / 2; ++c) { ffma2(L2[0][c], xx0, bet[i][c]); ffma2(L2[1][c], xx1, bet[i][c]); }
            }
            float L[2][kStreamCT];
#pragma unroll
            for (int c = 0; c < kStreamCT / 2; ++c) {
                unpack2(L2[0][c], L[0][2 * c], L[0][2 * c + 1]);
                unpack2(L2[1][c], L[1][2 * c], L[1][2 * c + 1]);
            }
            float M4[2][4], M2[2][2], eta[2];
            const bool hi = (j & 4) != 0, mid = (j & 2) != 0, lo = (j & 1) != 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int q2 = 0; q2 < 2; ++q2) {
                    const float send = hi ? L[q2][k] : L[q2][k + 4];
                    const float keep = hi ? L[q2][k + 4] : L[q2][k];
                    M4[q2][k] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 4);
                }
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
#pragma unroll
                for (int q2 = 0; q2 < 2; ++q2) {
                    const float send = mid ? M4[q2][k] : M4[q2][k + 2];
                    const float keep = mid ? M4[q2][k + 2] : M4[q2][k];
                    M2[q2][k] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 2);
                }
            }
#pragma unroll
            for (int q2 = 0; q2 < 2; ++q2) {
                const float send = lo ? M2[q2][0] : M2[q2][1];
                const float keep = lo ? M2[q2][1] : M2[q2][0];
                eta[q2] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 1);
            }
            float loss0, dl0, loss1, dl1;
            glm_loss_fast(p.fam.likelihood, eta[0], ya, loss0, dl0);
            glm_loss_fast(p.fam.likelihood, eta[1], yb, loss1, dl1);
            if (!va) { loss0 = 0.0f; dl0 = 0.0f; }
            if (!vb) { loss1 = 0.0f; dl1 = 0.0f; }
            nll_acc += loss0;
            nll_acc += loss1;
            float r[2][kStreamCT];
#pragma unroll
            for (int c = 0; c < kStreamCT; ++c) {
                r[0][c] = __shfl_sync(0xFFFFFFFFu, dl0, (lane & 24) | c);
                r[1][c] = __shfl_sync(0xFFFFFFFFu, dl1, (lane & 24) | c);
            }
#pragma unroll
            for (int i = 0; i < DPL; ++i) {
                const unsigned long long xx0 = pack2(x[0][i], x[0][i]), xx1 = pack2(x[1][i], x[1][i]);
#pragma unroll
                for (int c = 0; c < kStreamCT / 2; ++c) {
                    ffma2(acc[i][c], xx0, pack2(r[0][2 * c], r[0][2 * c + 1]));
                    ffma2(acc[i][c], xx1, pack2(r[1][2 * c], r[1][2 * c + 1]));
                }
            }
        };

        // ---- sweep this CTA's tiles
        for (int t = 0; t < n_tiles; ++t, ++cons_it) {
            const int st = cons_it % kStages; const uint32_t ph = (cons_it / kStages) & 1u;
            mbar_wait(&s.full[st], ph);
            const long long r0 = row0 + (long long)t * kTileRows;
            const int rows = (int)((row1 - r0 < kTileRows) ? (row1 - r0) : kTileRows);
            const float* tile = s.tiles + (size_t)st * s.tile_floats;
            const float* yt = s.ytiles + (size_t)st * kTileRows;
            // global step index (t * steps + u) == cw (mod 11): perfectly balanced over the pass
            const int steps = kTileRows / 4;
            const int u0 = (cw + kStreamWarps - (int)(((long long)t * steps) % kStreamWarps)) % kStreamWarps;
            for (int u = u0; u < steps; u += 2 * kStreamWarps) {
                const int blk = (rho == 1) ? u : (int)__umulhi((unsigned)u, rho_magic), sidx = u - blk * rho;   // u / rho
                const int row = blk * 4 * rho + sidx + rho * rsub;
                if (blk * 4 * rho >= rows) break;                 // warp-uniform
                const int ub = u + kStreamWarps;
                const int blkb = (rho == 1) ? ub : (int)__umulhi((unsigned)ub, rho_magic), sidxb = ub - blkb * rho;
                const int rowb = blkb * 4 * rho + sidxb + rho * rsub;
                if (ub < steps && blkb * 4 * rho < rows)
                    step2(tile + row * D + j, yt[row], row < rows, tile + rowb * D + j, yt[rowb], rowb < rows);
                else
                    step(tile + row * D + j, yt[row], row < rows, false);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.empty[st]);
        }
        if (has_tail && cw == 0) {                                // last N % 4 rows, straight from global
            const long long row = 4 * Q + rsub;
            const bool valid = row < N;
            const long long rr = valid ? row : (N - 1);
            step(p.fam.X + rr * D + j, p.fam.y[rr], valid, true);
        }

        // ---- reduce: row groups inside the warp, warps inside the CTA, then publish the partial
#pragma unroll
        for (int i = 0; i < DPL; ++i) {
#pragma unroll
            for (int c = 0; c < kStreamCT / 2; ++c) {
                float a, b; unpack2(acc[i][c], a, b);
                a += __shfl_xor_sync(0xFFFFFFFFu, a, 8);  b += __shfl_xor_sync(0xFFFFFFFFu, b, 8);
                a += __shfl_xor_sync(0xFFFFFFFFu, a, 16); b += __shfl_xor_sync(0xFFFFFFFFu, b, 16);
                if (rsub == 0) {
                    const int d = j + 8 * i;
                    s.red[((size_t)cw * kStreamCT + 2 * c) * kGStride + d] = a;
                    s.red[((size_t)cw * kStreamCT + 2 * c + 1) * kGStride + d] = b;
                }
            }
        }
        nll_acc += __shfl_xor_sync(0xFFFFFFFFu, nll_acc, 8);
        nll_acc += __shfl_xor_sync(0xFFFFFFFFu, nll_acc, 16);
        if (rsub == 0) s.red[((size_t)cw * kStreamCT + j) * kGStride + 64] = nll_acc;
        asm volatile("bar.sync 1, %0;" ::"n"(kStreamWarps * 32) : "memory");
        for (int o = ctid; o < kStreamCT * 65; o += kStreamWarps * 32) {
            const int c = o / 65, d = o - c * 65;
            float a = 0.0f;
#pragma unroll
            for (int w = 0; w < kStreamWarps; ++w) a += s.red[((size_t)w * kStreamCT + c) * kGStride + d];
            __stcg(p.partial + ((size_t)cta * kStreamCT + c) * kGStride + d, a);
        }
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(kStreamWarps * 32) : "memory");
        if (ctid == 0) { red_release_add(&p.sync->arrive, 1u); B2_DBG_LAP(1); }

        // ---- chain owner: sum the partials in fixed order, finish the potential, tick the chain
        if (is_tick) {
            if (ctid == 0) {
                ok = spin_ge(&p.sync->arrive, (unsigned)G * (pass + 1u), p.sync, p.spin_limit);
                s.flags[1] = ok ? 1 : 0;
                B2_DBG_LAP(2);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kStreamWarps * 32) : "memory");
            if (s.flags[1]) {
                // 5 segments x 65 outputs; each thread adds its segment's CTAs in ascending order.
                // Loads are issued 8 at a time (independent, L2 latency overlapped), adds stay ordered.
                const int o = ctid % 65, seg = ctid / 65;         // seg 0..5 (only 0..4 used)
                if (seg < 5) {
                    float a = 0.0f;
                    const int g0 = G * seg / 5, g1 = G * (seg + 1) / 5;
                    const float* src = p.partial + (size_t)cta * kGStride + o;
                    for (int g = g0; g < g1; g += 8) {
                        float v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            v[k] = (g + k < g1) ? __ldcg(src + (size_t)(g + k) * (kStreamCT * kGStride)) : 0.0f;
#pragma unroll
                        for (int k = 0; k < 8; ++k) a += v[k];
                    }
                    s.red[seg * kGStride + o] = a;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kStreamWarps * 32) : "memory");
                if (cw == 0) {
                    B2_DBG_LAP(3);
                    for (int d = lane; d < 65; d += 32)
                        s.gred[d] = (((s.red[d] + s.red[kGStride + d]) + s.red[2 * kGStride + d]) + s.red[3 * kGStride + d]) +
                                    s.red[4 * kGStride + d];
                    __syncwarp();
                    const float nll = s.gred[64];
                    float* gz = s.red + 8 * kGStride;             // scratch for the gradient wrt z (<= Dp floats)
                    float u;
                    bool finished = false;
                    if (p.mode == 1) {
                        const float* zin = p.z_in + (size_t)cta * p.cfg.D;
                        glm_finish(p.fam, zin, nll, s.gred, u, gz);
                        __syncwarp();
                        if (lane == 0) p.u_out[cta] = u;
                        for (int d = lane; d < p.cfg.D; d += 32) p.g_out[(size_t)cta * p.cfg.D + d] = gz[d];
                        finished = true;
                    } else if (s.ctl->phase != PH_DONE) {
                        ChainCtl c = *s.ctl;
                        __syncwarp();
                        glm_finish(p.fam, cv.v(V_ZS), nll, s.gred, u, gz);
                        __syncwarp();
                        Tick tk{p.cfg, c, cv, p.out, cta, p.C};
                        tk.advance(u, gz);
                        __syncwarp();
                        if (lane == 0) *s.ctl = c;
                        finished = (c.phase == PH_DONE);
                        const float* zs = cv.v(V_ZS);
                        float* bout = p.beta + (size_t)cta * 64;
                        for (int d = lane; d < 64; d += 32) {
                            float b = 0.0f;
                            if (!finished && d < D) b = glm_scale_at(p.fam, zs, d) * zs[p.fam.off_u + d];
                            __stcg(bout + d, b);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) {
                        if (finished) atomicAdd(&p.sync->done, 1u);
                        __threadfence();
                        red_release_add(&p.sync->ready, 1u);
                        B2_DBG_LAP(4);
                    }
                }
            }
        }
        ++pass;
    }

    // ---- shutdown: stop the producer, write the chain state back
    if (ctid == 0) { *(volatile int*)&s.flags[0] = 1; }
    asm volatile("bar.sync 1, %0;" ::"n"(kStreamWarps * 32) : "memory");
    if (is_tick) {
        if (p.vecs_in_smem && p.mode == 0) {
            for (int i = ctid; i < V_COUNT * p.Dp; i += kStreamWarps * 32) {
                const int f = i / p.Dp, d = i - f * p.Dp;
                p.vecs[((size_t)f * p.C + cta) * p.Dp + d] = s.cvecs[i];
            }
        }
        if (ctid == 0 && p.mode == 0) p.ctl[cta] = *s.ctl;
    }
    if (cta == 0 && ctid == 0) {
        p.sync->passes = pass;
        tdbg[5] = (unsigned long long)(clock64() - t_begin);
        for (int i = 0; i < 8; ++i) p.sync->dbg[i] = tdbg[i];
    }
    (void)t_red;
}

}  // namespace b2
