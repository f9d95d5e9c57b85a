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
// Data layout.  b200nuts_create repacks the caller's X [N, D] / y [N] once into an engine-owned
// *tile image* in HBM.  Tile i = rows 16i .. 16i+15 and their 16 responses, one contiguous 16-byte
// aligned block moved by ONE 1-D bulk async copy (cp.async.bulk + mbarrier complete_tx).  Inside a
// tile rows r and r+8 are interleaved ("pair rows"): the four floats
//     { X[r][c], X[r+8][c], X[r][c+1], X[r+8][c+1] }            (r < 8, c even)
// are adjacent.  That quad IS the A fragment of the forward mma.sync (rows g / g+8, k = t / t+4) and
// its two halves ARE B fragments of the backward one (k = t / t+4 rows, n = column), so every
// fragment is filled by a single 128-bit shared-memory load with no register shuffling.  Columns are
// padded to P = 8*odd floats and 16-byte chunks are XOR-swizzled by bit 1 of the pair row, which
// makes both access patterns bank-conflict free.
//
// CTA = 16 warps: warps 0..14 consumers, warp 15 tick.  Tile t of a CTA's slice belongs to consumer
// warp t mod 15 (static assignment: the summation order is fixed, results are bit-reproducible).
// Every consumer warp runs its own private ring of `stages` slots and refills a slot itself the
// moment it has finished reading it (one elected lane: expect_tx + bulk copy) -- measured on B200, a
// single issuing thread needs ~680 cycles per copy round trip and caps a CTA at ~3 TB/s chip-wide, 15
// independent issuers do not.  The tile sequence of a warp simply wraps around at the end of a pass
// (X is the same every pass), so the first `stages` tiles of the next pass are already in flight
// while the chains exchange gradients.
// Per tile (mma.sync m16n8k8, TF32 operands split hi/lo, fp32 accumulate -- fp32-level accuracy):
//   forward : logits[16 rows][8 chains] = X_t beta, 3 products (lo*hi, hi*lo, hi*hi) in independent
//             accumulator chains; beta lives in B fragments in registers for the whole pass;
//   link    : each lane owns 4 (row, chain) pairs: loss + residual (ex2/lg2/rcp);
//   shuffle : residuals move from C-fragment to A-fragment layout (8 shuffles), split hi/lo and
//             *stacked*: A = [r_hi ; r_lo]^T (16 x 8 rows);
//   backward: gbeta^T[chain][col] += A X_t, X as the B operand (hi then lo part): rows 0..7 of the
//             result collect r_hi x, rows 8..15 the r_lo x correction; 8 columns per MMA.  A chunk of 16
//             columns accumulates over the tile's 2 k-steps on the tensor core, then joins the
//             per-pass fp32 accumulators with FADDs.
// Inter-pass exchange (deterministic, no float atomics, no counters, no fences).  Everything that crosses
// CTAs carries its own sequence tag inside the same 8- or 16-byte word as the data ("flag-in-data"), so a reader
// simply polls the data it needs:
//   * every CTA reduces its warps' accumulators through the ring slots its warps just drained (thread i adds word i
//     of the 15 scratch slots: conflict-free) and publishes its partial as 16-byte words {v0, v1, v2, tag}, laid out
//     [group][chain][cta][word] so that what one owner needs from a range of CTAs is one contiguous block;
//   * chain c's CTA: every consumer warp fetches the block of ITS range of CTAs with one bulk copy into the scratch
//     slot it has just drained, checks the tags in shared memory (a stale block is fetched again), adds the rows in
//     CTA order; the 15 ranges are joined in warp order.  The tick warp finishes the potential (priors, Jacobians),
//     ticks the chain (tick.cuh) and publishes the next beta_c as tagged pairs in MMA-fragment order (4 replicas
//     to spread the L2 load);
//   * the tick warp of every CTA sleeps on a barrier until its own consumers have finished the sweep, then fetches
//     all chains' beta, stages them in shared memory and releases the CTA's consumer warps through a named barrier.
//     While the consumers sweep, an owner's tick warp runs the deferred half of the tick and looks ahead in the
//     chain's PRNG streams (Tick::prefetch) so the next tick finds its random numbers ready.
// Register discipline: the sweep body (stream_sweep) derives everything it needs from volatile thread / block id
// reads once per pass, and the exchange steps are out-of-line functions with small register footprints (the
// gather keeps its polled words in shared memory): ptxas otherwise moved spills into the MMA region.
#pragma once
#include <cuda_runtime.h>
#include "tick.cuh"
#include "families.cuh"
#include "linkfn.cuh"

namespace b2 {

using StreamTick = TickT<false>;         // no dense mass matrices in this regime (see TickT)

constexpr int kStreamCT = 8;             // chains per pass (= one chain group = the N of the MMAs)
constexpr int kMaxGroups = 20;           // chain groups per handle: passes rotate over the groups, so one group's exchange
constexpr int kMaxStreamChains = kMaxGroups * kStreamCT;   // ... hides behind the other groups' sweeps (every chain needs its own owner CTA)
constexpr int kConsWarps = 15;           // consumer warps
constexpr int kStreamThreads = 32 * (kConsWarps + 1);
constexpr int kTileRows = 16;            // rows per ring slot = one consumer warp's unit of work
constexpr int kMaxStages = 4;            // tile slots per consumer warp (4: tiles are consumed in pairs)
constexpr int kGStride = 72;             // floats per (cta, chain) partial: gbeta[<=64], nll, pad
constexpr int kRedWarps = (kConsWarps + 1) / 2;   // rows of the two-stage CTA reduction scratch
constexpr int kXSeg = kConsWarps;        // segments of the cross-CTA reduction (each adds a contiguous range of CTAs)
constexpr int kXStride = 66;             // floats per segment / per chain of the packed outputs (3 per 16-byte word, <= 22 words)
constexpr int kXScratch = kXSeg * kXStride;           // offset of the tick's gradient scratch behind the segment sums
constexpr int kXRedFloats = kXScratch + 192;          // owner CTA: segment sums + gradient scratch for the tick
// partials: 16-byte words {v0, v1, v2, tag}, NW = ceil((8 KS + 1) / 3) per (chain, cta) row; output e = column e, e = 8 KS: nll.
// Layout [group][chain][cta][NW]: everything one owner needs from a range of CTAs is ONE contiguous block (one bulk copy).
constexpr int kBarBeta = 1, kBarCons = 2, kBarTick = 3, kBarSwept = 4;
constexpr int kBetaCopies = 4;           // replicas of the published beta (CTA c fetches replica c mod 4)
constexpr int kBetaWords = 8 * kStreamCT * 4;     // 16-byte words per replica: [k-step][chain][t]
constexpr int kConsThreads = kConsWarps * 32, kTopThreads = kStreamThreads;
constexpr int kMaxShards = 16;           // row-sharded handles: ranks that sweep disjoint rows of the same pass
// mailbox of a row-sharded handle: [round parity][chain][source rank][kGStride] {value, tag}
constexpr size_t kMailFloat2 = (size_t)2 * kMaxStreamChains * kMaxShards * kGStride;

B2_HD constexpr int stream_pitch(int KS) { return (KS % 2 == 0) ? 8 * KS + 8 : 8 * KS; }     // 8 * odd
// float offset of element (row r, column c) of a tile, r in [0, 16), c in [0, P)
B2_HD int stream_tile_index(int P, int r, int c) {
    const int pr = r & 7, hi = (r >> 3) & 1;
    const int chunk = (c >> 1) ^ (((pr >> 1) & 1) << 1);
    return pr * 2 * P + 4 * chunk + 2 * (c & 1) + hi;
}
B2_HD int stream_tile_y_index(int P, int r) { return kTileRows * P + (r & 7) * 2 + ((r >> 3) & 1); }
B2_HD constexpr int stream_tile_floats(int KS) { return kTileRows * stream_pitch(KS) + kTileRows; }
// ring slot stride: a drained slot doubles as the warp's reduction scratch (18 values x 32 lanes)
B2_HD constexpr int stream_slot_floats(int KS) { return stream_tile_floats(KS) > 18 * 32 ? stream_tile_floats(KS) : 18 * 32; }
B2_HD constexpr int stream_ks_for(int D) { return D <= 8 ? 1 : D <= 16 ? 2 : D <= 32 ? 4 : D <= 56 ? 7 : 8; }

// dbg: clock64 totals on CTA 0 -- [0] wait for betas (incl. the owners' ticks), [1] sweep, [2] CTA reduction +
// publish partial, [3] (unused), [4] poll + sum the partials, [5] total loop, [6] tick warp busy,
// [7] beta -> fragments, [8] tick: finish potential, [9] tick: state machine, [10] tick: publish beta
struct StreamSync { unsigned int abort_flag, pad_[3]; unsigned long long passes; unsigned long long dbg[16];
                    unsigned long long tick_sum[kMaxStreamChains], tick_max[kMaxStreamChains], tick_lap[kMaxStreamChains][4];
                    unsigned int pre_hit[kMaxStreamChains][4], pre_miss[kMaxStreamChains][4];
                    unsigned long long laps[32];
                    unsigned int peek_hit[kMaxStreamChains], peek_fallback[kMaxStreamChains], peek_mismatch[kMaxStreamChains];
                    unsigned long long cta_lap[160][4]; unsigned long long wake[4]; };     // per CTA: wait for betas, sweep, CTA reduction + publish, poll + sum     // -DB2_TICK_LAPS builds only: [i] cycles, [16 + i] occurrences   // per owner CTA: tick cycles, look-ahead hits

struct StreamParams {
    TickCfg cfg; FamilySpec fam; OutBufs out;
    int C, Dp, mode;                     // mode 0: run chains, 1: evaluate potential at z_in
    int num_groups;                      // ceil(C / kStreamCT)
    int max_passes;                      // single group only: pause after this many sweeps (0 = run to the end)
    ChainCtl* ctl; float* vecs;          // [C], [V_COUNT][C][Dp]
    float2* partial;                     // 16-byte words {v0, v1, v2, tag = round of the group}: [num_groups][kStreamCT][grid][NW]
    uint4* beta;                         // [kBetaCopies][num_groups][8 k-steps][kStreamCT][4] {b0, tag, b1, tag}: beta in MMA-fragment order
    StreamSync* sync;
    const float* z_in; float* u_out; float* g_out;    // mode 1
    const float* img;                    // tile image of (X, y), see above
    long long n_tiles;                   // tiles in the image
    int pad_rows;                        // zero rows appended to fill the last tile
    int stages;                          // slots in every consumer warp's private ring
    int vecs_in_smem;
    int dbg_sweep;                       // timing experiments only: 1 = copies without compute, 2 = compute without copies
    int dbg_warps;                       // ... only the first dbg_warps consumer warps compute
    long long spin_limit;
    // row sharding (config 5): this rank sweeps its own rows; per-chain likelihood sums are exchanged through the
    // ranks' mailboxes (peer stores over NVLink) and added in rank order, identically on every rank
    int shard_rank, shard_count; unsigned int epoch;      // epoch: launch number, part of the exchange tag
    float2* mail[kMaxShards];            // mailbox of every rank ([shard_rank] = this rank's own)
    float nll_local_const;               // added to this rank's nll before the exchange (poisson: local sum lgamma(y+1))
    unsigned int* trace;                 // debug only (B200NUTS_TRACE): host-mapped [grid][8] progress words, see B2_TRACE
    int no_prefetch;                     // debug only: skip the PRNG look-ahead
    int no_peek;                         // debug only: never publish the next position ahead of the tick (Tick::peek_next)
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
// bounded variant: gives up after `limit` clocks and raises *abort_flag (returns false)
// (abort codes: 1 tile copy never landed, 2 beta fetch timed out, 3 partial poll timed out, 4 a peer rank's sums never arrived)
B2_D bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, long long limit, unsigned int* abort_flag) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    unsigned int it = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++it & 255u) == 0u) {
            if (*(volatile unsigned int*)abort_flag) return false;
            if (clock64() - t0 > limit) { atomicCAS(abort_flag, 0u, 1u); return false; }
        }
    }
    return true;
}
// ... for the sweep: the abort flag's address and the limit are looked up on the slow path only (no registers held for them)
B2_D bool mbar_wait_p(uint64_t* bar, uint32_t parity, const unsigned int* const* syncp) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    unsigned int it = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++it & 255u) == 0u) {
            unsigned int* af = const_cast<unsigned int*>(*syncp);
            if (*(volatile unsigned int*)af) return false;
            if (clock64() - t0 > 2000000000LL) { atomicCAS(af, 0u, 1u); return false; }
        }
    }
    return true;
}
B2_D void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Polling loads of the on-device exchange (partials, betas): relaxed, GPU scope (one 8- / 16-byte access: value and tag arrive
// together).  System scope is only needed for the peer mailboxes of a row-sharded handle (ld_sys_v2).
// 16-byte asynchronous copy global -> shared, L2 only (.cg): the data never occupies a register
B2_D void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
B2_D void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
B2_D uint4 ld_volatile_v4(const uint4* p) {
    uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v;
}
B2_D float2 ld_volatile_v2(const float2* p) {
    float2 v; asm volatile("ld.relaxed.gpu.global.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory"); return v;
}
B2_D float2 ld_sys_v2(const float2* p) {
    float2 v; asm volatile("ld.volatile.global.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory"); return v;
}
B2_D void st_sys_v2(float2* p, float2 v) {      // one 8-byte store, visible to peer GPUs (value and tag travel together)
    asm volatile("st.volatile.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
B2_D unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
B2_D void red_release_add(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// tf32 split: hi keeps the top 19 bits (exactly what the tensor core consumes), lo = x - hi (exact in fp32;
// its own low bits are dropped by the MMA: relative error 2^-22 of x).
B2_D void tf32_split(uint32_t& hi, uint32_t& lo) {
    const uint32_t h = hi & 0xFFFFE000u;
    lo = __float_as_uint(__uint_as_float(hi) - __uint_as_float(h));
    hi = h;
}
B2_D void split4(const float (&x)[4], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { hi[i] = __float_as_uint(x[i]); tf32_split(hi[i], lo[i]); }
}
// (not volatile: pure functions of their operands, the compiler may interleave independent chains)
B2_D void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
B2_D void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two fp32 -> packed bf16x2 (round to nearest even): first -> bits [15:0] (the lower k index of an MMA fragment)
B2_D uint32_t pack_bf16(float first, float second) {
    uint32_t d; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(second), "f"(first)); return d;
}
// lo part of the tf32 split: x - trunc_tf32(x), exact in fp32
B2_D float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// packed fp32 pairs (Blackwell add/sub.f32x2 = one issue slot for two lanes of work)
B2_D unsigned long long pack2f(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
B2_D void unpack2f(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
B2_D void tf32_lo2(float x0, float x1, float& l0, float& l1) {
    const float h0 = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack2f(x0, x1)), "l"(pack2f(h0, h1)));
    unpack2f(d, l0, l1);
}
B2_D void add2f(float& a0, float& a1, float b0, float b1) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack2f(a0, a1)), "l"(pack2f(b0, b1)));
    unpack2f(d, a0, a1);
}
// Named barriers count whole warps: a warp that reaches one while diverged (e.g. after per-lane polling loops)
// would be counted once per divergent group.  Reconverge first.
template <int ID, int N> B2_D void bar_sync() { __syncwarp(); asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
template <int ID, int N> B2_D void bar_arrive() { __syncwarp(); asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
// Warp barrier that cannot be optimised away (see the gred reduction in the tick warp): a shuffle whose result is consumed.
B2_D void warp_sync_hard() {
    unsigned int x = __shfl_sync(0xFFFFFFFFu, threadIdx.x, 0);
    asm volatile("" ::"r"(x) : "memory");
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

// progress marker of (CTA, role): role 0 = consumer thread 0, 1 = tick lane 0; word = pass << 8 | stage
#define B2_TRACE(role, stage) do { if (p.trace) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p.trace + 32 * blockIdx.x + (role)), "r"((pass << 8) | (unsigned)(stage)) : "memory"); } } while (0)

#define B2_TRACE_LANES(stage) do { if (lane == 0) B2_TRACE(1, stage); if (lane == 16) B2_TRACE(4, stage); if (lane == 31) B2_TRACE(5, stage); } while (0)

// Shared-memory layout: the fixed-size regions come FIRST (compile-time offsets: their addresses cost no registers in the
// sweep), then the tile ring (its size depends on the ring depth), then the chain vectors.
B2_HD constexpr size_t stream_head_smem(bool multi_group) {
    size_t b = 0;
    b += (size_t)kConsWarps * kMaxStages * 8;                      // mbarriers
    b += (size_t)(multi_group ? 2 : 1) * kBetaWords * 16;          // staged beta (two passes with several groups), [k-step][chain][t] {b0, tag, b1, tag}
    b += (size_t)kXRedFloats * 4;                                  // cross-CTA reduction + tick scratch
    b += 64 * 4 + 64 * 4 + 256 + 256;                              // gred(+nll), flags, timers
    b += 128 + 128;                                                // gather: one mbarrier per consumer warp, their phase bits
    b += (size_t)kStreamCT * kXStride * 4;                         // this CTA's reduced outputs on their way to the packed partial
    return (b + 127) / 128 * 128;
}
B2_HD size_t stream_fixed_smem(int Dp, bool vecs_in_smem, bool multi_group) {
    size_t b = stream_head_smem(multi_group);
    if (vecs_in_smem) b += (size_t)V_COUNT * Dp * 4;
    return b + 128;
}
B2_HD size_t stream_smem_bytes(int KS, int Dp, int stages, bool vecs_in_smem, bool multi_group) {
    return stream_fixed_smem(Dp, vecs_in_smem, multi_group) + (size_t)kConsWarps * stages * stream_slot_floats(KS) * 4;
}

// One-time repack of the caller's (X, y) into the tile image (see the header comment).
static __global__ void k_stream_repack(const float* __restrict__ X, const float* __restrict__ y, long long N, int D, int P,
                                long long n_tiles, float* __restrict__ img) {
    const long long tile_floats = (long long)kTileRows * P + kTileRows;
    const long long per_tile = (long long)kTileRows * (P + 1);          // (row, column) pairs incl. the y "column" P
    const long long total = n_tiles * per_tile;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long tile = i / per_tile; const int o = (int)(i - tile * per_tile);
        const int r = o / (P + 1), c = o - r * (P + 1);
        const long long row = tile * kTileRows + r;
        float v = 0.0f;
        if (c < P) {
            if (row < N && c < D) v = X[row * D + c];
            img[tile * tile_floats + stream_tile_index(P, r, c)] = v;
        } else {
            if (row < N) v = y[row];
            img[tile * tile_floats + stream_tile_y_index(P, r)] = v;
        }
    }
}

// beta_c = s(z) * u for the next sweep (zero when the chain needs no gradient), published in the order of the
// consumers' B fragments: word (k-step kk, chain, t) = { beta[8kk+2t], tag, beta[8kk+2t+1], tag }.  Each 8-byte
// half is written atomically, so a reader that sees the tag also sees the value next to it.
B2_D void stream_publish_beta(const StreamParams& p, int chain, const float* zsrc, bool active, uint32_t tag) {
    const int grp = chain / kStreamCT, slot = chain % kStreamCT;
    const int lane = threadIdx.x & 31;
    const int kk = lane >> 2, t = lane & 3;              // 8 k-steps x 4 = 32 lanes
    const int d0 = 8 * kk + 2 * t;
    float b0 = 0.0f, b1 = 0.0f;
    if (p.fam.off_lambda < 0 && p.fam.gscale == SCALE_NONE) {         // plain GLM: scale 1 (1.0f * u == u exactly)
        if (active && d0 < p.fam.Dx) b0 = zsrc[p.fam.off_u + d0];
        if (active && d0 + 1 < p.fam.Dx) b1 = zsrc[p.fam.off_u + d0 + 1];
    } else {
        if (active && d0 < p.fam.Dx) b0 = glm_scale_at(p.fam, zsrc, d0) * zsrc[p.fam.off_u + d0];
        if (active && d0 + 1 < p.fam.Dx) b1 = glm_scale_at(p.fam, zsrc, d0 + 1) * zsrc[p.fam.off_u + d0 + 1];
    }
    const uint4 w = make_uint4(__float_as_uint(b0), tag, __float_as_uint(b1), tag);
#pragma unroll
    for (int r = 0; r < kBetaCopies; ++r) __stcg(p.beta + ((size_t)r * p.num_groups + grp) * kBetaWords + (kk * kStreamCT + slot) * 4 + t, w);
    __syncwarp();
}

// The tick warp's work between two sweeps, for the chain owned by this CTA, in two parts.
// stream_tick_critical: finish the potential from the reduced likelihood sums and -- when Tick::peek_next can tell where the
// chain goes next without the full bookkeeping -- publish the next beta right away (returns true; `u` and gz keep what the
// deferred part needs).  stream_tick_deferred: advance the NUTS state machine; with a single chain group the caller runs it
// AFTER it has staged the next pass for its own consumers, i.e. while the grid already sweeps X again.
__device__ __forceinline__ bool stream_tick_critical(const StreamParams& p, StreamTick& tk, const float* gred, float* gz, float* zpeek, float nll,
                                                     int cta, unsigned long long* tdbg, uint32_t next_tag, float& u, unsigned int (&peek_stat)[3]) {
    const int lane = threadIdx.x & 31;
    const long long t0 = (lane == 0) ? clock64() : 0ll;
    if (p.mode == 1) {
        const float* zin = p.z_in + (size_t)cta * p.cfg.D;
        glm_finish(p.fam, zin, nll, gred, u, gz);
        warp_sync_hard();
        if (lane == 0) p.u_out[cta] = u;
        for (int d = lane; d < p.cfg.D; d += 32) p.g_out[(size_t)cta * p.cfg.D + d] = gz[d];
        return false;
    }
    glm_finish(p.fam, tk.v(V_ZS), nll, gred, u, gz);
    warp_sync_hard();                    // gz is read across lanes below
#if defined(__CUDA_ARCH__)
    const bool early = !p.no_peek && tk.peek_next(u, gz, zpeek);
#else
    const bool early = false;            // (host compilation pass only: peek_next is device code)
#endif
    if (early) {
        warp_sync_hard();                // the betas are gathered across lanes from zpeek
        stream_publish_beta(p, cta, zpeek, true, next_tag);
        peek_stat[0] += 1u;
    } else if (tk.c.phase == PH_LEAF) peek_stat[1] += 1u;
    warp_sync_hard();
    if (lane == 0) tdbg[8] += (unsigned long long)(clock64() - t0);
    return early;
}

// Returns true when the chain needs no further gradient.
__device__ __forceinline__ bool stream_tick_deferred(const StreamParams& p, StreamTick& tk, const float* gz, const float* zpeek, float u, bool early,
                                                     unsigned long long* tdbg, unsigned int (&peek_stat)[3]) {
    const int lane = threadIdx.x & 31;
    const long long t0 = (lane == 0) ? clock64() : 0ll;
    tk.advance(u, gz);
    warp_sync_hard();                    // the next beta is gathered across lanes from V_ZS
    if (early) {                         // the position published ahead must be the one the state machine arrived at
        const float* zs = tk.v(V_ZS);
        bool same = tk.c.phase != PH_DONE;
        for (int d = lane; d < p.cfg.D; d += 32) same = same && (__float_as_uint(zs[d]) == __float_as_uint(zpeek[d]));
        if (!__all_sync(0xFFFFFFFFu, same)) peek_stat[2] += 1u;
    }
    if (lane == 0) tdbg[9] += (unsigned long long)(clock64() - t0);
    return tk.c.phase == PH_DONE;
}

// Byte offsets of the fixed shared-memory regions (see stream_head_smem)
template <bool MG> struct StreamSmem {
    static constexpr size_t kFull = 0;
    static constexpr size_t kBs = kFull + (size_t)kConsWarps * kMaxStages * 8;
    static constexpr size_t kXred = kBs + (size_t)(MG ? 2 : 1) * kBetaWords * 16;
    static constexpr size_t kGred = kXred + (size_t)kXRedFloats * 4;
    static constexpr size_t kPout = kGred + 64 * 4 + 64 * 4;
    static constexpr size_t kFlags = kPout + (size_t)kStreamCT * kXStride * 4;
    static constexpr size_t kTdbg = kFlags + 256;
    static constexpr size_t kGbar = kTdbg + 256;
    static constexpr size_t kGpar = kGbar + 128;
    static constexpr size_t kTiles = stream_head_smem(MG);
    static_assert(kGpar + 128 <= kTiles, "fixed regions overlap the tile ring");
};
// column of X behind n index `n` of backward N-tile `nt` (see the kernel's backward MMAs)
template <int KS> B2_D int stream_bwd_col(int nt, int n) {
    return ((KS & 1) && nt == KS - 1) ? 16 * (KS / 2) + n : 16 * (nt >> 1) + 2 * n + (nt & 1);
}

// Exchange, step 1 (every CTA, consumer warps): the 15 warps' accumulators sit in their reduction scratch slots; add them in
// warp order, publish the CTA's partial.  Out of line on purpose, like stream_gather below: the register needs of the exchange
// must not leak into the allocation of the sweep loop (ptxas spilled inside the MMA region when this was inlined).
template <int KS, bool MG>
__device__ __noinline__ void stream_reduce_publish(const StreamParams& p, int nst, int grp, uint32_t tag) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using L = StreamSmem<MG>;
    constexpr int SLOT_FLOATS = stream_slot_floats(KS);
    constexpr int NOUT = 8 * KS + 1, NW = (NOUT + 2) / 3;
    const float* tiles = (const float*)(smem_raw + L::kTiles);
    const int* flags = (const int*)(smem_raw + L::kFlags);
    float* pout = (float*)(smem_raw + L::kPout);
    const int ctid = threadIdx.x, cta = blockIdx.x;
    bar_sync<kBarCons, kConsThreads>();
    {
        // thread i adds scratch word i = (value k, lane l) of the 15 warps in warp order (conflict-free: consecutive threads read
        // consecutive words) and knows where the sum belongs: value k = 2 nt + e of lane l = 4 c + (n >> 1) is (chain c, column
        // bwd_col(nt, n)) with n = 2 (l & 3) + e; values 16 + e of lanes 0..3 are the nll of chains 2 l + e (output 8 KS).
        int rs[kConsWarps];
#pragma unroll
        for (int w = 0; w < kConsWarps; ++w) rs[w] = (w * nst + flags[16 + w]) * SLOT_FLOATS;
        for (int i = ctid; i < 18 * 32; i += kConsThreads) {
            const int k = i >> 5, l = i & 31;
            int c = -1, e = 8 * KS;
            if (k < 2 * KS) { c = l >> 2; e = stream_bwd_col<KS>(k >> 1, 2 * (l & 3) + (k & 1)); }
            else if (k >= 16 && l < 4) c = 2 * l + (k & 1);
            if (c >= 0) {
                float a = 0.0f;
#pragma unroll
                for (int w = 0; w < kConsWarps; ++w) a += tiles[rs[w] + i];
                pout[c * kXStride + e] = a;
            }
        }
    }
    bar_sync<kBarCons, kConsThreads>();          // every scratch slot has been read (the caller refills them); the CTA's sums are in pout
    // publish: three outputs and the tag per 16-byte word (one store: a reader that sees the tag sees the values next to it)
    if (ctid < kStreamCT * NW) {
        const int c = ctid / NW, w = ctid - c * NW;
        const float* src = pout + c * kXStride + 3 * w;
        __stcg(reinterpret_cast<float4*>(p.partial) + (((size_t)grp * kStreamCT + c) * gridDim.x + cta) * NW + w,
               make_float4(src[0], src[1], src[2], __uint_as_float(tag)));
    }
}

// Exchange, step 2 (chain owners, consumer warps): poll the partials of all CTAs, add them in fixed order, leave the sums in gred.
template <int KS, bool MG>
__device__ __noinline__ void stream_gather(const StreamParams& p, int ggrp, uint32_t gtag) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using L = StreamSmem<MG>;
    constexpr int NOUT = 8 * KS + 1, NW = (NOUT + 2) / 3;
    float* xred = (float*)(smem_raw + L::kXred);
    float* gred = (float*)(smem_raw + L::kGred);
    unsigned long long* tdbg = (unsigned long long*)(smem_raw + L::kTdbg);
    const int ctid = threadIdx.x, cta = blockIdx.x, G = gridDim.x;
    const bool dbg = (ctid == 0);
    StreamSync* sy = p.sync;
    // kConsWarps segments x NW words: thread (seg, w) adds word w of its segment's CTAs in ascending order.  All loads of a
    // batch are in flight together (L2 latency overlapped); only the entries that were still stale are polled again
    // (the owner SM's load path moves about one 32-byte sector per cycle: a full round costs thousands of cycles).
    // (the same ranges and the same order of additions as stream_gather_land: a chain's sums must not depend on how many
    //  chain groups the handle has; five loads in flight per thread keep this function's register footprint small)
    constexpr int SEGS = kConsWarps;
    constexpr int NB = 5;
    const int seg = ctid / NW, w = ctid - seg * NW;
    if (seg < SEGS) {
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
        const int g0 = G * seg / SEGS, g1 = G * (seg + 1) / SEGS;
        const uint4* src = reinterpret_cast<const uint4*>(p.partial) + ((size_t)ggrp * kStreamCT + (MG ? cta % kStreamCT : cta)) * G * NW + w;
        constexpr size_t cta_stride = NW;
        const long long t_w = clock64();
        for (int gg = g0; gg < g1; gg += NB) {
            uint4 v[NB];
            const int nb = (g1 - gg < NB) ? (g1 - gg) : NB;
            const unsigned want = (1u << nb) - 1u;
            unsigned ready = 0u;
            while (true) {
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                    if (k < nb && !((ready >> k) & 1u)) {
                        v[k] = ld_volatile_v4(src + (size_t)(gg + k) * cta_stride);
                        if (v[k].w == gtag) ready |= 1u << k;
                    }
                }
                if (dbg) tdbg[3] += 1ull;                 // (poll rounds of thread 0)
                if (ready == want) break;
                if (ld_acquire(&sy->abort_flag)) break;
                if (clock64() - t_w > p.spin_limit) { atomicCAS(&sy->abort_flag, 0u, 3u); break; }
            }
#pragma unroll
            for (int k = 0; k < NB; ++k)
                if (k < nb) { a0 += __uint_as_float(v[k].x); a1 += __uint_as_float(v[k].y); a2 += __uint_as_float(v[k].z); }
        }
        float* dst = xred + seg * kXStride + 3 * w;
        dst[0] = a0; dst[1] = a1; dst[2] = a2;
    }
    bar_sync<kBarCons, kConsThreads>();      // the segment sums are in xred: join them in segment order
    if (ctid < NOUT) {
        float a = xred[ctid];
#pragma unroll
        for (int sgm = 1; sgm < SEGS; ++sgm) a += xred[sgm * kXStride + ctid];
        gred[ctid < 8 * KS ? ctid : 64] = a;
    }
    __threadfence_block();
}

// Single chain group: consumer warp w fetches the rows of ITS range of CTAs (one contiguous block of the owner's chain) with ONE
// bulk copy into the reduction scratch slot it has just drained (the owner CTAs refill those slots after the gather), checks
// the tags in shared memory (a stale block is simply fetched again), adds the rows in CTA order; the 15 ranges are joined in
// warp order.  The owner SM's load path moves only about one 32-byte sector per cycle for per-thread loads: this is ~3x faster.
template <int KS>
__device__ __noinline__ void stream_gather_land(const StreamParams& p, uint32_t gtag, float* land) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using L = StreamSmem<false>;
    constexpr int NOUT = 8 * KS + 1, NW = (NOUT + 2) / 3;
    constexpr int ROWS = stream_slot_floats(KS) * 4 / (NW * 16);       // rows of NW words that fit into a scratch slot
    static_assert(NW <= 32 && ROWS >= 1, "gather: a row must fit a warp and a slot");
    float* xred = (float*)(smem_raw + L::kXred);
    float* gred = (float*)(smem_raw + L::kGred);
    unsigned long long* tdbg = (unsigned long long*)(smem_raw + L::kTdbg);
    const int ctid = threadIdx.x, lane = ctid & 31, cw = ctid >> 5, cta = blockIdx.x, G = gridDim.x;
    uint64_t* gbar = (uint64_t*)(smem_raw + L::kGbar) + cw;
    unsigned int* gpar = (unsigned int*)(smem_raw + L::kGpar) + cw;
    const bool dbg = (ctid == 0);
    const int g0 = G * cw / kConsWarps, g1 = G * (cw + 1) / kConsWarps;
    const uint4* rows = reinterpret_cast<const uint4*>(p.partial) + (size_t)cta * G * NW;      // [cta][NW] of this owner's chain
    const uint4* land4 = reinterpret_cast<const uint4*>(land);
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    unsigned int par = *gpar;
    const long long t_w = clock64();
    for (int gg = g0; gg < g1; gg += ROWS) {
        const int nrow = (g1 - gg < ROWS) ? (g1 - gg) : ROWS;
        const uint32_t bytes = (uint32_t)nrow * NW * 16u;
        while (true) {
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy accesses of the slot come first
                mbar_expect_tx(gbar, bytes);
                bulk_g2s(land, rows + (size_t)gg * NW, bytes, gbar);
            }
            bool alive = true;
            {   // bounded wait (every lane: the data must be visible to each of them)
                unsigned int it = 0;
                while (!mbar_try_wait(gbar, par)) {
                    if ((++it & 255u) == 0u && (ld_acquire(&p.sync->abort_flag) || clock64() - t_w > p.spin_limit)) { alive = false; break; }
                }
            }
            par ^= 1u;
            bool ok = true;
            for (int i = lane; i < nrow * NW; i += 32) ok = ok && (reinterpret_cast<const volatile unsigned int*>(land4 + i)[3] == gtag);
            if (dbg) tdbg[3] += 1ull;                 // (poll rounds of thread 0)
            const bool all_ok = __all_sync(0xFFFFFFFFu, ok);
            if (all_ok) break;
            // (warp-uniform exits)
            bool give_up = !alive || ld_acquire(&p.sync->abort_flag) != 0u;
            if (clock64() - t_w > p.spin_limit) { atomicCAS(&p.sync->abort_flag, 0u, 3u); give_up = true; }
            if (__any_sync(0xFFFFFFFFu, give_up)) break;
        }
        if (lane < NW)
            for (int r = 0; r < nrow; ++r) {
                const uint4 v = land4[r * NW + lane];
                a0 += __uint_as_float(v.x); a1 += __uint_as_float(v.y); a2 += __uint_as_float(v.z);
            }
        __syncwarp();                                 // (a second block of rows lands in the same slot)
    }
    if (lane == 0) *gpar = par;
    if (lane < NW) { float* dst = xred + cw * kXStride + 3 * lane; dst[0] = a0; dst[1] = a1; dst[2] = a2; }
    bar_sync<kBarCons, kConsThreads>();      // the range sums are in xred: join them in warp order
    if (ctid < NOUT) {
        float a = xred[ctid];
#pragma unroll
        for (int sgm = 1; sgm < kConsWarps; ++sgm) a += xred[sgm * kXStride + ctid];
        gred[ctid < 8 * KS ? ctid : 64] = a;
    }
    __threadfence_block();
    if (ctid == 0) tdbg[19] = (unsigned long long)clock64();
    bar_arrive<kBarTick, kStreamThreads>();  // the tick warp takes over at once (the caller's slot refill is off the critical path)
}

struct StreamOne { static constexpr int value = 1; };
struct StreamTwo { static constexpr int value = 2; };

// KS = columns / 8 of the padded tile (compile time so every fragment stays in registers), LIK = likelihood.
// MG = several chain groups (more than 8 chains); with MG = false everything group-related folds to constants.
// (lap state lives in shared memory -- tdbg[14] last lap, tdbg[15] start -- so that it costs no registers in the sweep)
#define B2_DBG_LAP(k) do { if (dbg) { const unsigned long long t_now = (unsigned long long)clock64(); tdbg[k] += t_now - tdbg[14]; tdbg[14] = t_now; } } while (0)

// One pass of one consumer warp: beta fragments, the sweep over the warp's tiles, accumulators into the warp's reduction
// scratch slot.  Out of line on purpose: the sweep gets a register allocation of its own (inlined into the kernel, ptxas let
// unrelated code -- the exchange, the tick warp -- push spills into the MMA region).  `ring` = slot | parity << 2 of the next
// tile to consume; returns the new ring position | red_slot << 3 | (red_tile + 1) << 5.
template <int KS, int LIK, bool MG>
__device__ __forceinline__ uint32_t stream_sweep(const StreamParams& p, uint32_t ring, unsigned int pass, int n_tiles, unsigned int first_tile) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using L = StreamSmem<MG>;
    constexpr int P = stream_pitch(KS);            // row pitch of a tile, floats
    constexpr int TILE_FLOATS = stream_tile_floats(KS);   // floats moved per tile
    constexpr int SLOT_FLOATS = stream_slot_floats(KS);   // ring slot stride
    constexpr int NCH = KS / 2;                    // backward: 16-column chunks (two 8-column MMAs per 128-bit load)
    constexpr bool ODD = (KS & 1) != 0;            // ... + one 8-column chunk (64-bit loads) when KS is odd
    // (volatile reads: everything derived from them is recomputed per pass and dies with the sweep -- nothing the sweep needs
    //  stays live across the exchange that follows it, and nothing of the exchange reaches into the sweep's registers)
    int tid, cta;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(cta));
    const int cw = tid >> 5, lane = tid & 31, ctid = tid;
    const int nst = p.stages;
    uint64_t* full = (uint64_t*)(smem_raw + L::kFull);
    const uint4* bs = (const uint4*)(smem_raw + L::kBs);
    int* flags = (int*)(smem_raw + L::kFlags);
    unsigned long long* tdbg = (unsigned long long*)(smem_raw + L::kTdbg);
    float* tiles = (float*)(smem_raw + L::kTiles);
    // (StreamSync starts with abort_flag: &p.sync, read through the parameter, is all the bounded waits need)
    const unsigned int* const* syncp = reinterpret_cast<const unsigned int* const*>(&p.sync);
#ifdef B2_STREAM_SWEEP_KNOBS
    constexpr bool kKnobs = true;                  // timing experiments (B200NUTS_DEBUG_SWEEP / _WARPS): copies only, compute only
#else
    constexpr bool kKnobs = false;
#endif
    const bool dbg = (ctid == 0);
    const int g = lane >> 2, t = lane & 3;           // mma.sync fragment coordinates (groupID, threadID_in_group)
    // float offsets inside a tile (stream_tile_index): forward lane (g, t) reads pair row g, columns
    // 8kk + 2t, +1; backward lane (g, t) reads pair row 4ks + t, columns 16j + 2g, +1 (or 16*NCH + g)
    const int off_f = g * 2 * P + 4 * (t ^ (((g >> 1) & 1) << 1));
    const int off_b = t * 2 * P + 4 * (g ^ (((t >> 1) & 1) << 1));
    const int off_b1 = t * 2 * P + 4 * (8 * NCH + ((g >> 1) ^ (((t >> 1) & 1) << 1))) + 2 * (g & 1);
    const int src_ks0 = (t << 2) | (g >> 1), src_ks1 = ((4 + t) << 2) | (g >> 1);   // shuffle sources, see unit()
    uint32_t bhi[KS][2], bbf[KS][2];                 // beta as B fragments of the forward MMAs: tf32 (raw fp32 bits; the
                                                     // tensor core reads the top 19) and bf16x2 {hi part, lo part}
    float gacc[KS][4];                               // gbeta^T in C-fragment layout: chain g, columns bwd_col(nt, 2t / 2t+1);
                                                     // [0], [1] = r_hi part, [2], [3] = r_lo part
    float nll[2] = {0.0f, 0.0f};                     // loss of chains 2t, 2t+1 over this lane's rows

    const int n_mine = (n_tiles > cw) ? (n_tiles - cw + kConsWarps - 1) / kConsWarps : 0;
    float* my_tiles = tiles + (size_t)cw * nst * SLOT_FLOATS;
    uint64_t* my_full = full + cw * kMaxStages;
    const unsigned int my_first = first_tile + (unsigned int)cw;            // first tile of this warp (the image address is formed at issue time)
    auto issue = [&](int slot, int j) {              // one lane: tile j of this warp -> slot
        mbar_expect_tx(&my_full[slot], TILE_FLOATS * 4u);
        bulk_g2s(my_tiles + (size_t)slot * SLOT_FLOATS, p.img + (size_t)(my_first + (unsigned int)j * kConsWarps) * TILE_FLOATS, TILE_FLOATS * 4u, &my_full[slot]);
    };
    int slot = (int)(ring & 3u); uint32_t parity = (ring >> 2) & 1u;
    const bool pairs = (nst == 4);                   // consume two tiles at a time (two independent instruction streams)
    // A warp with no tiles still owns its (empty) slot 0 as reduction scratch; the others use the slot they drained last.
    int red_slot = 0, red_tile = -1;

    // NG 16-row tiles at once (NG = 1 or 2; with 2 the two tiles' instruction streams are independent and
    // interleave).  Products are split-precision: x = xh + xl (xh = the 19 bits a TF32 MMA reads), likewise beta
    // and r.  hi*hi runs as TF32 MMAs; the three small cross terms run as BF16 MMAs (k = 16), which is exact
    // enough because each is already ~2^-11 of the main term (total relative error ~1e-6, fp32 accumulate).
    auto unit = [&](auto ng_tag, const float* xa, const float* xb) {
        constexpr int NG = decltype(ng_tag)::value;
        constexpr int NC = (NG == 2) ? 1 : 2;        // accumulator chains per MMA kind and tile
        const float* xt[2] = {xa, xb};
        // ---- forward: logits
        float acc[NG][2 * NC][4];
#pragma unroll
        for (int gi = 0; gi < NG; ++gi)
#pragma unroll
            for (int k = 0; k < 2 * NC; ++k) { acc[gi][k][0] = 0.0f; acc[gi][k][1] = 0.0f; acc[gi][k][2] = 0.0f; acc[gi][k][3] = 0.0f; }
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
            for (int gi = 0; gi < NG; ++gi) {
                // (row g, col c), (row g+8, c), (row g, c+1), (row g+8, c+1) with c = 8kk + 2t
                const float4 v = *reinterpret_cast<const float4*>(xt[gi] + off_f + 16 * kk);
                const uint32_t ar[4] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
                mma_tf32(acc[gi][kk % NC], ar, bhi[kk][0], bhi[kk][1]);                      // xh * bh
                // k = (2t, 2t+1): xl at columns (c, c+1);  k = (2t+8, 2t+9): x at columns (c, c+1)
                float lx, ly, lz, lw;
                tf32_lo2(v.x, v.y, lx, ly); tf32_lo2(v.z, v.w, lz, lw);
                const uint32_t ab[4] = {pack_bf16(lx, lz), pack_bf16(ly, lw), pack_bf16(v.x, v.z), pack_bf16(v.y, v.w)};
                mma_bf16(acc[gi][NC + kk % NC], ab, bbf[kk][0], bbf[kk][1]);                // xl * bh + x * bl
            }
        }
        // ---- link: c0 (row g, chain 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1)
        float dl[NG][4];
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
            const float2 yy = *reinterpret_cast<const float2*>(xt[gi] + kTileRows * P + 2 * g);    // y[row g], y[row g+8]
            float ls[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float eta = acc[gi][0][i] + acc[gi][NC][i];
                if (NC == 2) eta = (acc[gi][1][i] + acc[gi][3][i]) + eta;
                link_fn<LIK>(eta, (i < 2) ? yy.x : yy.y, ls[i], dl[gi][i]);
            }
            nll[0] += ls[0] + ls[2]; nll[1] += ls[1] + ls[3];
        }
        // ---- residuals -> A fragments.  k-step ks covers pair rows 4ks..4ks+3; lane (g, t) needs
        //      r[row 4ks+t][chain g] and r[row 4ks+t+8][chain g], held by lane (4ks+t, g>>1) of the C layout.
        //      TF32: stacked A = [r_hi ; r_lo]^T: a0 = r_hi(low row), a1 = r_lo(low), a2 = r_hi(high), a3 = r_lo(high)
        //      BF16 (k = 16 = both k-steps): a0 = {r(ks0 low), r(ks0 high)}, a2 = {r(ks1 low), r(ks1 high)}, a1 = a3 = 0
        uint32_t ra[NG][2][4], rb[NG][4];
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
            rb[gi][1] = 0u; rb[gi][3] = 0u;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int src = ks ? src_ks1 : src_ks0;
                const float e0 = __shfl_sync(0xFFFFFFFFu, dl[gi][0], src), e1 = __shfl_sync(0xFFFFFFFFu, dl[gi][1], src);
                const float o0 = __shfl_sync(0xFFFFFFFFu, dl[gi][2], src), o1 = __shfl_sync(0xFFFFFFFFu, dl[gi][3], src);
                const float lo_row = (g & 1) ? e1 : e0, hi_row = (g & 1) ? o1 : o0;
                ra[gi][ks][0] = __float_as_uint(lo_row); ra[gi][ks][1] = __float_as_uint(tf32_lo(lo_row));
                ra[gi][ks][2] = __float_as_uint(hi_row); ra[gi][ks][3] = __float_as_uint(tf32_lo(hi_row));
                rb[gi][2 * ks] = pack_bf16(lo_row, hi_row);
            }
        }
        // ---- backward: gbeta^T[chain][col] += sum_rows r[row][chain] * x[row][col], 16 columns (2 MMAs wide) at a time
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            float ga[NG][2][4];
#pragma unroll
            for (int gi = 0; gi < NG; ++gi) {
#pragma unroll
                for (int k = 0; k < 2; ++k) { ga[gi][k][0] = 0.0f; ga[gi][k][1] = 0.0f; ga[gi][k][2] = 0.0f; ga[gi][k][3] = 0.0f; }
                float xl[2][4];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    // col c = 16j+2g: rows (4ks+t, +8); col c+1: rows (4ks+t, +8)
                    const float4 v = *reinterpret_cast<const float4*>(xt[gi] + off_b + ks * 8 * P + 32 * j);
                    mma_tf32(ga[gi][0], ra[gi][ks], __float_as_uint(v.x), __float_as_uint(v.y));   // [r_hi ; r_lo] * xh
                    mma_tf32(ga[gi][1], ra[gi][ks], __float_as_uint(v.z), __float_as_uint(v.w));
                    tf32_lo2(v.x, v.y, xl[ks][0], xl[ks][1]); tf32_lo2(v.z, v.w, xl[ks][2], xl[ks][3]);
                }
                mma_bf16(ga[gi][0], rb[gi], pack_bf16(xl[0][0], xl[0][1]), pack_bf16(xl[1][0], xl[1][1]));   // r * xl, 16 rows
                mma_bf16(ga[gi][1], rb[gi], pack_bf16(xl[0][2], xl[0][3]), pack_bf16(xl[1][2], xl[1][3]));
            }
#pragma unroll
            for (int gi = 0; gi < NG; ++gi)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    add2f(gacc[2 * j + k][0], gacc[2 * j + k][1], ga[gi][k][0], ga[gi][k][1]);
                    add2f(gacc[2 * j + k][2], gacc[2 * j + k][3], ga[gi][k][2], ga[gi][k][3]);
                }
        }
        if (ODD) {
            float ga[NG][4];
#pragma unroll
            for (int gi = 0; gi < NG; ++gi) {
                ga[gi][0] = 0.0f; ga[gi][1] = 0.0f; ga[gi][2] = 0.0f; ga[gi][3] = 0.0f;
                float xl[2][2];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const float2 v = *reinterpret_cast<const float2*>(xt[gi] + off_b1 + ks * 8 * P);
                    mma_tf32(ga[gi], ra[gi][ks], __float_as_uint(v.x), __float_as_uint(v.y));
                    tf32_lo2(v.x, v.y, xl[ks][0], xl[ks][1]);
                }
                mma_bf16(ga[gi], rb[gi], pack_bf16(xl[0][0], xl[0][1]), pack_bf16(xl[1][0], xl[1][1]));
            }
#pragma unroll
            for (int gi = 0; gi < NG; ++gi) {
                add2f(gacc[KS - 1][0], gacc[KS - 1][1], ga[gi][0], ga[gi][1]);
                add2f(gacc[KS - 1][2], gacc[KS - 1][3], ga[gi][2], ga[gi][3]);
            }
        }
    };


#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {                                   // beta -> B fragments (k = column, n = chain)
        const uint4 w = bs[(size_t)(MG ? (pass & 1u) : 0u) * kBetaWords + (kk * kStreamCT + g) * 4 + t];
        bhi[kk][0] = w.x; bhi[kk][1] = w.y;                             // staged pre-split by the tick warp
        bbf[kk][0] = w.z;                                               // hi parts: pairs with xl (k = 2t, 2t+1)
        bbf[kk][1] = w.w;                                               // lo parts: pairs with x  (k = 2t+8, 2t+9)
    }
#pragma unroll
    for (int nt = 0; nt < KS; ++nt) { gacc[nt][0] = 0.0f; gacc[nt][1] = 0.0f; gacc[nt][2] = 0.0f; gacc[nt][3] = 0.0f; }
    nll[0] = 0.0f; nll[1] = 0.0f;
    if (ctid == 0) B2_DBG_LAP(7);

    // ---- sweep: consume this warp's tiles in order (two at a time when the ring has 4 slots); a finished
    //      slot is refilled at once with the tile `nst` positions further down the (cyclic) sequence
    {
        int j = 0, jn = nst % (n_mine > 0 ? n_mine : 1);                // jn = (j + nst) mod n_mine
        while (j < n_mine) {
            const int s0 = slot; const uint32_t p0 = parity;
            if (++slot == nst) { slot = 0; parity ^= 1u; }
            if (pairs && j + 1 < n_mine) {
                const int s1 = slot; const uint32_t p1 = parity;
                if (++slot == nst) { slot = 0; parity ^= 1u; }
                if (!kKnobs || p.dbg_sweep != 2) { mbar_wait_p(&my_full[s0], p0, syncp); mbar_wait_p(&my_full[s1], p1, syncp); }
                if (!kKnobs || (p.dbg_sweep != 1 && cw < p.dbg_warps)) unit(StreamTwo{}, my_tiles + (size_t)s0 * SLOT_FLOATS, my_tiles + (size_t)s1 * SLOT_FLOATS);
                __syncwarp();
                const bool last = (j + 2 >= n_mine);
                if (lane == 0 && (!kKnobs || p.dbg_sweep != 2)) {
                    issue(s0, jn); if (++jn == n_mine) jn = 0;
                    if (!last) issue(s1, jn);
                }
                if (last) { red_slot = s1; red_tile = jn; }        // refilled after the CTA reduction below
                if (++jn == n_mine) jn = 0;
                j += 2;
            } else {
                if (!kKnobs || p.dbg_sweep != 2) mbar_wait_p(&my_full[s0], p0, syncp);
                if (!kKnobs || (p.dbg_sweep != 1 && cw < p.dbg_warps)) unit(StreamOne{}, my_tiles + (size_t)s0 * SLOT_FLOATS, nullptr);
                __syncwarp();
                const bool last = (j + 1 >= n_mine);
                if (lane == 0 && (!kKnobs || p.dbg_sweep != 2) && !last) issue(s0, jn);
                if (last) { red_slot = s0; red_tile = jn; }
                if (++jn == n_mine) jn = 0;
                j += 1;
            }
        }
    }
    // ---- reduce warps -> CTA through the ring slot every warp drained last ([value][lane] floats, conflict
    //      free), one barrier, fixed order => bit-reproducible; then publish {value, tag} pairs.
    {
        float val[16];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            val[2 * nt] = (nt < KS) ? gacc[nt < KS ? nt : 0][0] + gacc[nt < KS ? nt : 0][2] : 0.0f;
            val[2 * nt + 1] = (nt < KS) ? gacc[nt < KS ? nt : 0][1] + gacc[nt < KS ? nt : 0][3] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            nll[k] += __shfl_xor_sync(0xFFFFFFFFu, nll[k], 4);
            nll[k] += __shfl_xor_sync(0xFFFFFFFFu, nll[k], 8);
            nll[k] += __shfl_xor_sync(0xFFFFFFFFu, nll[k], 16);
        }
        float* scr = my_tiles + (size_t)red_slot * SLOT_FLOATS;      // >= 18 * 32 floats in every configuration
#pragma unroll
        for (int k = 0; k < 16; ++k) scr[k * 32 + lane] = val[k];
        scr[16 * 32 + lane] = nll[0]; scr[17 * 32 + lane] = nll[1];
        if (lane == 0) flags[16 + cw] = red_slot;
    }
    return (uint32_t)slot | (parity << 2) | ((uint32_t)red_slot << 3) | ((uint32_t)(red_tile + 1) << 5);
}

template <int KS, int LIK, bool MG>
__global__ void __launch_bounds__(kStreamThreads, 1) stream_engine_kernel(const __grid_constant__ StreamParams p) {
    const int NGRP = MG ? p.num_groups : 1;        // chain groups
    // With >= 3 groups the owners' consumers gather the partials of their chain one pass LATE (at the end of the next pass,
    // when every CTA has long published them), so no CTA ever waits for the slowest CTA of a pass.
    const bool DEFER = MG && NGRP >= 3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int P = stream_pitch(KS);            // row pitch of a tile, floats
    constexpr int TILE_FLOATS = stream_tile_floats(KS);   // floats moved per tile
    constexpr int SLOT_FLOATS = stream_slot_floats(KS);   // ring slot stride
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    const int nst = p.stages;

    // ---- carve shared memory
    using L = StreamSmem<MG>;
    uint64_t* full = (uint64_t*)(smem_raw + L::kFull);
    uint4* bs = (uint4*)(smem_raw + L::kBs);
    float* xred = (float*)(smem_raw + L::kXred);
    float* gred = (float*)(smem_raw + L::kGred);
    float* pout = (float*)(smem_raw + L::kPout);
    int* flags = (int*)(smem_raw + L::kFlags);       // per pass parity: [0..1] 0 go / 1 all chains done / 2 abort, [2..3] chain group, [4..5] its round;
                                                     // [8] last pass begun by the consumers, [16..31] reduction slot of warp w, [32..51] rounds per group
    unsigned long long* tdbg = (unsigned long long*)(smem_raw + L::kTdbg);
    float* tiles = (float*)(smem_raw + L::kTiles);
    float* cvecs = tiles + (size_t)kConsWarps * nst * SLOT_FLOATS;

    // ---- this CTA's slice of tiles
    const long long t_begin_tile = p.n_tiles * cta / G, t_end_tile = p.n_tiles * (cta + 1) / G;
    const int n_tiles = (int)(t_end_tile - t_begin_tile);
    const bool is_tick = cta < p.C;

    // ---- one-time setup: init barriers, stage the chain
    if (tid == 0) {
        for (int i = 0; i < kConsWarps * kMaxStages; ++i) mbar_init(&full[i], 1);
        for (int i = 0; i < kConsWarps; ++i) { mbar_init((uint64_t*)(smem_raw + L::kGbar) + i, 1); ((unsigned int*)(smem_raw + L::kGpar))[i] = 0u; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < 32; ++i) tdbg[i] = 0ull;
        for (int i = 0; i < 64; ++i) flags[i] = 0;
        for (int i = 0; i < 128; ++i) gred[i] = 0.0f;   // (columns 8 KS .. 63 are never written afterwards)
        for (int i = 0; i < kStreamCT * kXStride; ++i) pout[i] = 0.0f;   // (nor are the pad outputs of a chain's last word)
        flags[8] = -1;                               // last pass the consumers have begun
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
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    StreamSync* sy = p.sync;
    unsigned int pass = 0;

    // =============================================================== tick / fetch warp
    if (warp == kConsWarps) {
        const bool dbg = (cta == 0 && lane == 0);
        unsigned long long tk_sum = 0ull, tk_max = 0ull;
        unsigned int peek_stat[3] = {0u, 0u, 0u};      // early publishes, fall-backs, mismatches (must stay 0)
        bool deferred = false; float def_u = 0.0f;     // a tick whose next position is published but whose bookkeeping is pending
        unsigned long long* tlap = tdbg + 8;          // [0..1] finish / advance, [2] publish, [3] gred sum
        float loss0, dl0;
        link_fn<LIK>(0.0f, 0.0f, loss0, dl0);        // what every zero pad row adds to a chain's nll
        const float pad_nll = (float)p.pad_rows * loss0;
        // the chain's control block lives in this warp's registers for the whole launch (every lane holds a copy)
        ChainCtl c;
        if (is_tick && p.mode == 0) c = p.ctl[cta]; else memset(&c, 0, sizeof(c));
        StreamTick tk{p.cfg, c, cv, p.out, cta, p.C};
        bool chain_done = !is_tick || (p.mode == 0 && c.phase == PH_DONE);
        // Passes rotate over the chain groups (8 chains each) that still have work; group g's r-th sweep uses the betas its
        // owners published with tag r.  With several groups the owners' ticks overlap with the other groups' sweeps.
        const int my_group = MG ? cta / kStreamCT : 0;
        unsigned int* rounds = (unsigned int*)(flags + 32);       // sweeps staged so far, per group (this warp's private copy)
        volatile int* started = flags + 8;
        uint32_t done_mask = 0u;                     // groups whose chains have all finished
        uint32_t my_round = 0u;                      // round of my group's sweep that is waiting for its tick
        bool pending = false;                        // ... and whether there is one
        int pending_pass = 0;                        // ... and the pass that swept it
        int cur = 0, qpass = 0, status = 0;
        if (is_tick) {                               // prologue: the first beta
            const float* zsrc = (p.mode == 0) ? cv.v(V_ZS) : (p.z_in + (size_t)cta * p.cfg.D);
            stream_publish_beta(p, cta, zsrc, !chain_done, 1u | (chain_done ? 0x80000000u : 0u));
        }
        const int my_chain = (lane >> 2) & 7;         // chain slot of this lane's words (word w = lane + 32 kk)
#ifdef B2_TICK_LAPS
        if (cta == 0 && lane == 0) for (int k = 0; k < 34; ++k) b2_lap_store()[k] = 0ull;
        __syncwarp();
#endif

        uint32_t need = 0u;                          // round (= beta tag) of the sweep being staged
        int swept_sync = 0;                          // single group: last pass whose predecessor's sweep this warp has waited for
        int grp = -1;
        // Stage the next pass for this CTA's consumers: fetch the betas of group `grp` (tag need), put them into the pass's
        // staging buffer and release the consumers.  Returns 0 = staged, 1 = the group has finished (no pass), 2 = stop.
        auto stage = [&]() -> int {
            B2_TRACE_LANES(1);
            // pass-bounded launch (single group): every CTA stops at the same pass; the owners have ticked, so each chain
            // sits in front of its next gradient exactly as at the start of a launch
            if (!MG && p.max_passes > 0 && qpass >= p.max_passes && status == 0) status = 1;
            // the staging buffer and the barrier phase of pass qpass were last used by pass qpass - 2: the consumers must
            // have begun pass qpass - 1 before they are reused
            {
                const long long t_w = clock64();
                while (true) {                       // (lane 0's view decides: every lane runs the same number of iterations)
                    const int ready = __shfl_sync(0xFFFFFFFFu, (*started >= qpass - 1) ? 1 : 0, 0);
                    if (ready) break;
                    __nanosleep(256);                // (a spinning warp would take issue slots from the consumers of its scheduler)
                    bool give_up = ld_acquire(&sy->abort_flag) != 0u;
                    if (clock64() - t_w > 4 * p.spin_limit) { atomicCAS(&sy->abort_flag, 0u, 2u); give_up = true; }
                    if (__any_sync(0xFFFFFFFFu, give_up)) { status = 2; break; }
                }
            }
            uint4* bsq = bs + (size_t)(MG ? (qpass & 1) : 0) * kBetaWords;
            need = 0u;
            // (single group: no beta can arrive before this CTA's own consumers have finished the previous sweep -- sleep on a
            //  barrier until then instead of polling the L2 through the whole sweep)
            if (!MG && qpass > swept_sync) { bar_sync<kBarSwept, kStreamThreads>(); swept_sync = qpass; }    // (once per sweep: stage() can run twice for a pass)
            if (status == 0) {
                // fetch the betas of group `grp` for its next sweep (tags ride in the data: poll until they all match)
                need = rounds[grp] + 1u;
                const uint4* bsrc = p.beta + ((size_t)(cta % kBetaCopies) * NGRP + grp) * kBetaWords;
                const bool absent = grp * kStreamCT + my_chain >= p.C;
                uint4 w[KS];
                const long long t_w = clock64();
                while (true) {
                    bool ok = true;
                    // (the abort flag travels in the same batch of loads: one L2 round trip per poll, not two)
                    const unsigned int aborted = ld_acquire(&sy->abort_flag);
#pragma unroll
                    for (int kk = 0; kk < KS; ++kk) {
                        w[kk] = ld_volatile_v4(bsrc + lane + 32 * kk);
                        ok = ok && ((w[kk].y & 0x7FFFFFFFu) == need) && ((w[kk].w & 0x7FFFFFFFu) == need);
                    }
                    if (absent) ok = true;
                    if (p.trace) {
                        const unsigned int bal = __ballot_sync(0xFFFFFFFFu, ok);
                        if (lane == 0) { p.trace[32 * blockIdx.x + 2] = bal; p.trace[32 * blockIdx.x + 3] += 1u; }
                    }
                    if (dbg) tdbg[13] += 1ull;                     // (beta poll rounds, CTA 0)
                    if (__all_sync(0xFFFFFFFFu, ok)) break;
                    // (warp-uniform exits: a lane that left alone would deadlock the __all_sync above)
                    bool give_up = aborted != 0u;
                    if (clock64() - t_w > p.spin_limit) { atomicCAS(&sy->abort_flag, 0u, 2u); give_up = true; }
                    if (__any_sync(0xFFFFFFFFu, give_up)) { status = 2; break; }
                }
                if (dbg) tdbg[12] += (unsigned long long)(clock64() - t_w);      // (cycles spent polling for the betas, CTA 0)
                const bool done_bit = absent || ((w[0].y >> 31) != 0u);
                if (status == 0 && __all_sync(0xFFFFFFFFu, done_bit)) {      // nothing left to sweep for this group
                    done_mask |= 1u << grp; cur = (grp + 1) % NGRP;
                    return 1;
                }
                // staged in the form the consumers' B fragments need: {b0, b1 (raw fp32 bits = the tf32 operand),
                // bf16x2(hi parts), bf16x2(lo parts)} -- split once here instead of in every consumer lane
                const bool zero = absent || status == 2;
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                    const float b0 = __uint_as_float(w[kk].x), b1 = __uint_as_float(w[kk].z);
                    float l0, l1; tf32_lo2(b0, b1, l0, l1);
                    bsq[lane + 32 * kk] = zero ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(w[kk].x, w[kk].z, pack_bf16(b0 - l0, b1 - l1), pack_bf16(l0, l1));
                }
            }
            if (lane == 0) {
                flags[qpass & 1] = status; flags[2 + (qpass & 1)] = grp; flags[4 + (qpass & 1)] = (int)need;
                if (status == 0) rounds[grp] = need;
            }
            warp_sync_hard();
            if (dbg) tdbg[17] = (unsigned long long)clock64();
            bar_arrive<kBarBeta, kStreamThreads>();  // release the consumers (they read flags and the staged betas)
            B2_TRACE_LANES(2);
            return status ? 2 : 0;
        };

        while (true) {
            // ---- next group with work (all warps of all CTAs derive the same schedule from the same betas)
            grp = -1;
            for (int k = 0; k < NGRP; ++k) { const int cand = (cur + k) % NGRP; if (!((done_mask >> cand) & 1u)) { grp = cand; break; } }
            if (grp < 0) status = 1;
            // My own next beta must exist before my group can be swept again (always the case with a single group): tick first.
            // Otherwise stage the other group's pass first, so that my tick overlaps with its sweep.
            const bool tick_first = (status == 0) && pending && grp == my_group;
            int staged = -1;
            if (!tick_first) {
                staged = stage();
                if (staged == 1) continue;
                if (staged == 2) break;
            }
            if (tick_first && DEFER && pending_pass == qpass - 1) {
                // deferred gather, but no other group is left to sweep meanwhile: a flush pseudo-pass (status 3) makes this
                // CTA's consumers gather the partials of the sweep that just ended
                const long long t_w = clock64();
                while (true) {
                    const int ready = __shfl_sync(0xFFFFFFFFu, (*started >= qpass - 1) ? 1 : 0, 0);
                    if (ready) break;
                    __nanosleep(256);
                    bool give_up = ld_acquire(&sy->abort_flag) != 0u;
                    if (clock64() - t_w > 4 * p.spin_limit) { atomicCAS(&sy->abort_flag, 0u, 2u); give_up = true; }
                    if (__any_sync(0xFFFFFFFFu, give_up)) { status = 2; break; }
                }
                if (status == 0) {
                    if (lane == 0) flags[qpass & 1] = 3;
                    warp_sync_hard();
                    bar_arrive<kBarBeta, kStreamThreads>();
                    ++qpass; ++pass;
                }
            }
            // (deferred gather: the consumers deliver the sums at the end of the pass AFTER my group's sweep, so the tick waits
            //  until one more pass has been staged for them -- otherwise they would sit idle while this warp ticks)
            if (pending && status == 0 && (tick_first || !DEFER || qpass - pending_pass >= 2)) {
                // ---- the tick of this CTA's chain after a sweep of its group: reduced likelihood sums -> potential -> NUTS
                //      state machine -> next beta (tag my_round + 1)
                const uint32_t seq = my_round;
                B2_TRACE_LANES(3);
                float* gz = xred + kXScratch;        // scratch for the gradient wrt z (<= 64 floats when the early publish applies)
                float* zpeek = gz + 64;              // ... and for the position published ahead of the tick (<= 64 floats)
                bool early = false;
                bar_sync<kBarTick, kStreamThreads>();    // the sums over the CTAs' partials are in `gred`
                if (dbg) tdbg[16] += (unsigned long long)clock64() - tdbg[19];   // (hand-over latency consumers -> tick warp, CTA 0)
                B2_LAPQ(-1);
                B2_TRACE_LANES(4);
                const long long t_a = clock64();
                float a64 = gred[64];                // (the consumers left the sums of the 148 partials in gred: columns, [64] = nll)
                if (p.shard_count > 1) {
                    // ---- all-reduce over the row shards: store this rank's sums into every rank's mailbox, then add the
                    //      shard_count contributions of this chain in rank order (same order everywhere => identical bits)
                    const uint32_t xtag = (p.epoch << 24) ^ (seq & 0xFFFFFFu);
                    const size_t slot0 = ((size_t)(seq & 1u) * kMaxStreamChains + (size_t)cta) * kMaxShards * kGStride;
                    const long long t_x = clock64();
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int d = lane + 32 * k;
                        if (d < 65) {
                            float mine = gred[d];
                            if (d == 64) mine = (mine - pad_nll) + p.nll_local_const;
                            for (int q = 0; q < p.shard_count; ++q)
                                st_sys_v2(p.mail[q] + slot0 + (size_t)p.shard_rank * kGStride + d, make_float2(mine, __uint_as_float(xtag)));
                            float tot = 0.0f;
                            const float2* box = p.mail[p.shard_rank] + slot0 + d;
                            for (int sr = 0; sr < p.shard_count; ++sr) {
                                float2 v;
                                while (true) {
                                    v = ld_sys_v2(box + (size_t)sr * kGStride);
                                    if (__float_as_uint(v.y) == xtag) break;
                                    if (ld_acquire(&sy->abort_flag)) break;
                                    if (clock64() - t_x > 16 * p.spin_limit) { atomicCAS(&sy->abort_flag, 0u, 4u); break; }
                                }
                                tot += v.x;
                            }
                            gred[d] = tot;
                            if (k == 2) a64 = tot;
                        }
                    }
                }
                const float nll = __shfl_sync(0xFFFFFFFFu, a64, 0) - (p.shard_count > 1 ? 0.0f : pad_nll);
                if (lane == 0) tlap[3] += (unsigned long long)(clock64() - t_a);
                B2_LAPQ(6);
                if (!chain_done) {
                    early = stream_tick_critical(p, tk, gred, gz, zpeek, nll, cta, tdbg, seq + 1u, def_u, peek_stat);
                    if (p.mode == 1) chain_done = true;
                    else if (early && !MG) deferred = true;       // single group: advance() runs after the next pass has been staged
                    else chain_done = stream_tick_deferred(p, tk, gz, zpeek, def_u, early, tdbg, peek_stat);
                }
                B2_TRACE_LANES(5);
                const long long t_p = clock64();
                if (!early) stream_publish_beta(p, cta, cv.v(V_ZS), !chain_done, (seq + 1u) | (chain_done ? 0x80000000u : 0u));
                B2_LAPQ(7);
                if (lane == 0) {
                    const long long t_e = clock64();
                    tlap[2] += (unsigned long long)(t_e - t_p);
                    const unsigned long long dt = (unsigned long long)(t_e - t_a);
                    tk_sum += dt; if (dt > tk_max) tk_max = dt;
                    if (dbg) tdbg[6] += dt;
                }
                pending = false;
            }
            if (tick_first) staged = stage();
            if (deferred) {                          // the bookkeeping of the tick whose next position already went out (also
                                                     // when the launch stops here: the chain must sit in front of that position)
                chain_done = stream_tick_deferred(p, tk, xred + kXScratch, xred + kXScratch + 64, def_u, true, tdbg, peek_stat);
                deferred = false;
            }
            if (tick_first && staged == 2) break;
            if (tick_first && staged == 1) continue;
            if (is_tick && grp == my_group) {
                pending = true; my_round = need; pending_pass = qpass;
                if (p.mode == 0 && !chain_done && !p.no_prefetch) tk.prefetch();     // off the critical path: PRNG look-ahead
            }
            cur = (grp + 1) % NGRP; ++qpass; ++pass;
        }
        B2_TRACE_LANES(6);
#ifdef B2_TICK_LAPS
        if (cta == 0 && lane == 0) for (int k = 0; k < 32; ++k) sy->laps[k] = b2_lap_store()[k];
#endif
        if (is_tick && lane == 0) {
            if (p.mode == 0) p.ctl[cta] = c;
            sy->tick_sum[cta] = tk_sum; sy->tick_max[cta] = tk_max;
            sy->peek_hit[cta] = peek_stat[0]; sy->peek_fallback[cta] = peek_stat[1]; sy->peek_mismatch[cta] = peek_stat[2];
            for (int k = 0; k < 4; ++k) { sy->tick_lap[cta][k] = tlap[k]; sy->pre_hit[cta][k] = c.pre_hit[k]; sy->pre_miss[cta][k] = c.pre_miss[k]; }
        }
        __syncthreads();                             // pairs with the consumers' shutdown barrier: chain vectors are final
        return;
    }

    // =============================================================== consumer warps
    const int cw = warp;                             // 0..14
    const int ctid = tid;                            // 0..479
    // ---- this warp's private ring: its tiles are w, w + 15, w + 30, ... of the CTA's slice, over and over
    const int n_mine = (n_tiles > cw) ? (n_tiles - cw + kConsWarps - 1) / kConsWarps : 0;
    float* my_tiles = tiles + (size_t)cw * nst * SLOT_FLOATS;
    uint64_t* my_full = full + cw * kMaxStages;
    const float* my_src = p.img + (t_begin_tile + cw) * TILE_FLOATS;
    auto issue = [&](int slot, int j) {              // one lane: tile j of this warp -> slot
        mbar_expect_tx(&my_full[slot], TILE_FLOATS * 4u);
        bulk_g2s(my_tiles + (size_t)slot * SLOT_FLOATS, my_src + (size_t)j * kConsWarps * TILE_FLOATS, TILE_FLOATS * 4u, &my_full[slot]);
    };
    int slot = 0; uint32_t parity = 0;               // ring position of the next tile to consume
    if (lane == 0 && n_mine > 0)
        for (int s = 0; s < nst; ++s) issue(s, s % n_mine);
    // A warp with no tiles still owns its (empty) slot 0 as reduction scratch; the others use the slot they drained last.
    int red_slot = 0, red_tile = -1;

    const bool dbg = (ctid == 0);
    if (dbg) { tdbg[14] = (unsigned long long)clock64(); tdbg[15] = tdbg[14]; }

    while (true) {
        // ---- wait until this CTA's tick warp has staged every chain's beta of this pass
        if (ctid == 0) B2_TRACE(0, 1);
        bar_sync<kBarBeta, kStreamThreads>();
        const int st = flags[pass & 1u];
        if (st == 1 || st == 2) break;
        const bool real = !MG || st == 0;             // (3: flush pseudo-pass of the deferred gather -- nothing to sweep)
        if (ctid == 0) { *(volatile int*)(flags + 8) = (int)pass; B2_TRACE(0, 2); }
        if (ctid == 0) { tdbg[18] += (unsigned long long)clock64() - tdbg[17]; B2_DBG_LAP(0); }
        int grp = 0; uint32_t tag = 0u;
        if (real) {
        {
            const uint32_t r = stream_sweep<KS, LIK, MG>(p, (uint32_t)slot | (parity << 2), pass, n_tiles, (unsigned int)t_begin_tile);
            slot = (int)(r & 3u); parity = (r >> 2) & 1u; red_slot = (int)((r >> 3) & 3u); red_tile = (int)(r >> 5) - 1;
        }
        if (ctid == 0) { B2_DBG_LAP(1); B2_TRACE(0, 3); }
        // (read after the sweep so that they do not occupy registers during it; the tick warp rewrites this pass's words only
        //  after the consumers have begun the next pass)
        grp = MG ? flags[2 + (pass & 1u)] : 0;        // chain group served by this pass and the round (tag) of its sweep
        tag = (uint32_t)flags[4 + (pass & 1u)];

        stream_reduce_publish<KS, MG>(p, nst, grp, tag);
        if (!MG) bar_arrive<kBarSwept, kStreamThreads>();     // the tick warp may start polling for the next betas
        // (single group: an owner CTA gathers right away and uses the drained scratch slots as landing zones -- refilled after that)
        if (!(!MG && is_tick) && lane == 0 && red_tile >= 0 && p.dbg_sweep != 2) issue(red_slot, red_tile);
        if (ctid == 0) { B2_DBG_LAP(2); B2_TRACE(0, 4); }

        }   // real pass
        // ---- chain owner: poll the partials of all CTAs (tags ride in the data), sum them in fixed order,
        //      hand over to the tick warp.  Deferred mode: gather what the PREVIOUS pass left, note what this one leaves.
        int ggrp = grp; uint32_t gtag = tag;
        bool do_g = is_tick && real && (!MG || cta / kStreamCT == grp);
        if (DEFER) {
            const int* dprev = flags + 10 + 3 * (int)((pass + 1u) & 1u);
            do_g = is_tick && dprev[0] != 0; ggrp = dprev[1]; gtag = (uint32_t)dprev[2];
            if (ctid == 0) {
                int* dcur = flags + 10 + 3 * (int)(pass & 1u);
                dcur[0] = (is_tick && real && cta / kStreamCT == grp) ? 1 : 0; dcur[1] = grp; dcur[2] = (int)tag;
            }
        }
        if (do_g) {
            if constexpr (!MG) {
                stream_gather_land<KS>(p, gtag, my_tiles + (size_t)red_slot * SLOT_FLOATS);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes before the bulk copy reuses the slot
                if (lane == 0 && red_tile >= 0 && p.dbg_sweep != 2) issue(red_slot, red_tile);
            } else {
                stream_gather<KS, MG>(p, ggrp, gtag);
                bar_arrive<kBarTick, kStreamThreads>();  // tick warp takes over; consumers go wait for the next beta
            }
            if (ctid == 0) { B2_DBG_LAP(4); B2_TRACE(0, 5); }
        }
        ++pass;
    }

    // ---- shutdown: drain this warp's outstanding copies, write the chain state back
    if (ctid == 0) B2_TRACE(0, 6);
    if (n_mine > 0 && p.dbg_sweep != 2)
        for (int s = 0; s < nst; ++s) {
            mbar_wait_bounded(&my_full[slot], parity, p.spin_limit, &sy->abort_flag);
            if (++slot == nst) { slot = 0; parity ^= 1u; }
        }
    if (ctid == 0) B2_TRACE(0, 7);
    __syncthreads();                                 // tick warp included: the chain vectors are final
    if (ctid == 0) B2_TRACE(0, 8);
    if (is_tick && p.vecs_in_smem && p.mode == 0) {
        for (int i = ctid; i < V_COUNT * p.Dp; i += kConsThreads) {
            const int f = i / p.Dp, d = i - f * p.Dp;
            p.vecs[((size_t)f * p.C + cta) * p.Dp + d] = cvecs[i];
        }
    }
    if (ctid == 0 && cta < 160) { sy->cta_lap[cta][0] = tdbg[0]; sy->cta_lap[cta][1] = tdbg[1]; sy->cta_lap[cta][2] = tdbg[2]; sy->cta_lap[cta][3] = tdbg[4]; }
    if (cta == 0 && ctid == 0) {
        sy->passes = pass;
        tdbg[5] = (unsigned long long)clock64() - tdbg[15];
        for (int i = 0; i < 16; ++i) sy->dbg[i] = tdbg[i];
        sy->wake[0] = tdbg[16]; sy->wake[1] = tdbg[18];
    }
#undef B2_DBG_LAP
}

}  // namespace b2
