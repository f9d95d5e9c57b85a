// Regime R3: many-chain GEMM engine on the 5th-generation tensor cores (BASELINE configs 3 and 4).
//
// With hundreds to tens of thousands of chains the per-leapfrog likelihood gradient is two dense contractions,
//     L = B X^T   (chains x rows,   K = columns of X)         eta of every chain at every row
//     G = R X     (chains x columns, K = rows)                gbeta = X^T dl/deta
// with the link function in between.  The reference executes them as three XLA ops under vmap and materialises the
// [rows, chains] logits (numpyro/infer/hmc.py:790-798 vmapped sample_fn over potential_energy, infer/util.py:333-358).
// Here one pass of all chains is ONE kernel launch that never materialises L or R outside the SM:
//
//   * a CTA owns a chain tile of 128 chains (the M of every MMA, one chain per TMEM lane) and walks over chunks of 128 rows;
//   * forward: tcgen05.mma kind::tf32, A = betas [128 chains x 32 k] and B = X [128 rows x 32 k] from shared memory (K-major
//     SWIZZLE_128B tiles moved by cp.async.bulk = the TMA engine, 3-stage mbarrier ring), D = logits in tensor memory;
//   * epilogue warps pull the logits out of TMEM (tcgen05.ld), apply the link function (loss + residual), and write the
//     residuals back INTO tensor memory (tcgen05.st) -- they become the A operand of the backward MMAs straight from TMEM,
//     so R never touches shared or global memory; they are delivered in four row groups so that the backward MMAs of
//     group 0 start after a quarter of the epilogue;
//   * backward: tcgen05.mma with A = R from TMEM, B = X^T [<= 256 columns x 32 rows] from shared memory (a transposed tile
//     image of X, also K-major), D = gbeta^T [128 chains x <= 256 columns] in TMEM.
//
// fp32 parity on tf32 tensor cores (measured on B200, profiles/r02_umma_probe2.log):
//   * operands are split x = hi + lo (hi = the 19 bits kind::tf32 reads, lo = x - hi) and every product is three MMAs
//     hi*hi + lo*hi + hi*lo: error 8e-7 of sum|a b| against an exact product;
//   * the TMEM accumulator TRUNCATES on every accumulating MMA (-6e-8 relative per MMA, measured), so accumulation chains are
//     kept short: the forward uses two accumulators (hi*hi terms / cross terms), and gbeta is accumulated in TMEM for ONE row
//     chunk only (48 MMAs), then drained by the epilogue warps into fp32 registers (round-to-nearest adds) that carry the
//     sums over the whole unit.
//
// Work decomposition.  X is cut into S fixed row segments; a unit = (chain tile, segment, column block of <= 256).  Units
// are dealt round-robin to the CTAs in segment-major order, so all CTAs stream the same rows at the same time (X comes out
// of L2).  Every unit writes its own partial sums; the tick kernel adds the S partials of a chain in fixed order.  Results
// are therefore independent of scheduling and of which other chains are active: bit-reproducible, and identical between the
// run path and the potential hook (the parity tests rely on that).
// For more than 256 columns (horseshoe, 1000 columns) the backward product is done per column block, each block repeating
// the forward product (the TMEM accumulator holds 256 columns).
//
// Between two GEMM launches the tick kernel (one warp per chain) finishes the potential, advances the chain's NUTS state
// machine (tick.cuh) and writes the chain's next betas straight into the tile image.  The whole run is a CUDA graph with
// a device-side WHILE node (gemm -> tick -> schedule), so b200nuts_run only enqueues.
#pragma once
#include <cuda_runtime.h>
#include "tick.cuh"
#include "families.cuh"
#include "linkfn.cuh"
#include "umma.cuh"

namespace b2 {

constexpr int kGtChains = 128;                 // chains per tile = M of the MMAs = TMEM lanes
constexpr int kGtRows = 128;                   // rows per chunk = N of the forward MMA
constexpr int kGtNB = 256;                     // columns per column block = max N of the backward MMA
constexpr int kGtStages = 3;
constexpr int kGtStageBytes = 65536;           // forward: betas hi|lo (32 KB) + X hi|lo (32 KB); backward: X^T hi|lo (<= 64 KB)
constexpr int kGtEpiWarps = 16;                // 4 per TMEM lane quarter
constexpr int kGtThreads = 32 * (2 + kGtEpiWarps);
constexpr int kGtTileFloats = 128 * 32;        // one [128 x 32] fp32 tile
constexpr int kGtYSlots = 4;
// TMEM columns: [0,128) logits (hi*hi terms) then residual hi; [128,256) logits (cross terms) then residual lo; [256,512) gbeta
constexpr uint32_t kColLa = 0, kColLb = 128, kColG = 256;

struct GemmParams {
    const float* bimg;                         // betas   [CT][KB][hi|lo][128 chains x 32]
    const float* ximg;                         // X       [RC][KB][hi|lo][128 rows x 32]
    const float* xtimg;                        // X^T     column block j: [4 RC][hi|lo][NB_j columns x 32 rows]
    const float* yimg;                         // y       [RC][128]
    float* partial;                            // [CT][S][128][Dxp] likelihood gradient sums per unit
    float* pnll;                               // [CT][S][4 column groups of the epilogue][128]
    const int* active_tiles; const int* n_active;   // chain tiles that still need gradients (written by k_gemm_sched)
    int CT, RC, KB, S, cps, NDB, Dxp;          // cps = chunks per segment
    long long N;
    unsigned int* abort_flag; long long spin_limit;
    unsigned long long* dbg;                   // CTA 0: [0] cycles, [1] units, MMA thread waiting [2] for the epilogue, [3] for tile
                                               // copies, [4] for its own backward MMAs to complete
};

B2_HD size_t gemm_smem_bytes() { return (size_t)kGtStages * kGtStageBytes + kGtYSlots * 512 + 1024 /* alignment */ + 2048; }
B2_HD long long gemm_xt_block_floats(long long RC, int j) { return RC * 4 * 64 * (256ll * j); }   // float offset of column block j

// One pass: likelihood sums of every active chain tile.
template <int LIK>
__global__ void __launch_bounds__(kGtThreads, 1) gemm_pass_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ unsigned char gsm_raw[];
    unsigned char* sm = (unsigned char*)(((uintptr_t)gsm_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* stage = sm;                                         // kGtStages x 64 KB, 1024-byte aligned
    float* ysm = (float*)(sm + (size_t)kGtStages * kGtStageBytes);     // kGtYSlots x 128
    uint64_t* bars = (uint64_t*)(ysm + kGtYSlots * 128);
    uint64_t* full = bars;                     // [kGtStages]
    uint64_t* empty = bars + kGtStages;        // [kGtStages]
    uint64_t* ybar = bars + 2 * kGtStages;     // [kGtYSlots]
    uint64_t* l_full = ybar + kGtYSlots;       // forward MMAs of a chunk done
    uint64_t* g_full = l_full + 1;             // backward MMAs of a chunk done
    uint64_t* g_free = l_full + 2;             // gbeta drained
    uint64_t* r_ready = l_full + 3;            // [4] residuals of rows [32 rk, 32 rk + 32) of the chunk are in TMEM
    uint32_t* tmem_slot = (uint32_t*)(l_full + 7);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < kGtStages; ++i) { u_mbar_init(&full[i], 1); u_mbar_init(&empty[i], 1); }
        for (int i = 0; i < kGtYSlots; ++i) u_mbar_init(&ybar[i], 1);
        u_mbar_init(l_full, 1); u_mbar_init(g_full, 1); u_mbar_init(g_free, kGtEpiWarps);
        for (int i = 0; i < 4; ++i) u_mbar_init(&r_ready[i], kGtEpiWarps);
        u_fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int n_active = *p.n_active;
    const int per_block = n_active * p.S;                 // units = (segment, chain tile)
    // More than 256 columns: every chunk runs the forward product ONCE and then one backward product per column block
    // (the residuals stay in tensor memory); a block's sums are drained straight into the unit's partial sums in global memory.
    const int NJ = p.NDB;
    const int n_units = per_block;
    const long long t_start = clock64();
    unsigned long long wait_epi = 0ull, wait_tma = 0ull, wait_mma = 0ull, wait_gfree = 0ull;

    if (warp == 0) {
        // ================================================================= producer: one lane feeds the ring
        if (lane == 0) {
            int st = 0; uint32_t ph = 0; unsigned int chunk_no = 0; bool ok = true;
            for (int u = blockIdx.x; u < n_units && ok; u += gridDim.x) {
                const int s = u / n_active, t = p.active_tiles[u - s * n_active];
                const int c0 = s * p.cps, c1 = min(p.RC, c0 + p.cps);
                for (int c = c0; c < c1 && ok; ++c, ++chunk_no) {
                    {   // responses of this chunk
                        const int ys = chunk_no % kGtYSlots;
                        u_mbar_expect_tx(&ybar[ys], 512u);
                        u_bulk_g2s(ysm + ys * 128, p.yimg + (size_t)c * 128, 512u, &ybar[ys]);
                    }
                    for (int kb = 0; kb < p.KB && ok; ++kb) {
                        ok = u_mbar_wait(&empty[st], ph ^ 1u, p.spin_limit, p.abort_flag, 11u);
                        unsigned char* dst = stage + (size_t)st * kGtStageBytes;
                        u_mbar_expect_tx(&full[st], 65536u);
                        u_bulk_g2s(dst, p.bimg + ((size_t)t * p.KB + kb) * 2 * kGtTileFloats, 32768u, &full[st]);
                        u_bulk_g2s(dst + 32768, p.ximg + ((size_t)c * p.KB + kb) * 2 * kGtTileFloats, 32768u, &full[st]);
                        if (++st == kGtStages) { st = 0; ph ^= 1u; }
                    }
                    for (int j = 0; j < NJ && ok; ++j) {
                        const int nb = min(kGtNB, p.Dxp - kGtNB * j);
                        const float* xt_base = p.xtimg + gemm_xt_block_floats(p.RC, j);
                        for (int rk = 0; rk < 4 && ok; ++rk) {
                            ok = u_mbar_wait(&empty[st], ph ^ 1u, p.spin_limit, p.abort_flag, 12u);
                            unsigned char* dst = stage + (size_t)st * kGtStageBytes;
                            const uint32_t bytes = 2u * (uint32_t)nb * 128u;
                            u_mbar_expect_tx(&full[st], bytes);
                            u_bulk_g2s(dst, xt_base + ((size_t)c * 4 + rk) * 2 * nb * 32, bytes, &full[st]);
                            if (++st == kGtStages) { st = 0; ph ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer: one thread
        if (lane == 0) {
            int st = 0; uint32_t ph = 0; unsigned int chunk_no = 0, blk_no = 0; bool ok = true;      // blk_no: backward blocks issued
            const uint32_t idesc_f = umma_idesc_tf32(kGtChains, kGtRows);
            for (int u = blockIdx.x; u < n_units && ok; u += gridDim.x) {
                const int s = u / n_active;
                const int c0 = s * p.cps, c1 = min(p.RC, c0 + p.cps);
                for (int c = c0; c < c1 && ok; ++c, ++chunk_no) {
                    const uint32_t cpar = chunk_no & 1u;
                    // the previous chunk's backward MMAs read the residuals from the columns the logits go to: they must
                    // have completed (g_full of its last column block) before the forward MMAs of this chunk are issued
                    if (blk_no > 0) {
                        const long long tw = clock64();
                        ok = u_mbar_wait(g_full, (blk_no - 1u) & 1u, p.spin_limit, p.abort_flag, 13u);
                        wait_mma += (unsigned long long)(clock64() - tw);
                    }
                    tc_fence_after();
                    for (int kb = 0; kb < p.KB && ok; ++kb) {
                        {
                            const long long tw = clock64();
                            ok = u_mbar_wait(&full[st], ph, p.spin_limit, p.abort_flag, 14u);
                            wait_tma += (unsigned long long)(clock64() - tw);
                        }
                        tc_fence_after();
                        const uint32_t sa = u_smem(stage + (size_t)st * kGtStageBytes);
                        const uint64_t bh = umma_desc_k_sw128(sa), bl = umma_desc_k_sw128(sa + 16384u);
                        const uint64_t xh = umma_desc_k_sw128(sa + 32768u), xl = umma_desc_k_sw128(sa + 49152u);
#pragma unroll
                        for (uint32_t k4 = 0; k4 < 4; ++k4) {      // +32 bytes per k-step = +2 in the descriptor's address field
                            const uint32_t first = (kb == 0 && k4 == 0) ? 0u : 1u;
                            umma_tf32_ss(tmem + kColLa, bh + 2 * k4, xh + 2 * k4, idesc_f, first);     // hi * hi
                            umma_tf32_ss(tmem + kColLb, bl + 2 * k4, xh + 2 * k4, idesc_f, first);     // lo * hi
                            umma_tf32_ss(tmem + kColLb, bh + 2 * k4, xl + 2 * k4, idesc_f, 1u);        // hi * lo
                        }
                        umma_commit(&empty[st]);
                        if (++st == kGtStages) { st = 0; ph ^= 1u; }
                    }
                    umma_commit(l_full);
                    // the epilogue delivers the residuals in four groups of 32 rows (= one backward k-block each): the backward
                    // MMAs of a group start while the link function of the next groups is still being evaluated
                    for (int j = 0; j < NJ && ok; ++j, ++blk_no) {
                    const int nb = min(kGtNB, p.Dxp - kGtNB * j);
                    const uint32_t idesc_b = umma_idesc_tf32(kGtChains, nb);
                    for (uint32_t rk = 0; rk < 4 && ok; ++rk) {
                        {
                            const long long tw = clock64();
                            if (j == 0) ok = u_mbar_wait(&r_ready[rk], cpar, p.spin_limit, p.abort_flag, 15u);
                            const long long tg = clock64();
                            // the gbeta columns are free once the epilogue has drained the previous block
                            if (rk == 0 && blk_no > 0) ok = ok && u_mbar_wait(g_free, (blk_no - 1u) & 1u, p.spin_limit, p.abort_flag, 16u);
                            wait_gfree += (unsigned long long)(clock64() - tg);
                            wait_epi += (unsigned long long)(clock64() - tw);
                        }
                        {
                            const long long tw = clock64();
                            ok = ok && u_mbar_wait(&full[st], ph, p.spin_limit, p.abort_flag, 17u);
                            wait_tma += (unsigned long long)(clock64() - tw);
                        }
                        tc_fence_after();
                        const uint32_t sa = u_smem(stage + (size_t)st * kGtStageBytes);
                        const uint64_t th = umma_desc_k_sw128(sa), tl = umma_desc_k_sw128(sa + (uint32_t)nb * 128u);
#pragma unroll
                        for (uint32_t k4 = 0; k4 < 4; ++k4) {
                            const uint32_t col = 32u * rk + 8u * k4;
                            umma_tf32_ts(tmem + kColG, tmem + kColLa + col, th + 2 * k4, idesc_b, (rk == 0 && k4 == 0) ? 0u : 1u);   // r_hi * x_hi
                            umma_tf32_ts(tmem + kColG, tmem + kColLb + col, th + 2 * k4, idesc_b, 1u);                               // r_lo * x_hi
                            umma_tf32_ts(tmem + kColG, tmem + kColLa + col, tl + 2 * k4, idesc_b, 1u);                               // r_hi * x_lo
                        }
                        umma_commit(&empty[st]);
                        if (++st == kGtStages) { st = 0; ph ^= 1u; }
                    }
                    umma_commit(g_full);
                    }
                }
            }
        }
    } else {
        // ================================================================= epilogue warps
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int cg = (warp - 2) >> 2;            // which quarter of the columns it handles
        const uint32_t lane_base = (uint32_t)(32 * q) << 16;
        const int chain_row = 32 * q + lane;       // chain within the tile
        float gsum[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) gsum[i] = 0.0f;
        float nll_unit = 0.0f;
        unsigned int chunk_no = 0, blk_no = 0; bool ok = true;
        for (int u = blockIdx.x; u < n_units && ok; u += gridDim.x) {
            const int s = u / n_active, t = p.active_tiles[u - s * n_active];
            const int c0 = s * p.cps, c1 = min(p.RC, c0 + p.cps);
            float* const unit_dst = p.partial + (((size_t)t * p.S + s) * kGtChains + chain_row) * p.Dxp + 64 * cg;
            for (int c = c0; c < c1 && ok; ++c, ++chunk_no) {
                const uint32_t cpar = chunk_no & 1u;
                const int ys = chunk_no % kGtYSlots;
                const long long rows_left = p.N - (long long)c * kGtRows;
                const int n_valid = rows_left >= kGtRows ? kGtRows : (int)rows_left;
                ok = u_mbar_wait(&ybar[ys], (chunk_no / kGtYSlots) & 1u, p.spin_limit, p.abort_flag, 18u);
                ok = ok && u_mbar_wait(l_full, cpar, p.spin_limit, p.abort_flag, 19u);
                ok = __all_sync(0xFFFFFFFFu, ok);            // (tcgen05.ld / st are warp-collective: leave together)
                if (!ok) break;
                tc_fence_after();
                // ---- logits -> loss, residual (hi, lo) back into TMEM.  Row group rk (32 rows = one backward k-block) is
                //      shared by all 16 warps: this warp takes its chains' columns [32 rk + 8 cg, +8), so group 0 is complete
                //      after a quarter of the epilogue and the backward MMAs overlap with the rest.
                float nl0 = 0.0f, nl1 = 0.0f;
#pragma unroll
                for (int rk = 0; rk < 4; ++rk) {
                    const int col = 32 * rk + 8 * cg;
                    uint32_t a[8], b[8];
                    tmem_ld8(tmem + lane_base + kColLa + col, a);
                    tmem_ld8(tmem + lane_base + kColLb + col, b);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float eta = __uint_as_float(a[i]) + __uint_as_float(b[i]);
                        float loss, dl;
                        link_fn<LIK>(eta, ysm[ys * 128 + col + i], loss, dl);
                        if (col + i >= n_valid) { loss = 0.0f; dl = 0.0f; }
                        if (i & 1) nl1 += loss; else nl0 += loss;
                        a[i] = __float_as_uint(dl);
                        b[i] = __float_as_uint(u_tf32_lo(dl));
                    }
                    tmem_st8(tmem + lane_base + kColLa + col, a);
                    tmem_st8(tmem + lane_base + kColLb + col, b);
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) u_mbar_arrive(&r_ready[rk]);
                }
                nll_unit += nl0 + nl1;
                // ---- drain gbeta of every column block of this chunk (columns [64 cg, 64 cg + 64) of this thread's chain): into
                //      the fp32 register sums of the unit (one block), or -- several blocks -- into the unit's partial sums in
                //      global memory (first chunk stores, later chunks add; the unit belongs to this CTA alone)
                for (int j = 0; j < NJ && ok; ++j, ++blk_no) {
                    const int nb = min(kGtNB, p.Dxp - kGtNB * j);
                    ok = u_mbar_wait(g_full, blk_no & 1u, p.spin_limit, p.abort_flag, 20u);
                    ok = __all_sync(0xFFFFFFFFu, ok);
                    if (!ok) break;
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int col = 64 * cg + 16 * k;
                        if (col < nb) {
                            uint32_t v[16];
                            tmem_ld16(tmem + lane_base + kColG + col, v);
                            tmem_wait_ld();
                            if (NJ == 1) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) gsum[16 * k + i] += __uint_as_float(v[i]);
                            } else {
                                float* dst = unit_dst + kGtNB * j + 16 * k;
#pragma unroll
                                for (int i = 0; i < 16; i += 4) {
                                    float4 a = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
                                    if (c > c0) { const float4 o = *reinterpret_cast<const float4*>(dst + i); a.x = o.x + a.x; a.y = o.y + a.y; a.z = o.z + a.z; a.w = o.w + a.w; }
                                    *reinterpret_cast<float4*>(dst + i) = a;
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) u_mbar_arrive(g_free);
                }
            }
            // ---- end of the unit: write its partial sums
            if (ok) {
                if (NJ == 1) {
                    const int nb = min(kGtNB, p.Dxp);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (64 * cg + 16 * k < nb) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4)
                                *reinterpret_cast<float4*>(unit_dst + 16 * k + i) = make_float4(gsum[16 * k + i], gsum[16 * k + i + 1], gsum[16 * k + i + 2], gsum[16 * k + i + 3]);
                        }
                }
                p.pnll[(((size_t)t * p.S + s) * 4 + cg) * kGtChains + chain_row] = nll_unit;
            }
#pragma unroll
            for (int i = 0; i < 64; ++i) gsum[i] = 0.0f;
            nll_unit = 0.0f;
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
    if (blockIdx.x == 0 && tid == 32 && p.dbg) {
        p.dbg[0] += (unsigned long long)(clock64() - t_start);
        p.dbg[1] += (unsigned long long)((n_units + (int)gridDim.x - 1) / (int)gridDim.x);
        p.dbg[2] += wait_epi; p.dbg[3] += wait_tma; p.dbg[4] += wait_mma; p.dbg[5] += wait_gfree;
    }
}

// ---- tile images of the data (built once at create) -----------------------------------------------------------
// X [N, Dx] row-major -> ximg [RC][KB][hi|lo][128 rows x 32] (zero padded)
static __global__ void k_gemm_pack_x(const float* __restrict__ X, long long N, int Dx, int KB, long long RC, float* __restrict__ ximg) {
    const long long total = RC * KB * kGtTileFloats;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i & 31), r = (int)((i >> 5) & 127);
        const long long ck = i >> 12; const int kb = (int)(ck % KB); const long long c = ck / KB;
        const long long row = c * kGtRows + r; const int col = kb * 32 + k;
        const float v = (row < N && col < Dx) ? X[row * Dx + col] : 0.0f;
        float* tile = ximg + ck * 2 * kGtTileFloats;
        const int o = sw128_index(r, k);
        tile[o] = v;
        tile[kGtTileFloats + o] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
}
// X -> xtimg: column block j, row k-block rk (32 rows): [hi|lo][nb columns x 32 rows]
static __global__ void k_gemm_pack_xt(const float* __restrict__ X, long long N, int Dx, int Dxp, long long RC, float* __restrict__ xtimg) {
    const long long rows_p = RC * kGtRows;
    const long long total = rows_p * Dxp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / Dxp; const int col = (int)(i - row * Dxp);
        const int j = col / kGtNB, dl = col - j * kGtNB, nb = min(kGtNB, Dxp - kGtNB * j);
        const long long rk = row >> 5; const int k = (int)(row & 31);
        const float v = (row < N && col < Dx) ? X[row * Dx + col] : 0.0f;
        float* tile = xtimg + gemm_xt_block_floats(RC, j) + rk * 2 * nb * 32;
        const int o = sw128_index(dl, k);
        tile[o] = v;
        tile[nb * 32 + o] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
}
static __global__ void k_gemm_pack_y(const float* __restrict__ y, long long N, long long RC, float* __restrict__ yimg) {
    const long long total = RC * kGtRows;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) yimg[i] = i < N ? y[i] : 0.0f;
}

// ---- per-pass bookkeeping shared by the kernels of the graph ---------------------------------------------------
struct GemmCtx {                               // device memory, rewritten by b200nuts_run before every launch
    TickCfg cfg; OutBufs out;
    int max_passes;                            // 0 = run until every chain has finished
    unsigned int epoch;                        // launch number of a row-sharded handle (part of the exchange tags)
};
// Row sharding (BASELINE config 5): every rank runs the GEMM pass over its own rows with all chains replicated; the
// per-chain likelihood sums [Dx gradient columns + loss] are all-reduced inside the tick kernel: each rank stores its sums as
// {value, tag} 8-byte words into EVERY rank's mailbox (peer memory over NVLink) and adds the `count` contributions in rank
// order, so all ranks hold bit-identical (U, grad) and their replicated chains stay in lock-step.
constexpr int kGemmMaxShards = 16;
struct GemmShardDev {
    int rank, count;
    int stride;                                // float2 words per (parity, source rank, chain): Dxp + 32 (loss at index Dxp)
    float nll_local_const;                     // this rank's share of the likelihood constant (poisson: sum lgamma(y+1) over its rows)
    float2* mail[kGemmMaxShards];              // [2 parities][kGemmMaxShards source ranks][C][stride] of every rank
};
B2_HD size_t gemm_mail_words(int C, int Dxp) { return (size_t)2 * kGemmMaxShards * C * (Dxp + 32); }
struct GemmSched {                             // device memory
    int n_active; int pass; int pass_in_run; unsigned int abort_flag;
    unsigned long long passes_total; unsigned long long dbg[8];
};

// betas of one chain -> its row of the tile image (hi and lo parts)
B2_D void gemm_write_betas(const FamilySpec& f, const float* z, float* bimg, int KB, int chain) {
    const int t = chain / kGtChains, r = chain % kGtChains;
    const int lane = threadIdx.x & 31;
    for (int kb = 0; kb < KB; ++kb) {
        const int d = kb * 32 + lane;
        float b = 0.0f;
        if (d < f.Dx) b = glm_scale_at(f, z, d) * z[f.off_u + d];
        float* tile = bimg + ((size_t)t * KB + kb) * 2 * kGtTileFloats;
        const int o = sw128_index(r, lane);
        tile[o] = b;
        tile[kGtTileFloats + o] = u_tf32_lo(b);
    }
}

// Sum of the S unit partials of every chain, in segment order (fixed order => bit-reproducible): gbeta [C][Dxp], nll [C].
// A kernel of its own between the GEMM pass and the tick: the sums of a wide model are hundreds of MB per pass
// (horseshoe config: 79 segments x 1024 chains x 1024 columns), far too much for the one warp per chain of the tick.
static __global__ void __launch_bounds__(256) k_gemm_reduce(GemmParams gp, float* __restrict__ gbeta, float* __restrict__ nll_out, int C) {
    const int q4 = gp.Dxp >> 2;                                 // float4 words per chain
    const long long total = (long long)C * q4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int chain = (int)(i / q4), d4 = (int)(i - (long long)chain * q4);
        const int t = chain / kGtChains, r = chain % kGtChains;
        const float4* src = reinterpret_cast<const float4*>(gp.partial + ((size_t)t * gp.S * kGtChains + r) * gp.Dxp) + d4;
        const size_t stride = (size_t)kGtChains * q4;
        float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        for (int s = 0; s < gp.S; ++s) { const float4 v = src[(size_t)s * stride]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
        reinterpret_cast<float4*>(gbeta + (size_t)chain * gp.Dxp)[d4] = a;
        if (d4 == 0) {
            float nll = 0.0f;
            const float* pn = gp.pnll + (size_t)t * gp.S * 4 * kGtChains + r;
            for (int s = 0; s < gp.S * 4; ++s) nll += pn[(size_t)s * kGtChains];
            nll_out[chain] = nll;
        }
    }
}

B2_D void g_st_sys_v2(float2* p, float2 v) { asm volatile("st.volatile.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory"); }
B2_D float2 g_ld_volatile_v2(const float2* p) { float2 v; asm volatile("ld.volatile.global.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory"); return v; }
B2_HD uint32_t gemm_xtag(uint32_t epoch, uint32_t seq) { const uint32_t t = epoch * 0x9E3779B1u + seq; return t ? t : 1u; }

// All-reduce of one chain's likelihood sums over the row shards (one warp; see GemmShardDev).  gbeta[0..Dx) and the returned
// loss are replaced by the sums over all ranks, added in rank order.  Every wait is bounded (abort code 4).
B2_D float gemm_shard_allreduce(const GemmShardDev& sh, int C, int chain, float* gbeta, float nll, int Dx, int Dxp, uint32_t tag, int parity,
                                unsigned int* abort_flag, long long spin_limit) {
    const int lane = threadIdx.x & 31;
    const size_t per_rank = (size_t)C * sh.stride;
    const size_t base = ((size_t)parity * kGemmMaxShards) * per_rank + (size_t)chain * sh.stride;
    const float tagf = __uint_as_float(tag);
    for (int d = lane; d <= Dxp; d += 32) {
        if (d >= Dx && d != Dxp) continue;
        const float mine = (d == Dxp) ? (nll + sh.nll_local_const) : gbeta[d];
        for (int q = 0; q < sh.count; ++q) g_st_sys_v2(sh.mail[q] + base + (size_t)sh.rank * per_rank + d, make_float2(mine, tagf));
    }
    const long long t0 = clock64();
    float nll_tot = 0.0f;
    for (int d = lane; d <= Dxp; d += 32) {
        if (d >= Dx && d != Dxp) continue;
        const float2* box = sh.mail[sh.rank] + base + d;
        float tot = 0.0f;
        for (int sr = 0; sr < sh.count; ++sr) {
            float2 v;
            unsigned int it = 0;
            while (true) {
                v = g_ld_volatile_v2(box + (size_t)sr * per_rank);
                if (__float_as_uint(v.y) == tag) break;
                if ((++it & 63u) == 0u) {
                    if (*(volatile unsigned int*)abort_flag) break;
                    if (clock64() - t0 > spin_limit) { atomicCAS(abort_flag, 0u, 4u); break; }
                }
            }
            tot += v.x;
        }
        if (d == Dxp) nll_tot = tot; else gbeta[d] = tot;
    }
    // the loss lives in the lane that owns index Dxp
    return __shfl_sync(0xFFFFFFFFu, nll_tot, Dxp & 31);
}

// Tick of every chain after a pass (one warp per chain): likelihood sums -> potential -> NUTS state machine -> next betas.
// first != 0: start of a run -- no gradient yet, only publish the betas of the chains that wait for one.
static __global__ void __launch_bounds__(128) k_gemm_tick(GemmParams gp, const GemmCtx* ctx, GemmSched* sched, FamilySpec fam, ChainCtl* ctl,
                                                           float* vecs, float* gtmp, float* gbeta, const float* nll_in, float* bimg, int* tile_count,
                                                           int C, int Dp, int first, GemmShardDev sh, float* dense) {
    __shared__ TickCfg s_cfg; __shared__ OutBufs s_out;
    {
        const int* src = (const int*)&ctx->cfg; int* dst = (int*)&s_cfg;
        for (int i = threadIdx.x; i < (int)(sizeof(TickCfg) / 4); i += blockDim.x) dst[i] = src[i];
        if (threadIdx.x == 0) s_out = ctx->out;
    }
    __syncthreads();
    const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chain >= C) return;
    ChainCtl c = ctl[chain];
    if (c.phase == PH_DONE) return;
    const int par = first ? 0 : ((sched->pass + 1) & 1);          // the tile counts this tick fills (k_gemm_sched flips `pass`)
    ChainVecs cv; cv.base = vecs + (size_t)chain * Dp; cv.field_stride = C * Dp;
    if (dense) cv.dense = dense + (size_t)chain * 4 * s_cfg.D * s_cfg.D;
    if (!first) {
        Tick t{s_cfg, c, cv, s_out, chain, C};
        float* gb = gbeta + (size_t)chain * gp.Dxp; float* g = gtmp + (size_t)chain * Dp;
        float nll = nll_in[chain];                 // (k_gemm_reduce summed the unit partials)
        if (sh.count > 1) {
            nll = gemm_shard_allreduce(sh, C, chain, gb, nll, fam.Dx, gp.Dxp, gemm_xtag(ctx->epoch, (uint32_t)sched->pass_in_run + 1u),
                                       sched->pass_in_run & 1, gp.abort_flag, 8 * gp.spin_limit);
            __syncwarp(); lane_sync();
        }
        float u;
        const long long tq0 = clock64();
        glm_finish(fam, cv.v(V_ZS), nll, gb, u, g);
        __syncwarp(); lane_sync();
        const long long tq1 = clock64();
        t.advance(u, g);
        __syncwarp(); lane_sync();               // the betas below are gathered across lanes from V_ZS
        if ((threadIdx.x & 31) == 0) ctl[chain] = c;
        if (chain == 0 && (threadIdx.x & 31) == 0) { sched->dbg[6] += (unsigned long long)(tq1 - tq0); sched->dbg[7] += (unsigned long long)(clock64() - tq1); }
    }
    if (c.phase != PH_DONE) {
        gemm_write_betas(fam, cv.v(V_ZS), bimg, gp.KB, chain);
        if ((threadIdx.x & 31) == 0) atomicAdd(&tile_count[par * gp.CT + chain / kGtChains], 1);
    }
}

// After the tick: the list of chain tiles that still need gradients, pass counters, and the WHILE condition.
static __global__ void k_gemm_sched(const GemmCtx* ctx, GemmSched* sched, int* tile_count, int* active_tiles, int CT, int first,
                                    cudaGraphConditionalHandle cond, int use_cond) {
    if (threadIdx.x == 0) {
        const int par = first ? 0 : ((sched->pass + 1) & 1);
        int n = 0;
        for (int t = 0; t < CT; ++t) {
            if (tile_count[par * CT + t] > 0) active_tiles[n++] = t;
            tile_count[(par ^ 1) * CT + t] = 0;                  // the buffer the next tick fills
        }
        sched->n_active = n;
        if (first) { sched->pass = 0; sched->pass_in_run = 0; }
        else { sched->pass += 1; sched->pass_in_run += 1; sched->passes_total += 1ull; }
        bool go = n > 0 && sched->abort_flag == 0u;
        if (ctx->max_passes > 0 && sched->pass_in_run >= ctx->max_passes) go = false;
        if (use_cond) cudaGraphSetConditional(cond, go ? 1u : 0u);
    }
}

// ---- potential hook: betas from given positions, all tiles active; then the potential from the sums ----------------
static __global__ void __launch_bounds__(128) k_gemm_hook_begin(FamilySpec fam, const float* z_in, float* bimg, int KB, int C, int CT,
                                                                 int* active_tiles, GemmSched* sched) {
    const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) { for (int t = 0; t < CT; ++t) active_tiles[t] = t; sched->n_active = CT; }
    if (chain >= C) return;
    gemm_write_betas(fam, z_in + (size_t)chain * fam.D, bimg, KB, chain);
}
static __global__ void __launch_bounds__(128) k_gemm_hook_finish(GemmParams gp, FamilySpec fam, const float* z_in, float* U, float* g_out,
                                                                  float* gbeta, const float* nll_in, int C, GemmShardDev sh, unsigned int epoch) {
    const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chain >= C) return;
    float* gb = gbeta + (size_t)chain * gp.Dxp;
    float nll = nll_in[chain];
    if (sh.count > 1) {
        nll = gemm_shard_allreduce(sh, C, chain, gb, nll, fam.Dx, gp.Dxp, gemm_xtag(epoch, 0x7FFFFFu), 0, gp.abort_flag, 8 * gp.spin_limit);
        __syncwarp(); lane_sync();
    }
    float u;
    glm_finish(fam, z_in + (size_t)chain * fam.D, nll, gb, u, g_out + (size_t)chain * fam.D);
    if ((threadIdx.x & 31) == 0) U[chain] = u;
}

}  // namespace b2
