// Per-chain NUTS/HMC state machine ("tick"): everything numpyro does between two evaluations of
// the potential's gradient, for one chain, executed by one warp (or by the host simulator).
//
// The reference expresses a transition as three nested lax.while_loops around value_and_grad
// (numpyro/infer/hmc_util.py:1155-1179 doubling, :999-1065 leaves, :961-981 checkpoint scan) and
// batches chains with vmap, so every chain idles on the slowest one at every level.  Here the
// loops are turned inside out: a chain is an explicit state machine that is handed one
// (potential, gradient) pair per tick and answers with the next position to evaluate.  Chains are
// therefore fully asynchronous -- each rolls straight from one transition into the next -- while
// the expensive gradient can be produced for all chains together by whichever kernel suits the
// model (in-warp for tiny models, one streaming pass over X for tall data, a GEMM for many chains).
//
// Reference functions folded into this file:
//   a2  velocity_verlet            hmc_util.py:262-311      -> leap_begin / leap_finish
//   a4  euclidean_kinetic_energy   hmc_util.py:1183-1200    -> kinetic()
//   a5  momentum_generator         hmc.py:92-110            -> draw_momentum()
//   a6  build_tree                 hmc_util.py:1088-1180    -> begin_transition / begin_doubling
//   a7  _iterative_build_subtree   hmc_util.py:984-1085, _build_basetree :851-894 -> on_leaf()
//   a8  _combine_tree + kernels    hmc_util.py:749-848      -> on_leaf() / finish_doubling()
//   a9  ckpt index + U-turn checks hmc_util.py:710-746,941-981
//   a10 _nuts_next / sample_kernel hmc.py:416-530, _hmc_next :364-414 -> finish_transition()
//   a11 warmup_adapter, dual_averaging, welford_covariance   hmc_util.py:60-239,518-707
//   a12 find_reasonable_step_size  hmc_util.py:314-384      -> PH_HEUR
//   a13 init_kernel / find_valid_initial_params  hmc.py:193-362, infer/util.py:366-508 -> PH_INIT
//   a15 fori_collect index arithmetic  numpyro/util.py:368-403 -> collect()
#pragma once
#include "common.cuh"
#include "detmath.cuh"
#include "prng.cuh"

namespace b2 {

constexpr int kMaxDepthAlloc = 12;     // checkpoint rows allocated per chain (max_tree_depth <= 12)
constexpr int kMaxSites = 8;

enum VecField {
    V_Z = 0, V_G, V_IMM, V_SQRTM, V_WF_MEAN, V_WF_M2,
    V_ZL, V_RL, V_GL, V_ZR, V_RR, V_GR, V_ZP, V_GP, V_RSUM,
    V_ZS, V_RS, V_GS, V_ZPS, V_GPS, V_RSUMS,
    V_R0,                               // momentum at the start of the transition (HMC / heuristic)
    V_EPS,                              // look-ahead: unit normal draws of the next transition's momentum (prefetch())
    V_CKPT_R,                           // kMaxDepthAlloc rows
    V_CKPT_RSUM = V_CKPT_R + kMaxDepthAlloc,
    V_TMP0 = V_CKPT_RSUM + kMaxDepthAlloc, V_TMP1,   // dense mass matrix: M^-1 r products, Welford deltas
    V_COUNT
};

enum Phase { PH_DONE = 0, PH_INIT = 1, PH_LEAF = 2, PH_HMC = 3, PH_HEUR = 4 };

// Per-chain scalars (mirrors HMCState hmc.py:31-48 + HMCAdaptState hmc_util.py:18-30 + the live
// TreeInfo scalars hmc_util.py:36-57).  Plain ints/floats so the host can read it back verbatim.
struct ChainCtl {
    // HMCState
    int32_t i; uint32_t key[2];
    float pe, energy;
    int32_t num_steps; float accept_prob, mean_accept_prob; int32_t diverging;
    // HMCAdaptState
    float step_size;
    float da_x_t, da_x_avg, da_g_avg, da_prox; int32_t da_t;
    int32_t mm_n, window_idx; uint32_t wa_key[2];
    // transition in flight
    int32_t phase;
    float eps, energy0; int32_t max_depth;
    uint32_t key_next[2], k_loop[2];
    int32_t depth, n_total, turning, t_div; float weight, sum_acc, prop_pe, prop_energy;
    int32_t going_right, n_sub, sub_div; float sub_weight, sub_sum_acc, sub_prop_pe, sub_prop_energy;
    uint32_t k_sub[2], k_fin[2];
    // init / heuristic / HMC scratch
    int32_t init_tries; uint32_t k_init[2];
    int32_t hmc_left, hmc_n;
    float heur_step; int32_t heur_dir, heur_last; uint32_t heur_key[2]; int32_t heur_at_init, heur_t;
    // output cursor + totals
    int32_t n_collected; int32_t init_failed;
    unsigned long long total_leapfrogs;
    // PRNG look-ahead (Tick::prefetch): values that depend only on keys, each tagged with the key it was
    // derived from; a consumer uses an entry only when its tag equals the key it is about to expand, so a
    // stale or missing entry can never change a result -- it only costs the recomputation.
    uint32_t pre_mask;                                           // populated entries (bits below)
    uint32_t pl_from[2], pl_ksub[2]; float pl_u;                 // bit 0: on_leaf, from k_sub
    uint32_t pf_from[2]; float pf_u;                             // bit 1: finish_doubling, from k_fin
    uint32_t pd_from[2][2], pd_kloop[2][2], pd_ksub[2][2], pd_kfin[2][2]; int32_t pd_right[2];   // bits 2, 3: begin_doubling,
                                                                 // from k_loop ([0] this tree, [1] first doubling of the next transition)
    uint32_t pt_from[2], pt_keynext[2], pt_ktr[2];               // bit 4: begin_transition, from key (+ V_EPS)
    uint32_t pw_from[2], pw_next[2], pw_kss[2];                  // bit 5: adapt_update, from wa_key
    uint32_t pre_hit[4], pre_miss[4];                            // statistics: leaf, doubling, transition, adaptation
};

struct TickCfg {
    int32_t D;
    int32_t num_warmup, total_iters;        // a chain is PH_DONE once i == total_iters
    int32_t md_warm, md_post;               // max_tree_depth (warm-up, post warm-up)
    float target_accept, init_step_size;
    int32_t adapt_step, adapt_mass, regularize, model_built, find_heuristic;
    int32_t algo;                           // 0 NUTS, 1 HMC
    int32_t hmc_num_steps; float traj_len;  // HMC: fixed num_steps (>0) or trajectory_length
    int32_t num_windows; int32_t window_end[16];
    int32_t collect_start, thinning, S;     // fori_collect: start_idx, thinning, collection size
    int32_t init_given; float init_radius;
    int32_t dense;                          // dense_mass=True: one [D, D] block over all sites (hmc.py:759-769)
    int32_t imm_given;                      // inverse_mass_matrix= was supplied (pre-loaded into V_IMM / the dense block)
    int32_t n_sites; int32_t site_off[kMaxSites], site_size[kMaxSites];   // latent sites, trace order
};

struct OutBufs {                            // all [C][S] except z [C][S][D]; null = not collected
    float* z; int32_t* diverging; int32_t* num_steps;
    float* accept_prob; float* mean_accept_prob; float* pe; float* energy; float* step_size;
};

struct ChainVecs {
    float* base; int field_stride;
    // dense_mass=True: this chain's [4][D][D] block -- M^-1 | M^1/2 (lower triangular) | Welford m2 | Cholesky workspace
    float* dense = nullptr;
    B2_HD float* v(int f) const { return base + (size_t)f * field_stride; }
};

B2_HD Key mk(const uint32_t* p) { Key k; k.a = p[0]; k.b = p[1]; return k; }
B2_HD void st(uint32_t* p, Key k) { p[0] = k.a; p[1] = k.b; }
B2_HD bool key_is(const uint32_t* p, Key k) { return p[0] == k.a && p[1] == k.b; }
B2_HD float clip_max1(float p) { return (p > 1.0f) ? 1.0f : p; }      // jnp.clip(p, None, 1): NaN kept

B2_HD float kinetic(int D, const float* imm, const float* r) {
    if (D > 64) return 0.5f * lane_sum_wide(D, [&](int d) { return (imm[d] * r[d]) * r[d]; });
    return 0.5f * lane_sum(D, [&](int d) { return (imm[d] * r[d]) * r[d]; });
}

// hmc_util.py:710-746
B2_HD bool is_turning(int D, const float* imm, const float* r_left, const float* r_right, const float* r_sum) {
    float l, r;
    auto f = [&](int d, float& a, float& b) {
        const float s = r_sum[d] - (r_left[d] + r_right[d]) / 2.0f;
        a = (imm[d] * r_left[d]) * s;
        b = (imm[d] * r_right[d]) * s;
    };
    if (D > 64) lane_sum2_wide(D, f, l, r); else lane_sum2(D, f, l, r);
    return (l <= 0.0f) || (r <= 0.0f);
}

// first half of velocity_verlet (hmc_util.py:300-304): r_half and the new position
B2_HD void leap_begin(int D, float eps, const float* imm, const float* z, const float* r, const float* g,
                      float* z_out, float* r_half_out) {
    const float half = 0.5f * eps;
    if (D > 64) {
        for_d_wide(D, [&](int d) { Vals v; v.x[0] = r[d]; v.x[1] = g[d]; v.x[2] = z[d]; v.x[3] = imm[d]; return v; },
                   [&](int d, const Vals& v) { const float rh = v.x[0] - half * v.x[1]; r_half_out[d] = rh; z_out[d] = v.x[2] + eps * (v.x[3] * rh); });
        return;
    }
    B2_FOR_D(d, D) {
        const float rh = r[d] - half * g[d];
        r_half_out[d] = rh;
        z_out[d] = z[d] + eps * (imm[d] * rh);
    }
}

// development only (-DB2_TICK_LAPS): clock64 laps of chain 0's tick, see scripts/exchange_probe.py
#if defined(__CUDACC__) && defined(B2_TICK_LAPS)
// accumulators in shared memory (cheap: no global round trip inside the measured code); [i] cycles, [16 + i] occurrences
__device__ inline unsigned long long* b2_lap_store() { __shared__ unsigned long long s_[34]; return s_; }
__host__ __device__ inline void b2_lapq(int i) {
#if defined(__CUDA_ARCH__)
    if (blockIdx.x == 0 && (threadIdx.x & 31u) == 0u) {
        unsigned long long* s_ = b2_lap_store();
        const unsigned long long t_ = (unsigned long long)clock64();
        if (i >= 0) { s_[i] += t_ - s_[32]; s_[16 + i] += 1ull; }
        s_[32] = t_;
    }
#endif
}
#define B2_LAPQ(i) b2_lapq(i)
#else
#define B2_LAPQ(i) do {} while (0)
#endif

// ---- dense mass matrix kernels (out of line: rarely taken, and they must not take the Tick object -- its ChainCtl lives in
//      registers in the streaming engine) -----------------------------------------------------------------------------------
// out = A r for a row-major [D][D] matrix: row i (lane i mod 32) accumulates left to right from +0, one rounding per operation
B2_HD_COLD void dense_matvec(int Dn, const float* A, const float* r, float* out) {
    lane_sync();
    B2_FOR_D(i, Dn) {
        const float* row = A + (size_t)i * Dn;
        float a = 0.0f;
        for (int j = 0; j < Dn; ++j) a = a + row[j] * r[j];
        out[i] = a;
    }
    lane_sync();
}
// (mass_matrix_sqrt, mass_matrix_sqrt_inv) of the dense M^-1 in block 0 of `dense` (hmc_util.py:228-233, 499-509):
//   Lc = cholesky(sym(M^-1[::-1, ::-1])) -> block 3;  tril_inv[i][j] = Lc[D-1-j][D-1-i];  M^1/2 = tril_inv^-1 -> block 1.
// Element order as oracle/adapt.py cholesky_lower / solve_lower_identity (k ascending, one rounding per operation).
B2_HD_COLD void dense_roots_of(int Dn, float* dense, float* tmp) {
    const size_t DD = (size_t)Dn * Dn;
    const float* A = dense; float* W = dense + 3 * DD; float* X = dense + DD;
    for (int j = 0; j < Dn; ++j) {
        lane_sync();
        B2_FOR_D(ii, Dn - j) {
            const int i = j + ii;
            const int ri = Dn - 1 - i, rj = Dn - 1 - j;
            float s = (A[(size_t)ri * Dn + rj] + A[(size_t)rj * Dn + ri]) / 2.0f;
            for (int k = 0; k < j; ++k) s = s - W[(size_t)i * Dn + k] * W[(size_t)j * Dn + k];
            tmp[i] = s;
        }
        lane_sync();
        const float sj = tmp[j];
        const float piv = (sj > 0.0f) ? sqrtf(sj) : f_nan();
        B2_FOR_D(ii, Dn - j) {
            const int i = j + ii;
            W[(size_t)i * Dn + j] = (i == j) ? piv : tmp[i] / piv;
        }
        B2_FOR_D(ii, j) W[(size_t)ii * Dn + j] = 0.0f;           // (strict upper part: zero)
    }
    lane_sync();
    B2_FOR_D(cidx, Dn) {                                         // forward substitution, one column per lane
        for (int i = 0; i < Dn; ++i) {
            float s = (i == cidx) ? 1.0f : 0.0f;
            for (int k = 0; k < i; ++k) s = s - W[(size_t)(Dn - 1 - k) * Dn + (Dn - 1 - i)] * X[(size_t)k * Dn + cidx];
            X[(size_t)i * Dn + cidx] = s / W[(size_t)(Dn - 1 - i) * Dn + (Dn - 1 - i)];
        }
    }
    lane_sync();
}

// DENSE_OK = false compiles the dense-mass branches away (the streaming engine: its tick shares one register allocation with the
// sweep, and the out-of-line dense kernels cost the sweep ~4 % through spills around the call sites).
template <bool DENSE_OK>
struct TickT {
    const TickCfg& cfg;
    ChainCtl& c;
    ChainVecs vs;
    OutBufs out;
    int chain, C;

    B2_HD float* v(int f) const { return vs.v(f); }
    B2_HD bool dense_on() const { return DENSE_OK && cfg.dense != 0; }
    B2_HD int D() const { return cfg.D; }

    B2_HD void copy(int dst, int src) const {
        float* a = v(dst); const float* b = v(src);
        if (cfg.D > 64) { for_d_wide(cfg.D, [&](int d) { Vals x; x.x[0] = b[d]; return x; }, [&](int d, const Vals& x) { a[d] = x.x[0]; }); return; }
        B2_FOR_D(d, cfg.D) a[d] = b[d];
    }

    // ---------------------------------------------------------------- mass matrix (diagonal vector or dense block)
    // Dense (hmc_util.py:192-237, 439-515, 726-728, 1193-1194): every product with M^-1 / M^1/2 is a
    // [D, D] x [D] product in which row i (lane i mod 32) accumulates its terms left to right from +0, one rounding per
    // multiplication and per addition -- the order oracle/tree.py imm_apply restates.
    B2_HD float* dmat(int k) const { return vs.dense + (size_t)k * cfg.D * cfg.D; }
    B2_HD void matvec(const float* A, const float* r, float* out) const { dense_matvec(cfg.D, A, r, out); }
    B2_HD float kin(const float* r) const {                                   // euclidean_kinetic_energy
        if (!dense_on()) return kinetic(cfg.D, v(V_IMM), r);
        float* t = v(V_TMP0);
        matvec(dmat(0), r, t);
        return 0.5f * lane_sum(cfg.D, [&](int d) { return t[d] * r[d]; });
    }
    B2_HD bool turning_between(const float* r_left, const float* r_right, const float* r_sum) const {   // _is_turning
        if (!dense_on()) return is_turning(cfg.D, v(V_IMM), r_left, r_right, r_sum);
        float *vl = v(V_TMP0), *vr = v(V_TMP1);
        matvec(dmat(0), r_left, vl); matvec(dmat(0), r_right, vr);
        float l, r;
        lane_sum2(cfg.D, [&](int d, float& a, float& b) {
            const float s = r_sum[d] - (r_left[d] + r_right[d]) / 2.0f;
            a = vl[d] * s; b = vr[d] * s;
        }, l, r);
        return (l <= 0.0f) || (r <= 0.0f);
    }
    B2_HD void leap(float eps, const float* z, const float* r, const float* g, float* z_out, float* r_half_out) const {
        if (!dense_on()) { leap_begin(cfg.D, eps, v(V_IMM), z, r, g, z_out, r_half_out); return; }
        const float half = 0.5f * eps;
        B2_FOR_D(d, cfg.D) r_half_out[d] = r[d] - half * g[d];
        float* t = v(V_TMP0);
        matvec(dmat(0), r_half_out, t);
        B2_FOR_D(d, cfg.D) z_out[d] = z[d] + eps * t[d];
    }
    // r = S eps with S = M^1/2 (momentum_generator) or, for the step-size heuristic, S = M^-1 (the reference's quirk)
    B2_HD void scale_normals(int which, const float* eps, float* r) const {
        if (dense_on()) { matvec(dmat(which), eps, r); return; }
        const float* sc = v(which ? V_SQRTM : V_IMM);
        B2_FOR_D(d, cfg.D) r[d] = sc[d] * eps[d];
    }
    // (mass_matrix_sqrt, mass_matrix_sqrt_inv) of the dense M^-1 in dmat(0) (hmc_util.py:228-233, 499-509):
    //   Lc = cholesky(sym(M^-1[::-1, ::-1])) -> workspace dmat(3);  tril_inv[i][j] = Lc[D-1-j][D-1-i];  M^1/2 = tril_inv^-1
    // Element order as oracle/adapt.py cholesky_lower / solve_lower_identity (k ascending, one rounding per operation).
    B2_HD void dense_roots() const { dense_roots_of(cfg.D, vs.dense, v(V_TMP0)); }

    // ---------------------------------------------------------------- momentum (hmc.py:92-110)
    B2_HD void draw_momentum(Key k, float* r) const {
        if (cfg.model_built) k = split_at(k, 0);           // one-block dict -> split(key, 1)[0]
        if (dense_on()) {
            float* en = v(V_TMP1);
            B2_FOR_D(d, cfg.D) en[d] = normal_at(k, (uint32_t)d);
            scale_normals(1, en, r);
            return;
        }
        const float* sm = v(V_SQRTM);
        B2_FOR_D(d, cfg.D) r[d] = sm[d] * normal_at(k, (uint32_t)d);
    }

    // ---------------------------------------------------------------- init (infer/util.py:417-481)
    B2_HD void init_draw() {
        Key key = mk(c.k_init);
        Key sub = split_at(key, 1); key = split_at(key, 0);
        float* z = v(V_ZS);
        for (int s = 0; s < cfg.n_sites; ++s) {
            const int off = cfg.site_off[s], n = cfg.site_size[s];
            B2_FOR_D(j, n) z[off + j] = uniform_at(sub, (uint32_t)j, -cfg.init_radius, cfg.init_radius);
            sub = split_at(key, 1); key = split_at(key, 0);
        }
        st(c.k_init, key);
    }

    // Called once per chain before the first gradient.  ``chain_key`` is the row of
    // split(user_key, C) (mcmc.py:670-671); z0 != null means init_params were supplied.
    B2_HD void begin(Key chain_key, const float* z0) {
        Key rng = split_at(chain_key, 0), k_init = split_at(chain_key, 1);      // hmc.py:744-750
        st(c.key, rng); st(c.k_init, k_init);
        c.i = 0; c.init_tries = 0; c.init_failed = 0; c.n_collected = 0; c.total_leapfrogs = 0ull;
        c.num_steps = 0; c.accept_prob = 0.0f; c.mean_accept_prob = 0.0f; c.diverging = 0;
        if (z0) { float* z = v(V_ZS); B2_FOR_D(d, cfg.D) z[d] = z0[d]; }
        else init_draw();
        c.phase = PH_INIT;
    }

    // ---------------------------------------------------------------- init_kernel (hmc.py:193-362)
    B2_HD void finish_init(float u, const float* g) {
        copy(V_Z, V_ZS);
        float* gg = v(V_G);
        B2_FOR_D(d, cfg.D) gg[d] = g[d];
        c.pe = u;
        const Key rng = mk(c.key);
        const Key k_hmc = split_at(rng, 0), k_wa = split_at(rng, 1), k_mom = split_at(rng, 2);   // :335
        const Key wk = split_at(k_wa, 0), k_ss = split_at(k_wa, 1);                           // hmc_util.py:565
        st(c.key, k_hmc); st(c.wa_key, wk);
        float* imm = v(V_IMM); float* sm = v(V_SQRTM); float* mean = v(V_WF_MEAN); float* m2 = v(V_WF_M2);
        // _initialize_mass_matrix (hmc_util.py:439-515): identity, or the supplied matrix and its roots
        if (dense_on()) {
            const int Dn = cfg.D;
            float *A = dmat(0), *S = dmat(1), *M2 = dmat(2), *W = dmat(3);
            B2_FOR_D(i, Dn) {
                for (int j = 0; j < Dn; ++j) {
                    const float id = (i == j) ? 1.0f : 0.0f;
                    const size_t o = (size_t)i * Dn + j;
                    if (!cfg.imm_given) A[o] = id;
                    S[o] = id; W[o] = id; M2[o] = 0.0f;
                }
                mean[i] = 0.0f;
            }
            if (cfg.imm_given) dense_roots();
        } else {
            B2_FOR_D(d, cfg.D) {
                if (!cfg.imm_given) { imm[d] = 1.0f; sm[d] = 1.0f; }
                else sm[d] = 1.0f / sqrtf(imm[d]);
                mean[d] = 0.0f; m2[d] = 0.0f;
            }
        }
        c.mm_n = 0; c.window_idx = 0;
        c.step_size = cfg.init_step_size;
        // initial energy uses a throw-away momentum (hmc.py:340-343)
        float* r = v(V_R0);
        draw_momentum(k_mom, r);
        c.energy = c.pe + kin(r);
        if (cfg.adapt_step && cfg.find_heuristic) { heur_begin(k_ss, 1); return; }
        da_reinit(d_log(10.0f * c.step_size));                                               // :576
        begin_transition();
    }

    B2_HD void da_reinit(float prox) { c.da_x_t = 0.0f; c.da_x_avg = 0.0f; c.da_g_avg = 0.0f; c.da_t = 0; c.da_prox = prox; }

    // ---------------------------------------------------------------- find_reasonable_step_size
    // hmc_util.py:314-384 as a resumable loop: each trial needs one leapfrog from (z, g).
    B2_HD void heur_begin(Key k, int at_init) {
        c.heur_step = c.step_size; c.heur_dir = 0; c.heur_last = 0; st(c.heur_key, k); c.heur_at_init = at_init;
        heur_next();
    }
    B2_HD void heur_next() {
        const float tiny = 1.17549435e-38f, fmax = 3.40282347e+38f;
        const bool not_small = (c.heur_step > tiny) || (c.heur_dir >= 0);
        const bool not_large = (c.heur_step < fmax) || (c.heur_dir <= 0);
        const bool go = not_small && not_large && ((c.heur_last == 0) || (c.heur_dir == c.heur_last));
        if (!go) { heur_done(); return; }
        const Key key = mk(c.heur_key);
        const Key k_mom = split_at(key, 1); st(c.heur_key, split_at(key, 0));
        const float scale = (c.heur_dir > 0) ? 2.0f : ((c.heur_dir < 0) ? 0.5f : 1.0f);
        c.heur_step = scale * c.heur_step;
        // NB: the reference hands *inverse_mass_matrix* to momentum_generator here
        // (hmc_util.py:355 vs hmc.py:92), so r = M^-1 * eps, not M^1/2 * eps.
        float* r0 = v(V_R0);
        { Key km = cfg.model_built ? split_at(k_mom, 0) : k_mom;
          if (dense_on()) { float* en = v(V_TMP1); B2_FOR_D(d, cfg.D) en[d] = normal_at(km, (uint32_t)d); scale_normals(0, en, r0); }
          else { const float* imm = v(V_IMM); B2_FOR_D(d, cfg.D) r0[d] = imm[d] * normal_at(km, (uint32_t)d); } }
        leap(c.heur_step, v(V_Z), r0, v(V_G), v(V_ZS), v(V_RS));
        c.phase = PH_HEUR;
    }
    B2_HD void on_heur(float u, const float* g) {
        const float half = 0.5f * c.heur_step;
        float* rs = v(V_RS);
        B2_FOR_D(d, cfg.D) rs[d] = rs[d] - half * g[d];
        const float e_cur = kin(v(V_R0)) + c.pe;
        const float e_new = kin(rs) + u;
        const float delta = e_new - e_cur;
        const int new_dir = (d_log(0.8f) < -delta) ? 1 : -1;
        c.heur_last = c.heur_dir; c.heur_dir = new_dir;
        heur_next();
    }
    B2_HD void heur_done() {
        c.step_size = c.heur_step;
        if (c.heur_at_init) da_reinit(d_log(10.0f * c.step_size));
        else { da_reinit(d_log(10.0f) + d_log(c.step_size)); collect(c.heur_t); }
        begin_transition();
    }

    // ---------------------------------------------------------------- sample_kernel (hmc.py:459-530)
    B2_HD void begin_transition() {
        if (c.i >= cfg.total_iters) { c.phase = PH_DONE; return; }
        const Key key = mk(c.key);
        Key k_tr;
        float* r = v(V_R0);
        if ((c.pre_mask & 16u) && key_is(c.pt_from, key)) {          // looked ahead: same splits, same normals
            st(c.key_next, mk(c.pt_keynext)); k_tr = mk(c.pt_ktr);
            scale_normals(1, v(V_EPS), r); c.pre_hit[2] += 1u;
        } else { c.pre_miss[2] += 1u;
            const Key k_mom = split_at(key, 1);
            k_tr = split_at(key, 2);
            st(c.key_next, split_at(key, 0));
            draw_momentum(k_mom, r);
        }
        c.eps = c.step_size;
        c.energy0 = c.pe + kin(r);
        if (cfg.algo == 1) { hmc_begin(k_tr); return; }
        // build_tree root (hmc_util.py:1129-1153)
        const float* z = v(V_Z); const float* g = v(V_G);
        float *zl = v(V_ZL), *rl = v(V_RL), *gl = v(V_GL), *zr = v(V_ZR), *rr = v(V_RR), *gr = v(V_GR);
        float *zp = v(V_ZP), *gp = v(V_GP), *rs = v(V_RSUM);
        for_d_wide(cfg.D, [&](int d) { Vals x; x.x[0] = z[d]; x.x[1] = g[d]; x.x[2] = r[d]; return x; },
                   [&](int d, const Vals& x) {
            const float zz = x.x[0], gg = x.x[1], rv = x.x[2];
            zl[d] = zz; zr[d] = zz; zp[d] = zz; gl[d] = gg; gr[d] = gg; gp[d] = gg; rl[d] = rv; rr[d] = rv; rs[d] = rv;
        });
        c.depth = 0; c.n_total = 0; c.turning = 0; c.t_div = 0; c.weight = 0.0f; c.sum_acc = 0.0f;
        c.prop_pe = c.pe; c.prop_energy = c.energy0;
        c.max_depth = (c.i < cfg.num_warmup) ? cfg.md_warm : cfg.md_post;
        st(c.k_loop, k_tr);
        begin_doubling();            // max_tree_depth >= 1 is validated on the host (keeps the call graph acyclic)
    }

    // one iteration of build_tree's while loop up to the first leapfrog (hmc_util.py:1159-1162, :920)
    B2_HD void begin_doubling() {
        const Key k = mk(c.k_loop);
        if ((c.pre_mask & 4u) && key_is(c.pd_from[0], k)) {          // looked ahead (static indices: ChainCtl stays in registers)
            st(c.k_loop, mk(c.pd_kloop[0])); c.going_right = c.pd_right[0];
            st(c.k_sub, mk(c.pd_ksub[0])); st(c.k_fin, mk(c.pd_kfin[0])); c.pre_hit[1] += 1u;
        } else if ((c.pre_mask & 8u) && key_is(c.pd_from[1], k)) {
            st(c.k_loop, mk(c.pd_kloop[1])); c.going_right = c.pd_right[1];
            st(c.k_sub, mk(c.pd_ksub[1])); st(c.k_fin, mk(c.pd_kfin[1])); c.pre_hit[1] += 1u;
        } else { c.pre_miss[1] += 1u;
            const Key k_dir = split_at(k, 1), k_dbl = split_at(k, 2);
            st(c.k_loop, split_at(k, 0));
            c.going_right = (uniform01_at(k_dir, 0) < 0.5f) ? 1 : 0;
            st(c.k_sub, split_at(k_dbl, 0)); st(c.k_fin, split_at(k_dbl, 1));
        }
        c.n_sub = 0; c.sub_div = 0; c.sub_weight = 0.0f; c.sub_sum_acc = 0.0f;
        const float e = c.going_right ? c.eps : -c.eps;
        if (c.going_right) leap(e, v(V_ZR), v(V_RR), v(V_GR), v(V_ZS), v(V_RS));
        else leap(e, v(V_ZL), v(V_RL), v(V_GL), v(V_ZS), v(V_RS));
        c.phase = PH_LEAF;
    }

#if defined(__CUDA_ARCH__)
    // on_leaf for D <= 64 (two elements per lane) with the chain vectors held in registers: every vector is loaded once
    // (all loads in flight together), the leaf's arithmetic runs on registers, results are stored once.  Same expressions in
    // the same order per element and per reduction as the generic on_leaf below => identical bits (the GPU parity tests
    // compare whole runs with the oracle, which follows the generic form).
    __device__ __forceinline__ void on_leaf_fused(float u, const float* g) {
        const int Dn = cfg.D;
        const int lane = (int)(threadIdx.x & 31u);
        const int d0 = lane, d1 = lane + 32;
        const bool a0 = d0 < Dn, a1 = d1 < Dn;
        const float e = c.going_right ? c.eps : -c.eps;
        const float half = 0.5f * e;
        const float* imm = v(V_IMM);
        float *zs = v(V_ZS), *rs = v(V_RS), *gs = v(V_GS);
        float *zps = v(V_ZPS), *gps = v(V_GPS), *rsum_s = v(V_RSUMS);
        B2_LAPQ(-1);
        float g0 = 0.0f, g1 = 0.0f, r0 = 0.0f, r1 = 0.0f, i0 = 0.0f, i1 = 0.0f, z0 = 0.0f, z1 = 0.0f, q0 = 0.0f, q1 = 0.0f;
        if (a0) { g0 = g[d0]; r0 = rs[d0]; i0 = imm[d0]; z0 = zs[d0]; q0 = rsum_s[d0]; }
        if (a1) { g1 = g[d1]; r1 = rs[d1]; i1 = imm[d1]; z1 = zs[d1]; q1 = rsum_s[d1]; }
        r0 = r0 - half * g0; r1 = r1 - half * g1;
        c.total_leapfrogs += 1ull;
        float pk = 0.0f;                                           // kinetic(): lane partial in element order, then the butterfly
        if (a0) pk = pk + (i0 * r0) * r0;
        if (a1) pk = pk + (i1 * r1) * r1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) pk = pk + __shfl_xor_sync(0xFFFFFFFFu, pk, off);
        const float energy_new = u + 0.5f * pk;
        B2_LAPQ(0);
        float delta = energy_new - c.energy0;
        if (is_nan(delta)) delta = f_inf();
        const float leaf_w = -delta;
        const int leaf_div = (delta > 1000.0f) ? 1 : 0;
        // The three exponentials of a leaf -- exp(-delta) (accept prob), exp(-(w_leaf - w_sub)) (inside expit) and
        // exp(-|w_sub - w_leaf|) (inside logaddexp) -- run on three lanes at once instead of one after the other.
        const float x_mix = leaf_w - c.sub_weight, d_mix = c.sub_weight - leaf_w;
        const int sel = lane % 3;
        const float e_mine = d_exp(sel == 0 ? -delta : (sel == 1 ? -x_mix : -fabsf(d_mix)));
        const float e_acc = __shfl_sync(0xFFFFFFFFu, e_mine, 0), e_x = __shfl_sync(0xFFFFFFFFu, e_mine, 1),
                    e_abs = __shfl_sync(0xFFFFFFFFu, e_mine, 2);
        const float leaf_acc = clip_max1(e_acc);
        const Key ks = mk(c.k_sub);
        float u_leaf;
        if ((c.pre_mask & 1u) && key_is(c.pl_from, ks)) { st(c.k_sub, mk(c.pl_ksub)); u_leaf = c.pl_u; c.pre_hit[0] += 1u; }
        else { c.pre_miss[0] += 1u; const Key k_leaf = split_at(ks, 1); st(c.k_sub, split_at(ks, 0)); u_leaf = uniform01_at(k_leaf, 0); }
        B2_LAPQ(1);
        const int leaf_idx = c.n_sub;
        bool store_prop;
        if (leaf_idx == 0) {
            store_prop = true; q0 = r0; q1 = r1;
            c.sub_prop_pe = u; c.sub_prop_energy = energy_new;
            c.sub_weight = leaf_w; c.sub_sum_acc = leaf_acc;
        } else {                                                   // _combine_tree, uniform kernel
            const float p = 1.0f / (1.0f + e_x);                   // d_expit(leaf_w - c.sub_weight)
            store_prop = u_leaf < p;
            if (store_prop) { c.sub_prop_pe = u; c.sub_prop_energy = energy_new; }
            q0 = q0 + r0; q1 = q1 + r1;
            {   // d_logaddexp(c.sub_weight, leaf_w)
                const float a_ = c.sub_weight, b_ = leaf_w;
                c.sub_weight = is_nan(d_mix) ? (a_ + b_) : (((a_ >= b_) ? a_ : b_) + d_log1p(e_abs));
            }
            c.sub_sum_acc = c.sub_sum_acc + leaf_acc;
        }
        if (a0) { gs[d0] = g0; rsum_s[d0] = q0; if (store_prop) { zps[d0] = z0; gps[d0] = g0; } }
        if (a1) { gs[d1] = g1; rsum_s[d1] = q1; if (store_prop) { zps[d1] = z1; gps[d1] = g1; } }
        B2_LAPQ(2);
        c.sub_div = leaf_div;
        c.n_sub = leaf_idx + 1;
        const uint32_t n = (uint32_t)leaf_idx;
        const int idx_max = popc32(n >> 1);
        const int idx_min = idx_max - popc32((~n & (n + 1u)) - 1u) + 1;
        if ((leaf_idx & 1) == 0) {
            float* cr = v(V_CKPT_R + idx_max); float* cs = v(V_CKPT_RSUM + idx_max);
            if (a0) { cr[d0] = r0; cs[d0] = q0; }
            if (a1) { cr[d1] = r1; cs[d1] = q1; }
        }
        bool sub_turning = false;
        for (int i = idx_max; i >= idx_min && !sub_turning; --i) {
            const float* cr = v(V_CKPT_R + i); const float* cs = v(V_CKPT_RSUM + i);
            float pl = 0.0f, pr = 0.0f;
            if (a0) {
                const float c_r = cr[d0], c_s = cs[d0];
                const float sub = (q0 - c_s) + c_r;
                const float sm = sub - (c_r + r0) / 2.0f;
                pl = pl + (i0 * c_r) * sm; pr = pr + (i0 * r0) * sm;
            }
            if (a1) {
                const float c_r = cr[d1], c_s = cs[d1];
                const float sub = (q1 - c_s) + c_r;
                const float sm = sub - (c_r + r1) / 2.0f;
                pl = pl + (i1 * c_r) * sm; pr = pr + (i1 * r1) * sm;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                pl = pl + __shfl_xor_sync(0xFFFFFFFFu, pl, off);
                pr = pr + __shfl_xor_sync(0xFFFFFFFFu, pr, off);
            }
            sub_turning = (pl <= 0.0f) || (pr <= 0.0f);
        }
        B2_LAPQ(3);
        if (c.n_sub < (1 << c.depth) && !sub_turning && !c.sub_div) {      // next leaf of this subtree: leap_begin in registers
            const float h0 = r0 - half * g0, h1 = r1 - half * g1;
            if (a0) { rs[d0] = h0; zs[d0] = z0 + e * (i0 * h0); }
            if (a1) { rs[d1] = h1; zs[d1] = z1 + e * (i1 * h1); }
            B2_LAPQ(4);
            return;
        }
        if (a0) rs[d0] = r0;
        if (a1) rs[d1] = r1;
        finish_doubling(sub_turning);
        B2_LAPQ(5);
    }
#endif

#if defined(__CUDA_ARCH__)
    // "Where does this chain want its next gradient?" -- answered WITHOUT touching the chain state (streaming regime, D <= 64).
    // After a gradient arrives, the only thing the rest of the grid waits for is the chain's next position; everything
    // else a tick does (proposal bookkeeping, tree edges, root of the next tree, sample collection, ...) can run while the
    // grid already sweeps X again.  peek_next() evaluates just the decisions that select the next position -- the same
    // expressions, in the same order per element and per reduction, as on_leaf_fused / finish_doubling /
    // finish_transition_common / begin_transition / begin_doubling -- and writes that position to `zout`.  The caller
    // publishes it, then runs advance() as before, which recomputes the identical position into V_ZS (the streaming engine
    // counts mismatches in debug builds of the tests; there must be none).  Returns false when the case is not covered
    // (warm-up adaptation at a transition end, a PRNG look-ahead miss, the chain's last transition, HMC, D > 64): the
    // caller then falls back to "advance first, publish afterwards".
    __device__ __forceinline__ bool peek_next(float u, const float* g, float* zout) const {
        const int Dn = cfg.D;
        if (cfg.algo != 0 || c.phase != PH_LEAF || Dn > 64 || dense_on()) return false;
        const int lane = (int)(threadIdx.x & 31u);
        const int d0 = lane, d1 = lane + 32;
        const bool a0 = d0 < Dn, a1 = d1 < Dn;
        const float e = c.going_right ? c.eps : -c.eps;
        const float half = 0.5f * e;
        const float* imm = v(V_IMM);
        const float *zs = v(V_ZS), *rs = v(V_RS), *rsum_s = v(V_RSUMS);
        float g0 = 0.0f, g1 = 0.0f, r0 = 0.0f, r1 = 0.0f, i0 = 0.0f, i1 = 0.0f, z0 = 0.0f, z1 = 0.0f, q0 = 0.0f, q1 = 0.0f;
        if (a0) { g0 = g[d0]; r0 = rs[d0]; i0 = imm[d0]; z0 = zs[d0]; q0 = rsum_s[d0]; }
        if (a1) { g1 = g[d1]; r1 = rs[d1]; i1 = imm[d1]; z1 = zs[d1]; q1 = rsum_s[d1]; }
        r0 = r0 - half * g0; r1 = r1 - half * g1;
        float pk = 0.0f;
        if (a0) pk = pk + (i0 * r0) * r0;
        if (a1) pk = pk + (i1 * r1) * r1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) pk = pk + __shfl_xor_sync(0xFFFFFFFFu, pk, off);
        const float energy_new = u + 0.5f * pk;
        float delta = energy_new - c.energy0;
        if (is_nan(delta)) delta = f_inf();
        const float leaf_w = -delta;
        const bool leaf_div = delta > 1000.0f;
        const int leaf_idx = c.n_sub;
        if (leaf_idx == 0) { q0 = r0; q1 = r1; } else { q0 = q0 + r0; q1 = q1 + r1; }
        const uint32_t n = (uint32_t)leaf_idx;
        const int idx_max = popc32(n >> 1);
        const int idx_min = idx_max - popc32((~n & (n + 1u)) - 1u) + 1;
        bool sub_turning = false;
        for (int i = idx_max; i >= idx_min && !sub_turning; --i) {
            const float* cr = v(V_CKPT_R + i); const float* cs = v(V_CKPT_RSUM + i);
            float pl = 0.0f, pr = 0.0f;
            if (a0) {
                const float c_r = cr[d0], c_s = cs[d0];
                const float sub = (q0 - c_s) + c_r;
                const float sm = sub - (c_r + r0) / 2.0f;
                pl = pl + (i0 * c_r) * sm; pr = pr + (i0 * r0) * sm;
            }
            if (a1) {
                const float c_r = cr[d1], c_s = cs[d1];
                const float sub = (q1 - c_s) + c_r;
                const float sm = sub - (c_r + r1) / 2.0f;
                pl = pl + (i1 * c_r) * sm; pr = pr + (i1 * r1) * sm;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                pl = pl + __shfl_xor_sync(0xFFFFFFFFu, pl, off);
                pr = pr + __shfl_xor_sync(0xFFFFFFFFu, pr, off);
            }
            sub_turning = (pl <= 0.0f) || (pr <= 0.0f);
        }
        if (leaf_idx + 1 < (1 << c.depth) && !sub_turning && !leaf_div) {          // ---- next leaf of this subtree
            const float h0 = r0 - half * g0, h1 = r1 - half * g1;
            if (a0) zout[d0] = z0 + e * (i0 * h0);
            if (a1) zout[d1] = z1 + e * (i1 * h1);
            return true;
        }
        // ---- the subtree is complete: tree-level U-turn test of finish_doubling
        bool turning = sub_turning;
        if (!sub_turning) {
            const float* other = v(c.going_right ? V_RL : V_RR);
            const float* rsum = v(V_RSUM);
            float pl = 0.0f, pr = 0.0f;
            if (a0) {
                const float o0 = other[d0], t0 = rsum[d0] + q0;
                const float rl_ = c.going_right ? o0 : r0, rr_ = c.going_right ? r0 : o0;
                const float sm = t0 - (rl_ + rr_) / 2.0f;
                pl = pl + (i0 * rl_) * sm; pr = pr + (i0 * rr_) * sm;
            }
            if (a1) {
                const float o1 = other[d1], t1 = rsum[d1] + q1;
                const float rl_ = c.going_right ? o1 : r1, rr_ = c.going_right ? r1 : o1;
                const float sm = t1 - (rl_ + rr_) / 2.0f;
                pl = pl + (i1 * rl_) * sm; pr = pr + (i1 * rr_) * sm;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                pl = pl + __shfl_xor_sync(0xFFFFFFFFu, pl, off);
                pr = pr + __shfl_xor_sync(0xFFFFFFFFu, pr, off);
            }
            turning = (pl <= 0.0f) || (pr <= 0.0f);
        }
        if (c.depth + 1 < c.max_depth && !turning && !leaf_div) {                   // ---- next doubling of this tree
            if (!((c.pre_mask & 4u) && key_is(c.pd_from[0], mk(c.k_loop)))) return false;
            const bool right = c.pd_right[0] != 0;
            const float e2 = right ? c.eps : -c.eps;
            const float half2 = 0.5f * e2;
            float ez0 = z0, ez1 = z1, er0 = r0, er1 = r1, eg0 = g0, eg1 = g1;           // same direction: the leaf just built is the edge
            if (right != (c.going_right != 0)) {
                const float *ze = v(right ? V_ZR : V_ZL), *re = v(right ? V_RR : V_RL), *ge = v(right ? V_GR : V_GL);
                if (a0) { ez0 = ze[d0]; er0 = re[d0]; eg0 = ge[d0]; }
                if (a1) { ez1 = ze[d1]; er1 = re[d1]; eg1 = ge[d1]; }
            }
            const float h0 = er0 - half2 * eg0, h1 = er1 - half2 * eg1;
            if (a0) zout[d0] = ez0 + e2 * (i0 * h0);
            if (a1) zout[d1] = ez1 + e2 * (i1 * h1);
            return true;
        }
        // ---- the transition ends: which proposal becomes the new state, then the first leapfrog of the next transition
        if (c.i < cfg.num_warmup || c.i + 1 >= cfg.total_iters) return false;         // adaptation / the chain's last transition
        if (!((c.pre_mask & 2u) && key_is(c.pf_from, mk(c.k_fin)))) return false;
        if (!((c.pre_mask & 16u) && key_is(c.pt_from, mk(c.key_next)))) return false;
        if (!((c.pre_mask & 8u) && key_is(c.pd_from[1], mk(c.pt_ktr)))) return false;
        bool store_prop = true;
        float sw = leaf_w;                                                           // sub_weight after this leaf
        if (leaf_idx != 0) {
            if (!((c.pre_mask & 1u) && key_is(c.pl_from, mk(c.k_sub)))) return false;
            const float x_mix = leaf_w - c.sub_weight, d_mix = c.sub_weight - leaf_w;
            const float e_mine = d_exp((lane & 1) ? -fabsf(d_mix) : -x_mix);          // two exponentials on two lanes
            const float e_x = __shfl_sync(0xFFFFFFFFu, e_mine, 0), e_abs = __shfl_sync(0xFFFFFFFFu, e_mine, 1);
            const float pp = 1.0f / (1.0f + e_x);
            store_prop = c.pl_u < pp;
            const float a_ = c.sub_weight, b_ = leaf_w;
            sw = is_nan(d_mix) ? (a_ + b_) : (((a_ >= b_) ? a_ : b_) + d_log1p(e_abs));
        }
        float p2 = clip_max1(d_exp(sw - c.weight));
        if (sub_turning || leaf_div) p2 = 0.0f;
        const bool take = c.pf_u < p2;
        float zn0 = z0, zn1 = z1, gn0 = g0, gn1 = g1;                                 // take && store_prop: this leaf
        if (!(take && store_prop)) {
            const float *zq = v(take ? V_ZPS : V_ZP), *gq = v(take ? V_GPS : V_GP);
            if (a0) { zn0 = zq[d0]; gn0 = gq[d0]; }
            if (a1) { zn1 = zq[d1]; gn1 = gq[d1]; }
        }
        const float *sm = v(V_SQRTM), *en = v(V_EPS);
        const float e3 = c.pd_right[1] ? c.step_size : -c.step_size;
        const float half3 = 0.5f * e3;
        if (a0) { const float rr = sm[d0] * en[d0]; const float h = rr - half3 * gn0; zout[d0] = zn0 + e3 * (i0 * h); }
        if (a1) { const float rr = sm[d1] * en[d1]; const float h = rr - half3 * gn1; zout[d1] = zn1 + e3 * (i1 * h); }
        return true;
    }
#endif

    // one iteration of _iterative_build_subtree's loop body, after the gradient arrived
    B2_HD void on_leaf(float u, const float* g) {
#if defined(__CUDA_ARCH__)
        if (cfg.D <= 64 && !dense_on()) { on_leaf_fused(u, g); return; }
#endif
        const int Dn = cfg.D;
        const float e = c.going_right ? c.eps : -c.eps;
        const float half = 0.5f * e;
        const float* imm = v(V_IMM);
        float *zs = v(V_ZS), *rs = v(V_RS), *gs = v(V_GS);
        B2_LAPQ(-1);
        for_d_wide(Dn, [&](int d) { Vals x; x.x[0] = g[d]; x.x[1] = rs[d]; return x; },
                   [&](int d, const Vals& x) { gs[d] = x.x[0]; rs[d] = x.x[1] - half * x.x[0]; });
        lane_sync();
        c.total_leapfrogs += 1ull;
        // _build_basetree (hmc_util.py:866-875)
        const float energy_new = u + kin(rs);
        B2_LAPQ(0);
        float delta = energy_new - c.energy0;
        if (is_nan(delta)) delta = f_inf();
        const float leaf_w = -delta;
        const int leaf_div = (delta > 1000.0f) ? 1 : 0;
        const float leaf_acc = clip_max1(d_exp(-delta));
        const Key ks = mk(c.k_sub);
        float u_leaf;                                              // uniform01(split(k_sub)[1]), the proposal draw of this leaf
        if ((c.pre_mask & 1u) && key_is(c.pl_from, ks)) { st(c.k_sub, mk(c.pl_ksub)); u_leaf = c.pl_u; c.pre_hit[0] += 1u; }
        else { c.pre_miss[0] += 1u; const Key k_leaf = split_at(ks, 1); st(c.k_sub, split_at(ks, 0)); u_leaf = uniform01_at(k_leaf, 0); }
        B2_LAPQ(1);
        const int leaf_idx = c.n_sub;
        float *zps = v(V_ZPS), *gps = v(V_GPS), *rsum_s = v(V_RSUMS);
        if (leaf_idx == 0) {
            for_d_wide(Dn, [&](int d) { Vals x; x.x[0] = zs[d]; x.x[1] = gs[d]; x.x[2] = rs[d]; return x; },
                       [&](int d, const Vals& x) { zps[d] = x.x[0]; gps[d] = x.x[1]; rsum_s[d] = x.x[2]; });
            c.sub_prop_pe = u; c.sub_prop_energy = energy_new;
            c.sub_weight = leaf_w; c.sub_sum_acc = leaf_acc;
        } else {                                                   // _combine_tree, uniform kernel
            const float p = d_expit(leaf_w - c.sub_weight);
            const bool take = u_leaf < p;
            if (take) {
                for_d_wide(Dn, [&](int d) { Vals x; x.x[0] = zs[d]; x.x[1] = gs[d]; return x; },
                           [&](int d, const Vals& x) { zps[d] = x.x[0]; gps[d] = x.x[1]; });
                c.sub_prop_pe = u; c.sub_prop_energy = energy_new;
            }
            for_d_wide(Dn, [&](int d) { Vals x; x.x[0] = rsum_s[d]; x.x[1] = rs[d]; return x; },
                       [&](int d, const Vals& x) { rsum_s[d] = x.x[0] + x.x[1]; });
            c.sub_weight = d_logaddexp(c.sub_weight, leaf_w);
            c.sub_sum_acc = c.sub_sum_acc + leaf_acc;
        }
        B2_LAPQ(2);
        c.sub_div = leaf_div;
        c.n_sub = leaf_idx + 1;
        // checkpoints + iterative U-turn (hmc_util.py:941-981, 1034-1058)
        const uint32_t n = (uint32_t)leaf_idx;
        const int idx_max = popc32(n >> 1);
        const int idx_min = idx_max - popc32((~n & (n + 1u)) - 1u) + 1;
        if ((leaf_idx & 1) == 0) {
            float* cr = v(V_CKPT_R + idx_max); float* cs = v(V_CKPT_RSUM + idx_max);
            for_d_wide(Dn, [&](int d) { Vals x; x.x[0] = rs[d]; x.x[1] = rsum_s[d]; return x; },
                       [&](int d, const Vals& x) { cr[d] = x.x[0]; cs[d] = x.x[1]; });
        }
        bool sub_turning = false;
        if (dense_on() && idx_max >= idx_min) matvec(dmat(0), rs, v(V_TMP1));          // M^-1 r of the new leaf, once
        for (int i = idx_max; i >= idx_min && !sub_turning; --i) {
            const float* cr = v(V_CKPT_R + i); const float* cs = v(V_CKPT_RSUM + i);
            float l, r;
            if (dense_on()) {
                const float *vl = v(V_TMP0), *vr = v(V_TMP1);
                matvec(dmat(0), cr, v(V_TMP0));
                lane_sum2(Dn, [&](int d, float& a, float& b) {
                    const float sub = (rsum_s[d] - cs[d]) + cr[d];
                    const float s = sub - (cr[d] + rs[d]) / 2.0f;
                    a = vl[d] * s;
                    b = vr[d] * s;
                }, l, r);
            } else
            lane_sum2_wide(Dn, [&](int d, float& a, float& b) {
                const float sub = (rsum_s[d] - cs[d]) + cr[d];
                const float s = sub - (cr[d] + rs[d]) / 2.0f;
                a = (imm[d] * cr[d]) * s;
                b = (imm[d] * rs[d]) * s;
            }, l, r);
            sub_turning = (l <= 0.0f) || (r <= 0.0f);
        }
        B2_LAPQ(3);
        if (c.n_sub < (1 << c.depth) && !sub_turning && !c.sub_div) {      // next leaf of this subtree
            leap(e, zs, rs, gs, zs, rs);
            B2_LAPQ(4);
            return;
        }
        finish_doubling(sub_turning);
        B2_LAPQ(5);
    }

    // _combine_tree with the biased kernel (hmc_util.py:936-938, 767-848)
    B2_HD void finish_doubling(bool sub_turning) {
        const int Dn = cfg.D;
        const float *zs = v(V_ZS), *rs = v(V_RS), *gs = v(V_GS), *rsum_s = v(V_RSUMS);
        float *zo = v(c.going_right ? V_ZR : V_ZL), *ro = v(c.going_right ? V_RR : V_RL), *go = v(c.going_right ? V_GR : V_GL);
        float* rsum = v(V_RSUM);
        bool turning;
#if defined(__CUDA_ARCH__)
        if (Dn <= 64 && !dense_on()) {
            // two elements per lane, loads first; the tree-level U-turn test (is_turning) runs on the registers
            const int lane = (int)(threadIdx.x & 31u);
            const int d0 = lane, d1 = lane + 32;
            const bool a0 = d0 < Dn, a1 = d1 < Dn;
            const float* imm = v(V_IMM);
            const float* other = v(c.going_right ? V_RL : V_RR);        // the edge this doubling did not move
            float z0 = 0.0f, z1 = 0.0f, r0 = 0.0f, r1 = 0.0f, g0 = 0.0f, g1 = 0.0f, q0 = 0.0f, q1 = 0.0f, t0 = 0.0f, t1 = 0.0f,
                  o0 = 0.0f, o1 = 0.0f, i0 = 0.0f, i1 = 0.0f;
            if (a0) { z0 = zs[d0]; r0 = rs[d0]; g0 = gs[d0]; q0 = rsum_s[d0]; t0 = rsum[d0]; o0 = other[d0]; i0 = imm[d0]; }
            if (a1) { z1 = zs[d1]; r1 = rs[d1]; g1 = gs[d1]; q1 = rsum_s[d1]; t1 = rsum[d1]; o1 = other[d1]; i1 = imm[d1]; }
            t0 = t0 + q0; t1 = t1 + q1;
            if (a0) { zo[d0] = z0; ro[d0] = r0; go[d0] = g0; rsum[d0] = t0; }
            if (a1) { zo[d1] = z1; ro[d1] = r1; go[d1] = g1; rsum[d1] = t1; }
            turning = sub_turning;
            if (!sub_turning) {                                           // (the generic form short-circuits the same way)
                float pl = 0.0f, pr = 0.0f;
                if (a0) {
                    const float rl_ = c.going_right ? o0 : r0, rr_ = c.going_right ? r0 : o0;
                    const float sm = t0 - (rl_ + rr_) / 2.0f;
                    pl = pl + (i0 * rl_) * sm; pr = pr + (i0 * rr_) * sm;
                }
                if (a1) {
                    const float rl_ = c.going_right ? o1 : r1, rr_ = c.going_right ? r1 : o1;
                    const float sm = t1 - (rl_ + rr_) / 2.0f;
                    pl = pl + (i1 * rl_) * sm; pr = pr + (i1 * rr_) * sm;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    pl = pl + __shfl_xor_sync(0xFFFFFFFFu, pl, off);
                    pr = pr + __shfl_xor_sync(0xFFFFFFFFu, pr, off);
                }
                turning = (pl <= 0.0f) || (pr <= 0.0f);
            }
        } else
#endif
        {
            for_d_wide(Dn, [&](int d) { Vals x; x.x[0] = zs[d]; x.x[1] = rs[d]; x.x[2] = gs[d]; x.x[3] = rsum[d]; x.x[4] = rsum_s[d]; return x; },
                       [&](int d, const Vals& x) { zo[d] = x.x[0]; ro[d] = x.x[1]; go[d] = x.x[2]; rsum[d] = x.x[3] + x.x[4]; });
            lane_sync();
            turning = sub_turning || turning_between(v(V_RL), v(V_RR), rsum);
        }
        B2_LAPQ(11);
        float p = clip_max1(d_exp(c.sub_weight - c.weight));
        if (sub_turning || c.sub_div) p = 0.0f;
        const float u_fin = ((c.pre_mask & 2u) && key_is(c.pf_from, mk(c.k_fin))) ? c.pf_u : uniform01_at(mk(c.k_fin), 0);
        const bool take = u_fin < p;
        if (take) {
            copy(V_ZP, V_ZPS); copy(V_GP, V_GPS);
            c.prop_pe = c.sub_prop_pe; c.prop_energy = c.sub_prop_energy;
        }
        B2_LAPQ(12);
        c.depth += 1;
        c.weight = d_logaddexp(c.weight, c.sub_weight);
        B2_LAPQ(13);
        c.t_div = c.sub_div; c.turning = turning ? 1 : 0;
        c.sum_acc = c.sum_acc + c.sub_sum_acc;
        c.n_total += c.n_sub;
        if (c.depth < c.max_depth && !c.turning && !c.t_div) { begin_doubling(); B2_LAPQ(14); }
        else { finish_transition(); B2_LAPQ(15); }
    }

    // ---------------------------------------------------------------- plain HMC (hmc.py:364-414)
    B2_HD void hmc_begin(Key k_tr) {
        st(c.k_fin, k_tr);
        int n = cfg.hmc_num_steps;
        if (n <= 0) {
            n = (int)ceilf(cfg.traj_len / c.eps);
            c.eps = cfg.traj_len / (float)n;
        }
        if (!(n >= 1)) n = 1;         // keeps the call graph acyclic; ceil(L / eps) >= 1 for finite eps
        c.hmc_n = n; c.hmc_left = n;
        leap(c.eps, v(V_Z), v(V_R0), v(V_G), v(V_ZS), v(V_RS));
        c.phase = PH_HMC;
    }
    B2_HD void on_hmc(float u, const float* g) {
        const float half = 0.5f * c.eps;
        float *rs = v(V_RS), *gs = v(V_GS);
        B2_FOR_D(d, cfg.D) { const float gg = g[d]; gs[d] = gg; rs[d] = rs[d] - half * gg; }
        c.total_leapfrogs += 1ull;
        c.hmc_left -= 1;
        if (c.hmc_left > 0) { leap(c.eps, v(V_ZS), rs, gs, v(V_ZS), rs); return; }
        hmc_finish(u);
    }
    B2_HD void hmc_finish(float u_new) {
        const float e_old = c.energy0;
        const float e_new = u_new + kin(v(V_RS));
        float delta = e_new - e_old;
        if (is_nan(delta)) delta = f_inf();
        const float acc = clip_max1(d_exp(-delta));
        const bool take = uniform01_at(mk(c.k_fin), 0) < acc;
        if (take) { copy(V_ZP, V_ZS); copy(V_GP, V_GS); c.prop_pe = u_new; c.prop_energy = e_new; }
        else { copy(V_ZP, V_Z); copy(V_GP, V_G); c.prop_pe = c.pe; c.prop_energy = e_old; }
        c.t_div = (delta > 1000.0f) ? 1 : 0;
        c.n_total = c.hmc_n;
        finish_transition_common(acc);
    }

    // ---------------------------------------------------------------- end of sample_kernel
    B2_HD void finish_transition() {
        finish_transition_common(c.sum_acc / (float)c.n_total);     // hmc.py:441
    }
    B2_HD void finish_transition_common(float accept_prob) {
        copy(V_Z, V_ZP); copy(V_G, V_GP);
        c.pe = c.prop_pe; c.energy = c.prop_energy;
        c.num_steps = c.n_total; c.accept_prob = accept_prob; c.diverging = c.t_div;
        const int t = c.i;
        const bool in_warmup = t < cfg.num_warmup;
        const int itr = t + 1;
        const int n = in_warmup ? itr : (itr - cfg.num_warmup);
        c.mean_accept_prob = c.mean_accept_prob + (accept_prob - c.mean_accept_prob) / (float)n;   // :511-513
        c.i = itr; st(c.key, mk(c.key_next));
        c.heur_t = t;
        B2_LAPQ(8);
        if (in_warmup && adapt_update(t, accept_prob)) return;     // heuristic pending: heur_done() collects
        B2_LAPQ(9);
        collect(t);
        B2_LAPQ(10);
        begin_transition();
    }

    // fori_collect (numpyro/util.py:386-403): iteration t fills slot (t - start) / thinning; the
    // last writer of a slot wins, i.e. the iteration with (t - start + 1) % thinning == 0.
    B2_HD void collect(int t) {
        if (t < cfg.collect_start || cfg.S <= 0) return;
        const int rel = t - cfg.collect_start;
        if ((rel + 1) % cfg.thinning != 0) return;
        const int idx = rel / cfg.thinning;
        if (idx >= cfg.S) return;
        const size_t o = (size_t)chain * cfg.S + idx;
        if (out.z) { float* dst = out.z + o * cfg.D; const float* z = v(V_Z);
                     for_d_wide(cfg.D, [&](int d) { Vals x; x.x[0] = z[d]; return x; }, [&](int d, const Vals& x) { dst[d] = x.x[0]; }); }
        if (lane_first() == 0) {
            if (out.diverging) out.diverging[o] = c.diverging;
            if (out.num_steps) out.num_steps[o] = c.num_steps;
            if (out.accept_prob) out.accept_prob[o] = c.accept_prob;
            if (out.mean_accept_prob) out.mean_accept_prob[o] = c.mean_accept_prob;
            if (out.pe) out.pe[o] = c.pe;
            if (out.energy) out.energy[o] = c.energy;
            if (out.step_size) out.step_size[o] = c.step_size;
        }
        c.n_collected = idx + 1;
    }

    // ---------------------------------------------------------------- warmup_adapter.update_fn
    // hmc_util.py:637-705.  Returns true when the step-size heuristic took over (a gradient is
    // pending and begin_transition will be called from heur_done()).
    B2_HD bool adapt_update(int t, float accept_prob) {
        const Key wk = mk(c.wa_key);
        Key k_ss;
        if ((c.pre_mask & 32u) && key_is(c.pw_from, wk)) { k_ss = mk(c.pw_kss); st(c.wa_key, mk(c.pw_next)); }
        else { k_ss = split_at(wk, 1); st(c.wa_key, split_at(wk, 0)); }
        if (cfg.adapt_step) {
            // dual_averaging.update_fn (hmc_util.py:113-128), t0 = 10, kappa = 0.75, gamma = 0.05
            const float g = cfg.target_accept - accept_prob;
            const int tt = c.da_t + 1;
            const float tf = (float)(tt + 10);
            c.da_g_avg = (1.0f - 1.0f / tf) * c.da_g_avg + g / tf;
            c.da_x_t = c.da_prox - (sqrtf((float)tt) / 0.05f) * c.da_g_avg;
            const float w = d_pow((float)tt, -0.75f);
            c.da_x_avg = (1.0f - w) * c.da_x_avg + w * c.da_x_t;
            c.da_t = tt;
            const float ls = (t == cfg.num_warmup - 1) ? c.da_x_avg : c.da_x_t;
            float s = d_exp(ls);
            if (s < 1.17549435e-38f) s = 1.17549435e-38f;           // jnp.clip(step, tiny, max)
            if (s > 3.40282347e+38f) s = 3.40282347e+38f;
            c.step_size = s;
        }
        const bool middle = (c.window_idx > 0) && (c.window_idx < cfg.num_windows - 1);
        const int Dn = cfg.D;
        if (cfg.adapt_mass && middle) {                              // welford update (:178-195)
            const float* z = v(V_Z); float* mean = v(V_WF_MEAN); float* m2 = v(V_WF_M2);
            c.mm_n += 1;
            const float nf = (float)c.mm_n;
            if (dense_on()) {                                         // m2 + outer(delta_post, delta_pre)  (:193-194)
                float *pre = v(V_TMP0), *post = v(V_TMP1), *M2 = dmat(2);
                B2_FOR_D(d, Dn) {
                    const float p0 = z[d] - mean[d];
                    const float mu = mean[d] + p0 / nf;
                    pre[d] = p0; post[d] = z[d] - mu; mean[d] = mu;
                }
                lane_sync();
                B2_FOR_D(i, Dn) {
                    const float pi = post[i];
                    for (int j = 0; j < Dn; ++j) M2[(size_t)i * Dn + j] = M2[(size_t)i * Dn + j] + pi * pre[j];
                }
                lane_sync();
            } else
            for_d_wide(Dn, [&](int d) { Vals x; x.x[0] = z[d]; x.x[1] = mean[d]; x.x[2] = m2[d]; return x; },
                       [&](int d, const Vals& x) {
                const float pre = x.x[0] - x.x[1];
                const float mu = x.x[1] + pre / nf;
                const float post = x.x[0] - mu;
                mean[d] = mu; m2[d] = x.x[2] + pre * post;
            });
        }
        const int widx = (c.window_idx < cfg.num_windows) ? c.window_idx : (cfg.num_windows - 1);
        const bool at_end = (t == cfg.window_end[widx]);
        if (at_end) c.window_idx += 1;
        if (!(at_end && middle)) return false;
        if (cfg.adapt_mass) {                                        // mm_final (:197-237) + re-init
            float* imm = v(V_IMM); float* sm = v(V_SQRTM); float* mean = v(V_WF_MEAN); float* m2 = v(V_WF_M2);
            const float nf = (float)c.mm_n;
            const float nm1 = (float)(c.mm_n - 1);
            const float n5 = (float)(c.mm_n + 5);
            const float scale = nf / n5;
            const float shrink = 1e-3f * (5.0f / n5);
            if (dense_on()) {
                float *A = dmat(0), *M2 = dmat(2);
                B2_FOR_D(i, Dn) {
                    for (int j = 0; j < Dn; ++j) {
                        const size_t o = (size_t)i * Dn + j;
                        float cov = M2[o] / nm1;
                        if (cfg.regularize) cov = scale * cov + shrink * ((i == j) ? 1.0f : 0.0f);
                        A[o] = cov; M2[o] = 0.0f;
                    }
                    mean[i] = 0.0f;
                }
                dense_roots();
            } else
            B2_FOR_D(d, Dn) {
                float cov = m2[d] / nm1;
                if (cfg.regularize) cov = scale * cov + shrink;
                imm[d] = cov;
                sm[d] = 1.0f / sqrtf(cov);
                mean[d] = 0.0f; m2[d] = 0.0f;
            }
            c.mm_n = 0;
        }
        if (cfg.adapt_step) {
            if (cfg.find_heuristic) { heur_begin(k_ss, 0); return true; }
            da_reinit(d_log(10.0f) + d_log(c.step_size));
        }
        return false;
    }

    // ---------------------------------------------------------------- PRNG look-ahead
    // Everything the next advance() may draw from the PRNG depends only on keys that are already known while
    // the gradient is still being computed.  prefetch() derives those values ahead of time (the streaming
    // engine calls it while the grid sweeps X); advance() picks them up through the key tags.  Results are
    // identical with or without it (tests/hostsim never calls it, the GPU parity tests always do).
    template <int e> B2_HD void prefetch_doubling(Key k) {
        if ((c.pre_mask & (4u << e)) && key_is(c.pd_from[e], k)) return;
        const Key k_dir = split_at(k, 1), k_dbl = split_at(k, 2);
        st(c.pd_kloop[e], split_at(k, 0));
        c.pd_right[e] = (uniform01_at(k_dir, 0) < 0.5f) ? 1 : 0;
        st(c.pd_ksub[e], split_at(k_dbl, 0)); st(c.pd_kfin[e], split_at(k_dbl, 1));
        st(c.pd_from[e], k); c.pre_mask |= (4u << e);
    }
    B2_HD void prefetch() {
        if (c.phase != PH_LEAF) return;
        { const Key ks = mk(c.k_sub);                               // the leaf in flight
          if (!((c.pre_mask & 1u) && key_is(c.pl_from, ks))) {
              c.pl_u = uniform01_at(split_at(ks, 1), 0); st(c.pl_ksub, split_at(ks, 0)); st(c.pl_from, ks); c.pre_mask |= 1u;
          } }
        { const Key kf = mk(c.k_fin);                               // end of this doubling
          if (!((c.pre_mask & 2u) && key_is(c.pf_from, kf))) { c.pf_u = uniform01_at(kf, 0); st(c.pf_from, kf); c.pre_mask |= 2u; } }
        prefetch_doubling<0>(mk(c.k_loop));                         // the next doubling of this tree
        { const Key key = mk(c.key_next);                           // the next transition
          if (!((c.pre_mask & 16u) && key_is(c.pt_from, key))) {
              const Key k_mom = split_at(key, 1), k_tr = split_at(key, 2);
              st(c.pt_keynext, split_at(key, 0)); st(c.pt_ktr, k_tr);
              const Key km = cfg.model_built ? split_at(k_mom, 0) : k_mom;
              float* en = v(V_EPS);
              B2_FOR_D(d, cfg.D) en[d] = normal_at(km, (uint32_t)d);
              st(c.pt_from, key); c.pre_mask |= 16u;
          }
          prefetch_doubling<1>(mk(c.pt_ktr)); }
        { const Key wk = mk(c.wa_key);                              // warm-up adaptation key
          if (!((c.pre_mask & 32u) && key_is(c.pw_from, wk))) {
              st(c.pw_kss, split_at(wk, 1)); st(c.pw_next, split_at(wk, 0)); st(c.pw_from, wk); c.pre_mask |= 32u;
          } }
    }

    // ---------------------------------------------------------------- dispatcher
    // Feed the potential/gradient evaluated at V_ZS; on return either phase == PH_DONE or V_ZS
    // holds the next position to evaluate.
    B2_HD void advance(float u, const float* g) {
        switch (c.phase) {
        case PH_INIT: {
            const bool bad = !is_finite(u) || lane_any(cfg.D, [&](int d) { return !is_finite(g[d]); });
            c.init_tries += 1;
            if (bad && !cfg.init_given && c.init_tries < 100) { init_draw(); return; }
            // "Cannot find valid initial parameters" (infer/util.py:800-832): the chain stops here; b200nuts_sync /
            // b200nuts_get_state report B200NUTS_EINIT
            if (bad) { c.init_failed = 1; c.phase = PH_DONE; return; }
            finish_init(u, g);
            return;
        }
        case PH_LEAF: on_leaf(u, g); return;
        case PH_HMC: on_hmc(u, g); return;
        case PH_HEUR: on_heur(u, g); return;
        default: return;
        }
    }
};
using Tick = TickT<true>;

}  // namespace b2
