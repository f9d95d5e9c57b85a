// Host-side glue shared by the C ABI (b200nuts.cu) and the test-only host simulator
// (tests/hostsim/hostsim.cpp): config validation, latent layout, adaptation schedule.
#pragma once
#include <string>
#include <vector>
#include <utility>
#include "../../include/b200nuts.h"
#include "tick.cuh"
#include "families.cuh"

namespace b2 {

// build_adaptation_schedule (numpyro/infer/hmc_util.py:387-436): inclusive window ends.
inline std::vector<std::pair<int, int>> build_adaptation_schedule(int num_steps) {
    std::vector<std::pair<int, int>> s;
    if (num_steps < 20) { s.push_back({0, num_steps - 1}); return s; }
    int start_buf = 75, end_buf = 50, init_win = 25;
    if (start_buf + end_buf + init_win > num_steps) {
        start_buf = (int)(0.15 * num_steps);
        end_buf = (int)(0.1 * num_steps);
        init_win = num_steps - start_buf - end_buf;
    }
    s.push_back({0, start_buf - 1});
    const int end_start = num_steps - end_buf;
    int next_size = init_win, next_start = start_buf;
    while (next_start < end_start) {
        const int cur_start = next_start;
        int cur_size = next_size;
        if (3 * cur_size <= end_start - cur_start) next_size = 2 * cur_size;
        else cur_size = end_start - cur_start;
        next_start = cur_start + cur_size;
        s.push_back({cur_start, next_start - 1});
    }
    s.push_back({end_start, num_steps - 1});
    return s;
}

struct SiteLayout { int n_sites; int off[kMaxSites]; int size[kMaxSites]; };   // trace order

// Latent layout: flat order = site names sorted (hmc.py:765-768); init order = model trace order
// (infer/util.py:454-463).  Returns an error string or "".
inline std::string make_family(const B200NutsConfig& c, FamilySpec& f, SiteLayout& sites) {
    memset(&f, 0, sizeof(f));
    memset(&sites, 0, sizeof(sites));
    f.family = c.family;
    f.X = c.X; f.y = c.y; f.N = c.n_rows; f.Dx = c.n_cols;
    f.off_lambda = f.off_tau = f.off_prec = -1; f.off_u = 0;
    f.likelihood = c.likelihood; f.gscale = c.global_scale;
    f.tau_scale = c.tau_scale > 0 ? c.tau_scale : 1.0f;
    f.mu_scale = c.mu_scale > 0 ? c.mu_scale : 5.0f;
    if (c.num_chains <= 0) return "num_chains must be positive";
    switch (c.family) {
    case B200NUTS_FAMILY_DIAG_GAUSSIAN:
        if (c.n_rows <= 0 || !c.aux) return "diag_gaussian needs n_rows = D > 0 and aux = [mu, sigma]";
        f.D = (int)c.n_rows; f.aux0 = c.aux; f.aux1 = c.aux + f.D;
        sites.n_sites = 1; sites.off[0] = 0; sites.size[0] = f.D;
        return "";
    case B200NUTS_FAMILY_EIGHT_SCHOOLS:
        if (c.n_rows <= 0 || !c.aux || !c.y) return "eight_schools needs n_rows = J, y and aux = sigma";
        f.D = (int)c.n_rows + 2; f.aux0 = c.aux; f.aux1 = c.y;
        if (!(c.tau_scale > 0)) f.tau_scale = 5.0f;
        // trace order mu, tau, theta_base == sorted order
        sites.n_sites = 3; sites.off[0] = 0; sites.size[0] = 1; sites.off[1] = 1; sites.size[1] = 1;
        sites.off[2] = 2; sites.size[2] = (int)c.n_rows;
        return "";
    case B200NUTS_FAMILY_GLM: {
        if (c.n_rows <= 0 || c.n_cols <= 0 || !c.X || !c.y) return "glm needs X [n_rows, n_cols] and y";
        if (c.likelihood < 0 || c.likelihood > 2) return "unknown likelihood";
        if (c.global_scale < 0 || c.global_scale > 2) return "unknown global_scale prior";
        const int Dx = c.n_cols;
        f.g0 = c.group_col_begin; f.g1 = c.group_col_end;
        if (f.g0 == 0 && f.g1 == 0) f.g1 = Dx;
        if (f.g0 < 0 || f.g1 > Dx || f.g0 > f.g1) return "bad group column range";
        const bool local = c.local_scales != 0, glob = c.global_scale != 0, normal = c.likelihood == 2;
        int n = 0;
        if (local) {
            // horseshoe names, sorted: lambdas, (prec_obs), tau, unscaled_betas
            int off = 0;
            f.off_lambda = off; off += Dx;
            if (normal) { f.off_prec = off; off += 1; }
            if (glob) { f.off_tau = off; off += 1; }
            f.off_u = off; off += Dx;
            f.D = off;
            // trace order: lambdas, tau, unscaled_betas, prec_obs (horseshoe_regression.py:37-58)
            sites.off[n] = f.off_lambda; sites.size[n++] = Dx;
            if (glob) { sites.off[n] = f.off_tau; sites.size[n++] = 1; }
            sites.off[n] = f.off_u; sites.size[n++] = Dx;
            if (normal) { sites.off[n] = f.off_prec; sites.size[n++] = 1; }
        } else {
            // names, sorted: coefs, (prec_obs), (tau); trace order: (tau), coefs, (prec_obs)
            int off = 0;
            f.off_u = off; off += Dx;
            if (normal) { f.off_prec = off; off += 1; }
            if (glob) { f.off_tau = off; off += 1; }
            f.D = off;
            if (glob) { sites.off[n] = f.off_tau; sites.size[n++] = 1; }
            sites.off[n] = f.off_u; sites.size[n++] = Dx;
            if (normal) { sites.off[n] = f.off_prec; sites.size[n++] = 1; }
        }
        sites.n_sites = n;
        if (c.ecs_subsample_size > 0) {
            if (local || glob || normal) return "energy-conserving subsampling is implemented for the plain GLM (coefs only, Bernoulli / Poisson)";
            if (c.ecs_subsample_size > c.n_rows) return "ecs_subsample_size > n_rows";
            if (c.ecs_proxy_degree < 0 || c.ecs_proxy_degree > 2) return "ecs_proxy_degree must be 0, 1 or 2";
            f.ecs_m = c.ecs_subsample_size; f.ecs_degree = c.ecs_proxy_degree;
        }
        return "";
    }
    default: return "unknown family";
    }
}

inline std::string make_tick_cfg(const B200NutsConfig& c, const FamilySpec& f, const SiteLayout& sites,
                                 int num_warmup, bool init_given, TickCfg& t) {
    memset(&t, 0, sizeof(t));
    t.D = f.cond_Dfree > 0 ? f.cond_Dfree : f.D;
    t.num_warmup = num_warmup; t.total_iters = 0;
    t.md_warm = c.max_tree_depth_warmup > 0 ? c.max_tree_depth_warmup : 10;
    t.md_post = c.max_tree_depth > 0 ? c.max_tree_depth : 10;
    if (t.md_warm > kMaxDepthAlloc || t.md_post > kMaxDepthAlloc) return "max_tree_depth > 12 is not supported";
    if (t.md_warm < 1 || t.md_post < 1) return "max_tree_depth must be >= 1";
    t.target_accept = c.target_accept_prob > 0 ? c.target_accept_prob : 0.8f;
    t.init_step_size = c.step_size > 0 ? c.step_size : 1.0f;
    t.adapt_step = c.adapt_step_size; t.adapt_mass = c.adapt_mass_matrix; t.regularize = c.regularize_mass_matrix;
    t.model_built = c.model_built; t.find_heuristic = c.find_heuristic_step_size;
    t.algo = c.algo; t.hmc_num_steps = c.hmc_num_steps;
    t.traj_len = c.trajectory_length > 0 ? c.trajectory_length : 6.283185307179586f;
    auto sched = build_adaptation_schedule(num_warmup);
    if (sched.size() > 16) return "too many adaptation windows";
    t.num_windows = (int)sched.size();
    for (size_t i = 0; i < sched.size(); ++i) t.window_end[i] = sched[i].second;
    t.collect_start = 0; t.thinning = 1; t.S = 0;
    t.init_given = init_given ? 1 : 0;
    t.init_radius = c.init_radius > 0 ? c.init_radius : 2.0f;
    t.dense = c.dense_mass ? 1 : 0; t.imm_given = 0;        // (imm_given: set by the caller that pre-loaded the matrix)
    t.n_sites = sites.n_sites;
    for (int i = 0; i < sites.n_sites; ++i) { t.site_off[i] = sites.off[i]; t.site_size[i] = sites.size[i]; }
    return "";
}

inline void ctl_to_public(const ChainCtl& c, B200NutsChainState& s) {
    s.i = c.i; s.rng_key[0] = c.key[0]; s.rng_key[1] = c.key[1];
    s.potential_energy = c.pe; s.energy = c.energy; s.num_steps = c.num_steps;
    s.accept_prob = c.accept_prob; s.mean_accept_prob = c.mean_accept_prob; s.diverging = c.diverging;
    s.step_size = c.step_size;
    s.ss_x_t = c.da_x_t; s.ss_x_avg = c.da_x_avg; s.ss_g_avg = c.da_g_avg; s.ss_prox = c.da_prox; s.ss_t = c.da_t;
    s.mm_n = c.mm_n; s.window_idx = c.window_idx;
    s.adapt_rng_key[0] = c.wa_key[0]; s.adapt_rng_key[1] = c.wa_key[1];
    s.init_failed = c.init_failed; s.done = (c.phase == PH_DONE) ? 1 : 0;
    s.total_leapfrogs = c.total_leapfrogs;
}

inline void public_to_ctl(const B200NutsChainState& s, ChainCtl& c) {
    memset(&c, 0, sizeof(c));
    c.i = s.i; c.key[0] = s.rng_key[0]; c.key[1] = s.rng_key[1];
    c.pe = s.potential_energy; c.energy = s.energy; c.num_steps = s.num_steps;
    c.accept_prob = s.accept_prob; c.mean_accept_prob = s.mean_accept_prob; c.diverging = s.diverging;
    c.step_size = s.step_size;
    c.da_x_t = s.ss_x_t; c.da_x_avg = s.ss_x_avg; c.da_g_avg = s.ss_g_avg; c.da_prox = s.ss_prox; c.da_t = s.ss_t;
    c.mm_n = s.mm_n; c.window_idx = s.window_idx;
    c.wa_key[0] = s.adapt_rng_key[0]; c.wa_key[1] = s.adapt_rng_key[1];
    c.init_failed = s.init_failed; c.phase = PH_DONE;
    c.total_leapfrogs = s.total_leapfrogs;
}

}  // namespace b2
