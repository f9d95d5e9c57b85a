// TEST-ONLY host simulator of the per-chain state machine.
//
// g++ compiles the single-source numpyro_b200/csrc/tick.cuh (and families.cuh) for the CPU so that
// the `-m "not gpu"` tests can compare the engine's bookkeeping with the NumPy oracle bit for bit
// without a GPU.  This library is never loaded by the numpyro_b200 package: the product path is the
// CUDA engine only (no CPU fallback).
#include <vector>
#include <string>
#include <cstdio>
#include "../../numpyro_b200/csrc/engine.cuh"

using namespace b2;

typedef void (*potential_cb)(void* user, int chain, const float* z, int D, float* U, float* g);

struct HostSim {
    B200NutsConfig cfg; FamilySpec fam; SiteLayout sites; TickCfg tick;
    int C, D;
    std::vector<ChainCtl> ctl; std::vector<float> vecs, gtmp, scratch, ylgam, dense;
    std::string err; bool inited = false; bool lookahead = false; bool imm_given = false;
    ChainVecs cv(int chain) {
        ChainVecs v; v.base = vecs.data() + (size_t)chain * D; v.field_stride = C * D;
        v.dense = dense.empty() ? nullptr : dense.data() + (size_t)chain * 4 * D * D;
        return v;
    }
};

extern "C" {

int hostsim_create(const B200NutsConfig* cfg, HostSim** out) {
    HostSim* h = new HostSim();
    h->cfg = *cfg;
    std::string e = make_family(*cfg, h->fam, h->sites);
    if (!e.empty()) { fprintf(stderr, "hostsim: %s\n", e.c_str()); delete h; return B200NUTS_EINVAL; }
    h->C = cfg->num_chains; h->D = h->fam.D;
    h->ctl.assign(h->C, ChainCtl());
    h->vecs.assign((size_t)V_COUNT * h->C * h->D, 0.0f);
    h->gtmp.assign((size_t)h->C * h->D, 0.0f);
    if (cfg->dense_mass) h->dense.assign((size_t)h->C * 4 * h->D * h->D, 0.0f);
    if (h->fam.family == FAM_GLM) {
        h->scratch.assign((size_t)(h->fam.N + h->fam.Dx), 0.0f);
        if (h->fam.likelihood == LIK_POISSON) {
            double acc = 0.0;
            for (long long n = 0; n < h->fam.N; ++n) acc += lgamma((double)h->fam.y[n] + 1.0);
            h->fam.nll_const = (float)acc;
        }
    }
    *out = h;
    return 0;
}

void hostsim_destroy(HostSim* h) { delete h; }
// run Tick::prefetch() before every gradient, like the streaming engine does while it sweeps X
void hostsim_set_lookahead(HostSim* h, int on) { h->lookahead = on != 0; }
int hostsim_dim(HostSim* h) { return h->D; }

static void eval(HostSim* h, int chain, potential_cb cb, void* user, float& u) {
    float* z = h->cv(chain).v(V_ZS);
    float* g = h->gtmp.data() + (size_t)chain * h->D;
    if (cb) cb(user, chain, z, h->D, &u, g);
    else potential_inwarp(h->fam, z, h->scratch.data(), u, g);
}

int hostsim_init(HostSim* h, const uint32_t* keys, const float* z0, int num_warmup) {
    std::string e = make_tick_cfg(h->cfg, h->fam, h->sites, num_warmup, z0 != nullptr, h->tick);
    if (!e.empty()) { fprintf(stderr, "hostsim: %s\n", e.c_str()); return B200NUTS_EINVAL; }
    h->tick.imm_given = h->imm_given ? 1 : 0;
    OutBufs none; memset(&none, 0, sizeof(none));
    for (int c = 0; c < h->C; ++c) {
        memset(&h->ctl[c], 0, sizeof(ChainCtl));
        Tick t{h->tick, h->ctl[c], h->cv(c), none, c, h->C};
        Key k; k.a = keys[2 * c]; k.b = keys[2 * c + 1];
        t.begin(k, z0 ? z0 + (size_t)c * h->D : nullptr);
    }
    h->inited = true;
    return 0;
}

int hostsim_run(HostSim* h, const B200NutsRun* run, potential_cb cb, void* user) {
    if (!h->inited) return B200NUTS_ESTATE;
    h->tick.total_iters = run->upper;
    h->tick.collect_start = run->collect_start; h->tick.thinning = run->thinning > 0 ? run->thinning : 1;
    h->tick.S = run->collection_size;
    OutBufs out; out.z = run->z; out.diverging = run->diverging; out.num_steps = run->num_steps;
    out.accept_prob = run->accept_prob; out.mean_accept_prob = run->mean_accept_prob; out.pe = run->potential_energy;
    out.energy = run->energy; out.step_size = run->step_size;
    for (int c = 0; c < h->C; ++c) {
        Tick t{h->tick, h->ctl[c], h->cv(c), out, c, h->C};
        if (h->ctl[c].phase == PH_DONE && !h->ctl[c].init_failed) t.begin_transition();
        while (h->ctl[c].phase != PH_DONE) {
            float u;
            if (h->lookahead) t.prefetch();
            eval(h, c, cb, user, u);
            t.advance(u, h->gtmp.data() + (size_t)c * h->D);
        }
    }
    return 0;
}

int hostsim_get_state(HostSim* h, B200NutsChainState* st, float* z, float* g, float* imm, float* sqrtm) {
    for (int c = 0; c < h->C; ++c) {
        ctl_to_public(h->ctl[c], st[c]);
        ChainVecs v = h->cv(c);
        for (int d = 0; d < h->D; ++d) {
            if (z) z[(size_t)c * h->D + d] = v.v(V_Z)[d];
            if (g) g[(size_t)c * h->D + d] = v.v(V_G)[d];
            if (imm) imm[(size_t)c * h->D + d] = v.v(V_IMM)[d];
            if (sqrtm) sqrtm[(size_t)c * h->D + d] = v.v(V_SQRTM)[d];
        }
    }
    return 0;
}

// the kernel's inverse_mass_matrix= argument: [D] or [D][D], the same for every chain (call before hostsim_init)
int hostsim_set_inverse_mass_matrix(HostSim* h, const float* imm, int ndim) {
    const int D = h->D;
    for (int c = 0; c < h->C; ++c) {
        ChainVecs v = h->cv(c);
        if (h->cfg.dense_mass) {
            float* A = v.dense;
            for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j)
                A[(size_t)i * D + j] = ndim == 2 ? imm[(size_t)i * D + j] : (i == j ? imm[i] : 0.0f);
        } else {
            for (int d = 0; d < D; ++d) v.v(V_IMM)[d] = ndim == 2 ? imm[(size_t)d * D + d] : imm[d];
        }
    }
    h->imm_given = true;
    return 0;
}
// dense handles: [C][D][D] each (NULL = skip)
int hostsim_get_dense_state(HostSim* h, float* imm, float* sqrtm, float* sqrt_inv, float* m2) {
    if (h->dense.empty()) return B200NUTS_ESTATE;
    const int D = h->D; const size_t DD = (size_t)D * D;
    for (int c = 0; c < h->C; ++c) {
        const float* blk = h->dense.data() + (size_t)c * 4 * DD;
        for (size_t o = 0; o < DD; ++o) {
            if (imm) imm[c * DD + o] = blk[o];
            if (sqrtm) sqrtm[c * DD + o] = blk[DD + o];
            if (m2) m2[c * DD + o] = blk[2 * DD + o];
        }
        if (sqrt_inv) for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j)
            sqrt_inv[c * DD + (size_t)i * D + j] = blk[3 * DD + (size_t)(D - 1 - j) * D + (D - 1 - i)];     // tril_inv[i][j] = Lc[D-1-j][D-1-i]
    }
    return 0;
}

int hostsim_potential(HostSim* h, const float* z, float* U, float* g) {
    for (int c = 0; c < h->C; ++c) {
        float u;
        potential_inwarp(h->fam, z + (size_t)c * h->D, h->scratch.data(), u, g + (size_t)c * h->D);
        U[c] = u;
    }
    return 0;
}

// PRNG / det-math parity on the host build of the device headers
void hostsim_prng_split(const uint32_t* key, int num, uint32_t* out) {
    Key k; k.a = key[0]; k.b = key[1];
    for (int i = 0; i < num; ++i) { Key o = split_at(k, (uint32_t)i); out[2 * i] = o.a; out[2 * i + 1] = o.b; }
}
void hostsim_prng_uniform(const uint32_t* key, long long n, float lo, float hi, float* out) {
    Key k; k.a = key[0]; k.b = key[1];
    for (long long i = 0; i < n; ++i) out[i] = uniform_at(k, (uint32_t)i, lo, hi);
}
void hostsim_prng_normal(const uint32_t* key, long long n, float* out) {
    Key k; k.a = key[0]; k.b = key[1];
    for (long long i = 0; i < n; ++i) out[i] = normal_at(k, (uint32_t)i);
}
void hostsim_detmath(int op, const float* x, long long n, float* out) {
    for (long long i = 0; i < n; ++i) {
        switch (op) {
        case 0: out[i] = d_exp(x[i]); break;
        case 1: out[i] = d_log(x[i]); break;
        case 2: out[i] = d_log1p(x[i]); break;
        case 3: out[i] = d_expit(x[i]); break;
        default: out[i] = d_erfinv(x[i]); break;
        }
    }
}

}  // extern "C"
