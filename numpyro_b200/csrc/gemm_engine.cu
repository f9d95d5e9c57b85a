// Host side of the many-chain GEMM regime: tile images, the per-run CUDA graph (device-side WHILE loop), parity hook.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "gemm_engine.h"
#include "gemm_engine.cuh"

namespace b2 {

struct GemmRegime {
    FamilySpec fam;
    int C = 0, Dp = 0, CT = 0, num_sms = 0, grid = 0;
    GemmParams gp;
    float *bimg = nullptr, *ximg = nullptr, *xtimg = nullptr, *yimg = nullptr, *partial = nullptr, *pnll = nullptr;
    float *gtmp = nullptr, *gbeta = nullptr, *nll = nullptr;
    int *tile_count = nullptr, *active_tiles = nullptr;
    GemmSched* sched = nullptr; GemmCtx* ctx = nullptr;
    ChainCtl* ctl = nullptr; float* vecs = nullptr; float* dense = nullptr;
    bool use_graph = true;
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; cudaGraphConditionalHandle cond = 0;
    unsigned long long passes_seen = 0ull;
    size_t smem = 0;
    GemmShardDev shard;                        // count == 1: not row-sharded
    unsigned int epoch = 0;                    // launches of a row-sharded handle so far (every rank makes the same sequence)
};

#define GCK(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return std::string(#call) + ": " + cudaGetErrorString(e_);      \
    } while (0)

static const void* gemm_kernel_for(int lik) {
    return lik == LIK_BERNOULLI ? (const void*)gemm_pass_kernel<LIK_BERNOULLI>
         : lik == LIK_POISSON   ? (const void*)gemm_pass_kernel<LIK_POISSON>
                                : (const void*)gemm_pass_kernel<LIK_NORMAL>;
}

static std::string launch_pass(GemmRegime* g, cudaStream_t st) {
    void* args[] = {&g->gp};
    GCK(cudaLaunchKernel(gemm_kernel_for(g->fam.likelihood), dim3(g->grid), dim3(kGtThreads), args, g->smem, st));
    return "";
}

static int reduce_blocks(const GemmRegime* g) {
    const long long words = (long long)g->C * (g->gp.Dxp >> 2);
    const long long b = (words + 255) / 256;
    return (int)(b < 8ll * g->num_sms ? (b > 0 ? b : 1) : 8ll * g->num_sms);
}

static std::string launch_reduce(GemmRegime* g, cudaStream_t st) {
    k_gemm_reduce<<<reduce_blocks(g), 256, 0, st>>>(g->gp, g->gbeta, g->nll, g->C);
    GCK(cudaGetLastError());
    return "";
}

static std::string launch_tick(GemmRegime* g, int first, cudaStream_t st) {
    const int blocks = (g->C + 3) / 4;
    k_gemm_tick<<<blocks, 128, 0, st>>>(g->gp, g->ctx, g->sched, g->fam, g->ctl, g->vecs, g->gtmp, g->gbeta, g->nll, g->bimg, g->tile_count,
                                        g->C, g->Dp, first, g->shard, g->dense);
    GCK(cudaGetLastError());
    k_gemm_sched<<<1, 32, 0, st>>>(g->ctx, g->sched, g->tile_count, g->active_tiles, g->CT, first, g->cond, 0);
    GCK(cudaGetLastError());
    return "";
}

// graph = WHILE (cond) { gemm pass; reduce; tick; schedule (sets cond) }
static std::string build_graph(GemmRegime* g) {
    GCK(cudaGraphCreate(&g->graph, 0));
    GCK(cudaGraphConditionalHandleCreate(&g->cond, g->graph, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    np.conditional.handle = g->cond; np.conditional.type = cudaGraphCondTypeWhile; np.conditional.size = 1;
    cudaGraphNode_t wnode;
    GCK(cudaGraphAddNode(&wnode, g->graph, nullptr, 0, &np));
    cudaGraph_t body = np.conditional.phGraph_out[0];
    cudaGraphNode_t n_pass, n_red, n_tick, n_sched;
    {
        void* args[] = {&g->gp};
        cudaKernelNodeParams kp; memset(&kp, 0, sizeof(kp));
        kp.func = (void*)gemm_kernel_for(g->fam.likelihood); kp.gridDim = dim3(g->grid); kp.blockDim = dim3(kGtThreads);
        kp.sharedMemBytes = (unsigned int)g->smem; kp.kernelParams = args;
        GCK(cudaGraphAddKernelNode(&n_pass, body, nullptr, 0, &kp));
    }
    {
        void* args[] = {&g->gp, &g->gbeta, &g->nll, &g->C};
        cudaKernelNodeParams kp; memset(&kp, 0, sizeof(kp));
        kp.func = (void*)k_gemm_reduce; kp.gridDim = dim3(reduce_blocks(g)); kp.blockDim = dim3(256); kp.kernelParams = args;
        GCK(cudaGraphAddKernelNode(&n_red, body, &n_pass, 1, &kp));
    }
    {
        int first = 0;
        void* args[] = {&g->gp, &g->ctx, &g->sched, &g->fam, &g->ctl, &g->vecs, &g->gtmp, &g->gbeta, &g->nll, &g->bimg, &g->tile_count, &g->C, &g->Dp, &first, &g->shard, &g->dense};
        cudaKernelNodeParams kp; memset(&kp, 0, sizeof(kp));
        kp.func = (void*)k_gemm_tick; kp.gridDim = dim3((g->C + 3) / 4); kp.blockDim = dim3(128); kp.kernelParams = args;
        GCK(cudaGraphAddKernelNode(&n_tick, body, &n_red, 1, &kp));
    }
    {
        int first = 0, use_cond = 1;
        void* args[] = {&g->ctx, &g->sched, &g->tile_count, &g->active_tiles, &g->CT, &first, &g->cond, &use_cond};
        cudaKernelNodeParams kp; memset(&kp, 0, sizeof(kp));
        kp.func = (void*)k_gemm_sched; kp.gridDim = dim3(1); kp.blockDim = dim3(32); kp.kernelParams = args;
        GCK(cudaGraphAddKernelNode(&n_sched, body, &n_tick, 1, &kp));
    }
    GCK(cudaGraphInstantiate(&g->exec, g->graph, 0));
    return "";
}

std::string gemm_create(GemmRegime** out, const FamilySpec& fam, int C, int Dp, int num_sms, ChainCtl* ctl, float* vecs, float* dense,
                        long long* launches) {
    *out = nullptr;
    GemmRegime* g = new GemmRegime();
    g->fam = fam; g->C = C; g->Dp = Dp; g->num_sms = num_sms; g->ctl = ctl; g->vecs = vecs; g->dense = dense;
    g->grid = num_sms;
    if (const char* e = getenv("B200NUTS_GRID")) { const int v = atoi(e); if (v >= 1 && v < g->grid) g->grid = v; }
    g->use_graph = !getenv("B200NUTS_GEMM_HOSTLOOP");
    const int CT = (C + kGtChains - 1) / kGtChains;
    const long long RC = (fam.N + kGtRows - 1) / kGtRows;
    const int KB = (fam.Dx + 31) / 32, Dxp = KB * 32, NDB = (Dxp + kGtNB - 1) / kGtNB;
    if (RC > 0x7FFFFFFFll / 8) { delete g; return "gemm regime: too many rows for one handle"; }
    // units (chain tile x segment) per CTA: ~16 when there are many chain tiles; with a handful of tiles fewer, longer units keep
    // the number of partial sums a tick has to add per chain small (config 5: one tile, 2 units per CTA)
    long long S = (16ll * num_sms + CT - 1) / CT;
    if (S > 2ll * num_sms) S = 2ll * num_sms;
    if (const char* e = getenv("B200NUTS_GEMM_SEGMENTS")) S = atoll(e);
    if (S < 1) S = 1;
    if (S > RC) S = RC;
    const long long cps = (RC + S - 1) / S;
    S = (RC + cps - 1) / cps;
    g->CT = CT;
    GemmParams& gp = g->gp; memset(&gp, 0, sizeof(gp));
    gp.CT = CT; gp.RC = (int)RC; gp.KB = KB; gp.S = (int)S; gp.cps = (int)cps; gp.NDB = NDB; gp.Dxp = Dxp; gp.N = fam.N;
    gp.spin_limit = 4000000000ll;
    if (const char* e = getenv("B200NUTS_SPIN_LIMIT")) gp.spin_limit = atoll(e);
    auto fail = [&](const std::string& m) { gemm_destroy(g); return m; };
#define GALLOC(ptr, bytes) do { cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes)); if (e_ != cudaSuccess) return fail(std::string("cudaMalloc " #ptr ": ") + cudaGetErrorString(e_)); } while (0)
    const size_t img_x = (size_t)RC * KB * 2 * kGtTileFloats * 4;                 // = 8 bytes per element of the padded X
    GALLOC(g->bimg, (size_t)CT * KB * 2 * kGtTileFloats * 4);
    GALLOC(g->ximg, img_x);
    GALLOC(g->xtimg, (size_t)RC * kGtRows * Dxp * 2 * 4);
    GALLOC(g->yimg, (size_t)RC * kGtRows * 4);
    GALLOC(g->partial, (size_t)CT * S * kGtChains * Dxp * 4);
    GALLOC(g->pnll, (size_t)CT * S * 4 * kGtChains * 4);
    GALLOC(g->gtmp, (size_t)C * Dp * 4);
    GALLOC(g->gbeta, (size_t)C * Dxp * 4);
    GALLOC(g->nll, (size_t)C * 4);
    GALLOC(g->tile_count, (size_t)2 * CT * 4);
    GALLOC(g->active_tiles, (size_t)CT * 4);
    GALLOC(g->sched, sizeof(GemmSched));
    GALLOC(g->ctx, sizeof(GemmCtx));
#undef GALLOC
    cudaMemset(g->bimg, 0, (size_t)CT * KB * 2 * kGtTileFloats * 4);
    cudaMemset(g->sched, 0, sizeof(GemmSched));
    cudaMemset(g->ctx, 0, sizeof(GemmCtx));
    cudaMemset(g->tile_count, 0, (size_t)2 * CT * 4);
    cudaMemset(g->partial, 0, (size_t)CT * S * kGtChains * Dxp * 4);
    cudaMemset(g->pnll, 0, (size_t)CT * S * 4 * kGtChains * 4);
    gp.bimg = g->bimg; gp.ximg = g->ximg; gp.xtimg = g->xtimg; gp.yimg = g->yimg; gp.partial = g->partial; gp.pnll = g->pnll;
    gp.active_tiles = g->active_tiles; gp.n_active = &g->sched->n_active;
    gp.abort_flag = &g->sched->abort_flag; gp.dbg = g->sched->dbg;
    k_gemm_pack_x<<<num_sms * 8, 256>>>(fam.X, fam.N, fam.Dx, KB, RC, g->ximg);
    k_gemm_pack_xt<<<num_sms * 8, 256>>>(fam.X, fam.N, fam.Dx, Dxp, RC, g->xtimg);
    k_gemm_pack_y<<<num_sms * 2, 256>>>(fam.y, fam.N, RC, g->yimg);
    *launches += 3;
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return fail(std::string("gemm regime: tile images: ") + cudaGetErrorString(ce));
    g->smem = gemm_smem_bytes();
    const void* fn = gemm_kernel_for(fam.likelihood);
    if ((ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem)) != cudaSuccess)
        return fail(std::string("cudaFuncSetAttribute(gemm_pass_kernel): ") + cudaGetErrorString(ce));
    {   // load every kernel now (lazy module loading must not happen inside a captured / conditional launch)
        cudaFuncAttributes fa;
        const void* fns[] = {fn, (const void*)k_gemm_tick, (const void*)k_gemm_sched, (const void*)k_gemm_hook_begin, (const void*)k_gemm_hook_finish,
                             (const void*)k_gemm_reduce};
        for (const void* f : fns)
            if ((ce = cudaFuncGetAttributes(&fa, f)) != cudaSuccess) return fail(std::string("cudaFuncGetAttributes: ") + cudaGetErrorString(ce));
    }
    memset(&g->shard, 0, sizeof(g->shard));
    g->shard.count = 1;
    if (g->use_graph) {                        // (gemm_set_shards rebuilds it: the graph bakes the kernel arguments)
        std::string e = build_graph(g);
        if (!e.empty()) return fail("gemm regime: " + e);
    }
    if ((ce = cudaDeviceSynchronize()) != cudaSuccess) return fail(std::string("gemm regime create: ") + cudaGetErrorString(ce));
    *out = g;
    return "";
}

void gemm_destroy(GemmRegime* g) {
    if (!g) return;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    cudaFree(g->bimg); cudaFree(g->ximg); cudaFree(g->xtimg); cudaFree(g->yimg); cudaFree(g->partial); cudaFree(g->pnll);
    cudaFree(g->gtmp); cudaFree(g->gbeta); cudaFree(g->nll); cudaFree(g->tile_count); cudaFree(g->active_tiles); cudaFree(g->sched); cudaFree(g->ctx);
    delete g;
}

void gemm_describe(const GemmRegime* g, int* info8) {
    info8[0] = g->gp.CT; info8[1] = g->gp.RC; info8[2] = g->gp.KB; info8[3] = g->gp.S; info8[4] = g->gp.cps; info8[5] = g->gp.NDB;
    info8[6] = g->gp.Dxp; info8[7] = g->use_graph ? 1 : 0;
}

size_t gemm_mail_bytes(const GemmRegime* g) { return gemm_mail_words(g->C, g->gp.Dxp) * sizeof(float2); }

std::string gemm_set_shards(GemmRegime* g, int rank, int count, void* const* mail, long long n_rows_global, float nll_local_const) {
    if (count < 1 || count > kGemmMaxShards || rank < 0 || rank >= count) return "gemm regime: shard rank / count out of range";
    if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
    if (g->graph) { cudaGraphDestroy(g->graph); g->graph = nullptr; }
    memset(&g->shard, 0, sizeof(g->shard));
    g->shard.rank = rank; g->shard.count = count; g->shard.stride = g->gp.Dxp + 32; g->shard.nll_local_const = nll_local_const;
    for (int q = 0; q < count; ++q) g->shard.mail[q] = (float2*)mail[q];
    if (count > 1) { g->fam.N = n_rows_global; g->fam.nll_const = 0.0f; }     // priors / Normal-likelihood constants see the whole dataset
    if (g->use_graph) {
        std::string e = build_graph(g);
        if (!e.empty()) return "gemm regime: " + e;
    }
    GCK(cudaDeviceSynchronize());
    return "";
}

std::string gemm_run(GemmRegime* g, const TickCfg& cfg, const OutBufs& out, int max_passes, cudaStream_t st) {
    GemmCtx h; memset(&h, 0, sizeof(h));
    h.cfg = cfg; h.out = out; h.max_passes = max_passes; h.epoch = ++g->epoch;
    GCK(cudaMemcpyAsync(g->ctx, &h, sizeof(h), cudaMemcpyHostToDevice, st));        // (pageable source: staged before the call returns)
    GCK(cudaMemsetAsync(g->tile_count, 0, (size_t)2 * g->CT * 4, st));
    std::string e = launch_tick(g, 1, st);                                            // betas of the chains that wait for a gradient
    if (!e.empty()) return e;
    if (g->use_graph) { GCK(cudaGraphLaunch(g->exec, st)); return ""; }
    // debugging aid (B200NUTS_GEMM_HOSTLOOP=1): the same loop driven from the host, one synchronisation per pass
    for (;;) {
        GemmSched s;
        GCK(cudaMemcpyAsync(&s, g->sched, sizeof(s), cudaMemcpyDeviceToHost, st));
        GCK(cudaStreamSynchronize(st));
        if (s.n_active == 0 || s.abort_flag != 0u || (max_passes > 0 && s.pass_in_run >= max_passes)) break;
        if (!(e = launch_pass(g, st)).empty()) return e;
        if (!(e = launch_reduce(g, st)).empty()) return e;
        if (!(e = launch_tick(g, 0, st)).empty()) return e;
    }
    return "";
}

std::string gemm_potential(GemmRegime* g, const float* z, float* U, float* grad, cudaStream_t st, long long* launches) {
    const int blocks = (g->C + 3) / 4;
    k_gemm_hook_begin<<<blocks, 128, 0, st>>>(g->fam, z, g->bimg, g->gp.KB, g->C, g->CT, g->active_tiles, g->sched);
    GCK(cudaGetLastError());
    std::string e = launch_pass(g, st);
    if (!e.empty()) return e;
    if (!(e = launch_reduce(g, st)).empty()) return e;
    k_gemm_hook_finish<<<blocks, 128, 0, st>>>(g->gp, g->fam, z, U, grad, g->gbeta, g->nll, g->C, g->shard, ++g->epoch);
    GCK(cudaGetLastError());
    *launches += 4;
    return "";
}

std::string gemm_sync(GemmRegime* g, cudaStream_t st, GemmStatus* status, long long* launches) {
    GemmSched s;
    GCK(cudaMemcpyAsync(&s, g->sched, sizeof(s), cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    if (launches) *launches += 4ll * (long long)(s.passes_total - g->passes_seen);
    g->passes_seen = s.passes_total;
    if (status) {
        status->abort_flag = s.abort_flag; status->passes_total = s.passes_total; status->n_active = s.n_active;
        for (int i = 0; i < 8; ++i) status->dbg[i] = s.dbg[i];
    }
    if (s.abort_flag) {
        char buf[160];
        if (s.abort_flag == 4u) snprintf(buf, sizeof(buf), "gemm regime aborted: a peer rank's likelihood sums never arrived (row-sharded exchange timed out)");
        else snprintf(buf, sizeof(buf), "gemm regime aborted: wait %u inside gemm_pass_kernel timed out (11-12 producer, 13-17 MMA issuer, 18-20 epilogue)", s.abort_flag);
        cudaMemsetAsync(&g->sched->abort_flag, 0, 4, st);
        return buf;
    }
    return "";
}

}  // namespace b2
