// Host interface of the many-chain GEMM regime (gemm_engine.cuh / gemm_engine.cu), used by the C ABI in b200nuts.cu.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "tick.cuh"
#include "families.cuh"

namespace b2 {

struct GemmRegime;

struct GemmStatus {
    unsigned int abort_flag;              // 0, or the code of the wait that timed out inside gemm_pass_kernel
    unsigned long long passes_total;      // GEMM passes executed by this handle so far
    unsigned long long dbg[8];            // [0] cycles of CTA 0 in gemm_pass_kernel, [1] its units, [2] MMA thread waiting for the epilogue
    int n_active;                         // chain tiles that still wait for gradients (non-zero after a pass-bounded run)
};

// All functions return "" or an error message.  ctl / vecs are the handle's chain state (engine-owned, see b200nuts.cu).
std::string gemm_create(GemmRegime** out, const FamilySpec& fam, int C, int Dp, int num_sms, ChainCtl* ctl, float* vecs, float* dense, long long* launches);
void gemm_destroy(GemmRegime* g);
// Enqueue a whole run: every chain that waits for a gradient is advanced until it reaches cfg.total_iters (or for
// max_passes passes).  Only enqueues; gemm_sync reports the outcome.
std::string gemm_run(GemmRegime* g, const TickCfg& cfg, const OutBufs& out, int max_passes, cudaStream_t st);
std::string gemm_potential(GemmRegime* g, const float* z, float* U, float* grad, cudaStream_t st, long long* launches);
std::string gemm_sync(GemmRegime* g, cudaStream_t st, GemmStatus* status, long long* launches);
// Row sharding (config 5): bytes of this handle's mailbox (allocated by the caller, exported to the peers), and the wiring:
// mail[q] = rank q's mailbox as seen from this device, q < count.
size_t gemm_mail_bytes(const GemmRegime* g);
std::string gemm_set_shards(GemmRegime* g, int rank, int count, void* const* mail, long long n_rows_global, float nll_local_const);
void gemm_describe(const GemmRegime* g, int* info8);    // CT, RC, KB, S, cps, NDB, Dxp, graph mode

}  // namespace b2
