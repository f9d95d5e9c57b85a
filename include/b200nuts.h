/* b200nuts -- C ABI of the B200-native batched NUTS/HMC engine.
 *
 * This is the drop-in boundary for numpyro's MCMC hot path.  The entry points are
 * what an XLA-FFI / ctypes binding behind numpyro.infer.mcmc.MCMCKernel would call:
 *
 *   reference interface (numpyro 0.21.0)                          replaced by
 *   -----------------------------------------------------------  -------------------------------
 *   initialize_model -> potential_fn, ParamInfo                   b200nuts_create  (family spec +
 *     numpyro/infer/util.py:663-835                                 borrowed device data pointers)
 *   MCMCKernel.init(rng_key, num_warmup, init_params, ...)        b200nuts_init
 *     numpyro/infer/mcmc.py:90-108, hmc.py:740-799, :193-362
 *   MCMCKernel.sample(state, ...)  mcmc.py:110-124                b200nuts_transition (n_iter = 1)
 *   ... iterated by fori_collect                                  b200nuts_run (whole collection
 *     mcmc.py:466-521; numpyro/util.py:321-454                      loop, device resident) + b200nuts_sync
 *   MCMC.last_state / post_warmup_state (HMCState pytree)         b200nuts_get_state / set_state
 *     mcmc.py:558-587, hmc.py:31-48, hmc_util.py:18-30
 *   postprocess_fn (constrain + deterministic sites)              b200nuts_constrain
 *     mcmc.py:193-214, infer/util.py:177-191
 *   jax.value_and_grad(potential_fn), velocity_verlet             b200nuts_potential_and_grad,
 *     hmc_util.py:242-252, :262-311  (parity hooks)                 b200nuts_leapfrog
 *   jax.random.{split,bits,uniform,normal}  (parity hooks)        b200nuts_prng_*
 *
 * Conventions: plain C, no torch / XLA types.  Every function returns 0 on success or a negative
 * B200NUTS_E* code; b200nuts_last_error() gives the message.  Nothing throws across the ABI.
 * Device pointers are caller-owned and borrowed for the lifetime of the handle; the engine owns only
 * its per-chain scratch.  All work is enqueued on the caller's stream (cudaStream_t passed as
 * void*); calls return without synchronising unless stated: b200nuts_init / run / transition only
 * enqueue, b200nuts_sync waits for the last launch and reports its outcome (every call that needs
 * results -- get_state, the parity hooks -- synchronises by itself).  At most one launch of a handle
 * is in flight: a second b200nuts_run first collects the previous one.  There is no CPU fallback: a
 * missing GPU or an unsupported family/shape is an error.
 */
#ifndef B200NUTS_H_
#define B200NUTS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NUTS_OK 0
#define B200NUTS_EINVAL (-1)      /* bad argument / unsupported family or shape */
#define B200NUTS_ECUDA (-2)       /* CUDA runtime error (message has the cudaError string) */
#define B200NUTS_ESTATE (-3)      /* call order violated (e.g. run before init) */
#define B200NUTS_EINIT (-4)       /* "Cannot find valid initial parameters" (infer/util.py:800-832) */

enum { B200NUTS_FAMILY_DIAG_GAUSSIAN = 0, B200NUTS_FAMILY_EIGHT_SCHOOLS = 1, B200NUTS_FAMILY_GLM = 2 };
enum { B200NUTS_LIK_BERNOULLI_LOGIT = 0, B200NUTS_LIK_POISSON_LOG = 1, B200NUTS_LIK_NORMAL = 2 };
enum { B200NUTS_SCALE_NONE = 0, B200NUTS_SCALE_HALFCAUCHY = 1, B200NUTS_SCALE_EXPONENTIAL = 2 };
enum { B200NUTS_ALGO_NUTS = 0, B200NUTS_ALGO_HMC = 1 };
enum { B200NUTS_REGIME_AUTO = 0, B200NUTS_REGIME_WARP = 1, B200NUTS_REGIME_STREAM = 2, B200NUTS_REGIME_GEMM = 3 };

typedef struct B200Nuts B200Nuts;

/* Model family + sampler options.  Zero-initialise, then fill. */
typedef struct B200NutsConfig {
    /* --- model (what initialize_model would have traced) --- */
    int32_t family;            /* B200NUTS_FAMILY_* */
    int32_t num_chains;        /* chains owned by this handle (= this GPU) */
    int64_t n_rows;            /* GLM: rows of X held by this handle; eight schools: J; gaussian: D */
    int32_t n_cols;            /* GLM: columns of X */
    const float* X;            /* device, [n_rows, n_cols] row-major fp32 (GLM) */
    const float* y;            /* device, [n_rows] fp32 (GLM responses; eight schools: y_j) */
    const float* aux;          /* device; eight schools: sigma_j[J]; gaussian: mu[D] then sigma[D] */
    int32_t likelihood;        /* B200NUTS_LIK_* */
    int32_t local_scales;      /* 1: lambdas ~ HalfCauchy(1) per column (horseshoe) */
    int32_t global_scale;      /* B200NUTS_SCALE_*: prior of the global scale tau */
    int32_t group_col_begin, group_col_end;   /* columns multiplied by tau (0,0 => all) */
    float tau_scale;           /* HalfCauchy scale / 1/rate of tau (eight schools: 5) */
    float mu_scale;            /* eight schools: scale of mu's Normal prior (5) */
    /* --- kernel options (NUTS.__init__, numpyro/infer/hmc.py:916-951) --- */
    int32_t algo;              /* B200NUTS_ALGO_* */
    float step_size;           /* default 1.0 */
    int32_t adapt_step_size, adapt_mass_matrix, regularize_mass_matrix, find_heuristic_step_size;
    float target_accept_prob;  /* default 0.8 */
    int32_t max_tree_depth_warmup, max_tree_depth;   /* default 10, 10 */
    int32_t hmc_num_steps;     /* HMC: fixed number of leapfrogs (0 => use trajectory_length) */
    float trajectory_length;   /* HMC: default 2*pi */
    float init_radius;         /* init_to_uniform radius, default 2 */
    int32_t model_built;       /* 1 (default): kernel built from a model => one-block mass-matrix dict,
                                  momentum_generator splits its key once more (hmc.py:93-99) */
    int32_t regime;            /* B200NUTS_REGIME_* */
    /* --- row sharding (data-parallel likelihood, BASELINE config 5) --- */
    int32_t shard_rank, shard_count;   /* 0,1 when not sharded; <= B200NUTS_MAX_SHARDS */
    void* nccl_comm;           /* reserved (NULL): the per-gradient all-reduce runs inside the kernel over NVLink peer
                                  stores in fixed rank order (see b200nuts_shard_connect), not through NCCL */
    int64_t n_rows_global;     /* row-sharded handles: rows of the whole dataset (0 => n_rows) */
    /* --- mass matrix structure (hmc.py:916-951 dense_mass; hmc_util.py:439-515) --- */
    int32_t dense_mass;        /* 1: one dense [D, D] inverse mass matrix over all latent sites (dense_mass=True); 0: diagonal */
    /* --- energy-conserving subsampling, the inner potential of HMCECS (hmc_gibbs.py:502-690); plain GLM, warp regime --- */
    int32_t ecs_subsample_size;  /* m > 0: the likelihood is estimated from m of the n_rows rows (b200nuts_ecs_set_indices) */
    int32_t ecs_proxy_degree;    /* Taylor proxy degree 1 / 2 (b200nuts_ecs_set_proxy), 0 = no proxy (plate-scaled estimate) */
    int32_t reserved0;
    /* --- conditioning on Gibbs sites, the inner potential of HMCGibbs (hmc_gibbs.py:38-192); warp regime --- */
    const int32_t* cond_fixed;   /* HOST, [D of the full model] or NULL: 1 = coordinate belongs to a Gibbs site (whole sites only);
                                    the handle's latent dimension (b200nuts_dim) is then the number of free coordinates */
} B200NutsConfig;

/* Collection window of one run = fori_collect(lower, upper, thinning) (numpyro/util.py:321-454). */
typedef struct B200NutsRun {
    int32_t upper;             /* advance every chain until HMCState.i == upper */
    int32_t collect_start;     /* start_idx = lower + (upper - lower) % thinning */
    int32_t thinning;
    int32_t collection_size;   /* S = (upper - lower) / thinning; 0 => collect nothing */
    /* device output buffers, [num_chains, S(, D)]; NULL => field not collected */
    float* z;                  /* unconstrained samples */
    int32_t* diverging; int32_t* num_steps;
    float* accept_prob; float* mean_accept_prob; float* potential_energy; float* energy; float* step_size;
    /* streaming regime (<= 8 chains) and gemm regime: stop after this many passes over X (0 = run until every chain reached
     * `upper`).  Chains
     * pause wherever they are in their trees and the next b200nuts_run continues them -- results are identical to an
     * unbounded run; it lets a caller give every GPU the same amount of work per call. */
    int32_t max_passes;
} B200NutsRun;

/* Host mirror of HMCState + HMCAdaptState for one chain (hmc.py:31-48, hmc_util.py:18-30).
 * Vectors are [D] and live in the arrays passed to get/set_state. */
typedef struct B200NutsChainState {
    int32_t i; uint32_t rng_key[2];
    float potential_energy, energy;
    int32_t num_steps; float accept_prob, mean_accept_prob; int32_t diverging;
    float step_size;
    float ss_x_t, ss_x_avg, ss_g_avg, ss_prox; int32_t ss_t;
    int32_t mm_n, window_idx; uint32_t adapt_rng_key[2];
    int32_t init_failed; int32_t done;
    uint64_t total_leapfrogs;
} B200NutsChainState;

int b200nuts_create(const B200NutsConfig* cfg, B200Nuts** out);
void b200nuts_destroy(B200Nuts* h);
const char* b200nuts_last_error(const B200Nuts* h);      /* h may be NULL: last create() error */
int b200nuts_dim(const B200Nuts* h);                     /* latent dimension D (flat, sorted sites) */
int b200nuts_regime(const B200Nuts* h);

/* keys: host uint32 [num_chains][2] (rows of random.split(key, C), mcmc.py:670-671).
 * z0: device [num_chains][D] unconstrained init_params, or NULL => init_to_uniform. */
int b200nuts_init(B200Nuts* h, const uint32_t* keys, const float* z0, int32_t num_warmup, void* stream);
int b200nuts_run(B200Nuts* h, const B200NutsRun* run, void* stream);          /* enqueues; see b200nuts_sync */
/* MCMCKernel.sample (mcmc.py:110-124): advance every chain by n_iter transitions from its current state, collecting
 * nothing; read the new HMCState with b200nuts_get_state.  Enqueues. */
int b200nuts_transition(B200Nuts* h, int32_t n_iter, void* stream);
/* Wait for the last enqueued launch of this handle; returns its outcome (B200NUTS_ECUDA + message when a bounded wait
 * inside a persistent kernel timed out or the launch failed). */
int b200nuts_sync(B200Nuts* h);

/* Enqueue-only export of the HMCState fields MCMC collects, into caller-owned DEVICE buffers (any may be NULL):
 * z, z_grad [num_chains][D]; scalars [num_chains][B200NUTS_STATE_SCALARS] = {i, potential_energy, energy, num_steps,
 * accept_prob, mean_accept_prob, diverging, step_size} as floats.  Stream-ordered after b200nuts_transition / run; this
 * is what the XLA-FFI handler returns (ffi/b200nuts_ffi.cc). */
#define B200NUTS_STATE_SCALARS 8
int b200nuts_state_to_device(B200Nuts* h, float* z, float* z_grad, float* scalars, void* stream);

/* Synchronising. vectors (host, each [num_chains][D], may be NULL): z, z_grad, inverse_mass_matrix,
 * mass_matrix_sqrt, welford mean, welford m2. */
int b200nuts_get_state(B200Nuts* h, B200NutsChainState* states, float* z, float* z_grad, float* inv_mass,
                       float* mass_sqrt, float* wf_mean, float* wf_m2, void* stream);
int b200nuts_set_state(B200Nuts* h, const B200NutsChainState* states, const float* z, const float* z_grad,
                       const float* inv_mass, const float* wf_mean, const float* wf_m2, int32_t num_warmup,
                       void* stream);

/* Row-sharded handles (BASELINE config 5; the reference analogue is a likelihood over GSPMD-sharded model
 * arguments, numpyro/infer/mcmc.py:240-266, where XLA inserts the all-reduce of log-density and gradient).
 * Every rank creates a streaming-regime handle over ITS rows with the same chains, keys and options
 * (shard_rank / shard_count / n_rows_global set).  Each handle owns a small mailbox in device memory; after
 * b200nuts_shard_connect every rank stores its per-chain likelihood sums {value, tag} into every rank's
 * mailbox once per gradient and adds the shard_count contributions in rank order, so all ranks see
 * bit-identical (U, grad) and their replicated chains take identical decisions.  Priors are added after the sum.
 * export/connect are host calls: exchange the blobs with any host-side transport (torch.distributed
 * all_gather in numpyro_b200/engine.py).  Ranks may be processes (CUDA IPC) or threads of one process. */
#define B200NUTS_MAX_SHARDS 16
#define B200NUTS_SHARD_HANDLE_BYTES 128
int b200nuts_shard_export(B200Nuts* h, void* blob /* B200NUTS_SHARD_HANDLE_BYTES */);
int b200nuts_shard_connect(B200Nuts* h, const void* blobs /* [shard_count][B200NUTS_SHARD_HANDLE_BYTES], rank order */);

/* Parity hooks (device pointers). z,g: [num_chains][D]; U: [num_chains]. */
int b200nuts_potential_and_grad(B200Nuts* h, const float* z, float* U, float* g, void* stream);
/* n_steps velocity-Verlet steps with per-chain step size eps[C] and inverse mass inv_mass[C][D];
 * z, r updated in place; U, g receive the final potential / gradient. */
int b200nuts_leapfrog(B200Nuts* h, const float* eps, const float* inv_mass, float* z, float* r, float* U,
                      float* g, int32_t n_steps, void* stream);
/* unconstrained [n][D] -> constrained latent + deterministic sites, concatenated per row in the order
 * reported by b200nuts_constrained_layout (device pointers). */
int b200nuts_constrain(B200Nuts* h, const float* z, int64_t n, float* out, void* stream);
int b200nuts_constrained_dim(const B200Nuts* h);

/* The consumers of the collected samples (numpyro/infer/util.py log_likelihood :1133-1188,
 * Predictive :927-1131) for the registered families.  z: device [n][D] unconstrained samples; out: device
 * [n][b200nuts_obs_count] -- the observed site's log_prob per observation, or a draw from it with the per-sample key
 * keys[n][2] (device) that numpyro's seed handler would hand to that site.  Enqueue only. */
int b200nuts_log_likelihood(B200Nuts* h, const float* z, int64_t n, float* out, void* stream);
int b200nuts_predict(B200Nuts* h, const float* z, const uint32_t* keys, int64_t n, float* out, void* stream);
int64_t b200nuts_obs_count(const B200Nuts* h);

/* Mass matrix (numpyro/infer/hmc_util.py:439-515, hmc.py:759-769).
 * b200nuts_set_inverse_mass_matrix: the kernel's `inverse_mass_matrix=` argument, the same for every chain: host fp32,
 * ndim 1 => [D] (a dense handle puts it on the diagonal), ndim 2 => [D][D] row-major (a diagonal handle keeps its diagonal);
 * call before b200nuts_init.  b200nuts_get/set_dense_state: HMCAdaptState.inverse_mass_matrix / mass_matrix_sqrt /
 * mass_matrix_sqrt_inv / the Welford m2 of a dense handle, host fp32 [num_chains][D][D] each (NULL = skip); set ignores the
 * two roots and recomputes them from inverse_mass_matrix. */
int b200nuts_set_inverse_mass_matrix(B200Nuts* h, const float* imm, int32_t ndim, void* stream);
int b200nuts_get_dense_state(B200Nuts* h, float* inverse_mass_matrix, float* mass_matrix_sqrt, float* mass_matrix_sqrt_inv,
                             float* wf_m2, void* stream);
int b200nuts_set_dense_state(B200Nuts* h, const float* inverse_mass_matrix, const float* wf_m2, void* stream);

/* HMCECS inner potential (numpyro/infer/hmc_gibbs.py:577-682, contrib/ecs_proxies.py:23-300).
 * set_proxy: the Taylor proxy's reference point and the full-data terms at it -- device fp32: ref [n_cols], eta_ref [n_rows]
 * (= X ref), G [n_cols] and H [n_cols][n_cols] (gradient / Hessian of the full-data log-likelihood at ref; H may be NULL for
 * degree 1), L0 = the full-data log-likelihood at ref; the pointers are borrowed.  set_indices: the current subsample of every
 * chain, host int32 [num_chains][m] (the Gibbs site of HMCECS); every later potential / transition uses it. */
int b200nuts_ecs_set_proxy(B200Nuts* h, const float* ref, const float* eta_ref, const float* G, const float* H, float L0);
int b200nuts_ecs_set_indices(B200Nuts* h, const int32_t* idx, void* stream);

/* HMCGibbs inner potential (hmc_gibbs.py:153-186): current values of the Gibbs sites of every chain, host fp32
 * [num_chains][D_full + 1] -- unconstrained values at the fixed coordinates (free ones ignored) and, last, the amount added
 * to the potential (sum of the unconstrained values of fixed positive sites: the conditioned model has no Jacobian term for
 * them).  b200nuts_constrain of a conditioned handle takes vectors of the FULL model. */
int b200nuts_cond_set_values(B200Nuts* h, const float* values, void* stream);
int b200nuts_full_dim(const B200Nuts* h);

/* PRNG parity hooks: host in / host out, computed on the device, synchronising. */
int b200nuts_prng_split(const uint32_t* keys, int64_t n_keys, int32_t num, uint32_t* out);      /* out [n_keys][num][2] */
int b200nuts_prng_bits(const uint32_t* key, int64_t n, uint32_t* out);
int b200nuts_prng_uniform(const uint32_t* key, int64_t n, float lo, float hi, float* out);
int b200nuts_prng_normal(const uint32_t* key, int64_t n, float* out);
/* det-f32 math hooks: op 0 exp, 1 log, 2 log1p, 3 expit, 4 erfinv */
int b200nuts_detmath(int32_t op, const float* x, int64_t n, float* out);

/* number of kernels launched by this handle so far (bench.py's gpu_launches) */
int64_t b200nuts_launch_count(const B200Nuts* h);
/* streaming regime: sweeps over X executed so far by b200nuts_run (one sweep serves one gradient of every chain) */
int64_t b200nuts_pass_count(const B200Nuts* h);
/* gemm regime: {chain tiles, row chunks, k-blocks, row segments, chunks per segment, column blocks, padded columns,
 * 1 = device-side WHILE graph / 0 = host loop} */
int b200nuts_gemm_info(const B200Nuts* h, int32_t* out8);
/* streaming regime: clock64 totals of the last run on CTA 0 (see StreamSync.dbg in stream_engine.cuh); gemm regime: GemmStatus.dbg */
int b200nuts_debug_clocks(const B200Nuts* h, uint64_t* out16);

#ifdef __cplusplus
}
#endif
#endif /* B200NUTS_H_ */
