"""ctypes view of the C ABI declared in include/b200nuts.h.

Loads ``numpyro_b200/csrc/libb200nuts.so`` (built in-tree by ``__graft_entry__.build()`` /
``numpyro_b200.build``).  There is deliberately no fallback: if the CUDA library is missing or
cannot be loaded, importing the engine raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200NUTS_LIB") or os.path.join(_HERE, "csrc", "libb200nuts.so")   # (override: A/B runs of two builds)

i32, i64, u32, u64, f32 = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float
vp = C.c_void_p

OK, EINVAL, ECUDA, ESTATE, EINIT = 0, -1, -2, -3, -4
FAMILY_DIAG_GAUSSIAN, FAMILY_EIGHT_SCHOOLS, FAMILY_GLM = 0, 1, 2
LIK_BERNOULLI_LOGIT, LIK_POISSON_LOG, LIK_NORMAL = 0, 1, 2
SCALE_NONE, SCALE_HALFCAUCHY, SCALE_EXPONENTIAL = 0, 1, 2
ALGO_NUTS, ALGO_HMC = 0, 1
REGIME_AUTO, REGIME_WARP, REGIME_STREAM, REGIME_GEMM = 0, 1, 2, 3
MAX_SHARDS, SHARD_HANDLE_BYTES = 16, 128


class Config(C.Structure):
    """B200NutsConfig"""
    _fields_ = [
        ("family", i32), ("num_chains", i32), ("n_rows", i64), ("n_cols", i32),
        ("X", vp), ("y", vp), ("aux", vp),
        ("likelihood", i32), ("local_scales", i32), ("global_scale", i32),
        ("group_col_begin", i32), ("group_col_end", i32), ("tau_scale", f32), ("mu_scale", f32),
        ("algo", i32), ("step_size", f32),
        ("adapt_step_size", i32), ("adapt_mass_matrix", i32), ("regularize_mass_matrix", i32),
        ("find_heuristic_step_size", i32), ("target_accept_prob", f32),
        ("max_tree_depth_warmup", i32), ("max_tree_depth", i32),
        ("hmc_num_steps", i32), ("trajectory_length", f32), ("init_radius", f32),
        ("model_built", i32), ("regime", i32),
        ("shard_rank", i32), ("shard_count", i32), ("nccl_comm", vp),
        ("n_rows_global", i64),
        ("dense_mass", i32), ("ecs_subsample_size", i32), ("ecs_proxy_degree", i32), ("reserved0", i32),
        ("cond_fixed", vp),
    ]


class Run(C.Structure):
    """B200NutsRun"""
    _fields_ = [
        ("upper", i32), ("collect_start", i32), ("thinning", i32), ("collection_size", i32),
        ("z", vp), ("diverging", vp), ("num_steps", vp), ("accept_prob", vp),
        ("mean_accept_prob", vp), ("potential_energy", vp), ("energy", vp), ("step_size", vp),
        ("max_passes", i32),
    ]


class ChainState(C.Structure):
    """B200NutsChainState"""
    _fields_ = [
        ("i", i32), ("rng_key", u32 * 2), ("potential_energy", f32), ("energy", f32),
        ("num_steps", i32), ("accept_prob", f32), ("mean_accept_prob", f32), ("diverging", i32),
        ("step_size", f32),
        ("ss_x_t", f32), ("ss_x_avg", f32), ("ss_g_avg", f32), ("ss_prox", f32), ("ss_t", i32),
        ("mm_n", i32), ("window_idx", i32), ("adapt_rng_key", u32 * 2),
        ("init_failed", i32), ("done", i32), ("total_leapfrogs", u64),
    ]


def default_config(**kw) -> Config:
    """Defaults of ``NUTS.__init__`` (numpyro/infer/hmc.py:916-951)."""
    c = Config()
    c.step_size = 1.0
    c.adapt_step_size = 1
    c.adapt_mass_matrix = 1
    c.regularize_mass_matrix = 1
    c.find_heuristic_step_size = 0
    c.target_accept_prob = 0.8
    c.max_tree_depth_warmup = 10
    c.max_tree_depth = 10
    c.trajectory_length = 6.283185307179586
    c.init_radius = 2.0
    c.model_built = 1
    c.tau_scale = 1.0
    c.mu_scale = 5.0
    c.shard_count = 1
    for k, v in kw.items():
        setattr(c, k, v)
    return c


EXPORTS = {
    "b200nuts_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
    "b200nuts_destroy": (None, [vp]),
    "b200nuts_last_error": (C.c_char_p, [vp]),
    "b200nuts_dim": (C.c_int, [vp]),
    "b200nuts_regime": (C.c_int, [vp]),
    "b200nuts_init": (C.c_int, [vp, vp, vp, i32, vp]),
    "b200nuts_run": (C.c_int, [vp, C.POINTER(Run), vp]),
    "b200nuts_transition": (C.c_int, [vp, i32, vp]),
    "b200nuts_sync": (C.c_int, [vp]),
    "b200nuts_gemm_info": (C.c_int, [vp, vp]),
    "b200nuts_state_to_device": (C.c_int, [vp, vp, vp, vp, vp]),
    "b200nuts_get_state": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "b200nuts_set_state": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, i32, vp]),
    "b200nuts_shard_export": (C.c_int, [vp, vp]),
    "b200nuts_shard_connect": (C.c_int, [vp, vp]),
    "b200nuts_potential_and_grad": (C.c_int, [vp, vp, vp, vp, vp]),
    "b200nuts_leapfrog": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, i32, vp]),
    "b200nuts_constrain": (C.c_int, [vp, vp, i64, vp, vp]),
    "b200nuts_constrained_dim": (C.c_int, [vp]),
    "b200nuts_log_likelihood": (C.c_int, [vp, vp, i64, vp, vp]),
    "b200nuts_predict": (C.c_int, [vp, vp, vp, i64, vp, vp]),
    "b200nuts_obs_count": (i64, [vp]),
    "b200nuts_set_inverse_mass_matrix": (C.c_int, [vp, vp, i32, vp]),
    "b200nuts_get_dense_state": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "b200nuts_set_dense_state": (C.c_int, [vp, vp, vp, vp]),
    "b200nuts_ecs_set_proxy": (C.c_int, [vp, vp, vp, vp, vp, f32]),
    "b200nuts_ecs_set_indices": (C.c_int, [vp, vp, vp]),
    "b200nuts_cond_set_values": (C.c_int, [vp, vp, vp]),
    "b200nuts_full_dim": (C.c_int, [vp]),
    "b200nuts_prng_split": (C.c_int, [vp, i64, i32, vp]),
    "b200nuts_prng_bits": (C.c_int, [vp, i64, vp]),
    "b200nuts_prng_uniform": (C.c_int, [vp, i64, f32, f32, vp]),
    "b200nuts_prng_normal": (C.c_int, [vp, i64, vp]),
    "b200nuts_detmath": (C.c_int, [i32, vp, i64, vp]),
    "b200nuts_launch_count": (i64, [vp]),
    "b200nuts_pass_count": (i64, [vp]),
    "b200nuts_debug_clocks": (C.c_int, [vp, vp]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the engine; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA engine first "
                "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            if os.environ.get("B200NUTS_LIB") and not hasattr(lib, name):
                continue                 # (A/B runs against an older build: entry points added since are simply absent)
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
