"""numpyro's MCMC front end over the B200 engine.

Mirrors the user-facing surface of numpyro/infer/mcmc.py (``MCMC`` :225-809, ``MCMCKernel``
:33-159) and numpyro/infer/hmc.py (``HMC`` :533-822, ``NUTS`` :825-951, ``HMCState`` :31-48) for
the hot path only: same constructor arguments, same ``run / warmup / get_samples /
get_extra_fields / last_state / post_warmup_state / print_summary`` semantics and the same
collection layout ``[num_chains, num_samples // thinning, ...]``.  ``NUTS`` / ``HMC`` implement the
``MCMCKernel`` plug-in interface (``init`` / ``sample`` / ``postprocess_fn``, mcmc.py:79-124) over the
C ABI (b200nuts_init / b200nuts_transition / b200nuts_constrain); when numpyro is importable they
subclass the real ``MCMCKernel``.  The model argument is a declared family
(numpyro_b200.families); arrays are NumPy on the host, every computation runs in the CUDA engine.
``chain_method='parallel'`` shards chains over the GPUs visible to the process (or over
torch.distributed ranks under torchrun) with no communication except the final gather;
``'vectorized'`` keeps all chains on one GPU; ``'sequential'`` runs them one after the other.
"""
from __future__ import annotations

import os
import warnings
from collections import namedtuple
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _capi, diagnostics, families, random as b2random
from .engine import Engine

try:                                   # drop-in: be a real numpyro kernel where numpyro exists
    from numpyro.infer.mcmc import MCMCKernel as _KernelBase          # pragma: no cover (jax is not in this image)
except Exception:                      # noqa: BLE001
    class _KernelBase:                 # the interface of mcmc.py:33-159 is restated by HMC below
        pass

HMCState = namedtuple("HMCState", ["i", "z", "z_grad", "potential_energy", "energy", "r", "trajectory_length",
                                   "num_steps", "accept_prob", "mean_accept_prob", "diverging", "adapt_state", "rng_key"])
HMCAdaptState = namedtuple("HMCAdaptState", ["step_size", "inverse_mass_matrix", "mass_matrix_sqrt",
                                             "mass_matrix_sqrt_inv", "ss_state", "mm_state", "window_idx", "rng_key"])

_STATE_FIELDS = {"potential_energy": "potential_energy", "energy": "energy", "num_steps": "num_steps",
                 "accept_prob": "accept_prob", "mean_accept_prob": "mean_accept_prob", "diverging": "diverging",
                 "adapt_state.step_size": "step_size"}


# ---------------------------------------------------------------------- init strategies (initialization.py:88-160)
class _InitStrategy:
    def __init__(self, kind, radius=2.0, values=None):
        self.kind, self.radius, self.values = kind, float(radius), values


def init_to_uniform(radius=2):
    """initialization.py:88-122: uniform(-radius, radius) in unconstrained space (the default, radius 2)."""
    return _InitStrategy("uniform", radius=radius)


def init_to_feasible():
    """initialization.py:125-130 = init_to_uniform(radius=0): every unconstrained coordinate starts at 0."""
    return _InitStrategy("feasible", radius=0.0)


def init_to_value(values=None):
    """initialization.py:140-155: start from the given (constrained) site values.  Every latent site must be given:
    the reference's fallback for missing sites draws through the seed handler's key stream, which is not restated."""
    return _InitStrategy("value", values=dict(values or {}))


# ---------------------------------------------------------------------- HMCState <-> engine
def _state_from_engine(sts, vecs, bound, squeeze_one, trajectory_length, dense=None):
    """HMCState (hmc.py:31-48) of the chains behind one or more engine handles, field for field.  ``dense``: the handles'
    :meth:`Engine.dense_state` dicts when the kernel uses ``dense_mass=True`` (matrices [C, D, D] instead of vectors)."""
    cat = lambda name: np.concatenate([v[name] for v in vecs], axis=0)
    st = [s[k] for s in sts for k in range(len(s))]
    arr = lambda f, dt: np.array([getattr(s, f) for s in st], dt)
    unflat = lambda a: {s.name: a[:, s.z_offset:s.z_offset + s.size].reshape((a.shape[0],) + tuple(s.shape))
                        for s in bound.latent_sites}
    squeeze = (lambda a: a[0]) if squeeze_one else (lambda a: a)
    tree = lambda d: {k: squeeze(v) for k, v in d.items()}
    key = tuple(sorted(s.name for s in bound.latent_sites))        # one-block structured mass matrix (hmc.py:759-769)
    if dense is not None:
        dcat = lambda name: np.concatenate([d[name] for d in dense], axis=0)
        imm, msq, msq_inv, m2 = (dcat(n) for n in ("inverse_mass_matrix", "mass_matrix_sqrt", "mass_matrix_sqrt_inv", "wf_m2"))
    else:
        imm, msq, m2 = cat("inverse_mass_matrix"), cat("mass_matrix_sqrt"), cat("wf_m2")
        msq_inv = 1.0 / msq
    adapt = HMCAdaptState(
        squeeze(arr("step_size", np.float32)), {key: squeeze(imm)}, {key: squeeze(msq)}, {key: squeeze(msq_inv)},
        (squeeze(arr("ss_x_t", np.float32)), squeeze(arr("ss_x_avg", np.float32)), squeeze(arr("ss_g_avg", np.float32)),
         squeeze(arr("ss_t", np.int32)), squeeze(arr("ss_prox", np.float32))),
        {key: (squeeze(cat("wf_mean")), squeeze(m2), squeeze(arr("mm_n", np.int32)))},
        squeeze(arr("window_idx", np.int32)),
        squeeze(np.array([[s.adapt_rng_key[0], s.adapt_rng_key[1]] for s in st], np.uint32)))
    return HMCState(squeeze(arr("i", np.int32)), tree(unflat(cat("z"))), tree(unflat(cat("z_grad"))),
                    squeeze(arr("potential_energy", np.float32)), squeeze(arr("energy", np.float32)), None,
                    trajectory_length, squeeze(arr("num_steps", np.int32)),
                    squeeze(arr("accept_prob", np.float32)), squeeze(arr("mean_accept_prob", np.float32)),
                    squeeze(arr("diverging", np.int32).astype(bool)), adapt,
                    squeeze(np.array([[s.rng_key[0], s.rng_key[1]] for s in st], np.uint32)))


def _state_to_engine(state: HMCState, bound, engine: Engine, n_total: int, lo: int, hi: int, num_warmup: int, keys=None):
    """Load chains [lo, hi) of ``state`` (which holds ``n_total`` chains) into ``engine`` (b200nuts_set_state);
    ``keys`` replaces ``state.rng_key`` (mcmc.py:677-679)."""
    C = n_total
    D = engine.D

    def flat(tree):
        out = np.zeros((C, D), np.float32)
        for s in bound.latent_sites:
            out[:, s.z_offset:s.z_offset + s.size] = np.asarray(tree[s.name], np.float32).reshape(C, s.size)
        return out
    z, g = flat(state.z), flat(state.z_grad)
    a = state.adapt_state
    one = lambda d: np.asarray(next(iter(d.values())) if isinstance(d, dict) else d, np.float32).reshape(C, -1)
    imm = one(a.inverse_mass_matrix)
    mm = next(iter(a.mm_state.values())) if isinstance(a.mm_state, dict) else a.mm_state
    dense = imm.shape[1] == D * D and D > 1
    imm_full, m2_full = imm, np.asarray(mm[1], np.float32).reshape(C, -1)
    if dense:                       # the vector slots of the C ABI carry the diagonals; the matrices follow below
        imm = imm_full.reshape(C, D, D)[:, np.arange(D), np.arange(D)]
        mm = (mm[0], m2_full.reshape(C, D, D)[:, np.arange(D), np.arange(D)], mm[2])
    sc = lambda v, dt: np.asarray(v, dt).reshape(C)
    rk = np.asarray(state.rng_key if keys is None else keys, np.uint32).reshape(C, 2)
    akeys = np.asarray(a.rng_key, np.uint32).reshape(C, 2)
    n = hi - lo
    st = (_capi.ChainState * n)()
    for k in range(n):
        c = lo + k
        st[k].i = int(sc(state.i, np.int32)[c]); st[k].rng_key[0], st[k].rng_key[1] = int(rk[c][0]), int(rk[c][1])
        st[k].potential_energy = float(sc(state.potential_energy, np.float32)[c]); st[k].energy = float(sc(state.energy, np.float32)[c])
        st[k].num_steps = int(sc(state.num_steps, np.int32)[c]); st[k].accept_prob = float(sc(state.accept_prob, np.float32)[c])
        st[k].mean_accept_prob = float(sc(state.mean_accept_prob, np.float32)[c]); st[k].diverging = int(sc(state.diverging, np.int32)[c])
        st[k].step_size = float(sc(a.step_size, np.float32)[c])
        st[k].ss_x_t, st[k].ss_x_avg, st[k].ss_g_avg = (float(sc(a.ss_state[j], np.float32)[c]) for j in range(3))
        st[k].ss_t = int(sc(a.ss_state[3], np.int32)[c]); st[k].ss_prox = float(sc(a.ss_state[4], np.float32)[c])
        st[k].mm_n = int(sc(mm[2], np.int32)[c]); st[k].window_idx = int(sc(a.window_idx, np.int32)[c])
        st[k].adapt_rng_key[0], st[k].adapt_rng_key[1] = int(akeys[c][0]), int(akeys[c][1])
    vec = {"z": z[lo:hi], "z_grad": g[lo:hi], "inverse_mass_matrix": imm[lo:hi],
           "wf_mean": np.asarray(mm[0], np.float32).reshape(C, -1)[lo:hi],
           "wf_m2": np.asarray(mm[1], np.float32).reshape(C, -1)[lo:hi]}
    engine.set_state(st, vec, num_warmup)
    if dense:
        engine.set_dense_state(imm_full.reshape(C, D, D)[lo:hi], m2_full.reshape(C, D, D)[lo:hi])


def _flatten_init(init_params, bound, D):
    if isinstance(init_params, dict):
        lat = bound.latent_sites
        name0 = next(iter(init_params))
        first = np.asarray(init_params[name0])
        lead = first.shape[:first.ndim - len(next(s for s in lat if s.name == name0).shape)]
        n = int(np.prod(lead)) if lead else 1
        z = np.zeros((n, D), np.float32)
        for s in lat:
            z[:, s.z_offset:s.z_offset + s.size] = np.asarray(init_params[s.name], np.float32).reshape(n, s.size)
        return z
    return np.asarray(init_params, np.float32).reshape(-1, D)


class HMC(_KernelBase):
    """Hamiltonian Monte Carlo kernel (hmc.py:533-822) for a declared model family."""
    _algo = _capi.ALGO_HMC

    def __init__(self, model=None, potential_fn=None, kinetic_fn=None, step_size=1.0, inverse_mass_matrix=None,
                 adapt_step_size=True, adapt_mass_matrix=True, dense_mass=False, target_accept_prob=0.8,
                 num_steps=None, trajectory_length=2 * np.pi, init_strategy=None, find_heuristic_step_size=False,
                 forward_mode_differentiation=False, regularize_mass_matrix=True, max_tree_depth=10, regime="auto"):
        if potential_fn is not None or kinetic_fn is not None:
            raise NotImplementedError("the engine fuses the potential of registered model families; "
                                      "arbitrary potential_fn / kinetic_fn callables are not supported")
        if not isinstance(model, families.Model):
            raise TypeError("model must be a numpyro_b200.families.Model (a declared model family)")
        if not isinstance(dense_mass, (bool, np.bool_)):
            raise NotImplementedError("block-structured mass matrices (dense_mass as a list of site tuples) are not implemented; "
                                      "dense_mass=True / False and inverse_mass_matrix= (vector or matrix) are")
        if isinstance(inverse_mass_matrix, dict):
            if len(inverse_mass_matrix) != 1:
                raise NotImplementedError("inverse_mass_matrix: only one block over all latent sites is implemented")
            inverse_mass_matrix = next(iter(inverse_mass_matrix.values()))
        self._dense_mass = bool(dense_mass)
        self._inverse_mass_matrix = None if inverse_mass_matrix is None else np.asarray(inverse_mass_matrix, np.float32)
        if forward_mode_differentiation:
            raise NotImplementedError("gradients are hand-derived; forward_mode_differentiation does not apply")
        if init_strategy is None:
            init_strategy = init_to_uniform()
        if not isinstance(init_strategy, _InitStrategy):
            raise NotImplementedError("init_strategy must be numpyro_b200.infer.init_to_uniform / init_to_feasible / init_to_value")
        self._init_strategy = init_strategy
        self._model = model
        depth = max_tree_depth if isinstance(max_tree_depth, tuple) else (max_tree_depth, max_tree_depth)
        self._cfg = dict(algo=self._algo, step_size=float(step_size), adapt_step_size=int(adapt_step_size),
                         adapt_mass_matrix=int(adapt_mass_matrix), regularize_mass_matrix=int(regularize_mass_matrix),
                         find_heuristic_step_size=int(find_heuristic_step_size), target_accept_prob=float(target_accept_prob),
                         max_tree_depth_warmup=int(depth[0]), max_tree_depth=int(depth[1]),
                         hmc_num_steps=int(num_steps or 0),
                         trajectory_length=float(trajectory_length if num_steps is None else 0.0) or 2 * np.pi,
                         regime={"auto": 0, "warp": 1, "stream": 2, "gemm": 3}[regime], dense_mass=int(self._dense_mass))
        if init_strategy.kind == "uniform":
            self._cfg["init_radius"] = init_strategy.radius
        self._trajectory_length = None if num_steps is not None else trajectory_length
        # MCMCKernel state (init / sample driven one transition at a time)
        self._engine: Optional[Engine] = None
        self._bound = None
        self._num_warmup = 0
        self._single = True
        self._engine_state = None         # the HMCState object the engine's device state corresponds to

    # ------------------------------------------------------------------ kernel properties (hmc.py:715-738)
    @property
    def model(self):
        return self._model

    @property
    def sample_field(self):
        return "z"

    @property
    def default_fields(self):
        return ("z", "diverging")

    @property
    def is_ensemble_kernel(self):
        return False

    def get_diagnostics_str(self, state):
        return "{} steps of size {:.2e}. acc. prob={:.2f}".format(
            np.ravel(state.num_steps)[0], np.ravel(state.adapt_state.step_size)[0], np.ravel(state.mean_accept_prob)[0])

    # ------------------------------------------------------------------ init strategy -> unconstrained start
    def _strategy_z0(self, bound, D, n_chains):
        """None for a PRNG-drawn start, or the [n_chains, D] unconstrained start the strategy fixes."""
        st = self._init_strategy
        if st.kind == "uniform":
            return None
        if st.kind == "feasible":
            return np.zeros((n_chains, D), np.float32)
        missing = [s.name for s in bound.latent_sites if s.name not in st.values]
        if missing:
            raise NotImplementedError(f"init_to_value needs a value for every latent site (missing: {missing})")
        z = np.zeros((1, D), np.float32)
        for s in bound.latent_sites:
            v = np.asarray(st.values[s.name], np.float64).reshape(s.size)
            if s.positive:
                if np.any(v <= 0):
                    raise ValueError(f"init_to_value: site {s.name!r} must be positive")
                v = np.log(v)                           # inverse of ExpTransform (transforms.py:635-646)
            z[0, s.z_offset:s.z_offset + s.size] = v
        return np.repeat(z, n_chains, axis=0)

    def __getstate__(self):
        """hmc.py:818-822: the compiled closures (here: the engine handle and its device state) are not pickled."""
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_engine_state"] = None
        return state

    # ------------------------------------------------------------------ MCMCKernel (mcmc.py:79-124)
    def init(self, rng_key, num_warmup, init_params=None, model_args=(), model_kwargs=None):
        """``MCMCKernel.init`` (mcmc.py:90-108; hmc.py:740-799): bind the model, find valid initial parameters and
        return the initial ``HMCState``.  A batch of keys ``[C, 2]`` gives a vectorised state of C chains."""
        keys = np.asarray(rng_key, np.uint32)
        self._single = keys.ndim == 1
        keys = keys.reshape(-1, 2)
        bound = self._model.bind(*model_args, **(model_kwargs or {}))
        if self._engine is not None:
            self._engine.close()
        cfg = dict(self._cfg)
        cfg.update(bound.cfg)
        cfg["num_chains"] = keys.shape[0]
        e = Engine(device=torch.device("cuda", torch.cuda.current_device()), X=bound.X, y=bound.y, aux=bound.aux, **cfg)
        z0 = _flatten_init(init_params, bound, e.D) if init_params is not None else self._strategy_z0(bound, e.D, keys.shape[0])
        if self._inverse_mass_matrix is not None:
            e.set_inverse_mass_matrix(self._inverse_mass_matrix)
        e.init(keys, int(num_warmup), z0)
        e.run(0, 0, fields=())                      # evaluates the potential at the start (retrying invalid draws): HMCState.i == 0
        self._engine, self._bound, self._num_warmup = e, bound, int(num_warmup)
        self._engine_state = self._read_state()
        return self._engine_state

    def _read_state(self):
        e = self._engine
        st, vec = e.state()
        return _state_from_engine([st], [vec], self._bound, self._single, self._trajectory_length,
                                  dense=[e.dense_state()] if self._dense_mass else None)

    def sample(self, state, model_args=(), model_kwargs=None):
        """``MCMCKernel.sample`` (mcmc.py:110-124; hmc.py:801-816): one transition from ``state``."""
        e = self._engine
        if e is None:
            raise RuntimeError("sample() called before init()")
        if state is not self._engine_state:          # a state the engine does not hold (e.g. HMCGibbs edited z): load it
            _state_to_engine(state, self._bound, e, e.C, 0, e.C, self._num_warmup)
        e.transition(1)
        self._engine_state = self._read_state()
        return self._engine_state

    def postprocess_fn(self, model_args=(), model_kwargs=None):
        """mcmc.py:79-88 / hmc.py:712-713: unconstrained ``z`` dict -> constrained sites + deterministic sites."""
        bound = self._bound if self._bound is not None else self._model.bind(*model_args, **(model_kwargs or {}))
        e = self._engine
        if e is None:
            raise RuntimeError("postprocess_fn needs init() (the transform runs in the engine)")

        def fn(z):
            first = np.asarray(z[bound.latent_sites[0].name])
            lead = first.shape[:first.ndim - len(bound.latent_sites[0].shape)]
            n = int(np.prod(lead)) if lead else 1
            flat = np.zeros((n, e.D), np.float32)
            for s in bound.latent_sites:
                flat[:, s.z_offset:s.z_offset + s.size] = np.asarray(z[s.name], np.float32).reshape(n, s.size)
            con = e.constrain(torch.from_numpy(flat).to(e.device)).cpu().numpy()
            return {s.name: con[:, s.c_offset:s.c_offset + s.size].reshape(tuple(lead) + tuple(s.shape)) for s in bound.sites}
        return fn


class NUTS(HMC):
    """No-U-Turn sampler (hmc.py:825-951)."""
    _algo = _capi.ALGO_NUTS

    def __init__(self, model=None, potential_fn=None, kinetic_fn=None, step_size=1.0, inverse_mass_matrix=None,
                 adapt_step_size=True, adapt_mass_matrix=True, dense_mass=False, target_accept_prob=0.8,
                 trajectory_length=None, max_tree_depth=10, init_strategy=None, find_heuristic_step_size=False,
                 forward_mode_differentiation=False, regularize_mass_matrix=True, regime="auto"):
        super().__init__(model=model, potential_fn=potential_fn, kinetic_fn=kinetic_fn, step_size=step_size,
                         inverse_mass_matrix=inverse_mass_matrix, adapt_step_size=adapt_step_size,
                         adapt_mass_matrix=adapt_mass_matrix, dense_mass=dense_mass,
                         target_accept_prob=target_accept_prob, num_steps=None, trajectory_length=2 * np.pi,
                         init_strategy=init_strategy, find_heuristic_step_size=find_heuristic_step_size,
                         forward_mode_differentiation=forward_mode_differentiation,
                         regularize_mass_matrix=regularize_mass_matrix, max_tree_depth=max_tree_depth, regime=regime)
        self._trajectory_length = None


class _Shard:
    """The chains [lo, hi) living on one device."""

    def __init__(self, device, lo, hi):
        self.device, self.lo, self.hi = device, lo, hi
        self.engine: Optional[Engine] = None


class MCMC:
    """numpyro.infer.MCMC (mcmc.py:225-809) over the B200 engine."""

    def __init__(self, sampler, *, num_warmup, num_samples, num_chains=1, thinning=1, postprocess_fn=None,
                 chain_method="parallel", progress_bar=True, progress_rate=None, jit_model_args=False, row_shards=1):
        """``row_shards`` (extension; the reference's analogue is passing GSPMD-sharded model arguments,
        mcmc.py:240-266): split the rows of a tall GLM dataset over ``row_shards`` GPUs of this process; every GPU runs
        all chains over its rows and the per-gradient all-reduce happens inside the kernels (BASELINE config 5)."""
        self._generic = not isinstance(sampler, HMC)
        if self._generic and not (hasattr(sampler, "init") and hasattr(sampler, "sample") and hasattr(sampler, "inner_kernel")):
            raise TypeError("sampler must be numpyro_b200.infer.NUTS / HMC or a Gibbs kernel over one (numpyro_b200.hmc_gibbs)")
        if not isinstance(num_warmup, int) or num_warmup < 0:
            raise ValueError("num_warmup must be a non-negative integer")
        if thinning < 1:
            raise ValueError("thinning must be a positive integer")
        if chain_method not in ("parallel", "vectorized", "sequential"):
            raise ValueError("Only supporting the following methods to draw chains: 'sequential', 'parallel', or 'vectorized'")
        if postprocess_fn is not None:
            raise NotImplementedError("postprocess_fn is derived from the declared family")
        self.sampler, self.num_warmup, self.num_samples = sampler, num_warmup, num_samples
        self.num_chains, self.thinning, self.chain_method = num_chains, thinning, chain_method
        self.row_shards = int(row_shards)
        if self.row_shards < 1:
            raise ValueError("row_shards must be a positive integer")
        self.progress_bar = False          # the whole collection loop is device resident (util.py:411-416 path)
        self._shards: List[_Shard] = []
        self._bound = None
        self._states = None
        self._states_flat = None
        self._last_state = None
        self._warmup_state = None
        self._collect_warmup = False
        self._args, self._kwargs = (), {}
        self._dist = torch.distributed.is_available() and torch.distributed.is_initialized() and chain_method == "parallel"

    # ------------------------------------------------------------------ plumbing
    def _plan_shards(self):
        C = self.num_chains
        cur = torch.cuda.current_device() if torch.cuda.is_available() else 0
        if self.row_shards > 1:
            G, ndev = self.row_shards, torch.cuda.device_count()
            if ndev < G and not os.environ.get("B200NUTS_GRID"):
                raise ValueError(f"row_shards={G} needs {G} GPUs (found {ndev})")
            # (with B200NUTS_GRID set -- a testing aid -- the ranks share the current device, each on a part of its SMs)
            return [_Shard(torch.device("cuda", k if ndev >= G else cur), 0, C) for k in range(G)]
        if self._dist:
            W, r = torch.distributed.get_world_size(), torch.distributed.get_rank()
            if C % W:
                raise ValueError("num_chains must be divisible by the number of ranks")
            per = C // W
            return [_Shard(torch.device("cuda", cur), r * per, (r + 1) * per)]
        if self.chain_method == "sequential":
            return [_Shard(torch.device("cuda", cur), c, c + 1) for c in range(C)]
        ndev = torch.cuda.device_count() if self.chain_method == "parallel" else 1
        ndev = max(1, min(ndev, C))
        if self.chain_method == "parallel" and ndev < min(C, 2) and C > 1:
            warnings.warn("There are not enough devices to run parallel chains: the chains share one GPU "
                          "(equivalent to chain_method='vectorized').", stacklevel=3)
        first = cur if ndev == 1 else 0
        bounds = [C * k // ndev for k in range(ndev + 1)]
        return [_Shard(torch.device("cuda", first + k), bounds[k], bounds[k + 1]) for k in range(ndev)]

    def _ensure_engines(self, args, kwargs):
        bound = self.sampler.model.bind(*args, **kwargs)
        self._bound = bound
        if self._shards:
            for s in self._shards:
                if s.engine is not None:
                    s.engine.close()
        self._shards = self._plan_shards()
        G = self.row_shards
        if G > 1:
            if bound.X is None or bound.y is None:
                raise ValueError("row_shards needs a GLM family with a design matrix")
            N = int(bound.X.shape[0])
            cuts = [N * k // G for k in range(G + 1)]
        for k, s in enumerate(self._shards):
            cfg = dict(self.sampler._cfg)
            cfg.update(bound.cfg)
            cfg["num_chains"] = s.hi - s.lo
            if G > 1:
                cfg.update(shard_rank=k, shard_count=G, n_rows_global=N)      # (the engine picks the streaming or the GEMM regime)
                s.engine = Engine(device=s.device, X=bound.X[cuts[k]:cuts[k + 1]], y=bound.y[cuts[k]:cuts[k + 1]], aux=bound.aux, **cfg)
                s.stream = torch.cuda.Stream(device=s.device)     # the ranks' persistent kernels must run concurrently
            else:
                s.engine = Engine(device=s.device, X=bound.X, y=bound.y, aux=bound.aux, **cfg)
        if G > 1:
            blobs = [s.engine.shard_blob() for s in self._shards]
            for s in self._shards:
                s.engine.connect_shards(blobs)

    def _for_each_shard(self, fn):
        if len(self._shards) == 1 or (self.chain_method == "sequential" and self.row_shards == 1):
            return [fn(s) for s in self._shards]
        with ThreadPoolExecutor(len(self._shards)) as pool:
            return list(pool.map(fn, self._shards))

    @property
    def _local_lo(self):
        return self._shards[0].lo if (self._dist and self._shards) else 0

    @property
    def _local_chains(self):
        """Chains whose state this process holds (all of them except under torch.distributed)."""
        if self._dist and self._shards:
            return self._shards[0].hi - self._shards[0].lo
        return self.num_chains

    # ------------------------------------------------------------------ run / warmup
    def _chain_keys(self, rng_key):
        rng_key = np.asarray(rng_key, np.uint32)
        if rng_key.ndim == 2:
            if rng_key.shape[0] != self.num_chains:
                raise ValueError("a batch of keys must have num_chains rows")
            return rng_key
        return rng_key[None] if self.num_chains == 1 else b2random.split(rng_key, self.num_chains)   # mcmc.py:670-671

    def warmup(self, rng_key, *args, extra_fields=(), collect_warmup=False, init_params=None, **kwargs):
        """mcmc.py:589-633: run the adaptation phase only and keep its last state."""
        self._warmup_state = None
        self._collect_warmup = collect_warmup
        self._run(rng_key, args, kwargs, extra_fields, init_params, lower=0 if collect_warmup else self.num_warmup,
                  upper=self.num_warmup)
        self._warmup_state = self._last_state

    def run(self, rng_key, *args, extra_fields=(), init_params=None, **kwargs):
        """mcmc.py:635-729.  With a ``post_warmup_state`` the run starts from that state (its ``rng_key`` replaced by the
        new one, mcmc.py:677-679) and draws ``num_samples`` MORE samples, whatever the state's iteration counter is."""
        if self._warmup_state is not None:
            self._run(rng_key, args, kwargs, extra_fields, init_params, resume=self._warmup_state)
        else:
            self._run(rng_key, args, kwargs, extra_fields, init_params, lower=self.num_warmup,
                      upper=self.num_warmup + self.num_samples)

    def _run_generic(self, rng_key, args, kwargs, extra_fields, init_params, lower, upper, resume):
        """mcmc.py:466-521 for a kernel that is driven one ``sample`` at a time from the host (HMCGibbs / HMCECS): the chains
        of the run are one vectorised kernel state; collection as fori_collect (util.py:368-403)."""
        keys = self._chain_keys(rng_key)
        C = self.num_chains
        k = self.sampler
        if resume is None:
            state = k.init(keys if C > 1 else keys[0], self.num_warmup, init_params, args, kwargs)
            i0 = 0
        else:
            state = resume._replace(rng_key=(keys if C > 1 else keys[0]))
            i0 = int(np.ravel(state.hmc_state.i)[0])
            lower, upper = i0, i0 + self.num_samples
        post = k.postprocess_fn(args, kwargs)
        S = max((upper - lower) // self.thinning, 0)
        start = lower + (upper - lower) % self.thinning
        rows, extra = [], {f: [] for f in extra_fields}

        def field(st, path):
            for part in path.split("."):
                st = st[part] if isinstance(st, dict) else getattr(st, part)
            return np.asarray(st)
        for i in range(i0, upper):
            state = k.sample(state, args, kwargs)
            if i >= start and (i - start + 1) % self.thinning == 0:
                rows.append({n: np.array(v) for n, v in post(state.z).items()})
                for f in extra_fields:
                    extra[f].append(field(state, f))
        lead = (lambda a: a) if C > 1 else (lambda a: a[None])
        stack = lambda seq: np.moveaxis(np.stack([lead(a) for a in seq]), 0, 1) if seq else np.zeros((C, 0))
        self._states = {"z": {n: stack([r[n] for r in rows]) for n in (rows[0] if rows else {})}}
        for f in extra_fields:
            self._states[f] = stack(extra[f])
        div = np.zeros((C, S), bool)
        self._states.setdefault("diverging", div)
        self._states_flat = None
        self._last_state = state
        self._bound = getattr(k, "_hmc_bound", None)
        self.total_grad_evals = None

    def _run(self, rng_key, args, kwargs, extra_fields, init_params, lower=None, upper=None, resume=None):
        if self._generic:
            return self._run_generic(rng_key, args, kwargs, extra_fields, init_params, lower, upper, resume)
        import time as _time
        t_begin = _time.perf_counter()
        keys = self._chain_keys(rng_key)
        fresh = resume is None
        same = lambda a, b: len(a) == len(b) and all(x is y for x, y in zip(a, b))
        same_data = bool(self._shards) and same(args, self._args) and sorted(kwargs) == sorted(self._kwargs) and \
            same([kwargs[k] for k in sorted(kwargs)], [self._kwargs[k] for k in sorted(kwargs)])
        if fresh or not same_data:
            self._ensure_engines(args, kwargs)
        self._args, self._kwargs = args, kwargs
        t_engines = _time.perf_counter()
        bound = self._bound
        D = self._shards[0].engine.D
        z0 = None
        if fresh:
            z0 = _flatten_init(init_params, bound, D) if init_params is not None else self.sampler._strategy_z0(bound, D, self.num_chains)
        else:
            i0 = np.unique(np.asarray(resume.i))
            if i0.size != 1:
                raise ValueError("post_warmup_state: every chain must be at the same iteration")
            lower, upper = int(i0[0]), int(i0[0]) + self.num_samples       # fori_collect(0, num_samples) from that state
        fields = ["z", "diverging"]
        unc_sites, remove = [], set()
        for f in tuple(extra_fields):
            if f.startswith("~z."):
                remove.add(f[3:])
            elif f.startswith("z."):
                unc_sites.append(f[2:])
            elif f in _STATE_FIELDS:
                if _STATE_FIELDS[f] not in fields:
                    fields.append(_STATE_FIELDS[f])
            else:
                raise ValueError(f"unsupported extra field {f!r}")
        n_local, lo0 = self._local_chains, self._local_lo

        def work(s: _Shard):
            e = s.engine
            with torch.cuda.device(s.device), torch.cuda.stream(getattr(s, "stream", None)):
                if fresh:
                    if self.sampler._inverse_mass_matrix is not None:
                        e.set_inverse_mass_matrix(self.sampler._inverse_mass_matrix)
                    e.init(keys[s.lo:s.hi], self.num_warmup, None if z0 is None else z0[s.lo:s.hi])
                else:
                    _state_to_engine(resume, bound, e, n_local, s.lo - lo0, s.hi - lo0, self.num_warmup,
                                     keys=keys[lo0:lo0 + n_local])
                out = e.run(upper, lower, self.thinning, fields=fields)
                con = e.constrain(out["z"]).view(e.C, -1, e.Dc)
                st, vec = e.state()                                  # raises "Cannot find valid initial parameters" (EINIT)
                host = {k: v.cpu().numpy() for k, v in out.items()}
                host["_constrained"] = con.cpu().numpy()
                return host, st, vec, (e.dense_state() if self.sampler._dense_mass else None)

        results = self._for_each_shard(work)
        t_work = _time.perf_counter()
        if self.row_shards > 1:
            results = results[:1]              # every rank holds the same (bit-identical) chains
        #: gradient evaluations (leapfrogs) spent so far by every local chain, warm-up included
        self.total_grad_evals = int(sum(int(r[1][k].total_leapfrogs) for r in results for k in range(len(r[1]))))
        host = {k: np.concatenate([r[0][k] for r in results], axis=0) for k in results[0][0]}
        if self._dist:
            host = self._all_gather(host)
        states: Dict[str, object] = {}
        zdict = {}
        for s in bound.sites:
            if s.name in remove:
                continue
            block = host["_constrained"][:, :, s.c_offset:s.c_offset + s.size]
            zdict[s.name] = block.reshape(block.shape[:2] + tuple(s.shape))
        states["z"] = zdict
        states["diverging"] = host["diverging"].astype(bool)
        for f in tuple(extra_fields):
            if f in _STATE_FIELDS:
                states[f] = host[_STATE_FIELDS[f]]
        for name in unc_sites:
            s = next(x for x in bound.latent_sites if x.name == name)
            block = host["z"][:, :, s.z_offset:s.z_offset + s.size]
            states["z." + name] = block.reshape(block.shape[:2] + tuple(s.shape))
        self._states = states
        self._states_flat = None
        self._last_state = _state_from_engine([r[1] for r in results], [r[2] for r in results], bound,
                                              self.num_chains == 1 and not self._dist, self.sampler._trajectory_length,
                                              dense=[r[3] for r in results] if self.sampler._dense_mass else None)
        #: wall-clock split of this call (ms): H2D of the data + engine creation (tile images), chain init + sampling +
        #: constrain + D2H of the samples, host-side assembly of the result dicts
        self.timings = {"h2d_and_engine_create": 1e3 * (t_engines - t_begin), "init_sample_d2h": 1e3 * (t_work - t_engines),
                        "assemble": 1e3 * (_time.perf_counter() - t_work)}

    def _all_gather(self, host):
        dist = torch.distributed
        out = {}
        on_gpu = dist.get_backend() == "nccl"
        for k, v in host.items():
            t = torch.from_numpy(np.ascontiguousarray(v))
            t = t.cuda() if on_gpu else t
            parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
            dist.all_gather(parts, t)                 # the only communication of the chain-sharded mode
            out[k] = torch.cat(parts, dim=0).cpu().numpy()
        return out

    def __getstate__(self):
        """mcmc.py:806-809 (test_pickle.py:88-95): samples, states and settings travel; engine handles and device buffers
        are rebuilt by the next ``run``."""
        state = self.__dict__.copy()
        state["_shards"] = []
        state["_args"], state["_kwargs"] = (), {}
        return state

    # ------------------------------------------------------------------ results
    @property
    def last_state(self):
        return self._last_state

    @property
    def post_warmup_state(self):
        return self._warmup_state

    @post_warmup_state.setter
    def post_warmup_state(self, state):
        """mcmc.py:558-587: continue sampling from a previous (post warm-up) state; ``run`` loads it into the engine."""
        self._warmup_state = state

    def get_samples(self, group_by_chain=False):
        """mcmc.py:549-556."""
        if self._states is None:
            raise RuntimeError("run() has not been called")
        z = self._states["z"]
        if group_by_chain:
            return dict(z)
        return {k: v.reshape((-1,) + v.shape[2:]) for k, v in z.items()}

    def get_extra_fields(self, group_by_chain=False):
        if self._states is None:
            raise RuntimeError("run() has not been called")
        out = {k: v for k, v in self._states.items() if k != "z"}
        if group_by_chain:
            return out
        return {k: v.reshape((-1,) + v.shape[2:]) for k, v in out.items()}

    def print_summary(self, prob=0.9, exclude_deterministic=True):
        """mcmc.py:769-797 (table of mean / std / n_eff / r_hat + number of divergences)."""
        det = {s.name for s in self._bound.sites if s.deterministic}
        sites = {k: v for k, v in self._states["z"].items() if not (exclude_deterministic and k in det)}
        summ = diagnostics.summary(sites, group_by_chain=True)
        print("\n{:>16} {:>9} {:>9} {:>9} {:>9}".format("", "mean", "std", "n_eff", "r_hat"))
        for name, s in summ.items():
            mean, std, neff, rhat = (np.atleast_1d(s[k]).ravel() for k in ("mean", "std", "n_eff", "r_hat"))
            for j in range(mean.shape[0]):
                label = name if mean.shape[0] == 1 else f"{name}[{j}]"
                print("{:>16} {:>9.2f} {:>9.2f} {:>9.2f} {:>9.2f}".format(label, mean[j], std[j], neff[j], rhat[j]))
        print("\nNumber of divergences: {}".format(int(np.sum(self._states["diverging"]))))
