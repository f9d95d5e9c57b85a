"""Thin Python handle over the C ABI (include/b200nuts.h).

PyTorch is used only as plumbing: it owns the device buffers (dataset, outputs) and the CUDA
stream; every computation happens inside libb200nuts.so.  There is no CPU path -- constructing an
``Engine`` without a GPU or without the built library raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _capi

_F32_FIELDS = ("accept_prob", "mean_accept_prob", "potential_energy", "energy", "step_size")
_I32_FIELDS = ("diverging", "num_steps")
ALL_FIELDS = ("z",) + _I32_FIELDS + _F32_FIELDS


class EngineError(RuntimeError):
    pass


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _on_device(fn):
    """Run a method with the handle's GPU current (the C ABI launches on the calling thread's current device) and
    restore the caller's device afterwards."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        with torch.cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapped


class Engine:
    """One handle = the chains that live on one GPU."""

    def __init__(self, device="cuda:0", X=None, y=None, aux=None, **cfg):
        if not torch.cuda.is_available():
            raise EngineError("numpyro_b200 needs a CUDA device: the engine has no CPU fallback")
        self.lib = _capi.load()
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        to = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float32).to(self.device).contiguous()
        self.X, self.y, self.aux = to(X), to(y), to(aux)            # borrowed by the handle: keep alive
        cond_fixed = cfg.pop("cond_fixed", None)
        c = _capi.default_config(**cfg)
        if cond_fixed is not None:                                  # HMCGibbs: host int32 mask over the full model's coordinates
            self._cond_fixed = np.ascontiguousarray(cond_fixed, np.int32)
            c.cond_fixed = self._cond_fixed.ctypes.data_as(C.c_void_p)
        if self.X is not None:
            c.X = self.X.data_ptr()
            c.n_rows, c.n_cols = self.X.shape
        if self.y is not None:
            c.y = self.y.data_ptr()
        if self.aux is not None:
            c.aux = self.aux.data_ptr()
        self.cfg = c
        self.h = C.c_void_p()
        with torch.cuda.device(self.device):      # (the caller's current device is left as it was)
            rc = self.lib.b200nuts_create(C.byref(c), C.byref(self.h))
        if rc != 0:
            raise EngineError(f"b200nuts_create failed ({rc}): {self.lib.b200nuts_last_error(None).decode()}")
        self.C = int(c.num_chains)
        self.D = self.lib.b200nuts_dim(self.h)
        self.Dfull = self.lib.b200nuts_full_dim(self.h)      # (== D unless the handle is conditioned on Gibbs sites)
        self.Dc = self.lib.b200nuts_constrained_dim(self.h)
        self.regime = self.lib.b200nuts_regime(self.h)
        self.shard_rank, self.shard_count = int(c.shard_rank), int(c.shard_count)
        self._shard_group = None

    # ------------------------------------------------------------------ row-sharded handles (config 5)
    @_on_device
    def shard_blob(self) -> bytes:
        """This rank's mailbox handle for :meth:`connect_shards` (b200nuts_shard_export)."""
        buf = C.create_string_buffer(_capi.SHARD_HANDLE_BYTES)
        self._check(self.lib.b200nuts_shard_export(self.h, buf), "b200nuts_shard_export")
        return buf.raw

    @_on_device
    def connect_shards(self, blobs=None, group=None):
        """Wire the per-gradient all-reduce of a row-sharded handle.  ``blobs``: the ranks' :meth:`shard_blob` in rank
        order (ranks = threads of this process), or None to exchange them over ``torch.distributed`` (ranks = processes;
        every later launch is then preceded by a barrier on ``group`` so that the ranks' kernels start together)."""
        if blobs is None:
            import torch.distributed as dist
            blobs = [None] * dist.get_world_size(group)
            dist.all_gather_object(blobs, self.shard_blob(), group=group)
            self._shard_group = group if group is not None else dist.group.WORLD
        if len(blobs) != self.shard_count:
            raise EngineError(f"connect_shards: {len(blobs)} handles for shard_count={self.shard_count}")
        self._check(self.lib.b200nuts_shard_connect(self.h, b"".join(blobs)), "b200nuts_shard_connect")

    def _align_ranks(self):
        if self._shard_group is not None:
            import torch.distributed as dist
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self._shard_group)

    # ------------------------------------------------------------------ lifetime
    @_on_device
    def close(self):
        if getattr(self, "h", None) and self.h.value:
            torch.cuda.synchronize(self.device)
            self.lib.b200nuts_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): {self.lib.b200nuts_last_error(self.h).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ MCMCKernel.init
    @_on_device
    def init(self, keys, num_warmup: int, z0=None):
        keys = np.ascontiguousarray(keys, np.uint32).reshape(self.C, 2)
        self._z0 = None if z0 is None else torch.as_tensor(z0, dtype=torch.float32).to(self.device).contiguous().view(self.C, self.D)
        self._check(self.lib.b200nuts_init(self.h, keys.ctypes.data_as(C.c_void_p), _ptr(self._z0), int(num_warmup),
                                           self._stream()), "b200nuts_init")
        self.num_warmup = int(num_warmup)

    # ------------------------------------------------------------------ fori_collect
    @_on_device
    def run(self, upper: int, lower: int, thinning: int = 1, fields: Sequence[str] = ALL_FIELDS, max_passes: int = 0,
            out: Optional[Dict[str, torch.Tensor]] = None, sync: bool = True) -> Dict[str, torch.Tensor]:
        """``max_passes`` (streaming regime with <= 8 chains, gemm regime): stop after that many passes; call again with the
        same window and ``out=`` the returned buffers to continue -- the chains resume exactly where they paused.
        ``b200nuts_run`` only enqueues; with ``sync=True`` (default) the outcome is collected before returning
        (:meth:`sync`), otherwise the caller collects it later."""
        S = max((upper - lower) // thinning, 0)
        start = lower + (upper - lower) % thinning
        reuse = out
        out = {} if out is None else out
        run = _capi.Run(upper=int(upper), collect_start=int(start), thinning=int(thinning), collection_size=int(S),
                        max_passes=int(max_passes))
        for f in fields:
            if reuse is not None:
                t = reuse[f]
            elif f == "z":
                t = torch.zeros((self.C, S, self.D), dtype=torch.float32, device=self.device)
            elif f in _I32_FIELDS:
                t = torch.zeros((self.C, S), dtype=torch.int32, device=self.device)
            elif f in _F32_FIELDS:
                t = torch.zeros((self.C, S), dtype=torch.float32, device=self.device)
            else:
                raise ValueError(f"unknown field {f!r}")
            out[f] = t
            setattr(run, f, t.data_ptr() if S > 0 else None)
        self._align_ranks()
        self._check(self.lib.b200nuts_run(self.h, C.byref(run), self._stream()), "b200nuts_run")
        if sync:
            self.sync()
        return out

    @_on_device
    def transition(self, n_iter: int = 1, sync: bool = True):
        """``MCMCKernel.sample`` granularity: advance every chain by ``n_iter`` transitions (b200nuts_transition)."""
        self._check(self.lib.b200nuts_transition(self.h, int(n_iter), self._stream()), "b200nuts_transition")
        if sync:
            self.sync()

    @_on_device
    def state_to_device(self):
        """Enqueue-only export of (z, z_grad, scalars[C, 8]) into device tensors (b200nuts_state_to_device)."""
        z = torch.empty((self.C, self.D), dtype=torch.float32, device=self.device)
        g = torch.empty_like(z)
        sc = torch.empty((self.C, 8), dtype=torch.float32, device=self.device)
        self._check(self.lib.b200nuts_state_to_device(self.h, _ptr(z), _ptr(g), _ptr(sc), self._stream()), "b200nuts_state_to_device")
        return z, g, sc

    def sync(self):
        """Wait for the last enqueued launch and raise if it failed (b200nuts_sync)."""
        self._check(self.lib.b200nuts_sync(self.h), "b200nuts_sync")

    def gemm_info(self) -> Dict[str, int]:
        out = np.zeros(8, np.int32)
        self._check(self.lib.b200nuts_gemm_info(self.h, out.ctypes.data_as(C.c_void_p)), "b200nuts_gemm_info")
        return dict(zip(("chain_tiles", "row_chunks", "k_blocks", "segments", "chunks_per_segment", "column_blocks",
                         "padded_columns", "device_while_graph"), (int(v) for v in out)))

    # ------------------------------------------------------------------ HMCState in / out
    @_on_device
    def state(self):
        st = (_capi.ChainState * self.C)()
        names = ("z", "z_grad", "inverse_mass_matrix", "mass_matrix_sqrt", "wf_mean", "wf_m2")
        vec = {n: np.zeros((self.C, self.D), np.float32) for n in names}
        rc = self.lib.b200nuts_get_state(self.h, C.cast(st, C.c_void_p), *[vec[n].ctypes.data_as(C.c_void_p) for n in names],
                                         self._stream())
        if rc == _capi.EINIT:
            raise RuntimeError("Cannot find valid initial parameters. Please check your model again.")
        self._check(rc, "b200nuts_get_state")
        return st, vec

    @_on_device
    def set_state(self, st, vec, num_warmup: int):
        arr = lambda n: None if vec.get(n) is None else np.ascontiguousarray(vec[n], np.float32).ctypes.data_as(C.c_void_p)
        keep = {n: np.ascontiguousarray(v, np.float32) for n, v in vec.items() if v is not None}
        ptr = lambda n: keep[n].ctypes.data_as(C.c_void_p) if n in keep else None
        self._check(self.lib.b200nuts_set_state(self.h, C.cast(st, C.c_void_p), ptr("z"), ptr("z_grad"),
                                                ptr("inverse_mass_matrix"), ptr("wf_mean"), ptr("wf_m2"),
                                                int(num_warmup), self._stream()), "b200nuts_set_state")
        self.num_warmup = int(num_warmup)

    # ------------------------------------------------------------------ mass matrix structure
    @_on_device
    def set_inverse_mass_matrix(self, imm):
        """The kernel's ``inverse_mass_matrix=`` argument ([D] or [D, D], the same for every chain); before :meth:`init`."""
        imm = np.ascontiguousarray(imm, np.float32)
        if imm.ndim not in (1, 2) or any(n != self.D for n in imm.shape):
            raise ValueError(f"inverse_mass_matrix must be [{self.D}] or [{self.D}, {self.D}]")
        self._check(self.lib.b200nuts_set_inverse_mass_matrix(self.h, imm.ctypes.data_as(C.c_void_p), imm.ndim, self._stream()),
                    "b200nuts_set_inverse_mass_matrix")

    @_on_device
    def dense_state(self) -> Dict[str, np.ndarray]:
        """dense_mass handles: HMCAdaptState's matrices, [C, D, D] each (b200nuts_get_dense_state)."""
        names = ("inverse_mass_matrix", "mass_matrix_sqrt", "mass_matrix_sqrt_inv", "wf_m2")
        out = {n: np.zeros((self.C, self.D, self.D), np.float32) for n in names}
        self._check(self.lib.b200nuts_get_dense_state(self.h, *[out[n].ctypes.data_as(C.c_void_p) for n in names], self._stream()),
                    "b200nuts_get_dense_state")
        return out

    @_on_device
    def set_dense_state(self, inverse_mass_matrix, wf_m2=None):
        imm = np.ascontiguousarray(inverse_mass_matrix, np.float32).reshape(self.C, self.D, self.D)
        m2 = None if wf_m2 is None else np.ascontiguousarray(wf_m2, np.float32).reshape(self.C, self.D, self.D)
        self._check(self.lib.b200nuts_set_dense_state(self.h, imm.ctypes.data_as(C.c_void_p),
                                                      None if m2 is None else m2.ctypes.data_as(C.c_void_p), self._stream()),
                    "b200nuts_set_dense_state")

    # ------------------------------------------------------------------ HMCECS inner potential
    @_on_device
    def ecs_set_proxy(self, ref, eta_ref, G, H, L0: float):
        """Device tensors (borrowed: the caller keeps them alive): reference point, X ref, gradient and Hessian of the
        full-data log-likelihood at it; L0 its value (b200nuts_ecs_set_proxy)."""
        self._ecs_keep = (ref, eta_ref, G, H)
        self._check(self.lib.b200nuts_ecs_set_proxy(self.h, _ptr(ref), _ptr(eta_ref), _ptr(G), _ptr(H), C.c_float(L0)),
                    "b200nuts_ecs_set_proxy")

    @_on_device
    def ecs_set_indices(self, idx):
        """The current subsample of every chain, int32 [C, m] (b200nuts_ecs_set_indices)."""
        idx = np.ascontiguousarray(idx, np.int32).reshape(self.C, -1)
        self._check(self.lib.b200nuts_ecs_set_indices(self.h, idx.ctypes.data_as(C.c_void_p), self._stream()), "b200nuts_ecs_set_indices")

    # ------------------------------------------------------------------ parity hooks
    @_on_device
    def potential_and_grad(self, z):
        z = torch.as_tensor(z, dtype=torch.float32).to(self.device).contiguous().view(self.C, self.D)
        U = torch.zeros(self.C, dtype=torch.float32, device=self.device)
        g = torch.zeros((self.C, self.D), dtype=torch.float32, device=self.device)
        self._align_ranks()
        self._check(self.lib.b200nuts_potential_and_grad(self.h, _ptr(z), _ptr(U), _ptr(g), self._stream()),
                    "b200nuts_potential_and_grad")
        return U, g

    @_on_device
    def leapfrog(self, eps, inv_mass, z, r, n_steps: int):
        dev = lambda a, shape: torch.as_tensor(a, dtype=torch.float32).to(self.device).contiguous().view(*shape).clone()
        eps, inv_mass = dev(eps, (self.C,)), dev(inv_mass, (self.C, self.D))
        z, r = dev(z, (self.C, self.D)), dev(r, (self.C, self.D))
        U = torch.zeros(self.C, dtype=torch.float32, device=self.device)
        g = torch.zeros((self.C, self.D), dtype=torch.float32, device=self.device)
        self._check(self.lib.b200nuts_leapfrog(self.h, _ptr(eps), _ptr(inv_mass), _ptr(z), _ptr(r), _ptr(U), _ptr(g),
                                               int(n_steps), self._stream()), "b200nuts_leapfrog")
        return z, r, U, g

    @_on_device
    def cond_set_values(self, values):
        """Conditioned handles: [C, Dfull + 1] unconstrained values of the Gibbs coordinates + the potential shift."""
        v = np.ascontiguousarray(values, np.float32).reshape(self.C, self.Dfull + 1)
        self._check(self.lib.b200nuts_cond_set_values(self.h, v.ctypes.data_as(C.c_void_p), self._stream()), "b200nuts_cond_set_values")

    @_on_device
    def constrain(self, z: torch.Tensor) -> torch.Tensor:
        z = z.contiguous().view(-1, self.Dfull)
        out = torch.empty((z.shape[0], self.Dc), dtype=torch.float32, device=self.device)
        if z.shape[0] == 0:
            return out
        self._check(self.lib.b200nuts_constrain(self.h, _ptr(z), z.shape[0], _ptr(out), self._stream()), "b200nuts_constrain")
        return out

    @_on_device
    def log_likelihood(self, z) -> torch.Tensor:
        """[n, D] unconstrained samples -> [n, n_obs] log_prob of the observed site per observation (b200nuts_log_likelihood)."""
        z = torch.as_tensor(z, dtype=torch.float32).to(self.device).contiguous().view(-1, self.D)
        out = torch.empty((z.shape[0], int(self.lib.b200nuts_obs_count(self.h))), dtype=torch.float32, device=self.device)
        self._check(self.lib.b200nuts_log_likelihood(self.h, _ptr(z), z.shape[0], _ptr(out), self._stream()), "b200nuts_log_likelihood")
        return out

    @_on_device
    def predict(self, z, keys) -> torch.Tensor:
        """[n, D] unconstrained samples + [n, 2] uint32 keys of the observed site -> [n, n_obs] draws (b200nuts_predict)."""
        z = torch.as_tensor(z, dtype=torch.float32).to(self.device).contiguous().view(-1, self.D)
        k = torch.from_numpy(np.ascontiguousarray(keys, np.uint32).reshape(-1, 2).view(np.int32)).to(self.device)
        if k.shape[0] != z.shape[0]:
            raise ValueError("one key per sample")
        out = torch.empty((z.shape[0], int(self.lib.b200nuts_obs_count(self.h))), dtype=torch.float32, device=self.device)
        self._check(self.lib.b200nuts_predict(self.h, _ptr(z), _ptr(k), z.shape[0], _ptr(out), self._stream()), "b200nuts_predict")
        return out

    @property
    def launch_count(self) -> int:
        return int(self.lib.b200nuts_launch_count(self.h))

    @property
    def pass_count(self) -> int:
        return int(self.lib.b200nuts_pass_count(self.h))

    def debug_clocks(self) -> np.ndarray:
        out = np.zeros(16, np.uint64)
        self._check(self.lib.b200nuts_debug_clocks(self.h, out.ctypes.data_as(C.c_void_p)), "b200nuts_debug_clocks")
        return out


# ---------------------------------------------------------------------- PRNG / det-math hooks
def prng_split(keys, num: int) -> np.ndarray:
    keys = np.ascontiguousarray(keys, np.uint32).reshape(-1, 2)
    out = np.zeros((keys.shape[0], num, 2), np.uint32)
    rc = _capi.load().b200nuts_prng_split(keys.ctypes.data_as(C.c_void_p), keys.shape[0], num, out.ctypes.data_as(C.c_void_p))
    if rc:
        raise EngineError(f"b200nuts_prng_split failed ({rc})")
    return out


def _draw(fn, key, n, *extra, dtype=np.float32):
    key = np.ascontiguousarray(key, np.uint32).reshape(2)
    out = np.zeros(n, dtype)
    rc = fn(key.ctypes.data_as(C.c_void_p), n, *extra, out.ctypes.data_as(C.c_void_p))
    if rc:
        raise EngineError(f"prng hook failed ({rc})")
    return out


def prng_bits(key, n):
    return _draw(_capi.load().b200nuts_prng_bits, key, n, dtype=np.uint32)


def prng_uniform(key, n, lo=0.0, hi=1.0):
    return _draw(_capi.load().b200nuts_prng_uniform, key, n, C.c_float(lo), C.c_float(hi))


def prng_normal(key, n):
    return _draw(_capi.load().b200nuts_prng_normal, key, n)


def detmath(op: int, x) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32).ravel()
    out = np.zeros_like(x)
    rc = _capi.load().b200nuts_detmath(op, x.ctypes.data_as(C.c_void_p), x.shape[0], out.ctypes.data_as(C.c_void_p))
    if rc:
        raise EngineError(f"b200nuts_detmath failed ({rc})")
    return out
