"""Builds and wraps the TEST-ONLY host simulator (tests/hostsim/hostsim.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from numpyro_b200 import _capi

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
_SO = os.path.join(_DIR, "libhostsim.so")
_CSRC = os.path.join(os.path.dirname(_DIR), "..", "numpyro_b200", "csrc")

CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float))


def build():
    srcs = [os.path.join(_DIR, "hostsim.cpp")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) > os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                           srcs[0], "-o", _SO])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.hostsim_create.argtypes = [C.POINTER(_capi.Config), C.POINTER(C.c_void_p)]
        _lib.hostsim_destroy.argtypes = [C.c_void_p]
        _lib.hostsim_dim.argtypes = [C.c_void_p]
        _lib.hostsim_set_lookahead.argtypes = [C.c_void_p, C.c_int]
        _lib.hostsim_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.hostsim_run.argtypes = [C.c_void_p, C.POINTER(_capi.Run), CB, C.c_void_p]
        _lib.hostsim_get_state.argtypes = [C.c_void_p] * 6
        _lib.hostsim_potential.argtypes = [C.c_void_p] * 4
        _lib.hostsim_set_inverse_mass_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib.hostsim_get_dense_state.argtypes = [C.c_void_p] * 5
        _lib.hostsim_prng_split.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.hostsim_prng_uniform.argtypes = [C.c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_void_p]
        _lib.hostsim_prng_normal.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
        _lib.hostsim_detmath.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostSim:
    """Same call sequence as the C ABI (create / init / run / get_state) on host arrays."""

    def __init__(self, cfg, keep=()):
        self._keep = list(keep)             # host arrays the config points into
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = lib().hostsim_create(C.byref(cfg), C.byref(self.h))
        assert rc == 0, rc
        self.C = cfg.num_chains
        self.D = lib().hostsim_dim(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().hostsim_destroy(self.h)
            self.h = None

    def set_lookahead(self, on=True):
        """Run Tick::prefetch() (PRNG look-ahead) before every gradient, as the streaming engine does."""
        lib().hostsim_set_lookahead(self.h, 1 if on else 0)

    def init(self, keys, num_warmup, z0=None):
        keys = np.ascontiguousarray(keys, np.uint32).reshape(self.C, 2)
        z0 = None if z0 is None else np.ascontiguousarray(z0, np.float32)
        assert lib().hostsim_init(self.h, _p(keys), _p(z0), num_warmup) == 0

    def run(self, upper, lower, thinning=1, potential=None):
        S = (upper - lower) // thinning
        start = lower + (upper - lower) % thinning
        out = {"z": np.zeros((self.C, S, self.D), np.float32)}
        for f in ("diverging", "num_steps"):
            out[f] = np.zeros((self.C, S), np.int32)
        for f in ("accept_prob", "mean_accept_prob", "potential_energy", "energy", "step_size"):
            out[f] = np.zeros((self.C, S), np.float32)
        run = _capi.Run(upper=upper, collect_start=start, thinning=thinning, collection_size=S)
        for f, a in out.items():
            setattr(run, f, _p(a))
        if potential is None:
            cb = C.cast(None, CB)
        else:
            D = self.D

            def _cb(user, chain, z, d, u, g):
                zz = np.ctypeslib.as_array(z, (D,)).copy()
                uu, gg = potential(chain, zz)
                u[0] = np.float32(uu)
                np.ctypeslib.as_array(g, (D,))[:] = np.asarray(gg, np.float32)
            cb = CB(_cb)
        assert lib().hostsim_run(self.h, C.byref(run), cb, None) == 0
        return out

    def state(self):
        st = (_capi.ChainState * self.C)()
        z, g, imm, sm = (np.zeros((self.C, self.D), np.float32) for _ in range(4))
        lib().hostsim_get_state(self.h, C.cast(st, C.c_void_p), _p(z), _p(g), _p(imm), _p(sm))
        return st, z, g, imm, sm

    def set_inverse_mass_matrix(self, imm):
        imm = np.ascontiguousarray(imm, np.float32)
        assert lib().hostsim_set_inverse_mass_matrix(self.h, _p(imm), imm.ndim) == 0

    def dense_state(self):
        a = [np.zeros((self.C, self.D, self.D), np.float32) for _ in range(4)]
        assert lib().hostsim_get_dense_state(self.h, *[_p(x) for x in a]) == 0
        return dict(zip(("inverse_mass_matrix", "mass_matrix_sqrt", "mass_matrix_sqrt_inv", "wf_m2"), a))

    def potential(self, z):
        z = np.ascontiguousarray(z, np.float32).reshape(self.C, self.D)
        U = np.zeros(self.C, np.float32)
        g = np.zeros((self.C, self.D), np.float32)
        lib().hostsim_potential(self.h, _p(z), _p(U), _p(g))
        return U, g
