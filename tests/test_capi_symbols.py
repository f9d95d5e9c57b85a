"""The C-ABI library loads on a CPU-only box and exports every symbol include/b200nuts.h declares;
without a GPU the product path fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200nuts.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200nuts_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from numpyro_b200 import _capi, build
    build.build_engine()
    lib = _capi.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200nuts.h but not exported"
    assert set(names) == set(_capi.EXPORTS), "ctypes table and header disagree"


def test_struct_sizes_match_header_layout():
    from numpyro_b200 import _capi
    assert C.sizeof(_capi.Run) == 4 * 4 + 8 * 8 + 8          # + max_passes (int32, padded to the struct alignment)
    assert C.sizeof(_capi.ChainState) % 8 == 0
    assert _capi.Config.X.offset % 8 == 0 and _capi.Config.nccl_comm.offset % 8 == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from numpyro_b200 import _capi, engine
    with pytest.raises(engine.EngineError, match="no CPU fallback"):
        engine.Engine(family=_capi.FAMILY_DIAG_GAUSSIAN, num_chains=1, n_rows=2, aux=[0.0, 0.0, 1.0, 1.0])
    h = C.c_void_p()
    cfg = _capi.default_config(family=_capi.FAMILY_DIAG_GAUSSIAN, num_chains=1, n_rows=2)
    assert _capi.load().b200nuts_create(C.byref(cfg), C.byref(h)) == _capi.ECUDA
    assert b"no CPU fallback" in _capi.load().b200nuts_last_error(None)
