"""The C-ABI shared library: loads, exports every symbol include/pdabm.h declares, agrees with the
ctypes mirror on the struct layout, and fails with an error code (not a crash, not a CPU fallback)
when no CUDA device is present.  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from pdabm_b200 import _lib
    return _lib.load()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "pdabm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pd_[a-z_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    from pdabm_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 15
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"libpdabm.so does not export {n}"
    assert set(_lib.EXPORTS) == set(names), "ctypes binding and header disagree"


def test_abi_version_and_struct_layout(lib):
    from pdabm_b200._lib import pd_config
    assert lib.pd_abi_version() == 1
    cfg = pd_config()
    assert lib.pd_default_config(C.byref(cfg), 512) == 0
    assert cfg.struct_size == C.sizeof(pd_config)          # C sizeof == ctypes sizeof
    # reference defaults (model.py:39-169)
    assert (cfg.width, cfg.rows, cfg.row0) == (512, 512, 0)
    assert cfg.max_agents == 131072 and cfg.seed == 12345
    assert (cfg.payoff_cc, cfg.payoff_dc, cfg.payoff_cd, cfg.payoff_dd) == (3.0, 5.0, -1.0, 0.0)
    assert (cfg.cost_of_living, cfg.travel_cost, cfg.max_energy) == (1.0, 0.5, 150.0)
    assert (cfg.reproduce_min_energy, cfg.reproduce_cost, cfg.reproduction_inheritence) == (100.0, 50.0, 0.0)
    assert (cfg.init_energy_mu, cfg.init_energy_sigma, cfg.init_energy_min) == (50.0, 10.0, 5.0)
    assert list(cfg.strategy_weights) == [0.25] * 4
    assert (cfg.strategy_pure, cfg.strategy_per_trait, cfg.max_children_per_step) == (0, 0, 1)


def test_error_conventions(lib):
    from pdabm_b200._lib import pd_config
    cfg = pd_config()
    lib.pd_default_config(C.byref(cfg), 512)
    h = C.c_void_p()
    cfg.struct_size = 3
    assert lib.pd_create(C.byref(cfg), C.byref(h)) == -1           # PD_ERR_INVALID
    assert b"struct_size" in lib.pd_last_error()
    lib.pd_default_config(C.byref(cfg), 500)                        # not a power of two
    assert lib.pd_create(C.byref(cfg), C.byref(h)) == -1
    assert b"power of two" in lib.pd_last_error()
    assert lib.pd_step(None, 1, None, 0) == -3                      # PD_ERR_STATE
    assert lib.pd_destroy(None) == 0


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pdabm_b200._lib import pd_config
    from pdabm_b200 import Simulation, SimConfig, PdabmError
    cfg = pd_config()
    lib.pd_default_config(C.byref(cfg), 64)
    h = C.c_void_p()
    rc = lib.pd_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and not h.value                                  # PD_ERR_CUDA, no handle
    with pytest.raises(PdabmError):
        Simulation(SimConfig(width=64))
