"""CPU: the C-ABI library loads and exports every symbol include/bayesvg_b200.h declares; without a
device it fails loudly (no CPU fallback).  No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bayesvg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bayesvg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bvg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 25
    L = _lib.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names
    assert L.bvg_version() == _lib.ABI_VERSION


def test_config_default_matches_reference_defaults():
    """findSpatiallyVariableFeaturesBayes.R:78-97 and CmdStan's ADVI defaults"""
    cfg = _lib.Config()
    _lib.lib().bvg_config_default(C.byref(cfg))
    assert (cfg.n_iter, cfg.mle_init, cfg.mle_iter, cfg.n_draws, cfg.elbo_samples) == (3000, 1, 1000, 1000, 150)
    assert (cfg.eval_elbo, cfg.adapt_iter, cfg.lbfgs_history, cfg.seed) == (100, 50, 25, 312)
    assert cfg.tol_rel_obj == 0.01 and cfg.eta == 0.0 and cfg.world_size == 1
    assert C.sizeof(_lib.Config) == 136  # layout of bvg_config (guards accidental ABI drift)


def test_no_device_is_an_error_not_a_fallback():
    L = _lib.lib()
    if L.bvg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    cfg = _lib.Config()
    L.bvg_config_default(C.byref(cfg))
    h = C.c_void_p()
    rc = L.bvg_fit_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and b"no CPU fallback" in L.bvg_last_error()
    out = np.zeros((4, 2), order="F")
    xy = np.asfortranarray(np.random.default_rng(0).normal(size=(4, 2)))
    dp = C.POINTER(C.c_double)
    rc = L.bvg_basis_build(0, xy.ctypes.data_as(dp), 4, xy.ctypes.data_as(dp), 2, 1.0, 0, 1.5, 1.0, out.ctypes.data_as(dp))
    assert rc == -2


def test_argument_validation_happens_before_any_device_work():
    L = _lib.lib()
    cfg = _lib.Config()
    L.bvg_config_default(C.byref(cfg))
    h = C.c_void_p()
    cfg.abi_version = 99
    assert L.bvg_fit_create(C.byref(cfg), C.byref(h)) == -1 and b"ABI" in L.bvg_last_error()
    L.bvg_config_default(C.byref(cfg))
    cfg.likelihood = 7
    assert L.bvg_fit_create(C.byref(cfg), C.byref(h)) == -1
    assert L.bvg_last_error() == b"Please provide a valid likelihood."
    L.bvg_config_default(C.byref(cfg))
    cfg.precision, cfg.engine = _lib.PREC_FP64, _lib.ENGINE_TCGEN05
    assert L.bvg_fit_create(C.byref(cfg), C.byref(h)) == -1 and b"fp64" in L.bvg_last_error()
    dp = C.POINTER(C.c_double)
    z = np.zeros(8)
    assert L.bvg_basis_build(0, z.ctypes.data_as(dp), 4, z.ctypes.data_as(dp), 2, 1.0, 1, 2.0, 1.0, z.ctypes.data_as(dp)) == -1
    assert b"smoothness" in L.bvg_last_error()
    assert L.bvg_basis_build(0, z.ctypes.data_as(dp), 4, z.ctypes.data_as(dp), 2, 1.0, 5, 1.5, 1.0, z.ctypes.data_as(dp)) == -1
    assert L.bvg_last_error() == b"Please provide a valid covariance kernel."
