"""CPU-only checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/futbol_b200.h declares; structs match their Python mirrors; the product refuses to run
without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "futbol_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(futbol_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gym_futbol_b200 import _lib
    L = _lib.load()
    names = _declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(L, name), name
    assert set(names) == set(_lib.EXPORTS)
    assert L.futbol_abi_version() == 2


def test_struct_mirrors_match_header_sizes():
    from gym_futbol_b200 import _lib
    assert C.sizeof(_lib.FutbolConfig) == 64
    assert _lib.V0_ENV_STATE.itemsize == 224
    assert _lib.STATS_DTYPE.itemsize == C.sizeof(_lib.FutbolStats) == 64
    assert _lib.V0_ENV_STATE.fields["t_total"][1] == 200 and _lib.V0_ENV_STATE.fields["owner"][1] == 220
    assert _lib.V1_ENV_STATE.itemsize == 21 * 6 * 8 + 24 and _lib.V1_ENV_STATE.fields["owner_side"][1] == 21 * 6 * 8 + 16
    for N, P in ((1, 3 + 36), (2, 10 + 60), (5, 55 + 132), (10, 210 + 252)):      # header + jn[P] + last[P], padded to 8 bytes
        dt = _lib.v1_env_state_dtype(N)
        assert dt.itemsize == (1032 + 12 * P + 7) // 8 * 8 and dt.fields["jn"][1] == 1032 and dt.fields["last"][1] == 1032 + 8 * P


def test_null_arguments_return_error_codes_without_a_gpu():
    from gym_futbol_b200 import _lib
    L = _lib.load()
    assert L.futbol_create(None, None) == -1
    assert b"null" in L.futbol_last_error()
    assert L.futbol_reset(None, None, None, None, 0, None) == -1
    assert L.futbol_rollout(None, None, 4, None, None, None, None, None, None) == -1
    assert L.futbol_state_bytes(None) == 0
    assert L.futbol_set_rollout_slices(None, 4) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gym_futbol_b200 import FutbolError, FutbolVecEnv
    with pytest.raises(FutbolError):
        FutbolVecEnv(4)
    from gym_futbol_b200 import _lib
    cfg = _lib.FutbolConfig(_lib.ABI_VERSION, 0, 4, 0, 0, 2, 1, 0, 0, 1, 20, 40.0, 12.0)
    h = C.c_void_p()
    assert _lib.load().futbol_create(C.byref(cfg), C.byref(h)) == -2      # FUTBOL_ERR_CUDA
    assert b"no CPU fallback" in _lib.load().futbol_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gym_futbol_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|from\s+\.+\s*oracle|oracle/|libfutbol_oracle)", src), \
                    os.path.join(dirpath, f)


def test_builtin_spaces():
    from gym_futbol_b200 import spaces
    d = spaces.Discrete(16)
    assert d.contains(d.sample()) and not d.contains(16)
    m = spaces.MultiDiscrete([5, 5, 5, 5])
    assert m.contains(m.sample())
    b = spaces.Box(low=-1.0, high=1.0, shape=(20,), dtype=np.float32)
    assert b.shape == (20,) and b.contains(b.sample())
