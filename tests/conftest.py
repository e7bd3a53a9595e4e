import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden_v0():
    import json

    import numpy as np

    z = np.load(os.path.join(ROOT, "tests", "golden", "v0_golden.npz"))
    cases = {}
    for key in z.files:
        case, field = key.split("/")
        cases.setdefault(case, {})[field] = z[key]
    for case in cases.values():
        case["meta"] = json.loads(str(case["meta"]))
    fp = cases.pop("libm_fingerprint")
    from oracle.v0 import lib

    sq = lib().futbol_oracle_libm_sq
    import ctypes

    sq.restype, sq.argtypes = ctypes.c_double, [ctypes.c_double]
    same_libm = all(sq(float(x)) == float(y) for x, y in zip(fp["x"], fp["y"]))
    return {"cases": cases, "same_libm": same_libm}
