"""The kernel's guard-free fp64 division / square root (csrc/ieee_fast.cuh) against nvcc's correctly rounded
builtins, bit for bit, on operands drawn from the simulator's domain."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fast_div_and_sqrt_match_the_builtins():
    import torch
    from gym_futbol_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(0)
    n = 1 << 24
    mism = torch.zeros(3, dtype=torch.int64, device="cuda")

    def run(a, b):
        _lib.check(lib.futbol_selftest_arith(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(mism.data_ptr()),
                                             a.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    for rep in range(8):                          # 2^27 pairs
        # magnitudes log-uniform over 1e-17 .. 1e5 with random signs; numerators include exact zeros
        ea = torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * 22 - 17
        eb = torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * 22 - 17
        a = torch.pow(10.0, ea) * (torch.randint(0, 2, (n,), generator=g, device="cuda") * 2 - 1)
        b = torch.pow(10.0, eb) * (torch.randint(0, 2, (n,), generator=g, device="cuda") * 2 - 1)
        a[::97] = 0.0
        run(a, b)
        # pitch-scale values: differences of positions, squared distances, speeds
        p = torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * 105
        q = torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * 68
        run((p - 52.5) * 0.1, torch.sqrt((p - 52.5) ** 2 + (q - 34.0) ** 2) + 1e-12)
    # the constants the kernels divide by, against many numerators; perfect squares and their neighbours
    x = torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * 200
    for c in (0.1, 180.0, 2.0, 10.0, 52.5, 55.5, 34.0, 25.0):
        run(x, torch.full_like(x, c))
    k = torch.arange(1, n + 1, device="cuda", dtype=torch.float64)
    sq = (k % 4096) ** 2 + 1.0
    run(k, sq)
    run(k, torch.nextafter(sq, torch.zeros_like(sq)))
    run(k, torch.nextafter(sq, torch.full_like(sq, 1e30)))
    torch.cuda.synchronize()
    assert mism.tolist() == [0, 0, 0], mism.tolist()
