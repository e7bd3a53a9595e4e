"""Rollout-buffer glue behind the simulator: generalised advantage estimation on the device (futbol_gae).

Mirrors what stable-baselines' PPO2 runner does with the reference env's outputs (colab_notebook.ipynb:852;
gamma 0.99 / lambda 0.95 in the saved models' JSON), on the ``[T, n]`` tensors the vectorised env produces,
without leaving the GPU.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def gae(reward, done, value, gamma=0.99, lam=0.95, out=None):
    """reward f32 [T, n], done u8 [T, n], value f32 [T + 1, n] (CUDA, contiguous) -> (advantage, return) f32 [T, n]."""
    T, n = reward.shape
    if reward.dtype != torch.float32 or value.dtype != torch.float32 or done.dtype != torch.uint8:
        raise ValueError("gae expects float32 reward/value and uint8 done")
    if tuple(done.shape) != (T, n) or tuple(value.shape) != (T + 1, n):
        raise ValueError("shape mismatch: reward %s done %s value %s" % (tuple(reward.shape), tuple(done.shape), tuple(value.shape)))
    if not (reward.is_cuda and done.is_cuda and value.is_cuda):
        raise _lib.FutbolError("gae needs CUDA tensors; there is no CPU fallback")
    reward, done, value = reward.contiguous(), done.contiguous(), value.contiguous()
    adv, ret = out if out is not None else (torch.empty_like(reward), torch.empty_like(reward))
    lib = _lib.load()
    with torch.cuda.device(reward.device):
        stream = C.c_void_p(torch.cuda.current_stream(reward.device).cuda_stream)
        _lib.check(lib.futbol_gae(C.c_void_p(reward.data_ptr()), C.c_void_p(done.data_ptr()), C.c_void_p(value.data_ptr()),
                                  float(gamma), float(lam), C.c_void_p(adv.data_ptr()), C.c_void_p(ret.data_ptr()), T, n, stream))
    return adv, ret
