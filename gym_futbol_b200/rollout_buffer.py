"""Rollout-buffer glue around the simulator: action sampling from the policy's logits (futbol_sample_actions), generalised
advantage estimation (futbol_gae) and the minibatch gather (futbol_gather_minibatch) on the device.

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


def sample_actions(logits, seed=0, t=0, t_base=None, out=None):
    """One categorical draw per row of ``logits`` ([n, A] float32 or bfloat16 CUDA tensor, A <= 32) in ONE launch:
    returns ``(actions uint8 [n], log_prob float32 [n])`` -- what softmax + multinomial + gather + cast compute.  The
    uniform of row i is Philox(seed, t_base[0] + t, i): ``t_base`` is an optional one-element int64 CUDA tensor, a counter
    the caller advances (``t_base += T``) so that a captured CUDA graph draws fresh numbers on every replay."""
    if not logits.is_cuda:
        raise _lib.FutbolError("sample_actions needs CUDA tensors; there is no CPU fallback")
    if logits.dim() != 2 or logits.dtype not in (torch.float32, torch.bfloat16) or not 1 <= logits.shape[1] <= 32:
        raise ValueError("logits must be [n, A] float32 or bfloat16 with A <= 32")
    if t_base is not None and (t_base.dtype != torch.int64 or t_base.numel() != 1 or t_base.device != logits.device):
        raise ValueError("t_base must be a one-element int64 tensor on the logits' device")
    logits = logits.contiguous()
    n, A = logits.shape
    act, logp = out if out is not None else (torch.empty(n, dtype=torch.uint8, device=logits.device),
                                             torch.empty(n, dtype=torch.float32, device=logits.device))
    if act.dtype != torch.uint8 or logp.dtype != torch.float32 or act.numel() != n or logp.numel() != n or not (
            act.is_contiguous() and logp.is_contiguous()) or act.device != logits.device or logp.device != logits.device:
        raise ValueError("out must be (uint8 [n], float32 [n]), contiguous, on the logits' device")
    lib = _lib.load()
    with torch.cuda.device(logits.device):
        stream = C.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream)
        _lib.check(lib.futbol_sample_actions(C.c_void_p(logits.data_ptr()), 1 if logits.dtype == torch.bfloat16 else 0, n, A,
                                             int(seed) & (2 ** 64 - 1), None if t_base is None else C.c_void_p(t_base.data_ptr()),
                                             int(t), C.c_void_p(act.data_ptr()), C.c_void_p(logp.data_ptr()), stream))
    return act, logp


def gather_minibatch(obs, idx, act=None, cols=(), out=None):
    """Rows ``idx`` (int64 CUDA tensor, values in [0, rows)) of a flattened rollout buffer in ONE launch.

    obs: float32 ``[..., D]`` (flattened to ``[rows, D]``; e.g. the ``[T, n, 30]`` observation buffer), act: optional uint8
    ``[...]`` with the same leading shape, cols: up to four float32 tensors of that leading shape (old log-prob, advantage,
    return, value).  Returns ``obs_mb`` if only obs is given, else ``(obs_mb, act_mb, *cols_mb)``.  ``out``: optional
    destination(s) of the same structure (contiguous, on the same device).  Equals torch indexing of each tensor.
    """
    if not obs.is_cuda:
        raise _lib.FutbolError("gather_minibatch needs CUDA tensors; there is no CPU fallback")
    if obs.dtype != torch.float32 or idx.dtype != torch.int64 or not idx.is_cuda:
        raise ValueError("gather_minibatch expects float32 observations and an int64 CUDA index")
    if len(cols) > 4:
        raise ValueError("at most four scalar columns")
    D = obs.shape[-1]
    obs2 = obs.contiguous().view(-1, D)
    rows, m = obs2.shape[0], idx.numel()
    idx = idx.contiguous().view(-1)
    srcs = [None if act is None else act.contiguous().view(-1)] + [c.contiguous().view(-1) for c in cols]
    for t, dt in zip(srcs, [torch.uint8] + [torch.float32] * len(cols)):
        if t is not None and (t.dtype != dt or t.numel() != rows or t.device != obs.device):
            raise ValueError("columns must have the observation buffer's leading shape (uint8 actions, float32 scalars)")
    if out is None:
        outs = [torch.empty((m, D), dtype=torch.float32, device=obs.device)] + \
               [None if t is None else torch.empty(m, dtype=t.dtype, device=obs.device) for t in srcs]
    else:
        outs = list(out) if isinstance(out, (tuple, list)) else [out]
        outs += [None] * (1 + len(srcs) - len(outs))
        for o, ref in zip(outs, [obs2] + srcs):
            if (o is None) != (ref is None) or (o is not None and (o.dtype != ref.dtype or o.numel() != m * (D if ref is obs2 else 1)
                                                                 or not o.is_contiguous() or o.device != obs.device)):
                raise ValueError("out must mirror the inputs: contiguous tensors of %d rows on %s" % (m, obs.device))
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
    pairs = []
    for src, dst in zip(srcs + [None] * (5 - len(srcs)), outs[1:] + [None] * (5 - len(srcs))):
        pairs += [p(src), p(dst)]
    lib = _lib.load()
    with torch.cuda.device(obs.device):
        stream = C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)
        _lib.check(lib.futbol_gather_minibatch(p(idx), m, rows, p(obs2), D, p(outs[0]), *pairs, None, stream))
    if act is None and not cols:
        return outs[0]
    return tuple(o for o in outs if o is not None)
