"""Host-buffer front end of the fused rollout: what a HOST-side consumer of the reference API receives.

The reference's callers live on the host (stable-baselines' runner, colab_notebook.ipynb:852): they hand numpy
actions to ``env.step`` and get numpy observations back.  ``HostRollout`` is that contract for the batched
simulator: actions come from pinned host memory and EVERY observation, reward and done flag of a K-step
rollout is delivered into pinned host memory.  The K steps are issued as ``chunks`` launches so that the
device->host copy of one chunk (copy stream) overlaps the simulation of the next (compute stream); two device
buffers alternate.  The link, not the simulator, bounds this path (120 B of observation per env-step against
~55 GB/s of PCIe), which is why the zero-copy device path (``FutbolVecEnv.rollout`` / ``step``) is the intended
use; ``bench.py`` reports both.

``bind_to_local_cpus`` pins the calling process to the CPUs next to its GPU before the pinned buffers are
allocated (first touch puts them on the GPU's NUMA node); ``measure_d2h_peak`` times plain pinned device->host
copies in this process -- with every rank of a multi-GPU job calling it at the same moment it measures the
box's shared host-link ceiling that the end-to-end number is then reported against.
"""
from __future__ import annotations

import os

import torch


def gpu_local_cpus(device_index):
    """CPUs on the NUMA node of the GPU (from sysfs), or None when the platform does not say."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return sorted(cpus) or None
    except Exception:  # noqa: BLE001
        return None


def bind_to_local_cpus(device_index):
    """Restrict this process to the GPU's local CPUs; returns the CPU list used (None = left unchanged)."""
    cpus = gpu_local_cpus(device_index)
    if cpus:
        try:
            os.sched_setaffinity(0, cpus)
        except Exception:  # noqa: BLE001
            return None
    return cpus


def measure_d2h_peak(device, nbytes=1 << 30, reps=3, barrier=None):
    """Best-of-``reps`` bandwidth (GB/s) of a pinned device->host copy of ``nbytes`` on ``device``.  ``barrier``: optional
    callable run before every repetition (multi-rank: all ranks copy at once -> the shared-link ceiling)."""
    dev = torch.device(device)
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    stream = torch.cuda.current_stream(dev)
    dst.copy_(src, non_blocking=True)
    stream.synchronize()
    best = 0.0
    for _ in range(reps):
        if barrier is not None:
            barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        dst.copy_(src, non_blocking=True)
        b.record(stream)
        stream.synchronize()
        best = max(best, nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
    return best


class HostRollout:
    """K-step rollouts of a ``FutbolVecEnv`` / ``FutbolV1VecEnv`` with host-resident actions and results."""

    def __init__(self, env, K, chunks=8):
        if K % chunks:
            raise ValueError("K must be a multiple of chunks")
        self.env, self.K, self.chunks, self.Kc = env, int(K), int(chunks), int(K) // int(chunks)
        dev, n, D = env.device, env.num_envs, env.obs_dim
        act_shape = (n,) + tuple(env.act_shape)
        self.h_actions = torch.zeros((K,) + act_shape, dtype=torch.uint8).pin_memory()
        self.h_obs = torch.empty((K, n, D), dtype=torch.float32).pin_memory()
        self.h_reward = torch.empty((K, n), dtype=torch.float32).pin_memory()
        self.h_done = torch.empty((K, n), dtype=torch.uint8).pin_memory()
        self.h_stats = torch.empty(env.stats.numel(), dtype=torch.uint8).pin_memory()
        self._d = [(torch.empty((self.Kc,) + act_shape, dtype=torch.uint8, device=dev),
                    torch.empty((self.Kc, n, D), dtype=torch.float32, device=dev),
                    torch.empty((self.Kc, n), dtype=torch.float32, device=dev),
                    torch.empty((self.Kc, n), dtype=torch.uint8, device=dev)) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(dev)
        self._simulated = [torch.cuda.Event() for _ in range(2)]     # chunk simulated (compute stream)
        self._copied = [torch.cuda.Event() for _ in range(2)]        # chunk copied out (copy stream)
        for ev in self._copied:
            ev.record(self.copy_stream)
        self.h2d_bytes = self.h_actions.numel()
        self.d2h_bytes = self.h_obs.numel() * 4 + self.h_reward.numel() * 4 + self.h_done.numel() + self.h_stats.numel()

    def run(self, synchronize=True):
        """One K-step rollout from ``self.h_actions``; results land in ``h_obs`` / ``h_reward`` / ``h_done`` / ``h_stats``.
        With ``synchronize`` (default) the host holds the whole rollout on return."""
        env, Kc = self.env, self.Kc
        stream = torch.cuda.current_stream(env.device)
        for c in range(self.chunks):
            b = c & 1
            da, do, dr, dd = self._d[b]
            ks = slice(c * Kc, (c + 1) * Kc)
            stream.wait_event(self._copied[b])
            da.copy_(self.h_actions[ks], non_blocking=True)
            env.rollout(Kc, actions=da, out=(do, dr, dd))
            self._simulated[b].record(stream)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self._simulated[b])
                self.h_obs[ks].copy_(do, non_blocking=True)
                self.h_reward[ks].copy_(dr, non_blocking=True)
                self.h_done[ks].copy_(dd, non_blocking=True)
                self._copied[b].record(self.copy_stream)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._simulated[(self.chunks - 1) & 1])
            self.h_stats.copy_(env.stats, non_blocking=True)
        if synchronize:
            self.copy_stream.synchronize()
        return self.h_obs, self.h_reward, self.h_done

    def run_resident(self):
        """The same loop with the results left in HBM for an on-device consumer: host actions in, statistics out.  The
        upload of chunk c + 1 (copy stream) overlaps the simulation of chunk c; two device action buffers alternate."""
        env, Kc = self.env, self.Kc
        stream = torch.cuda.current_stream(env.device)
        if not hasattr(self, "_uploaded"):
            self._uploaded = [torch.cuda.Event() for _ in range(2)]      # chunk's actions are in HBM (copy stream)
            self._consumed = [torch.cuda.Event() for _ in range(2)]      # the rollout that read them is done (compute stream)
        for ev in self._consumed:
            ev.record(stream)                      # also orders this call behind whatever used the buffers before it
        for c in range(self.chunks):
            b = c & 1
            da = self._d[b][0]
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self._consumed[b])
                da.copy_(self.h_actions[c * Kc:(c + 1) * Kc], non_blocking=True)
                self._uploaded[b].record(self.copy_stream)
            stream.wait_event(self._uploaded[b])
            env.rollout(Kc, actions=da)
            self._consumed[b].record(stream)
        self.h_stats.copy_(env.stats, non_blocking=True)
        stream.synchronize()
