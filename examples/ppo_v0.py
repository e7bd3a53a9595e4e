#!/usr/bin/env python
"""PPO on the v0 FutbolEnv (2v2 vs the hard-coded opponents) with the policy consuming the simulator's CUDA
tensors in place -- BASELINE.json configs[3]: 65,536 envs, n_steps 128, 4 minibatches x 4 epochs, gamma 0.99,
lambda 0.95, entropy 0.01, value 0.5, clip 0.2, max-grad-norm 0.5 (the hyper-parameters recorded in the
reference's saved PPO2 models, trained_model_2v2/model*.zip `data`), MLP shaped like the notebook's custom
policy ([256, 256] shared, [128, 128] policy / value heads, colab_notebook.ipynb:782-783).

    python examples/ppo_v0.py [--envs 65536] [--iters 3] [--steps 128] [--fused 0|1]

The policy needs obs_t to pick a_t, so the rollout uses the per-step API (one launch per step, state round-trips
HBM), replayed as one CUDA graph from the second iteration on.  With ``--fused 1`` (default) the step kernel writes
observation, reward and done flag of step t STRAIGHT into row t + 1 / t of the rollout buffers (``step(out=...)``:
no per-step copies), the action and its log-probability come straight from the logits in ONE launch
(``rollout_buffer.sample_actions``), and every minibatch is built by ONE gather launch over the six buffers
(``rollout_buffer.gather_minibatch``); ``--fused 0`` is the previous flow (copy the env's buffers every step, six torch
indexing launches per minibatch) kept for the before/after number.  GAE runs on the device (futbol_gae).  Prints
env-steps/s of the rollout alone and of rollout + update.  torch is the policy / optimiser library here; the
simulator, the advantage kernel and the gather are this repository's CUDA.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn as nn

from gym_futbol_b200 import FutbolVecEnv
from gym_futbol_b200.rollout_buffer import gae, gather_minibatch, sample_actions


class Policy(nn.Module):
    def __init__(self, obs_dim=30, n_actions=16):
        super().__init__()
        self.shared = nn.Sequential(nn.Linear(obs_dim, 256), nn.Tanh(), nn.Linear(256, 256), nn.Tanh())
        self.pi = nn.Sequential(nn.Linear(256, 128), nn.Tanh(), nn.Linear(128, 128), nn.Tanh(), nn.Linear(128, n_actions))
        self.vf = nn.Sequential(nn.Linear(256, 128), nn.Tanh(), nn.Linear(128, 128), nn.Tanh(), nn.Linear(128, 1))
        # observations are raw pitch coordinates (0..105); a fixed scale keeps the first layer in range
        self.register_buffer("scale", torch.tensor([105.0, 68.0, 105.0, 68.0, 20.0] * 5 + [10.0] * 5).reciprocal())

    def forward(self, obs):
        h = self.shared(obs * self.scale)
        return self.pi(h), self.vf(h).squeeze(-1)


class PPO:
    """Collection (T x (policy forward, sampling, env step) + GAE) and the clipped-surrogate update."""

    def __init__(self, n_envs=65536, n_steps=128, minibatches=4, epochs=4, seed=0, device="cuda:0", bf16=True, fused=True, graph=True,
                 sampler=None):
        self.dev = dev = torch.device(device)
        torch.manual_seed(seed)
        torch.backends.cuda.matmul.allow_tf32 = True      # the policy is library code; the simulator stays fp64
        self.n, self.T, self.minibatches, self.epochs, self.fused, self.use_graph = n_envs, n_steps, minibatches, epochs, bool(fused), bool(graph)
        # "kernel": actions and their log-probabilities straight from the logits in one launch (futbol_sample_actions, the fused
        # flow's default); "torch": log_softmax + multinomial + gather + cast
        self.sampler = sampler or ("kernel" if fused else "torch")
        self.seed = seed
        self.t_base = torch.zeros(1, dtype=torch.int64, device=dev)   # device counter of the sampler: advanced inside the graph
        n, T = n_envs, n_steps
        self.env = FutbolVecEnv(n, device=dev, seed=seed, random_opp=False)
        self.policy = Policy().to(dev)
        self.opt = torch.optim.Adam(self.policy.parameters(), lr=2.5e-4, eps=1e-5, fused=True)
        self.amp = lambda: torch.autocast("cuda", dtype=torch.bfloat16, enabled=bool(bf16))
        self.obs_buf = torch.empty((T + 1, n, 30), device=dev)        # row t: the observation action t was chosen from
        self.act_buf = torch.empty((T, n), dtype=torch.uint8, device=dev)
        self.logp_buf = torch.empty((T, n), device=dev)
        self.rew_buf = torch.empty((T, n), device=dev)
        self.done_buf = torch.empty((T, n), dtype=torch.uint8, device=dev)
        self.val_buf = torch.empty((T + 1, n), device=dev)
        self.adv = torch.empty((T, n), device=dev)
        self.ret = torch.empty((T, n), device=dev)
        obs = self.env.reset()
        assert obs.data_ptr() == self.env.obs.data_ptr()    # the policy reads the simulator's buffer in place
        self.obs_buf[0].copy_(obs)
        self.graph, self.iters, self.last_loss = None, 0, float("nan")

    def _collect(self):
        """All enqueued on the current stream, no host synchronisation (so it captures into a CUDA graph as is)."""
        env, policy, T = self.env, self.policy, self.T
        with torch.no_grad():
            if self.fused and self.iters >= 1:
                self.obs_buf[0].copy_(self.obs_buf[T])     # continue from the last observation of the previous collection
            obs = self.obs_buf[0] if self.fused else env.obs
            for t in range(T):
                if not self.fused:
                    self.obs_buf[t].copy_(obs)
                with self.amp():
                    logits, v = policy(obs)
                self.val_buf[t] = v.float()
                if self.sampler == "kernel":
                    sample_actions(logits, seed=self.seed, t=t, t_base=self.t_base, out=(self.act_buf[t], self.logp_buf[t]))
                else:
                    logp_all = torch.log_softmax(logits.float(), dim=-1)
                    a = torch.multinomial(logp_all.exp(), 1).squeeze(-1)
                    self.act_buf[t] = a.to(torch.uint8)
                    self.logp_buf[t] = logp_all.gather(1, a.unsqueeze(1)).squeeze(1)
                if self.fused:      # the kernel writes row t + 1 of the observation buffer and row t of reward / done itself
                    obs, _, _, _ = env.step(self.act_buf[t], out=(self.obs_buf[t + 1], self.rew_buf[t], self.done_buf[t]))
                    assert obs.data_ptr() == self.obs_buf[t + 1].data_ptr()
                else:
                    obs, rew, done, _ = env.step(self.act_buf[t])
                    assert obs.data_ptr() == env.obs.data_ptr()
                    self.rew_buf[t].copy_(rew)
                    self.done_buf[t].copy_(done)
            with self.amp():
                self.val_buf[T] = policy(obs)[1].float()
            self.t_base += T                                   # the next collection (or graph replay) draws fresh numbers
            gae(self.rew_buf, self.done_buf, self.val_buf, 0.99, 0.95, out=(self.adv, self.ret))

    def prepare_graph(self):
        """Capture the collection phase (~3000 launches) as ONE CUDA graph; call after at least one eager collection
        (warm-up of cuBLAS / the allocator).  Capturing enqueues nothing.  The optimiser updates the weights in place, so
        replays see the current policy; the C ABI only enqueues on the caller's stream and every buffer is persistent,
        so the whole phase captures as is."""
        if self.use_graph and self.graph is None and self.iters >= 1:
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._collect()
            torch.cuda.synchronize()

    def collect(self):
        """One collection phase: a graph replay once ``prepare_graph`` has run, eager before."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._collect()
        self.iters += 1

    def update(self):
        T, n, dev = self.T, self.n, self.dev
        flat = lambda x: x.reshape(T * n, *x.shape[2:])  # noqa: E731
        b_obs, b_act, b_logp = flat(self.obs_buf[:T]), flat(self.act_buf), flat(self.logp_buf)
        b_adv, b_ret, b_val = flat(self.adv), flat(self.ret), flat(self.val_buf[:T])
        mb = T * n // self.minibatches
        policy, opt = self.policy, self.opt
        for _ in range(self.epochs):
            perm = torch.randperm(T * n, device=dev)
            for k in range(self.minibatches):
                idx = perm[k * mb:(k + 1) * mb]
                if self.fused:
                    m_obs, m_act, m_logp, a_, m_ret, m_val = gather_minibatch(b_obs, idx, act=b_act, cols=(b_logp, b_adv, b_ret, b_val))
                else:
                    m_obs, m_act, m_logp, a_, m_ret, m_val = b_obs[idx], b_act[idx], b_logp[idx], b_adv[idx], b_ret[idx], b_val[idx]
                with self.amp():
                    logits, v = policy(m_obs)
                logits, v = logits.float(), v.float()
                dist = torch.distributions.Categorical(logits=logits)
                logp = dist.log_prob(m_act.long())
                a_ = (a_ - a_.mean()) / (a_.std() + 1e-8)
                ratio = (logp - m_logp).exp()
                pg = torch.max(-a_ * ratio, -a_ * ratio.clamp(0.8, 1.2)).mean()
                v_clip = m_val + (v - m_val).clamp(-0.2, 0.2)
                vl = 0.5 * torch.max((v - m_ret) ** 2, (v_clip - m_ret) ** 2).mean()
                loss = pg - 0.01 * dist.entropy().mean() + 0.5 * vl
                opt.zero_grad(set_to_none=True)
                loss.backward()
                nn.utils.clip_grad_norm_(policy.parameters(), 0.5)
                opt.step()
        self.last_loss = loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--minibatches", type=int, default=4)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--graph", type=int, default=1, help="replay the collection phase as one CUDA graph from the second iteration on")
    ap.add_argument("--bf16", type=int, default=1, help="run the torch policy under bf16 autocast (the simulator is fp64 either way)")
    ap.add_argument("--fused", type=int, default=1, help="step(out=...) into the rollout buffers + one-launch minibatch gather")
    args = ap.parse_args()
    ppo = PPO(args.envs, args.steps, args.minibatches, args.epochs, args.seed, bf16=args.bf16, fused=args.fused, graph=args.graph)
    n, T = args.envs, args.steps
    for it in range(args.iters):
        ppo.prepare_graph()                                # captures before the second iteration (not timed)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ppo.collect()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ppo.update()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("iter %d: mean step reward %.3f | rollout %.3e env-steps/s (%.1f ms) | rollout+update %.3e env-steps/s (%.1f ms) | loss %.3f"
              % (it, ppo.rew_buf.mean().item(), T * n / (t1 - t0), (t1 - t0) * 1e3, T * n / (t2 - t0), (t2 - t0) * 1e3, ppo.last_loss.item()), flush=True)


if __name__ == "__main__":
    main()
