#!/usr/bin/env python
"""PPO on the v0 FutbolEnv (2v2 vs the hard-coded opponents) with the policy consuming the simulator's CUDA
tensors in place -- BASELINE.json configs[3]: 65,536 envs, n_steps 128, 4 minibatches x 4 epochs, gamma 0.99,
lambda 0.95, entropy 0.01, value 0.5, clip 0.2, max-grad-norm 0.5 (the hyper-parameters recorded in the
reference's saved PPO2 models, trained_model_2v2/model*.zip `data`), MLP shaped like the notebook's custom
policy ([256, 256] shared, [128, 128] policy / value heads, colab_notebook.ipynb:782-783).

    python examples/ppo_v0.py [--envs 65536] [--iters 3] [--steps 128]

The policy needs obs_t to pick a_t, so the rollout uses the per-step API (one launch per step, state
round-trips HBM), replayed as one CUDA graph from the second iteration on; observations, rewards and dones never leave the device and the env's own output buffers
are what the policy reads (asserted by data_ptr identity).  GAE runs on the device (futbol_gae).  Prints
env-steps/s of the rollout alone and of rollout + update.  torch is the policy/optimiser library here; the
simulator and the advantage kernel are this repository's CUDA.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn as nn

from gym_futbol_b200 import FutbolVecEnv
from gym_futbol_b200.rollout_buffer import gae


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
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(args.seed)
    torch.backends.cuda.matmul.allow_tf32 = True          # the policy is library code; the simulator stays fp64
    n, T = args.envs, args.steps
    env = FutbolVecEnv(n, device=dev, seed=args.seed, random_opp=False)
    policy = Policy().to(dev)
    opt = torch.optim.Adam(policy.parameters(), lr=2.5e-4, eps=1e-5, fused=True)
    amp = lambda: torch.autocast("cuda", dtype=torch.bfloat16, enabled=bool(args.bf16))  # noqa: E731
    obs_buf = torch.empty((T, n, 30), device=dev)
    act_buf = torch.empty((T, n), dtype=torch.uint8, device=dev)
    logp_buf = torch.empty((T, n), device=dev)
    rew_buf = torch.empty((T, n), device=dev)
    done_buf = torch.empty((T, n), dtype=torch.uint8, device=dev)
    val_buf = torch.empty((T + 1, n), device=dev)
    obs = env.reset()
    assert obs.data_ptr() == env.obs.data_ptr()            # the policy reads the simulator's buffer in place
    def collect():
        """T policy steps + T env steps + GAE, all enqueued on the current stream (no host synchronisation)."""
        obs = env.obs
        with torch.no_grad():
            for t in range(T):
                obs_buf[t].copy_(obs)
                with amp():
                    logits, v = policy(obs)
                val_buf[t] = v.float()
                logp_all = torch.log_softmax(logits.float(), dim=-1)
                a = torch.multinomial(logp_all.exp(), 1).squeeze(-1)
                act_buf[t] = a.to(torch.uint8)
                logp_buf[t] = logp_all.gather(1, a.unsqueeze(1)).squeeze(1)
                obs, rew, done, _ = env.step(act_buf[t])
                assert obs.data_ptr() == env.obs.data_ptr()
                rew_buf[t].copy_(rew)
                done_buf[t].copy_(done)
            with amp():
                val_buf[T] = policy(obs)[1].float()
            return gae(rew_buf, done_buf, val_buf, 0.99, 0.95)

    graph = None
    for it in range(args.iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if args.graph and it == 1:
            # The whole collection phase (T x (policy forward, sampling, env step) + GAE: ~3000 launches) as ONE CUDA
            # graph: the C ABI only enqueues on the caller's stream and every buffer is persistent, so it captures as is.
            # Captured after one eager iteration (warm-up of cuBLAS / the allocator); the optimiser updates the weights
            # in place, so replays see the current policy.  The capture itself is not timed.
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                adv, ret = collect()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        if graph is not None:
            graph.replay()
        else:
            adv, ret = collect()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        flat = lambda x: x.reshape(T * n, *x.shape[2:])
        b_obs, b_act, b_logp, b_adv, b_ret, b_val = flat(obs_buf), flat(act_buf).long(), flat(logp_buf), flat(adv), flat(ret), flat(val_buf[:T])
        mb = T * n // args.minibatches
        for _ in range(args.epochs):
            perm = torch.randperm(T * n, device=dev)
            for k in range(args.minibatches):
                idx = perm[k * mb:(k + 1) * mb]
                with amp():
                    logits, v = policy(b_obs[idx])
                logits, v = logits.float(), v.float()
                dist = torch.distributions.Categorical(logits=logits)
                logp = dist.log_prob(b_act[idx])
                a_ = b_adv[idx]
                a_ = (a_ - a_.mean()) / (a_.std() + 1e-8)
                ratio = (logp - b_logp[idx]).exp()
                pg = torch.max(-a_ * ratio, -a_ * ratio.clamp(0.8, 1.2)).mean()
                v_clip = b_val[idx] + (v - b_val[idx]).clamp(-0.2, 0.2)
                vl = 0.5 * torch.max((v - b_ret[idx]) ** 2, (v_clip - b_ret[idx]) ** 2).mean()
                loss = pg - 0.01 * dist.entropy().mean() + 0.5 * vl
                opt.zero_grad(set_to_none=True)
                loss.backward()
                nn.utils.clip_grad_norm_(policy.parameters(), 0.5)
                opt.step()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("iter %d: mean step reward %.3f | rollout %.3e env-steps/s (%.1f ms) | rollout+update %.3e env-steps/s (%.1f ms) | loss %.3f"
              % (it, rew_buf.mean().item(), T * n / (t1 - t0), (t1 - t0) * 1e3, T * n / (t2 - t0), (t2 - t0) * 1e3, loss.item()), flush=True)


if __name__ == "__main__":
    main()
