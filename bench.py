#!/usr/bin/env python
"""Headline benchmark: env-steps/s of the v0 FutbolEnv step, 2v2 vs hard-coded opponents, 2^20 envs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One bench "step" = one fused rollout launch: every env of the job advanced ROLLOUT_K = 64 env-steps
(BASELINE.json configs[2]).  The 2^20 envs of the metric are sharded over the N ranks by global env
id (no data-path collective), so total work is fixed as N grows: "scaling": "strong".

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TOTAL_ENVS = 1 << 20
ROLLOUT_K = 64
E2E_CHUNKS = 8                                 # launches per step in the end-to-end loop (copy/compute overlap)
STATE_BYTES_PER_ENV = 223                      # SoA state, v0_kernels.cu
BYTES_PER_ENV_STEP = 120 + 4 + 1 + 1 + 2.0 * STATE_BYTES_PER_ENV / ROLLOUT_K   # obs f32x30, reward, done, action, state/K
METRIC = "env-steps/sec (whole box) 2v2 at 2^20 envs"
UNIT = "env-steps/s"


def measured_traffic(n_local, K):
    """dram__bytes_read.sum + dram__bytes_write.sum of one rollout launch of this size, from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            for rec in json.load(f)["launches"]:
                if rec["envs"] == n_local and rec["K"] == K:
                    return rec["dram_bytes"]
    except Exception:  # noqa: BLE001
        pass
    return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


# ----------------------------------------------------------------------------------- CPU legs
def cpu_port_baseline(budget_s=12.0, threads=None):
    """The C oracle port (oracle/futbol_v0_oracle.c) on all host cores, bounded sample of the same workload."""
    from oracle.v0 import OracleV0
    threads = threads or os.cpu_count() or 1
    n = 8192
    orc = OracleV0(n, seed=0, random_opp=False, arith=0)
    orc.rollout(ROLLOUT_K, autoreset=2, n_threads=threads, record=False)      # warm-up
    t0, steps = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s:
        orc.rollout(ROLLOUT_K, autoreset=2, n_threads=threads, record=False)
        steps += n * ROLLOUT_K
    dt = time.perf_counter() - t0
    return {"value": steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d envs x %d env-steps (%d rollouts of K=%d) in %.1f s, random_opp=False, C oracle, %d pthreads"
                      % (n, steps // n, steps // (n * ROLLOUT_K), ROLLOUT_K, dt, threads)}


def _py_ref_worker(widx, n_envs, env_steps, q_in, q_out):
    from oracle.ref_harness import RefEnvV0
    import random
    envs = [RefEnvV0(seed=widx, env_id=i, random_opp=False, rng="mt") for i in range(n_envs)]
    rnd = random.Random(widx)
    q_out.put("ready")
    while True:
        cmd = q_in.get()
        if cmd is None:
            return
        done_steps = 0
        while done_steps < env_steps:
            for e in envs:
                _, _, d, _ = e.step(rnd.randrange(16))
                if d:
                    e.reset()
            done_steps += n_envs
        q_out.put(done_steps)


class PythonReferencePool:
    """SubprocVecEnv-style pool over the UNMODIFIED reference FutbolEnv: P processes x M envs each."""

    def __init__(self, procs, envs_per_proc, env_steps_per_step):
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        self.procs, self.q_in, self.q_out = [], [], ctx.Queue()
        for w in range(procs):
            qi = ctx.Queue()
            p = ctx.Process(target=_py_ref_worker, args=(w, envs_per_proc, env_steps_per_step, qi, self.q_out), daemon=True)
            p.start()
            self.procs.append(p); self.q_in.append(qi)
        for _ in range(procs):
            assert self.q_out.get(timeout=120) == "ready"

    def step(self):
        for qi in self.q_in:
            qi.put(1)
        return sum(self.q_out.get() for _ in self.procs)

    def close(self):
        for qi in self.q_in:
            qi.put(None)
        for p in self.procs:
            p.join(timeout=5)


def python_reference_available():
    from oracle.ref_harness import find_reference_root
    return find_reference_root()


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    base = {"metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "impl": "reference", "gpu_launches": 0,
            "config": {"workload": "FutbolEnv v0 2v2 vs hard-coded opponents (random_opp=False), uniform random AI actions, "
                                   "reset on done; bounded CPU sample of the 2^20-env x K=64 workload"}}
    root = python_reference_available()
    if root is not None:
        per_proc_steps = 1500                     # env-steps per process per bench step (~0.15 s)
        pool = PythonReferencePool(cores, 4, per_proc_steps)
        for _ in range(args.warmup):
            pool.step()
        t0 = time.perf_counter()
        total = sum(pool.step() for _ in range(args.steps))
        dt = time.perf_counter() - t0
        pool.close()
        val = total / dt
        kind = "reference"
        sample = ("unmodified reference FutbolEnv (%s) with gym/matplotlib stand-ins and an injected stdlib-MT RNG; "
                  "%d processes x 4 envs, %d env-steps per bench step" % (root, cores, total // max(1, args.steps)))
    else:
        port = cpu_port_baseline(budget_s=max(5.0, 1.0 * args.steps), threads=cores)
        val, dt, kind, sample = port["value"], None, "port", port["sample"]
    base.update({"value": val, "ms_per_step": (dt / args.steps * 1e3) if dt else None,
                 "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                 "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    _emit(json.dumps(base))


_RESULT_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner from C when
    NCCL_DEBUG is set on the box), so file descriptor 1 is pointed at stderr for the run and the result line goes to the
    saved original."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (line + "\n").encode())


# ----------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=TOTAL_ENVS, help="total envs of the job (default 2^20)")
    ap.add_argument("--random-opp", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the non-headline configurations (configs object)")
    ap.add_argument("--bind-cpus", type=int, default=1, help="pin each rank to its GPU's local CPUs before allocating pinned memory")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from gym_futbol_b200 import FutbolVecEnv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from gym_futbol_b200.sharding import shard
    assert args.envs % world == 0, "the metric's env count must split evenly over the ranks"
    first_env, n_local = shard(args.envs, rank, world)
    K = ROLLOUT_K
    env = FutbolVecEnv(n_local, device=dev, seed=0, env_id_offset=first_env, random_opp=bool(args.random_opp))
    env.reset()
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    acts = torch.randint(0, 16, (K, n_local), dtype=torch.uint8, device=dev, generator=g)   # resident in HBM
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput (`value`) ----
    for _ in range(max(3, args.warmup)):
        env.rollout(K, actions=acts)
    barrier()
    launches0 = env.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for a, b in evs:
        a.record(stream)
        env.rollout(K, actions=acts)
        b.record(stream)
    t_end.record(stream)
    barrier()
    launches = env.launch_count - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    kernel_ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    t = torch.tensor([elapsed_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, kernel_ms = t.tolist()
    value = args.envs * K * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the public API with HOST buffers (`e2e`) ----
    # What a host-side consumer of the reference API gets (gym_futbol_b200/host_io.py): actions come from pinned host
    # memory, and every observation, reward and done flag of the step is delivered back into pinned host memory, the copy
    # of one chunk overlapped with the simulation of the next.
    from gym_futbol_b200 import host_io
    local_cpus = host_io.bind_to_local_cpus(local_rank) if args.bind_cpus else None   # before the pinned allocations
    pipe = host_io.HostRollout(env, K, chunks=E2E_CHUNKS)
    pipe.h_actions.copy_(torch.randint(0, 16, (K, n_local), dtype=torch.uint8))
    d2h_peak = host_io.measure_d2h_peak(dev, nbytes=1 << 30, reps=3, barrier=barrier)   # all ranks copy at once
    for _ in range(2):
        pipe.run()
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        pipe.run()                                   # the host consumer holds the whole step before the next begins
    e1.record(pipe.copy_stream)                      # the last device->host copy of the last step ends here
    barrier()
    t = torch.tensor([e0.elapsed_time(e1), -d2h_peak], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms, d2h_peak_min = t[0].item(), -t[1].item()
    e2e_value = args.envs * K * e2e_steps / (e2e_ms * 1e-3)
    e2e_launches = E2E_CHUNKS * e2e_steps
    d2h_rate = pipe.d2h_bytes * e2e_steps / (e2e_ms * 1e-3) / 1e9       # per GPU

    for _ in range(2):
        pipe.run_resident()
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(stream)
    for _ in range(e2e_steps):
        pipe.run_resident()
    r1.record(stream)
    barrier()
    t = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    resident_value = args.envs * K * e2e_steps / (t.item() * 1e-3)

    stats = torch.from_numpy(env.stats.cpu().numpy().view("int64").copy()).to(dev)   # optional statistics gather
    if world > 1:
        dist.all_reduce(stats[1:6], op=dist.ReduceOp.SUM)
    slices, kernel_name = env.rollout_slices(K), env.rollout_kernel(K)
    h2d_bytes, d2h_bytes = pipe.h2d_bytes, pipe.d2h_bytes
    del pipe, acts
    env.close()
    del env
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configurations, timed in the same clock-sampled run (rank 0 of a 1-GPU run) ----
    peak, peak_src = measured_peaks()
    configs = None
    if world == 1 and not args.no_configs:
        configs = {}
        for name, fn in (("step_api_4096", lambda: bench_step_api(torch, dev, peak)),
                         ("v1_2v2_2p20", lambda: bench_v1_rollout(torch, dev, peak, 2, 1 << 20)),
                         ("v1_5v5_2p18", lambda: bench_v1_rollout(torch, dev, peak, 5, 1 << 18)),
                         ("v1_10v10_2p16", lambda: bench_v1_rollout(torch, dev, peak, 10, 1 << 16)),
                         ("v1_5v5_2p18_k256", lambda: bench_v1_rollout(torch, dev, peak, 5, 1 << 18, K=256, reps=4, cpu_seconds=0.0)),
                         ("ppo_65536", lambda: bench_ppo(torch, dev))):
            try:
                configs[name] = fn()
            except Exception as exc:  # noqa: BLE001  (a failing side configuration must not lose the headline line)
                configs[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            torch.cuda.empty_cache()
    clocks = sampler.stop() if rank == 0 else None      # sampled across every timed region of this process

    if rank == 0:
        achieved = n_local * K * BYTES_PER_ENV_STEP / (kernel_ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "FutbolEnv v0 2v2 vs hard-coded opponents (random_opp=%s), %d envs total, fused K=%d "
                                   "rollout per step, uniform random actions read from an HBM-resident [K,n] u8 tensor"
                                   % (bool(args.random_opp), args.envs, K),
                       "envs_per_gpu": n_local, "rollout_k": K, "parallelism": "env-sharded x%d, no collective" % world,
                       "l2": "each step writes %.2f GB per GPU (obs/reward/done), larger than the 126 MB L2"
                             % (n_local * K * 125 / 1e9)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes) * world,
                    "d2h_bytes_per_step": int(d2h_bytes) * world, "launches_per_step": E2E_CHUNKS,
                    "note": "pinned-host actions in; EVERY observation, reward and done flag of the step copied back to "
                            "pinned host memory (host-link bound), copy of chunk c overlapped with simulation of chunk c+1 "
                            "(gym_futbol_b200.host_io.HostRollout)",
                    "roofline": {"bound": "host link (device->host)", "achieved": d2h_rate, "peak": d2h_peak_min, "unit": "GB/s per GPU",
                                 "frac": d2h_rate / d2h_peak_min if d2h_peak_min > 0 else None,
                                 "peak_source": "pinned 1 GiB device->host copies timed in this process, all %d ranks copying at "
                                                "the same time, slowest rank's best of 3" % world,
                                 "cpu_affinity": "GPU-local CPUs %s" % (("%d-%d" % (local_cpus[0], local_cpus[-1])) if local_cpus else "not set")},
                    "obs_resident_in_hbm": {"value": resident_value, "unit": UNIT,
                                            "note": "same loop, observations consumed on the device (zero-copy policy): "
                                                    "host actions in, statistics out"}},
            "gpu_launches": int(launches), "gpu_launches_e2e": int(e2e_launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(n_local, K), "peak_source": peak_src,
                         "kernel": ("%s (the same step code; (time slice, env block) units from a work queue, %d slices)"
                                    % (kernel_name, slices)) if slices > 1 else kernel_name,
                         "bytes_per_env_step": BYTES_PER_ENV_STEP, "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": n_local * K * BYTES_PER_ENV_STEP,
                         "note": "the kernel is instruction-issue bound (fp64 IEEE sqrt/div sequences, selects, Philox), "
                                 "not HBM bound: see profiles/README.md; traffic = ncu dram bytes of one launch of this "
                                 "size (profiles/traffic.json), null if not captured for this size"},
            "rollout_stats": {"episodes": int(stats[2]), "goals_ai": int(stats[3]), "goals_opp": int(stats[4]),
                              "out_of_field": int(stats[5])},
        }
        if configs is not None:
            out["configs"] = configs
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_port_baseline()
            out["cpu_baseline_reference"] = cpu_reference_baseline()
        _emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------- the other configs
def _timed(torch, fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def bench_step_api(torch, dev, peak, n=4096, steps=1000):
    """BASELINE.json configs[1]: v0 vs hard-coded opponents, 4096 envs, 100 warm-up + 1000 timed steps through the PER-STEP
    API (one launch per step, state round-trips HBM: 572 B per env-step), eager and as CUDA graphs of 100 steps, and the
    same 1000 steps as one fused rollout."""
    from gym_futbol_b200 import FutbolVecEnv
    env = FutbolVecEnv(n, device=dev, seed=0, random_opp=False)
    env.reset()
    acts = torch.randint(0, 16, (steps, n), dtype=torch.uint8, device=dev)
    for t in range(100):
        env.step(acts[t])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for t in range(steps):
        env.step(acts[t])
    b.record()
    torch.cuda.synchronize()
    wall_us = (time.perf_counter() - t0) / steps * 1e6
    eager_us = a.elapsed_time(b) / steps * 1e3
    g, side = torch.cuda.CUDAGraph(), torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for t in range(100):
                env.step(acts[t])
    torch.cuda.current_stream(dev).wait_stream(side)
    g.replay()
    graph_us = _timed(torch, g.replay, 20) / 100 * 1e3
    env.rollout(steps, actions=acts)
    fused_ms = _timed(torch, lambda: env.rollout(steps, actions=acts), 10)
    bpe = 126 + 2 * STATE_BYTES_PER_ENV
    out = {"workload": "v0 2v2 vs hard-coded opponents, %d envs, per-step API" % n, "timed_steps": steps,
           "eager": {"us_per_step_device": eager_us, "us_per_step_wall": wall_us, "env_steps_per_s": n / (wall_us * 1e-6),
                     "note": "one Python call + one launch per step; bound by the host call"},
           "cuda_graph_100_steps": {"us_per_step": graph_us, "env_steps_per_s": n / (graph_us * 1e-6), "replays_timed": 20,
                                    "hbm_frac": n / (graph_us * 1e-6) * bpe / 1e9 / peak},
           "fused_rollout_k1000": {"ms": fused_ms, "env_steps_per_s": n * steps / (fused_ms * 1e-3), "launches_timed": 10},
           "bytes_per_env_step": bpe, "grid": "%d blocks on the GPU's SMs" % ((n + 127) // 128)}
    env.close()
    return out


def bench_v1_rollout(torch, dev, peak, N, n, K=ROLLOUT_K, reps=10, cpu_seconds=2.0):
    """BASELINE.json configs[4] (5v5 at 2^18 envs) and its siblings: the v1 N-vs-N rigid-body variant, fused K = 64 rollouts
    with given uniform random left-team actions (HBM resident), random right team drawn in-kernel."""
    from gym_futbol_b200 import FutbolV1VecEnv
    env = FutbolV1VecEnv(n, number_of_player=N, device=dev, seed=0)
    env.reset()
    # a pool of distinct action tables, cycled (period 8 K steps > one episode): one table reused every launch would repeat each
    # player's K actions for ever -- players then pile up at the walls and a step has 40 % more contacts than random play
    tables = [torch.randint(0, 5, (K, n, 2 * N), dtype=torch.uint8, device=dev) for _ in range(8)]
    turn = [0]

    def launch():
        env.rollout(K, actions=tables[turn[0] % len(tables)])
        turn[0] += 1
    kernel, slices = env.rollout_kernel(K), env.rollout_slices(K)
    for _ in range(3):
        launch()
    env.read_stats(clear=True)
    # the automatic launch and, for the record, the plain launch of the same batch, alternating (the envs' episodes run in
    # step, so the cost of a step depends on where in the episode the launch falls: both see the same mix)
    ms = ms_plain = 0.0
    for _ in range(reps):
        env.set_rollout_slices(0)
        ms += _timed(torch, launch, 1) / reps
        env.set_rollout_slices(1)
        ms_plain += _timed(torch, launch, 1) / reps
    env.set_rollout_slices(0)
    st = env.read_stats()
    B = 2 * N + 1
    P = B * (B - 1) // 2 + 12 * B
    state_bytes = 48 * B + 18                                     # bodies + scalars; the arbiter cache is touched only by contacts
    bpe = (4 + 8 * N) * 4 + 4 + 1 + 2 * N + 2.0 * state_bytes / K   # of the plain launch: the fraction below does not credit hand-overs
    rate = n * K / (ms * 1e-3)
    cpu = None
    if cpu_seconds > 0:                                        # the C restatement (oracle/futbol_v1_oracle.c) on the host cores, as context
        from oracle.v1 import OracleV1
        threads = os.cpu_count() or 1
        orc = OracleV1(2048, seed=0, number_of_player=N)
        orc.rollout(K, autoreset=2, n_threads=threads, record=False)
        t0, steps = time.perf_counter(), 0
        while time.perf_counter() - t0 < cpu_seconds:
            orc.rollout(K, autoreset=2, n_threads=threads, record=False)
            steps += 2048 * K
        cpu = {"value": steps / (time.perf_counter() - t0), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "2048 envs x %d env-steps, C oracle, %d pthreads (the reference's own v1 needs pymunk, absent here)" % (steps // 2048, threads)}
    out = {"workload": "v1 Futbol %dv%d, %d envs, fused K=%d rollout, given uniform random left actions (8 distinct HBM-resident tables, cycled)" % (N, N, n, K), "launches_timed": reps,
           "kernel": kernel, "time_slices": slices, "bytes_per_env_step_with_handover": bpe + 2.0 * state_bytes * (slices - 1) / K, "env_steps_per_s_plain_launch": n * K / (ms_plain * 1e-3),
           "ms_per_launch": ms, "env_steps_per_s": rate, "bytes_per_env_step": bpe, "hbm_gbs": rate * bpe / 1e9,
           "hbm_frac": rate * bpe / 1e9 / peak, "contacts_per_env_step": st["contacts"] / max(1, st["env_steps"]),
           "contacts_dropped": st["contacts_dropped"], "arbiter_cache_bytes_per_env": 16 * P, "cpu_baseline": cpu,
           "note": "latency bound (sequential turns, Gauss-Seidel contact solver, divergent contacts): see profiles/r2_v1_history.md"}
    env.close()
    return out


def bench_ppo(torch, dev, n=65536, T=128, iters=4):
    """BASELINE.json configs[3]: PPO collection and collection + update at 65,536 envs (examples/ppo_v0.py), with the fused
    glue (step(out=...) into the rollout buffers, one-launch minibatch gather) and, for the before/after, without."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ppo_v0", os.path.join(ROOT, "examples", "ppo_v0.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {"workload": "PPO on v0 2v2 vs hard-coded opponents, %d envs, n_steps %d, 4 minibatches x 4 epochs, MLP [256,256]+[128,128] "
                       "heads (bf16 autocast), collection replayed as one CUDA graph" % (n, T), "iterations_timed": iters}
    for label, fused in (("fused", True), ("unfused", False)):
        ppo = mod.PPO(n_envs=n, n_steps=T, seed=0, device=str(dev), fused=fused)
        ppo.collect(); ppo.update()                      # eager warm-up iteration
        ppo.prepare_graph()
        ppo.collect(); ppo.update()
        torch.cuda.synchronize()
        c_ms = u_ms = 0.0
        for _ in range(iters):
            c_ms += _timed(torch, ppo.collect, 1)
            u_ms += _timed(torch, ppo.update, 1)
        out[label] = {"collect_env_steps_per_s": n * T * iters / (c_ms * 1e-3), "collect_ms": c_ms / iters,
                      "collect_update_env_steps_per_s": n * T * iters / ((c_ms + u_ms) * 1e-3), "update_ms": u_ms / iters,
                      "obs_zero_copy": True}
        ppo.env.close()
        del ppo
        torch.cuda.empty_cache()
    out["note"] = ("fused = the step kernel writes obs/reward/done into the rollout-buffer rows (no per-step copies) and one gather launch "
                   "builds each minibatch; unfused = copy per step + six torch indexing launches per minibatch; the torch MLP is the cost")
    return out


def cpu_reference_baseline(budget_s=8.0):
    """The UNMODIFIED Python reference FutbolEnv, SubprocVecEnv style over the host cores, timed in this same run."""
    root = python_reference_available()
    cores = os.cpu_count() or 1
    if root is None:
        return {"unavailable": "reference package not found (looked for baseline/_ref)"}
    pool = PythonReferencePool(cores, 4, 1500)
    pool.step()
    t0, total = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget_s:
        total += pool.step()
    dt = time.perf_counter() - t0
    pool.close()
    return {"value": total / dt, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "unmodified reference FutbolEnv (%s), gym/matplotlib stand-ins + injected stdlib-MT RNG, random_opp=False; %d processes x 4 envs, "
                      "%d env-steps in %.1f s" % (root, cores, total, dt)}


if __name__ == "__main__":
    main()
