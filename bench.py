#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched Drone2dEnv step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--impl ours|reference] [--no-extras]

Headline workload (BASELINE.json configs[2]): mode='curriculum', scenario='stage_5' random obstacle fields,
65 536 environments per GPU, uniform random actions in [-1,1]^2 read from device-resident action buffers,
auto-reset on.  One "step" = one launch of the fused step kernel over all environments of the rank (one env-step
each).  Ranks shard the environments (contiguous global env ids) with no data-path collective: scaling is weak.

Timing: W warm-up steps, then blocks of EXACTLY K steps, each block bracketed by a barrier + synchronize and timed
with CUDA events (max over ranks); blocks are repeated until at least 0.5 s has been timed, `ms_per_step` / `value`
are the MEDIAN block, the spread is reported next to them.

Keys beyond the base contract: `roofline`, `cpu_baseline` (N = 1), `e2e` (C ABI with host buffers),
`e2e_python_api` (Drone2dVecEnv.step, what an SB3 user calls), `configs` (BASELINE configs 2, 4 and 5 on the same
box, same timing rules).  `--impl reference` times the CPU port of the reference environment in the reference's own
arrangement (one process per host core behind pipes, SB3 SubprocVecEnv protocol, main.py:88-101,190).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"
METRIC = "env-steps/sec (stage_5, 64K envs/GPU)"
UNIT = "env-steps/s"
# Algorithmic bytes per env-step, SURVEY.md §8(d): dynamic state read+written (35 words) + I/O (121 B)
# + per-env tables (waypoints 96 B + obstacles 42 B at 4-byte words); body state, impulses and
# accumulators are stored as float64 here, which adds 2*32*4 = 256 B  ->  540 + 256 = 796 B (stage_5),
# 401 + 256 = 657 B for the shared-table scenarios.
ALGO_BYTES_STAGE5 = 796.0
ALGO_BYTES_SCENARIO = 657.0
MIN_TIMED_S = float(os.environ.get("D2D_BENCH_MIN_S", "0.5"))    # lowered only for profiler runs (ncu replays every launch)
ACTION_RING = 256
SWEEP_SCENARIOS = ("S_corridor", "S_parallel", "perpendicular", "impossible")
AGENT = os.path.join(REPO, "tests", "golden", "agent_19_policy.npz")


def make_config(envs_per_gpu, world):
    """The `config` object of the JSON line -- identical for both arms (--impl ours / reference)."""
    return {"workload": "stage_5 random obstacle field, %d envs/GPU, random actions, auto-reset" % envs_per_gpu,
            "envs_per_gpu": envs_per_gpu, "parallelism": "env-sharded x%d" % world,
            "actions": "uniform random in [-1,1]^2 (ours: device-resident ring of %d batches)" % ACTION_RING,
            "l2": "state + tables per GPU = 284 MB > L2 (126 MB): inputs larger than L2, no flush"}


def kernel_source_hash():
    """Hash of the CUDA sources with comments and white space removed: identifies the CODE an ncu capture was taken on
    (a reworded comment does not make a capture stale, a changed statement does)."""
    import re
    h = hashlib.sha1()
    d = os.path.join(REPO, PKG, "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            src = open(os.path.join(d, f), "r", encoding="utf-8", errors="replace").read()
            src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)          # block comments
            src = re.sub(r"(?m)//[^\n]*$", " ", src)                   # line comments (no string literal of these files holds //)
            h.update(re.sub(r"\s+", " ", src).encode())
    return h.hexdigest()[:12]


def profiled_traffic(envs):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel from the newest committed ncu
    summary (profiles/*_step_kernel.json, tools/make_profile_summary.py).  The summary records the hash of the kernel
    sources it was taken on; if the sources changed since, the figure is reported as stale (traffic = null)."""
    try:
        d = os.path.join(REPO, "profiles")
        best = None
        for f in sorted(os.listdir(d)):
            if f.endswith("_step_kernel.json") and "interim" not in f:
                best = os.path.join(d, f)
        if best is None:
            return None, None, None
        with open(best) as fh:
            j = json.load(fh)
        if int(j.get("envs_per_launch", -1)) != int(envs):
            return None, os.path.basename(best), None
        L = j["launch"]
        limiter = {
            "note": "what actually bounds the kernel (same ncu capture): float64 latency/issue, not HBM",
            "fp64_pipe_active_pct": L.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_slots_active_pct": L.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct_of_max": L.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "active_lanes_per_warp_instruction": L.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
        }
        stale = j.get("kernel_source_hash") not in (None, kernel_source_hash())
        src = os.path.basename(best) + (" (STALE: kernel sources changed since the capture)" if stale else "")
        return (None if stale else float(j["dram_bytes_per_launch"])), src, limiter
    except Exception:  # noqa: BLE001
        return None, None, None


def peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def bind_rank_to_core_slice(local, world):
    """One process per GPU.  Each rank gets its own, DISJOINT slice of the cores this job may use (preferring the
    cores NVML reports as local to the rank's GPU when that still leaves a disjoint slice), so that the host ends
    of the ranks' PCIe traffic and their Python loops do not share cores.  Returns the number of cores bound."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        per = max(1, len(allowed) // world)
        mine = set(allowed[local * per:(local + 1) * per]) or set(allowed)
        os.sched_setaffinity(0, mine)
        return len(mine)
    except Exception:  # noqa: BLE001
        return 0


# --------------------------------------------------------------------------- #
# reference arm / CPU baseline: the reference's arrangement on the host cores #
# --------------------------------------------------------------------------- #
def _port_worker(remote, seed, mode, scenario):
    """SB3 SubprocVecEnv worker protocol (one env per process, pipe per worker): ('step', action) -> (obs, reward,
    done, info) with the env reset inside the worker when an episode ends; ('close', None)."""
    import random
    import numpy as np
    from oracle import ref_port
    random.seed(seed)
    np.random.seed(seed)
    env = ref_port.RefPortEnv(mode=mode, scenario=scenario)
    obs = env.reset()
    remote.send(obs)
    while True:
        cmd, data = remote.recv()
        if cmd == "step":
            obs, rew, done, info = env.step(data)
            if done:
                info = dict(info)
                info["terminal_observation"] = obs
                obs = env.reset()
            remote.send((obs, rew, done, info))
        else:
            remote.close()
            return


class PortSubprocVecEnv:
    """The reference's CPU arrangement (main.py:88-101,181-190: `SubprocVecEnv([make_mp_env(...)] * num_cpu)`) with the
    Python port of the reference env (oracle/ref_port.py: scipy fminbound twice per step, Python obstacle loop, C
    rigid-body step through ctypes like pymunk's cffi) in place of the pymunk env, which is not installable."""

    def __init__(self, workers, mode="curriculum", scenario="stage_5"):
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        self.remotes, self.procs = [], []
        for w in range(workers):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_port_worker, args=(b, 1000 + w, mode, scenario), daemon=True)
            p.start()
            b.close()
            self.remotes.append(a)
            self.procs.append(p)
        for r in self.remotes:
            r.recv()
        self.n = workers

    def step(self, actions):
        for r, a in zip(self.remotes, actions):
            r.send(("step", a))
        return [r.recv() for r in self.remotes]

    def close(self):
        for r in self.remotes:
            r.send(("close", None))
        for p in self.procs:
            p.join(timeout=5)


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def reference_steps(steps, warmup, budget_s):
    """W warm-up + K timed steps of the reference arrangement; a step is a bounded sample: S rounds of the vector
    env (one env-step per worker per round), S sized from the warm-up rate so that the K steps take ~budget_s."""
    import numpy as np
    workers = host_cores()
    venv = PortSubprocVecEnv(workers)
    rng = np.random.RandomState(0)

    def rounds(s):
        for _ in range(s):
            venv.step(rng.uniform(-1, 1, (workers, 2)).astype(np.float32))

    rounds(10)
    t0 = time.perf_counter()
    probe = 40
    rounds(probe)
    rate = probe * workers / (time.perf_counter() - t0)
    S = int(min(5000, max(1, round(budget_s * rate / (steps * workers)))))
    for _ in range(warmup):
        rounds(min(S, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        rounds(S)
    dt = time.perf_counter() - t0
    venv.close()
    n = steps * S * workers
    return {"value": n / dt, "unit": UNIT, "cores": workers, "kind": "port", "seconds": dt, "env_steps_per_step": S * workers,
            "sample": "%d env-steps of curriculum/stage_5 (%d steps x %d rounds x %d single-env worker processes behind pipes, "
                      "SB3 SubprocVecEnv protocol) in %.1f s; env = Python port of the reference (scipy fminbound x2 per "
                      "step + C rigid-body step); pymunk / gym / SB3 are not installed on this box" % (n, steps, S, workers, dt)}


def run_reference(args):
    rank, world, _ = _dist_env()
    if rank != 0:
        return
    budget = float(os.environ.get("D2D_CPU_BUDGET_S", "25"))
    cb = reference_steps(args.steps, args.warmup, budget)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            # measured: wall time of one timed step of THIS arm = one bounded sample of the workload
            "ms_per_step": 1000.0 * cb["seconds"] / args.steps, "env_steps_per_step": cb["env_steps_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(args.envs_per_gpu, world),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- #
# our arm                                                                     #
# --------------------------------------------------------------------------- #
class Timer:
    """Blocks of exactly K steps, each bracketed by barrier + synchronize, repeated until >= MIN_TIMED_S."""

    def __init__(self, torch, dist, dev):
        self.torch, self.dist, self.dev = torch, dist, dev

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def run(self, fn, steps, min_s=MIN_TIMED_S, max_reps=200):
        torch = self.torch
        ms, total = [], 0.0
        while True:
            self.barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            fn(steps)
            ev1.record()
            self.barrier()
            t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=self.dev)
            if self.dist is not None:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms.append(float(t.item()))
            total += ms[-1] / 1000.0
            if total >= min_s or len(ms) >= max_reps:
                break
        s = sorted(ms)
        return {"median": s[len(s) // 2], "min": s[0], "max": s[-1], "reps": len(s)}


def rate_record(envs_total, steps, t, algo_bytes, envs_per_gpu, peak):
    per_step_ms = t["median"] / steps
    ach = envs_per_gpu * algo_bytes / (per_step_ms / 1000.0) / 1e9
    return {"value": envs_total * steps / (t["median"] / 1000.0), "unit": UNIT, "ms_per_step": per_step_ms,
            "ms_per_step_min": t["min"] / steps, "ms_per_step_max": t["max"] / steps, "reps": t["reps"],
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "algorithmic_bytes_per_env_step": algo_bytes}}


def run_ours(args):
    import importlib
    import numpy as np
    import torch
    rank, world, local = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    cpu_base = None
    if args.cpu_baseline and world == 1:
        # before any CUDA / NCCL work: the host cores are otherwise idle
        cpu_base = reference_steps(8, 1, float(os.environ.get("D2D_CPU_BUDGET_S", "12")))
    cores_bound = bind_rank_to_core_slice(local, world) if world > 1 else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module(PKG)
    policy = importlib.import_module(PKG + ".policy")
    E, K = args.envs_per_gpu, args.steps
    peak, peak_src = peaks()
    timer = Timer(torch, dist, dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def ring_for(n, seed):
        g = torch.Generator(device=dev)
        g.manual_seed(seed + rank)
        return (torch.rand((ACTION_RING, n, 2), device=dev, generator=g) * 2 - 1).contiguous()

    # ---------------- headline: stage_5, E envs per GPU, device-resident ----------------
    env = pkg.Drone2dVecEnv(E, device=dev, seed=args.seed, env_id_base=rank * E, info_mode="none",
                            mode="curriculum", scenario="stage_5")
    env.reset_device()
    act_ring = ring_for(E, 1234)
    state_mb = env.state.numel() / 1e6
    cursor = [0]

    def steps_fn(n):
        k0 = cursor[0]
        for k in range(n):
            env.step_device(act_ring[(k0 + k) % ACTION_RING], terminal_obs=False)
        cursor[0] = k0 + n

    # Untimed pre-steps bring the batch to its steady state (episodes of different ages, auto-resets and episode
    # prefetching active) whatever --warmup is: the first ~100 steps after a global reset have no finished
    # episodes and run ~15 % faster than the sustained rate.
    settle = max(0, 300 - args.warmup)
    steps_fn(settle)
    steps_fn(args.warmup)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = env.launch_count
    t_main = timer.run(steps_fn, K)
    launches = env.launch_count - l0
    clocks = sampler.stop()
    stats = env.episode_stats(reset=True)

    # ---------------- end to end through the C ABI with host buffers ("e2e") ----------------
    e2e_steps = max(8, min(K, 200))
    t_act = torch.from_numpy(np.random.RandomState(rank).uniform(-1, 1, (4, E, 2)).astype(np.float32)).pin_memory()
    t_obs = torch.empty((E, 27), dtype=torch.float32).pin_memory()
    t_rew = torch.empty(E, dtype=torch.float32).pin_memory()
    t_done = torch.empty(E, dtype=torch.uint8).pin_memory()
    h_act, h_obs, h_rew, h_done = t_act.numpy(), t_obs.numpy(), t_rew.numpy(), t_done.numpy()
    L = env._L

    def host_steps(n):
        for k in range(n):
            rc = L.d2d_step_host(env._h, h_act[k % 4].ctypes.data, h_obs.ctypes.data, h_rew.ctypes.data, h_done.ctypes.data, 1, st)
            if rc:
                raise RuntimeError(L.d2d_last_error().decode())

    def wall(fn, n):
        timer.barrier()
        t0 = time.perf_counter()
        fn(n)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    host_steps(3)
    e2e_s = wall(host_steps, e2e_steps)
    env.close()
    del env, act_ring

    # ---------------- end to end through the Python VecEnv API ("e2e_python_api") ----------------
    api_steps = max(8, min(K, 40))
    venv = pkg.Drone2dVecEnv(E, device=dev, seed=args.seed, env_id_base=rank * E, info_mode="episode",
                             mode="curriculum", scenario="stage_5")
    venv.reset()
    np_act = np.random.RandomState(7 + rank).uniform(-1, 1, (4, E, 2)).astype(np.float32)
    for k in range(300 // 10):                      # bring episodes to mixed ages cheaply on the device first
        for _ in range(10):
            venv.step_device(None, terminal_obs=False)
    n_term = [0]

    def api_fn(n):
        for k in range(n):
            _, _, dones, infos = venv.step(np_act[k % 4])
            n_term[0] += int(dones.sum())

    api_fn(3)
    n_term[0] = 0
    api_s = wall(api_fn, api_steps)
    venv.close()
    del venv

    extras = {}
    if args.extras:
        # ---------------- BASELINE config 2: corridor, 4096 envs ----------------
        e2 = pkg.Drone2dVecEnv(4096, device=dev, seed=args.seed, env_id_base=rank * 4096, info_mode="none",
                               mode="test", scenario="corridor")
        e2.reset_device()
        r2 = ring_for(4096, 99)
        c2 = [0]

        def f2(n):
            for k in range(n):
                e2.step_device(r2[(c2[0] + k) % ACTION_RING], terminal_obs=False)
            c2[0] += n

        f2(300)
        K2 = max(K, 200)
        rec = rate_record(4096 * world, K2, timer.run(f2, K2, min_s=0.2), ALGO_BYTES_SCENARIO, 4096, peak)
        rec["workload"] = "config 2: corridor scenario (18 circles), 4096 envs/GPU, random actions, auto-reset"
        extras["config2_corridor_4096"] = rec
        e2.close()
        del e2, r2

        # ---------------- BASELINE config 4: large scenario, agent_19 in the loop ----------------
        e4 = pkg.Drone2dVecEnv(E, device=dev, seed=args.seed, env_id_base=rank * E, info_mode="none",
                               mode="test", scenario="large")
        e4.reset_device()
        actor = policy.FusedActor.load(AGENT, e4)
        actor.rollout(300)
        l4 = e4.launch_count
        K4 = max(K, 200)
        t4 = timer.run(lambda n: actor.rollout(n), K4, min_s=0.3)
        rec = rate_record(E * world, K4, t4, ALGO_BYTES_SCENARIO, E, peak)
        s4 = e4.episode_stats()
        rec.update({"workload": "config 4: large scenario, agent_19 PPO actor (tcgen05 policy kernel) driving %d envs/GPU, "
                                "stochastic predict, closed loop on the device (d2d_rollout: 2 launches per step)" % E,
                    "launches_per_step": (e4.launch_count - l4) / (K4 * t4["reps"]),
                    "episodes": s4["episodes"], "success_rate": s4["n_successful_runs"] / max(s4["episodes"], 1)})
        extras["config4_large_agent19"] = rec
        e4.close()
        del e4, actor

        # ---------------- BASELINE config 5: 4-scenario sweep, 131072 envs per GPU (1 M on 8 GPUs) ----------------
        per = 32768
        envs5, rings5, streams5 = [], [], []
        for si, scen in enumerate(SWEEP_SCENARIOS):
            e5 = pkg.Drone2dVecEnv(per, device=dev, seed=args.seed + si, env_id_base=(rank * 4 + si) * per, info_mode="none",
                                   mode="test", scenario=scen)
            e5.reset_device()
            envs5.append(e5)
            rings5.append(ring_for(per, 500 + si)[:64].contiguous())
            streams5.append(torch.cuda.Stream(device=dev))
        c5 = [0]
        torch.cuda.synchronize(dev)

        def f5(n):
            cur = torch.cuda.current_stream(dev)
            for s in streams5:
                s.wait_stream(cur)
            for k in range(n):
                for e5, r5, s in zip(envs5, rings5, streams5):
                    with torch.cuda.stream(s):
                        e5.step_device(r5[(c5[0] + k) % 64], terminal_obs=False)
            for s in streams5:
                cur.wait_stream(s)
            c5[0] += n

        f5(300)
        K5 = max(K, 100)
        rec = rate_record(4 * per * world, K5, timer.run(f5, K5, min_s=0.3), ALGO_BYTES_SCENARIO, 4 * per, peak)
        rec["workload"] = ("config 5: %s, %d envs each = %d envs/GPU (x%d GPUs = %d envs), random actions, auto-reset; "
                           "one step = all four scenarios advance once (4 launches on 4 streams)"
                           % ("/".join(SWEEP_SCENARIOS), per, 4 * per, world, 4 * per * world))
        extras["config5_sweep_131072_per_gpu"] = rec
        for e5 in envs5:
            e5.close()

    # ---------------- report ----------------
    total_envs = E * world
    if dist is not None:
        st_t = torch.zeros(16, dtype=torch.float64, device=dev)
        dist.all_reduce(st_t, op=dist.ReduceOp.SUM)              # exercises the per-rollout statistics reduction (NCCL)
    if rank == 0:
        head = rate_record(total_envs, K, t_main, ALGO_BYTES_STAGE5, E, peak)
        traffic, traffic_src, limiter = profiled_traffic(E)
        roof = head["roofline"]
        roof.update({"traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": E * ALGO_BYTES_STAGE5, "kernel": "d2d_step_kernel<true>",
                     "kernel_ms": head["ms_per_step"], "kernel_source_hash": kernel_source_hash(), "limiter": limiter})
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "ms_per_step_min": head["ms_per_step_min"],
            "ms_per_step_max": head["ms_per_step_max"], "timed_blocks": head["reps"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(E, world),
            "run": {"state_mb_per_gpu": state_mb, "steady_state_presteps": settle, "host_cores_bound_per_rank": cores_bound,
                    "timing": "blocks of exactly `steps` steps, CUDA events, barrier+sync around each, max over ranks, "
                              "repeated until >= %.1f s; median block reported" % MIN_TIMED_S},
            "clocks": clocks, "gpu_launches": launches // max(t_main["reps"], 1), "gpu_launches_all_blocks": launches,
            "e2e": {"value": total_envs * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": E * 2 * 4,
                    "d2h_bytes_per_step": E * (27 * 4 + 4 + 1), "steps": e2e_steps, "ms_per_step": 1000.0 * e2e_s / e2e_steps,
                    "api": "d2d_step_host (C ABI; page-locked host buffers mapped into the step kernel: actions are read and "
                           "obs/reward/done written over PCIe inside every step, then the stream is synchronised)"},
            "e2e_python_api": {"value": total_envs * api_steps / api_s, "unit": UNIT, "steps": api_steps,
                               "ms_per_step": 1000.0 * api_s / api_steps,
                               "h2d_bytes_per_step": E * 2 * 4, "d2h_bytes_per_step": E * (27 * 4 + 4 + 1),
                               "terminal_observations_per_step": n_term[0] / api_steps,
                               "api": "Drone2dVecEnv.step(np.ndarray): SB3 VecEnv call, info_mode='episode' (info dicts + "
                                      "terminal_observation for every finished env), outputs copied to fresh numpy arrays"},
            "roofline": roof,
            "episodes_finished": stats["episodes"],
        }
        if extras:
            line["configs"] = extras
        if cpu_base is not None:
            line["cpu_baseline"] = {k: cpu_base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--envs-per-gpu", type=int, default=65536)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip BASELINE configs 2, 4, 5")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
