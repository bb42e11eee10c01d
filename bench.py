#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched Drone2dEnv step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--impl ours|reference]

Workload (BASELINE.json configs[2]): mode='curriculum', scenario='stage_5' random
obstacle fields, 65 536 environments per GPU, uniform random actions in [-1,1]^2
read from device-resident action buffers, auto-reset on.  One "step" = one launch
of the fused step kernel over all environments of the rank (one env-step each).
Ranks shard the environments (contiguous global env ids) with no data-path
collective: scaling is weak.  Timing: CUDA events around K steps after W warm-up
steps, barrier + synchronize on both sides, max over ranks.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU port of the
reference environment (oracle/ref_port.py: Python glue + scipy fminbound + C
rigid-body step, one process per host core) on the same workload.
"""
import argparse
import ctypes as C
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
# accumulators are stored as float64 here, which adds 2*32*4 = 256 B  ->  540 + 256 = 796 B (stage_5).
ALGO_BYTES_STAGE5 = 796.0
ALGO_BYTES_SCENARIO = 657.0
WORKLOAD = "stage_5 random obstacle field, 65536 envs/GPU, random actions, auto-reset"


def _profiled_traffic(envs):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel, from the committed ncu
    summary of this workload (profiles/*_step_kernel.json, written by tools/make_profile_summary.py)."""
    try:
        d = os.path.join(REPO, "profiles")
        best = None
        for f in sorted(os.listdir(d)):
            if f.endswith("_step_kernel.json") and "interim" not in f:
                best = os.path.join(d, f)
        if best is None:
            return None, None
        with open(best) as fh:
            j = json.load(fh)
        if int(j.get("envs_per_launch", -1)) != int(envs):
            return None, None
        _profiled_traffic.limiter = {
            "note": "what actually bounds the kernel (same ncu capture): float64 latency/issue, not HBM",
            "fp64_pipe_active_pct": j["launch"].get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_slots_active_pct": j["launch"].get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct_of_max": j["launch"].get("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "active_lanes_per_warp_instruction": j["launch"].get("smsp__thread_inst_executed_per_inst_executed.ratio"),
        }
        return float(j["dram_bytes_per_launch"]), os.path.basename(best)
    except Exception:  # noqa: BLE001
        return None, None


_profiled_traffic.limiter = None


def _peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback"


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
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def bind_to_gpu_numa_node(local):
    """One process per GPU: run on (and first-touch the page-locked buffers from) the CPU cores NVML reports as
    local to this rank's GPU, so that the host end of the PCIe traffic stays on the GPU's NUMA node.  Best effort:
    returns the number of cores bound to, or 0 if NVML / the affinity call is unavailable."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w in range(words) for b in range(64) if (mask[w] >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001
        return 0


def cpu_baseline_sample(budget_s, scenario="stage_5", mode="curriculum", workers=None):
    """Time the Python port of the reference env (oracle/ref_port.py) on the host cores:
    one process per usable core, random actions, auto-reset, ~budget_s seconds each."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workers = workers or cores
    ctx = mp.get_context("fork")
    q = ctx.Queue()

    procs = [ctx.Process(target=_cpu_worker, args=(w, budget_s, mode, scenario, q)) for w in range(workers)]
    t0 = time.time()
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    wall = time.time() - t0
    steps = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return {"value": steps / busy, "unit": UNIT, "cores": workers, "kind": "port",
            "sample": "%d env-steps of %s/%s in %.1f s on %d processes (Python port of the reference env: "
                      "scipy fminbound x2 per step + C rigid-body step; pymunk itself is not installable here)"
                      % (steps, mode, scenario, busy, workers), "wall_s": wall}


def _cpu_worker(w, budget_s, mode, scenario, q):
    import random
    import numpy as np
    from oracle import ref_port
    random.seed(1000 + w)
    np.random.seed(1000 + w)
    env = ref_port.RefPortEnv(mode=mode, scenario=scenario)
    env.reset()
    rng = np.random.RandomState(w)
    for _ in range(20):                                    # warm-up
        _, _, d, _ = env.step(rng.uniform(-1, 1, 2).astype(np.float32))
        if d:
            env.reset()
    n = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        _, _, d, _ = env.step(rng.uniform(-1, 1, 2).astype(np.float32))
        n += 1
        if d:
            env.reset()
    q.put((n, time.perf_counter() - t0))


def run_reference(args):
    rank, world, _ = _dist_env()
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    # each "step" of the reference arm is a bounded sample: total CPU budget ~ 20 s
    budget = float(os.environ.get("D2D_CPU_BUDGET_S", "20"))
    cb = cpu_baseline_sample(budget)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm,
            # same meaning as in our arm: the time one pass over one batch of envs_per_gpu environments takes
            "ms_per_step": 1000.0 * args.envs_per_gpu / cb["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": args.envs_per_gpu},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import importlib
    import numpy as np
    import torch
    rank, world, local = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cores = bind_to_gpu_numa_node(local) if world > 1 else 0
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module(PKG)
    E = args.envs_per_gpu
    env = pkg.Drone2dVecEnv(E, device=dev, seed=args.seed, env_id_base=rank * E, info_mode="none",
                            mode="curriculum", scenario="stage_5")
    env.reset_device()
    # device-resident action ring: R distinct uniform(-1,1) batches, rotated every step.  R exceeds the
    # typical episode length, so every env sees i.i.d. actions within an episode (a short ring repeats
    # the same torque pattern and ends episodes ~2.5x sooner than action_space.sample() would).
    R = 256
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    act_ring = (torch.rand((R, E, 2), device=dev, generator=g) * 2 - 1).contiguous()
    state_mb = env.state.numel() / 1e6

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident throughput ("value") ----------------
    # Untimed pre-steps bring the batch to its steady state (episodes of different ages, auto-resets
    # and episode prefetching active) whatever --warmup is: the first ~100 steps after a global
    # reset have no finished episodes and run ~15 % faster than the sustained rate.
    settle = max(0, 300 - args.warmup)
    for k in range(settle):
        env.step_device(act_ring[k % R], terminal_obs=False)
    for k in range(args.warmup):
        env.step_device(act_ring[(settle + k) % R], terminal_obs=False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = env.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(args.steps):
        env.step_device(act_ring[k % R], terminal_obs=False)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = env.launch_count - l0
    clocks = sampler.stop()
    stats = env.episode_stats(reset=True)

    # ---------------- end-to-end through the host boundary ("e2e") ----------------
    e2e_steps = max(8, min(args.steps, 200))
    # the user-side host buffers are page-locked (torch pin_memory), as the boundary recommends
    t_act = torch.from_numpy(np.random.RandomState(rank).uniform(-1, 1, (4, E, 2)).astype(np.float32)).pin_memory()
    t_obs = torch.empty((E, 27), dtype=torch.float32).pin_memory()
    t_rew = torch.empty(E, dtype=torch.float32).pin_memory()
    t_done = torch.empty(E, dtype=torch.uint8).pin_memory()
    h_act, h_obs, h_rew, h_done = t_act.numpy(), t_obs.numpy(), t_rew.numpy(), t_done.numpy()
    L = env._L
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def host_step(k):
        a = h_act[k % 4]
        rc = L.d2d_step_host(env._h, a.ctypes.data, h_obs.ctypes.data, h_rew.ctypes.data, h_done.ctypes.data, 1, st)
        if rc:
            raise RuntimeError(L.d2d_last_error().decode())

    for k in range(3):
        host_step(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        host_step(k)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0

    # ---------------- reduce over ranks ----------------
    t = torch.tensor([ms, e2e_s * 1000.0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        st_t = env.episode_stats_tensor()
        dist.all_reduce(st_t, op=dist.ReduceOp.SUM)          # the per-rollout statistics reduction (NCCL)
    ms_max, e2e_ms_max = t.tolist()
    total_envs = E * world
    value = total_envs * args.steps / (ms_max / 1000.0)
    e2e_value = total_envs * e2e_steps / (e2e_ms_max / 1000.0)

    if rank == 0:
        peak, peak_src = _peaks()
        per_launch_ms = ms / max(launches, 1)
        achieved = E * ALGO_BYTES_STAGE5 / (per_launch_ms / 1000.0) / 1e9
        traffic, traffic_src = _profiled_traffic(E)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": E, "l2_note": "state+tables per GPU = %.0f MB > L2 "
                       "(126 MB): inputs larger than L2, no flush" % state_mb, "parallelism": "env-sharded x%d" % world, "host_cores_bound_per_rank": numa_cores,
                       "actions": "device-resident ring of %d random batches" % R,
                       "steady_state_presteps": settle},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": E * 2 * 4,
                    "d2h_bytes_per_step": E * (27 * 4 + 4 + 1), "steps": e2e_steps,
                    "api": "d2d_step_host (C ABI; page-locked host buffers mapped into the step kernel: actions are read and obs/reward/done written over PCIe inside every step, then the stream is synchronised)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": E * ALGO_BYTES_STAGE5,
                         "algorithmic_bytes_per_env_step": ALGO_BYTES_STAGE5,
                         "kernel": "d2d_step_kernel<true>", "kernel_ms": per_launch_ms,
                         "limiter": _profiled_traffic.limiter},
            "episodes_finished": stats["episodes"],
        }
        if args.cpu_baseline:
            budget = float(os.environ.get("D2D_CPU_BUDGET_S", "15"))
            cb = cpu_baseline_sample(budget)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    env.close()
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
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
