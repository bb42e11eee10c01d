"""End-to-end A/B at N ranks (run under torchrun, or alone): what bounds `e2e` when several GPUs share the host?

  1. ceiling: every rank copies its 7.4 MB of step outputs device -> page-locked host memory with plain
     cudaMemcpyAsync, all ranks at once -> aggregate GB/s (the box's concurrent D2H ceiling for this payload);
  2. d2d_step_host through mapped host memory (the kernel's own stores cross PCIe; default path);
  3. d2d_step_host through device staging buffers + copy engine (D2D_HOST_COPY=1 in a fresh process).

usage: [torchrun --nproc-per-node N] python tools/e2e_ab.py            (prints one JSON line on rank 0)
"""
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1 and os.environ.get("E2E_BIND", "1") == "1":
    allowed = sorted(os.sched_getaffinity(0))
    per = max(1, len(allowed) // world)
    os.sched_setaffinity(0, set(allowed[local * per:(local + 1) * per]))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
pkg = importlib.import_module("drone_2d_custom_gym_env_for_reinforcement_learning_b200")
E = 65536


def barrier():
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)


def max_over_ranks(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {"world": world, "cores_bound": len(os.sched_getaffinity(0)), "host_copy_env": os.environ.get("D2D_HOST_COPY", "0")}
# 1. plain D2H ceiling
d_obs = torch.zeros((E, 27), dtype=torch.float32, device=dev)
d_rew = torch.zeros(E, dtype=torch.float32, device=dev)
d_done = torch.zeros(E, dtype=torch.uint8, device=dev)
h_obs, h_rew, h_done = (torch.empty_like(t, device="cpu").pin_memory() for t in (d_obs, d_rew, d_done))
for _ in range(5):
    h_obs.copy_(d_obs, non_blocking=True); h_rew.copy_(d_rew, non_blocking=True); h_done.copy_(d_done, non_blocking=True)
barrier()
K = 200
t0 = time.perf_counter()
for _ in range(K):
    h_obs.copy_(d_obs, non_blocking=True); h_rew.copy_(d_rew, non_blocking=True); h_done.copy_(d_done, non_blocking=True)
    torch.cuda.synchronize(dev)
dt = max_over_ranks(time.perf_counter() - t0)
nbytes = E * (27 * 4 + 4 + 1)
out["d2h_ceiling"] = {"us_per_step": dt / K * 1e6, "GBps_per_gpu": nbytes * K / dt / 1e9, "GBps_aggregate": world * nbytes * K / dt / 1e9}

# 2./3. the step through the host boundary
env = pkg.Drone2dVecEnv(E, device=dev, seed=0, env_id_base=rank * E, info_mode="none", mode="curriculum", scenario="stage_5")
env.reset_device()
for _ in range(300):
    env.step_device(None, terminal_obs=False)
acts = torch.from_numpy(np.random.RandomState(rank).uniform(-1, 1, (4, E, 2)).astype(np.float32)).pin_memory()
L, st = env._L, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
a_np, o_np, r_np, d_np = acts.numpy(), h_obs.numpy(), h_rew.numpy(), h_done.numpy()


def host_steps(n):
    for k in range(n):
        rc = L.d2d_step_host(env._h, a_np[k % 4].ctypes.data, o_np.ctypes.data, r_np.ctypes.data, d_np.ctypes.data, 1, st)
        assert rc == 0, L.d2d_last_error()


host_steps(5)
barrier()
t0 = time.perf_counter()
host_steps(K)
torch.cuda.synchronize(dev)
dt = max_over_ranks(time.perf_counter() - t0)
out["d2d_step_host"] = {"us_per_step": dt / K * 1e6, "env_steps_per_s_aggregate": world * E * K / dt,
                        "GBps_aggregate": world * nbytes * K / dt / 1e9,
                        "path": "copy engine (staging buffers)" if out["host_copy_env"] == "1" else "mapped host memory (zero copy)"}
# device-only step for reference
barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(K):
    env.step_device(None, terminal_obs=False)
ev1.record()
torch.cuda.synchronize(dev)
out["device_only_us_per_step"] = max_over_ranks(ev0.elapsed_time(ev1) / K * 1e3)
if rank == 0:
    print(json.dumps(out), flush=True)
env.close()
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
