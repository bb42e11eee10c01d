"""Episode statistics across ranks.

Environments shard by contiguous global index over the GPUs of a box (one process
per GPU); the step path has NO collective.  The only exchange is one SUM
all-reduce per rollout of the 16-double episode-statistics vector each rank's step
kernel accumulates (fields of env:575-613, main.py:272-282,
tensorboardlogger.py:80-108) -- NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU
tests.  Latency-bound: 128 bytes.
"""
import numpy as np
import torch

from .vec_env import STATS_KEYS


def shard_range(total_envs, rank, world_size):
    """Contiguous [lo, hi) slice of the global env index space owned by `rank`."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_stats_tensor(t, group=None):
    """In-place SUM all-reduce of a [16] float64 statistics tensor (no-op without a process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def summarize(vec):
    """Reference-style summary (main.py:305-327 results.txt + tensorboardlogger means) of a statistics vector."""
    v = np.asarray(vec, dtype=np.float64)
    d = dict(zip(STATS_KEYS, v.tolist()))
    n = max(d["episodes"], 1.0)
    runs = max(d["n_successful_runs"] + d["n_failed_runs"], 1.0)
    return {
        "episodes": int(d["episodes"]),
        "Successes": int(d["n_successful_runs"]), "Fails": int(d["n_failed_runs"]), "Collisions": int(d["n_collisions"]),
        "Success rate": d["n_successful_runs"] / runs, "Collision rate": d["n_collisions"] / runs,
        "Average APE": d["APE"] / n, "Average flight time": d["env_steps"] / n,
        "episodes/avg_reward": d["reward"] / n, "episodes/avg_length": d["env_steps"] / n,
        "episodes/avg_collision_reward": d["collision_reward"] / n,
        "episodes/avg_collision_avoidance_reward": d["collision_avoidance_reward"] / n,
        "episodes/avg_path_adherence_reward": d["path_adherence"] / n,
        "episodes/avg_path_progression_reward": d["path_progression"] / n,
        "episodes/avg_reach_end_reward": d["reach_end_reward"] / n,
        "episodes/avg_agressive_alpha_reward": d["agressive_alpha_reward"] / n,
        "mean_total_reward": d["total_reward"] / n,
    }


def global_episode_stats(env, group=None, reset=True):
    """All-reduce this rank's device-side accumulator and return the global summary."""
    t = env.episode_stats_tensor().clone()
    reduce_stats_tensor(t, group)
    if reset:
        env.episode_stats(reset=True)
    torch.cuda.synchronize(env.device)
    return summarize(t.cpu().numpy())
