"""N>1 host logic on CPU with the gloo backend (world_size 2): env sharding and the
per-rollout episode-statistics reduction."""
import importlib
import os
import socket

import torch
import torch.multiprocessing as mp

PKG = "drone_2d_custom_gym_env_for_reinforcement_learning_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stats = importlib.import_module(PKG + ".stats")
    lo, hi = stats.shard_range(1000, rank, world)
    t = torch.zeros(16, dtype=torch.float64)
    t[0] = hi - lo                       # pretend every env of the shard finished one episode
    t[1] = (hi - lo) // 2                # successes
    t[2] = (hi - lo) - (hi - lo) // 2    # failures
    t[4] = 100.0 * (hi - lo)             # env_steps
    stats.reduce_stats_tensor(t)
    q.put((rank, lo, hi, t.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_index_space():
    stats = importlib.import_module(PKG + ".stats")
    for total, world in ((65536 * 8, 8), (1000, 3), (7, 8), (1048576, 8)):
        r = [stats.shard_range(total, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == total
        assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_stats_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort()
    assert (out[0][1], out[0][2], out[1][1], out[1][2]) == (0, 500, 500, 1000)
    for _, _, _, v in out:
        assert v[0] == 1000 and v[1] == 500 and v[2] == 500 and v[4] == 100000.0
    stats = importlib.import_module(PKG + ".stats")
    s = stats.summarize(out[0][3])
    assert s["episodes"] == 1000 and s["Success rate"] == 0.5 and s["Average flight time"] == 100.0
