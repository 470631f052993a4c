"""CPU, gloo, world size 2: the N>1 host logic of bench.py / the pipeline -- rank-strided sharding, the single counter all_gather,
max-over-ranks timing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cartoonsegmentation_b200.utils.dist import gather_counters, max_over_ranks, shard_indices


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(11, rank, world)
    frames = float(len(mine))
    ms = 10.0 + 5.0 * rank
    g = gather_counters(torch.tensor([frames, ms], dtype=torch.float64))
    slowest = max_over_ranks(ms, torch.device("cpu"))
    q.put((rank, mine, g.tolist(), slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_counter_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, g0, s0), (r1, m1, g1, s1) = res
    assert sorted(m0 + m1) == list(range(11)) and not set(m0) & set(m1)            # every image exactly once
    assert g0 == g1 == [[6.0, 10.0], [5.0, 15.0]]                                  # both ranks see the same gathered counters
    assert s0 == s1 == 15.0                                                        # time = max over ranks
    total_frames, fps = sum(r[0] for r in g0), sum(r[0] for r in g0) / (s0 * 1e-3)
    assert total_frames == 11 and abs(fps - 11 / 0.015) < 1e-6


def test_single_process_paths():
    assert shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert gather_counters(torch.tensor([1.0, 2.0], dtype=torch.float64)).tolist() == [[1.0, 2.0]]
    assert max_over_ranks(3.5, torch.device("cpu")) == 3.5
