"""The N>1 host logic on CPU: world_size-2 gloo processes partition the objects like the reference
(gpu = id mod #GPUs), meet only at barriers, and reduce their timings with max / their work with sum."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from ro_map_b200 import partition


def test_round_robin_rule():
    assert partition.assign_objects(4, 1) == [[0, 1, 2, 3]]
    assert partition.assign_objects(8, 8) == [[k] for k in range(8)]
    assert partition.assign_objects(4, 2) == [[0, 2], [1, 3]]
    assert partition.assign_objects(3, 4) == [[0], [1], [2], []]          # ragged: an idle rank
    assert partition.assign_objects(0, 2) == [[], []]                      # empty
    for n, w in [(7, 3), (16, 8), (5, 5)]:
        parts = partition.assign_objects(n, w)
        assert sorted(sum(parts, [])) == list(range(n))
        assert all(partition.owner(k, w) == r for r, ks in enumerate(parts) for k in ks)
    with pytest.raises(ValueError):
        partition.assign_objects(1, 0)


def _worker(rank, world, port, n_objects, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = partition.assign_objects(n_objects, world)[rank]
    dist.barrier()
    ms = 10.0 + 5.0 * rank                       # pretend device time of the timed region on this rank
    worst = partition.reduce_max(ms)
    total_iters = partition.reduce_sum(100.0 * len(mine))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    dist.barrier()
    if rank == 0:
        out.put((worst, total_iters, gathered))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_objects", [4, 3])
def test_two_rank_gloo(n_objects):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_objects, q)) for r in range(2)]
    for p in procs:
        p.start()
    worst, total, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert worst == 15.0                                         # max over ranks, never a mean
    assert total == 100.0 * n_objects                            # whole-job work = sum over ranks
    assert sorted(sum(gathered, [])) == list(range(n_objects))   # every object trained exactly once
