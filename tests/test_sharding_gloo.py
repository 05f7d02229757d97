"""N > 1 host logic on CPU: two ``gloo`` ranks shard an environment batch and reduce the statistics record."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from free_range_zoo_b200.distributed import all_reduce_statistics, episode_statistics, shard


def test_shards_partition_the_batch():
    for total, world in ((524288, 8), (10, 4), (7, 8)):
        blocks = [shard(total, rank, world) for rank in range(world)]
        assert blocks[0][0] == 0 and sum(count for _, count in blocks) == total
        for (offset, count), (next_offset, _) in zip(blocks, blocks[1:]):
            assert offset + count == next_offset


def _worker(rank, world, port, total_envs, results):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    offset, count = shard(total_envs, rank, world)
    generator = torch.Generator().manual_seed(0)
    everything = torch.rand((total_envs, 3), generator=generator)  # the same "global" rollout result on every rank
    mine = everything[offset:offset + count]
    record = episode_statistics(mine, mine[:, 0] > 0.5, mine[:, 1] > 0.5, torch.full((count, ), 10))
    all_reduce_statistics(record)
    if rank == 0:
        expected = episode_statistics(everything, everything[:, 0] > 0.5, everything[:, 1] > 0.5,
                                      torch.full((total_envs, ), 10))
        results.put(bool(torch.allclose(record, expected)))
    dist.barrier()
    dist.destroy_process_group()


def test_statistics_all_reduce_over_two_gloo_ranks():
    context = mp.get_context('spawn')
    results = context.Queue()
    port = 29500 + os.getpid() % 2000
    workers = [context.Process(target=_worker, args=(rank, 2, port, 1001, results)) for rank in range(2)]
    for worker in workers:
        worker.start()
    assert results.get(timeout=120)
    for worker in workers:
        worker.join(timeout=60)
        assert worker.exitcode == 0
