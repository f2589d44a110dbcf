"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous batch sharding, per-rank seeds, the single
all-gather of final latents, max-over-ranks timing."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ctrl_adapter_b200.distributed import gather_latents, max_over_ranks, rank_seed, shard_range


def test_shard_range_partitions():
    for gb in (1, 2, 7, 8, 64):
        for world in (1, 2, 4, 8):
            spans = [shard_range(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, gb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s, e = shard_range(gb, rank, world)
        # each rank "denoises" its own samples: value = global sample index, seeded per rank
        g = torch.Generator().manual_seed(rank_seed(1234, rank))
        local = torch.arange(s, e, dtype=torch.float32)[:, None, None].expand(-1, 4, 3).contiguous()
        noise = torch.randn(1, generator=g)
        full = gather_latents(local, gb)
        t = max_over_ranks(float(rank + 1), torch.device("cpu"))
        q.put((rank, full[:, 0, 0].tolist(), t, float(noise)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("gb", [5, 8])
def test_gather_and_timing_world2(gb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + gb
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gb, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, vals, t, _ in res:
        assert vals == [float(i) for i in range(gb)], f"rank {rank} gathered {vals}"
        assert t == 2.0
    assert res[0][3] != res[1][3]  # different per-rank seeds
