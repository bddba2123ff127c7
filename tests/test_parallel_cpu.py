"""Host-side multi-process logic on CPU (gloo, world_size 2): shard ownership, the final gather of sharded sampling and
the in-place bucketed gradient all-reduce that the training executor drives.  The CUDA kernels are not involved here;
the N>1 GPU path itself is timed by ``bench.py --gpus N``."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, fn_name: str, out_dir: str) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world)
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def _spawn(fn_name: str, tmp_path, world: int = 2) -> None:
    mp.spawn(_worker, args=(world, _free_port(), fn_name, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(tmp_path, f"ok{r}")) for r in range(world))


def test_shard_bounds_cover_the_batch_exactly():
    from dmme_b200.parallel import shard_bounds
    for total in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _gather_case(rank: int, world: int) -> None:
    from dmme_b200.parallel import gather_shards, shard_bounds
    for total in (8, 7):  # even and ragged
        full = torch.arange(total * 6, dtype=torch.float32).reshape(total, 2, 3)
        lo, hi = shard_bounds(total, rank, world)
        got = gather_shards(full[lo:hi].clone(), total)
        assert torch.equal(got, full)
        got0 = gather_shards(full[lo:hi].clone(), total, dst=0)
        assert (got0 is None) == (rank != 0)
        if rank == 0:
            assert torch.equal(got0, full)


def test_gather_shards_gloo(tmp_path):
    _spawn("_gather_case", tmp_path)


def _bucket_case(rank: int, world: int) -> None:
    from dmme_b200.parallel import GradBucketer, allreduce_gradients
    n = 1000
    arena = torch.full((n + 24,), float(rank + 1))
    arena[:n] += torch.arange(n, dtype=torch.float32)
    b = GradBucketer(arena, bucket_bytes=4 * 256)  # 256-element buckets
    b.mark(100)       # nothing complete yet
    assert b.buckets == []
    b.mark(600)       # two complete buckets
    assert b.buckets == [(0, 256), (256, 512)]
    b.finish(n)       # third bucket + tail
    assert b.buckets == [(0, 256), (256, 512), (512, 768), (768, 1000)]
    want = torch.arange(n, dtype=torch.float32) + sum(range(1, world + 1)) / world
    assert torch.allclose(arena[:n], want)
    assert torch.equal(arena[n:], torch.full((24,), float(rank + 1)))  # beyond the cursor: untouched
    # parameter-list variant
    ps = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7))]
    for i, p in enumerate(ps):
        p.grad = torch.full_like(p, float(rank + i))
    allreduce_gradients(ps, bucket_bytes=64)
    for i, p in enumerate(ps):
        assert torch.allclose(p.grad, torch.full_like(p, i + (world - 1) / 2))


def test_grad_bucketer_gloo(tmp_path):
    _spawn("_bucket_case", tmp_path)


def test_grad_bucketer_single_process_is_a_no_op():
    from dmme_b200.parallel import GradBucketer
    arena = torch.arange(10, dtype=torch.float32)
    b = GradBucketer(arena, bucket_bytes=16)
    b.finish(10)
    assert torch.equal(arena, torch.arange(10, dtype=torch.float32))
    assert b.buckets[0] == (0, 4) and b.buckets[-1][1] == 10
