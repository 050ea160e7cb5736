"""Host logic of the multi-GPU paths, on CPU: partition arithmetic and the world_size-2 timing
reduction bench.py uses (gloo backend)."""
import os
import socket

import numpy as np
import pytest

from goofy_b200 import sharding


def test_strips_cover_image_exactly_once():
    for w, h, stride in ((16384, 16384, 65792), (8192, 8192, 32768), (512, 36, 2048), (64, 4, 256)):
        for n in (1, 2, 4, 8):
            ss = sharding.strips(w, h, stride, n)
            assert sum(s.rows for s in ss) == h
            assert sum(s.dst_bytes for s in ss) == w * h // 2
            pos = 0
            for s in ss:
                assert s.first_row == pos and s.rows % 4 == 0
                assert s.src_offset == s.first_row * stride
                assert s.dst_offset == (s.first_row // 4) * (w // 4) * 8
                pos += s.rows


def test_config5_strip_offsets():
    """16384^2 with stride 65792 over 8 GPUs: strip g = rows [2048g, 2048(g+1)), output offset g*512*4096*8."""
    ss = sharding.strips(16384, 16384, 65792, 8)
    for g, s in enumerate(ss):
        assert (s.first_row, s.rows) == (2048 * g, 2048)
        assert s.dst_offset == g * 512 * 4096 * 8


def test_batch_partition_is_balanced_and_complete():
    for n_images in (0, 1, 7, 4096):
        for n in (1, 2, 4, 8):
            parts = [sharding.batch_partition(n_images, n, g) for g in range(n)]
            assert sum(len(p) for p in parts) == n_images
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            flat = [i for p in parts for i in p]
            assert flat == list(range(n_images))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    for k in list(os.environ):   # a launcher's rendezvous settings must not leak into this private group
        if k.startswith(("TORCHELASTIC", "TORCH_NCCL", "GROUP_", "ROLE_")) or k in ("MASTER_ADDR", "MASTER_PORT", "RANK",
                                                                                    "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE"):
            os.environ.pop(k, None)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    # what bench.py does with its per-rank device time: barrier, then MAX over ranks; shards are disjoint
    dist.barrier()
    t = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mine = list(sharding.batch_partition(10, world, rank))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, float(t.item()), gathered))
    dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_world_size_2_gloo_timing_reduction_and_disjoint_shards():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, tmax, gathered in res:
        assert tmax == 15.0
        assert sorted(i for part in gathered for i in part) == list(range(10))
