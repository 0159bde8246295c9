"""N > 1 host logic on CPU: two gloo ranks run the bench's progression split and framebuffer reduce (corona-13_b200/progressive.py)
on synthetic per-progression framebuffers.  Checks: the ranks' index ranges tile [0, K*N*P) exactly like a 1-GPU run's
progressions, and rank 0 ends up with the sum over ALL progressions while the other rank ends up with nothing to report."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

P = importlib.import_module("corona-13_b200.progressive")
H, W, K, PATHS = 8, 16, 5, 1000


def fake_progression(first, count):
    """stand-in for cb200_render_pass: a framebuffer that is a pure function of the index range (like the real paths)"""
    g = torch.Generator().manual_seed(int(first) % (2**31))
    return torch.rand(H, W, 3, generator=g) * (count / PATHS)


def worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    red = P.FramebufferReducer(H, W, "cpu", rank, world, dist)
    ranges = []
    for s in range(K):
        first, count = P.progression_range(s, rank, world, PATHS)
        ranges.append((first, count))
        fb = red.acquire(s)
        assert float(fb.abs().sum()) == 0.0, "buffer handed out before its reduce was retired and cleared"
        fb.add_(fake_progression(first, count))
        red.submit(s)
    total = red.finish()
    # the outlier rejection cascade: per-rank level buffers, summed once at export
    lv = P.reduce_dbor(torch.stack([fake_progression(1000 * rank + l, PATHS) for l in range(3)]), rank, world, dist)
    assert (lv is None) == (rank != 0)
    if lv is not None:
        want = sum(torch.stack([fake_progression(1000 * r + l, PATHS) for l in range(3)]) for r in range(world))
        assert torch.allclose(lv, want)
    out[rank] = (ranges, None if total is None else total.clone().numpy())
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(120)
def test_two_ranks_tile_the_progressions_and_sum_on_root():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(worker, args=(world, free_port(), out), nprocs=world, join=True)
        res = dict(out)
    firsts = sorted(f for r in range(world) for f, _ in res[r][0])
    assert firsts == [k * PATHS for k in range(K * world)]            # progressions 0..K*N-1, each exactly once
    assert res[1][1] is None
    want = sum(fake_progression(k * PATHS, PATHS) for k in range(K * world)).numpy()
    assert np.allclose(res[0][1], want, rtol=1e-6, atol=1e-6)


def test_single_rank_is_a_plain_accumulator():
    red = P.FramebufferReducer(H, W, "cpu")
    for s in range(3):
        first, count = P.progression_range(s, 0, 1, PATHS)
        assert first == s * PATHS
        red.acquire(s).add_(fake_progression(first, count))
        red.submit(s)
    want = sum(fake_progression(k * PATHS, PATHS) for k in range(3))
    assert torch.allclose(red.finish(), want)
