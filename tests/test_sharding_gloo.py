"""CPU, world_size 2 over gloo: the host logic of the two sharded drivers (b200gs/parallel.py).
View-sharded rendering must cover every view exactly once; an N-rank image-parallel step must leave on every
rank the same fused gradient buffer a single rank gets by accumulating all views itself."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers  # noqa: F401  (sys.path)
from b200gs import parallel

P = 257
N_VIEWS = 5


def fake_view_grads(view):
    """Deterministic stand-in for one view's rasterizer backward (the kwargs of accumulate_view)."""
    g = torch.Generator().manual_seed(1234 + view)
    r = lambda *s: torch.randn(*s, generator=g)
    radii = (torch.rand(P, generator=g) > 0.3).to(torch.int32) * torch.randint(1, 40, (P,), generator=g, dtype=torch.int32)
    return dict(d_xyz=r(P, 3), d_shs=r(P, 16, 3), d_opacity=r(P, 1), d_scaling=r(P, 3), d_rotation=r(P, 4),
                d_feature=r(P, 3), viewspace_grad=r(P, 3), radii=radii)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # view-sharded rendering: every view rendered exactly once, gathered on rank 0
        res = parallel.render_views_sharded(lambda i: (i, rank), N_VIEWS, gather=True)
        bucket = parallel.FusedGradBuffer(P, "cpu")
        parallel.image_parallel_step(fake_view_grads, list(range(N_VIEWS)), bucket)
        out[rank] = dict(views=res, flat=bucket.flat.clone().numpy(), radii=bucket.max_radii2D.clone().numpy())
    finally:
        dist.destroy_process_group()


def test_shard_views_partition():
    for n in (1, 2, 3, 8):
        for v in (0, 1, 7, 200):
            got = sorted(i for r in range(n) for i in parallel.shard_views(v, r, n))
            assert got == list(range(v))


def test_fused_buffer_layout():
    b = parallel.FusedGradBuffer(10, "cpu")
    assert b.flat.numel() == 12 * 64 and parallel.FUSED_WIDTH == 64  # segment stride = P rounded up to a multiple of 4
    # segments are contiguous, 16-byte aligned tensors the rasterizer backward can write into directly (any P)
    for name, w in parallel.SLOTS:
        seg = b.segment(name)
        assert seg.is_contiguous() and seg.numel() == 10 * w and seg.data_ptr() % 16 == 0
    assert b.grads_flat.numel() == 62 * 12 and b.stats_flat.numel() == 2 * 12
    b.segment("shs")[3, 2, 1] = 5.0
    assert b.flat.sum() == 5.0


def test_image_parallel_matches_single_rank():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    # single-rank reference: accumulate every view locally
    ref = parallel.FusedGradBuffer(P, "cpu")
    for v in range(N_VIEWS):
        ref.accumulate_view(**fake_view_grads(v))
    for r in range(world):
        np.testing.assert_allclose(out[r]["flat"], ref.flat.numpy(), rtol=1e-5, atol=1e-5)  # fp32 reassociation
        np.testing.assert_array_equal(out[r]["radii"], ref.max_radii2D.numpy())
    np.testing.assert_array_equal(out[0]["flat"], out[1]["flat"])  # identical on every rank -> identical Adam/densify
    merged = out[0]["views"]
    assert sorted(merged) == list(range(N_VIEWS))
    assert all(merged[i] == (i, i % world) for i in range(N_VIEWS))
