"""World-size-2 test of the multi-GPU host logic on CPU (gloo): strided viewpoint sharding, per-rank draw-list
recording, and the host-side gather of per-frame checksums.  No collective touches the draw path."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import common
from common import drr, orc, synth_wad
from doom_rust_renderer_b200 import shard

W, H, N = 96, 64, 6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _frame_sums(path, views):
    """Per-frame checksums via the product host path (front-end + binning) replayed on the CPU."""
    ctx = drr.Context(W, H, 0, max(len(views), 1), _host_only=True)
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    assert scene.emit_views(ctx, views) == []
    return np.array([drr.checksum_numpy(common.replay_binned_frame(ctx, k)) for k in range(len(views))], np.uint64)


def _worker(rank, world, port, path, views, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.shard_indices(len(views), rank, world)
    local = _frame_sums(path, views[mine])
    allsums = shard.gather_checksums(local)  # rank order: rank 0's viewpoints, then rank 1's, ...
    counts = [len(shard.shard_indices(len(views), r, world)) for r in range(world)]
    parts = np.split(allsums, np.cumsum(counts)[:-1])
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), shard.unshard(parts, len(views)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 4096, 65536):
        for world in (1, 2, 3, 8):
            r = [shard.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_strided_shards_partition_the_batch():
    for n in (0, 1, 7, 4096, 65537):
        for world in (1, 2, 3, 8):
            parts = [shard.shard_indices(n, k, world) for k in range(world)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            if n:
                assert (shard.unshard([p * 3 for p in parts], n) == np.arange(n) * 3).all()
    with pytest.raises(ValueError):
        shard.shard_indices(10, 2, 2)


def test_two_rank_sharding_gathers_the_same_checksums(tmp_path):
    path, gm = common.wad("tiny")
    game = orc.Game(path, "E1M1", W, H)
    views = common.usable_views(game, synth_wad.scatter_viewpoints(gm, 64), N)
    want = np.array([drr.checksum_numpy(game.render(float(v[0]), float(v[1]), float(v[2]))) for v in views], np.uint64)
    port = _free_port()
    mp.start_processes(_worker, args=(2, port, path, views, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % rank))
        assert (got == want).all()
    assert shard.checksum_of_checksums(want) == shard.checksum_of_checksums(got)
