"""world_size-2 gloo test of the frame-sharding path: two CPU processes each render their share
of a 4-frame batch (the oracle stands in for the rasteriser), sum into one flat bucket, all-reduce once; the result
must equal the single-process gradient of the batch-mean loss."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from exavatar_release_b200.camera import look_at_cam_param
from exavatar_release_b200.renderer import GaussianRenderer
from exavatar_release_b200.sharding import GradBucket, shard_frames, sharded_step
from exavatar_release_b200.synthetic import make_assets, make_grad_image

KEYS = ("mean_3d", "scale", "rotation", "opacity", "rgb", "mean_2d")


def _frame_grads(assets, yaw, scale):
    from oracle import oracle as O
    leaves = {k: v.clone().requires_grad_() for k, v in assets.items()}
    r = GaussianRenderer(rasterizer_cls=O.OracleRasterizer, settings_cls=O.OracleSettings)
    out = r(leaves, (64, 64), look_at_cam_param(yaw, (64, 64)), torch.ones(3))
    loss = (out["img"] * make_grad_image("T0", 0)).sum() * scale
    loss.backward()
    g = {k: leaves[k].grad for k in leaves}
    g["mean_2d"] = out["mean_2d"].grad
    return g


def _shapes(assets):
    s = {k: tuple(v.shape) for k, v in assets.items()}
    s["mean_2d"] = (assets["mean_3d"].shape[0], 3)
    return s


def _worker(rank, world, port, yaws, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    assets = make_assets("T0", seed=11)
    bucket = GradBucket(_shapes(assets))
    flat = sharded_step(yaws, lambda yaw, sc: _frame_grads(assets, yaw, sc), bucket, rank, world)
    ret[rank] = flat.clone().numpy()
    dist.destroy_process_group()


def test_shard_frames_round_robin():
    assert shard_frames(8, 0, 2) == [0, 2, 4, 6] and shard_frames(8, 1, 2) == [1, 3, 5, 7]
    assert shard_frames(5, 3, 4) == [3] and shard_frames(2, 3, 4) == []
    assert sorted(sum((shard_frames(8, r, 8) for r in range(8)), [])) == list(range(8))
    with pytest.raises(ValueError):
        shard_frames(4, 2, 2)


def test_bucket_views_alias_flat_storage():
    b = GradBucket({"a": (3, 2), "b": (4,)})
    b.views["a"].fill_(1.0)
    b.views["b"].fill_(2.0)
    assert b.flat.tolist() == [1.0] * 6 + [2.0] * 4
    b.add_({"a": torch.ones(3, 2), "missing": None})
    assert b.views["a"].sum() == 12.0


def test_two_ranks_equal_single_process():
    yaws = [-15.0, -5.0, 5.0, 15.0]
    assets = make_assets("T0", seed=11)
    single = GradBucket(_shapes(assets))
    ref = sharded_step(yaws, lambda yaw, sc: _frame_grads(assets, yaw, sc), single, 0, 1).clone().numpy()
    assert np.abs(ref).max() > 0

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, yaws, ret), nprocs=2, join=True)
    assert np.array_equal(ret[0], ret[1])  # every rank holds the same reduced bucket
    assert np.allclose(ret[0], ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max())
