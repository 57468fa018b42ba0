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
from exavatar_release_b200.sharding import (GradBucket, reduce_densify_stats, shard_frames, sharded_step,
                                            split_noise_generator)
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


def _stats_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)  # every rank saw different frames
    stats = {"grad_accum": torch.rand(50, generator=g), "count": torch.randint(0, 3, (50,), generator=g).float(),
             "radius_max": torch.randint(0, 30, (50,), generator=g).float()}
    local = {k: v.clone() for k, v in stats.items()}
    reduce_densify_stats(stats)
    noise = torch.normal(torch.zeros(7, 3), torch.ones(7, 3), generator=split_noise_generator(step=1234))
    ret[rank] = ({k: v.numpy() for k, v in local.items()}, {k: v.numpy() for k, v in stats.items()}, noise.numpy())
    dist.destroy_process_group()


def test_densify_statistics_reduce_sum_sum_max_and_split_noise_is_identical_on_every_rank():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_stats_worker, args=(2, port, ret), nprocs=2, join=True)
    (l0, r0, n0), (l1, r1, n1) = ret[0], ret[1]
    assert np.allclose(r0["grad_accum"], l0["grad_accum"] + l1["grad_accum"]) and np.array_equal(r0["grad_accum"], r1["grad_accum"])
    assert np.array_equal(r0["count"], l0["count"] + l1["count"]) and np.array_equal(r0["count"], r1["count"])
    assert np.array_equal(r0["radius_max"], np.maximum(l0["radius_max"], l1["radius_max"]))
    assert np.array_equal(n0, n1)  # same step -> same split offsets on both replicas
    other = torch.normal(torch.zeros(7, 3), torch.ones(7, 3), generator=split_noise_generator(step=1235)).numpy()
    assert not np.array_equal(n0, other)
