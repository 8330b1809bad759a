"""Host-side logic of the depth-sharded build on CPU: world_size-2 gloo, with the CPU oracle standing
in for the CUDA operator (the data path itself is covered by the -m gpu tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from satmvs_b200 import sharded, synth


def test_plane_range_partitions_every_plane_once():
    for D in (1, 7, 8, 48, 64, 192):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                d0, d1 = sharded.plane_range(D, r, world)
                assert 0 <= d0 <= d1 <= D
                seen += list(range(d0, d1))
            assert seen == list(range(D))
            sizes = [sharded.plane_range(D, r, world)[1] - sharded.plane_range(D, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, D, mode, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import volume
        B, V, C, H, W = 1, 3, 4, 10, 14
        fe = synth.make_features(B, V, C, H, W, seed=3)
        rp = synth.make_rpc_stack(B, V, H, W)
        dv = synth.make_depth_planes(B, D, H, W)

        def cpu_builder(ref, srcs, ref_cam, src_cams, depth, geo):
            cams = torch.stack([ref_cam] + list(src_cams), 1)
            return volume.variance_cost_volume([ref] + list(srcs), cams, depth, geo)

        got = sharded.build_cost_volume_sharded(fe[0], fe[1:], rp[:, 0], [rp[:, 1], rp[:, 2]], dv, "rpc",
                                                mode=mode, builder=cpu_builder)
        want = volume.variance_cost_volume(fe, rp, dv, "rpc")
        if mode == "none":
            d0, d1 = sharded.plane_range(D, rank, world)
            want = want[:, :, d0:d1]
        ret[rank] = bool(torch.equal(got, want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("D,mode", [(8, "nccl"), (6, "nccl"), (7, "none")])
def test_sharded_build_world2(D, mode):
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), D, mode, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _worker_reduced(rank, world, port, D, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import regress, volume
        B, V, C, H, W = 1, 3, 4, 10, 14
        fe = synth.make_features(B, V, C, H, W, seed=3)
        rp = synth.make_rpc_stack(B, V, H, W)
        dv = synth.make_depth_planes(B, D, H, W)

        def cpu_builder(ref, srcs, ref_cam, src_cams, depth, geo):
            cams = torch.stack([ref_cam] + list(src_cams), 1)
            return volume.variance_cost_volume([ref] + list(srcs), cams, depth, geo)

        depth, conf = sharded.sweep_depth_sharded(fe[0], fe[1:], rp[:, 0], [rp[:, 1], rp[:, 2]], dv, "rpc", scale=-2.0,
                                                  builder=cpu_builder, head=regress.StreamingSoftArgminState(B, H, W))
        one = regress.StreamingSoftArgminState(B, H, W)
        one.update_volume(volume.variance_cost_volume(fe, rp, dv, "rpc"), dv, -2.0)
        wd, wc = one.finish()
        # fp64 sums re-associated across ranks: equal to the last bits of the fp32 result
        ret[rank] = bool(torch.allclose(depth, wd, rtol=1e-6, atol=0) and torch.allclose(conf, wc, rtol=1e-6, atol=0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("D", [8, 7])
def test_sharded_sweep_exchanging_only_the_softargmin_sums_world2(D):
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker_reduced, args=(world, _free_port(), D, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_gather_requires_even_split():
    # uneven D cannot use the single all-gather (documented); the caller pads or uses mode "none"
    with pytest.raises(Exception):
        sharded.gather_slabs(torch.zeros(1, 1, 3, 2, 2), 7, None)
