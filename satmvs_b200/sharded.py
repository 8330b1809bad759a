"""Depth-sharded cost-volume build across the GPUs of one node.

Every (d, h, w) cell of the sweep is independent, so rank r builds planes [r*D/G, (r+1)*D/G) of
var[B, C, D, H, W]; features and cameras are replicated.  Reassembly:

  mode "nccl"   one `all_gather_into_tensor` of the slabs (the north_star baseline) into a
                [G, B, C, D/G, H, W] buffer, then a device copy into the reference's [B, C, D, H, W] layout;
  mode "fused"  the sweep kernel itself stores every value into each peer's full volume through
                peer-mapped pointers (torch symmetric memory = CUDA IPC over NVLink), so the gather
                overlaps the sweep and needs no second pass; one barrier at the end;
  mode "multimem"  like "fused", but through the NVLink-switch multicast mapping of the symmetric volume (NVLS): every
                value leaves the SM once as a `multimem.st` and the switch replicates it, so egress is 1x instead of (G-1)x;
  mode "none"   no exchange: returns the local slab (build-only scaling).

The "fused" / "multimem" modes return the cached symmetric-memory volume itself: a BORROWED buffer, valid until the next
call with the same shape on the same group (clone it to keep it).

The plumbing is torch.distributed (NCCL on GPUs, gloo in the CPU tests); the data path is the C ABI's
`satmvs_cost_volume_*_fwd_sharded`.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .warping import _check_cam, _depth_arg, _dptr, build_cost_volume, host_f64, sweep_workspace


def plane_range(D: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous planes of `rank`; D need not divide evenly (the first D % world ranks take one more)."""
    base, extra = divmod(D, world)
    d0 = rank * base + min(rank, extra)
    return d0, d0 + base + (1 if rank < extra else 0)


def gather_slabs(slab: torch.Tensor, D: int, group=None) -> torch.Tensor:
    """All-gather per-rank slabs [B, C, Dl, H, W] (equal Dl) into the reference layout [B, C, D, H, W]."""
    world = dist.get_world_size(group)
    B, C, Dl, H, W = slab.shape
    if Dl * world != D:
        raise ValueError("gather_slabs needs D divisible by the world size")
    buf = torch.empty((world * B, C, Dl, H, W), dtype=slab.dtype, device=slab.device)   # rank-major concatenation
    dist.all_gather_into_tensor(buf, slab.contiguous(), group=group)
    # [G, B, C, Dl, H, W] -> [B, C, G*Dl, H, W]
    return buf.view(world, B, C, Dl, H, W).permute(1, 2, 0, 3, 4, 5).reshape(B, C, D, H, W)


_SYMM_CACHE: dict = {}


def _symmetric_volume(shape, device, group):
    """A [B, C, D, H, W] volume every rank can write into: returns (local tensor, peer pointers, handle)."""
    import torch.distributed._symmetric_memory as symm_mem
    grp = group if group is not None else dist.group.WORLD
    key = (tuple(shape), device.index, grp.group_name)       # the group's name, not id(): ids are recycled after destroy
    hit = _SYMM_CACHE.get(key)
    if hit is None:
        t = symm_mem.empty(shape, dtype=torch.float32, device=device)
        hdl = symm_mem.rendezvous(t, group=grp)
        hit = (t, [int(p) for p in hdl.buffer_ptrs], hdl)
        _SYMM_CACHE[key] = hit
    return hit


def build_cost_volume_sharded(ref_fea, src_feas, ref_cam, src_cams, depth_values, geo_model="rpc", *,
                              group=None, mode="nccl", builder=None):
    """Depth-sharded `build_cost_volume`.  Returns the full [B, C, D, H, W] volume on every rank
    (modes "nccl", "fused") or this rank's slab (mode "none").  `builder` replaces the CUDA operator
    in the CPU (gloo) tests of the host logic."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    D = depth_values.shape[1]
    d0, d1 = plane_range(D, rank, world)
    shard = depth_values[:, d0:d1].contiguous()
    if mode not in ("fused", "multimem"):
        fn = builder or build_cost_volume
        slab = fn(ref_fea, src_feas, ref_cam, src_cams, shard, geo_model)
        if mode == "none" or world == 1:
            return slab
        if mode != "nccl":
            raise ValueError(f"unknown mode {mode!r}")
        return gather_slabs(slab, D, group)

    # fused: the kernel writes its planes into every peer's volume
    ref = _lib.require_cuda(ref_fea, "ref_fea")
    srcs = [_lib.require_cuda(s, "src_fea") for s in src_feas]
    B, C, H, W = ref.shape
    if isinstance(src_cams, torch.Tensor):
        src_cams = list(torch.unbind(src_cams, 1))
    r = _check_cam(host_f64(ref_cam), B, geo_model, "ref camera")
    s = np.stack([_check_cam(host_f64(c), B, geo_model, "src camera") for c in src_cams])
    depth, per_pixel = _depth_arg(shard, B, H, W)
    vol, peers, hdl = _symmetric_volume((B, C, D, H, W), ref.device, group)
    mc = 0
    if mode == "multimem":
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        if mc == 0:
            raise RuntimeError("mode='multimem' needs an NVLS multicast mapping of the symmetric volume "
                               "(NVSwitch + multicast support); use mode='fused'")
    fn = _lib.lib().satmvs_cost_volume_rpc_fwd_sharded if geo_model == "rpc" else _lib.lib().satmvs_cost_volume_homo_fwd_sharded
    per_b = C * D * H * W * 4
    ws = sweep_workspace(len(srcs) * C * H * W * 4, ref.device)
    hdl.barrier()                      # nobody is still reading the previous contents
    with torch.cuda.device(ref.device):
        st = _lib.stream_ptr(ref.device)
        for b in range(B):
            outs = _lib.ptr_array([mc + b * per_b] if mc else [p + b * per_b for p in peers])
            cams = np.ascontiguousarray(s[:, b])
            _lib.check(fn(ref[b].data_ptr(), _lib.ptr_array([x[b].data_ptr() for x in srcs]), len(srcs), _dptr(r[b]),
                          _dptr(cams), depth[b].data_ptr(), per_pixel, C, d1 - d0, H, W, d0, D, outs, -1 if mc else len(peers),
                          ws.data_ptr(), ws.numel(), st),
                       "cost_volume_fwd_sharded")
    hdl.barrier()                      # every rank's planes have landed everywhere
    return vol


def combine_stream_states(state: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce of the streaming soft-argmin state [B, 3, H, W] fp64 of depth-sharded planes: (sum e, sum d*e) add,
    max e takes the maximum -- 24 bytes per pixel on the wire instead of 4 * C * D / G bytes per pixel of variance slab."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return state
    sums = state[:, :2].contiguous()
    mx = state[:, 2].contiguous()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    state[:, :2] = sums
    state[:, 2] = mx
    return state


def sweep_depth_sharded(ref_fea, src_feas, ref_cam, src_cams, depth_values, geo_model="rpc", *, group=None, scale=-1.0,
                        builder=None, head=None):
    """Depth-sharded plane sweep WITHOUT a regulariser: every rank sweeps its planes, folds the matching cost
    reg_k = scale * mean_c var[c, k] of its slab into the fp64 streaming soft-argmin sums (`casred.py:218-236` arithmetic) and
    the ranks exchange only those sums.  Returns (depth [B,H,W], confidence [B,H,W]) on every rank.
    `builder` / `head` replace the CUDA operators in the CPU (gloo) tests of the host logic."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    D = depth_values.shape[1]
    d0, d1 = plane_range(D, rank, world)
    shard = depth_values[:, d0:d1].contiguous()
    slab = (builder or build_cost_volume)(ref_fea, src_feas, ref_cam, src_cams, shard, geo_model)
    B, _, _, H, W = slab.shape
    if head is None:
        from .regress import StreamingSoftArgmin
        head = StreamingSoftArgmin(B, H, W, slab.device)
    head.update_volume(slab, shard, scale)
    combine_stream_states(head.state, group)
    return head.finish()
