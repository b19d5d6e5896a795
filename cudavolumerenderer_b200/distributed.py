"""Multi-GPU plumbing: one process per GPU, volume replicated, work sharded by image
tiles or by sample index, ONE collective per render (sum of the resolved framebuffers)
through torch.distributed (NCCL on GPUs; gloo in the CPU tests).  SURVEY.md section 8(e).

Sharding rules (the library's planner, cvr_shard_plan; pure functions, tested on CPU):
  tiles:    rank r renders tiles k = r, r+G, ... of the reference's tile list -> disjoint
            pixels, the image does not depend on G.
  spp:      rank r renders sample indices [first, first+count) of every pixel; stream ids
            are seed + sample*npix + pixel, so the multiset of paths equals the 1-GPU run.
  balanced: equal work on every rank: the sample split when iterations is a multiple of G (one launch, the
            same number of paths through the same pixels on every rank); otherwise the complete rounds of the
            tile interleave by tile and the n_tiles mod G left-over tiles by sample index.
Every rank resolves with scale = TOTAL iterations, so the sum over ranks is the image (rgb; the alpha channel,
"some path of this pixel escaped" / iterations, is clamped back to that value after the sum).
The one-process form of the same thing (N devices, one host thread each, ncclReduce called by
the library itself) is cudavolumerenderer_b200.DeviceGroup / cvr_group_render_image.
"""
from __future__ import annotations


def spp_shard(total_spp: int, rank: int, world: int) -> tuple[int, int]:
    """(first, count) of the sample indices rank `rank` renders; remainders go to the
    low ranks so that counts differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(total_spp, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def tile_shard(n_tiles: int, rank: int, world: int) -> list[int]:
    """Global tile indices of rank `rank` (interleaved: k = rank (mod world))."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return list(range(rank, n_tiles, world))


def render_sharded(launcher, res, n_tiles, iterations: int, mode: str, d_image, fov_x: float = 0.7,
                   inv_view=None, fuse_tiles: bool = True, reduce_to: int | None = None):
    """Render this rank's share into the torch tensor `d_image` (H, W, 4 float32, on the
    launcher's device; zero-filled here first) and combine the ranks' images: all-reduce, or
    -- reduce_to = rank -- a reduce to the one rank that needs the image (half the traffic).
    Returns d_image (complete on every rank / on rank `reduce_to`)."""
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    d_image.zero_()
    if d_image.is_cuda:
        # the zero fill is on torch's current stream; the render must not overtake it.  Ordered by
        # the stream itself when the launcher was given that stream (setStream), by a host-side
        # wait otherwise (the launcher's own stream is non-blocking).
        import torch

        cur = torch.cuda.current_stream(d_image.device)
        if launcher.streamPtr() != (cur.cuda_stream or 1):
            cur.synchronize()
    kw = dict(fov_x=fov_x, inv_view=inv_view, d_image=d_image.data_ptr(), host_image=None, fuse_tiles=fuse_tiles)
    if mode == "spp":
        first, count = spp_shard(iterations, rank, world)
        if count:
            launcher.renderImage(res, n_tiles, iterations, sample_first=first, sample_count=count, **kw)
    elif mode == "tiles":
        launcher.renderImage(res, n_tiles, iterations, tile_first=rank, tile_stride=world, **kw)
    elif mode == "balanced":
        from . import abi

        sh = abi.shard_plan(n_tiles[0] * n_tiles[1], iterations, rank, world, "balanced")
        launcher.renderImageSharded(res, n_tiles, iterations, sh, fov_x=fov_x, inv_view=inv_view, fuse_tiles=fuse_tiles,
                                    d_image=d_image.data_ptr())
    else:
        raise ValueError("mode must be 'spp', 'tiles' or 'balanced'")
    if world > 1:
        if reduce_to is None:
            dist.all_reduce(d_image)
        else:
            dist.reduce(d_image, dst=reduce_to)
        if reduce_to is None or reduce_to == rank:
            # alpha is not a sum: an escaping path STORES w = 1 (Q12) and the resolve divides by the iteration count, so
            # every sample-sharded rank contributes 1 / iterations where one of ITS paths escaped (k_clamp_alpha in the
            # library's own group path); min(sum, 1 / iterations) is what one device produces
            import numpy as np

            d_image[..., 3].clamp_(max=float(np.float32(1.0) / np.float32(iterations)))
    return d_image
