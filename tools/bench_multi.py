"""BASELINE configs 3 and 5 across N GPUs (strong scaling: the image and its spp are fixed, the
work is split by tiles or by sample index, volume replicated, ONE NCCL all-reduce of the
framebuffer inside the timed region).  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P tools/bench_multi.py [c3] [c5]
Rank 0 prints one JSON line per (config, mode)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cudavolumerenderer_b200 as cvr  # noqa: E402
from cudavolumerenderer_b200.distributed import render_sharded  # noqa: E402

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    saved = os.dup(1)  # NCCL's banner goes to stderr, stdout carries JSON lines only
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.all_reduce(torch.zeros(1, device=dev))
    torch.cuda.synchronize(dev)
    os.dup2(saved, 1)
    os.close(saved)
which = [a for a in sys.argv[1:]] or ["c3", "c5"]
CONFIGS = {
    "c3": ("C3 manix 256x230x256, 1024x1024, 256 spp, 10x10 tiles, regenerationSK", lambda: cvr.scenes.manix(), 1024, 256, (10, 10)),
    "c5": ("C5 sparse 2048^3, 4096x4096, 16 of 1024 spp, 8x8 tiles, regenerationSK", lambda: cvr.scenes.sparse_fbm(2048), 4096, 16, (8, 8)),
}
for key in which:
    name, mk, res, spp, tiles = CONFIGS[key]
    sc = mk()
    kl = cvr.createLauncher("regenerationSK", local)
    stream = torch.cuda.current_stream(dev)
    kl.setStream(stream.cuda_stream)
    kl.setScene(sc)
    d_img = torch.zeros((res, res, 4), dtype=torch.float32, device=dev)
    for mode in ("tiles", "spp", "balanced"):
        def step():
            kl.setSeed(0)
            render_sharded(kl, (res, res), tiles, spp, mode, d_img, fov_x=sc.fov_x)
        for _ in range(2):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record(stream)
        for _ in range(reps):
            step()
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        tw = (res // tiles[0]) * tiles[0]
        th = (res // tiles[1]) * tiles[1]
        paths = tw * th * spp  # tile_dim floors (Q6): the remainder pixels are never rendered
        if rank == 0:
            print(json.dumps({"config": name, "n_gpus": world, "sharding": mode, "scaling": "strong", "ms_per_render": float(ms.item()),
                              "msamples_per_s": paths / float(ms.item()) / 1e3, "paths": paths,
                              "image_mean": float(torch.nanmean(d_img[..., :3]).item()),
                              "options": {k: kl.getOption(k) for k in ("sched", "warp_slots", "skip", "tracking")}}), flush=True)
    kl.close()
    del d_img
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
