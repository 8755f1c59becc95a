#!/usr/bin/env python
"""The other BASELINE.json configurations (bench.py measures configs[1]); run plain or under torchrun.

  python tools/workloads.py vol512   [--steps 50]            config 3: one full 512x512 slab, encode + DDIM-50 + decode
  torchrun ... tools/workloads.py ddpm1000 [--batch 32]      config 4: DDPM-1000, batch 32 sharded over the ranks
  torchrun ... tools/workloads.py sweep [--volumes 64]       config 5: V full 512x512 slabs, 25 patches each, sliding-
                                                             window stitching per volume, NCCL gather of the results
Prints one JSON line on rank 0 (device-timed, max over ranks).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from v2v_b200.dist import gather_slabs, shard_range  # noqa: E402
from v2v_b200.inference.volume import generate_volume, window_starts  # noqa: E402
from v2v_b200.models import VideoToVideoDiffusion  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["vol512", "ddpm1000", "sweep"])
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--volumes", type=int, default=64)
    ap.add_argument("--ddpm-timesteps", type=int, default=1000)
    a = ap.parse_args()
    rank, world, lr = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device(f"cuda:{lr}")
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    cfg = bench.load_cfg()
    if a.what == "ddpm1000":
        cfg = dict(cfg, diffusion_timesteps=a.ddpm_timesteps)
    torch.manual_seed(0)
    m = VideoToVideoDiffusion(cfg).eval().to(dev)
    g = torch.Generator().manual_seed(1234 + rank)
    torch.manual_seed(42 + rank)

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, out

    if a.what == "vol512":
        v = (torch.rand((1, 1, 8, 512, 512), generator=g) * 2 - 1).to(dev)
        m.generate(v, "ddim", 2, target_depth=48)  # plan + capture
        ms, out = timed(lambda: m.generate(v, "ddim", a.steps, target_depth=48))
        res = {"workload": "config 3: (1,1,8,512,512) -> (1,1,48,512,512), encode + DDIM-%d + decode" % a.steps,
               "seconds": ms / 1e3, "volumes_per_s": 1e3 / ms, "algorithmic_tflop": 1781.2 * (a.steps + 1) / 51,
               "out_shape": list(out.shape)}
    elif a.what == "ddpm1000":
        lo, hi = shard_range(a.batch, rank, world)
        b = hi - lo
        v = (torch.rand((b, 1, 8, 192, 192), generator=g) * 2 - 1).to(dev)
        m.generate(v, "ddim", 2, target_depth=48)  # one-time weight repack, planning and graph capture
        ms, out = timed(lambda: gather_slabs(m.generate(v, "ddpm", target_depth=48)))
        res = {"workload": f"config 4: DDPM-{a.ddpm_timesteps}, batch {a.batch} x (1,1,8,192,192) over {world} GPU(s) "
                           f"({b} per rank) + decode + gather", "seconds": ms / 1e3,
               "patch_volumes_per_s": a.batch * 1e3 / ms, "out_shape": list(out.shape)}
    else:
        lo, hi = shard_range(a.volumes, rank, world)  # whole volumes per rank: stitching stays local
        vols = (torch.rand((hi - lo, 1, 8, 512, 512), generator=g) * 2 - 1).to(dev)
        n_win = len(window_starts(8, 512, 512))
        m.generate(vols[:1, :, :, :192, :192].repeat(4, 1, 1, 1, 1).contiguous(), "ddim", 2, target_depth=48)

        def run():
            outs = [generate_volume(m, vols[i:i + 1], "ddim", a.steps, batch=4) for i in range(hi - lo)]
            local = torch.cat(outs) if outs else torch.empty((0, 1, 48, 512, 512), device=dev)
            counts = [shard_range(a.volumes, r, world)[1] - shard_range(a.volumes, r, world)[0] for r in range(world)]
            return gather_slabs(local, counts if len(set(counts)) > 1 else None)
        ms, out = timed(run)
        res = {"workload": f"config 5: {a.volumes} x (1,1,8,512,512), {n_win} windows each, DDIM-{a.steps}, "
                           f"stitched per volume, gathered over {world} GPU(s)", "seconds": ms / 1e3,
               "volumes_per_s": a.volumes * 1e3 / ms, "patch_volumes_per_s": a.volumes * n_win * 1e3 / ms,
               "out_shape": list(out.shape)}
    if rank == 0:
        res["n_gpus"] = world
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
