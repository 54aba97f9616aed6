#!/usr/bin/env python
"""Multi-GPU correctness check (run under torchrun on N GPUs of one node):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 tools/multi_gpu_check.py

Every rank renders its share of the samples (spp partition) or of the rows (tile partition) of the same
scene with the same seed, the accumulation buffers are merged with one NCCL all-reduce, and rank 0 compares
the merged image with a single-GPU render of all samples: the tile partition must be bit-identical, the spp
partition identical up to float summation order.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crender_b200 import api, scenes  # noqa: E402
from crender_b200 import distributed as D  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    desc = scenes.mesh_scene(200, 100)
    w, h, spp, bounces = 640, 360, 16, 6
    g = api.scene(device=local)
    scenes.load(desc, g)
    g.commit()
    results = {}
    for partition in ("spp", "tile"):
        r = api.renderer(w, h, bounces, g, seed=4)
        if partition == "spp":
            lo, hi = D.sample_range(rank, world, spp)
            r.render(hi - lo, first_sample=lo)
        else:
            for y0, y1 in D.row_bands(rank, world, h, band=32):
                r.set_rows(y0, y1)
                r.render(spp, first_sample=0)
            r.set_rows(0, h)
        passes = D.merge_renderer(r, partition, passes_local=None if partition == "spp" else spp)
        results[partition] = (r.raw_sum().copy(), r.current_progress().copy(), passes)
        del r
    if rank == 0:
        r = api.renderer(w, h, bounces, g, seed=4)
        r.render(spp)
        ref_raw, ref_disp = r.raw_sum(), r.current_progress()
        ok = True
        for partition, (raw, disp, passes) in results.items():
            d = np.abs(raw[..., :3] - ref_raw[..., :3])
            rel = float(np.sqrt(np.mean(d**2)) / np.sqrt(np.mean(ref_raw[..., :3] ** 2)))
            exact = bool(np.array_equal(raw[..., :3], ref_raw[..., :3]))
            disp_err = float(np.abs(disp - ref_disp).max())
            print(f"{partition}: world {world} passes {passes} relRMSE vs 1-GPU {rel:.3e} bit-identical {exact} max display diff {disp_err:.3e}")
            ok &= passes == spp and rel < 1e-5 and (exact or partition == "spp")
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
