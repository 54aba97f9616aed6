"""N>1 host logic on CPU: world_size-2 gloo. The compute stand-in is the oracle (allowed in tests); what
is under test is crender_b200.distributed: the partition of the sample range / row bands and the merge."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crender_b200 import distributed as D


def test_sample_range_partitions_exactly():
    for world in (1, 2, 3, 4, 8):
        for n in (1, 7, 16, 255, 1024):
            got = []
            for r in range(world):
                lo, hi = D.sample_range(r, world, n, first_sample=5)
                got.extend(range(lo, hi))
            assert got == list(range(5, 5 + n))


def test_row_bands_cover_once():
    for world in (1, 2, 4, 8):
        for h in (1, 63, 64, 1080, 2160):
            rows = np.zeros(h, int)
            for r in range(world):
                for y0, y1 in D.row_bands(r, world, h):
                    rows[y0:y1] += 1
            assert np.all(rows == 1)


def _worker(rank, world, port, partition, out_dir):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import oracle_binding as ob
    from crender_b200 import scenes

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    desc = scenes.cornell()
    s = ob.scene()
    scenes.load(desc, s)
    s.commit()
    w, h, spp = 24, 20, 6
    r = ob.renderer(w, h, 4, s, seed=1)
    if partition == "spp":
        lo, hi = D.sample_range(rank, world, spp)
        r.render(hi - lo, first_sample=lo, nthreads=1)
        passes_local = hi - lo
    else:
        for y0, y1 in D.row_bands(rank, world, h, band=4):
            r.set_rows(y0, y1)
            r.render(spp, first_sample=0, nthreads=1)
        passes_local = spp
    raw = torch.from_numpy(r.raw_sum()[..., :3].copy())
    passes = D.merge_partial_sums(raw, passes_local, partition)
    if rank == 0:
        np.save(os.path.join(out_dir, f"{partition}.npy"), raw.numpy())
        np.save(os.path.join(out_dir, f"{partition}_passes.npy"), np.asarray([passes]))
    dist.destroy_process_group()


@pytest.mark.parametrize("partition", ["spp", "tile"])
def test_two_rank_merge_equals_single(oracle, tmp_path, partition):
    from crender_b200 import scenes

    port = 29650 + (os.getpid() % 200) + (0 if partition == "spp" else 1)
    mp.spawn(_worker, args=(2, port, partition, str(tmp_path)), nprocs=2, join=True)
    merged = np.load(tmp_path / f"{partition}.npy")
    passes = int(np.load(tmp_path / f"{partition}_passes.npy")[0])
    s = oracle.scene()
    scenes.load(scenes.cornell(), s)
    s.commit()
    r = oracle.renderer(24, 20, 4, s, seed=1)
    r.render(6, nthreads=1)
    single = r.raw_sum()[..., :3]
    assert passes == 6
    if partition == "tile":
        np.testing.assert_array_equal(merged, single)  # disjoint rows: sums are untouched
    else:
        np.testing.assert_allclose(merged, single, rtol=1e-5, atol=1e-6)  # float summation order across ranks
