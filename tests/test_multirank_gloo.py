"""N > 1 host logic on the CPU: two gloo ranks play two GPUs.  Each rank owns interleaved rows,
counts its rows' hit nodes, the counts are all_gathered (the one exchange of the reference-LCG
mode), every rank derives its rows' AO-stream bases, renders its rows, and rank 0 gathers the
bands.  The renderer here is the T1 oracle (no GPU in this container); the partition / prefix /
gather code is the product's (580-raytracer_b200/__init__.py), the same bench.py drives."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ASSETS, ROOT, load_golden

W, H, SPP, DEPTH, SCENE = 64, 46, 3, 4, "simpleSphereScene.json"


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    import oracle
    pkg = load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle.Oracle(oracle.load_scene_json(ASSETS, SCENE))
    first, step, n = pkg.rows_for_rank(H, rank, world)
    rows = np.arange(first, first + n * step, step)
    pix = (rows[:, None] * W + np.arange(W)[None, :]).reshape(-1).astype(np.int32)
    # structure pass: hit nodes per owned pixel (no AO rays needed)
    _, _, hits = orc.render(W, H, 1, DEPTH, pix=pix, ao_base=np.zeros(pix.size, np.uint64))
    counts = hits.reshape(n, W).sum(axis=1).astype(np.int64)
    max_rows = (H + world - 1) // world
    mine = torch.zeros(max_rows, dtype=torch.int64)
    mine[:n] = torch.from_numpy(counts)
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    bases = pkg.row_bases_from_counts(H, world, [c.numpy().astype(np.uint64) for c in allc])[rank]
    within = np.cumsum(hits.reshape(n, W), axis=1) - hits.reshape(n, W)
    ao_base = (bases[:, None] + within.astype(np.uint64)).reshape(-1)       # one ambient light in this scene
    band, rays, _ = orc.render(W, H, SPP, DEPTH, pix=pix, ao_base=ao_base)
    padded = torch.zeros((max_rows, W, 3), dtype=torch.int16)
    padded[:n] = torch.from_numpy(band.reshape(n, W, 3))
    # int16 is not a collective dtype (neither gloo nor NCCL): the bands travel as bytes
    pbytes = padded.view(torch.uint8)
    gl = [torch.zeros_like(pbytes) for _ in range(world)] if rank == 0 else None
    dist.gather(pbytes, gl, dst=0)
    tr = torch.tensor([rays], dtype=torch.int64)
    dist.all_reduce(tr)
    if rank == 0:
        fb = pkg.interleave_rows(H, W, world, [g.view(torch.int16).numpy() for g in gl])
        q.put((fb, int(tr[0])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_ranks_reproduce_the_single_stream(oracle, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    fb, rays = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    orc = oracle.Oracle(oracle.load_scene_json(ASSETS, SCENE))
    ref, ref_rays, _ = orc.render(W, H, SPP, DEPTH, nthreads=1)      # one engine, scanline order
    assert np.array_equal(fb, ref)
    assert rays == ref_rays
