"""N > 1 path on CPU: world_size-2 gloo processes shard the image units, "generate" their images from the unit seeds
and all-gather them; the result must equal the single-process sweep."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_image(unit: int):
    from faceposegenerator_b200.parallel import unit_seed
    g = torch.Generator().manual_seed(unit_seed(unit, 1000))
    return torch.randint(0, 256, (8, 8, 3), generator=g, dtype=torch.uint8)


def _worker(rank, world, port, n_units, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from faceposegenerator_b200.parallel import gather_images, shard_units
        mine = shard_units(n_units, rank, world)
        local = torch.stack([_fake_image(u) for u in mine]) if mine else torch.zeros((0, 8, 8, 3), dtype=torch.uint8)
        full = gather_images(local, n_units, rank, world)
        # the asynchronous double-buffered gather of bench.py: three submissions, each result read one call later
        from faceposegenerator_b200.parallel import ImageGather, unit_index
        ig = ImageGather(n_units, rank, world, (8, 8, 3), "cpu")
        rounds = []
        for k in range(3):
            ig.submit((local.int() + k).clamp(max=255).to(torch.uint8))
            if k > 0:
                assert ig.last is not None
        buf = ig.wait()
        q.put((rank, mine, full.clone(), buf[unit_index(n_units, world)].clone()))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(n_units, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return results


def test_shard_units_partition():
    from faceposegenerator_b200.parallel import shard_units
    for n, g in ((1024, 8), (7, 2), (3, 4), (0, 2)):
        parts = [shard_units(n, r, g) for r in range(g)]
        assert sorted(u for p in parts for u in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_gloo_sweep_matches_single_process():
    for n_units in (6, 7):   # even and ragged
        ref = torch.stack([_fake_image(u) for u in range(n_units)])
        res = _run(n_units)
        ranks = sorted(r[0] for r in res)
        assert ranks == [0, 1]
        for rank, mine, full, last_async in res:
            assert mine == list(range(rank, n_units, 2))
            assert full.shape == ref.shape and torch.equal(full, ref)
            assert torch.equal(last_async, (ref.int() + 2).clamp(max=255).to(torch.uint8))   # third submission, unit order
