"""CPU, world_size 2 (gloo): the row partition + all_gather of output tiles that bench.py uses for N>1."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    from panogrf_b200 import sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, W = bench.H, bench.W
    rows = sharded.row_block(H, rank, world)
    que, ref = bench.make_inputs(torch, rows)
    coords = que["coords"][0]
    # every rank gets a contiguous block of rows, (x,y) pixel order, identical source maps
    assert coords.shape[0] == (rows[1] - rows[0]) * W
    assert float(coords[0, 1]) == rows[0] and float(coords[-1, 1]) == rows[1] - 1 and float(coords[-1, 0]) == W - 1
    assert torch.equal(sharded.shard_coords(H, W, rank, world), que["coords"])
    ref_sum = torch.tensor([float(ref["imgs"].sum())])
    lst = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(lst, ref_sum)
    assert all(float(x) == float(lst[0]) for x in lst)
    # stand-in for the rendered tile: (r, g, b, depth) as functions of the pixel id; ONE gather reproduces the full image
    pid = coords[:, 1] * W + coords[:, 0]
    tile = sharded.pack_tile(torch.stack([pid, 2 * pid, 3 * pid], -1)[None], (4 * pid)[None])
    assert tile.shape == (coords.shape[0], 4)
    full = sharded.gather_tiles(tile, H, W)
    want = torch.arange(H * W, dtype=torch.float32)[:, None] * torch.tensor([1.0, 2.0, 3.0, 4.0])
    ok = torch.equal(full, want)
    # ragged partition (H not divisible by the world size): blocks differ by one row, padded for the collective
    Hr = 7
    r0, r1 = sharded.row_block(Hr, rank, world)
    c = sharded.shard_coords(Hr, 8, rank, world)[0]
    t2 = (c[:, 1] * 8 + c[:, 0])[:, None].repeat(1, 4).contiguous()
    ok = ok and torch.equal(sharded.gather_tiles(t2, Hr, 8)[:, 0], torch.arange(Hr * 8, dtype=torch.float32))
    # cost-volume items: dealt without overlap, every item exactly once
    b0, b1 = sharded.item_block(8, rank, world)
    cover = torch.zeros(8)
    cover[b0:b1] = 1
    dist.all_reduce(cover)
    ok = ok and bool((cover == 1).all())
    vol, blk = sharded.cost_volume_sharded(lambda im, tr, ro: im * 2, torch.arange(8.0)[:, None], torch.zeros(8, 1), torch.zeros(8, 1))
    ok = ok and blk == (b0, b1) and torch.equal(vol[:, 0], 2 * torch.arange(float(b0), float(b1)))
    # device-time reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ret[rank] = bool(ok) and float(t) == world
    dist.destroy_process_group()


def test_row_block_partition_properties():
    sys.path.insert(0, ROOT)
    from panogrf_b200 import sharded
    for n in (1, 7, 8, 512, 513):
        for world in (1, 2, 3, 4, 8):
            blocks = [sharded.row_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_row_sharding_and_gather_world2():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = 29500 + os.getpid() % 2000
        procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
            assert p.exitcode == 0
        assert ret[0] and ret[1]


def test_reference_arm_prints_contract_line(capsys):
    """`bench.py --impl reference` (oracle port on the host cores) on a tiny sample."""
    sys.path.insert(0, ROOT)
    import bench
    v, t = bench.oracle_rays_per_s(64)
    assert v > 0 and t > 0
