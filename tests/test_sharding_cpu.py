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
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, W = bench.H, bench.W
    rows = (rank * H // world, (rank + 1) * H // world)
    que, ref = bench.make_inputs(torch, rows)
    coords = que["coords"][0]
    # every rank gets a contiguous block of rows, (x,y) pixel order, identical source maps
    assert coords.shape[0] == (rows[1] - rows[0]) * W
    assert float(coords[0, 1]) == rows[0] and float(coords[-1, 1]) == rows[1] - 1 and float(coords[-1, 0]) == W - 1
    ref_sum = torch.tensor([float(ref["imgs"].sum())])
    lst = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(lst, ref_sum)
    assert all(float(x) == float(lst[0]) for x in lst)
    # stand-in for the rendered tile: a function of the pixel id; gather reproduces the full image in row order
    tile = (coords[:, 1] * W + coords[:, 0]).contiguous()
    full = torch.empty(world * tile.numel())
    dist.all_gather_into_tensor(full, tile)
    ok = torch.equal(full.reshape(-1), torch.arange(H * W, dtype=torch.float32))
    # device-time reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ret[rank] = bool(ok) and float(t) == world
    dist.destroy_process_group()


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
