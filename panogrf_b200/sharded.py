"""Multi-GPU sharding of the hot path: one process per GPU, rays / batch items split across ranks.

* Renderer: the rows of the target ERP view are split into `world` contiguous blocks; every rank holds the source maps and
  the weights and renders its block; the ONLY collective is one `all_gather_into_tensor` of the per-rank output tile
  `(rays, 4) = (r, g, b, depth)` — 16 B per ray over NCCL / NVLink.
* Cost volume: voxels are independent; batch items are dealt to the ranks (configs[2]: one item per GPU), no exchange.

The functions only touch `torch.distributed` when a process group is initialised and `world > 1`; the partition arithmetic is
plain Python so that the CPU (gloo) tests exercise exactly what the GPU path runs.
"""
import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def row_block(height, rank, world):
    """Contiguous block of ERP rows of `rank`: [r0, r1).  Rows are spread as evenly as possible (block sizes differ by <= 1)."""
    base, rem = divmod(int(height), int(world))
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def item_block(n_items, rank, world):
    """Batch items of `rank` for the cost volume: same dealing rule as `row_block`."""
    return row_block(n_items, rank, world)


def shard_coords(height, width, rank, world, device=None):
    """(1, rows*W, 2) pixel coordinates (x, y) of the rank's row block, row-major — the layout of que_imgs_info['coords']."""
    r0, r1 = row_block(height, rank, world)
    ys, xs = torch.meshgrid(torch.arange(r0, r1), torch.arange(width), indexing="ij")
    c = torch.stack([xs, ys], -1).reshape(1, -1, 2).float()
    return c.to(device) if device is not None else c


def pack_tile(rgb, depth):
    """(1,rn,3), (1,rn) -> (rn,4) contiguous tile (r, g, b, depth)."""
    return torch.cat([rgb.reshape(-1, 3), depth.reshape(-1, 1)], 1).contiguous()


def gather_tiles(tile, height, width, group=None, out=None):
    """All-gather the per-rank `(rows*W, 4)` tiles into the full `(H*W, 4)` image (every rank gets it).  Equal blocks use one
    `all_gather_into_tensor`; ragged blocks (H % world != 0) are padded to the largest block for the collective."""
    rank, world = world_info(group)
    if world == 1:
        return tile
    sizes = [(row_block(height, r, world)[1] - row_block(height, r, world)[0]) * width for r in range(world)]
    big = max(sizes)
    send = tile
    if tile.shape[0] != big:
        send = tile.new_zeros(big, tile.shape[1])
        send[:tile.shape[0]] = tile
    if out is None or out.shape[0] != world * big:
        out = tile.new_empty(world * big, tile.shape[1])
    dist.all_gather_into_tensor(out, send, group=group)
    if all(s == big for s in sizes):
        return out
    return torch.cat([out[r * big:r * big + sizes[r]] for r in range(world)], 0)


def render_view_sharded(net, que_imgs_info, ref_imgs_info, height, width, group=None, out=None, suffix="_fine"):
    """Render this rank's row block of the (height, width) query view with `net.render` and all-gather (rgb, depth).

    `que_imgs_info` carries the pose (`c2w`, `depth_range`); its `coords` (if any) are replaced by the rank's block.
    Returns the full image as a `(H*W, 4)` tensor on every rank."""
    rank, world = world_info(group)
    q = dict(que_imgs_info)
    q["coords"] = shard_coords(height, width, rank, world, ref_imgs_info["imgs"].device)
    o = net.render(q, ref_imgs_info, False)
    key_d = "render_depth" + suffix
    if key_d not in o:
        raise KeyError(f"{key_d}: the sharded view gathers (rgb, depth) tiles; set cfg['render_depth'] = True")
    return gather_tiles(pack_tile(o["pixel_colors_nr" + suffix], o[key_d]), height, width, group, out)


def cost_volume_sharded(fn, images, trans, rots, group=None, **kw):
    """Run `fn(images[b0:b1], trans[b0:b1], rots[b0:b1], **kw)` (e.g. a partial of `calculate_cost_volume_erp`) on this
    rank's batch items; returns (volume of the local items, (b0, b1)).  No collective: the regulariser consumes the volume per
    item."""
    rank, world = world_info(group)
    b0, b1 = item_block(images.shape[0], rank, world)
    if b1 == b0:
        return None, (b0, b1)
    return fn(images[b0:b1], trans[b0:b1], rots[b0:b1], **kw), (b0, b1)
