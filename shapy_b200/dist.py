"""Batch sharding across the GPUs of one NVSwitch node.

The path is embarrassingly parallel over the batch (eval-mode BatchNorm, no cross-sample term), so the only
collectives are the ones the north star names: a scatter of the image batch from rank 0 and a gather of
the per-body results back to rank 0, both over NCCL / NVLink (`torch.distributed`, one process per GPU).
The reference has nothing to mirror here: its inference is single-GPU (evaluation.py:641-642).
"""
import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous shards in global order: rank r owns [r * total / world, (r + 1) * total / world)."""
    per = (total + world - 1) // world
    lo = min(rank * per, total)
    return lo, min(lo + per, total)


def scatter_images(images, per_rank: int, shape, device, src: int = 0):
    """rank `src` passes the full (world * per_rank, 3, H, W) batch, the others None; returns the local shard."""
    world, rank = dist.get_world_size(), dist.get_rank()
    local = torch.empty((per_rank,) + tuple(shape), dtype=torch.float32, device=device)
    if world == 1:
        local.copy_(images)
        return local
    chunks = list(images.chunk(world, dim=0)) if rank == src else None
    dist.scatter(local, chunks, src=src)
    return local


def gather_results(tensors: dict, dst: int = 0):
    """Gathers every (B_local, ...) tensor of `tensors` to rank `dst`, keeping the global batch order."""
    world, rank = dist.get_world_size(), dist.get_rank()
    out = {}
    for k, t in tensors.items():
        t = t.contiguous()
        if world == 1:
            out[k] = t
            continue
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, bufs, dst=dst)
        out[k] = torch.cat(bufs, dim=0) if rank == dst else None
    return out


def sharded_forward(model, images, per_rank: int, shape=(3, 224, 224), device=None, src: int = 0):
    """images: full batch on rank `src` (device tensor), None elsewhere.  Returns the gathered result dict on
    rank `src` (None on the others): vertices, v_shaped, betas, joints, measurements (B, 5)."""
    device = device or torch.device('cuda', torch.cuda.current_device())
    local = scatter_images(images, per_rank, shape, device, src)
    with torch.no_grad():
        out = model(local)
    st = out['stage_02']
    meas = torch.stack([out['measurements'][k] for k in ('mass', 'height', 'chest', 'waist', 'hips')], dim=1)
    res = dict(vertices=st['vertices'], v_shaped=st['v_shaped'], betas=st['betas'], joints=st['joints']._t,
               measurements=meas)
    return gather_results(res, dst=src)


def scatter_bytes(data_u8, nbytes_per_rank: int, device, src: int = 0):
    """rank `src` passes a uint8 tensor of world * nbytes_per_rank bytes (rank r's shard at [r * n, (r + 1) * n)), the
    others None; returns the local shard as a flat uint8 tensor on `device`."""
    world, rank = dist.get_world_size(), dist.get_rank()
    local = torch.empty(nbytes_per_rank, dtype=torch.uint8, device=device)
    if world == 1:
        local.copy_(data_u8.reshape(-1))
        return local
    chunks = list(data_u8.reshape(world, nbytes_per_rank).unbind(0)) if rank == src else None
    dist.scatter(local, chunks, src=src)
    return local


def sharded_forward_u8(model, input_stage, images_u8, per_rank: int, height: int, width: int, device=None, src: int = 0,
                       keys=('vertices', 'betas', 'measurements')):
    """BASELINE configs[4] as a call: rank `src` holds the whole batch as uint8 images (world * per_rank, H, W, 3)
    on its device (None elsewhere).  Scatter ships ONE byte per channel over NVLink (4x less rank-0 egress than fp32
    crops: 512 images = 77 MB instead of 308 MB), every rank crops / resizes / normalises its shard on the device
    (`shapy_preprocess_forward`), runs the regressor, and the per-body results named by `keys` are gathered back to
    `src` in global batch order.  Returns the gathered dict on `src`, None elsewhere."""
    device = device or torch.device('cuda', torch.cuda.current_device())
    local = scatter_bytes(images_u8, per_rank * height * width * 3, device, src)
    key = (per_rank, height, width, str(device))
    cache = getattr(input_stage, '_dist_tables', None)
    if cache is None:
        cache = input_stage._dist_tables = {}
    if key not in cache:
        cache[key] = input_stage.uniform_table(per_rank, height, width).to(device)
    with torch.no_grad():
        x = input_stage.run_device(local, cache[key], per_rank)
        out = model(x)
    st = out[out['stage_keys'][-1]]
    res = dict(vertices=st['vertices'], v_shaped=st['v_shaped'], betas=st['betas'], joints=st['joints']._t,
               measurements=torch.stack([out['measurements'][k] for k in ('mass', 'height', 'chest', 'waist', 'hips')], dim=1))
    return gather_results({k: res[k] for k in keys}, dst=src)
