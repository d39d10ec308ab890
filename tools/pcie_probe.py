"""Host <-> device copy bandwidth per rank, alone and with every rank copying at once (pinned memory, one copy stream
per direction).  Explains the e2e leg of bench.py at N > 1: usage
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py"""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
MB = 256
h = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(MB << 20, dtype=torch.uint8, device='cuda')


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def bw(direction, both=False):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    sync()
    t0 = time.perf_counter()
    for _ in range(4):
        with torch.cuda.stream(s1):
            (d if direction == 'h2d' else h).copy_(h if direction == 'h2d' else d, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return 4 * MB / 1024 / dt * (2 if both else 1)


h2 = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
d2 = torch.empty(MB << 20, dtype=torch.uint8, device='cuda')
res = {}
for name, fn in (('h2d', lambda: bw('h2d')), ('d2h', lambda: bw('d2h')), ('h2d+d2h', lambda: bw('h2d', True))):
    fn()
    # all ranks at once
    v = torch.tensor([fn()], device='cuda', dtype=torch.float64)
    if world > 1:
        allv = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(allv, v)
        res[name + ' all ranks'] = [round(float(x), 1) for x in allv]
    else:
        res[name] = round(float(v), 1)
    # one rank at a time
    if world > 1:
        alone = []
        for r in range(world):
            sync()
            x = fn() if r == rank else 0.0
            t = torch.tensor([x], device='cuda', dtype=torch.float64)
            dist.all_reduce(t)
            alone.append(round(float(t), 1))
        res[name + ' one rank at a time'] = alone
if rank == 0:
    for k, v in res.items():
        print(f'{k:28s} GB/s per rank: {v}' + (f'  sum {sum(v):.1f}' if isinstance(v, list) else ''))
if world > 1:
    dist.destroy_process_group()
