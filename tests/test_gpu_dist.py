"""2-GPU NCCL test of the batch-sharded forward (skipped on a single-GPU box): gather(outputs of the shards)
must equal the single-GPU outputs of the same images bit for bit (same per-rank batch size => same launch
configuration; the path has no cross-sample term)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    from shapy_b200 import dist as sdist, synth
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        model = synth.build_synthetic_regressor().to(dev).eval()
        per = 2
        full = torch.randn(world * per, 3, 224, 224, generator=torch.Generator().manual_seed(0))
        res = sdist.sharded_forward(model, full.to(dev) if rank == 0 else None, per, device=dev)
        if rank == 0:
            ok = True
            for r in range(world):
                with torch.no_grad():
                    o = model(full[r * per:(r + 1) * per].to(dev))
                ok = ok and torch.equal(res['vertices'][r * per:(r + 1) * per], o['stage_02']['vertices'])
                ok = ok and torch.equal(res['betas'][r * per:(r + 1) * per], o['stage_02']['betas'])
                ok = ok and torch.equal(res['measurements'][r * per:(r + 1) * per, 2], o['measurements']['chest'])
            ret[0] = bool(ok)
        # configs[4] path: rank 0 holds uint8 images, scatter ships bytes, every rank preprocesses on the device
        from shapy_b200.preprocess import InputStage
        stage = InputStage(dev)
        u8 = torch.randint(0, 256, (world * per, 224, 224, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
        res8 = sdist.sharded_forward_u8(model, stage, u8.to(dev) if rank == 0 else None, per, 224, 224, device=dev)
        if rank == 0:
            ok8 = True
            for r in range(world):
                shard = u8[r * per:(r + 1) * per].to(dev)
                x = stage.run_device(shard.reshape(-1), stage.uniform_table(per, 224, 224).to(dev), per)
                with torch.no_grad():
                    o = model(x)
                ok8 = ok8 and torch.equal(res8['vertices'][r * per:(r + 1) * per], o['stage_02']['vertices'])
                ok8 = ok8 and torch.equal(res8['betas'][r * per:(r + 1) * per], o['stage_02']['betas'])
            ret[1] = bool(ok8)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_sharded_forward_matches_single_gpu():
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, 29500 + (os.getpid() % 2000), ret), nprocs=2, join=True)
    assert ret.get(0) is True
    assert ret.get(1) is True
