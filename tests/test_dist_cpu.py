"""CPU suite, part 4: the N>1 path — sharding rules and the final-latent all-gather over gloo, world_size 2."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stable_diffusion_pytorch_b200.dist import gather_latents, shard_inputs, shard_range


def test_shard_range_covers_batch():
    for total in (1, 2, 7, 64, 256):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_shard_inputs_keeps_cfg_pair_together():
    B = 6
    lat = torch.arange(B, dtype=torch.float32).view(B, 1, 1, 1).expand(B, 4, 2, 2).contiguous()
    ctx = torch.arange(2 * B, dtype=torch.float32).view(2 * B, 1, 1).expand(2 * B, 77, 8).contiguous()   # rows [uncond ; cond]
    l1, c1 = shard_inputs(lat, ctx, 1, 4)
    lo, hi = shard_range(B, 1, 4)
    assert l1[:, 0, 0, 0].tolist() == list(range(lo, hi))
    n = hi - lo
    assert c1[:n, 0, 0].tolist() == list(range(lo, hi))                 # uncond rows of these images
    assert c1[n:, 0, 0].tolist() == [B + i for i in range(lo, hi)]      # their cond rows
    with pytest.raises(ValueError):
        shard_inputs(lat, ctx[:B], 0, 2)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    full = torch.randn((total, 4, 8, 8), generator=g)              # full batch from ONE generator, then sliced
    ctx = torch.randn((2 * total, 77, 16), generator=g)
    local, _ = shard_inputs(full, ctx, rank, world)
    out = gather_latents(local * 2.0, total)                       # stand-in for the per-rank denoising result
    ok = torch.equal(out, full * 2.0)
    q.put((rank, bool(ok), tuple(out.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [4, 5])
def test_gather_latents_gloo_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and all(shape == (total, 4, 8, 8) for _, _, shape in res)
