"""GPU parity at the shapes the multi-GPU BASELINE configurations put on ONE GPU (where the GEMM tilings change: multi-wave
persistent grids, co-resident CTAs, other tune-cache entries), checked against the CPU oracle on the same seeded inputs:

  config 3  SD1.5-arch 64x64, batch 64 over 8 GPUs  -> UNet batch 16 per GPU (8 images x CFG pair)
  config 4  SD2.1-arch 96x96, batch 16 over 8 GPUs  -> UNet batch 4 per GPU
  config 5  SD2.1-arch 64x64 one-step, batch 256 over 8 GPUs -> UNet batch 32 per GPU, no CFG, then x0 = (x - sigma eps)/alpha

and "a sharded run equals the full-batch run" (SURVEY 8(e)): shard_inputs -> DenoiseLoop per shard -> concat == the loop on
the whole batch.  Gates (north_star): rel-L2 <= 1e-4 in fp32 mode, <= 1e-2 in bf16, bookkeeping bit-exact.
The oracle is evaluated sample by sample (every sample is independent through the UNet), which bounds its memory."""
import numpy as np
import pytest
import torch

from oracle import sampler_oracle as SO
from oracle import unet_oracle as UO
from stable_diffusion_pytorch_b200 import DDIMSampler, UNet
from stable_diffusion_pytorch_b200.dist import shard_inputs
from stable_diffusion_pytorch_b200.pipeline import DenoiseLoop, one_step

pytestmark = pytest.mark.gpu
FP32_TOL, BF16_TOL = 1e-4, 1e-2


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _net(cfg, seed, dev):
    sd = UO.make_state_dict(seed, **cfg)
    net = UNet(attention_head_dim=cfg["attention_head_dim"], cross_attention_dim=cfg["cross_attention_dim"])
    net.load_state_dict(sd, strict=True)
    return net.to(dev).eval(), sd


@pytest.fixture(scope="module")
def sd15(dev):
    return _net(UO.SD15, 0, dev)


@pytest.fixture(scope="module")
def sd21(dev):
    return _net(UO.SD21, 1, dev)


_ORACLE_CACHE = {}


def oracle_forward_per_sample(sd, arch, x, t, ctx, ctx_bcast=False):
    """Oracle UNet on a batch, one sample at a time (SDPA attention as the reference, attention.py:37-43); memoised on the
    inputs so the fp32 and bf16 variants of a test share one CPU evaluation."""
    key = (id(sd), tuple(x.shape), float(x.double().sum()), int(t[0]), float(ctx.double().sum()), ctx_bcast)
    if key in _ORACLE_CACHE:
        return _ORACLE_CACHE[key]
    out = _oracle_forward_per_sample(sd, arch, x, t, ctx, ctx_bcast)
    _ORACLE_CACHE[key] = out
    return out


def _oracle_forward_per_sample(sd, arch, x, t, ctx, ctx_bcast):
    old = UO.USE_SDPA
    UO.USE_SDPA = True
    torch.set_num_threads(max(1, torch.get_num_threads()))
    outs = []
    try:
        with torch.no_grad():
            for i in range(x.shape[0]):
                c = ctx[:1] if ctx_bcast else ctx[i:i + 1]
                outs.append(UO.unet_forward(sd, x[i:i + 1], t, c, **arch))
    finally:
        UO.USE_SDPA = old
    return torch.cat(outs, 0)


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_config3_unet_batch16_forward(sd15, dev, precision, tol):
    """Config 3 per-GPU shape at 8 GPUs: 8 images -> UNet batch 16 at 64x64 (multi-wave grids)."""
    net, sd = sd15
    net.set_precision(precision)
    g = torch.Generator().manual_seed(303)
    lat = torch.randn((8, 4, 64, 64), generator=g)
    ctx = torch.randn((16, 77, 768), generator=g)
    t = torch.tensor([981])
    x = lat.repeat(2, 1, 1, 1)
    with torch.no_grad():
        got = net(x.to(dev), t.to(dev), ctx.to(dev)).cpu()
    # the oracle on a subset of samples is enough to pin every tile class (first / middle / last samples, both CFG halves)
    idx = [0, 5, 7, 8, 13, 15]
    want = oracle_forward_per_sample(sd, UO.SD15, x[idx], t, ctx[idx])
    e = rel_l2(got[idx].numpy(), want.numpy())
    print(f"config 3 shape (UNet batch 16, 64x64) {precision}: rel-L2 {e:.3e} on samples {idx}")
    assert e < tol
    # remaining samples: batch independence inside OUR program -- the same sample at another batch position gives the same row
    with torch.no_grad():
        x2 = x.clone()
        x2[3] = x[0]
        c2 = ctx.clone()
        c2[3] = ctx[0]
        got2 = net(x2.to(dev), t.to(dev), c2.to(dev)).cpu()
    assert rel_l2(got2[3].numpy(), got[0].numpy()) < (1e-5 if precision == "fp32" else 2e-3)
    net.set_precision("bf16")


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_config4_unet_batch4_forward(sd21, dev, precision, tol):
    """Config 4 per-GPU shape at 8 GPUs: SD2.1-arch 96x96, 2 images -> UNet batch 4, v-prediction step afterwards."""
    net, sd = sd21
    net.set_precision(precision)
    g = torch.Generator().manual_seed(404)
    lat = torch.randn((2, 4, 96, 96), generator=g)
    ctx = torch.randn((4, 77, 1024), generator=g)
    t = torch.tensor([961])
    x = lat.repeat(2, 1, 1, 1)
    with torch.no_grad():
        got = net(x.to(dev), t.to(dev), ctx.to(dev))
    want = oracle_forward_per_sample(sd, UO.SD21, x, t, ctx)
    e = rel_l2(got.cpu().numpy(), want.numpy())
    print(f"config 4 shape (UNet batch 4, 96x96) {precision}: UNet output rel-L2 {e:.3e}")
    assert e < tol
    # CFG + v-prediction DDIM update on top: kernel (ours) vs oracle statements on the ORACLE's model output -> the latent gate
    smp = DDIMSampler(prediction_type="v_prediction")
    smp._set_inference_steps(50)
    y = smp.reverse_process(lat.to(dev), t.to(dev), got, cfg_scale=7.5).cpu().numpy()
    _, alphas, a_hat = SO.schedule_fp32()
    u, c = SO.cfg_blend(want.numpy())
    ref = SO.ddim_reverse(lat.numpy(), 961, SO.cfg_combine(u, c, 7.5), alphas, a_hat, 1000, 50, "v_prediction")
    e2 = rel_l2(y, ref)
    print(f"config 4 shape {precision}: latent after one CFG + v-pred DDIM step rel-L2 {e2:.3e}")
    assert e2 < tol
    net.set_precision("bf16")


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_config5_one_step_batch32(sd21, dev, precision, tol):
    """Config 5 per-GPU shape: SwiftBrush one-step, SD2.1-arch 64x64, 32 images, no CFG, ONE broadcast prompt, t = 999,
    x0 = (x - sigma_T eps)/alpha_T (models/diffusion.py:106-113)."""
    net, sd = sd21
    net.set_precision(precision)
    g = torch.Generator().manual_seed(505)
    lat = torch.randn((32, 4, 64, 64), generator=g)
    ctx = torch.randn((1, 77, 1024), generator=g)
    smp = DDIMSampler()
    with torch.no_grad():
        x0 = one_step(net, smp, lat.to(dev), ctx.to(dev)).cpu().numpy()
    idx = [0, 11, 31]
    t = torch.tensor([int(smp.timesteps[0])])
    assert int(t) == 999
    eps = oracle_forward_per_sample(sd, UO.SD21, lat[idx], t, ctx, ctx_bcast=True).numpy()
    want = SO.x0_from_eps(lat[idx].numpy(), eps)
    e = rel_l2(x0[idx], want)
    print(f"config 5 shape (one-step, batch 32, 64x64) {precision}: x0 rel-L2 {e:.3e} on samples {idx}")
    assert e < tol
    net.set_precision("bf16")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sharded_run_equals_full_batch(sd15, dev, precision):
    """full-batch tensors from ONE generator, sliced per rank with the CFG pair kept together; every shard's
    DenoiseLoop (what each rank of bench.py runs) must reproduce the rows of the full-batch loop.  The shards use other
    GEMM tilings than the full batch, so fp32 agrees to summation order and bf16 to the bf16 gate; the timestep walk is
    bit-exact."""
    net, _ = sd15
    net.set_precision(precision)
    B, hw, steps = 4, 16, 4
    g = torch.Generator().manual_seed(1234)
    lat = torch.randn((B, 4, hw, hw), generator=g)
    ctx = torch.randn((2 * B, 77, 768), generator=g)
    smp = DDIMSampler()
    smp._set_inference_steps(10)
    with torch.no_grad():
        full_loop = DenoiseLoop(net, smp, B, hw, hw)
        full = full_loop.run(lat.to(dev), ctx.to(dev), steps=steps).cpu()
        t_full = int(full_loop.prog.t_in.item())
        for world in (2, 4):
            parts = []
            for rank in range(world):
                l, c = shard_inputs(lat, ctx, rank, world, do_cfg=True)
                loop = DenoiseLoop(net, smp, l.shape[0], hw, hw)
                parts.append(loop.run(l.to(dev), c.to(dev), steps=steps).cpu())
                assert int(loop.prog.t_in.item()) == t_full                  # same last timestep walked: bit-exact bookkeeping
            got = torch.cat(parts, 0)
            e = rel_l2(got.numpy(), full.numpy())
            print(f"sharded x{world} vs full batch ({precision}): rel-L2 {e:.3e}")
            assert e < (1e-5 if precision == "fp32" else BF16_TOL)
    net.set_precision("bf16")


def test_inference_mode_callers(sd15, dev):
    """The reference's inpaint path runs under torch.inference_mode() (models/diffusion.py:328): a context produced there has no
    version counter; plans built there must stay usable outside it."""
    net, _ = sd15
    net.set_precision("bf16")
    net.invalidate()
    g = torch.Generator().manual_seed(7)
    lat = torch.randn((1, 4, 8, 8), generator=g)
    with torch.inference_mode():
        ctx = torch.randn((2, 77, 768), generator=g).to(dev)          # inference tensor
        x = lat.repeat(2, 1, 1, 1).to(dev)
        t = torch.tensor([981], device=dev)
        a = net(x, t, ctx)                                              # builds packed weights + plan inside inference mode
        b = net(x, t, ctx)
        smp = DDIMSampler()
        smp._set_inference_steps(10)
        loop = DenoiseLoop(net, smp, 1, 8, 8)
        y1 = loop.run(lat.to(dev), ctx, steps=2)
    with torch.no_grad():                                               # same plan, outside inference mode
        c = net(x.clone(), t.clone(), ctx.clone())
        y2 = loop.run(lat.to(dev), ctx.clone(), steps=2)
    assert torch.equal(a, b) and torch.equal(a, c) and torch.equal(y1, y2)
    net.invalidate()


def test_context_batch_must_match_or_broadcast(sd15, dev):
    """A context batch outside {1, UNet batch} used to broadcast the first 77 rows to every sample silently."""
    net, _ = sd15
    smp = DDIMSampler()
    smp._set_inference_steps(10)
    with pytest.raises(RuntimeError, match="context batch"):
        DenoiseLoop(net, smp, 2, 8, 8, do_cfg=True, context_batch=2)     # UNet batch 4, context (2,77,D)
    with pytest.raises(RuntimeError):
        net(torch.zeros((4, 4, 8, 8), device=dev), torch.tensor([1], device=dev), torch.zeros((2, 77, 768), device=dev))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_non_current_device(dev):
    """UNet + sampler on cuda:1 while cuda:0 is the current device (per-device smem opt-in, launches on the tensor's stream)."""
    d1 = torch.device("cuda:1")
    torch.cuda.set_device(0)
    net, sd = _net(UO.SD15, 0, d1)
    lat, ctx = UO.synthetic_inputs(1, 8, 8, 768, seed=3)
    t = torch.tensor([981])
    with torch.no_grad():
        want = UO.unet_forward(sd, lat.repeat(2, 1, 1, 1), t, ctx, **UO.SD15)
        got = net(lat.repeat(2, 1, 1, 1).to(d1), t.to(d1), ctx.to(d1))
        smp = DDIMSampler()
        smp._set_inference_steps(10)
        y = smp.reverse_process(lat.to(d1), t.to(d1), got, cfg_scale=7.5)
    assert torch.cuda.current_device() == 0
    assert rel_l2(got.cpu().numpy(), want.numpy()) < BF16_TOL and torch.isfinite(y).all()
