"""GPU tests of the plan-level C entry (sdk_plan_*): the launch lists behind the C ABI must (1) replay to the same bits as the
Python host's own launches, (2) survive sdk_plan_save -> sdk_plan_load (every pointer relocated, TMA descriptors re-encoded), and
(3) run from a host that has neither Python nor PyTorch: tools/c_host/denoise.c, compiled here with gcc, runs UNet.forward and the
whole DDIM loop from an engine file and must reproduce the Python results bit for bit (same kernels, same tilings, same order)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import unet_oracle as UO
from stable_diffusion_pytorch_b200 import DDIMSampler, UNet, _lib
from stable_diffusion_pytorch_b200._lib import BF16_T, F32_T, ConvParams
from stable_diffusion_pytorch_b200.pipeline import DenoiseLoop
from stable_diffusion_pytorch_b200.unet import StepProgram

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def net(dev):
    sd = UO.make_state_dict(3, **UO.SD15)
    n = UNet(attention_head_dim=UO.SD15["attention_head_dim"], cross_attention_dim=UO.SD15["cross_attention_dim"])
    n.load_state_dict(sd, strict=True)
    return n.to(dev).eval().set_precision("bf16")


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def need_disk(path, gib):
    """Engine files of the full SD-1.5 architecture carry 1.8 GB of packed weights (the reference only works with its fixed
    block_out_channels, so there is no small model to export): skip rather than fail on a box without scratch space."""
    import shutil
    free = shutil.disk_usage(str(path)).free
    if free < gib * (1 << 30):
        pytest.skip(f"needs {gib} GiB of scratch space under {path}, {free / (1 << 30):.1f} GiB free")


def test_plan_replay_equals_python_launches(net, dev):
    """One sdk_plan_launch per program == the same launches issued one by one from Python."""
    pw = net._weights(dev)
    prog = StepProgram(net, pw, 2, 16, 16, 1, 2, 77)
    g = torch.Generator().manual_seed(5)
    prog.x_in.copy_(torch.randn(prog.x_in.shape, generator=g))
    prog.cond_in.copy_(torch.randn(prog.cond_in.shape, generator=g))
    prog.t_in.fill_(481)
    ctx_list, ops_list = list(prog.ctx_ops), list(prog.ops)         # ad-hoc copies: launched from Python, op by op
    prog.launch(ctx_list)
    prog.launch(ops_list)
    torch.cuda.synchronize()
    ref = prog.out.clone()
    prog.out.zero_()
    prog.launch(prog.ctx_ops)                                       # the plan's programs
    prog.launch(prog.ops)
    torch.cuda.synchronize()
    assert prog.plan is not None
    assert prog.lib.sdk_plan_num_launches(prog.plan, 0) == len(prog.ops) and prog.lib.sdk_plan_num_launches(prog.plan, 1) == len(prog.ctx_ops)
    assert torch.isfinite(ref).all() and torch.equal(prog.out, ref)


def test_engine_roundtrip_in_process(net, dev, tmp_path):
    """save -> load into a second plan with its own device memory -> same bits."""
    need_disk(tmp_path, 4)
    lib = _lib.lib()
    g = torch.Generator().manual_seed(6)
    x = torch.randn((2, 4, 16, 16), generator=g).to(dev)
    ctx = torch.randn((2, 77, 768), generator=g).to(dev)
    t = torch.tensor([301], device=dev)
    with torch.no_grad():
        ref = net(x, t, ctx)
    prog = next(p.prog for k, p in net._plans.items() if k[2:5] == (2, 16, 16))
    path = str(tmp_path / "unet.engine")
    net.export_engine(path, 2, 16, 16)                              # the plan forward() just used
    assert os.path.getsize(path) > 1_000_000_000                   # the packed bf16 weights travel with the launch lists
    h = C.c_void_p()
    _lib.check(lib.sdk_plan_load(path.encode(), C.byref(h)))
    os.remove(path)
    try:
        assert lib.sdk_plan_num_launches(h, 0) == len(prog.ops)
        s = stream()
        xs, cs, ts = x.cpu().contiguous(), ctx.cpu().contiguous(), torch.tensor([301], dtype=torch.int64)
        _lib.check(lib.sdk_plan_upload(h, b"x", xs.data_ptr(), xs.numel() * 4, s))
        _lib.check(lib.sdk_plan_upload(h, b"context", cs.data_ptr(), cs.numel() * 4, s))
        _lib.check(lib.sdk_plan_upload(h, b"timestep", ts.data_ptr(), 8, s))
        torch.cuda.synchronize()
        _lib.check(lib.sdk_plan_launch(h, 1, s))
        _lib.check(lib.sdk_plan_launch(h, 0, s))
        out = torch.empty((2, 4, 16, 16), dtype=torch.float32)
        _lib.check(lib.sdk_plan_download(h, b"out", out.data_ptr(), out.numel() * 4, s))
        torch.cuda.synchronize()
        assert torch.equal(out, ref.cpu())
        # the loaded plan owns different memory than the exporting one
        p0, p1, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(lib.sdk_plan_region(h, b"out", C.byref(p0), C.byref(n)))
        assert p0.value != prog.out.data_ptr() and n.value == out.numel() * 4
        # graph capture inside the plan: replay gives the same bits
        side = torch.cuda.Stream()
        sp = C.c_void_p(side.cuda_stream)
        _lib.check(lib.sdk_plan_capture(h, 0, sp))
        out2 = torch.zeros_like(out)
        _lib.check(lib.sdk_plan_launch(h, 0, sp))
        _lib.check(lib.sdk_plan_download(h, b"out", out2.data_ptr(), out2.numel() * 4, sp))
        side.synchronize()
        assert torch.equal(out2, ref.cpu())
    finally:
        _lib.check(lib.sdk_plan_destroy(h))


def test_c_host_runs_forward_and_loop(net, dev, tmp_path):
    """tools/c_host/denoise.c (plain C, sdb200.h only) == UNet.forward and DenoiseLoop.run of the Python host, bit for bit."""
    need_disk(tmp_path, 4)
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler on this box (the in-process reload test covers the engine format)")
    exe = str(tmp_path / "denoise")
    libdir = os.path.join(ROOT, "stable-diffusion-pytorch_b200")
    cc = subprocess.run(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "c_host", "denoise.c"), "-o", exe,
                         "-L", libdir, "-lsdb200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    g = torch.Generator().manual_seed(7)
    lat = torch.randn((1, 4, 16, 16), generator=g)
    ctx = torch.randn((2, 77, 768), generator=g)
    smp = DDIMSampler()
    smp._set_inference_steps(10)
    steps = 6
    with torch.no_grad():
        loop = DenoiseLoop(net, smp, 1, 16, 16, do_cfg=True, cfg_scale=7.5)
        ref = loop.run(lat.to(dev), ctx.to(dev), steps=steps).cpu()
        loop.reset(lat.to(dev), ctx.to(dev))
        eng = str(tmp_path / "loop.engine")
        loop.export_engine(eng)
        # one forward of the same program (CFG batch 2 from 1 stored latent: latent.repeat(2,...) is folded into the program)
        loop.prog.t_in.fill_(781)
        loop.prog.launch(loop.prog.ops)
        torch.cuda.synchronize()
        fwd_ref = loop.prog.out.cpu().clone()
    lat.numpy().tofile(str(tmp_path / "lat.f32"))
    ctx.numpy().tofile(str(tmp_path / "ctx.f32"))
    env = dict(os.environ)
    r = subprocess.run([exe, eng, "loop", str(tmp_path / "lat.f32"), str(tmp_path / "ctx.f32"), str(steps), str(tmp_path / "out_loop.f32")],
                       capture_output=True, text=True, env=env, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stderr
    got = torch.from_numpy(np.fromfile(str(tmp_path / "out_loop.f32"), dtype=np.float32).reshape(1, 4, 16, 16))
    assert torch.isfinite(got).all() and torch.equal(got, ref)
    r = subprocess.run([exe, eng, "forward", str(tmp_path / "lat.f32"), str(tmp_path / "ctx.f32"), "781", str(tmp_path / "out_fwd.f32")],
                       capture_output=True, text=True, env=env, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stderr
    os.remove(eng)
    got = torch.from_numpy(np.fromfile(str(tmp_path / "out_fwd.f32"), dtype=np.float32).reshape(2, 4, 16, 16))
    assert torch.equal(got, fwd_ref)


def test_vae_and_text_encoder_engines(dev, tmp_path):
    """The stages either side of the loop export the same way: VAE.decode ("z" -> "out") and the text encoder ("ids" -> "out")
    reloaded from their engine files reproduce the Python host's results bit for bit."""
    from oracle import vae_oracle as VO
    from stable_diffusion_pytorch_b200 import VAE, TextEncoder
    lib = _lib.lib()
    vae = VAE()
    vae.load_state_dict(VO.make_state_dict(3), strict=True)
    vae = vae.to(dev).eval()
    g = torch.Generator().manual_seed(9)
    z = torch.randn((1, 4, 8, 8), generator=g)
    enc = TextEncoder(n_vocab=1000, embed_dim=768, max_len=77, num_layers=2).to(dev).eval()
    ids = torch.randint(0, 1000, (2, 77), generator=g)
    with torch.no_grad():
        img = vae.decode(z.to(dev)).cpu()
        emb = enc(ids.to(dev)).cpu()
    cases = [(next(iter(vae._plans.values())).prog, "z", z, img), (next(iter(enc._plans.values())).prog, "ids", ids, emb)]
    for prog, name, inp, ref in cases:
        path = str(tmp_path / f"{name}.engine")
        prog.export_engine(path)
        h = C.c_void_p()
        _lib.check(lib.sdk_plan_load(path.encode(), C.byref(h)))
        os.remove(path)
        s = stream()
        src = inp.contiguous()
        _lib.check(lib.sdk_plan_upload(h, name.encode(), src.data_ptr(), src.numel() * src.element_size(), s))
        _lib.check(lib.sdk_plan_launch(h, 0, s))
        out = torch.empty_like(ref)
        _lib.check(lib.sdk_plan_download(h, b"out", out.data_ptr(), out.numel() * 4, s))
        torch.cuda.synchronize()
        _lib.check(lib.sdk_plan_destroy(h))
        assert torch.isfinite(out).all() and torch.equal(out, ref), name


def test_plan_with_parameter_struct_launch(dev, tmp_path):
    """The exact-fp32 GEMM takes a host parameter struct: the plan copies it, relocates its pointers and replays it after a reload."""
    lib = _lib.lib()
    g = torch.Generator().manual_seed(8)
    x = torch.randn((2 * 8 * 8, 64), generator=g).to(dev)
    w = (torch.randn((32, 9 * 64), generator=g) / 24).to(dev)
    b = torch.randn((32,), generator=g).to(dev)
    out = torch.zeros((2 * 8 * 8, 32), device=dev)
    p = ConvParams()
    p.src0, p.C0, p.weight, p.bias, p.out = x.data_ptr(), 64, w.data_ptr(), b.data_ptr(), out.data_ptr()
    p.B, p.Hin, p.Win, p.Hout, p.Wout, p.ksize, p.stride, p.N = 2, 8, 8, 8, 8, 3, 1, 32
    p.in_dtype, p.out_dtype = F32_T, F32_T
    _lib.check(lib.sdk_conv_gemm_f32(C.byref(p), stream()))
    torch.cuda.synchronize()
    ref = out.clone()
    h = C.c_void_p()
    _lib.check(lib.sdk_plan_create(C.byref(h)))
    _lib.check(lib.sdk_plan_add_launch(h, 0, b"sdk_conv_gemm_f32", (C.c_uint64 * 1)(C.addressof(p)), 1))
    p.N = 0                                                           # the plan holds its own copy of the struct
    for t_, kind, name in ((x, 0, b""), (w, 0, b""), (b, 0, b""), (out, 2, b"out")):
        _lib.check(lib.sdk_plan_add_region(h, t_.data_ptr(), t_.numel() * 4, kind, name))
    path = str(tmp_path / "conv.engine")
    _lib.check(lib.sdk_plan_save(h, path.encode()))
    # a launch whose pointer is in no region cannot be saved
    stray = torch.zeros(16, device=dev)
    _lib.check(lib.sdk_plan_add_launch(h, 1, b"sdk_zero", (C.c_uint64 * 2)(stray.data_ptr(), 64), 2))
    assert lib.sdk_plan_save(h, (path + ".bad").encode()) == -1 and b"outside every registered region" in lib.sdk_last_error()
    _lib.check(lib.sdk_plan_destroy(h))
    h2 = C.c_void_p()
    _lib.check(lib.sdk_plan_load(path.encode(), C.byref(h2)))
    _lib.check(lib.sdk_plan_launch(h2, 0, stream()))
    got = torch.empty((2 * 8 * 8, 32), dtype=torch.float32)
    _lib.check(lib.sdk_plan_download(h2, b"out", got.data_ptr(), got.numel() * 4, stream()))
    torch.cuda.synchronize()
    _lib.check(lib.sdk_plan_destroy(h2))
    assert torch.equal(got, ref.cpu())
