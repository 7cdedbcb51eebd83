"""Single-layer A/B of the weight-stationary persistent walk: times sdk_tc_gemm for a few short-K shapes with
weight_stationary = 1 (never) / 2 (whenever it fits) and several N tiles.  python tools/ws_probe.py"""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stable_diffusion_pytorch_b200 import _lib
from stable_diffusion_pytorch_b200._lib import BF16_T, F32_T, TcGemmDesc

dev = torch.device("cuda:0")
lib = _lib.lib()
stream = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def kmajor(w):
    n, k = w.shape
    return w.view(n, k // 64, 64).permute(1, 0, 2).contiguous()


def run(name, M, K, N, geglu, odt, res):
    a = (torch.randn((M, K), device=dev)).bfloat16()
    w = (torch.randn((N, K), device=dev) / math.sqrt(K)).bfloat16()
    wk = kmajor(w)
    bias = torch.randn((N,), device=dev) * 0.1
    Nout = N // 2 if geglu else N
    resid = torch.randn((M, Nout), device=dev) if res else None
    out = torch.empty((M, Nout), device=dev, dtype=torch.float32 if odt == F32_T else torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for bn in (64, 128, 160, 256):
        if N % bn:
            continue
        for ws in (1, 2):
            d = TcGemmDesc()
            d.w_kmajor, d.w_const = 1, 1
            d.a[0], d.w[0], d.C[0], d.ksize[0], d.nseg = a.data_ptr(), wk.data_ptr(), K, 1, 1
            d.B, d.H, d.W, d.N = 1, 1, M, N
            d.bias, d.out, d.out_dtype, d.geglu = bias.data_ptr(), out.data_ptr(), odt, int(geglu)
            d.residual = resid.data_ptr() if res else 0
            d.block_n, d.splits, d.weight_stationary = bn, 1, ws
            h = C.c_void_p()
            if lib.sdk_tc_gemm_create(C.byref(d), C.byref(h)) != 0:
                continue
            info = (C.c_int * 11)()
            lib.sdk_tc_gemm_info(h, info, 11)
            wsb = torch.zeros(max(int(lib.sdk_tc_gemm_workspace_bytes(h)), 256), dtype=torch.uint8, device=dev)
            lib.sdk_tc_gemm_set_workspace(h, wsb.data_ptr())
            for _ in range(3):
                lib.sdk_tc_gemm_launch(h, stream())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                lib.sdk_tc_gemm_launch(h, stream())
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 50
            gf = 2.0 * M * N * K / 1e9
            print(f"{name:14s} M={M} K={K} N={N} bn={bn:3d} ws={ws} kernel={('plain','persistent','persistent-WS')[info[10]]:14s} grid={info[2]:4d} "
                  f"{us:7.2f} us  {gf / us * 1e-3:7.1f} TFLOP/s")
            lib.sdk_tc_gemm_destroy(h)


run("ff0_L0", 8192, 320, 2560, True, BF16_T, False)
run("qkv_L0", 8192, 320, 960, False, BF16_T, False)
run("ff0_L1", 2048, 640, 5120, True, BF16_T, False)
run("o1_L0_res", 8192, 320, 320, False, F32_T, True)
run("ff0_L0_b16", 65536, 320, 2560, True, BF16_T, False)
run("qkv_L0_b16", 65536, 320, 960, False, BF16_T, False)
