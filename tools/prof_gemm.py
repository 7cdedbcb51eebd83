"""Launch representative tcgen05 implicit-GEMM shapes of the SD1.5 step (UNet batch 2, 64x64 latent)
for ncu / event timing.   python tools/prof_gemm.py [--iters N] [--block-n BN] [--splits S] [--only name]"""
import argparse
import ctypes as C
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stable_diffusion_pytorch_b200 import _lib  # noqa: E402
from stable_diffusion_pytorch_b200._lib import BF16_T, F32_T, TcGemmDesc  # noqa: E402

SHAPES = [
    # name, B, H, W, C, N, k
    ("conv3_L0_320", 2, 64, 64, 320, 320, 3),
    ("conv3_L0_960", 2, 64, 64, 960, 320, 3),
    ("conv3_L1_640", 2, 32, 32, 640, 640, 3),
    ("conv3_L2_1280", 2, 16, 16, 1280, 1280, 3),
    ("conv3_L3_1280", 2, 8, 8, 1280, 1280, 3),
    ("conv3_L3_2560", 2, 8, 8, 2560, 1280, 3),
    ("lin_L0_qkv", 1, 1, 8192, 320, 960, 1),
    ("lin_L0_geglu", 1, 1, 8192, 320, 2560, 1),
    ("lin_L0_ff1", 1, 1, 8192, 1280, 320, 1),
    ("lin_L1_geglu", 1, 1, 2048, 640, 5120, 1),
    ("lin_L2_geglu", 1, 1, 512, 1280, 10240, 1),
    ("lin_L2_ff1", 1, 1, 512, 5120, 1280, 1),
    ("lin_L0_proj", 1, 1, 8192, 320, 320, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--block-n", type=int, default=0)
    ap.add_argument("--splits", type=int, default=0)
    ap.add_argument("--only", default="")
    ap.add_argument("--rowmajor", action="store_true")
    ap.add_argument("--two-cta", type=int, default=0)
    ap.add_argument("--stats", action="store_true", help="also produce the per-channel statistics table (sdk_tc_gemm_set_stats)")
    ap.add_argument("--stamps", action="store_true", help="print in-kernel %%globaltimer stamps of CTA (0,0,0)")
    ap.add_argument("--warm", action="store_true", help="no L2 flush between launches; time 20 back-to-back launches")
    ap.add_argument("--shape", default="", help="custom: name,B,H,W,C,N,k")
    args = ap.parse_args()
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = SHAPES
    if args.shape:
        f = args.shape.split(",")
        shapes = [(f[0],) + tuple(int(x) for x in f[1:])]
    for name, B, H, W, Cc, N, k in shapes:
        if args.only and args.only not in name:
            continue
        a = torch.randn((B, H, W, Cc), device=dev).bfloat16()
        w = (torch.randn((N, k * k * Cc), device=dev) / math.sqrt(k * k * Cc)).bfloat16()
        bias = torch.randn((N,), device=dev)
        geglu = "geglu" in name
        out = torch.empty((B, H, W, N // 2 if geglu else N), device=dev, dtype=torch.bfloat16 if geglu else torch.float32)
        d = TcGemmDesc()
        if not args.rowmajor:
            w = w.view(N, k * k * Cc // 64, 64).permute(1, 0, 2).contiguous()
        d.w_kmajor = 0 if args.rowmajor else 1
        d.two_cta = args.two_cta
        d.a[0], d.w[0], d.C[0], d.ksize[0], d.nseg = a.data_ptr(), w.data_ptr(), Cc, k, 1
        d.B, d.H, d.W, d.N = B, H, W, N
        d.bias, d.out = bias.data_ptr(), out.data_ptr()
        d.out_dtype, d.geglu, d.block_n, d.splits = (BF16_T if geglu else F32_T), int(geglu), args.block_n, args.splits
        h = C.c_void_p()
        _lib.check(lib.sdk_tc_gemm_create(C.byref(d), C.byref(h)))
        info = (C.c_int * 9)()
        lib.sdk_tc_gemm_info(h, info, 9)
        ws = torch.zeros(max(int(lib.sdk_tc_gemm_workspace_bytes(h)), 256), dtype=torch.uint8, device=dev)
        lib.sdk_tc_gemm_set_workspace(h, ws.data_ptr())
        if args.stats and not geglu:
            cst = torch.zeros((B, N, 2), device=dev, dtype=torch.float64)
            if lib.sdk_tc_gemm_set_stats(h, cst.data_ptr()) != 0:
                print(f"{name}: statistics unsupported")
        _lib.check(lib.sdk_tc_gemm_launch(h, stream))
        torch.cuda.synchronize()
        if args.stamps:
            st = torch.zeros(16, dtype=torch.int64, device=dev)
            lib.sdk_tc_gemm_set_debug(h, st.data_ptr())
            for _ in range(3):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); _lib.check(lib.sdk_tc_gemm_launch(h, stream)); e1.record()
                torch.cuda.synchronize()
                v = st.tolist()
                print("   stamps (ns since entry): prologue %d | operands landed %d | last MMA issued %d | accumulator ready %d | epilogue done %d | exit %d ; event time %.1f us"
                      % tuple([v[i] - v[0] for i in range(1, 7)] + [e0.elapsed_time(e1) * 1e3]))
                if v[8]:
                    print("      epilogue detail (ns since accumulator ready): chunk0 done %d | chunk1 done %d | all chunks %d | partials fenced %d | ticket taken %d"
                          % tuple(v[i] - v[4] for i in (8, 9, 10, 11, 12)))
            lib.sdk_tc_gemm_set_debug(h, 0)
        ts = []
        if args.warm:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                _lib.check(lib.sdk_tc_gemm_launch(h, stream))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 20)
        for _ in range(0 if args.warm else args.iters):
            flush.zero_()                                   # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.sdk_tc_gemm_launch(h, stream))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        gf = 2.0 * B * H * W * N * k * k * Cc / 1e9
        wmb = N * k * k * Cc * 2 / 1e6
        us = min(ts)
        print(f"{name:16s} cg={info[8]} bn={info[0]:3d} splits={info[1]:2d} grid=({info[2]},{info[3]}) tile=({info[4]},{info[5]},{info[6]}) "
              f"{gf:7.2f} GF  w={wmb:6.1f} MB  {us:8.1f} us  {gf / us * 1e3:7.1f} TF/s  w-stream {wmb / us * 1e3:7.1f} GB/s")
        lib.sdk_tc_gemm_destroy(h)


if __name__ == "__main__":
    main()
