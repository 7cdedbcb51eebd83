"""Launch the GroupNorm-apply kernel (statistics from per-channel sums) on UNet shapes for ncu / event timing.
   python tools/prof_gn.py [B]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stable_diffusion_pytorch_b200 import _lib  # noqa: E402
from stable_diffusion_pytorch_b200._lib import BF16_T  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, HW, C0, C1, raw in (("L0_320", 4096, 320, 0, False), ("L0_640+320_raw", 4096, 640, 320, True), ("L1_640", 1024, 640, 0, False),
                                  ("L2_1280+1280_raw", 256, 1280, 1280, True), ("L3_1280", 64, 1280, 0, False)):
        s0 = torch.randn((B, HW, C0), device=dev)
        s1 = torch.randn((B, HW, C1), device=dev) if C1 else None
        Ct = C0 + C1
        g, b = torch.ones(Ct, device=dev), torch.zeros(Ct, device=dev)
        cs0 = torch.zeros((B, C0, 2), device=dev, dtype=torch.float64)
        cs1 = torch.zeros((B, max(C1, 1), 2), device=dev, dtype=torch.float64)
        lib.sdk_channel_stats(s0.data_ptr(), B, HW, C0, cs0.data_ptr(), stream)
        if C1:
            lib.sdk_channel_stats(s1.data_ptr(), B, HW, C1, cs1.data_ptr(), stream)
        out = torch.empty((B, HW, Ct), device=dev, dtype=torch.bfloat16)
        rawt = torch.empty((B, HW, Ct), device=dev, dtype=torch.bfloat16) if raw else None

        def run():
            _lib.check(lib.sdk_groupnorm_apply_cs(s0.data_ptr(), C0, cs0.data_ptr(), s1.data_ptr() if C1 else 0, C1, cs1.data_ptr() if C1 else 0,
                                                  B, HW, 1e-5, g.data_ptr(), b.data_ptr(), 1, out.data_ptr(), rawt.data_ptr() if raw else 0, BF16_T, stream))
        run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        nbytes = B * HW * Ct * (4 + 2 + (2 if raw else 0))
        # warm: 20 back-to-back launches (data L2-resident when it fits), and the same with precomputed group statistics
        stats = torch.zeros((B, 32, 2), device=dev)
        stats[..., 1] = 1.0

        def run_pre():
            _lib.check(lib.sdk_groupnorm_apply(s0.data_ptr(), C0, s1.data_ptr() if C1 else 0, C1, B, HW, stats.data_ptr(), g.data_ptr(), b.data_ptr(), 1,
                                               out.data_ptr(), rawt.data_ptr() if raw else 0, BF16_T, stream))
        warm = []
        for fn in (run, run_pre):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record(); torch.cuda.synchronize()
            warm.append(e0.elapsed_time(e1) * 1e3 / 20)
        print(f"{name:18s} B={B}: cold {min(ts):7.1f} us {nbytes / min(ts) / 1e6:6.2f} TB/s | warm x20: {warm[0]:6.1f} us {nbytes / warm[0] / 1e6:6.2f} TB/s, "
              f"with precomputed group stats {warm[1]:6.1f} us  ({nbytes / 1e6:.1f} MB)")


if __name__ == "__main__":
    main()
