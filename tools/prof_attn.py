"""Launch the bf16 flash-attention kernel on the UNet's shapes (UNet batch 2) for ncu / event timing."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stable_diffusion_pytorch_b200 import _lib  # noqa: E402

SHAPES = [("self_L0", 2, 8, 4096, 4096, 40), ("self_L1", 2, 8, 1024, 1024, 80), ("self_L2", 2, 8, 256, 256, 160),
          ("cross_L0", 2, 8, 4096, 77, 40), ("cross_L1", 2, 8, 1024, 77, 80), ("sd21_L0", 2, 5, 9216, 9216, 64)]


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for name, B, Hh, Sq, Sk, D in SHAPES:
        if only and only not in name:
            continue
        Cc = Hh * D
        q = torch.randn((B, Sq, Cc), device=dev).bfloat16()
        kv = torch.randn((B, Sk, 2 * Cc), device=dev).bfloat16()
        out = torch.empty((B, Sq, Cc), device=dev, dtype=torch.bfloat16)
        args = (q.data_ptr(), Cc, Sq * Cc, kv.data_ptr(), 2 * Cc, Sk * 2 * Cc, kv.data_ptr() + Cc * 2, 2 * Cc, Sk * 2 * Cc,
                out.data_ptr(), Cc, Sq * Cc, B, Hh, Sq, Sk, D, float(D ** -0.5), stream)
        use_tc = os.environ.get("ATTN_TC", "0") == "1" and D in (40, 64)
        if use_tc:
            h = C.c_void_p()
            _lib.check(lib.sdk_attention_tc_create(*args[:-1], C.byref(h)))
            run = lambda: _lib.check(lib.sdk_attention_tc_launch(h, stream))
        else:
            run = lambda: _lib.check(lib.sdk_attention_bf16(*args))
        run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        name = name + ("[tc]" if use_tc else "")
        gf = 4.0 * B * Hh * Sq * Sk * D / 1e9
        print(f"{name:14s} B={B} H={Hh} Sq={Sq} Sk={Sk} D={D}: {min(ts):8.1f} us  {gf / min(ts) * 1e3:7.1f} TF/s  exps/us={B * Hh * Sq * Sk / min(ts) / 1e6:6.2f} G/s")


if __name__ == "__main__":
    main()
