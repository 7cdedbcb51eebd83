"""One GEGLU-in GEMM (level 0: M = 8192, K = 320, N = 2560, block_n 256) launched a few times -- target of `ncu --set full`."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stable_diffusion_pytorch_b200 import _lib
from stable_diffusion_pytorch_b200._lib import BF16_T, TcGemmDesc
dev = torch.device("cuda:0")
lib = _lib.lib()
M, K, N = 8192, 320, 2560
geglu = int(os.environ.get("GEGLU", "1"))
bn = int(os.environ.get("BN", "256"))
a = torch.randn((M, K), device=dev).bfloat16()
w = (torch.randn((N, K), device=dev) / math.sqrt(K)).bfloat16()
wk = w.view(N, K // 64, 64).permute(1, 0, 2).contiguous()
bias = torch.randn((N,), device=dev) * 0.1
out = torch.empty((M, N // 2 if geglu else N), device=dev, dtype=torch.bfloat16)
d = TcGemmDesc()
d.w_kmajor, d.w_const = 1, 1
d.a[0], d.w[0], d.C[0], d.ksize[0], d.nseg = a.data_ptr(), wk.data_ptr(), K, 1, 1
d.B, d.H, d.W, d.N = 1, 1, M, N
d.bias, d.out, d.out_dtype, d.geglu, d.block_n, d.splits = bias.data_ptr(), out.data_ptr(), BF16_T, geglu, bn, 1
h = C.c_void_p()
_lib.check(lib.sdk_tc_gemm_create(C.byref(d), C.byref(h)))
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(5):
    _lib.check(lib.sdk_tc_gemm_launch(h, s))
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
