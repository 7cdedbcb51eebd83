"""Run the fp32 sharded-loop scenario; prints whether the result is finite (used with SDB200_PDL_MASK to bisect kernel families)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import unet_oracle as UO
from stable_diffusion_pytorch_b200 import DDIMSampler, UNet
from stable_diffusion_pytorch_b200.pipeline import DenoiseLoop
dev = torch.device("cuda:0")
net = UNet(); net.load_state_dict(UO.make_state_dict(0, **UO.SD15), strict=True); net = net.to(dev).eval().set_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")
g = torch.Generator().manual_seed(1234)
lat = torch.randn((4, 4, 16, 16), generator=g); ctx = torch.randn((8, 77, 768), generator=g)
smp = DDIMSampler(); smp._set_inference_steps(10)
bad = 0
with torch.no_grad():
    for B in (4, 2, 1):
        for rep in range(2):
            loop = DenoiseLoop(net, smp, B, 16, 16)
            out = loop.run(lat[:B].to(dev), torch.cat([ctx[:B], ctx[4:4 + B]]).to(dev), steps=4)
            ok = bool(torch.isfinite(out).all())
            bad += not ok
            print(f"B={B} rep={rep} finite={ok} norm={float(out.norm()):.6f}")
print("BAD" if bad else "OK", os.environ.get("SDB200_PDL_MASK"), os.environ.get("SDB200_PDL"))
