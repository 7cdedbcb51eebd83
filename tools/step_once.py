"""Minimal driver for profilers: a few eager (no CUDA graph) sampling steps of one BASELINE configuration, nothing else.

    ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip S --launch-count C --csv \
        --log-file out.csv python tools/step_once.py [--config 2] [--steps 4] [--batch B]

Prints the number of kernel launches of one step so that --launch-skip can be set to (steps - 1) * launches (+ set-up launches)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--graph", action="store_true")
    args = ap.parse_args()
    cfg = dict(bench.CONFIGS[args.config])
    if args.batch:
        cfg["batch"] = args.batch
    from stable_diffusion_pytorch_b200 import DDIMSampler, UNet
    from stable_diffusion_pytorch_b200.pipeline import DenoiseLoop
    dev = torch.device("cuda:0")
    arch, sd, lat, ctx = bench.build_oracle_inputs(cfg, cfg["batch"])
    net = UNet(attention_head_dim=arch["attention_head_dim"], cross_attention_dim=arch["cross_attention_dim"])
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    smp = DDIMSampler(prediction_type=cfg["ptype"])
    smp._set_inference_steps(50)
    loop = DenoiseLoop(net, smp, cfg["batch"], cfg["hw"], cfg["hw"], do_cfg=cfg["cfg"], use_cuda_graph=args.graph)
    with torch.no_grad():
        loop.reset(lat.to(dev), ctx.to(dev))
        torch.cuda.synchronize()
        print(f"launches_per_step {loop.launches_per_step}", flush=True)
        for _ in range(args.steps):
            loop.step()
        torch.cuda.synchronize()
    print("finite", bool(torch.isfinite(loop.latent).all()))


if __name__ == "__main__":
    main()
