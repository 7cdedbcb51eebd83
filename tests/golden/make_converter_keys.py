"""Extract the (reference key -> diffusers key) tables of the UNMODIFIED reference converter
(utils/model_converter.py: load_unet_weights_v1_5 / load_unet_weights_v2_1) by running it on a recording fake
checkpoint.  Writes tests/golden/converter_keys_{v15,v21}.json.  Run in the build container only."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("SD_REFERENCE", "/root/reference"))
import utils.model_converter as MC      # noqa: E402  (reference)


class Recorder(dict):
    def __init__(self):
        super().__init__()
        self.names = []

    def __getitem__(self, k):
        self.names.append(k)
        return torch.full((1, 1), float(len(self.names) - 1))


for tag, fn in (("v15", MC.load_unet_weights_v1_5), ("v21", MC.load_unet_weights_v2_1)):
    rec = Recorder()
    MC.load_file = lambda path, device=None, rec=rec: rec
    out = fn("fake.safetensors", device="cpu")["unet"]
    table = {k: [rec.names[int(v.flatten()[0])], v.dim() == 4] for k, v in out.items()}
    with open(os.path.join(HERE, f"converter_keys_{tag}.json"), "w") as f:
        json.dump(table, f, indent=0, sort_keys=True)
    print(tag, len(table), sum(1 for v in table.values() if v[1]), "unsqueezed")
