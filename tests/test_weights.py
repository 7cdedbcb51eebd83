"""CPU suite: the rule-based diffusers -> reference key mapper reproduces the reference's hand-unrolled converter tables
(tests/golden/converter_keys_*.json, extracted from utils/model_converter.py by make_converter_keys.py)."""
import json
import os

import pytest
import torch

from stable_diffusion_pytorch_b200.arch import build_arch, param_spec
from stable_diffusion_pytorch_b200.weights import convert_state_dict, key_map

CFGS = {"v15": dict(attention_head_dim=8, cross_attention_dim=768), "v21": dict(attention_head_dim=[5, 10, 20, 20], cross_attention_dim=1024)}


@pytest.mark.parametrize("tag", ["v15", "v21"])
def test_key_map_matches_reference_converter(golden_dir, tag):
    table = json.load(open(os.path.join(golden_dir, f"converter_keys_{tag}.json")))
    a = build_arch(**CFGS[tag])
    names = [n for n, _ in param_spec(a)]
    assert sorted(names) == sorted(table)                               # same 686 reference keys
    km = key_map(a, names)
    assert {k: v[0] for k, v in table.items()} == km                    # same diffusers source for every key
    # the reference unsqueezes exactly the SD-2.1 linear proj_in / proj_out weights
    unsq = sorted(k for k, v in table.items() if v[1])
    assert all(k.endswith(("conv_input.weight", "conv_output.weight")) for k in unsq)
    assert (len(unsq) == 32) == (tag == "v21")


def test_convert_state_dict_shapes():
    a = build_arch(**CFGS["v21"])
    spec = param_spec(a)
    km = key_map(a, [n for n, _ in spec])
    fake = {}
    for n, shape in spec:                                               # diffusers side: proj_in/out are Linear [C, C]
        shp = shape[:2] if n.endswith(("conv_input.weight", "conv_output.weight")) else shape
        fake[km[n]] = torch.empty(shp, device="meta")
    out = convert_state_dict(a, fake, spec)
    assert all(tuple(out[n].shape) == tuple(s) for n, s in spec)
    fake[km[spec[0][0]]] = torch.empty((3, 3), device="meta")
    with pytest.raises(ValueError):
        convert_state_dict(a, fake, spec)
