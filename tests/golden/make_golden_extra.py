"""Round-2 fixtures from the EXECUTED reference (same recipe as make_golden.py; run in the build container only):

    python tests/golden/make_golden_extra.py   ->  tests/golden/unet_extra.npz

  tiny_cba   Unet(resnet_block_order='conv_bn_act'): final_conv = [ResnetBlock, Conv2d 1x1] (reference modules/unet.py:112-116)
  cfg2_cls   the CIFAR-shape U-Net (dim 128, mults 1,2,2,2) with num_classes=10: class embedding added to the stem output
             (reference modules/unet.py:118-120,134-141), labels [3, 10 (= the padding / null row)]
  cfg2_b4    teacher-forced eps of the CIFAR-shape U-Net at batch 4 with mixed timesteps (a second, larger pin of the tcgen05 engine)

Every output is asserted equal to the CPU oracle (oracle/ref_port.py) before it is stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests", "_shim"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)

import diffusion_model_nemo.modules as M  # noqa: E402  (the unmodified reference)
from oracle import ref_port as O  # noqa: E402

torch.set_num_threads(8)

CFGS = {
    "tiny_cba": (dict(dim=32, dim_mults=[1, 2], channels=3, groups=8, order="conv_bn_act"), 16, 2),
    "cfg2_cls": (dict(dim=128, dim_mults=[1, 2, 2, 2], channels=3, groups=8, num_classes=10), 32, 2),
    "cfg2_b4": (dict(dim=128, dim_mults=[1, 2, 2, 2], channels=3, groups=8), 32, 4),
}


def main():
    out = {}
    for name, (cfg, size, b) in CFGS.items():
        sd = O.random_state_dict(cfg, seed=0)
        ref = M.Unet(input_dim=None, dim=cfg["dim"], dim_mults=cfg["dim_mults"], channels=cfg["channels"], use_convnext=False,
                     resnet_block_groups=cfg["groups"], dropout=0.0, num_classes=cfg.get("num_classes"),
                     resnet_block_order=cfg.get("order", "bn_act_conv")).eval()
        res = ref.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        g = torch.Generator().manual_seed(11)
        x = torch.randn(b, cfg["channels"], size, size, generator=g)
        t = torch.tensor([7, 513, 999, 0][:b])
        kw = {"classes": torch.tensor([3, 10][:b])} if cfg.get("num_classes") is not None else {}
        with torch.no_grad():
            y_ref = ref(x, t, **kw)
        y = O.unet_forward(sd, cfg, x, t.float(), kw.get("classes"))
        d = float((y - y_ref).abs().max())
        print(f"{name}: oracle vs executed reference max-abs diff {d:.3e}")
        assert d == 0.0
        out[f"{name}/t"] = t.numpy()
        out[f"{name}/y"] = y_ref.numpy()
        if kw:
            out[f"{name}/classes"] = kw["classes"].numpy()
    np.savez_compressed(os.path.join(HERE, "unet_extra.npz"), **out)


if __name__ == "__main__":
    main()
