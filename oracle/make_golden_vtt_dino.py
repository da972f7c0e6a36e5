"""TEST INFRASTRUCTURE — writes tests/golden/vtt_dino/*.npz by running the UNMODIFIED reference
/root/reference/models/VTT.py::VTT (imported through oracle/stubs) on seeded inputs.  Build container only:

    python -m oracle.make_golden_vtt_dino

Each case stores the configuration, the weights (perturbed away from the timm init so every term matters), the
inputs, the shared keep-index masks, the three outputs of forward_features and, for a fixed scalar objective, the
gradient norm of every parameter plus two full gradients."""
from __future__ import annotations

import json
import sys
from dataclasses import asdict
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import reference_adapter as R  # noqa: E402
from oracle import vtt_dino_oracle as VD  # noqa: E402

OUT = ROOT / "tests" / "golden" / "vtt_dino"
CASES = {   # name: (config overrides, batch, number of masks, kept tokens per mask)
    "regs1_nomask": (dict(num_register_tokens=1), 2, 0, 0),
    "regs1_masks2": (dict(num_register_tokens=1), 2, 2, 24),
    "regs0_mask1": (dict(num_register_tokens=0, heads=3, depth=2), 3, 1, 9),
}
BASE = dict(image_size=(64, 64), tactile_size=(32, 32), image_patch_size=8, tactile_patch_size=4, dim=128, depth=1,
            heads=2, dim_head=64, mlp_dim=256, image_channels=3, tactile_channels=3, num_tactiles=2)


def objective(out, w1, w2):
    return (out["x_norm_patchtokens"] * w1).sum() + (out["x_norm_regtokens"] ** 2).sum() + 0.5 * (out["x_prenorm"] * w2).sum()


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    for name, (over, B, n_masks, keep) in CASES.items():
        cfg = VD.VTTDinoConfig(**dict(BASE, **over))
        ref = R.build_reference_vtt_dino(cfg, seed=11)
        g = torch.Generator().manual_seed(5)
        with torch.no_grad():
            for k, p in ref.named_parameters():
                if k == "register_tokens":
                    p.copy_(torch.randn(p.shape, generator=g) * 0.5)
                elif k.endswith(".bias"):
                    p.copy_(torch.randn(p.shape, generator=g) * 0.05)
                elif p.dim() == 1:
                    p.add_(torch.randn(p.shape, generator=g) * 0.1)
        x = {"image": torch.rand(B, 3, 64, 64, generator=g), "tactile1": torch.rand(B, 3, 32, 32, generator=g),
             "tactile2": torch.rand(B, 3, 32, 32, generator=g)}
        masks = [torch.stack([torch.randperm(64, generator=g)[:keep] for _ in range(B)]) for _ in range(n_masks)] or None
        out = ref.forward_features(x, masks)
        w1 = torch.randn(out["x_norm_patchtokens"].shape, generator=g)
        w2 = torch.randn(out["x_prenorm"].shape, generator=g)
        objective(out, w1, w2).backward()
        blob = {"config_json": np.frombuffer(json.dumps(asdict(cfg)).encode(), dtype=np.uint8), "batch": np.int64(B),
                "n_masks": np.int64(n_masks), "w1": w1.numpy(), "w2": w2.numpy()}
        for k, v in ref.state_dict().items():
            blob["w." + k] = v.detach().numpy()
        for k, v in x.items():
            blob["x." + k] = v.numpy()
        for i, m in enumerate(masks or []):
            blob[f"mask.{i}"] = m.numpy()
        for k in ("x_norm_regtokens", "x_norm_patchtokens", "x_prenorm"):
            blob["out." + k] = out[k].detach().numpy()
        for k, p in ref.named_parameters():
            blob["ghas." + k] = np.bool_(p.grad is not None and float(p.grad.abs().max()) > 0)
            if p.grad is not None:
                blob["gnorm." + k] = np.float64(p.grad.double().norm())
        for k in ("register_tokens", "norm.weight", "tactile_to_patch_embedding_2.2.weight", "transformer.layers.0.0.to_qkv.weight"):
            p = dict(ref.named_parameters()).get(k)
            if p is not None and p.grad is not None:
                blob["grad." + k] = p.grad.numpy()
        np.savez_compressed(OUT / f"{name}.npz", **blob)
        print(name, "->", OUT / f"{name}.npz")


if __name__ == "__main__":
    main()
