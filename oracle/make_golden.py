"""TEST INFRASTRUCTURE — writes the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (/root/reference/models/pretrain_models.py, imported through oracle/stubs) on seeded
synthetic inputs with externally supplied mask noise.  Run in the build container only:

    python -m oracle.make_golden

The reference has no golden vectors of its own; these freeze what its code computes
here (torch CPU fp32) so the oracle restatement and the CUDA path can be checked anywhere,
including on the GPU box where /root/reference does not exist.

Cases
  tiny_*    reduced model (dim 128, 2+1 layers, 16+2x16 tokens); tiny_nt2 stores its weights in the
            fixture (RNG-independent anchor), the others derive them from the seed like canon_*.
  canon_*   the canonical train.py model at B=2; weights come from
            oracle.init_state_dict(cfg, seed=0) (a checksum is stored to detect RNG drift).
Each case stores inputs, noise, mask indices, loss, encoder/decoder activations, embeddings,
per-parameter gradient norms, a few full gradients and the parameters after one
clip(0.5)+AdamW(lr=1e-4) step.
"""
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
from oracle import vtmae_oracle as O  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

TINY = dict(image_size=(32, 32), tactile_size=(16, 16), image_patch_size=8, tactile_patch_size=4,
            dim=128, depth=2, heads=2, dim_head=64, mlp_dim=256, image_channels=3, tactile_channels=3,
            frame_stack=1, decoder_dim=128, decoder_depth=1, decoder_heads=2, decoder_dim_head=64,
            masking_ratio=0.75)

CASES = {
    "tiny_nt2": (dict(TINY, num_tactiles=2), 3, True),
    "tiny_vision": (dict(TINY, num_tactiles=0), 3, False),
    "tiny_ecm": (dict(TINY, num_tactiles=2, early_conv_masking=True, image_size=(64, 64),
                      tactile_size=(32, 32), masking_ratio=0.9), 2, False),
    "tiny_learnedpos": (dict(TINY, num_tactiles=2, use_sincosmod_encodings=False), 2, False),
    "canon_nt2": (dict(), 2, False),
    "canon_vision": (dict(num_tactiles=0), 2, False),
}

FULL_GRAD_KEYS = ("mask_token", "encoder_modality_embedding.weight", "decoder_modality_embedding.weight",
                  "to_tactiles.bias", "decoder.norm.weight", "encoder.transformer.layers.0.0.to_qkv.weight")


def synth_inputs(cfg: O.VTMAEConfig, batch: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    x = {"image": torch.rand(batch, cfg.image_channels, *O._pair(cfg.image_size), generator=g)}
    for i in range(cfg.num_tactiles):
        x[f"tactile{i + 1}"] = torch.rand(batch, cfg.tactile_channels, *O._pair(cfg.tactile_size), generator=g)
    n = cfg.n_img + cfg.num_tactiles * cfg.n_tac
    noise = O.tie_free_noise(batch, n, g, [cfg.n_img] + [cfg.n_tac] * cfg.num_tactiles)
    return x, noise


def weights_checksum(sd) -> float:
    return float(sum(v.double().abs().sum() for k, v in sorted(sd.items())))


def make_case(name: str, overrides: dict, batch: int, store_weights: bool):
    cfg = O.VTMAEConfig(**overrides)
    sd0 = O.init_state_dict(cfg, seed=0)
    mae = R.build_reference_model(cfg, seed=0)
    missing = mae.load_state_dict(O.expand_aliases(sd0), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    x, noise = synth_inputs(cfg, batch, seed=1234)
    out = {"config_json": np.frombuffer(json.dumps(asdict(cfg)).encode(), dtype=np.uint8),
           "batch": np.int64(batch), "noise": noise.numpy(),
           "weights_checksum": np.float64(weights_checksum(sd0))}
    for k, v in x.items():
        out["x." + k] = v.numpy()
    if store_weights:
        for k, v in sd0.items():
            out["w." + k] = v.numpy()

    # ---- forward + backward through the reference, hooks capture activations
    captured = {}
    h1 = mae.encoder.transformer.register_forward_hook(lambda m, i, o: captured.__setitem__("enc", (i[0], o)))
    h2 = mae.decoder.register_forward_hook(lambda m, i, o: captured.__setitem__("dec", (i[0], o)))
    mae.train()
    mae.zero_grad()
    with R.injected_noise(R.split_noise(noise, cfg, True, cfg.num_tactiles > 0)):
        loss = mae(x)
    loss.backward()
    h1.remove(); h2.remove()
    out["loss"] = loss.detach().numpy()
    out["enc_in"] = captured["enc"][0].detach().numpy()
    out["encoded"] = captured["enc"][1].detach().numpy()
    if store_weights:
        out["decoder_in"] = captured["dec"][0].detach().numpy()
    out["decoded"] = captured["dec"][1].detach().numpy()
    # indices: recompute with the reference's own statement (argsort of the injected noise)
    nm_img, nm_tac = O.mask_counts(cfg.masking_ratio, cfg.n_img, cfg.num_tactiles * cfg.n_tac, cfg.num_tactiles)
    masked, unmasked = [], []
    off = 0
    for seg, nm in [(cfg.n_img, nm_img)] + [(cfg.n_tac, nm_tac)] * cfg.num_tactiles:
        perm = noise[:, off:off + seg].argsort(dim=-1) + off
        masked.append(perm[:, :nm]); unmasked.append(perm[:, nm:]); off += seg
    out["masked_indices"] = torch.cat(masked, 1).numpy()
    out["unmasked_indices"] = torch.cat(unmasked, 1).numpy()
    # cross-check: the gather the reference performed equals tokens at those indices
    named = dict(mae.named_parameters())
    for k, p in named.items():
        if k.startswith("encoder.image_to_patch_embedding") or k.startswith("encoder.tactile_to_patch_embedding"):
            continue
        out["gnorm." + k] = np.float64(0.0 if p.grad is None else p.grad.double().norm().item())
        out["ghas." + k] = np.bool_(p.grad is not None)
    for k in FULL_GRAD_KEYS:
        if k in named and named[k].grad is not None:
            out["grad." + k] = named[k].grad.numpy().copy()

    # ---- one optimizer step exactly as train_iterations does (pretrain_models.py:707-711)
    opt = torch.optim.AdamW(mae.parameters(), lr=1e-4)
    total_norm = torch.nn.utils.clip_grad_norm_(mae.parameters(), 0.5)
    opt.step()
    out["grad_total_norm"] = total_norm.detach().numpy()
    sd1 = mae.state_dict()
    for k in FULL_GRAD_KEYS + ("to_pixels.weight",):
        if k in sd1:
            out["after." + k] = sd1[k].numpy().copy()
    out["after_checksum"] = np.float64(weights_checksum(O.canonical(sd1)))
    with R.injected_noise(R.split_noise(noise, cfg, True, cfg.num_tactiles > 0)):
        out["loss_after_step"] = mae(x).detach().numpy()

    # ---- embeddings with the ORIGINAL weights (rollout path, pretrain_models.py:588-668)
    mae.load_state_dict(O.expand_aliases(sd0), strict=True)
    with torch.no_grad():
        out["embeddings"] = mae.get_embeddings(x, eval=False)[:1].numpy()
        if cfg.num_tactiles > 0:
            out["embeddings_vision_only"] = mae.get_embeddings(x, eval=False, use_tactile=False)[:1].numpy()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / f"{name}.npz", **out)
    print(f"{name}: loss={float(loss):.6f} total_norm={float(total_norm):.4f} "
          f"-> {(GOLDEN / (name + '.npz')).stat().st_size / 1e6:.2f} MB")


def main():
    if not R.reference_available():
        raise SystemExit("reference tree not present; golden fixtures are generated in the build container")
    torch.set_num_threads(8)
    for name, (ov, b, store) in CASES.items():
        make_case(name, ov, b, store)


if __name__ == "__main__":
    main()
