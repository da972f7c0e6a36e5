"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference module
/root/reference/models/pretrain_models.py (when that tree is present, i.e. in the build container,
never on the GPU box) behind the restated third-party stubs in oracle/stubs, so the oracle and the
golden vectors can be pinned against the reference's own code.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import sys
from pathlib import Path

import torch

REFERENCE_ROOT = Path("/root/reference")
STUBS = Path(__file__).resolve().parent / "stubs"


def reference_available() -> bool:
    return (REFERENCE_ROOT / "models" / "pretrain_models.py").exists()


def load_reference_module():
    """Returns the imported reference `models.pretrain_models` module (or raises if absent)."""
    if not reference_available():
        raise FileNotFoundError("reference tree not present")
    for p in (str(STUBS), str(REFERENCE_ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module("models.pretrain_models")


def build_reference_model(cfg, seed: int = 0):
    """Constructs reference VTT + VTMAE for an oracle VTMAEConfig (ctor calls as train.py:128-153)."""
    ref = load_reference_module()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        enc = ref.VTT(image_size=cfg.image_size, tactile_size=cfg.tactile_size,
                      image_patch_size=cfg.image_patch_size, tactile_patch_size=cfg.tactile_patch_size,
                      dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, mlp_dim=cfg.mlp_dim,
                      image_channels=cfg.image_channels, tactile_channels=cfg.tactile_channels,
                      dim_head=cfg.dim_head, num_tactiles=cfg.num_tactiles, frame_stack=cfg.frame_stack)
        mae = ref.VTMAE(encoder=enc, decoder_dim=cfg.decoder_dim, masking_ratio=cfg.masking_ratio,
                        decoder_depth=cfg.decoder_depth, decoder_heads=cfg.decoder_heads,
                        decoder_dim_head=cfg.decoder_dim_head, num_tactiles=cfg.num_tactiles,
                        early_conv_masking=cfg.early_conv_masking,
                        use_sincosmod_encodings=cfg.use_sincosmod_encodings, frame_stack=cfg.frame_stack)
    return mae


@contextlib.contextmanager
def injected_noise(segments):
    """Replaces torch.rand by a function that pops the pre-drawn (B, n) tensors in call order
    (image, tactile1, tactile2: pretrain_models.py:229,237)."""
    queue = list(segments)
    real = torch.rand

    def fake(*size, **kw):
        t = queue.pop(0)
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        assert tuple(t.shape) == shape, (t.shape, shape)
        return t.clone()

    torch.rand = fake
    try:
        yield
    finally:
        torch.rand = real


def split_noise(noise, cfg, use_vision=True, use_tactile=True):
    segs, off = [], 0
    if use_vision:
        segs.append(noise[:, off:off + cfg.n_img]); off += cfg.n_img
    if use_tactile:
        for _ in range(cfg.num_tactiles):
            segs.append(noise[:, off:off + cfg.n_tac]); off += cfg.n_tac
    return segs


def load_reference_vtt_module():
    """The unmodified /root/reference/models/VTT.py (DINO-side encoder), behind the omegaconf / lightning shims."""
    if not (REFERENCE_ROOT / "models" / "VTT.py").exists():
        raise FileNotFoundError("reference tree not present")
    for p in (str(STUBS), str(REFERENCE_ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("models.VTT")


def build_reference_vtt_dino(cfg, seed: int = 0):
    """Reference models/VTT.py::VTT for an oracle VTTDinoConfig (ctor call as models/ppo_dino.py:563-578)."""
    ref = load_reference_vtt_module()
    torch.manual_seed(seed)
    return ref.VTT(image_size=cfg.image_size, tactile_size=cfg.tactile_size, image_patch_size=cfg.image_patch_size,
                   tactile_patch_size=cfg.tactile_patch_size, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads,
                   mlp_dim=cfg.mlp_dim, num_tactiles=cfg.num_tactiles, image_channels=cfg.image_channels,
                   tactile_channels=cfg.tactile_channels, dim_head=cfg.dim_head,
                   num_register_tokens=cfg.num_register_tokens, pos_embed_fn="sinusoidal")
