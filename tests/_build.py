"""Test helper: build the product modules from an oracle VTMAEConfig."""
import torch


def build_product(cfg, device="cuda", weights=None):
    from m3l_b200 import VTT, VTMAE
    from oracle import vtmae_oracle as O
    enc = VTT(image_size=cfg.image_size, tactile_size=cfg.tactile_size, image_patch_size=cfg.image_patch_size,
              tactile_patch_size=cfg.tactile_patch_size, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads,
              mlp_dim=cfg.mlp_dim, image_channels=cfg.image_channels, tactile_channels=cfg.tactile_channels,
              dim_head=cfg.dim_head, num_tactiles=cfg.num_tactiles, frame_stack=cfg.frame_stack)
    mae = VTMAE(encoder=enc, decoder_dim=cfg.decoder_dim, masking_ratio=cfg.masking_ratio,
                decoder_depth=cfg.decoder_depth, decoder_heads=cfg.decoder_heads,
                decoder_dim_head=cfg.decoder_dim_head, num_tactiles=cfg.num_tactiles,
                early_conv_masking=cfg.early_conv_masking, use_sincosmod_encodings=cfg.use_sincosmod_encodings,
                frame_stack=cfg.frame_stack)
    if weights is not None:
        mae.load_state_dict(O.expand_aliases({k: v.detach().clone() for k, v in weights.items()}), strict=True)
    return mae.to(device)
