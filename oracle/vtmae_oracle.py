"""TEST INFRASTRUCTURE — CPU oracle for the VTMAE/VTT hot path.  NOT part of the product.

Only tests/, __graft_entry__.smoke() and bench.py's `cpu_baseline` / `--impl reference` legs may
import this module.  The shipped package (m3l_b200/) never does; it fails loudly without its CUDA
library.

This is a functional restatement (plain torch fp32 ops over a flat `state_dict`) of the
reference algorithm; each function cites the reference lines it follows
(paths relative to /root/reference):

  vt_load            utils/pretrain_utils.py:7-57
  patchify           models/pretrain_models.py:768,775   (einops Rearrange
                     'b c (h p1) (w p2) -> b (h w) (p1 p2 c)')
  token embedding    models/pretrain_models.py:194,198,202-219
  mask sampling      models/pretrain_models.py:223-248
  forward / loss     models/pretrain_models.py:146-342
  get_embeddings     models/pretrain_models.py:588-668
  EarlyCNN           models/pretrain_models.py:37-56
  extractor          models/pretrain_models.py:819-841
  train step         models/pretrain_models.py:707-711 (+ torch.optim.AdamW, clip_grad_norm_)
  Transformer        vit-pytorch==1.6.4 (requirements.txt:160)          } third-party, absent from
  sin-cos table      positional-encodings==6.0.1 (requirements.txt:107) } the tree: restated

PINNING STATUS.  The reference has no tests, golden vectors or fixtures for this path
(SURVEY.md §4, §8c), and the two third-party packages cannot be installed offline, so:
  * everything that lives in /root/reference (forward, masking, loss, get_embeddings, EarlyCNN,
    vt_load, train step) IS pinned: tests/test_oracle_vs_reference.py imports the unmodified
    reference file through oracle/stubs and checks this restatement against it bit-for-bit on
    CPU, and oracle/make_golden.py freezes reference outputs into tests/golden/;
  * the Transformer block arithmetic and the sin-cos table are "PARITY UNPINNED": restated from
    the packages' published algorithms (oracle/stubs/vit_pytorch, oracle/stubs/
    positional_encodings), with only their call signature and parameter names pinned by the
    reference tree.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# configuration (constructor arguments of VTT / VTMAE: pretrain_models.py:60-73,719-734)
# --------------------------------------------------------------------------------------------
def _pair(t):
    return t if isinstance(t, tuple) else (t, t)


@dataclass
class VTMAEConfig:
    image_size: Tuple[int, int] = (64, 64)
    tactile_size: Tuple[int, int] = (32, 32)
    image_patch_size: int = 8
    tactile_patch_size: int = 4
    dim: int = 256
    depth: int = 4
    heads: int = 4
    dim_head: int = 64
    mlp_dim: int = 512
    image_channels: int = 12          # 3 * frame_stack
    tactile_channels: int = 12
    num_tactiles: int = 2
    frame_stack: int = 4
    decoder_dim: int = 256
    decoder_depth: int = 3
    decoder_heads: int = 4
    decoder_dim_head: int = 64
    masking_ratio: float = 0.95
    early_conv_masking: bool = False
    use_sincosmod_encodings: bool = True

    # derived geometry -----------------------------------------------------------------
    @property
    def image_grid(self):
        (h, w), (p, q) = _pair(self.image_size), _pair(self.image_patch_size)
        return h // p, w // q

    @property
    def tactile_grid(self):
        (h, w), (p, q) = _pair(self.tactile_size), _pair(self.tactile_patch_size)
        return h // p, w // q

    @property
    def n_img(self):
        return self.image_grid[0] * self.image_grid[1]

    @property
    def n_tac(self):  # per sensor
        return self.tactile_grid[0] * self.tactile_grid[1]

    @property
    def p_img(self):
        p, q = _pair(self.image_patch_size)
        return self.image_channels * p * q

    @property
    def p_tac(self):
        p, q = _pair(self.tactile_patch_size)
        return self.tactile_channels * p * q

    @property
    def decoder_mlp_dim(self):
        return 4 * self.decoder_dim  # pretrain_models.py:113


# --------------------------------------------------------------------------------------------
# elementary pieces
# --------------------------------------------------------------------------------------------
def vt_load(x: dict, frame_stack: int = 1, image_normalization=(0, 1), tactile_normalization=(-1, 1)):
    """obs dict -> model input dict (utils/pretrain_utils.py:7-57). Consumes `x['tactile']`."""
    out = {}
    if "image" in x:
        img = torch.as_tensor(x["image"], dtype=torch.float32)
        if img.dim() == 3:
            img = img[None]
        assert img.shape[-1] == 3 * frame_stack
        img = img.permute(0, 3, 1, 2)
        out["image"] = (img - image_normalization[0]) / (image_normalization[1] - image_normalization[0])
    if "tactile" in x:
        tac = torch.as_tensor(x["tactile"], dtype=torch.float32)
        if tac.dim() == 3:
            tac = tac[None]
        ch = tac.shape[1]
        assert ch in (3 * frame_stack, 6 * frame_stack, 12 * frame_stack)
        per_frame = ch // frame_stack
        base = []
        for i in range(frame_stack):
            base += [i * per_frame + 0, i * per_frame + 1, i * per_frame + 2]
        base = torch.tensor(base)
        for s in range(per_frame // 3):
            t = tac[:, base + 3 * s]
            out[f"tactile{s + 1}"] = (t - tactile_normalization[0]) / (
                tactile_normalization[1] - tactile_normalization[0])
    return out


def patchify(x: torch.Tensor, p1: int, p2: int) -> torch.Tensor:
    """'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' (pretrain_models.py:768,775)."""
    b, c, H, W = x.shape
    h, w = H // p1, W // p2
    x = x.reshape(b, c, h, p1, w, p2).permute(0, 2, 4, 3, 5, 1)  # b h w p1 p2 c
    return x.reshape(b, h * w, p1 * p2 * c)


def sincos_2d(nx: int, ny: int, channels: int) -> torch.Tensor:
    """positional_encodings.PositionalEncoding2D(channels)(zeros(1,nx,ny,channels)).flatten(1,2)
    -> (nx*ny, channels).  Restated (package absent): see oracle/stubs/positional_encodings."""
    ch = int(math.ceil(channels / 4) * 2)
    inv_freq = 1.0 / (10000 ** (torch.arange(0, ch, 2).float() / ch))

    def enc(n):
        ang = torch.arange(n, dtype=torch.float32)[:, None] * inv_freq[None, :]
        return torch.stack((ang.sin(), ang.cos()), dim=-1).flatten(-2, -1)  # (n, ch)

    emb = torch.zeros(nx, ny, 2 * ch)
    emb[:, :, :ch] = enc(nx)[:, None, :]
    emb[:, :, ch:] = enc(ny)[None, :, :]
    return emb[:, :, :channels].reshape(nx * ny, min(channels, 2 * ch))


def layer_norm(x, sd, prefix, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def linear(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def transformer(x, sd, prefix, depth, heads, dim_head):
    """vit_pytorch.vit.Transformer (v1.6.4, restated): pre-norm MHA + FF residual blocks, final LN."""
    b, n, _ = x.shape
    scale = dim_head ** -0.5
    for l in range(depth):
        pa, pf = f"{prefix}.layers.{l}.0", f"{prefix}.layers.{l}.1"
        h = layer_norm(x, sd, pa + ".norm")
        q, k, v = F.linear(h, sd[pa + ".to_qkv.weight"]).chunk(3, dim=-1)
        q, k, v = (t.reshape(b, n, heads, dim_head).transpose(1, 2) for t in (q, k, v))
        attn = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * scale, dim=-1)
        o = torch.matmul(attn, v).transpose(1, 2).reshape(b, n, heads * dim_head)
        if (pa + ".to_out.0.weight") in sd:
            o = linear(o, sd, pa + ".to_out.0")
        x = o + x
        h = layer_norm(x, sd, pf + ".net.0")
        h = F.gelu(linear(h, sd, pf + ".net.1"))
        x = linear(h, sd, pf + ".net.4") + x
    return layer_norm(x, sd, prefix + ".norm")


def early_cnn(x, sd, prefix, key):
    """EarlyCNN (pretrain_models.py:37-56)."""
    x = F.relu(F.conv2d(x, sd[prefix + ".conv1.weight"], sd[prefix + ".conv1.bias"], stride=2, padding=1))
    x = F.relu(F.conv2d(x, sd[prefix + ".conv2.weight"], sd[prefix + ".conv2.bias"], stride=2, padding=1))
    if key == "image":
        x = F.relu(F.conv2d(x, sd[prefix + ".conv3.weight"], sd[prefix + ".conv3.bias"], stride=2, padding=1))
    else:
        x = F.relu(F.conv2d(x, sd[prefix + ".conv3.weight"], sd[prefix + ".conv3.bias"], stride=1, padding=1))
    x = F.conv2d(x, sd[prefix + ".conv4.weight"], sd[prefix + ".conv4.bias"])
    return x.flatten(2).transpose(1, 2)


# --------------------------------------------------------------------------------------------
# masking (pretrain_models.py:223-248)
# --------------------------------------------------------------------------------------------
def mask_counts(masking_ratio: float, n_img: int, n_tac_total: int, num_tactiles: int):
    """Python-float truncations exactly as the reference performs them."""
    n = n_img + n_tac_total
    num_masked = int(masking_ratio * n)
    image_perc = n_img / n
    n_mask_img = int(num_masked * image_perc)
    n_mask_tac = (num_masked - n_mask_img) // num_tactiles if (num_tactiles > 0 and n_tac_total > 0) else 0
    return n_mask_img, n_mask_tac


def mask_indices(noise: torch.Tensor, n_img: int, n_tac: int, num_tactiles: int,
                 n_mask_img: int, n_mask_tac: int):
    """noise (B, n_img + nt*n_tac): the uniform draws the reference makes with torch.rand, in call
    order image, tactile1, tactile2 (pretrain_models.py:229,237).  Returns int64 (masked, unmasked).
    argsort is ascending; the harness supplies tie-free rows so any correct sort agrees."""
    segs = [(0, n_img, n_mask_img)] if n_img > 0 else []
    for i in range(num_tactiles if n_tac > 0 else 0):
        segs.append((n_img + i * n_tac, n_tac, n_mask_tac))
    masked, unmasked = [], []
    for off, n, nm in segs:
        perm = noise[:, off:off + n].argsort(dim=-1) + off
        masked.append(perm[:, :nm])
        unmasked.append(perm[:, nm:])
    return torch.cat(masked, dim=1), torch.cat(unmasked, dim=1)


def tie_free_noise(batch: int, n: int, generator: torch.Generator, seg_sizes: Optional[List[int]] = None):
    """U[0,1) fp32 noise whose rows have no duplicate within any segment."""
    noise = torch.rand(batch, n, generator=generator)
    seg_sizes = seg_sizes or [n]
    for _ in range(100):
        bad = torch.zeros(batch, dtype=torch.bool)
        off = 0
        for s in seg_sizes:
            srt = noise[:, off:off + s].sort(dim=-1).values
            bad |= (srt[:, 1:] == srt[:, :-1]).any(dim=-1)
            off += s
        if not bad.any():
            return noise
        noise[bad] = torch.rand(int(bad.sum()), n, generator=generator)
    raise RuntimeError("could not draw tie-free noise")


# --------------------------------------------------------------------------------------------
# token embedding (shared by forward and get_embeddings)
# --------------------------------------------------------------------------------------------
def _tokens(sd, cfg: VTMAEConfig, x: dict, use_vision: bool, use_tactile: bool):
    if "image" not in x:
        use_vision = False
    has_tac = cfg.num_tactiles > 0 and use_tactile
    ref = x["image"] if "image" in x else x["tactile1"]
    b = ref.shape[0]
    pi, pj = _pair(cfg.image_patch_size)
    ti, tj = _pair(cfg.tactile_patch_size)
    D = cfg.dim
    img_patches = patchify(x["image"], pi, pj) if use_vision else ref.new_zeros(b, 0, 3)
    tac_patches = (torch.cat([patchify(x[f"tactile{i}"], ti, tj) for i in range(1, cfg.num_tactiles + 1)], dim=1)
                   if has_tac else ref.new_zeros(b, 0, 3))
    n_img, n_tac_total = img_patches.shape[1], tac_patches.shape[1]

    def embed(p, prefix):  # LayerNorm(P) -> Linear(P, D) -> LayerNorm(D)
        h = layer_norm(p, sd, prefix + ".0")
        h = linear(h, sd, prefix + ".1")
        return layer_norm(h, sd, prefix + ".2")

    if cfg.early_conv_masking:
        img_tok = early_cnn(x["image"], sd, "early_conv_vision", "image") if use_vision else ref.new_zeros(b, 0, D)
        tac_tok = (torch.cat([early_cnn(x[f"tactile{i}"], sd, "early_conv_tactile", "tactile")
                              for i in range(1, cfg.num_tactiles + 1)], dim=1)
                   if has_tac else ref.new_zeros(b, 0, D))
    else:
        img_tok = embed(img_patches, "image_patch_to_emb") if use_vision else ref.new_zeros(b, 0, D)
        tac_tok = embed(tac_patches, "tactile_patch_to_emb") if has_tac else ref.new_zeros(b, 0, D)

    if cfg.use_sincosmod_encodings:
        mod = sd["encoder_modality_embedding.weight"]
        if use_vision:
            img_tok = img_tok + mod[0] + sd["image_enc_pos_embedding"]
        if has_tac:
            n1 = n_tac_total // cfg.num_tactiles
            tac_tok = torch.cat([tac_tok[:, i * n1:(i + 1) * n1] + mod[1 + i] for i in range(cfg.num_tactiles)], dim=1)
            tac_tok = tac_tok + sd["tactile_enc_pos_embedding"]
    tokens = torch.cat((img_tok, tac_tok), dim=1)
    if not cfg.use_sincosmod_encodings:
        tokens = tokens + sd["encoder.pos_embedding"][:, 1:(n_img + n_tac_total + 1)]
    return tokens, img_patches, tac_patches, use_vision, has_tac


def vtmae_embeddings(sd, cfg: VTMAEConfig, x: dict, use_vision=True, use_tactile=True):
    """VTMAE.get_embeddings (pretrain_models.py:588-668): encoder over all tokens, no masking."""
    tokens, *_ = _tokens(sd, cfg, x, use_vision, use_tactile)
    return transformer(tokens, sd, "encoder.transformer", cfg.depth, cfg.heads, cfg.dim_head)


def vtmae_forward(sd, cfg: VTMAEConfig, x: dict, noise: torch.Tensor, use_vision=True, use_tactile=True,
                  intermediates: Optional[dict] = None):
    """VTMAE.forward (pretrain_models.py:146-342) with the mask noise supplied externally.
    `noise` has one column per token actually present (image first, then each sensor)."""
    tokens, img_patches, tac_patches, use_vision, has_tac = _tokens(sd, cfg, x, use_vision, use_tactile)
    b, n, _ = tokens.shape
    n_img, n_tac_total = img_patches.shape[1], tac_patches.shape[1]
    nt = cfg.num_tactiles if has_tac else 0
    n_tac = n_tac_total // nt if nt else 0
    nm_img, nm_tac = mask_counts(cfg.masking_ratio, n_img, n_tac_total, cfg.num_tactiles)
    masked, unmasked = mask_indices(noise, n_img, n_tac, nt, nm_img, nm_tac)
    masked_img, masked_tac = masked[:, :nm_img], masked[:, nm_img:]
    br = torch.arange(b)[:, None]

    enc_in = tokens[br, unmasked]
    encoded = transformer(enc_in, sd, "encoder.transformer", cfg.depth, cfg.heads, cfg.dim_head)
    dec_tok = linear(encoded, sd, "enc_to_dec") if "enc_to_dec.weight" in sd else encoded

    Dd = cfg.decoder_dim
    mask_tokens = sd["mask_token"][None, None, :].expand(b, masked.shape[1], Dd)
    if not cfg.use_sincosmod_encodings:
        dec_tok = dec_tok + sd["decoder_pos_emb.weight"][unmasked]
        mask_tokens = mask_tokens + sd["decoder_pos_emb.weight"][masked]
    z = torch.zeros(b, n, Dd)
    z = z.index_put((br, unmasked), dec_tok)
    z = z.index_put((br, masked), mask_tokens)
    if cfg.use_sincosmod_encodings:
        mod = sd["decoder_modality_embedding.weight"]
        parts = []
        if use_vision:
            parts.append(z[:, :n_img] + mod[0] + sd["image_dec_pos_embedding"])
        if has_tac:
            zt = torch.cat([z[:, n_img + i * n_tac:n_img + (i + 1) * n_tac] + mod[1 + i] for i in range(nt)], dim=1)
            parts.append(zt + sd["tactile_dec_pos_embedding"])
        z = torch.cat(parts, dim=1)
    decoded = transformer(z, sd, "decoder", cfg.decoder_depth, cfg.decoder_heads, cfg.decoder_dim_head)

    loss = 0
    if cfg.early_conv_masking:  # loss over ALL patches (pretrain_models.py:311-322)
        if has_tac:
            loss = loss + 10 * F.mse_loss(linear(decoded[:, n_img:], sd, "to_tactiles"), tac_patches)
        if use_vision:
            loss = loss + F.mse_loss(linear(decoded[:, :n_img], sd, "to_pixels"), img_patches)
    else:                       # masked patches only (pretrain_models.py:324-340)
        if has_tac:
            pred = linear(decoded[br, masked_tac], sd, "to_tactiles")
            loss = loss + 10 * F.mse_loss(pred, tac_patches[br, masked_tac - n_img])
        if use_vision:
            pred = linear(decoded[br, masked_img], sd, "to_pixels")
            loss = loss + F.mse_loss(pred, img_patches[br, masked_img])
    if intermediates is not None:
        intermediates.update(masked_indices=masked, unmasked_indices=unmasked, enc_in=enc_in, encoded=encoded,
                             decoder_in=z, decoded=decoded)
    return loss


def mask_counts_reconstruct(mask_ratio: float, n_img: int, n_tac_total: int, num_tactiles: int):
    """reconstruct() masks per modality (pretrain_models.py:425,433) - a different rule from forward()."""
    nm_img = int(mask_ratio * n_img) if n_img else 0
    nm_tac = int(mask_ratio * n_tac_total / num_tactiles) if n_tac_total else 0
    return nm_img, nm_tac


def unpatchify_image(p, gh, gw, ph, pw):
    """Rearrange('b (h w) (p1 p2 c) -> b c (h p1) (w p2)') (pretrain_models.py:463-465)."""
    b = p.shape[0]
    return p.reshape(b, gh, gw, ph, pw, -1).permute(0, 5, 1, 3, 2, 4).reshape(b, -1, gh * ph, gw * pw)


def unpatchify_tactile(p, n, gh, gw, ph, pw):
    """Rearrange('b (n h w) (p1 p2 c) -> b (n c) (h p1) (w p2)') (pretrain_models.py:476-478)."""
    b = p.shape[0]
    return p.reshape(b, n, gh, gw, ph, pw, -1).permute(0, 1, 6, 2, 4, 3, 5).reshape(b, -1, gh * ph, gw * pw)


def vtmae_reconstruct(sd, cfg: VTMAEConfig, x: dict, noise: torch.Tensor, mask_ratio=None, use_vision=True,
                      use_tactile=True):
    """VTMAE.reconstruct (pretrain_models.py:344-586), mask noise supplied externally (image, tactile1,
    tactile2 order as the torch.rand calls at :426,439).  With early_conv_masking the heads run on all
    tokens and the reconstruction is the prediction itself (:560-575)."""
    if mask_ratio is None:
        mask_ratio = cfg.masking_ratio
    tokens, img_patches, tac_patches, use_vision, has_tac = _tokens(sd, cfg, x, use_vision, use_tactile)
    b, n, _ = tokens.shape
    n_img, n_tac_total = img_patches.shape[1], tac_patches.shape[1]
    nt = cfg.num_tactiles if has_tac else 0
    n_tac = n_tac_total // nt if nt else 0
    nm_img, nm_tac = mask_counts_reconstruct(mask_ratio, n_img, n_tac_total, cfg.num_tactiles)
    masked, unmasked = mask_indices(noise, n_img, n_tac, nt, nm_img, nm_tac)
    masked_img, masked_tac = masked[:, :nm_img], masked[:, nm_img:]
    br = torch.arange(b)[:, None]
    encoded = transformer(tokens[br, unmasked], sd, "encoder.transformer", cfg.depth, cfg.heads, cfg.dim_head)
    dec_tok = linear(encoded, sd, "enc_to_dec") if "enc_to_dec.weight" in sd else encoded
    Dd = cfg.decoder_dim
    mask_tokens = sd["mask_token"][None, None, :].expand(b, masked.shape[1], Dd)
    if not cfg.use_sincosmod_encodings:
        dec_tok = dec_tok + sd["decoder_pos_emb.weight"][unmasked]
        mask_tokens = mask_tokens + sd["decoder_pos_emb.weight"][masked]
    z = torch.zeros(b, n, Dd).index_put((br, unmasked), dec_tok).index_put((br, masked), mask_tokens)
    if cfg.use_sincosmod_encodings:
        mod = sd["decoder_modality_embedding.weight"]
        parts = []
        if use_vision:
            parts.append(z[:, :n_img] + mod[0] + sd["image_dec_pos_embedding"])
        if has_tac:
            zt = torch.cat([z[:, n_img + i * n_tac:n_img + (i + 1) * n_tac] + mod[1 + i] for i in range(nt)], dim=1)
            parts.append(zt + sd["tactile_dec_pos_embedding"])
        z = torch.cat(parts, dim=1)
    decoded = transformer(z, sd, "decoder", cfg.decoder_depth, cfg.decoder_heads, cfg.decoder_dim_head)
    out = {}
    if use_vision:
        (H, W), (ph, pw) = _pair(cfg.image_size), _pair(cfg.image_patch_size)
        vis, rec = img_patches.clone(), img_patches.clone()
        vis[br, masked_img] = 0.5
        if cfg.early_conv_masking:
            rec = linear(decoded[:, :n_img], sd, "to_pixels")
            loss_img = F.mse_loss(rec, img_patches)
        else:
            pred = linear(decoded[br, masked_img], sd, "to_pixels")
            rec[br, masked_img] = pred
            loss_img = F.mse_loss(pred, img_patches[br, masked_img])
        out["image_rec"] = unpatchify_image(rec, H // ph, W // pw, ph, pw)
        out["image_masked"] = unpatchify_image(vis, H // ph, W // pw, ph, pw)
        out["recon_loss_image"] = loss_img
    if has_tac:
        (H, W), (ph, pw) = _pair(cfg.tactile_size), _pair(cfg.tactile_patch_size)
        vis, rec = tac_patches.clone(), tac_patches.clone()
        vis[br, masked_tac - n_img] = float("inf")
        if cfg.early_conv_masking:
            rec = linear(decoded[:, n_img:], sd, "to_tactiles")
            loss_tac = F.mse_loss(rec, tac_patches)
        else:
            pred = linear(decoded[br, masked_tac], sd, "to_tactiles")
            rec[br, masked_tac - n_img] = pred
            loss_tac = F.mse_loss(pred, tac_patches[br, masked_tac - n_img])
        out["tactile_rec"] = unpatchify_tactile(rec, nt, H // ph, W // pw, ph, pw)
        out["tactile_masked"] = unpatchify_tactile(vis, nt, H // ph, W // pw, ph, pw)
        out["recon_loss_tactile"] = loss_tac
    return out


def extractor_forward(sd_mae, cfg: VTMAEConfig, sd_vit, observations: dict, vision_only_control=False):
    """MAEExtractor.forward (pretrain_models.py:819-841): 5-D obs -> (B, dim).
    `sd_vit` holds the extra 1-layer `vit_layer.transformer` (dim, 1, 4, 64, 2*dim)."""
    obs = dict(observations)
    if "image" in obs and obs["image"].dim() == 5:      # (B, F, H, W, 3) -> (B, H, W, 3F)
        im = obs["image"].permute(0, 2, 3, 1, 4)
        obs["image"] = im.reshape(im.shape[0], im.shape[1], im.shape[2], -1)
    if "tactile" in obs and obs["tactile"].dim() == 5:  # (B, F, 6, h, w) -> (B, 6F, h, w)
        t = obs["tactile"]
        obs["tactile"] = t.reshape(t.shape[0], -1, t.shape[3], t.shape[4])
    x = vt_load(obs, frame_stack=cfg.frame_stack)
    emb = vtmae_embeddings(sd_mae, cfg, x, use_tactile=not vision_only_control)
    emb = transformer(emb, sd_vit, "transformer", 1, 4, 64)
    return emb.mean(dim=1)


# --------------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------------
def _linear_init(out_f, in_f, g, bias=True, fan_in=None):
    bound = 1.0 / math.sqrt(fan_in or in_f)
    w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
    if not bias:
        return w, None
    return w, (torch.rand(out_f, generator=g) * 2 - 1) * bound


def _transformer_params(sd, prefix, dim, depth, heads, dim_head, mlp_dim, g):
    inner = heads * dim_head
    for l in range(depth):
        pa, pf = f"{prefix}.layers.{l}.0", f"{prefix}.layers.{l}.1"
        sd[pa + ".norm.weight"], sd[pa + ".norm.bias"] = torch.ones(dim), torch.zeros(dim)
        sd[pa + ".to_qkv.weight"], _ = _linear_init(3 * inner, dim, g, bias=False)
        if not (heads == 1 and dim_head == dim):
            sd[pa + ".to_out.0.weight"], sd[pa + ".to_out.0.bias"] = _linear_init(dim, inner, g)
        sd[pf + ".net.0.weight"], sd[pf + ".net.0.bias"] = torch.ones(dim), torch.zeros(dim)
        sd[pf + ".net.1.weight"], sd[pf + ".net.1.bias"] = _linear_init(mlp_dim, dim, g)
        sd[pf + ".net.4.weight"], sd[pf + ".net.4.bias"] = _linear_init(dim, mlp_dim, g)
    sd[prefix + ".norm.weight"], sd[prefix + ".norm.bias"] = torch.ones(dim), torch.zeros(dim)


def init_state_dict(cfg: VTMAEConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Canonical (alias-free) parameter + buffer dict with the reference's names, shapes and
    default-init distributions.  `expand_aliases` adds the duplicate
    `encoder.{image,tactile}_to_patch_embedding.{1,2,3}` names the reference state_dict carries."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    D, Dd = cfg.dim, cfg.decoder_dim
    n_total = cfg.n_img + cfg.num_tactiles * cfg.n_tac
    for name, P in (("image_patch_to_emb", cfg.p_img), ("tactile_patch_to_emb", cfg.p_tac)):
        sd[name + ".0.weight"], sd[name + ".0.bias"] = torch.ones(P), torch.zeros(P)
        sd[name + ".1.weight"], sd[name + ".1.bias"] = _linear_init(D, P, g)
        sd[name + ".2.weight"], sd[name + ".2.bias"] = torch.ones(D), torch.zeros(D)
    sd["encoder.pos_embedding"] = torch.randn(1, n_total + 1, D, generator=g)
    _transformer_params(sd, "encoder.transformer", D, cfg.depth, cfg.heads, cfg.dim_head, cfg.mlp_dim, g)
    if D != Dd:
        sd["enc_to_dec.weight"], sd["enc_to_dec.bias"] = _linear_init(Dd, D, g)
    sd["mask_token"] = torch.randn(Dd, generator=g)
    _transformer_params(sd, "decoder", Dd, cfg.decoder_depth, cfg.decoder_heads, cfg.decoder_dim_head,
                        cfg.decoder_mlp_dim, g)
    sd["decoder_pos_emb.weight"] = torch.randn(n_total, Dd, generator=g)
    sd["to_pixels.weight"], sd["to_pixels.bias"] = _linear_init(cfg.p_img, Dd, g)
    sd["to_tactiles.weight"], sd["to_tactiles.bias"] = _linear_init(cfg.p_tac, Dd, g)
    sd["encoder_modality_embedding.weight"] = torch.randn(1 + cfg.num_tactiles, D, generator=g)
    sd["decoder_modality_embedding.weight"] = torch.randn(1 + cfg.num_tactiles, Dd, generator=g)
    if cfg.early_conv_masking:
        for name, cin, key in (("early_conv_vision", cfg.image_channels, "image"),
                               ("early_conv_tactile", cfg.tactile_channels, "tactile")):
            chans = [cin, D // 8, D // 4, D // 2, D]
            ks = [4, 4, 4 if key == "image" else 3, 1]
            for i in range(4):
                w = torch.empty(chans[i + 1], chans[i], ks[i], ks[i])
                bound = 1.0 / math.sqrt(chans[i] * ks[i] * ks[i])
                w.copy_((torch.rand(w.shape, generator=g) * 2 - 1) * bound)
                sd[f"{name}.conv{i + 1}.weight"] = w
                sd[f"{name}.conv{i + 1}.bias"] = (torch.rand(chans[i + 1], generator=g) * 2 - 1) * bound
    sd.update(position_buffers(cfg))
    return sd


def position_buffers(cfg: VTMAEConfig) -> Dict[str, torch.Tensor]:
    """The four sin-cos buffers (pretrain_models.py:120-140).  The encoder-dim generator is used for
    the decoder tables too (Appendix A.3), hence min(decoder_dim, 2*ch) channels."""
    gh, gw = cfg.image_grid
    th, tw = cfg.tactile_grid
    out = {}
    ch2 = 2 * int(math.ceil(cfg.dim / 4) * 2)
    for tag, C in (("enc", cfg.dim), ("dec", cfg.decoder_dim)):
        c_eff = min(C, ch2)
        full = sincos_2d(gh, gw, cfg.dim) if C == cfg.dim else _sincos_with_generator_dim(gh, gw, cfg.dim, c_eff)
        tac = sincos_2d(th, tw, cfg.dim) if C == cfg.dim else _sincos_with_generator_dim(th, tw, cfg.dim, c_eff)
        out[f"image_{tag}_pos_embedding"] = full[None]
        out[f"tactile_{tag}_pos_embedding"] = tac.repeat(cfg.num_tactiles, 1)[None]
    return out


def _sincos_with_generator_dim(nx, ny, gen_channels, out_channels):
    ch = int(math.ceil(gen_channels / 4) * 2)
    return sincos_2d(nx, ny, 2 * ch)[:, :out_channels]


BUFFER_KEYS = ("image_enc_pos_embedding", "tactile_enc_pos_embedding",
               "image_dec_pos_embedding", "tactile_dec_pos_embedding")


def expand_aliases(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Adds `encoder.*_to_patch_embedding.{1,2,3}.*` aliases of `*_patch_to_emb.{0,1,2}.*`
    (the reference registers the same modules twice: pretrain_models.py:99-105)."""
    out = dict(sd)
    for mod in ("image", "tactile"):
        for i in range(3):
            for leaf in ("weight", "bias"):
                k = f"{mod}_patch_to_emb.{i}.{leaf}"
                if k in sd:
                    out[f"encoder.{mod}_to_patch_embedding.{i + 1}.{leaf}"] = sd[k]
    return out


def canonical(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Reference-style state_dict -> alias-free dict (drops the encoder.*_to_patch_embedding copies)."""
    return {k: v for k, v in sd.items()
            if not (k.startswith("encoder.image_to_patch_embedding") or
                    k.startswith("encoder.tactile_to_patch_embedding"))}


def param_keys(sd) -> List[str]:
    return [k for k in sd if k not in BUFFER_KEYS]


# --------------------------------------------------------------------------------------------
# train step (pretrain_models.py:707-711): zero_grad, fwd, bwd, clip_grad_norm_(0.5), AdamW.step
# --------------------------------------------------------------------------------------------
@dataclass
class AdamWState:
    step: int = 0
    m: Dict[str, torch.Tensor] = field(default_factory=dict)
    v: Dict[str, torch.Tensor] = field(default_factory=dict)


def clip_and_adamw(sd, grads: Dict[str, Optional[torch.Tensor]], st: AdamWState, lr=1e-4, betas=(0.9, 0.999),
                   eps=1e-8, weight_decay=0.01, max_norm=0.5):
    """torch.nn.utils.clip_grad_norm_(params, 0.5) then torch.optim.AdamW(lr).step() (defaults),
    restated; params whose grad is None are skipped entirely (no decay)."""
    live = {k: g for k, g in grads.items() if g is not None}
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in live.values()]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    st.step += 1
    b1, b2 = betas
    bc1, bc2 = 1 - b1 ** st.step, 1 - b2 ** st.step
    with torch.no_grad():
        for k, g in live.items():
            g = g * coef
            p = sd[k]
            if k not in st.m:
                st.m[k], st.v[k] = torch.zeros_like(p), torch.zeros_like(p)
            p.mul_(1 - lr * weight_decay)
            st.m[k].lerp_(g, 1 - b1)
            st.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (st.v[k].sqrt() / math.sqrt(bc2)).add_(eps)
            p.addcdiv_(st.m[k], denom, value=-lr / bc1)
    return total


def train_step(sd, cfg, x, noise, st: AdamWState, lr=1e-4):
    """One reference train step on an alias-free state_dict of leaf tensors. Returns (loss, grad_norm, grads)."""
    keys = param_keys(sd)
    for k in keys:
        sd[k].requires_grad_(True)
        sd[k].grad = None
    loss = vtmae_forward(sd, cfg, x, noise)
    loss.backward()
    grads = {k: sd[k].grad for k in keys}
    norm = clip_and_adamw(sd, grads, st, lr=lr)
    return loss.detach(), norm, grads
